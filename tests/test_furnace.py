"""The reference's own rendering tests (Tests/RaytracingTests.cpp:263-523: EmptyScene, BackgroundLightOnly and the
furnace tests), restated through this repo's host model with the reference's scene, pass counts and tolerances.
They are statistical: a unit sphere inside a constant background light must come out as light * albedo."""
import json

import numpy as np
import pytest

from raytracer_b200 import host_api

pytestmark = pytest.mark.gpu

SIZE = 32                                   # RenderingTest::ViewportSize
LIGHT = [1.0, 2.0, 3.0]
RENDERERS = ["Path Tracer", "Path Tracer MIS", "VCM"]     # RaytracingTests.cpp:17-22: the reference runs every test on all three


def furnace_scene(tmp_path, material=None, fov=10.0, sphere=True):
    scene = {"lights": [{"type": "background", "color": LIGHT}],
             "camera": {"transform": {"translation": [0.0, 0.0, -3.0]}, "fieldOfView": fov}}
    if material is not None:
        material = dict(material, name="m")
        scene["materials"] = [material]
    if sphere:
        scene["objects"] = [{"type": "sphere", "radius": 1.0, "material": "m", "transform": {"translation": [0.0, 0.0, 0.0]}}]
    p = tmp_path / "furnace.json"
    p.write_text(json.dumps(scene))
    return str(p)


def render_mean(path, renderer, passes):
    scene = host_api.Scene(path, "")
    vp = host_api.Viewport(scene, device=0)
    try:
        vp.set_renderer(renderer)
        # the reference's default RenderingParams (Context.h:54-90): depth 20, Russian roulette from depth 1, `Single`
        # light sampling; 256 sampler dimensions instead of 64 because this library refuses a pass that could run out
        # of seeded dimensions, where the reference silently falls back to an RNG (GenericSampler.cpp:106-109)
        vp.set_params(dimensions=256, max_ray_depth=20, min_rr_depth=1, strategy="single")
        vp.resize(SIZE, SIZE)
        vp.reset()
        vp.render(passes)
        return vp.read_sum() / np.float32(passes)
    finally:
        vp.close()
        scene.close()


def validate(img, expected, max_error):
    """ValidateBitmap, RaytracingTests.cpp:241-259: EVERY pixel within max_error of the expected colour."""
    err = np.abs(img - np.asarray(expected, dtype=np.float32)[None, None, :])
    assert err.max() <= max_error, "max error %.4f > %.4f at %s" % (err.max(), max_error, np.unravel_index(err.argmax(), err.shape))


@pytest.mark.parametrize("renderer", RENDERERS)
def test_background_light_only(tmp_path, renderer):
    # RaytracingTests.cpp:287-315 (camera FoV 90 degrees, one pass, 0.01)
    validate(render_mean(furnace_scene(tmp_path, sphere=False, fov=90.0), renderer, 1), LIGHT, 0.01)


@pytest.mark.parametrize("renderer", RENDERERS)
def test_furnace_diffuse(tmp_path, renderer):
    # RaytracingTests.cpp:317-365
    albedo = [0.4, 0.6, 0.8]
    img = render_mean(furnace_scene(tmp_path, {"bsdf": "diffuse", "baseColor": albedo}), renderer, 100)
    validate(img, np.array(LIGHT) * np.array(albedo), 0.05)


@pytest.mark.parametrize("renderer", RENDERERS)
def test_furnace_emissive(tmp_path, renderer):
    # RaytracingTests.cpp:367-416: a null BSDF emitter shows exactly its emission colour
    img = render_mean(furnace_scene(tmp_path, {"bsdf": "null", "baseColor": [0.0, 0.0, 0.0], "emissionColor": [3.0, 2.0, 1.0]}), renderer, 1)
    validate(img, [3.0, 2.0, 1.0], 0.0)


@pytest.mark.parametrize("renderer", RENDERERS)
def test_furnace_metal(tmp_path, renderer):
    # RaytracingTests.cpp:418-467
    albedo = [0.4, 0.6, 0.8]
    img = render_mean(furnace_scene(tmp_path, {"bsdf": "metal", "baseColor": albedo, "IoR": 0.0, "K": 100.0}), renderer, 20)
    validate(img, np.array(LIGHT) * np.array(albedo), 0.05)


@pytest.mark.parametrize("renderer", RENDERERS)
def test_furnace_dielectric(tmp_path, renderer):
    # RaytracingTests.cpp:469-516
    img = render_mean(furnace_scene(tmp_path, {"bsdf": "dielectric", "baseColor": [1.0, 1.0, 1.0]}), renderer, 1000)
    validate(img, LIGHT, 0.075)
