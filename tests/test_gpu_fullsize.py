"""BASELINE.json's full-size configuration (sponza-class scene, 1920x1080, PathTracerMIS depth 8, All lights) on the GPU,
checked through size-independent properties (determinism, batched == pass-by-pass, sharded == unsharded) AND bit for
bit against the live reference at the sizes BASELINE.json states: config 1 (sphere + area light, 256x256 x 16 spp),
config 2 (Cornell box, 512x512 x 256 spp, depth 8) and config 3 (the bench scene at 1920x1080; two passes, the CPU
reference needs well under a second per 1080p pass per 16 cores)."""
import os

import numpy as np
import pytest

import bench
from raytracer_b200 import capi, host_api
from tests import util

pytestmark = pytest.mark.gpu
W, H, PASSES = bench.WIDTH, bench.HEIGHT, 4


@pytest.fixture(scope="module")
def scene():
    s = host_api.Scene(bench.ensure_scene(), bench.SCENE_DIR + "/")
    yield s
    s.close()


def render(scene, passes_per_call, num_shards=1, shard_index=0, w=W, h=H):
    vp = host_api.Viewport(scene, device=0, num_shards=num_shards, shard_index=shard_index)
    try:
        vp.set_renderer("Path Tracer MIS")
        vp.set_params(dimensions=bench.DIMS, max_ray_depth=bench.MAX_DEPTH, strategy="all", seed=12345)
        vp.resize(w, h)
        done = 0
        while done < PASSES:
            n = min(passes_per_call, PASSES - done)
            vp.render(n)
            done += n
        return vp.read_sum(), vp.read_sum(1), vp.counters()
    finally:
        vp.close()


@pytest.fixture(scope="module")
def batched(scene):
    return render(scene, PASSES)


def test_full_size_render_is_deterministic(scene, batched):
    again = render(scene, PASSES)
    util.assert_bit_identical(again[0], batched[0], "sum, second run")
    assert again[2]["numRays"] == batched[2]["numRays"] and again[2]["numShadowRays"] == batched[2]["numShadowRays"]


def test_full_size_batched_equals_pass_by_pass(scene, batched):
    single = render(scene, 1)
    util.assert_bit_identical(single[0], batched[0], "sum")
    util.assert_bit_identical(single[1], batched[1], "secondary sum")
    for k in ("numRays", "numShadowRays", "numShadowRaysHit", "numPrimaryRays"):
        assert single[2][k] == batched[2][k], k
    assert batched[2]["numPrimaryRays"] == W * H * PASSES
    assert batched[2]["numKernelLaunches"] < single[2]["numKernelLaunches"]


def test_full_size_image_is_sane(batched):
    img = batched[0] / np.float32(PASSES)
    assert np.isfinite(img).all() and (img >= 0).all()
    assert 0.01 < img.mean() < 100.0
    # secondary sum holds the even passes only (Film.cpp:31-39)
    assert (batched[1] <= batched[0] + 1e-6).all()


def test_full_size_shards_reassemble(scene, batched):
    acc = np.zeros_like(batched[0])
    rays = 0
    for s in range(2):
        part, _, c = render(scene, PASSES, num_shards=2, shard_index=s)
        assert ((acc != 0) & (part != 0)).sum() == 0, "shards overlap"
        acc += part
        rays += c["numRays"] + c["numShadowRays"]
    util.assert_bit_identical(acc, batched[0], "2-shard film")
    assert rays == batched[2]["numRays"] + batched[2]["numShadowRays"]


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
def test_full_size_scene_against_the_live_reference(scene, tmp_path):
    """The bench scene itself (256 768 triangles, loaded and BVH-built by this repo's host code) through the real
    reference on this box's CPU, at a resolution the CPU finishes in seconds: bit-identical film and ray counters."""
    w, h, passes = 320, 180, 2
    sj, data = bench.ensure_scene(), bench.SCENE_DIR + "/"
    sump, passp = str(tmp_path / "o.sum"), str(tmp_path / "o.passes")
    meta = util.run_rtref(["render", "--scene", sj, "--data", data, "--width", w, "--height", h, "--passes", passes, "--dims", bench.DIMS,
                           "--depth", bench.MAX_DEPTH, "--strategy", "all", "--dump-sum", sump, "--dump-passes", passp], seed="4242")
    flat = str(tmp_path / "host.rtb")
    scene.save_flat(flat, w, h)                       # the HOST model's flattening, not the oracle's
    sf = capi.SceneFile(flat)
    ps = capi.PassScript(passp)
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf)
        ctx.resize(w, h)
        ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[:passes], ps.offsets[:passes], 0)
        util.assert_bit_identical(ctx.read_sum(), util.load_sum(sump, w, h), "bench scene vs live reference")
        c = ctx.counters()
    assert (c["numRays"], c["numShadowRays"], c["numShadowRaysHit"]) == (meta["numRays"], meta["numShadowRays"], meta["numShadowRaysHit"])


def live_reference_case(tmp_path, scene_json, data_dir, w, h, passes, dims, host_scene=None, seed="12345", timeout=3000):
    """Render `passes` passes with the real reference (exact-math build) on this box's CPU and with the CUDA path from the
    reference's own pass script (Halton seed vectors + AA offsets); return both films and counter sets."""
    sump, sum2p, passp = str(tmp_path / "o.sum"), str(tmp_path / "o.sum2"), str(tmp_path / "o.passes")
    meta = util.run_rtref(["render", "--scene", scene_json, "--data", data_dir, "--width", w, "--height", h, "--passes", passes, "--dims", dims,
                           "--depth", bench.MAX_DEPTH, "--strategy", "all", "--dump-sum", sump, "--dump-sum2", sum2p, "--dump-passes", passp],
                          seed=seed, timeout=timeout)
    flat = str(tmp_path / "host.rtb")
    own = host_scene is None
    if own:
        host_scene = host_api.Scene(scene_json, data_dir)
    try:
        host_scene.save_flat(flat, w, h)                  # the HOST model's flattening, not the oracle's
    finally:
        if own:
            host_scene.close()
    sf = capi.SceneFile(flat)
    ps = capi.PassScript(passp)
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf)
        ctx.resize(w, h)
        ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[:passes], ps.offsets[:passes], 0)
        got, got2, c = ctx.read_sum(), ctx.read_sum(1), ctx.counters()
    util.assert_bit_identical(got, util.load_sum(sump, w, h), "sum vs live reference")
    util.assert_bit_identical(got2, util.load_sum(sum2p, w, h), "secondary sum vs live reference")
    for k in ("numRays", "numShadowRays", "numShadowRaysHit", "numPrimaryRays"):
        assert c[k] == meta[k], "%s: %d vs reference %d" % (k, c[k], meta[k])
    assert c["numPrimaryRays"] == w * h * passes
    return meta


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
def test_config1_sphere_area_light_256x256_16spp_against_the_live_reference(tmp_path):
    """BASELINE config 1 at its stated size: single sphere + area light, 256x256, 16 spp, PathTracerMIS, fixed seed."""
    scn = os.path.join(util.GOLDEN, "scenes")
    live_reference_case(tmp_path, os.path.join(scn, "c1_sphere_area.json"), scn + "/", 256, 256, 16, 128)


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
def test_config2_cornell_box_512x512_256spp_against_the_live_reference(tmp_path):
    """BASELINE config 2 at its stated size: Cornell box, 512x512, 256 spp, max depth 8 (67 M samples; the film after
    256 passes accumulated in pass order must equal the reference's bit for bit)."""
    scn = os.path.join(util.GOLDEN, "scenes")
    live_reference_case(tmp_path, os.path.join(scn, "cornell.json"), scn + "/", 512, 512, 256, 128)


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("shadow_kernel", ["simple", "wide", "split"])
def test_config3_bench_scene_1920x1080_against_the_live_reference(scene, tmp_path, shadow_kernel, monkeypatch):
    """BASELINE config 3 at its stated resolution: the 256 768-triangle bench scene, 1920x1080, depth 8, both lights
    sampled per vertex; two passes (4.1 M samples, ~33 M rays) against the reference on this box's CPU - with the default
    any-hit kernel (binary tree) and with k_shadow_wide on the 8-wide quantised tree (18 M shadow rays, same booleans)."""
    monkeypatch.setenv("RTB_SHADOW_KERNEL", shadow_kernel)
    live_reference_case(tmp_path, bench.ensure_scene(), bench.SCENE_DIR + "/", bench.WIDTH, bench.HEIGHT, 2, bench.DIMS, host_scene=scene, seed="4242")
