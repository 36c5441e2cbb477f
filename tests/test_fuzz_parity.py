"""Randomised parity: seeded random scenes in the reference's JSON schema (every shape, BSDF, light type, procedural
textures, a random triangle mesh, depth of field) through

  * the host model (loader + BVH builder + flattening)  vs  the reference's own Scene object, byte for byte,
  * the device headers (tests/hostsim on CPU, the CUDA kernels through the C ABI on a GPU)  vs  the reference's render of
    the same passes, bit for bit (sum buffer, secondary sum, ray counters).

The fixed fixtures walk the branches their author thought of; these walk combinations nobody picked by hand
(a rotated box light seen through rough glass, a spot light inside a mesh, near-specular roughness ...).
The live reference (oracle/_ref) is required; without it the tests skip.
"""
import json
import math
import os
import random
import subprocess

import numpy as np
import pytest

from raytracer_b200 import capi, host_api
from tests import util

W, H, PASSES, DIMS, DEPTH = 40, 30, 2, 256, 5
SEEDS = list(range(1, 13))
BSDFS = ["diffuse", "roughDiffuse", "dielectric", "roughDielectric", "metal", "roughMetal", "plastic", "roughPlastic", "null"]
HOSTSIM_DIR = os.path.join(util.ROOT, "tests", "hostsim")
HOSTSIM = os.path.join(HOSTSIM_DIR, "hostsim")

needs_oracle = pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")


def _xform(rng, spread=3.0, y=(-1.0, 2.5), rotate=True):
    d = {"translation": [round(rng.uniform(-spread, spread), 3), round(rng.uniform(*y), 3), round(rng.uniform(-spread, spread), 3)]}
    if rotate and rng.random() < 0.7:
        d["orientation"] = [round(rng.uniform(-180.0, 180.0), 2) for _ in range(3)]
    return d


def _color(rng, lo=0.05, hi=1.0):
    return [round(rng.uniform(lo, hi), 3) for _ in range(3)]


def _shape(rng, area_light=False):
    kind = rng.choice(["sphere", "box", "plane" if area_light else "rect"])
    if kind == "sphere":
        return {"type": "sphere", "radius": round(rng.uniform(0.2, 1.2), 3)}
    if kind == "box":
        return {"type": "box", "size": [round(rng.uniform(0.15, 1.5), 3) for _ in range(3)]}
    return {"type": kind, "size": [round(rng.uniform(0.3, 2.5), 3) for _ in range(2)]}


def write_random_mesh(rng, path):
    """A small height field plus a few free triangles, two materials from a .mtl."""
    n = rng.randint(3, 7)
    verts, faces = [], {"a": [], "b": []}
    for j in range(n + 1):
        for i in range(n + 1):
            verts.append((i * 0.6 - 0.3 * n, 0.35 * math.sin(1.3 * i + rng.random()) * math.cos(0.9 * j), j * 0.6 - 0.3 * n))
    for j in range(n):
        for i in range(n):
            a, b, c, d = j * (n + 1) + i + 1, j * (n + 1) + i + 2, (j + 1) * (n + 1) + i + 2, (j + 1) * (n + 1) + i + 1
            m = "a" if (i + j) % 2 else "b"
            faces[m].append((a, b, c))
            faces[m].append((a, c, d))
    for _ in range(rng.randint(0, 6)):
        base = len(verts)
        p = [rng.uniform(-1.5, 1.5), rng.uniform(0.3, 1.5), rng.uniform(-1.5, 1.5)]
        for _ in range(3):
            verts.append(tuple(p[k] + rng.uniform(-0.6, 0.6) for k in range(3)))
        faces["a"].append((base + 1, base + 2, base + 3))
    name = os.path.basename(path)[:-4]
    with open(path, "w") as f:
        f.write("mtllib %s.mtl\n" % name)
        f.write("".join("v %.5f %.5f %.5f\n" % v for v in verts))
        for m, fs in faces.items():
            f.write("usemtl %s\n" % m)
            f.write("".join("f %d %d %d\n" % t for t in fs))
    with open(path[:-4] + ".mtl", "w") as f:
        f.write("newmtl a\nKd 0.8 0.4 0.3\nKe 0 0 0\n\nnewmtl b\nKd 0.3 0.5 0.8\nKe 0 0 0\n\n")


def write_random_bitmaps(rng, out_dir):
    """24-bit colour, 8-bit palette and 8-bit grey BMPs of odd sizes (the three layouts the BMP reader accepts)."""
    from tests.scenes.make_scenes import write_bmp
    tex = os.path.join(out_dir, "TEXTURES")
    os.makedirs(tex, exist_ok=True)
    a, b, c = (rng.randint(1, 97) for _ in range(3))
    w, h = rng.randint(1, 19), rng.randint(1, 13)
    write_bmp(os.path.join(tex, "color.bmp"), w, h, 24, lambda x, y: ((x * a + y * 7) % 256, (y * b + x * 3) % 256, (x * y * c + 11) % 256))
    write_bmp(os.path.join(tex, "normal.bmp"), rng.randint(2, 9), rng.randint(2, 9), 24,
              lambda x, y: (255, int(128 + 90 * math.cos(y * 0.9 + x * 0.3 + a)), int(128 + 90 * math.sin(x * 0.8 + b))))
    write_bmp(os.path.join(tex, "env.bmp"), rng.randint(4, 24), rng.randint(2, 12), 24,
              lambda x, y: (min(255, 40 + (x * a) % 60 + 10 * y), min(255, 60 + 8 * y), min(255, 30 + (y * b) % 90 + 4 * x)))
    palette = [((i * 53 + a) % 256, (i * 97 + 30) % 256, (255 - i * 16) % 256, (i * 17) % 256) for i in range(16)]
    write_bmp(os.path.join(tex, "pal.bmp"), rng.randint(1, 15), rng.randint(1, 9), 8, lambda x, y: (x * 3 + y * 5 + c) % 16, palette)
    write_bmp(os.path.join(tex, "grey.bmp"), rng.randint(1, 9), rng.randint(1, 9), 8, lambda x, y: (x * 40 + y * 9 + a) % 256)


def write_random_textured_mesh(rng, path):
    """A folded strip with texture coordinates and (sometimes) vertex normals; the .mtl binds textures by file name."""
    n = rng.randint(2, 8)
    with_normals = rng.random() < 0.5
    lines = ["mtllib %s.mtl" % os.path.basename(path)[:-4]]
    for i in range(n + 1):
        x = -1.5 + 3.0 * i / n
        for j, z in enumerate((-0.6, 0.6)):
            lines.append("v %.6f %.6f %.6f" % (x, 0.3 * math.sin(i * 1.3 + rng.random()) + 0.1 * j, z))
    for i in range(n + 1):
        for j in range(2):
            lines.append("vt %.6f %.6f" % (rng.uniform(-1.5, 2.5), rng.uniform(-1.5, 2.5)))
    if with_normals:
        for i in range(2 * (n + 1)):
            nx, ny, nz = rng.uniform(-0.4, 0.4), 1.0, rng.uniform(-0.4, 0.4)
            l = math.sqrt(nx * nx + ny * ny + nz * nz)
            lines.append("vn %.6f %.6f %.6f" % (nx / l, ny / l, nz / l))
    lines.append("usemtl cloth")
    fmt = "%d/%d/%d" if with_normals else "%d/%d"
    for i in range(n):
        quad = (2 * i + 1, 2 * i + 2, 2 * i + 4, 2 * i + 3)
        for tri in ((0, 1, 2), (0, 2, 3)):
            lines.append("f " + " ".join(fmt % ((quad[k],) * (3 if with_normals else 2)) for k in tri))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(path[:-4] + ".mtl", "w") as f:
        f.write("newmtl cloth\nKd 0.9 0.8 0.7\nKe 0 0 0\nmap_Kd ../TEXTURES/%s\n" % rng.choice(["color.bmp", "pal.bmp"]))
        if rng.random() < 0.6:
            f.write("norm ../TEXTURES/normal.bmp\n")
        f.write("\n")


def random_scene(seed, out_dir):
    rng = random.Random(seed)
    write_random_bitmaps(rng, out_dir)
    textures = [
        {"name": "checker", "type": "checkerboard", "colorA": _color(rng), "colorB": _color(rng)},
        {"name": "clouds", "type": "noise", "colorA": _color(rng, 0.0, 0.3), "colorB": _color(rng, 0.6, 1.0), "octaves": rng.randint(1, 5)},
        {"name": "blend", "type": "mix", "textureA": "checker", "textureB": "clouds", "weight": "clouds"},
        {"name": "photo", "type": "bitmap", "path": "TEXTURES/color.bmp"},
        {"name": "blend2", "type": "mix", "textureA": "blend", "textureB": "photo", "weight": "TEXTURES/grey.bmp"},
    ]
    materials = [{"name": "ground", "bsdf": "diffuse", "baseColor": _color(rng, 0.4, 0.9)}]
    for i in range(rng.randint(3, 7)):
        bsdf = rng.choice(BSDFS)
        m = {"name": "m%d" % i, "bsdf": bsdf, "baseColor": _color(rng)}
        if bsdf.startswith("rough"):
            m["roughness"] = rng.choice([0.001, 0.004, 0.006, round(rng.uniform(0.02, 1.0), 3)])      # < 0.005 falls back to specular
        if "ielectric" in bsdf or "lastic" in bsdf:
            m["IoR"] = round(rng.uniform(1.05, 2.3), 3)
        if "etal" in bsdf:
            m["IoR"] = round(rng.uniform(0.1, 2.5), 3)
            m["K"] = round(rng.uniform(0.5, 5.0), 3)
        if rng.random() < 0.2:
            m["emissionColor"] = _color(rng, 0.2, 3.0)
        if rng.random() < 0.3:
            m["baseColorTexture"] = rng.choice(["checker", "clouds", "blend", "photo", "blend2", "TEXTURES/pal.bmp"])
        if rng.random() < 0.15 and "roughness" in m:
            m["roughnessTexture"] = "clouds"
        if rng.random() < 0.15:
            m["metalness"] = round(rng.random(), 3)
            if rng.random() < 0.5:
                m["metalnessTexture"] = "TEXTURES/grey.bmp"
        if rng.random() < 0.25:
            m["normalMap"] = "TEXTURES/normal.bmp"
            m["normalMapStrength"] = round(rng.uniform(0.1, 1.0), 3)
        if "emissionColor" in m and rng.random() < 0.4:
            m["emissionTexture"] = rng.choice(["checker", "photo"])
        materials.append(m)
    names = [m["name"] for m in materials[1:]]
    objects = [{"type": "box", "size": [6.0, 0.2, 6.0], "transform": {"translation": [0.0, -1.6, 0.0]}, "material": "ground"}]
    for _ in range(rng.randint(2, 8)):
        o = _shape(rng)
        o["transform"] = _xform(rng)
        o["material"] = rng.choice(names)
        if o["type"] == "rect" and rng.random() < 0.5:
            o["textureScale"] = [round(rng.uniform(0.2, 3.0), 3), round(rng.uniform(0.2, 3.0), 3)]
        objects.append(o)
    if rng.random() < 0.6:
        os.makedirs(os.path.join(out_dir, "MODELS"), exist_ok=True)
        write_random_mesh(rng, os.path.join(out_dir, "MODELS", "fuzz%d.obj" % seed))
        objects.append({"type": "mesh", "path": "MODELS/fuzz%d.obj" % seed, "transform": _xform(rng, 1.5, (-1.0, 0.5)), "material": rng.choice(names)})
    if rng.random() < 0.4:
        os.makedirs(os.path.join(out_dir, "MODELS"), exist_ok=True)
        write_random_textured_mesh(rng, os.path.join(out_dir, "MODELS", "cloth%d.obj" % seed))
        objects.append({"type": "mesh", "path": "MODELS/cloth%d.obj" % seed, "transform": _xform(rng, 1.5, (-0.5, 1.5)), "material": rng.choice(names)})
    lights = []
    for _ in range(rng.randint(1, 4)):
        kind = rng.choice(["area", "area", "point", "spot", "directional", "background"])
        if kind == "background" and any(l["type"] == "background" for l in lights):
            kind = "point"
        if kind == "area":
            lights.append({"type": "area", "color": _color(rng, 2.0, 12.0), "transform": _xform(rng, 3.0, (1.0, 4.0)), "shape": _shape(rng, True)})
        elif kind == "point":
            lights.append({"type": "point", "color": _color(rng, 4.0, 20.0), "transform": _xform(rng, 3.0, (0.5, 4.0), rotate=False)})
        elif kind == "spot":
            lights.append({"type": "spot", "color": _color(rng, 10.0, 60.0), "angle": round(rng.uniform(5.0, 60.0), 2), "transform": _xform(rng, 3.0, (1.0, 4.0))})
        elif kind == "directional":
            lights.append({"type": "directional", "color": _color(rng, 1.0, 20.0), "angle": round(rng.uniform(0.0, 10.0), 2),
                           "transform": {"orientation": [round(rng.uniform(20.0, 160.0), 2), round(rng.uniform(-180.0, 180.0), 2), 0.0]}})
        else:
            lights.append({"type": "background", "color": _color(rng, 0.05, 1.5)})
            if rng.random() < 0.5:
                lights[-1]["texture"] = "TEXTURES/env.bmp"
    camera = {"transform": {"translation": [round(rng.uniform(-1.0, 1.0), 3), round(rng.uniform(0.0, 2.0), 3), round(rng.uniform(-8.0, -5.0), 3)],
                            "orientation": [round(rng.uniform(0.0, 20.0), 2), round(rng.uniform(-10.0, 10.0), 2), 0.0]},
              "fieldOfView": round(rng.uniform(35.0, 80.0), 2)}
    if rng.random() < 0.3:
        camera.update({"enableDOF": True, "aperture": round(rng.uniform(0.02, 0.3), 3), "focalPlaneDistance": round(rng.uniform(3.0, 9.0), 3)})
    scene = {"textures": textures, "materials": materials, "objects": objects, "lights": lights, "camera": camera}
    path = os.path.join(out_dir, "fuzz%d.json" % seed)
    with open(path, "w") as f:
        json.dump(scene, f, indent=1)
    return path


def reference_render(tmp_path, sj, data, renderer):
    sump, sum2p, passp = (str(tmp_path / ("ref_%s.%s" % (renderer.replace(" ", ""), e))) for e in ("sum", "sum2", "passes"))
    meta = util.run_rtref(["render", "--scene", sj, "--data", data, "--width", W, "--height", H, "--passes", PASSES, "--dims", DIMS, "--depth", DEPTH,
                           "--strategy", "all", "--renderer", renderer, "--dump-sum", sump, "--dump-sum2", sum2p, "--dump-passes", passp])
    return meta, sump, sum2p, passp


def assert_same_flat(mine, ref):
    from tests.test_host_scene import nodes_view, normalise_nodes
    assert (mine.num_objects, mine.num_shapes, mine.num_materials, mine.num_lights, mine.num_global_lights, len(mine.meshes)) == \
           (ref.num_objects, ref.num_shapes, ref.num_materials, ref.num_lights, ref.num_global_lights, len(ref.meshes))
    assert (normalise_nodes(nodes_view(mine.tlas)) == normalise_nodes(nodes_view(ref.tlas))).all(), "top-level BVH differs"
    for what in ("objects", "shapes", "materials", "lights", "global_lights"):
        assert getattr(mine, what).tobytes() == getattr(ref, what).tobytes(), "%s differ" % what
    for a, b in zip(mine.meshes, ref.meshes):
        for k in ("numNodes", "numTriangles", "numVertices", "numMaterials"):
            assert a[k] == b[k], k
        assert (normalise_nodes(nodes_view(a["nodes"])) == normalise_nodes(nodes_view(b["nodes"]))).all(), "mesh BVH differs"
        for k in ("triangles", "indices", "shading", "materialIds"):
            assert a[k].tobytes() == b[k].tobytes(), "mesh %s differ" % k
    assert bytes(mine.camera) == bytes(ref.camera), "camera differs"
    assert len(mine.textures) == len(ref.textures)
    for i, (a, b) in enumerate(zip(mine.textures, ref.textures)):
        for k in ("type", "width", "height", "numOctaves", "texA", "texB", "texWeight"):
            assert a[k] == b[k], "texture %d: %s" % (i, k)
        assert a["colorA"].tobytes() == b["colorA"].tobytes() and a["colorB"].tobytes() == b["colorB"].tobytes(), "texture %d colours" % i
        assert (a["texels"] is None) == (b["texels"] is None)
        if a["texels"] is not None:
            assert a["texels"].tobytes() == b["texels"].tobytes(), "texture %d: decoded texels differ" % i
    for what in ("material_textures", "light_textures"):
        a, b = getattr(mine, what), getattr(ref, what)
        assert (a is None) == (b is None), what
        if a is not None:
            assert a.tobytes() == b.tobytes(), "%s differ" % what


@pytest.fixture(scope="module")
def hostsim(tmp_path_factory):
    subprocess.run(["bash", os.path.join(HOSTSIM_DIR, "build.sh")], check=True)
    bn = tmp_path_factory.mktemp("bn") / "bn.dat"
    capi.load_blue_noise().tofile(str(bn))     # interleaved RGBA16 layout the sampler reads
    return str(bn)


def prepare(tmp_path, seed):
    """Random scene -> (json path, data dir, the HOST model's flat file); the host's flattening is checked against the
    reference's on the way."""
    data = str(tmp_path) + "/"
    sj = random_scene(seed, str(tmp_path))
    refflat, mineflat = str(tmp_path / "ref.rtb"), str(tmp_path / "mine.rtb")
    util.run_rtref(["flat", "--scene", sj, "--data", data, "--width", W, "--height", H, "--out", refflat])
    scene = host_api.Scene(sj, data)
    try:
        scene.save_flat(mineflat, W, H)
    finally:
        scene.close()
    assert_same_flat(capi.SceneFile(mineflat), capi.SceneFile(refflat))
    return sj, data, mineflat


@needs_oracle
@pytest.mark.parametrize("seed", SEEDS)
def test_random_scene_on_the_device_headers(hostsim, tmp_path, seed):
    sj, data, flat = prepare(tmp_path, seed)
    for renderer, rid in (("Path Tracer MIS", 1), ("Path Tracer", 0)):
        meta, sump, _, passp = reference_render(tmp_path, sj, data, renderer)
        out = str(tmp_path / "mine.sum")
        r = subprocess.run([HOSTSIM, "render", flat, passp, str(W), str(H), out, str(rid), str(DEPTH), "0", "1", hostsim], stdout=subprocess.PIPE, text=True, check=True)
        util.assert_bit_identical(util.load_sum(out, W, H), util.load_sum(sump, W, H), "seed %d %s" % (seed, renderer))
        mine = json.loads(r.stdout.strip().splitlines()[-1])
        assert (mine["numRays"], mine["numShadowRays"], mine["numShadowRaysHit"]) == (meta["numRays"], meta["numShadowRays"], meta["numShadowRaysHit"])
    # ILight::Emit of every light (csrc/rtb_bidir.h)
    from tests.test_hostsim_parity import emission_samples
    sp, refp, minep = str(tmp_path / "samples.f32"), str(tmp_path / "emit_ref.f32"), str(tmp_path / "emit_mine.f32")
    emission_samples(64, seed).tofile(sp)
    util.run_rtref(["emit", "--scene", sj, "--data", data, "--samples", sp, "--out", refp])
    subprocess.run([HOSTSIM, "emit", flat, sp, minep], check=True)
    assert open(refp, "rb").read() == open(minep, "rb").read(), "seed %d: ILight::Emit differs" % seed
    # ILight::Illuminate as the bidirectional renderers call it
    from tests.test_hostsim_parity import shading_points
    pp, refp, minep = str(tmp_path / "points.f32"), str(tmp_path / "ill_ref.f32"), str(tmp_path / "ill_mine.f32")
    shading_points(100, seed).tofile(pp)
    util.run_rtref(["illuminate", "--scene", sj, "--data", data, "--points", pp, "--out", refp])
    subprocess.run([HOSTSIM, "illuminate", flat, pp, minep], check=True)
    assert open(refp, "rb").read() == open(minep, "rb").read(), "seed %d: ILight::Illuminate (area sampling) differs" % seed
    # BSDF::Pdf, forward and reverse, of every material
    from tests.test_hostsim_parity import bsdf_direction_pairs
    dp, refp, minep = str(tmp_path / "dirs.f32"), str(tmp_path / "pdf_ref.f32"), str(tmp_path / "pdf_mine.f32")
    bsdf_direction_pairs(300, seed).tofile(dp)
    util.run_rtref(["bsdfpdf", "--scene", sj, "--data", data, "--dirs", dp, "--out", refp])
    subprocess.run([HOSTSIM, "bsdfpdf", flat, dp, minep], check=True)
    assert open(refp, "rb").read() == open(minep, "rb").read(), "seed %d: BSDF::Pdf differs" % seed
    # DebugRenderer: one mode per scene (normals / tangents / texture coordinates / material parameters ...)
    mode = 1 + seed % 12
    sump, passp, out = str(tmp_path / "dbg.sum"), str(tmp_path / "dbg.passes"), str(tmp_path / "dbg_mine.sum")
    util.run_rtref(["render", "--scene", sj, "--data", data, "--width", W, "--height", H, "--passes", 1, "--dims", DIMS, "--renderer", "Debug",
                    "--debug-mode", mode, "--dump-sum", sump, "--dump-passes", passp])
    subprocess.run([HOSTSIM, "render", flat, passp, str(W), str(H), out, "2", "8", str(mode), "1", hostsim], stdout=subprocess.PIPE, check=True)
    util.assert_bit_identical(util.load_sum(out, W, H), util.load_sum(sump, W, H), "seed %d debug mode %d" % (seed, mode))


@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("seed", SEEDS)
def test_random_scene_on_the_gpu(tmp_path, seed):
    sj, data, flat = prepare(tmp_path, seed)
    sf = capi.SceneFile(flat)
    cam = capi.camera_for(sf, W, H)
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf)
        ctx.resize(W, H)
        for renderer, rid in (("Path Tracer MIS", capi.RTB_RENDERER_PATH_TRACER_MIS), ("Path Tracer", capi.RTB_RENDERER_PATH_TRACER)):
            meta, sump, sum2p, passp = reference_render(tmp_path, sj, data, renderer)
            ps = capi.PassScript(passp)
            ctx.reset()
            ctx.render_passes(cam, ps.seeds[:PASSES], ps.offsets[:PASSES], 0, renderer=rid, max_depth=DEPTH)
            util.assert_bit_identical(ctx.read_sum(), util.load_sum(sump, W, H), "seed %d %s sum" % (seed, renderer))
            util.assert_bit_identical(ctx.read_sum(1), util.load_sum(sum2p, W, H), "seed %d %s secondary sum" % (seed, renderer))
            c = ctx.counters()
            assert (c["numRays"], c["numShadowRays"], c["numShadowRaysHit"]) == (meta["numRays"], meta["numShadowRays"], meta["numShadowRaysHit"])
