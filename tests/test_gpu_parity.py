"""GPU parity tests proper: the CUDA path, called through the C ABI (include/rtb200.h), against the golden fixtures
generated from the real reference (tests/golden/make_golden.py) and, where the prebuilt oracle binary travelled to
this box (oracle/_ref), against the reference itself run live on the same inputs.

Bar: BIT-EXACT float32 sum buffers, hit records, sampler integers and ray counters (oracle = exact-math build of the
reference: IEEE div/sqrt, no FMA contraction; see DESIGN.md "Oracle").
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from raytracer_b200 import capi
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(device=0)
    yield c
    c.close()


def render_golden_case(ctx, scene, renderer, which=0):
    meta = json.load(open(util.golden("%s_%s.json" % (scene, renderer))))
    w, h = meta["width"], meta["height"]
    sf = capi.SceneFile(util.golden(scene + ".rtb"))
    ps = capi.PassScript(util.golden(scene + ".passes"))
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    rid, dbg = util.RENDERER_IDS[renderer]
    for p in range(meta["passes"]):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p, renderer=rid, debug_mode=dbg, max_depth=8))
    return meta, ctx.read_sum(which), ctx.counters()


@pytest.mark.parametrize("scene,renderer", util.GOLDEN_RENDERS)
def test_sum_buffer_bit_exact(ctx, scene, renderer):
    meta, img, counters = render_golden_case(ctx, scene, renderer)
    want = util.load_sum(util.golden("%s_%s.sum" % (scene, renderer)), meta["width"], meta["height"])
    util.assert_bit_identical(img, want, "%s/%s sum buffer" % (scene, renderer))
    assert counters["numPrimaryRays"] == meta["numPrimaryRays"]
    if renderer in ("mis", "pt"):
        assert counters["numRays"] == meta["numRays"]
        assert counters["numShadowRays"] == meta["numShadowRays"]
        assert counters["numShadowRaysHit"] == meta["numShadowRaysHit"]


@pytest.mark.parametrize("scene,renderer", util.REFERENCE_SCENE_RENDERS)
def test_reference_test_scenes_bit_exact(ctx, scene, renderer):
    """The reference's OWN Data/TestScenes (every one its loader accepts), flattened from the reference's Scene object and
    rendered by the reference (tests/golden/refscenes, make_golden.py): same sum buffer and ray counters, bit for bit."""
    e = util.reference_scene_manifest()[scene]
    base = os.path.join(util.GOLDEN, "refscenes", scene)
    sf = capi.SceneFile(base + ".rtb")
    ps = capi.PassScript(base + ".passes")
    ctx.upload_scene(sf)
    ctx.resize(e["width"], e["height"])
    cam = capi.camera_for(sf, e["width"], e["height"])
    rid, dbg = util.RENDERER_IDS[renderer]
    for p in range(e["passes"]):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p, renderer=rid, debug_mode=dbg, max_depth=8))
    util.assert_bit_identical(ctx.read_sum(), util.load_sum("%s_%s.sum" % (base, renderer), e["width"], e["height"]), "%s/%s" % (scene, renderer))
    counters = ctx.counters()
    for k, v in e[renderer].items():
        assert counters[k] == v, k


@pytest.mark.parametrize("scene", ["c1_sphere_area", "cornell", "materials", "mesh_small", "textures"])
def test_secondary_sum_bit_exact(ctx, scene):
    meta, img, _ = render_golden_case(ctx, scene, "mis", which=1)
    want = util.load_sum(util.golden("%s_mis.sum2" % scene), meta["width"], meta["height"])
    util.assert_bit_identical(img, want, "%s secondary sum" % scene)


@pytest.mark.parametrize("scene", ["cornell", "materials", "mesh_small"])
def test_trace_batch_closest(ctx, scene):
    ctx.upload_scene(capi.SceneFile(util.golden(scene + ".rtb")))
    rays = np.fromfile(util.golden(scene + ".rays"), dtype=capi.RAY_DTYPE)
    want = np.fromfile(util.golden(scene + ".hits"), dtype=capi.HIT_DTYPE)
    got = ctx.trace_batch(rays, capi.RTB_TRACE_CLOSEST)
    assert (got["objectId"] == want["objectId"]).all()
    hit = want["objectId"] != capi.RTB_INVALID_OBJECT
    assert hit.sum() > 100, "fixture should contain hits"
    assert (got["subObjectId"][hit] == want["subObjectId"][hit]).all()
    util.assert_bit_identical(got["distance"], want["distance"], "%s hit distance" % scene)
    mesh = hit & (want["subObjectId"] < 0xFFFFFFF0) & np.isin(want["objectId"], mesh_object_ids(scene))
    util.assert_bit_identical(got["u"][mesh], want["u"][mesh], "%s hit u" % scene)
    util.assert_bit_identical(got["v"][mesh], want["v"][mesh], "%s hit v" % scene)


def mesh_object_ids(scene):
    sf = capi.SceneFile(util.golden(scene + ".rtb"))
    objs = np.frombuffer(sf.objects.tobytes(), dtype=np.uint32).reshape(sf.num_objects, 36)
    shapes = np.frombuffer(sf.shapes.tobytes(), dtype=np.uint32).reshape(sf.num_shapes, 16)
    ids = []
    for i in range(sf.num_objects):
        kind, shape_index = objs[i, 0], objs[i, 1]
        if kind == 0 and shapes[shape_index, 0] == 3:
            ids.append(i)
    return np.array(ids, dtype=np.uint32)


@pytest.mark.parametrize("scene", ["cornell", "materials", "mesh_small"])
def test_trace_batch_shadow(ctx, scene):
    ctx.upload_scene(capi.SceneFile(util.golden(scene + ".rtb")))
    rays = np.fromfile(util.golden(scene + ".rays"), dtype=capi.RAY_DTYPE)
    tmax = np.fromfile(util.golden(scene + ".tmax"), dtype=np.float32)
    want = np.fromfile(util.golden(scene + ".shadow_hits"), dtype=capi.HIT_DTYPE)
    got = ctx.trace_batch(rays, capi.RTB_TRACE_SHADOW, tmax=tmax)
    assert 0 < want["occluded"].sum() < len(want)
    assert (got["occluded"] == want["occluded"]).all()


def test_texture_evaluation_bit_exact(ctx):
    """rtb_texture_batch: ITexture::Evaluate of every texture of the textured scene (bitmaps in three BMP layouts,
    checkerboard, noise, nested mix) at 1 100 adversarial coordinates against the reference's own Evaluate."""
    sf = capi.SceneFile(util.golden("textures.rtb"))
    ctx.upload_scene(sf)
    uv = np.fromfile(util.golden("textures.uv"), dtype=np.float32).reshape(-1, 2)
    want = np.fromfile(util.golden("textures.texeval"), dtype=np.float32).reshape(len(sf.textures), len(uv), 4)
    for t in range(len(sf.textures)):
        got = ctx.texture_batch(t, uv)
        util.assert_bit_identical(got, want[t], "texture %d" % t)
    with pytest.raises(capi.RtbError, match="out of range"):
        ctx.texture_batch(len(sf.textures), uv[:1])


def test_trace_batch_empty(ctx):
    ctx.upload_scene(capi.SceneFile(util.golden("cornell.rtb")))
    got = ctx.trace_batch(np.zeros(0, dtype=capi.RAY_DTYPE))
    assert len(got) == 0


@pytest.mark.parametrize("blue", [True, False])
def test_sampler_streams_integer_exact(ctx, blue):
    ps = capi.PassScript(util.golden("cornell.passes"))
    want = np.fromfile(util.golden("sampler_bn.u32" if blue else "sampler_nobn.u32"), dtype=np.uint32).reshape(ps.num_passes, 20, 24, 40)
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ctx.upload_scene(sf)
    ctx.resize(64, 64)
    cam = capi.camera_for(sf, 64, 64)
    ys, xs = np.mgrid[0:20, 0:24]
    xy = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.uint32)
    for p in range(ps.num_passes):
        got = ctx.sample_batch(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p, blue_noise=blue), xy, 40)
        assert (got.reshape(20, 24, 40) == want[p]).all(), "pass %d" % p


@pytest.mark.parametrize("scene", ["cornell", "mesh_small"])
def test_primary_rays_bit_exact(ctx, scene):
    sf = capi.SceneFile(util.golden(scene + ".rtb"))
    ps = capi.PassScript(util.golden(scene + ".passes"))
    ctx.upload_scene(sf)
    ctx.resize(64, 64)
    cam = capi.camera_for(sf, 64, 64)
    got = ctx.primary_rays(ctx.make_pass(cam, ps.seeds[1], ps.offsets[1], 1))
    want = np.fromfile(util.golden(scene + ".primary"), dtype=capi.RAY_DTYPE)
    for f in ("origin", "dir", "invDir", "originDivDir"):
        util.assert_bit_identical(got[f][:, :3], want[f][:, :3], "%s primary %s" % (scene, f))


def test_error_paths(ctx):
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    ctx.upload_scene(sf)
    with pytest.raises(capi.RtbError):
        ctx.resize(0, 10)                      # Viewport::Resize rejects empty sizes (Viewport.cpp:52-58)
    ctx.resize(32, 32)
    cam = capi.camera_for(sf, 32, 32)
    with pytest.raises(capi.RtbError):         # not enough seeded dimensions -> refuse instead of an RNG fallback
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[0][:8], ps.offsets[0], 0))
    with pytest.raises(capi.RtbError):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[0], ps.offsets[0], 0, renderer=7))
    c2 = capi.Context(device=0)
    c2.resize(16, 16)
    with pytest.raises(capi.RtbError):         # no scene uploaded (Viewport: "Missing renderer", Viewport.cpp:194-198)
        c2.render_pass(c2.make_pass(cam, ps.seeds[0], ps.offsets[0], 0))
    c2.close()


def test_empty_scene_is_black(ctx, tmp_path):
    """RenderingTest.EmptyScene (Tests/RaytracingTests.cpp:263-285): an empty scene renders exactly zero."""
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    sf.num_tlas = sf.num_objects = sf.num_lights = sf.num_global_lights = 0
    ctx.upload_scene(sf)
    ctx.resize(32, 32)
    cam = capi.camera_for(sf, 32, 32)
    for rid in (0, 1):
        ctx.reset()
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[0], ps.offsets[0], 0, renderer=rid))
        assert (ctx.read_sum() == 0.0).all()


def test_sharded_render_is_bit_identical(ctx):
    """Tile-interleaved shards (the multi-GPU decomposition) reassemble to exactly the single-context image."""
    scene = "mesh_small"
    meta, full, _ = render_golden_case(ctx, scene, "mis")
    w, h = meta["width"], meta["height"]
    sf = capi.SceneFile(util.golden(scene + ".rtb"))
    ps = capi.PassScript(util.golden(scene + ".passes"))
    cam = capi.camera_for(sf, w, h)
    acc = np.zeros_like(full)
    n = 3
    for s in range(n):
        c = capi.Context(device=0, num_shards=n, shard_index=s, tile_size=16)
        c.upload_scene(sf)
        c.resize(w, h)
        for p in range(meta["passes"]):
            c.render_pass(c.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
        part = c.read_sum()
        assert ((acc != 0) & (part != 0)).sum() == 0, "shards overlap"
        acc += part
        c.close()
    util.assert_bit_identical(acc, full, "sharded film")


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("scene,w,h,passes", [("cornell", 200, 152, 6), ("mesh_small", 160, 90, 5), ("materials", 96, 128, 3)])
def test_live_oracle_larger_ragged_sizes(ctx, tmp_path, scene, w, h, passes):
    """Same inputs through the real reference, live on this box's CPU, at sizes that are not tile multiples."""
    sj = os.path.join(util.GOLDEN, "scenes", scene + ".json")
    data = os.path.join(util.GOLDEN, "scenes") + "/"
    dims = 256 if scene == "materials" else 128
    sump, passp, flatp = str(tmp_path / "o.sum"), str(tmp_path / "o.passes"), str(tmp_path / "o.rtb")
    meta = util.run_rtref(["render", "--scene", sj, "--data", data, "--width", w, "--height", h, "--passes", passes, "--dims", dims,
                           "--depth", 8, "--strategy", "all", "--dump-sum", sump, "--dump-passes", passp], seed="777")
    util.run_rtref(["flat", "--scene", sj, "--data", data, "--width", w, "--height", h, "--out", flatp])
    sf = capi.SceneFile(flatp)
    ps = capi.PassScript(passp)
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    for p in range(passes):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
    util.assert_bit_identical(ctx.read_sum(), util.load_sum(sump, w, h), "%s live oracle" % scene)
    c = ctx.counters()
    assert (c["numRays"], c["numShadowRays"], c["numShadowRaysHit"]) == (meta["numRays"], meta["numShadowRays"], meta["numShadowRaysHit"])


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("mode", range(13))
def test_every_debug_mode_against_the_live_reference(ctx, tmp_path, mode):
    """DebugRenderer::RenderPixel, all 13 DebugRenderingMode values (camera light, triangle id, depth, position, frame
    vectors, texture coordinates, material parameters), on the textured scene: bit-exact against the reference run here."""
    scn = os.path.join(util.GOLDEN, "scenes")
    w, h = 56, 40
    sump, passp = str(tmp_path / "o.sum"), str(tmp_path / "o.passes")
    util.run_rtref(["render", "--scene", os.path.join(scn, "textures.json"), "--data", scn + "/", "--width", w, "--height", h, "--passes", 2, "--dims", 64,
                    "--renderer", "Debug", "--debug-mode", mode, "--dump-sum", sump, "--dump-passes", passp])
    sf = capi.SceneFile(util.golden("textures.rtb"))
    ps = capi.PassScript(passp)
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    for p in range(2):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p, renderer=capi.RTB_RENDERER_DEBUG, debug_mode=mode))
    util.assert_bit_identical(ctx.read_sum(), util.load_sum(sump, w, h), "debug mode %d" % mode)


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("bokeh,barrel", [(1, 0.01), (2, 0.01), (0, 0.08)])
def test_bokeh_shapes_and_barrel_distortion_against_the_live_reference(ctx, tmp_path, bokeh, barrel):
    """Camera::GenerateRay with hexagon / square bokeh and another barrel distortion factor (camera state the JSON schema
    cannot express; the harness sets it on the reference's Camera object, RtbCamera carries it to the device)."""
    scn = os.path.join(util.GOLDEN, "scenes")
    w, h, passes = 72, 56, 3
    common = ["--scene", os.path.join(scn, "cornell_dof.json"), "--data", scn + "/", "--width", w, "--height", h, "--bokeh", bokeh, "--barrel", barrel]
    sump, passp, flatp = str(tmp_path / "o.sum"), str(tmp_path / "o.passes"), str(tmp_path / "o.rtb")
    meta = util.run_rtref(["render", *common, "--passes", passes, "--dims", 128, "--depth", 8, "--strategy", "all", "--dump-sum", sump, "--dump-passes", passp])
    util.run_rtref(["flat", *common, "--out", flatp])
    sf = capi.SceneFile(flatp)
    ps = capi.PassScript(passp)
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    assert cam.bokehShape == bokeh
    for p in range(passes):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
    util.assert_bit_identical(ctx.read_sum(), util.load_sum(sump, w, h), "bokeh %d barrel %g" % (bokeh, barrel))
    assert ctx.counters()["numRays"] == meta["numRays"]


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
def test_texture_shaped_bokeh_against_the_live_reference(ctx, tmp_path):
    """Depth of field whose lens point is drawn from a bokeh bitmap's importance map (RtbSceneDesc::bokehCdf)."""
    scn = os.path.join(util.GOLDEN, "scenes")
    w, h, passes = 80, 60, 3
    common = ["--scene", os.path.join(scn, "cornell_bokeh.json"), "--data", scn + "/", "--width", w, "--height", h]
    sump, passp, flatp = str(tmp_path / "o.sum"), str(tmp_path / "o.passes"), str(tmp_path / "o.rtb")
    meta = util.run_rtref(["render", *common, "--passes", passes, "--dims", 128, "--depth", 8, "--strategy", "all", "--dump-sum", sump, "--dump-passes", passp])
    util.run_rtref(["flat", *common, "--out", flatp])
    sf = capi.SceneFile(flatp)
    assert sf.bokeh_cdf is not None and sf.camera.bokehShape == 4
    ps = capi.PassScript(passp)
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    for p in range(passes):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
    util.assert_bit_identical(ctx.read_sum(), util.load_sum(sump, w, h), "texture-shaped bokeh")
    assert ctx.counters()["numRays"] == meta["numRays"]
    # the same camera on a scene without a bokeh distribution is refused
    ctx.upload_scene(capi.SceneFile(util.golden("cornell.rtb")))
    with pytest.raises(capi.RtbError, match="no bokeh distribution"):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[0], ps.offsets[0], 0))


def test_gathered_shards_unshard_to_the_full_frame(ctx):
    """What bench.py does at N > 1, emulated on one GPU: per-shard compact films concatenated (the all-gather result)
    and scattered back to raster order by rtb_unshard_film; also checks the layout against raytracer_b200.sharding."""
    import torch
    from raytracer_b200 import sharding
    from bench import CudaArray

    scene = "cornell"
    meta, full, _ = render_golden_case(ctx, scene, "mis")
    w, h = meta["width"], meta["height"]
    sf = capi.SceneFile(util.golden(scene + ".rtb"))
    ps = capi.PassScript(util.golden(scene + ".passes"))
    cam = capi.camera_for(sf, w, h)
    n = 4
    parts, ctxs = [], []
    for s in range(n):
        c = capi.Context(device=0, num_shards=n, shard_index=s, tile_size=8)
        c.upload_scene(sf)
        c.resize(w, h)
        for p in range(meta["passes"]):
            c.render_pass(c.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
        c.synchronize()
        ptr, cnt = c.film_device_buffer(0)
        parts.append(torch.as_tensor(CudaArray(ptr, cnt), device="cuda").clone())
        ctxs.append(c)
    gathered = torch.cat(parts)
    torch.cuda.synchronize()
    img = ctxs[0].unshard_film(gathered.data_ptr())
    util.assert_bit_identical(img, full, "unsharded film")
    img2 = sharding.unshard(gathered.cpu().numpy().reshape(-1, 3), w, h, n, tile_size=8)
    util.assert_bit_identical(img2, full, "python unshard of the same buffers")
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("scene", ["cornell", "mesh_small"])
def test_batched_passes_are_bit_identical_to_single_passes(ctx, scene):
    """rtb_render_passes traces several passes in one wavefront; the film must equal pass-by-pass rendering exactly."""
    meta = json.load(open(util.golden("%s_mis.json" % scene)))
    w, h = meta["width"], meta["height"]
    sf = capi.SceneFile(util.golden(scene + ".rtb"))
    ps = capi.PassScript(util.golden(scene + ".passes"))
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    ctx.render_passes(cam, ps.seeds[:meta["passes"]], ps.offsets[:meta["passes"]], 0)
    for which, name in ((0, "sum"), (1, "sum2")):
        want = util.load_sum(util.golden("%s_mis.%s" % (scene, name)), w, h)
        util.assert_bit_identical(ctx.read_sum(which), want, "%s batched %s" % (scene, name))
    c = ctx.counters()
    assert (c["numRays"], c["numShadowRays"], c["numShadowRaysHit"], c["numPrimaryRays"]) == \
           (meta["numRays"], meta["numShadowRays"], meta["numShadowRaysHit"], meta["numPrimaryRays"])


@pytest.mark.parametrize("slots", ["1", "2", "4"])
@pytest.mark.parametrize("scene", ["mesh_small", "textures"])
def test_wavefronts_in_flight_are_bit_identical(scene, slots, monkeypatch):
    """Pass-by-pass rendering against ONE batched wavefront of the same passes: film, secondary film and ray counters must
    be equal, every time.  With RTB_SLOTS = 2 / 4 every pass goes to its own wavefront slot / stream and several
    wavefronts overlap on the GPU - the configuration in which round 1 saw a few hundred paths per frame go wrong.  Round
    2 bisected that to k_tail and then to the code ptxas generates for it at -O2 / -O3: the kernel is
    built at ptxas -O1 and stays in the launch chain of every context, and this test keeps watch (a library built with
    RTB_TAIL_PTXAS=-O3 reproduces the problem here)."""
    monkeypatch.setenv("RTB_SLOTS", slots)
    meta = json.load(open(util.golden("%s_mis.json" % scene)))
    sf = capi.SceneFile(util.golden(scene + ".rtb"))
    w = h = 256
    n = 8
    rng = np.random.default_rng(5)
    seeds = rng.integers(0, 2**32, (n, meta["dims"]), dtype=np.uint32)
    offs = (rng.normal(size=(n, 2)) * 0.5).astype(np.float32)

    def run(single, reads_between=False):
        with capi.Context(device=0) as c:
            c.upload_scene(sf)
            c.resize(w, h)
            cam = capi.camera_for(sf, w, h)
            if single:
                for p in range(n):
                    c.render_pass(c.make_pass(cam, seeds[p], offs[p], p))
                    if reads_between and p == 3:
                        c.read_sum()                # a join point in the middle of the frame
            else:
                c.render_passes(cam, seeds, offs, 0)
            return c.read_sum(), c.read_sum(1), c.counters()

    want = run(False)
    for rep in range(3):
        got = run(True, reads_between=(rep == 2))
        util.assert_bit_identical(got[0], want[0], "%s sum, repetition %d" % (scene, rep))
        util.assert_bit_identical(got[1], want[1], "%s secondary sum, repetition %d" % (scene, rep))
        for k in ("numRays", "numShadowRays", "numShadowRaysHit", "numPrimaryRays"):
            assert got[2][k] == want[2][k], k


@pytest.mark.parametrize("scene", ["mesh_small", "cornell"])
def test_film_snapshots_of_a_batched_call_equal_pass_by_pass_frames(ctx, scene):
    """rtb_film_snapshots: one batched rtb_render_passes call leaves the film as of EVERY pass; each snapshot must be
    the frame Viewport::GetSumBuffer would have returned after that pass (= pass-by-pass rendering), bit for bit."""
    meta = json.load(open(util.golden("%s_mis.json" % scene)))
    w, h, n = meta["width"], meta["height"], meta["passes"]
    sf = capi.SceneFile(util.golden(scene + ".rtb"))
    ps = capi.PassScript(util.golden(scene + ".passes"))
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    want = []
    for p in range(n):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
        want.append(ctx.read_sum().copy())
    ctx.reset()
    ptr, stride = ctx.film_snapshots(n - 1)                    # one entry short: the last pass of the call gets no snapshot
    ctx.render_passes(cam, ps.seeds[:n], ps.offsets[:n], 0)
    for p in range(n - 1):
        got = ctx.unshard_film(ptr + 4 * stride * p)           # a snapshot has the layout of the (single-shard) film
        util.assert_bit_identical(got, want[p], "%s: snapshot after pass %d" % (scene, p))
    util.assert_bit_identical(ctx.read_sum(), want[-1], "%s: film after the batched call" % scene)
    ctx.film_snapshots(0)


def test_async_read_back_snapshots_every_pass(ctx):
    """rtb_read_sum_async after pass k must deliver the film as of pass k even though pass k+1 is already running
    (double-buffered snapshot on the render stream, copy on a second stream); three read-backs in flight exercise
    the buffer reuse."""
    meta = json.load(open(util.golden("cornell_mis.json")))
    w, h = meta["width"], meta["height"]
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    n = min(3, meta["passes"])
    want = []
    for p in range(n):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
        want.append(ctx.read_sum().copy())
    ctx.reset()
    pinned = [capi.PinnedImage(w, h) for _ in range(n)]
    try:
        for p in range(n):
            ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
            ctx.read_sum_async(pinned[p])
        ctx.synchronize()
        for p in range(n):
            util.assert_bit_identical(pinned[p].array, want[p], "async read-back after pass %d" % p)
    finally:
        for im in pinned:
            im.close()


# ---- front buffer: Viewport::SetPostprocessParams / GetFrontBuffer (rtb_read_front) ---------------------------------------

@pytest.mark.parametrize("scene", ["c1_sphere_area", "cornell", "cornell_dof", "materials", "mesh_small", "textures"])
def test_front_buffer_golden(ctx, scene):
    """The reference's front buffer (default PostprocessParams, dithering off) of the golden renders, bit for bit."""
    meta, _, _ = render_golden_case(ctx, scene, "mis")
    got = ctx.read_front(capi.postprocess_desc(meta["passes"], dithering=0.0))
    want = np.fromfile(util.golden(scene + "_mis.front"), dtype=np.uint32).reshape(meta["height"], meta["width"])
    assert np.array_equal(got, want), "%d of %d front-buffer pixels differ" % ((got != want).sum(), want.size)


POSTPROCESS_CASES = [
    # scene, w, h, passes, exposure, contrast, saturation, tonemapper, filter, bloom
    ("cornell", 64, 48, 3, 0.0, 0.8, 0.98, 0, (1, 1, 1), 0.0),
    ("cornell", 64, 48, 3, 0.0, 0.8, 0.98, 1, (1, 1, 1), 0.0),
    ("cornell", 64, 48, 3, 0.0, 0.8, 0.98, 2, (1, 1, 1), 0.0),
    ("materials", 64, 48, 3, 1.5, 1.2, 0.3, 3, (0.9, 0.5, 1.25), 0.0),
    ("mesh_small", 40, 30, 2, 9.0, 0.8, 1.0, 1, (3, 0, 1e-30), 0.0),
    ("cornell", 33, 17, 1, 30.0, 40.0, 5.0, 2, (1, 1, 1), 0.0),       # overflowing curve: inf / nan lanes through ToBGR
    ("cornell", 33, 17, 1, -20.0, 3.0, 1.0, 0, (1, 1, 1), 0.0),
    ("cornell", 200, 196, 2, 0.0, 0.8, 0.98, 3, (1, 1, 1), 0.3),      # bloom: smallest image its widest box blur fits in
    ("materials", 256, 200, 1, 1.0, 0.8, 0.98, 3, (1, 1, 1), 0.8),
]


def reference_front_buffer(tmp_path, scene, w, h, passes, exposure, contrast, saturation, tonemapper, filt, bloom):
    """sum buffer, front buffer, pass script and flat scene of one live run of the reference."""
    sj = os.path.join(util.GOLDEN, "scenes", scene + ".json")
    data = os.path.join(util.GOLDEN, "scenes") + "/"
    dims = 256 if scene == "materials" else 128
    sump, frontp, passp, flatp = str(tmp_path / "o.sum"), str(tmp_path / "o.front"), str(tmp_path / "o.passes"), str(tmp_path / "o.rtb")
    util.run_rtref(["render", "--scene", sj, "--data", data, "--width", w, "--height", h, "--passes", passes, "--dims", dims, "--depth", 8,
                    "--strategy", "all", "--dump-sum", sump, "--dump-passes", passp, "--dump-front", frontp, "--pp-exposure", exposure,
                    "--pp-contrast", contrast, "--pp-saturation", saturation, "--pp-tonemapper", tonemapper,
                    "--pp-filter", "%g,%g,%g" % filt, "--pp-bloom", bloom])
    util.run_rtref(["flat", "--scene", sj, "--data", data, "--width", w, "--height", h, "--out", flatp])
    return sump, np.fromfile(frontp, dtype=np.uint32).reshape(h, w), passp, flatp


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("case", POSTPROCESS_CASES, ids=lambda c: "%s-%dx%d-tm%d-bloom%g-exp%g" % (c[0], c[1], c[2], c[7], c[9], c[4]))
def test_front_buffer_against_the_live_reference(ctx, tmp_path, case):
    """Every tone-mapping curve, colour filter / exposure / contrast / saturation extremes and bloom: the 8-bit image
    equals the reference's Viewport::GetFrontBuffer bit for bit."""
    scene, w, h, passes, exposure, contrast, saturation, tonemapper, filt, bloom = case
    sump, want, passp, flatp = reference_front_buffer(tmp_path, *case)
    sf = capi.SceneFile(flatp)
    ps = capi.PassScript(passp)
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    for p in range(passes):
        ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
    util.assert_bit_identical(ctx.read_sum(), util.load_sum(sump, w, h), "sum buffer")
    post = capi.postprocess_desc(passes, color_filter=filt, exposure=exposure, contrast=contrast, saturation=saturation, dithering=0.0,
                                 bloom=bloom, tonemapper=tonemapper)
    got = ctx.read_front(post)
    assert np.array_equal(got, want), "%d of %d front-buffer pixels differ" % ((got != want).sum(), want.size)
    if bloom > 0.0:
        post.bloomFactor = 0.0
        assert not np.array_equal(ctx.read_front(post), want), "bloom left the image unchanged"


def test_front_buffer_of_gathered_shards_with_bloom(ctx):
    """The all-gathered film of 4 shards through rtb_read_front equals the single-context front buffer, bloom included
    (bloom needs the whole image: one shard alone is refused)."""
    import torch
    from bench import CudaArray

    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    w, h, passes = 208, 200, 2
    cam = capi.camera_for(sf, w, h)

    def render(c):
        c.upload_scene(sf)
        c.resize(w, h)
        for p in range(passes):
            c.render_pass(c.make_pass(cam, ps.seeds[p], ps.offsets[p], p))

    render(ctx)
    post = capi.postprocess_desc(passes, dithering=0.0, bloom=0.4, exposure=0.5)
    plain = capi.postprocess_desc(passes, dithering=0.0)
    want, want_plain = ctx.read_front(post), ctx.read_front(plain)
    n, parts, ctxs = 4, [], []
    for s in range(n):
        c = capi.Context(device=0, num_shards=n, shard_index=s)
        render(c)
        c.synchronize()
        ptr, cnt = c.film_device_buffer(0)
        parts.append(torch.as_tensor(CudaArray(ptr, cnt), device="cuda").clone())
        ctxs.append(c)
    gathered = torch.cat(parts)
    torch.cuda.synchronize()
    assert np.array_equal(ctxs[1].read_front(post, gathered.data_ptr()), want)
    assert np.array_equal(ctxs[2].read_front(plain, gathered.data_ptr()), want_plain)
    # one shard on its own: its pixels tone-mapped, the others 0
    own = ctxs[3].read_front(plain)
    mask = own != 0
    assert mask.any() and not mask.all() and np.array_equal(own[mask], want_plain[mask])
    with pytest.raises(capi.RtbError, match="all-gathered"):
        ctxs[3].read_front(post)
    for c in ctxs:
        c.close()


def test_front_buffer_dithering_and_errors(ctx):
    meta, _, _ = render_golden_case(ctx, "cornell", "mis")
    clean = ctx.read_front(capi.postprocess_desc(meta["passes"], dithering=0.0))
    noisy = ctx.read_front(capi.postprocess_desc(meta["passes"], dithering=0.005, dither_seed=3))
    again = ctx.read_front(capi.postprocess_desc(meta["passes"], dithering=0.005, dither_seed=3))
    other = ctx.read_front(capi.postprocess_desc(meta["passes"], dithering=0.005, dither_seed=4))
    assert np.array_equal(noisy, again) and not np.array_equal(noisy, other) and not np.array_equal(noisy, clean)
    for shift in (0, 8, 16):
        d = ((noisy >> shift) & 255).astype(np.int32) - ((clean >> shift) & 255).astype(np.int32)
        assert np.abs(d).max() <= 2            # 0.005 * 255 = 1.3 levels either way
        assert abs(d.mean()) < 0.1             # zero-mean noise
    assert (noisy >> 24).max() == 0
    with pytest.raises(capi.RtbError, match="tonemapper"):
        ctx.read_front(capi.postprocess_desc(4, tonemapper=7))
    with pytest.raises(capi.RtbError, match="numPasses"):
        ctx.read_front(capi.postprocess_desc(0))
    with pytest.raises(capi.RtbError, match="widest box blur"):
        ctx.read_front(capi.postprocess_desc(4, bloom=0.5))          # 64 x 64 image
    ctx.resize(202, 200)
    with pytest.raises(capi.RtbError, match="multiple of 4"):
        ctx.read_front(capi.postprocess_desc(4, bloom=0.5))


# ---- adaptive rendering: block errors, active tiles, the host's block list (Viewport.cpp:552-732) ------------------------

@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("scene", ["cornell", "mesh_small"])
def test_block_errors_against_the_live_reference(ctx, tmp_path, scene):
    """rtb_block_errors == Viewport::ComputeBlockError for the whole image, single pixels / rows / columns and random
    rectangles, from this context's film and from the all-gathered films of 3 shards."""
    import torch
    from bench import CudaArray
    from tests.test_hostsim_parity import error_test_blocks

    w, h, passes = 100, 76, 4
    sj = os.path.join(util.GOLDEN, "scenes", scene + ".json")
    data = os.path.join(util.GOLDEN, "scenes") + "/"
    blocks = error_test_blocks(w, h)
    bp, refp, passp, flatp = (str(tmp_path / n) for n in ("blocks.bin", "o.err", "o.passes", "o.rtb"))
    blocks.tofile(bp)
    util.run_rtref(["render", "--scene", sj, "--data", data, "--width", w, "--height", h, "--passes", passes, "--dims", 128, "--depth", 8,
                    "--strategy", "all", "--dump-passes", passp, "--block-errors", bp, "--dump-block-errors", refp])
    util.run_rtref(["flat", "--scene", sj, "--data", data, "--width", w, "--height", h, "--out", flatp])
    want = np.fromfile(refp, dtype=np.float32)
    sf, ps = capi.SceneFile(flatp), capi.PassScript(passp)
    cam = capi.camera_for(sf, w, h)

    def render(c):
        c.upload_scene(sf)
        c.resize(w, h)
        c.render_passes(cam, ps.seeds[:passes], ps.offsets[:passes])

    render(ctx)
    assert ctx.block_errors(passes, blocks).tobytes() == want.tobytes()
    n, parts, parts2, ctxs = 3, [], [], []
    for s in range(n):
        c = capi.Context(device=0, num_shards=n, shard_index=s, tile_size=16)
        render(c)
        c.synchronize()
        for which, dst in ((0, parts), (1, parts2)):
            ptr, cnt = c.film_device_buffer(which)
            dst.append(torch.as_tensor(CudaArray(ptr, cnt), device="cuda").clone())
        ctxs.append(c)
    g, g2 = torch.cat(parts), torch.cat(parts2)
    torch.cuda.synchronize()
    assert ctxs[1].block_errors(passes, blocks, g.data_ptr(), g2.data_ptr()).tobytes() == want.tobytes()
    with pytest.raises(capi.RtbError, match="all-gathered"):
        ctxs[0].block_errors(passes, blocks)
    with pytest.raises(capi.RtbError, match="even"):
        ctx.block_errors(3, blocks)
    with pytest.raises(capi.RtbError, match="outside"):
        ctx.block_errors(4, [(0, w + 1, 0, h)])
    for c in ctxs:
        c.close()


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("subdivision,convergence,per_call", [(0.05, 0.01, 1), (0.3, 0.08, 1), (0.3, 0.08, 5)])
def test_adaptive_rendering_against_the_live_reference(tmp_path, subdivision, convergence, per_call):
    """AdaptiveRenderingSettings through the host viewport, which draws its own pass script (seeded like the oracle's
    fixed-seed build, so the two renders see the same random numbers): after every pass the progress and the block list are the
    reference's; at the end so are the sums (converged blocks stop receiving samples), the ray counters and the front
    buffer (a converged pixel keeps the value it had when it was last rendered).  The second parameter set converges
    the whole image: the last passes render nothing."""
    from raytracer_b200 import host_api

    w, h, passes, init, minb, maxb = 96, 80, 16, 2, 4, 32
    scn = os.path.join(util.GOLDEN, "scenes")
    sj = os.path.join(scn, "cornell.json")
    sump, sum2p, frontp, progp = (str(tmp_path / n) for n in ("o.sum", "o.sum2", "o.front", "o.progress"))
    meta = util.run_rtref(["render", "--scene", sj, "--data", scn + "/", "--width", w, "--height", h, "--passes", passes, "--dims", 256, "--depth", 8,
                           "--strategy", "all", "--adaptive", 1, "--adaptive-initial", init, "--adaptive-min-block", minb, "--adaptive-max-block", maxb,
                           "--adaptive-subdivision", subdivision, "--adaptive-convergence", convergence,
                           "--dump-progress", progp, "--dump-sum", sump, "--dump-sum2", sum2p, "--dump-front", frontp])
    records = util.read_progress(progp)
    scene = host_api.Scene(sj, scn + "/")
    vp = host_api.Viewport(scene, device=0)
    try:
        vp.set_renderer("Path Tracer MIS")
        vp.set_params(dimensions=256, max_ray_depth=8, strategy="all", seed=12345)
        vp.set_adaptive(True, init, minb, maxb, subdivision, convergence)
        vp.set_postprocess(dithering=0.0)
        vp.resize(w, h)
        vp.reset()              # the harness resets once more after Resize: the Halton scramble moves on (see test_host_scene)
        done = 0
        while done < passes:
            n = min(per_call, passes - done)
            vp.render(n)
            done += n
            r, mine = records[done - 1], vp.progress()
            assert mine["passesFinished"] == r["passesFinished"] == done
            assert (mine["activeBlocks"], mine["activePixels"]) == (r["activeBlocks"], r["activePixels"]), "pass %d" % done
            assert mine["converged"] == r["converged"] and np.isinf(mine["averageError"])
            assert np.array_equal(vp.blocks(), r["blocks"]), "block list after pass %d" % done
        util.assert_bit_identical(vp.read_sum(), util.load_sum(sump, w, h), "adaptive sum")
        util.assert_bit_identical(vp.read_sum(1), util.load_sum(sum2p, w, h), "adaptive secondary sum")
        c = vp.counters()
        assert (c["numPrimaryRays"], c["numRays"], c["numShadowRays"], c["numShadowRaysHit"]) == \
               (meta["numPrimaryRays"], meta["numRays"], meta["numShadowRays"], meta["numShadowRaysHit"])
        assert c["numPrimaryRays"] < w * h * passes
        got, want = vp.read_front(), np.fromfile(frontp, dtype=np.uint32).reshape(h, w)
        assert np.array_equal(got, want), "%d of %d front-buffer pixels differ" % ((got != want).sum(), want.size)
        # Viewport::Reset: every block is back
        vp.reset()
        assert len(vp.blocks()) == 9
        vp.render(2)
        assert vp.counters()["numPrimaryRays"] == 2 * w * h
    finally:
        vp.close()
        scene.close()


def test_active_tiles(ctx):
    """rtb_set_active_tiles: the passes that follow touch only the pixels inside the tiles."""
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    w = h = 64
    cam = capi.camera_for(sf, w, h)
    ctx.upload_scene(sf)
    ctx.resize(w, h)
    first_two = lambda: ctx.render_passes(cam, ps.seeds[0:2], ps.offsets[0:2], first_pass=0)
    last_two = lambda: ctx.render_passes(cam, ps.seeds[2:4], ps.offsets[2:4], first_pass=2)
    first_two()
    two = ctx.read_sum().copy()
    last_two()
    four = ctx.read_sum().copy()
    ctx.reset()
    first_two()
    ctx.set_active_tiles([(8, 40, 0, 17), (40, 64, 3, 4), (0, 1, 63, 64)])
    last_two()
    mask = np.zeros((h, w), dtype=bool)
    mask[0:17, 8:40] = True; mask[3:4, 40:64] = True; mask[63:64, 0:1] = True
    got = ctx.read_sum()
    util.assert_bit_identical(got[mask], four[mask], "active pixels")
    util.assert_bit_identical(got[~mask], two[~mask], "inactive pixels")
    assert ctx.counters()["numPrimaryRays"] == 2 * w * h + 2 * int(mask.sum())
    # nothing active: passes are counted, nothing is rendered
    ctx.set_active_tiles(np.zeros((0, 4), dtype=np.uint32))
    last_two()
    util.assert_bit_identical(ctx.read_sum(), got, "no active pixels")
    with pytest.raises(capi.RtbError, match="overlaps"):
        ctx.set_active_tiles([(0, 10, 0, 10), (9, 12, 9, 12)])
    with pytest.raises(capi.RtbError, match="outside"):
        ctx.set_active_tiles([(0, 65, 0, 10)])
    # the whole image again
    ctx.set_active_tiles(None)
    ctx.reset()
    ctx.render_passes(cam, ps.seeds[0:4], ps.offsets[0:4])
    util.assert_bit_identical(ctx.read_sum(), four, "all pixels active again")


@pytest.mark.parametrize("env", [{"RTB_SHADOW_KERNEL": "simple"}, {"RTB_SHADOW_KERNEL": "wide", "RTB_HOT_NODES": "0"}, {"RTB_SHADOW_KERNEL": "wide", "RTB_HOT_NODES": "64"},
                                 {"RTB_SHADOW_KERNEL": "wide", "RTB_HOT_NODES": "-1"}, {"RTB_SHADOW_KERNEL": "wide", "RTB_NO_WIDE": "1"},
                                 {"RTB_SHADOW_KERNEL": "split"}, {"RTB_TRAV_KERNEL": "simple"}, {"RTB_SORT_SHADE": "0"}])
def test_any_hit_walks_agree(env, monkeypatch):
    """The any-hit walks - the reference's binary tree in k_shadow and the 8-wide quantised tree (rtb_wide.h) in the
    persistent k_shadow_wide (the default) with 0 / 64 / all-that-fit nodes of the tree top staged in shared memory by a
    bulk copy, the split pair of kernels - as well as the one-ray-per-thread closest-hit kernel (k_traverse instead of the persistent k_traverse_p) and k_shade without
    its material binning give the reference's film and ray counters, bit for bit.  (rtb_trace_batch and the CPU build of
    the headers exercise the wide walk inline as well.)"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for scene in ("mesh_small", "textures", "materials"):
        with capi.Context(device=0) as c:
            meta, img, counters = render_golden_case(c, scene, "mis")
        want = util.load_sum(util.golden("%s_mis.sum" % scene), meta["width"], meta["height"])
        util.assert_bit_identical(img, want, "%s/mis with %s" % (scene, env))
        assert counters["numShadowRays"] == meta["numShadowRays"] and counters["numShadowRaysHit"] == meta["numShadowRaysHit"]
        assert counters["numRays"] == meta["numRays"]


@pytest.mark.parametrize("graph", ["1", "0"])
def test_wavefront_launch_chain_as_a_cuda_graph(graph, monkeypatch):
    """A context on a stream of its own replays the launch chain of a wavefront as a CUDA graph (captured on the first
    pass, relaunched while the parameters stay the same): the film and counters are the golden ones, graph or no graph,
    pass by pass and batched, MIS and PT."""
    import torch
    monkeypatch.setenv("RTB_GRAPH", graph)
    stream = torch.cuda.Stream()
    for scene, renderer in (("mesh_small", "mis"), ("cornell", "pt"), ("materials", "mis")):
        meta = json.load(open(util.golden("%s_%s.json" % (scene, renderer))))
        w, h = meta["width"], meta["height"]
        sf = capi.SceneFile(util.golden(scene + ".rtb"))
        ps = capi.PassScript(util.golden(scene + ".passes"))
        rid, dbg = util.RENDERER_IDS[renderer]
        want = util.load_sum(util.golden("%s_%s.sum" % (scene, renderer)), w, h)
        for batched in (False, True):
            with capi.Context(device=0, stream=stream.cuda_stream) as c:
                c.upload_scene(sf)
                c.resize(w, h)
                cam = capi.camera_for(sf, w, h)
                if batched:
                    c.render_passes(cam, ps.seeds[:meta["passes"]], ps.offsets[:meta["passes"]], 0, renderer=rid, debug_mode=dbg, max_depth=8)
                else:
                    for p in range(meta["passes"]):
                        c.render_pass(c.make_pass(cam, ps.seeds[p], ps.offsets[p], p, renderer=rid, debug_mode=dbg, max_depth=8))
                util.assert_bit_identical(c.read_sum(), want, "%s/%s graph=%s batched=%s" % (scene, renderer, graph, batched))
                counters = c.counters()
            assert counters["numRays"] == meta["numRays"] and counters["numShadowRays"] == meta["numShadowRays"]


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("scene,passes", [("materials", 400), ("mesh_small", 300)])
def test_single_light_sampling_statistics_against_the_live_reference(tmp_path, scene, passes):
    """LightSamplingStrategy::Single (the reference's default, Context.h:82, PathTracerMIS.cpp:128-146): one light per
    vertex, picked uniformly, contribution scaled by the number of lights.  The reference picks with its per-thread
    generator, so the bar is statistical: the converged mean per channel and the image at 8 x 8-pixel blocks agree with
    the reference's `single` render, both agree with the `all` estimator (same expectation), and the shadow-ray count is
    the reference's to within the sampling noise (one shadow ray per vertex at most, instead of one per light)."""
    w = h = 48
    scn = os.path.join(util.GOLDEN, "scenes")
    common = ["--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", w, "--height", h]
    refp, passp, flatp = (str(tmp_path / n) for n in ("ref.sum", "ref.passes", "scene.rtb"))
    meta = util.run_rtref(["render", *common, "--passes", passes, "--dims", 256, "--depth", 8, "--strategy", "single", "--dump-sum", refp, "--dump-passes", passp])
    util.run_rtref(["flat", *common, "--out", flatp])
    ref = util.load_sum(refp, w, h) / passes
    sf = capi.SceneFile(flatp)
    ps = capi.PassScript(passp)
    imgs, counters = {}, {}
    for name, strategy in (("single", capi.RTB_LIGHT_SAMPLING_SINGLE), ("all", capi.RTB_LIGHT_SAMPLING_ALL)):
        with capi.Context(device=0) as c:
            c.upload_scene(sf)
            c.resize(w, h)
            c.render_passes(capi.camera_for(sf, w, h), ps.seeds[:passes], ps.offsets[:passes], 0, max_depth=8, strategy=strategy)
            imgs[name], counters[name] = c.read_sum() / passes, c.counters()
    blocks = lambda x: x.reshape(6, 8, 6, 8, 3).mean(axis=(1, 3)).ravel()
    for name in ("single", "all"):
        got = imgs[name]
        assert np.isfinite(got).all() and (got >= 0).all()
        ma, mb = ref.mean(axis=(0, 1)), got.mean(axis=(0, 1))
        assert (np.abs(ma - mb) <= 0.03 * ma).all(), (name, ma, mb)
        assert np.corrcoef(blocks(ref), blocks(got))[0, 1] >= 0.995, name
    # the camera paths are the reference's (same sampler), so the path-ray count is close; shadow rays: at most one per vertex
    assert counters["single"]["numPrimaryRays"] == meta["numPrimaryRays"] == w * h * passes
    assert abs(counters["single"]["numShadowRays"] - meta["numShadowRays"]) <= 0.02 * meta["numShadowRays"]
    assert counters["single"]["numShadowRays"] < counters["all"]["numShadowRays"]


@pytest.mark.parametrize("w,h,tile", [(64, 64, 8), (96, 40, 32), (62, 50, 8)])
def test_shards_scattered_into_one_mapped_host_frame(w, h, tile):
    """rtb_scatter_shard - what every rank of an N-GPU end-to-end run does with its own pixels: four shard contexts write
    into ONE page-locked, device-mapped raster frame; the frame equals the unsharded film bit for bit.  Widths that are a
    multiple of four take k_scatter_shard_rows (16-byte stores along the raster rows), 62 the pixel-by-pixel fallback."""
    lib = capi.load_library()
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    cam = capi.camera_for(sf, w, h)
    passes = 3
    with capi.Context(device=0) as c:
        c.upload_scene(sf); c.resize(w, h)
        c.render_passes(cam, ps.seeds[:passes], ps.offsets[:passes], 0)
        full, full2 = c.read_sum(), c.read_sum(1)
    frame, frame2 = capi.PinnedImage(w, h), capi.PinnedImage(w, h)
    frame.array[:] = -1.0; frame2.array[:] = -1.0
    n = 4
    for s in range(n):
        with capi.Context(device=0, num_shards=n, shard_index=s, tile_size=tile) as c:
            c.upload_scene(sf); c.resize(w, h)
            c.render_passes(cam, ps.seeds[:passes], ps.offsets[:passes], 0)
            for which, f in ((0, frame), (1, frame2)):
                dev = lib.rtb_host_device_pointer(C.c_void_p(f.ptr))
                assert dev, lib.rtb_last_error()
                rc = lib.rtb_scatter_shard(c.h, None, which, C.c_void_p(dev))
                assert rc == 0, lib.rtb_last_error()
            c.synchronize()
    util.assert_bit_identical(frame.array, full, "scattered host frame")
    util.assert_bit_identical(frame2.array, full2, "scattered host frame, secondary sum")
    frame.close(); frame2.close()


def test_snapshots_scattered_on_the_second_stream_under_the_next_passes():
    """rtb_scatter_shard_async - the end-to-end path of an N-GPU run: two shard contexts render groups of passes with film
    snapshots; after every group each pass' frame is scattered into a page-locked raster frame on the context's SECOND stream
    while the next group already renders (the next k_film waits for the scatter before it overwrites the snapshots).  The
    frame of every pass of every group equals the unsharded film after that many passes, bit for bit."""
    lib = capi.load_library()
    sf = capi.SceneFile(util.golden("mesh_small.rtb"))
    ps = capi.PassScript(util.golden("mesh_small.passes"))
    w, h, group, groups, n = 128, 96, 2, 2, 2
    cam = capi.camera_for(sf, w, h)
    want = []
    with capi.Context(device=0) as c:
        c.upload_scene(sf); c.resize(w, h)
        for p in range(group * groups):
            c.render_passes(cam, ps.seeds[p:p + 1], ps.offsets[p:p + 1], p)
            want.append(c.read_sum())
    frames = [capi.PinnedImage(w, h) for _ in range(group * groups)]
    for f in frames:
        f.array[:] = -1.0
    for s in range(n):
        with capi.Context(device=0, num_shards=n, shard_index=s) as c:
            c.upload_scene(sf); c.resize(w, h)
            snap, stride = C.c_void_p(), C.c_uint64()
            assert lib.rtb_film_snapshots(c.h, group, C.byref(snap), C.byref(stride)) == 0, lib.rtb_last_error()
            for g in range(groups):
                c.render_passes(cam, ps.seeds[g * group:(g + 1) * group], ps.offsets[g * group:(g + 1) * group], g * group)
                for i in range(group):
                    dev = lib.rtb_host_device_pointer(C.c_void_p(frames[g * group + i].ptr))
                    rc = lib.rtb_scatter_shard_async(c.h, C.c_void_p(snap.value + 4 * stride.value * i), 0, C.c_void_p(dev))
                    assert rc == 0, lib.rtb_last_error()
            c.synchronize()
            lib.rtb_film_snapshots(c.h, 0, None, None)
    for p, f in enumerate(frames):
        util.assert_bit_identical(f.array, want[p], "frame of pass %d, scattered asynchronously" % p)
        f.close()
