"""CPU-side checks (no GPU): the DEVICE headers compiled with g++ (tests/hostsim, a test tool that is not part of the
product) reproduce the golden fixtures of the real reference bit for bit.  This is how the numerics of the CUDA code
are debugged in the authoring container, where there is no GPU; the `-m gpu` tests then check the real kernels."""
import json
import os
import subprocess

import numpy as np
import pytest

from raytracer_b200 import capi
from tests import util

HOSTSIM_DIR = os.path.join(util.ROOT, "tests", "hostsim")
HOSTSIM = os.path.join(HOSTSIM_DIR, "hostsim")


@pytest.fixture(scope="module")
def hostsim(tmp_path_factory):
    subprocess.run(["bash", os.path.join(HOSTSIM_DIR, "build.sh")], check=True)
    bn = tmp_path_factory.mktemp("bn") / "bn.dat"
    capi.load_blue_noise().tofile(str(bn))     # interleaved RGBA16 layout the sampler reads
    return str(bn)


@pytest.mark.parametrize("scene,renderer", util.GOLDEN_RENDERS)
def test_hostsim_sum_buffer_bit_exact(hostsim, tmp_path, scene, renderer):
    meta = json.load(open(util.golden("%s_%s.json" % (scene, renderer))))
    rid, dbg = util.RENDERER_IDS[renderer]
    out = str(tmp_path / "out.sum")
    r = subprocess.run([HOSTSIM, "render", util.golden(scene + ".rtb"), util.golden(scene + ".passes"), str(meta["width"]), str(meta["height"]),
                        out, str(rid), "8", str(dbg), "1", hostsim], stdout=subprocess.PIPE, text=True, check=True)
    counters = json.loads(r.stdout.strip().splitlines()[-1])
    want = util.load_sum(util.golden("%s_%s.sum" % (scene, renderer)), meta["width"], meta["height"])
    util.assert_bit_identical(util.load_sum(out, meta["width"], meta["height"]), want, "%s/%s" % (scene, renderer))
    if renderer in ("mis", "pt"):
        assert counters["numRays"] == meta["numRays"]
        assert counters["numShadowRays"] == meta["numShadowRays"]
        assert counters["numShadowRaysHit"] == meta["numShadowRaysHit"]


@pytest.mark.parametrize("scene,renderer", util.REFERENCE_SCENE_RENDERS)
def test_hostsim_reference_test_scenes(hostsim, tmp_path, scene, renderer):
    """The device headers on the reference's own Data/TestScenes (tests/golden/refscenes)."""
    e = util.reference_scene_manifest()[scene]
    base = os.path.join(util.GOLDEN, "refscenes", scene)
    out = str(tmp_path / "out.sum")
    rid, dbg = util.RENDERER_IDS[renderer]
    r = subprocess.run([HOSTSIM, "render", base + ".rtb", base + ".passes", str(e["width"]), str(e["height"]), out, str(rid), "8", str(dbg), "1", hostsim],
                       stdout=subprocess.PIPE, text=True, check=True)
    counters = json.loads(r.stdout.strip().splitlines()[-1])
    util.assert_bit_identical(util.load_sum(out, e["width"], e["height"]), util.load_sum("%s_%s.sum" % (base, renderer), e["width"], e["height"]), "%s/%s" % (scene, renderer))
    for k in ("numRays", "numShadowRays", "numShadowRaysHit"):
        assert counters[k] == e[renderer][k], k


@pytest.mark.parametrize("scene", ["cornell", "materials", "mesh_small"])
def test_hostsim_trace_batch(hostsim, tmp_path, scene):
    out = str(tmp_path / "hits")
    subprocess.run([HOSTSIM, "trace", util.golden(scene + ".rtb"), util.golden(scene + ".rays"), out], check=True)
    got = np.fromfile(out, dtype=capi.HIT_DTYPE)
    want = np.fromfile(util.golden(scene + ".hits"), dtype=capi.HIT_DTYPE)
    assert (got["objectId"] == want["objectId"]).all()
    util.assert_bit_identical(got["distance"], want["distance"], "distance")
    subprocess.run([HOSTSIM, "trace", util.golden(scene + ".rtb"), util.golden(scene + ".rays"), out, util.golden(scene + ".tmax")], check=True)
    got = np.fromfile(out, dtype=capi.HIT_DTYPE)
    want = np.fromfile(util.golden(scene + ".shadow_hits"), dtype=capi.HIT_DTYPE)
    assert (got["occluded"] == want["occluded"]).all()


def test_hostsim_texture_evaluation(hostsim, tmp_path):
    """ITexture::Evaluate of all 12 textures of the textured scene at 1 100 adversarial coordinates (negative, exactly on
    texel / tile boundaries, huge, tiny): the device header against the reference's own Evaluate (textures.texeval)."""
    out = str(tmp_path / "t.f32")
    subprocess.run([HOSTSIM, "texeval", util.golden("textures.rtb"), util.golden("textures.uv"), out], check=True)
    got = np.fromfile(out, dtype=np.uint32)
    want = np.fromfile(util.golden("textures.texeval"), dtype=np.uint32)
    assert got.shape == want.shape
    bad = np.argwhere(got != want)
    assert len(bad) == 0, "%d of %d floats differ, first at texture %d point %d" % (len(bad), got.size, bad[0][0] // (4 * 1100), (bad[0][0] // 4) % 1100)


@pytest.mark.parametrize("blue", [True, False])
def test_hostsim_sampler(hostsim, tmp_path, blue):
    out = str(tmp_path / "s.u32")
    subprocess.run([HOSTSIM, "sampler", util.golden("cornell.passes"), "24", "20", "40", out, hostsim if blue else "-"], check=True)
    want = np.fromfile(util.golden("sampler_bn.u32" if blue else "sampler_nobn.u32"), dtype=np.uint32)
    assert (np.fromfile(out, dtype=np.uint32) == want).all()


@pytest.mark.parametrize("scene", ["cornell", "mesh_small"])
def test_hostsim_primary_rays(hostsim, tmp_path, scene):
    out = str(tmp_path / "p.rays")
    subprocess.run([HOSTSIM, "primary", util.golden(scene + ".rtb"), util.golden(scene + ".passes"), "1", "64", "64", out], check=True)
    got = np.fromfile(out, dtype=capi.RAY_DTYPE)
    want = np.fromfile(util.golden(scene + ".primary"), dtype=capi.RAY_DTYPE)
    for f in ("origin", "dir", "invDir", "originDivDir"):
        util.assert_bit_identical(got[f][:, :3], want[f][:, :3], f)


DEBUG_MODES = {0: "CameraLight", 1: "TriangleID", 2: "Depth", 3: "Position", 4: "Normals", 5: "Tangents", 6: "Bitangents",
               7: "TexCoords", 8: "BaseColor", 9: "Emission", 10: "Roughness", 11: "Metalness", 12: "IoR"}


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("mode", sorted(DEBUG_MODES))
def test_hostsim_every_debug_mode_against_the_live_reference(hostsim, tmp_path, mode):
    """DebugRenderer::RenderPixel (DebugRenderer.cpp:26-178), all 13 DebugRenderingMode values, on the textured scene."""
    scn = os.path.join(util.GOLDEN, "scenes")
    sump, passp = str(tmp_path / "o.sum"), str(tmp_path / "o.passes")
    util.run_rtref(["render", "--scene", os.path.join(scn, "textures.json"), "--data", scn + "/", "--width", 48, "--height", 40, "--passes", 2, "--dims", 64,
                    "--renderer", "Debug", "--debug-mode", mode, "--dump-sum", sump, "--dump-passes", passp])
    out = str(tmp_path / "mine.sum")
    subprocess.run([HOSTSIM, "render", util.golden("textures.rtb"), passp, "48", "40", out, "2", "8", str(mode), "1", hostsim], stdout=subprocess.PIPE, check=True)
    util.assert_bit_identical(util.load_sum(out, 48, 40), util.load_sum(sump, 48, 40), "debug mode %s" % DEBUG_MODES[mode])


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("bokeh,barrel", [(1, 0.01), (2, 0.01), (0, 0.08), (1, 0.0)])
def test_hostsim_bokeh_shapes_and_barrel_distortion(hostsim, tmp_path, bokeh, barrel):
    """Camera::GenerateRay with the hexagon / square bokeh (Camera.cpp:195-217, SamplingHelpers.cpp:40-57) and other
    barrel distortion factors (Camera.cpp:87-92): camera state the JSON schema cannot express, set on the reference's
    Camera object by the harness and carried to the device in RtbCamera."""
    scn = os.path.join(util.GOLDEN, "scenes")
    common = ["--scene", os.path.join(scn, "cornell_dof.json"), "--data", scn + "/", "--width", 48, "--height", 48, "--bokeh", bokeh, "--barrel", barrel]
    sump, passp, flatp = str(tmp_path / "o.sum"), str(tmp_path / "o.passes"), str(tmp_path / "o.rtb")
    meta = util.run_rtref(["render", *common, "--passes", 3, "--dims", 128, "--depth", 8, "--strategy", "all", "--dump-sum", sump, "--dump-passes", passp])
    util.run_rtref(["flat", *common, "--out", flatp])
    out = str(tmp_path / "mine.sum")
    r = subprocess.run([HOSTSIM, "render", flatp, passp, "48", "48", out, "1", "8", "0", "1", hostsim], stdout=subprocess.PIPE, text=True, check=True)
    util.assert_bit_identical(util.load_sum(out, 48, 48), util.load_sum(sump, 48, 48), "bokeh %d barrel %g" % (bokeh, barrel))
    assert json.loads(r.stdout.strip().splitlines()[-1])["numRays"] == meta["numRays"]


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
def test_hostsim_texture_shaped_bokeh(hostsim, tmp_path):
    """Depth of field with a bokeh bitmap (Camera::GenerateBokeh, BokehShape::Texture -> BitmapTexture::Sample ->
    Distribution::SampleDiscrete): the importance map the host builds equals the reference's, and so does the render."""
    from raytracer_b200 import host_api
    scn = os.path.join(util.GOLDEN, "scenes")
    sj = os.path.join(scn, "cornell_bokeh.json")
    common = ["--scene", sj, "--data", scn + "/", "--width", 48, "--height", 48]
    sump, passp, flatp, minep = str(tmp_path / "o.sum"), str(tmp_path / "o.passes"), str(tmp_path / "o.rtb"), str(tmp_path / "mine.rtb")
    meta = util.run_rtref(["render", *common, "--passes", 3, "--dims", 128, "--depth", 8, "--strategy", "all", "--dump-sum", sump, "--dump-passes", passp])
    util.run_rtref(["flat", *common, "--out", flatp])
    host_api.Scene(sj, scn + "/").save_flat(minep, 48, 48)
    ref, mine = capi.SceneFile(flatp), capi.SceneFile(minep)
    assert ref.bokeh_size == mine.bokeh_size == (11, 9) and ref.camera.bokehShape == mine.camera.bokehShape == 4
    assert ref.bokeh_cdf.tobytes() == mine.bokeh_cdf.tobytes(), "importance map differs from BitmapTexture::MakeSamplable"
    out = str(tmp_path / "mine.sum")
    r = subprocess.run([HOSTSIM, "render", minep, passp, "48", "48", out, "1", "8", "0", "1", hostsim], stdout=subprocess.PIPE, text=True, check=True)
    util.assert_bit_identical(util.load_sum(out, 48, 48), util.load_sum(sump, 48, 48), "texture-shaped bokeh")
    assert json.loads(r.stdout.strip().splitlines()[-1])["numRays"] == meta["numRays"]


def hostsim_front(sump, w, h, passes, out, exposure=0.0, contrast=0.8, saturation=0.98, dither=0.0, tonemapper=3, filt=(1, 1, 1), bloom=0.0):
    subprocess.run([HOSTSIM, "front", sump, str(w), str(h), str(passes), repr(float(exposure)), repr(float(contrast)), repr(float(saturation)), repr(float(dither)),
                    str(tonemapper), "%r,%r,%r" % tuple(float(f) for f in filt), out, repr(float(bloom))], check=True)
    return np.fromfile(out, dtype=np.uint32).reshape(h, w)


@pytest.mark.parametrize("scene", ["c1_sphere_area", "cornell", "cornell_dof", "materials", "mesh_small", "textures"])
def test_hostsim_front_buffer_golden(hostsim, tmp_path, scene):
    """Viewport::PostProcessTile (Viewport.cpp:492-552) with the default PostprocessParams on the golden sum buffers:
    the device header's arithmetic gives the reference's 8-bit front buffer."""
    meta = json.load(open(util.golden(scene + "_mis.json")))
    w, h = meta["width"], meta["height"]
    got = hostsim_front(util.golden(scene + "_mis.sum"), w, h, meta["passes"], str(tmp_path / "mine.front"))
    want = np.fromfile(util.golden(scene + "_mis.front"), dtype=np.uint32).reshape(h, w)
    assert np.array_equal(got, want), "%d of %d front-buffer pixels differ" % ((got != want).sum(), want.size)
    assert len(np.unique(want)) > 50 and (want >> 24).max() == 0


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("case", [
    # scene, w, h, passes, exposure, contrast, saturation, tonemapper, filter, bloom
    ("cornell", 64, 48, 3, 0.0, 0.8, 0.98, 0, (1, 1, 1), 0.0),
    ("cornell", 64, 48, 3, 0.0, 0.8, 0.98, 1, (1, 1, 1), 0.0),
    ("cornell", 64, 48, 3, 0.0, 0.8, 0.98, 2, (1, 1, 1), 0.0),
    ("materials", 64, 48, 3, 1.5, 1.2, 0.3, 3, (0.9, 0.5, 1.25), 0.0),
    ("textures", 50, 37, 2, -2.0, 0.5, 1.6, 2, (1, 1, 1), 0.0),
    ("mesh_small", 40, 30, 2, 9.0, 0.8, 1.0, 1, (3, 0, 1e-30), 0.0),
    ("cornell", 33, 17, 1, 30.0, 40.0, 5.0, 2, (1, 1, 1), 0.0),
    ("cornell", 33, 17, 1, 30.0, 40.0, 5.0, 3, (1, 1, 1), 0.0),
    ("cornell", 33, 17, 1, -20.0, 3.0, 1.0, 0, (1, 1, 1), 0.0),
    ("cornell", 200, 196, 2, 0.0, 0.8, 0.98, 3, (1, 1, 1), 0.3),
    ("materials", 256, 200, 1, 1.0, 0.8, 0.98, 3, (1, 1, 1), 0.8),
], ids=lambda c: "%s-%dx%d-tm%d-bloom%g-exp%g" % (c[0], c[1], c[2], c[7], c[9], c[4]))
def test_hostsim_front_buffer_against_the_live_reference(hostsim, tmp_path, case):
    """Tone-mapping curves, exposure / contrast / saturation / colour-filter extremes (overflowing lanes included) and
    bloom (Bitmap::GaussianBlur's running sums, five levels): bit-identical to Viewport::GetFrontBuffer."""
    scene, w, h, passes, exposure, contrast, saturation, tonemapper, filt, bloom = case
    scn = os.path.join(util.GOLDEN, "scenes")
    sump, frontp = str(tmp_path / "o.sum"), str(tmp_path / "o.front")
    util.run_rtref(["render", "--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", w, "--height", h, "--passes", passes,
                    "--dims", 256, "--depth", 8, "--dump-sum", sump, "--dump-front", frontp, "--pp-exposure", exposure, "--pp-contrast", contrast,
                    "--pp-saturation", saturation, "--pp-tonemapper", tonemapper, "--pp-filter", "%g,%g,%g" % filt, "--pp-bloom", bloom])
    want = np.fromfile(frontp, dtype=np.uint32).reshape(h, w)
    got = hostsim_front(sump, w, h, passes, str(tmp_path / "mine.front"), exposure, contrast, saturation, 0.0, tonemapper, filt, bloom)
    assert np.array_equal(got, want), "%d of %d front-buffer pixels differ" % ((got != want).sum(), want.size)
    if bloom > 0.0:
        assert not np.array_equal(hostsim_front(sump, w, h, passes, str(tmp_path / "nobloom.front"), exposure, contrast, saturation, 0.0, tonemapper, filt), want)


def error_test_blocks(w, h, seed=3):
    """The whole image (Viewport::ComputeError), single pixels, single rows / columns and random rectangles."""
    rng = np.random.default_rng(seed)
    blocks = [(0, w, 0, h), (0, 1, 0, 1), (w - 1, w, h - 1, h), (0, w, 5, 6), (7, 8, 0, h)]
    while len(blocks) < 48:
        x0, x1 = sorted(int(v) for v in rng.integers(0, w + 1, 2))
        y0, y1 = sorted(int(v) for v in rng.integers(0, h + 1, 2))
        if x0 < x1 and y0 < y1:
            blocks.append((x0, x1, y0, y1))
    return np.array(blocks, dtype=np.uint32)


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("scene", ["cornell", "mesh_small", "materials"])
def test_hostsim_block_errors_against_the_live_reference(hostsim, tmp_path, scene):
    """Viewport::ComputeBlockError (Viewport.cpp:552-582): per-pixel estimate, row sums left to right, block sum top to
    bottom - the float sums are order dependent, so the device code adds in the reference's order."""
    w, h, passes = 100, 76, 4
    scn = os.path.join(util.GOLDEN, "scenes")
    blocks = error_test_blocks(w, h)
    bp, sump, sum2p, refp, minep, progp = (str(tmp_path / n) for n in ("blocks.bin", "o.sum", "o.sum2", "o.err", "mine.err", "o.progress"))
    blocks.tofile(bp)
    util.run_rtref(["render", "--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", w, "--height", h, "--passes", passes, "--dims", 256,
                    "--dump-sum", sump, "--dump-sum2", sum2p, "--block-errors", bp, "--dump-block-errors", refp, "--dump-progress", progp])
    subprocess.run([HOSTSIM, "blockerrors", sump, sum2p, str(w), str(h), str(passes), bp, minep], check=True)
    want, got = np.fromfile(refp, dtype=np.float32), np.fromfile(minep, dtype=np.float32)
    assert len(want) == len(blocks) and want.tobytes() == got.tobytes()
    assert np.isfinite(want).all() and want[0] > 0
    # RenderingProgress::averageError after the last (even) pass is the whole image's error (Viewport.cpp:179-183)
    assert util.read_progress(progp)[-1]["averageError"] == want[0]


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("subdivision,convergence", [(0.05, 0.01), (0.3, 0.08)])
def test_block_list_revision_against_the_live_reference(tmp_path, subdivision, convergence):
    """Viewport::UpdateBlocksList (Viewport.cpp:647-732), host logic only: fed the reference's own block lists and block
    errors pass after pass, ViewportB200::ReviseBlocks produces the reference's next list (order included)."""
    from raytracer_b200 import host_api
    w, h, passes, init, minb, maxb = 96, 80, 16, 2, 4, 32
    scn = os.path.join(util.GOLDEN, "scenes")
    progp = str(tmp_path / "o.progress")
    util.run_rtref(["render", "--scene", os.path.join(scn, "cornell.json"), "--data", scn + "/", "--width", w, "--height", h, "--passes", passes, "--dims", 256,
                    "--adaptive", 1, "--adaptive-initial", init, "--adaptive-min-block", minb, "--adaptive-max-block", maxb,
                    "--adaptive-subdivision", subdivision, "--adaptive-convergence", convergence, "--dump-progress", progp])
    records = util.read_progress(progp)
    assert len(records) == passes and len(records[0]["blocks"]) == 9
    previous, revisions, dropped = records[0]["blocks"], 0, False
    for r in records[1:]:
        if len(r["previousErrors"]) and r["passesFinished"] >= init:
            mine = host_api.revise_blocks(previous, r["previousErrors"], minb, subdivision, convergence)
            assert np.array_equal(mine, r["blocks"]), "block list after pass %d" % r["passesFinished"]
            area = int(((mine[:, 1] - mine[:, 0]) * (mine[:, 3] - mine[:, 2])).sum())
            assert area == r["activePixels"] and len(mine) == r["activeBlocks"]
            revisions += 1
            dropped = dropped or area < w * h
        else:
            assert np.array_equal(previous, r["blocks"])
        previous = r["blocks"]
    assert revisions >= 5 and dropped


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
def test_active_block_overlay_against_the_live_reference(tmp_path):
    """Viewport::VisualizeActiveBlocks (Viewport.cpp:734-771): the reference's front buffer + its final block list through
    ViewportB200::VisualizeBlocks gives the reference's overlay image (8-bit blend, borders, overlapping order)."""
    from raytracer_b200 import host_api
    w, h, passes = 96, 80, 12
    scn = os.path.join(util.GOLDEN, "scenes")
    progp, frontp, overp = str(tmp_path / "o.progress"), str(tmp_path / "o.front"), str(tmp_path / "o.overlay")
    util.run_rtref(["render", "--scene", os.path.join(scn, "cornell.json"), "--data", scn + "/", "--width", w, "--height", h, "--passes", passes, "--dims", 256,
                    "--adaptive", 1, "--adaptive-initial", 2, "--adaptive-min-block", 4, "--adaptive-max-block", 32, "--adaptive-subdivision", 0.05,
                    "--adaptive-convergence", 0.01, "--dump-progress", progp, "--dump-front", frontp, "--dump-overlay", overp])
    blocks = util.read_progress(progp)[-1]["blocks"]
    front = np.fromfile(frontp, dtype=np.uint32).reshape(h, w)
    want = np.fromfile(overp, dtype=np.uint32).reshape(h, w)
    assert 9 < len(blocks) and not np.array_equal(front, want)
    assert np.array_equal(host_api.visualize_blocks(front, blocks), want)


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
def test_golden_fixtures_reproduce_from_the_reference(tmp_path):
    """The committed fixtures are what the real reference produces today (pins the oracle: fixed seed, exact math)."""
    scn = os.path.join(util.GOLDEN, "scenes")
    out = str(tmp_path / "c.sum")
    meta = util.run_rtref(["render", "--scene", os.path.join(scn, "cornell.json"), "--data", scn + "/", "--width", 64, "--height", 64,
                           "--passes", 4, "--dims", 128, "--depth", 8, "--strategy", "all", "--dump-sum", out])
    util.assert_bit_identical(util.load_sum(out, 64, 64), util.load_sum(util.golden("cornell_mis.sum"), 64, 64), "cornell golden")
    assert meta["numRays"] == json.load(open(util.golden("cornell_mis.json")))["numRays"]


@pytest.mark.parametrize("scene", ["cornell", "mesh_small"])
def test_hostsim_block_errors_golden(hostsim, tmp_path, scene):
    """The device header's block-error arithmetic on the golden sum / secondary-sum buffers (no live oracle needed)."""
    meta = json.load(open(util.golden(scene + "_mis.json")))
    out = str(tmp_path / "mine.err")
    subprocess.run([HOSTSIM, "blockerrors", util.golden(scene + "_mis.sum"), util.golden(scene + "_mis.sum2"), str(meta["width"]), str(meta["height"]),
                    str(meta["passes"]), util.golden(scene + ".blocks"), out], check=True)
    assert open(out, "rb").read() == open(util.golden(scene + "_mis.blockerr"), "rb").read()


def emission_samples(n=400, seed=2):
    """(position sample x3, direction sample x2) rows: corners of the unit cube first, then random."""
    rng = np.random.default_rng(seed)
    samples = rng.random((n, 5), dtype=np.float32)
    samples[:8] = [[0, 0, 0, 0, 0], [1, 1, 1, 1, 1], [0.5] * 5, [0, 1, 0, 1, 0], [0.25, 0.75, 0.1, 0, 0.5], [0.3, 0.3, 0.34, 0.999, 0.001],
                   [0.1, 0.2, 0.67, 0.5, 0.25], [0.9, 0.8, 0.99, 0.75, 0.5]]
    return np.minimum(samples, np.float32(0.99999994))          # what GenericSampler::GetFloat can return


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("scene", ["c1_sphere_area", "cornell", "materials", "mesh_small", "textures"])
def test_hostsim_light_emission_against_the_live_reference(hostsim, tmp_path, scene):
    """ILight::Emit of every light of the fixture scenes (sphere / rect / box area lights, point, spot, directional,
    constant and environment-mapped background): position, direction, both pdfs, cosAtLight and radiance, bit for bit.
    Building block of the light sub-paths; see csrc/rtb_bidir.h."""
    scn = os.path.join(util.GOLDEN, "scenes")
    sp, refp, minep = str(tmp_path / "samples.f32"), str(tmp_path / "ref.f32"), str(tmp_path / "mine.f32")
    emission_samples().tofile(sp)
    util.run_rtref(["emit", "--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--samples", sp, "--out", refp])
    subprocess.run([HOSTSIM, "emit", util.golden(scene + ".rtb"), sp, minep], check=True)
    want, got = np.fromfile(refp, dtype=np.uint32).reshape(-1, 12), np.fromfile(minep, dtype=np.uint32).reshape(-1, 12)
    assert want.shape == got.shape and len(want) % 400 == 0 and len(want) > 0
    bad = np.argwhere((want != got).any(axis=1)).ravel()
    assert len(bad) == 0, "light %d sample %d: %s vs %s" % (bad[0] // 400, bad[0] % 400, got[bad[0]].view(np.float32), want[bad[0]].view(np.float32))
    assert np.isfinite(want.view(np.float32)).all()


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("scene,w,h", [("cornell", 64, 64), ("materials", 40, 30), ("mesh_small", 1920, 1080), ("cornell_dof", 33, 77)])
def test_hostsim_camera_connection_against_the_live_reference(hostsim, tmp_path, scene, w, h):
    """Camera::WorldToFilm (through the world-to-screen matrix Camera::SetPerspective builds, SSE lane quirks included)
    and Camera::PdfW for points all around the camera: in-front flag, film coordinates and pdf, bit for bit.
    Building block of the camera connection of the light sub-paths (csrc/rtb_bidir.h)."""
    scn = os.path.join(util.GOLDEN, "scenes")
    pp, refp, flatp, minep = (str(tmp_path / n) for n in ("points.f32", "ref.bin", "scene.rtb", "mine.bin"))
    rng = np.random.default_rng(4)
    (rng.normal(size=(600, 3)) * 6).astype(np.float32).tofile(pp)
    common = ["--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", w, "--height", h]
    util.run_rtref(["camera", *common, "--points", pp, "--out", refp])
    util.run_rtref(["flat", *common, "--out", flatp])
    subprocess.run([HOSTSIM, "camera", flatp, str(w), str(h), pp, refp, minep], check=True)
    want, got = np.fromfile(refp, dtype=np.uint32).reshape(-1, 7), np.fromfile(minep, dtype=np.uint32).reshape(-1, 7)
    assert want.shape == got.shape == (600, 7)
    assert 0 < want[:, 0].sum() < 600, "points on both sides of the camera expected"
    assert np.array_equal(want, got), "%d of 600 records differ" % (want != got).any(axis=1).sum()
    inside = want[want[:, 0] == 1][:, 1:3].view(np.float32)
    assert ((inside >= 0) & (inside <= 1)).all(axis=1).any(), "no point inside the frame"


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("scene,passes,mean_tolerance,min_correlation", [("cornell", 200, 0.02, 0.995), ("c1_sphere_area", 200, 0.03, 0.99), ("textures", 300, 0.12, 0.90)])
def test_hostsim_light_tracer_statistics(hostsim, tmp_path, scene, passes, mean_tolerance, min_correlation):
    """LightTracer::RenderPixel (LightTracer.cpp:27-183) assembled from the checked parts in csrc/rtb_bidir.h (emission,
    camera connection, film splat) and the path tracer's BSDF / traversal code.  The reference draws every number of
    this renderer from per-thread generators, so - like its own renderer tests - the comparison is statistical: ray
    count, mean colour and the image at 8 x 8-pixel blocks."""
    w = h = 48
    depth = 6
    scn = os.path.join(util.GOLDEN, "scenes")
    common = ["--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", w, "--height", h]
    refp, flatp, minep = str(tmp_path / "ref.sum"), str(tmp_path / "scene.rtb"), str(tmp_path / "mine.sum")
    meta = util.run_rtref(["render", *common, "--passes", passes, "--dims", 256, "--depth", depth, "--renderer", "Light Tracer", "--dump-sum", refp])
    util.run_rtref(["flat", *common, "--out", flatp])
    r = subprocess.run([HOSTSIM, "lighttrace", flatp, str(w), str(h), str(passes), str(depth), "7", minep], stdout=subprocess.PIPE, text=True, check=True)
    mine = json.loads(r.stdout.strip().splitlines()[-1])
    assert abs(mine["numRays"] - meta["numRays"]) <= 0.003 * meta["numRays"], (mine["numRays"], meta["numRays"])
    a, b = util.load_sum(refp, w, h) / passes, util.load_sum(minep, w, h) / passes
    ma, mb = a.mean(axis=(0, 1)), b.mean(axis=(0, 1))
    assert (np.abs(ma - mb) <= mean_tolerance * ma).all(), (ma, mb)
    blocks = lambda x: x.reshape(6, 8, 6, 8, 3).mean(axis=(1, 3)).ravel()
    assert np.corrcoef(blocks(a), blocks(b))[0, 1] >= min_correlation


def bsdf_direction_pairs(n=1500, seed=6):
    """(local outgoing, local incoming) unit vector pairs: reflection and transmission configurations, arbitrary ones
    (mostly clipped to 0), and a few axis-aligned cases."""
    rng = np.random.default_rng(seed)

    def unit(k):
        d = rng.normal(size=(k, 3))
        return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)

    out, inn = unit(n), unit(n)
    out[:n // 3, 2] = np.abs(out[:n // 3, 2]); inn[:n // 3, 2] = -np.abs(inn[:n // 3, 2])
    a, b = n // 3, n // 3 + n // 8
    out[a:b, 2] = np.abs(out[a:b, 2]); inn[a:b, 2] = np.abs(inn[a:b, 2])
    out[0], inn[0] = [0, 0, 1], [0, 0, -1]
    out[1], inn[1] = [1, 0, 0], [0, 0, -1]
    out[2], inn[2] = [0, 0, 1], [0.6, 0, -0.8]
    return np.hstack([out, inn]).astype(np.float32)


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("scene", ["materials", "cornell", "textures", "mesh_small"])
def test_hostsim_bsdf_pdfs_against_the_live_reference(hostsim, tmp_path, scene):
    """BSDF::Pdf(ctx, ForwardPdf / ReversePdf) of every material (all nine BSDFs, incl. the specular fallback of the rough
    ones) for 1 500 direction pairs: what VCM's vertex connection evaluates (csrc/rtb_bidir.h bsdf_pdf), bit for bit."""
    scn = os.path.join(util.GOLDEN, "scenes")
    dp, refp, minep = str(tmp_path / "dirs.f32"), str(tmp_path / "ref.f32"), str(tmp_path / "mine.f32")
    bsdf_direction_pairs().tofile(dp)
    util.run_rtref(["bsdfpdf", "--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--dirs", dp, "--out", refp])
    subprocess.run([HOSTSIM, "bsdfpdf", util.golden(scene + ".rtb"), dp, minep], check=True)
    want = np.fromfile(refp, dtype=np.float32)
    assert len(want) % 3000 == 0 and (want > 0).sum() > 1000 and np.isfinite(want).all()
    assert want.tobytes() == open(minep, "rb").read()


def shading_points(n=500, seed=8):
    """(position, orthonormal frame, 3 samples) rows for ILight::Illuminate."""
    rng = np.random.default_rng(seed)
    pos = rng.normal(size=(n, 3)) * 3
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    t = np.cross(nrm, rng.normal(size=(n, 3))); t /= np.linalg.norm(t, axis=1, keepdims=True)
    return np.hstack([pos, t, np.cross(nrm, t), nrm, np.minimum(rng.random((n, 3)), 0.99999994)]).astype(np.float32)


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("scene", ["materials", "cornell", "textures", "mesh_small", "c1_sphere_area"])
def test_hostsim_bidirectional_illuminate_against_the_live_reference(hostsim, tmp_path, scene):
    """ILight::Illuminate with rendererSupportsSolidAngleSampling = false (how VCM calls it: area lights are sampled by
    area, every light reports its emission pdf) for 500 shading points x every light: direction, distance, both pdfs,
    cosAtLight and radiance, incl. the early-out cases, bit for bit (csrc/rtb_bidir.h light_illuminate_bidir)."""
    scn = os.path.join(util.GOLDEN, "scenes")
    pp, refp, minep = str(tmp_path / "points.f32"), str(tmp_path / "ref.f32"), str(tmp_path / "mine.f32")
    shading_points().tofile(pp)
    util.run_rtref(["illuminate", "--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--points", pp, "--out", refp])
    subprocess.run([HOSTSIM, "illuminate", util.golden(scene + ".rtb"), pp, minep], check=True)
    want = np.fromfile(refp, dtype=np.float32).reshape(-1, 10)
    assert len(want) % 500 == 0 and (want[:, 7:].sum(axis=1) > 0).sum() > 20
    assert want.tobytes() == open(minep, "rb").read()


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("scene", ["materials", "cornell", "textures", "mesh_small"])
def test_hostsim_bidirectional_radiance_against_the_live_reference(hostsim, tmp_path, scene):
    """ILight::GetRadiance with rendererSupportsSolidAngleSampling = false and both pdf outputs (VCM's EvaluateLight) for
    every light a ray can hit (area, directional, constant / environment-mapped background) x 600 light-space rays with
    cosAtLight on both sides of the cut-off: radiance, directPdfA, emissionPdfW, bit for bit."""
    scn = os.path.join(util.GOLDEN, "scenes")
    rp, refp, minep = str(tmp_path / "rays.f32"), str(tmp_path / "ref.f32"), str(tmp_path / "mine.f32")
    rng = np.random.default_rng(9)
    d = rng.normal(size=(600, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:200, 2] = -np.abs(d[:200, 2]) * 50; d[:200] /= np.linalg.norm(d[:200], axis=1, keepdims=True)     # inside small directional cones too
    cos = rng.uniform(-0.2, 1.0, size=(600, 1)); cos[:5, 0] = [0.0, 1e-6, 9e-7, 1.0, -1.0]
    np.hstack([d, cos]).astype(np.float32).tofile(rp)
    util.run_rtref(["radiance", "--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--rays", rp, "--out", refp])
    subprocess.run([HOSTSIM, "radiance", util.golden(scene + ".rtb"), rp, minep], check=True)
    want = np.fromfile(refp, dtype=np.float32).reshape(-1, 5)
    assert len(want) % 600 == 0 and len(want) > 0 and (want[:, :3].sum(axis=1) > 0).any()
    assert want.tobytes() == open(minep, "rb").read()


@pytest.mark.skipif(not util.oracle_available("exact"), reason="oracle/_ref/exact not built")
@pytest.mark.parametrize("scene,passes,connection,merging,mean_tolerance,min_correlation", [
    ("cornell", 30, 1, 1, 0.02, 0.999),          # VCM as the reference configures it
    ("materials", 30, 1, 1, 0.02, 0.999),        # all nine BSDFs, six lights of five types
    ("textures", 40, 1, 1, 0.03, 0.995),         # textured materials, environment map
    ("caustics", 120, 1, 1, 0.04, 0.99),         # BASELINE config 5 in miniature: rough glass + glass block, caustics
    ("cornell", 30, 1, 0, 0.02, 0.999),          # vertex connection only (bidirectional path tracing)
    ("cornell", 150, 0, 1, 0.06, 0.99),          # vertex merging only (photon mapping; noisy at this photon density)
])
def test_hostsim_vcm_statistics(hostsim, tmp_path, scene, passes, connection, merging, mean_tolerance, min_correlation):
    """The reference's VCM renderer (VertexConnectionAndMerging.cpp) against the CPU model assembled from the checked
    parts (tests/hostsim/hostsim_vcm.h): light sub-paths with dVCM / dVC / dVM bookkeeping, camera connection, next-event
    estimation, vertex connection and vertex merging with the one-pass delay, in the reference's three modes.  The
    reference draws its light paths from per-thread generators, so the comparison is statistical (mean colour and the
    image at 8 x 8-pixel blocks)."""
    w = h = 48
    scn = os.path.join(util.GOLDEN, "scenes")
    common = ["--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", w, "--height", h]
    refp, passp, flatp, minep = (str(tmp_path / n) for n in ("ref.sum", "ref.passes", "scene.rtb", "mine.sum"))
    util.run_rtref(["render", *common, "--passes", passes, "--dims", 256, "--renderer", "VCM", "--vcm-connection", connection, "--vcm-merging", merging,
                    "--dump-sum", refp, "--dump-passes", passp])
    util.run_rtref(["flat", *common, "--out", flatp])
    subprocess.run([HOSTSIM, "vcm", flatp, passp, str(w), str(h), "7", str(connection), str(merging), minep, hostsim], check=True)
    a, b = util.load_sum(refp, w, h) / passes, util.load_sum(minep, w, h) / passes
    ma, mb = a.mean(axis=(0, 1)), b.mean(axis=(0, 1))
    assert np.isfinite(b).all() and (np.abs(ma - mb) <= mean_tolerance * ma).all(), (ma, mb)
    blocks = lambda x: x.reshape(6, 8, 6, 8, 3).mean(axis=(1, 3)).ravel()
    assert np.corrcoef(blocks(a), blocks(b))[0, 1] >= min_correlation


def write_pass_script(path, seeds, offsets):
    with open(path, "wb") as f:
        np.array([len(seeds), seeds.shape[1]], dtype=np.uint32).tofile(f)
        for s, o in zip(seeds, offsets):
            s.astype(np.uint32).tofile(f)
            o.astype(np.float32).tofile(f)


@pytest.mark.parametrize("name,material,fov,sphere,passes,expected,tolerance", [
    # RaytracingTests.cpp:287-516 with its third renderer, "VCM": scene, pass count, expected colour, tolerance
    ("background", None, 90.0, False, 1, [1.0, 2.0, 3.0], 0.01),
    ("diffuse", {"bsdf": "diffuse", "baseColor": [0.4, 0.6, 0.8]}, 10.0, True, 100, [0.4, 1.2, 2.4], 0.05),
    ("emissive", {"bsdf": "null", "baseColor": [0.0, 0.0, 0.0], "emissionColor": [3.0, 2.0, 1.0]}, 10.0, True, 1, [3.0, 2.0, 1.0], 0.0),
    ("metal", {"bsdf": "metal", "baseColor": [0.4, 0.6, 0.8], "IoR": 0.0, "K": 100.0}, 10.0, True, 20, [0.4, 1.2, 2.4], 0.05),
    ("dielectric", {"bsdf": "dielectric", "baseColor": [1.0, 1.0, 1.0]}, 10.0, True, 1000, [1.0, 2.0, 3.0], 0.075),
])
def test_hostsim_vcm_passes_the_reference_renderer_tests(hostsim, tmp_path, name, material, fov, sphere, passes, expected, tolerance):
    """The reference's own renderer tests (furnace, background light) run on "VCM" too (RaytracingTests.cpp:17-22): the
    CPU model of VCM passes them with the reference's scenes, pass counts and per-pixel tolerances."""
    from raytracer_b200 import host_api
    from tests.test_furnace import furnace_scene, validate, SIZE
    sj = furnace_scene(tmp_path, material, fov=fov, sphere=sphere)
    flatp, passp, outp = str(tmp_path / "scene.rtb"), str(tmp_path / "script.passes"), str(tmp_path / "mine.sum")
    scene = host_api.Scene(sj, "")
    try:
        scene.save_flat(flatp, SIZE, SIZE)
    finally:
        scene.close()
    write_pass_script(passp, *host_api.pass_script(passes, SIZE, SIZE, dimensions=64))
    subprocess.run([HOSTSIM, "vcm", flatp, passp, str(SIZE), str(SIZE), "11", "1", "1", outp, hostsim], check=True)
    validate(util.load_sum(outp, SIZE, SIZE) / np.float32(passes), expected, tolerance)


def test_bidirectional_functions_compile_for_the_device(tmp_path):
    """csrc/rtb_bidir.h is written for host and device; the CPU model exercises the host instantiation, this compiles the
    device one (nvcc, sm_100a, the product's arithmetic flags) inside a kernel that is never launched."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--fmad=false", "-ftz=true", "-prec-div=true", "-prec-sqrt=true",
                        "-c", os.path.join(HOSTSIM_DIR, "bidir_device_check.cu"), "-o", str(tmp_path / "check.o")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
