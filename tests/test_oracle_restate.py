"""Pins oracle/restate (the plain-C restatement of the sampler, ray generation, intersectors and two-level traversal)
against the golden vectors produced by the REAL reference (tests/golden/make_golden.py).  No GPU, no product code:
the restatement shares nothing with raytracer_b200/csrc, so agreement of (reference, restatement, CUDA path) on the same
fixtures is a three-way check."""
import json
import os
import subprocess

import numpy as np
import pytest

from raytracer_b200 import capi
from tests import util

RESTATE_DIR = os.path.join(util.ROOT, "oracle", "restate")
RESTATE = os.path.join(util.ROOT, "oracle", "_ref", "restate", "rt_restate")


@pytest.fixture(scope="module")
def restate(tmp_path_factory):
    subprocess.run(["bash", os.path.join(RESTATE_DIR, "build.sh")], check=True)
    bn = tmp_path_factory.mktemp("bn") / "bn.dat"
    capi.load_blue_noise().tofile(str(bn))
    return str(bn)


@pytest.mark.parametrize("scene", ["cornell", "materials", "mesh_small"])
def test_restate_closest_hits(restate, tmp_path, scene):
    out = str(tmp_path / "hits")
    subprocess.run([RESTATE, "trace", util.golden(scene + ".rtb"), util.golden(scene + ".rays"), out], check=True)
    got = np.fromfile(out, dtype=capi.HIT_DTYPE)
    want = np.fromfile(util.golden(scene + ".hits"), dtype=capi.HIT_DTYPE)
    assert len(got) == len(want) > 0
    assert (got["objectId"] == want["objectId"]).all()
    hit = want["objectId"] != 0xFFFFFFFF
    assert (got["subObjectId"][hit] == want["subObjectId"][hit]).all()
    util.assert_bit_identical(got["distance"], want["distance"], "distance")
    mesh = hit & (want["subObjectId"] < 0xFFFFFFF0) & (want["u"] != 0)
    util.assert_bit_identical(got["u"][mesh], want["u"][mesh], "u")
    util.assert_bit_identical(got["v"][mesh], want["v"][mesh], "v")


@pytest.mark.parametrize("scene", ["cornell", "materials", "mesh_small"])
def test_restate_shadow_hits(restate, tmp_path, scene):
    out = str(tmp_path / "hits")
    subprocess.run([RESTATE, "trace", util.golden(scene + ".rtb"), util.golden(scene + ".rays"), out, util.golden(scene + ".tmax")], check=True)
    got = np.fromfile(out, dtype=capi.HIT_DTYPE)
    want = np.fromfile(util.golden(scene + ".shadow_hits"), dtype=capi.HIT_DTYPE)
    assert (got["occluded"] == want["occluded"]).all()
    assert 0 < want["occluded"].sum() < len(want)


@pytest.mark.parametrize("blue", [True, False])
def test_restate_sampler(restate, tmp_path, blue):
    out = str(tmp_path / "s.u32")
    subprocess.run([RESTATE, "sampler", util.golden("cornell.passes"), "24", "20", "40", out, restate if blue else "-"], check=True)
    want = np.fromfile(util.golden("sampler_bn.u32" if blue else "sampler_nobn.u32"), dtype=np.uint32)
    assert (np.fromfile(out, dtype=np.uint32) == want).all()


@pytest.mark.parametrize("scene", ["cornell", "mesh_small"])
def test_restate_primary_rays(restate, tmp_path, scene):
    out = str(tmp_path / "p.rays")
    subprocess.run([RESTATE, "primary", util.golden(scene + ".rtb"), util.golden(scene + ".passes"), "1", "64", "64", out], check=True)
    got = np.fromfile(out, dtype=capi.RAY_DTYPE)
    want = np.fromfile(util.golden(scene + ".primary"), dtype=capi.RAY_DTYPE)
    for f in ("origin", "dir", "invDir", "originDivDir"):
        util.assert_bit_identical(got[f][:, :3], want[f][:, :3], f)


@pytest.mark.parametrize("scene,renderer,mode", [("cornell", "dbgtri", 1), ("cornell", "dbgdepth", 2),
                                                 ("mesh_small", "dbgtri", 1), ("mesh_small", "dbgdepth", 2)])
def test_restate_debug_renderer_sum(restate, tmp_path, scene, renderer, mode):
    """Whole-frame check: raygen + traversal + DebugRenderer colour + film accumulation over all passes."""
    meta = json.load(open(util.golden("%s_%s.json" % (scene, renderer))))
    out = str(tmp_path / "d.sum")
    subprocess.run([RESTATE, "debug", util.golden(scene + ".rtb"), util.golden(scene + ".passes"), str(mode),
                    str(meta["width"]), str(meta["height"]), out], check=True)
    want = util.load_sum(util.golden("%s_%s.sum" % (scene, renderer)), meta["width"], meta["height"])
    util.assert_bit_identical(util.load_sum(out, meta["width"], meta["height"]), want, "%s/%s" % (scene, renderer))


def test_restate_texture_evaluation(restate, tmp_path):
    """ITexture::Evaluate (bitmap with smooth-step bilinear filtering, checkerboard, simplex noise octaves, nested mix)
    restated in plain C, at the 1 100 adversarial coordinates of textures.uv, against the reference's own Evaluate."""
    out = str(tmp_path / "t.f32")
    subprocess.run([RESTATE, "texeval", util.golden("textures.rtb"), util.golden("textures.uv"), out], check=True)
    got = np.fromfile(out, dtype=np.uint32)
    want = np.fromfile(util.golden("textures.texeval"), dtype=np.uint32)
    assert got.shape == want.shape and got.size > 0
    bad = np.argwhere(got != want)
    assert len(bad) == 0, "%d of %d floats differ (first: texture %d, point %d)" % (len(bad), got.size, bad[0][0] // 4400, (bad[0][0] // 4) % 1100)


@pytest.mark.parametrize("scene", ["c1_sphere_area", "cornell", "cornell_dof", "materials", "mesh_small", "textures"])
def test_restate_front_buffer(restate, tmp_path, scene):
    """Viewport::PostProcessTile with the default PostprocessParams (saturation, FastLog / FastExp contrast, ACES curve,
    sRGB approximation, ToBGR) restated in plain C: the reference's 8-bit front buffer of the golden renders."""
    meta = json.load(open(util.golden(scene + "_mis.json")))
    out = str(tmp_path / "front.bgra")
    subprocess.run([RESTATE, "front", util.golden(scene + "_mis.sum"), str(meta["width"]), str(meta["height"]), str(meta["passes"]), out], check=True)
    assert open(out, "rb").read() == open(util.golden(scene + "_mis.front"), "rb").read()


@pytest.mark.parametrize("scene", ["cornell", "mesh_small"])
def test_restate_block_errors(restate, tmp_path, scene):
    """Viewport::ComputeBlockError (row sums left to right, rows top to bottom) for the whole image, single pixels,
    rows, columns and random rectangles."""
    meta = json.load(open(util.golden(scene + "_mis.json")))
    out = str(tmp_path / "block.err")
    subprocess.run([RESTATE, "blockerr", util.golden(scene + "_mis.sum"), util.golden(scene + "_mis.sum2"), str(meta["width"]), str(meta["height"]),
                    str(meta["passes"]), util.golden(scene + ".blocks"), out], check=True)
    want = np.fromfile(util.golden(scene + "_mis.blockerr"), dtype=np.float32)
    assert len(want) == 40 and np.isfinite(want).all() and (want[:1] > 0).all()
    assert open(out, "rb").read() == want.tobytes()


@pytest.mark.parametrize("scene,dims", [("c1_sphere_area", 128), ("cornell", 128), ("materials", 256), ("mesh_small", 128), ("textures", 192)])
def test_restate_pass_script(restate, tmp_path, scene, dims):
    """The host prologue of Viewport::Render (entropy stream -> Random -> scrambled Halton; the 4-wide generator, FastLog
    and Sin behind the Box-Muller anti-aliasing offset) restated in plain C: the committed pass scripts, byte for byte."""
    out = str(tmp_path / "script.passes")
    subprocess.run([RESTATE, "passes", "12345", str(dims), "64", "64", "4", out], check=True)
    assert open(out, "rb").read() == open(util.golden(scene + ".passes"), "rb").read()
