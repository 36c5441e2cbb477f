"""The "VCM" renderer on the device (csrc/rtb_vcm.cuh; reference: Rendering/VertexConnectionAndMerging.cpp) through the
C ABI, against the reference itself run live on this box's CPU.

The reference draws its light sub-paths (and the film jitter of their splats) from per-thread generators, so even the
reference does not reproduce its own VCM image from run to run: the bar here is statistical, as in the CPU model's tests
(tests/test_hostsim_parity.py::test_hostsim_vcm_statistics) - mean colour per channel and the image at 8 x 8-pixel
blocks - plus the reference's own renderer tests on "VCM" (tests/test_furnace.py, RaytracingTests.cpp:287-516)."""
import os

import numpy as np
import pytest

from raytracer_b200 import capi
from tests import util

pytestmark = pytest.mark.gpu


def render_vcm(flatp, passp, w, h, passes, **vcm):
    sf = capi.SceneFile(flatp)
    ps = capi.PassScript(passp)
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf)
        ctx.resize(w, h)
        ctx.set_vcm_params(**vcm)
        ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[:passes], ps.offsets[:passes], 0, renderer=capi.RTB_RENDERER_VCM)
        return ctx.read_sum(), ctx.read_sum(1), ctx.counters()


def reference_vcm(tmp_path, scene, w, h, passes, connection, merging, dims=256):
    scn = os.path.join(util.GOLDEN, "scenes")
    common = ["--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", w, "--height", h]
    refp, ref2p, passp, flatp = (str(tmp_path / n) for n in ("ref.sum", "ref.sum2", "ref.passes", "scene.rtb"))
    meta = util.run_rtref(["render", *common, "--passes", passes, "--dims", dims, "--renderer", "VCM", "--vcm-connection", connection, "--vcm-merging", merging,
                           "--dump-sum", refp, "--dump-sum2", ref2p, "--dump-passes", passp])
    util.run_rtref(["flat", *common, "--out", flatp])
    return util.load_sum(refp, w, h), util.load_sum(ref2p, w, h), passp, flatp, meta


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("scene,passes,connection,merging,mean_tolerance,min_correlation", [
    # (tolerances: the camera paths of both sides draw the SAME pass script, which takes most of the noise out of the
    #  difference - over 8 reference seeds the CPU model's mean colour stays within 0.5 % of the reference's in the 2 % rows)
    ("cornell", 30, 1, 1, 0.02, 0.999),          # VCM as the reference configures it
    ("materials", 30, 1, 1, 0.02, 0.999),        # all nine BSDFs, six lights of five types
    ("textures", 40, 1, 1, 0.03, 0.995),         # textured materials, environment map
    ("caustics", 120, 1, 1, 0.04, 0.99),         # BASELINE config 5 in miniature: rough glass + glass block, caustics
    ("cornell", 30, 1, 0, 0.02, 0.999),          # vertex connection only (bidirectional path tracing)
    ("cornell", 150, 0, 1, 0.06, 0.99),          # vertex merging only (photon mapping; noisy at this photon density)
])
def test_vcm_statistics_against_the_live_reference(tmp_path, scene, passes, connection, merging, mean_tolerance, min_correlation):
    w = h = 48
    ref, ref2, passp, flatp, meta = reference_vcm(tmp_path, scene, w, h, passes, connection, merging)
    got, got2, counters = render_vcm(flatp, passp, w, h, passes, connection=bool(connection), merging=bool(merging))
    a, b = ref / passes, got / passes
    ma, mb = a.mean(axis=(0, 1)), b.mean(axis=(0, 1))
    assert np.isfinite(b).all() and (b >= 0).all()
    assert (np.abs(ma - mb) <= mean_tolerance * ma).all(), (ma, mb)
    blocks = lambda x: x.reshape(6, 8, 6, 8, 3).mean(axis=(1, 3)).ravel()
    assert np.corrcoef(blocks(a), blocks(b))[0, 1] >= min_correlation
    # the secondary sum holds the even passes (camera paths and splats alike)
    m2a, m2b = ref2.mean(axis=(0, 1)), got2.mean(axis=(0, 1))
    assert (np.abs(m2a - m2b) <= 2.0 * mean_tolerance * m2a).all(), (m2a, m2b)
    assert counters["numPrimaryRays"] == w * h * passes
    # (the reference's VCM renderer does not touch the ray counters - its meta reports 0 - so only sanity here: at least one
    #  light and one camera segment per pixel and pass, shadow rays only with vertex connection)
    assert counters["numRays"] >= w * h * passes
    assert (counters["numShadowRays"] > 0) == bool(connection) and counters["numShadowRaysHit"] <= counters["numShadowRays"]


def test_vcm_is_reproducible_and_needs_the_previous_pass(tmp_path):
    """Same seeds, same image (the light paths draw from a counter-seeded generator; only the order of the float atomics
    of the splats can differ); a merging pass whose predecessor was not rendered is refused."""
    scn = os.path.join(util.GOLDEN, "scenes")
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    w = h = 64
    imgs = []
    for _ in range(2):
        with capi.Context(device=0) as ctx:
            ctx.upload_scene(sf); ctx.resize(w, h)
            ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[:4], ps.offsets[:4], 0, renderer=capi.RTB_RENDERER_VCM)
            imgs.append(ctx.read_sum())
    assert np.allclose(imgs[0], imgs[1], rtol=1e-4, atol=1e-5)
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf); ctx.resize(w, h)
        with pytest.raises(capi.RtbError):
            ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[2:3], ps.offsets[2:3], 2, renderer=capi.RTB_RENDERER_VCM)
        # ... but vertex connection alone has no such dependency
        ctx.set_vcm_params(connection=True, merging=False)
        ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[2:3], ps.offsets[2:3], 2, renderer=capi.RTB_RENDERER_VCM)
        assert np.isfinite(ctx.read_sum()).all()


def test_vcm_sharded_connection_only_matches_unsharded(tmp_path):
    """Two shards, vertex connection only: each shard's film plus the splats of both shards' light paths (summed like the
    all-reduce a multi-GPU caller performs) gives the single-context image up to the order of the float additions."""
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    w = h = 64
    n = 4
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf); ctx.resize(w, h)
        ctx.set_vcm_params(connection=True, merging=False)
        ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[:n], ps.offsets[:n], 0, renderer=capi.RTB_RENDERER_VCM)
        whole = ctx.read_sum()
    import torch
    ctxs = [capi.Context(device=0, num_shards=2, shard_index=s) for s in range(2)]
    try:
        bufs = []
        for c in ctxs:
            c.upload_scene(sf); c.resize(w, h)
            c.set_vcm_params(connection=True, merging=False)
            c.render_passes(capi.camera_for(sf, w, h), ps.seeds[:n], ps.offsets[:n], 0, renderer=capi.RTB_RENDERER_VCM)
            c.synchronize()
            bufs.append(c.vcm_splat_buffers())
        total = [bufs[0][k] + bufs[1][k] for k in range(2)]
        acc = np.zeros_like(whole)
        for c, b in zip(ctxs, bufs):
            b[0].copy_(total[0]); b[1].copy_(total[1])
            torch.cuda.synchronize()
            c.vcm_fold_splats()
            acc += c.read_sum()
        assert np.allclose(acc, whole, rtol=2e-4, atol=1e-5)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.skipif(not util.oracle_available("exact"), reason="prebuilt oracle/_ref/exact did not travel to this box")
@pytest.mark.parametrize("scene,passes,mean_tolerance,min_correlation", [
    ("cornell", 200, 0.03, 0.995),
    # point / spot lights and specular surfaces, and no shared camera samples to correlate the two sides: at 200 passes two
    # runs of the REFERENCE with different seeds differ by 3 % (one standard deviation) in the mean colour and correlate at
    # 0.87 (a 200-pass version of this row failed on one box in three) - hence 3 200 passes, 3.7 standard deviations
    ("materials", 3200, 0.04, 0.8),
])
def test_light_tracer_statistics_against_the_live_reference(tmp_path, scene, passes, mean_tolerance, min_correlation):
    """RTB_RENDERER_LIGHT_TRACER (k_light_tracer; reference: LightTracer::RenderPixel, LightTracer.cpp:26-175): one light
    sub-path per pixel and pass, its vertices connected to the camera and splatted.  Statistical bar, as for VCM (the
    reference draws from per-thread generators): mean colour per channel, 8 x 8-block image, secondary sum, and the number
    of path rays within the sampling noise of the reference's."""
    w = h = 48
    scn = os.path.join(util.GOLDEN, "scenes")
    common = ["--scene", os.path.join(scn, scene + ".json"), "--data", scn + "/", "--width", w, "--height", h]
    refp, ref2p, passp, flatp = (str(tmp_path / n) for n in ("ref.sum", "ref.sum2", "ref.passes", "scene.rtb"))
    meta = util.run_rtref(["render", *common, "--passes", passes, "--dims", 256, "--depth", 8, "--renderer", "Light Tracer",
                           "--dump-sum", refp, "--dump-sum2", ref2p, "--dump-passes", passp])
    util.run_rtref(["flat", *common, "--out", flatp])
    ref, ref2 = util.load_sum(refp, w, h), util.load_sum(ref2p, w, h)
    sf = capi.SceneFile(flatp)
    ps = capi.PassScript(passp)
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf)
        ctx.resize(w, h)
        ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[:passes], ps.offsets[:passes], 0, renderer=capi.RTB_RENDERER_LIGHT_TRACER, max_depth=8)
        got, got2, counters = ctx.read_sum(), ctx.read_sum(1), ctx.counters()
    a, b = ref / passes, got / passes
    assert np.isfinite(b).all() and (b >= 0).all() and b.mean() > 0
    ma, mb = a.mean(axis=(0, 1)), b.mean(axis=(0, 1))
    assert (np.abs(ma - mb) <= mean_tolerance * ma).all(), (ma, mb)
    blocks = lambda x: x.reshape(6, 8, 6, 8, 3).mean(axis=(1, 3)).ravel()
    assert np.corrcoef(blocks(a), blocks(b))[0, 1] >= min_correlation
    m2a, m2b = ref2.mean(axis=(0, 1)), got2.mean(axis=(0, 1))
    assert (np.abs(m2a - m2b) <= 2.0 * mean_tolerance * m2a).all(), (m2a, m2b)
    assert counters["numPrimaryRays"] == w * h * passes
    assert counters["numRays"] >= w * h * passes * 0.5 and counters["numShadowRaysHit"] <= counters["numShadowRays"]


def test_light_tracer_is_reproducible_and_shards_add_up():
    """Same seeds, same image (up to the order of the float atomics); the splats of two shards' light paths, summed like the
    all-reduce of a multi-GPU caller and folded, give the unsharded image."""
    import torch
    sf = capi.SceneFile(util.golden("cornell.rtb"))
    ps = capi.PassScript(util.golden("cornell.passes"))
    w = h = 64
    n = 4
    imgs = []
    for _ in range(2):
        with capi.Context(device=0) as ctx:
            ctx.upload_scene(sf); ctx.resize(w, h)
            ctx.render_passes(capi.camera_for(sf, w, h), ps.seeds[:n], ps.offsets[:n], 0, renderer=capi.RTB_RENDERER_LIGHT_TRACER)
            imgs.append(ctx.read_sum())
    whole = imgs[0]
    assert np.allclose(imgs[0], imgs[1], rtol=1e-4, atol=1e-5) and whole.mean() > 0
    ctxs = [capi.Context(device=0, num_shards=2, shard_index=s) for s in range(2)]
    try:
        bufs = []
        for c in ctxs:
            c.upload_scene(sf); c.resize(w, h)
            c.render_passes(capi.camera_for(sf, w, h), ps.seeds[:n], ps.offsets[:n], 0, renderer=capi.RTB_RENDERER_LIGHT_TRACER)
            c.synchronize()
            bufs.append(c.vcm_splat_buffers())
        total = [bufs[0][k] + bufs[1][k] for k in range(2)]
        acc = np.zeros_like(whole)
        for c, b in zip(ctxs, bufs):
            b[0].copy_(total[0]); b[1].copy_(total[1])
            torch.cuda.synchronize()
            c.vcm_fold_splats()
            acc += c.read_sum()
        assert np.allclose(acc, whole, rtol=2e-4, atol=1e-5)
    finally:
        for c in ctxs:
            c.close()
