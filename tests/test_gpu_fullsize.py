"""BASELINE.json's full-size configuration (sponza-class scene, 1920x1080, PathTracerMIS depth 8, All lights) on the GPU,
checked through size-independent properties: determinism, batched == pass-by-pass, sharded == unsharded, and the
live reference on a crop-free lower resolution of the same scene (the CPU oracle needs minutes at 1080p)."""
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
