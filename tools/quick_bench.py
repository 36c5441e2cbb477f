#!/usr/bin/env python3
"""Development probe (not the contract bench): render the sponza-class scene from a prebuilt RTBS file and print
per-kernel timings.  Usage: python tools/quick_bench.py [width height passes]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from raytracer_b200 import capi  # noqa: E402
from tests import util  # noqa: E402


def main():
    w = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
    h = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
    passes = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    flags = capi.RTB_FLAG_COUNT_TESTS if os.environ.get("RTB_COUNT") else 0
    sf = capi.SceneFile(os.path.join(ROOT, "bench_data", "sponza_class.rtb"))
    rng = np.random.default_rng(1)
    with capi.Context(device=0, flags=flags) as ctx:
        ctx.upload_scene(sf)
        # parity at 480x270 against the oracle dump, if present
        sump = os.path.join(ROOT, "bench_data", "sponza_480.sum")
        if os.path.exists(sump):
            ps = capi.PassScript(os.path.join(ROOT, "bench_data", "sponza_480.passes"))
            ctx.resize(480, 270)
            cam = capi.camera_for(sf, 480, 270)
            for p in range(ps.num_passes):
                ctx.render_pass(ctx.make_pass(cam, ps.seeds[p], ps.offsets[p], p))
            got = ctx.read_sum()
            want = util.load_sum(sump, 480, 270)
            same = (got.view(np.uint32) == want.view(np.uint32)).all(axis=2)
            print("parity 480x270x%d vs oracle: %.4f%% identical pixels, counters %s" % (ps.num_passes, 100.0 * same.mean(), ctx.counters()))
        ctx.resize(w, h)
        cam = capi.camera_for(sf, w, h)
        ctx.set_profiling(True)
        agg = {}
        nb = int(os.environ.get("QB_BATCH", "8"))
        for p in range(2):
            ctx.render_passes(cam, rng.integers(0, 2**32, (nb, 128), dtype=np.uint32), rng.normal(size=(nb, 2)) * 0.5, p * nb)
            t = ctx.pass_timings()
            print("profiled %d passes: %s" % (nb, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in t.items()}))
        ctx.set_profiling(False)
        ctx.reset()
        ctx.synchronize()
        t0 = time.perf_counter()
        ctx.render_passes(cam, rng.integers(0, 2**32, (passes, 128), dtype=np.uint32), rng.normal(size=(passes, 2)) * 0.5, 0)
        ctx.synchronize()
        dt = time.perf_counter() - t0
        c = ctx.counters()
        rays = c["numRays"] + c["numShadowRays"]
        print("%dx%d x %d passes: %.3f s, %.1f Mray/s (path+shadow), %.2f Msamples/s, rays/sample %.2f + %.2f" % (
            w, h, passes, dt, rays / dt * 1e-6, c["numPrimaryRays"] / dt * 1e-6, c["numRays"] / c["numPrimaryRays"], c["numShadowRays"] / c["numPrimaryRays"]))
        if flags:
            print("box tests/ray %.2f tri tests/ray %.2f" % (c["numRayBoxTests"] / rays, c["numRayTriangleTests"] / rays))


if __name__ == "__main__":
    main()
