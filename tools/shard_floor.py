#!/usr/bin/env python3
"""Development probe: how long does ONE pass of a 1/N shard take when passes are rendered one at a time (the e2e
pattern at N GPUs), against the same passes batched?  Usage: python tools/shard_floor.py [num_shards]"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench
from raytracer_b200 import host_api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scene = host_api.Scene(bench.ensure_scene(), bench.SCENE_DIR + "/")
vp = host_api.Viewport(scene, device=0, num_shards=n, shard_index=0)
vp.set_renderer("Path Tracer MIS")
vp.set_params(dimensions=bench.DIMS, max_ray_depth=bench.MAX_DEPTH, strategy="all", seed=12345)
vp.resize(bench.WIDTH, bench.HEIGHT)
vp.render(8); vp.synchronize()
for label, per_call, sync_each in (("batched (64 passes, one call)", 64, False), ("one pass per call, no sync", 1, False), ("one pass per call, sync after each", 1, True)):
    vp.reset(); vp.synchronize()
    t0 = time.perf_counter()
    done = 0
    while done < 64:
        vp.render(per_call); done += per_call
        if sync_each: vp.synchronize()
    vp.synchronize()
    dt = time.perf_counter() - t0
    c = vp.counters()
    print("shard 1/%d  %-36s %.3f ms per pass, %.1f Mray/s (this shard), launches %d" % (n, label, dt / 64 * 1e3, (c["numRays"] + c["numShadowRays"]) / dt * 1e-6, c["numKernelLaunches"]))
