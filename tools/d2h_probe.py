#!/usr/bin/env python3
"""Development probe: D2H rate of rtb_read_sum into pageable vs page-locked host frames."""
import os, sys, time
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from raytracer_b200 import capi
from tests import util
sf = capi.SceneFile(util.golden("cornell.rtb"))
with capi.Context(device=0) as ctx:
    ctx.upload_scene(sf)
    ctx.resize(1920, 1080)
    pin = capi.PinnedImage(1920, 1080)
    page = np.empty((1080, 1920, 3), np.float32)
    for name, arr in (("pageable", page), ("pinned", pin.array)):
        ctx.read_sum(out=arr)
        t0 = time.perf_counter()
        for _ in range(20):
            ctx.read_sum(out=arr)
        dt = (time.perf_counter() - t0) / 20
        print("%s: %.3f ms per 24.9 MB frame = %.1f GB/s" % (name, dt * 1e3, 24.8832e-3 / dt))
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.read_sum_async(pin)
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("async pinned: %.3f ms per frame = %.1f GB/s" % (dt * 1e3, 24.8832e-3 / dt))
    pin.close()
import torch
x = torch.empty(24883200 // 4, device="cuda"); h = torch.empty(24883200 // 4, pin_memory=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): h.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print("torch pinned D2H: %.3f ms = %.1f GB/s" % (dt * 1e3, 24.8832e-3 / dt))
