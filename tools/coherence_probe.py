#!/usr/bin/env python3
"""Experiment: how much does ray ordering matter for k_trace_batch on the sponza-class scene?
Primary rays (ordered / shuffled) and synthetic bounce rays (origin = primary hit point, random direction) in
natural pixel order, shuffled, and sorted by (direction octant, Morton code of the origin)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from raytracer_b200 import capi


def make_rays(o, d):
    n = len(o)
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    f = np.float32
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + (d[:, 2] * d[:, 2]), dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (f(1) / d).astype(np.float32)
        nd = (d / ln[:, None]).astype(np.float32)
        inv_dir = (ln[:, None] * inv).astype(np.float32)
    rays["origin"][:, :3] = o; rays["dir"][:, :3] = nd; rays["invDir"][:, :3] = inv_dir; rays["originDivDir"][:, :3] = o * inv_dir
    return rays


def morton(q):
    def spread(v):
        v = v.astype(np.uint64)
        v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; v = (v | (v << 2)) & 0x09249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def run(ctx, rays, label, mode=0):
    t = torch.from_numpy(rays.view(np.uint8).reshape(len(rays), 64)).cuda()
    hits = torch.zeros(len(rays), 24, dtype=torch.uint8, device="cuda")
    if mode == 1:
        h = np.zeros(len(rays), dtype=capi.HIT_DTYPE); h["distance"] = 3.0e38
        hits = torch.from_numpy(h.view(np.uint8).reshape(len(rays), 24)).cuda()
    ms = []
    for _ in range(4):
        ms.append(ctx.trace_batch_device(t.data_ptr(), len(rays), hits.data_ptr(), mode))
    best = min(ms[1:])
    print("%-46s %8.3f ms  %7.1f Mray/s" % (label, best, len(rays) / best * 1e-3))
    return hits.cpu().numpy().view(capi.HIT_DTYPE).reshape(-1)


def main():
    sf = capi.SceneFile(os.path.join(ROOT, "bench_data", "sponza_class.rtb"))
    rng = np.random.default_rng(3)
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf)
        ctx.resize(1920, 1080)
        cam = capi.camera_for(sf, 1920, 1080)
        prim = ctx.primary_rays(ctx.make_pass(cam, rng.integers(0, 2**32, 128, dtype=np.uint32), (0.1, 0.2), 0))
        h = run(ctx, prim, "primary, raster order")
        perm = rng.permutation(len(prim))
        run(ctx, prim[perm], "primary, shuffled")
        hit = h["objectId"] != 0xFFFFFFFF
        o = (prim["origin"][hit, :3] + prim["dir"][hit, :3] * (h["distance"][hit, None] - 1e-3)).astype(np.float32)
        d = rng.normal(size=o.shape).astype(np.float32)
        b = make_rays(o, d)
        for mode, name in ((0, "closest"), (1, "shadow(tmax=inf)")):
            run(ctx, b, "bounce %s, pixel order" % name, mode)
            run(ctx, b[rng.permutation(len(b))], "bounce %s, shuffled" % name, mode)
            octant = ((d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2))
            lo, hi = o.min(0), o.max(0)
            q = np.clip(((o - lo) / (hi - lo + 1e-6) * 1024).astype(np.int64), 0, 1023)
            key = (octant << 30) | morton(q)
            run(ctx, b[np.argsort(key, kind="stable")], "bounce %s, sorted octant+morton30" % name, mode)
            q5 = np.clip(((o - lo) / (hi - lo + 1e-6) * 32).astype(np.int64), 0, 31)
            key = (morton(q5) << 3) | octant
            run(ctx, b[np.argsort(key, kind="stable")], "bounce %s, sorted morton15 then octant" % name, mode)
            # direction-major: 6x? quantized direction (8x8x8) then origin
            dq = np.clip(((d / np.linalg.norm(d, axis=1, keepdims=True) * 0.5 + 0.5) * 8).astype(np.int64), 0, 7)
            key = (morton(dq) << 30) | morton(q)
            run(ctx, b[np.argsort(key, kind="stable")], "bounce %s, sorted dir9 then morton30" % name, mode)


if __name__ == "__main__":
    main()
