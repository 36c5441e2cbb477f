"""Development probe: the kernels OUTSIDE the pass loop at the bench size (1920 x 1080) - front buffer without and with bloom,
block errors of a 16 x 16-pixel block grid - for `ncu --metrics gpu__time_duration.sum -k regex:"k_front|k_blur|k_.*_error|k_unshard"`.
Usage: python tools/aux_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from raytracer_b200 import capi

w, h, passes = 1920, 1080, 4
sf = capi.SceneFile(os.path.join(ROOT, "bench_data", "sponza_class.rtb"))
rng = np.random.default_rng(1)
with capi.Context(device=0) as ctx:
    ctx.upload_scene(sf); ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    ctx.render_passes(cam, rng.integers(0, 2**32, (passes, 128), dtype=np.uint32), rng.normal(size=(passes, 2)) * 0.5, 0)
    ctx.synchronize()
    for bloom in (0.0, 0.1):
        t0 = time.perf_counter(); img = ctx.read_front(capi.postprocess_desc(passes, bloom=bloom)); dt = time.perf_counter() - t0
        print("read_front bloom=%.1f: %.2f ms wall (kernels + 8.3 MB read-back), mean byte %.1f" % (bloom, dt * 1e3, float((img & 0xFF).mean())))
    bs = 16
    blocks = np.array([(x, min(x + bs, w), y, min(y + bs, h)) for y in range(0, h, bs) for x in range(0, w, bs)], dtype=np.uint32)
    t0 = time.perf_counter(); err = ctx.block_errors(passes, blocks); dt = time.perf_counter() - t0
    print("block_errors of %d blocks: %.2f ms wall, mean %.4g" % (len(blocks), dt * 1e3, float(err.mean())))
