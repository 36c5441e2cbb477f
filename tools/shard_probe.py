"""Development probe: what a 1/N shard of the 1080p bench frame costs on ONE GPU (the per-rank work of an N-GPU run),
with and without the CUDA-graph replay of the launch chain.  Usage: python tools/shard_probe.py [numShards] [steps]"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import bench
from raytracer_b200 import host_api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
scene = host_api.Scene(bench.ensure_scene(), bench.SCENE_DIR + "/")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
vp = host_api.Viewport(scene, device=0, num_shards=n, shard_index=0, stream=stream.cuda_stream)
vp.set_renderer("Path Tracer MIS"); vp.set_params(dimensions=bench.DIMS, max_ray_depth=bench.MAX_DEPTH, strategy="all", seed=12345)
vp.resize(bench.WIDTH, bench.HEIGHT)
vp.render(5); vp.synchronize(); vp.reset()
best = 1e9
for rep in range(5):
    vp.reset(); vp.synchronize(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); vp.render(steps); vp.context(); 
    from raytracer_b200 import capi
    capi.load_library().rtb_join(vp.context()); e1.record(); vp.synchronize(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
if os.environ.get("SHARD_PROBE_PROFILE"):      # per-depth breakdown of one wavefront (set RTB_PROFILE_DEPTHS=1 as well)
    lib = capi.load_library()
    vp.reset(); lib.rtb_set_profiling(vp.context(), 1); vp.render(steps); vp.synchronize(); lib.rtb_set_profiling(vp.context(), 0)
c = vp.counters()
print("shard 0 of %d, %d steps: best %.3f ms (%.3f ms per step), graph=%s" % (n, steps, best, best / steps, os.environ.get("RTB_GRAPH", "1")))
vp.close(); scene.close()
