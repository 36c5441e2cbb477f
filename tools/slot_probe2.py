"""Overlap probe, part 2: render 8 single-pass wavefronts of a golden scene at 256 x 256 and save the
film; run once per configuration (RTB_SLOTS, RTB_TAIL_WITH_SLOTS, RTB200_LIB ... from the environment) and compare the files.
   python tools/slot_probe2.py <scene> <out.npy>            python tools/slot_probe2.py --cmp a.npy b.npy"""
import sys, json, numpy as np
sys.path.insert(0, "/root/repo")
if sys.argv[1] == "--cmp":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = a.view(np.uint32) != b.view(np.uint32)
    px = np.argwhere(bad.any(axis=2))
    print("%s vs %s: %d bad floats, %d bad pixels%s" % (sys.argv[2], sys.argv[3], int(bad.sum()), len(px),
          (", max rel %.3g" % float((np.abs(a - b) / np.maximum(np.abs(a), 1e-6))[bad].max())) if bad.any() else ""))
    sys.exit(0)
from raytracer_b200 import capi
from tests import util
scene, out = sys.argv[1], sys.argv[2]
meta = json.load(open(util.golden("%s_mis.json" % scene)))
w = h = 256
sf = capi.SceneFile(util.golden(scene + ".rtb"))
rng = np.random.default_rng(5)
seeds = rng.integers(0, 2**32, (8, meta["dims"]), dtype=np.uint32); offs = (rng.normal(size=(8, 2)) * 0.5).astype(np.float32)
with capi.Context(device=0) as ctx:
    ctx.upload_scene(sf); ctx.resize(w, h)
    cam = capi.camera_for(sf, w, h)
    for rep in range(3):
        ctx.reset()
        for p in range(8): ctx.render_pass(ctx.make_pass(cam, seeds[p], offs[p], p, renderer=1, debug_mode=0))
        img = ctx.read_sum()
        np.save(out.replace(".npy", "_%d.npy" % rep), img)
    print(out, ctx.counters())
