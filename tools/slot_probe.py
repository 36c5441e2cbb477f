import sys, numpy as np
sys.path.insert(0, "/root/repo")
from raytracer_b200 import capi
from tests import util
import json
scene = sys.argv[1] if len(sys.argv) > 1 else "mesh_small"
RID, DBG = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 0)
meta = json.load(open(util.golden("%s_mis.json" % scene))) if not scene.endswith(".rtb") else {"dims": 128}
w, h = 256, 256
sf = capi.SceneFile(scene if scene.endswith(".rtb") else util.golden(scene + ".rtb"))
rng = np.random.default_rng(5)
seeds = rng.integers(0, 2**32, (8, meta["dims"]), dtype=np.uint32); offs = (rng.normal(size=(8, 2)) * 0.5).astype(np.float32)
def run(single):
    with capi.Context(device=0) as ctx:
        ctx.upload_scene(sf); ctx.resize(w, h)
        cam = capi.camera_for(sf, w, h)
        if single:
            for p in range(8): ctx.render_pass(ctx.make_pass(cam, seeds[p], offs[p], p, renderer=RID, debug_mode=DBG))
        else:
            ctx.render_passes(cam, seeds, offs, 0, renderer=RID, debug_mode=DBG)
        return ctx.read_sum(), ctx.counters()
a, ca = run(False)
for k in range(3):
    b, cb = run(True)
    print("bad floats", int((a.view(np.uint32) != b.view(np.uint32)).sum()), ca["numRays"] == cb["numRays"], ca["numShadowRays"] == cb["numShadowRays"])
bad = np.argwhere((a.view(np.uint32) != b.view(np.uint32)).any(axis=2))
print("bad pixels", len(bad), bad[:40].tolist())
if len(bad):
    print("a", a[tuple(bad[0])], "b", b[tuple(bad[0])], "diff", (b - a)[tuple(bad[0])])
    ys, xs = bad[:, 0], bad[:, 1]
    print("rows", np.unique(ys)[:30].tolist(), "x blocks(8)", np.unique(xs // 8)[:30].tolist())
