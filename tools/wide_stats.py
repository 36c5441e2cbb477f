"""CPU study of the any-hit tree (rtb_wide.h) on the bench scene: the 8-wide walk must return the binary walk's
booleans for every ray; prints node visits / triangle tests per ray for both (hostsim = the device headers built with g++)."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench
from raytracer_b200 import host_api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
tmp = "/tmp/wide_stats"; os.makedirs(tmp, exist_ok=True)
flat = os.path.join(tmp, "bench.rtb")
if not os.path.exists(flat):
    s = host_api.Scene(bench.ensure_scene(), bench.SCENE_DIR + "/"); s.save_flat(flat, 1920, 1080); s.close()
rng = np.random.default_rng(7)
o = np.zeros((n, 4), np.float32); d = np.zeros((n, 4), np.float32)
o[:, :3] = rng.uniform([-14, 0.05, -6], [14, 11, 6], (n, 3))
v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
# half of the rays point to the upper hemisphere (sky / sun samples), like the bench's shadow rays
v[: n // 2, 1] = np.abs(v[: n // 2, 1])
# WIDE_STATS_NEAR_AXIS=f: that fraction of the rays gets one direction component shrunk to ~1e-5 (|invDir| ~ 1e5 on that axis:
# the case that made a shared per-ray slack pass every box, see rtb_wide.h wide_ray_eps)
frac = float(os.environ.get("WIDE_STATS_NEAR_AXIS", "0"))
if frac > 0:
    k = int(n * frac)
    ax = rng.integers(0, 3, k)
    v[np.arange(k), ax] *= 1e-5
    v[:k] /= np.linalg.norm(v[:k], axis=1, keepdims=True)
d[:, :3] = v
with np.errstate(divide="ignore"):
    inv = np.zeros((n, 4), np.float32); inv[:, :3] = (np.float32(1.0) / d[:, :3]).astype(np.float32)
odd = (o * inv).astype(np.float32)
rays = np.concatenate([o, d, inv, odd], axis=1).astype(np.float32)
rays.tofile(os.path.join(tmp, "rays.bin"))
tmax = np.where(rng.uniform(size=n) < 0.5, np.float32(3.0e38), rng.uniform(0.2, 40.0, n)).astype(np.float32)
tmax.tofile(os.path.join(tmp, "tmax.bin"))
hs = os.path.join(ROOT, "tests", "hostsim", "hostsim")
out = {}
for name, env in (("binary", {"HOSTSIM_NO_WIDE": "1"}), ("wide", {})):
    e = dict(os.environ, HOSTSIM_COUNT="1", **env)
    r = subprocess.run([hs, "trace", flat, os.path.join(tmp, "rays.bin"), os.path.join(tmp, name + ".hits"), os.path.join(tmp, "tmax.bin")], env=e, stdout=subprocess.PIPE, text=True, check=True)
    st = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    hits = np.fromfile(os.path.join(tmp, name + ".hits"), dtype=np.uint32).reshape(n, -1)
    out[name] = (st, hits)
    print(name, "box tests/ray %.2f  tri tests/ray %.2f  occluded %.3f" % (st["boxTests"] / n, st["triTests"] / n, hits[:, -1].mean() if False else 0), st)
a, b = out["binary"][1], out["wide"][1]
print("rays", n, "different records:", int((a != b).any(axis=1).sum()))
