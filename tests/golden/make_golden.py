#!/usr/bin/env python3
"""Generate the committed golden fixtures from the REAL reference (oracle/_ref/exact/rtref, built by
oracle/build_ref.sh from /root/reference).  Run in the authoring container only:

    oracle/build_ref.sh exact && python tests/golden/make_golden.py

For every test scene (tests/scenes/make_scenes.py) it stores
    <scene>.rtb            the reference's own Scene (after BuildBVH) flattened to the RtbSceneDesc file format
    <scene>.passes         per-pass Halton seed vectors + AA offsets the reference's Viewport used (RT_FIXED_SEED=12345)
    <scene>_<r>.sum        Viewport::GetSumBuffer() after N passes, float32 RGB, renderer r in {mis, pt, dbgtri, dbgdepth}
    <scene>_<r>.json       the reference's ray counters for that render
    <scene>.rays/.hits/.shadow_hits/.tmax   Scene::Traverse / Traverse_Shadow on a ray batch
    <scene>.primary        Camera::GenerateRay for every pixel of pass 0
    sampler_bn.u32 / sampler_nobn.u32       GenericSampler streams
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "tests", "scenes"))
import make_scenes  # noqa: E402

RTREF_DIR = os.path.join(ROOT, "oracle", "_ref", "exact")
RUN_DIR = os.path.join(RTREF_DIR, "run")
SEED = "12345"
W = H = 64
PASSES = 4
RAY_DTYPE = np.dtype([("origin", np.float32, 4), ("dir", np.float32, 4), ("invDir", np.float32, 4), ("originDivDir", np.float32, 4)])

SCENES = {
    # name: (sampler dimensions, renderers)
    "c1_sphere_area": (128, ["mis"]),
    "cornell": (128, ["mis", "pt", "dbgtri", "dbgdepth"]),
    "cornell_dof": (128, ["mis"]),      # depth of field: 2 extra sampler dimensions per primary ray
    "materials": (256, ["mis", "pt"]),
    "mesh_small": (128, ["mis", "pt", "dbgtri", "dbgdepth"]),
    # every texture kind on every material slot + normal maps + environment map (RTBS v2 file with float texels)
    "textures": (192, ["mis", "pt", "dbgbase", "dbgnorm", "dbgrough", "dbguv"]),
}
RENDERERS = {"mis": ("Path Tracer MIS", 0), "pt": ("Path Tracer", 0), "dbgtri": ("Debug", 1), "dbgdepth": ("Debug", 2),
             "dbgbase": ("Debug", 8), "dbgnorm": ("Debug", 4), "dbgrough": ("Debug", 10), "dbguv": ("Debug", 7)}


def rtref(*args):
    env = dict(os.environ, RT_FIXED_SEED=SEED)
    r = subprocess.run([os.path.join(RTREF_DIR, "rtref")] + [str(a) for a in args], cwd=RUN_DIR, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise SystemExit("rtref failed: %s\n%s" % (" ".join(str(a) for a in args), r.stderr[-2000:]))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def make_rays(rng, n, lo, hi):
    """Ray(origin, dir) exactly as the reference constructs it (Ray.h:23-28), in float32."""
    rays = np.zeros(n, dtype=RAY_DTYPE)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    # edge cases: axis-aligned directions (zero components -> infinite invDir, NaN slabs), rays starting on box planes
    d[: n // 16, 0] = 0.0
    d[n // 16: n // 8, 1] = 0.0
    d[n // 8: n // 8 + n // 32, 1:] = 0.0
    o[n // 4: n // 4 + n // 16, 1] = np.float32(-5.0)
    f = np.float32
    ln = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + (d[:, 2] * d[:, 2] + f(0)), dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (f(1) / d).astype(np.float32)
        il = (f(1) / ln).astype(np.float32)
        inv_dir = (ln[:, None] * inv).astype(np.float32)
        nd = (d * il[:, None]).astype(np.float32)
        odd = (o * inv_dir).astype(np.float32)
    rays["origin"][:, :3] = o
    rays["dir"][:, :3] = nd
    rays["invDir"][:, :3] = inv_dir
    rays["originDivDir"][:, :3] = odd
    return rays


REFERENCE_DATA = "/root/reference/Data"


def make_reference_scenes(out):
    """The reference's OWN test scenes (Data/TestScenes/*.json, read in place, never copied): for every scene its
    loader accepts, the flattened Scene + MIS / PT sum buffers at 32x32 x 3 passes -> tests/golden/refscenes/."""
    import glob
    if not os.path.isdir(REFERENCE_DATA):
        print("no %s: tests/golden/refscenes left as committed" % REFERENCE_DATA)
        return
    dst = os.path.join(out, "refscenes")
    os.makedirs(dst, exist_ok=True)
    w = h = 32
    manifest = {}
    for path in sorted(glob.glob(os.path.join(REFERENCE_DATA, "TestScenes", "*.json"))):
        name = os.path.basename(path)[:-5]
        base = os.path.join(dst, name)
        common = ["--scene", path, "--data", REFERENCE_DATA + "/", "--width", w, "--height", h]
        try:
            rtref("flat", *common, "--out", base + ".rtb")
        except SystemExit:
            manifest[name] = {"loads": False}        # the reference itself rejects (or crashes on) this scene
            continue
        entry = {"loads": True, "width": w, "height": h, "passes": 3, "dims": 256}
        for r in ("mis", "pt"):
            c = rtref("render", *common, "--passes", 3, "--dims", 256, "--renderer", RENDERERS[r][0], "--depth", 8, "--strategy", "all",
                      "--dump-sum", "%s_%s.sum" % (base, r), "--dump-passes", base + ".passes")
            entry[r] = {k: c[k] for k in ("numPrimaryRays", "numRays", "numShadowRays", "numShadowRaysHit")}
        manifest[name] = entry
    json.dump(manifest, open(os.path.join(dst, "manifest.json"), "w"), indent=1, sort_keys=True)


def main():
    out = HERE
    scn = os.path.join(out, "scenes")
    make_scenes.write_all(scn)
    rng = np.random.default_rng(7)
    for name, (dims, renderers) in SCENES.items():
        sj = os.path.join(scn, name + ".json")
        common = ["--scene", sj, "--data", scn + "/", "--width", W, "--height", H]
        rtref("flat", *common, "--out", os.path.join(out, name + ".rtb"))
        for r in renderers:
            rname, dbg = RENDERERS[r]
            extra = ["--dump-passes", os.path.join(out, name + ".passes")] if r == renderers[0] else []
            if name in ("cornell", "mesh_small") and r == "mis":
                # Viewport::ComputeBlockError of the whole image, single pixels / rows / columns and random rectangles
                brng = np.random.default_rng(11)
                blocks = [(0, W, 0, H), (0, 1, 0, 1), (W - 1, W, H - 1, H), (0, W, 5, 6), (7, 8, 0, H)]
                while len(blocks) < 40:
                    x0, x1 = sorted(int(v) for v in brng.integers(0, W + 1, 2))
                    y0, y1 = sorted(int(v) for v in brng.integers(0, H + 1, 2))
                    if x0 < x1 and y0 < y1:
                        blocks.append((x0, x1, y0, y1))
                np.array(blocks, dtype=np.uint32).tofile(os.path.join(out, name + ".blocks"))
                extra += ["--block-errors", os.path.join(out, name + ".blocks"), "--dump-block-errors", os.path.join(out, name + "_mis.blockerr")]
            c = rtref("render", *common, "--passes", PASSES, "--dims", dims, "--renderer", rname, "--debug-mode", dbg,
                      "--depth", 8, "--strategy", "all", "--dump-sum", os.path.join(out, "%s_%s.sum" % (name, r)),
                      "--dump-sum2", os.path.join(out, "%s_%s.sum2" % (name, r)) if r == "mis" else "/dev/null",
                      # Viewport::GetFrontBuffer with the default PostprocessParams, dithering off (not reproducible)
                      *(["--dump-front", os.path.join(out, "%s_%s.front" % (name, r))] if r == "mis" else []), *extra)
            keep = {k: c[k] for k in ("numPrimaryRays", "numRays", "numShadowRays", "numShadowRaysHit", "passes", "width", "height")}
            keep["dims"] = dims
            json.dump(keep, open(os.path.join(out, "%s_%s.json" % (name, r)), "w"))
        if name in ("cornell", "mesh_small", "materials"):
            n = 3072
            lo, hi = (-6.0, 6.0) if name != "mesh_small" else (-17.0, 17.0)
            rays = make_rays(rng, n, lo, hi)
            if name == "mesh_small":
                rays["origin"][:, 1] = np.abs(rays["origin"][:, 1]) * np.float32(0.6)
                rays["originDivDir"][:, :3] = rays["origin"][:, :3] * rays["invDir"][:, :3]
            rays.tofile(os.path.join(out, name + ".rays"))
            tmax = rng.uniform(0.05, 30.0, size=n).astype(np.float32)
            tmax[::7] = np.float32(3.4028234e38) * np.float32(0.999)
            tmax.tofile(os.path.join(out, name + ".tmax"))
            rtref("trace", "--scene", sj, "--data", scn + "/", "--rays", os.path.join(out, name + ".rays"), "--out", os.path.join(out, name + ".hits"))
            rtref("trace", "--scene", sj, "--data", scn + "/", "--rays", os.path.join(out, name + ".rays"), "--shadow", 1,
                  "--tmax", os.path.join(out, name + ".tmax"), "--out", os.path.join(out, name + ".shadow_hits"))
        if name == "textures":
            # ITexture::Evaluate at adversarial coordinates: negative, exactly on texel / tile boundaries, just below 1,
            # huge (floor loses the fraction), tiny, and a dense random cloud
            f = np.float32
            special = [0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 0.25, 0.75, 1.0 - 2.0 ** -24, 2.0 ** -24, -2.0 ** -24, 2.0 ** -149, 1.0 / 16.0, 1.0 / 8.0, 7.0 / 8.0,
                       15.0 / 16.0, 3.0, -3.0, 17.5, -17.5, 1024.25, -1024.25, 8388608.0, 16777216.0, -16777217.0, 1.0e9, -1.0e9, 0.333333343, 12.0 / 13.0, 4.0 / 5.0]
            grid = np.array([(a, b) for a in special for b in special[::3]], dtype=f)
            cloud = np.concatenate([rng.uniform(-2.0, 3.0, size=(600, 2)), rng.normal(size=(200, 2)) * 40.0]).astype(f)
            uv = np.concatenate([grid, cloud]).astype(f)
            uv.tofile(os.path.join(out, "textures.uv"))
            rtref("texeval", "--scene", sj, "--data", scn + "/", "--uv", os.path.join(out, "textures.uv"), "--out", os.path.join(out, "textures.texeval"))
        if name in ("cornell", "mesh_small"):
            rtref("primary", *common, "--passes", os.path.join(out, name + ".passes"), "--pass", 1, "--out", os.path.join(out, name + ".primary"))
    make_reference_scenes(out)
    # sampler streams: 16x16 pixels (offset into the 128-tile by using a 200-wide virtual row is not needed: x,y < 16)
    p = os.path.join(out, "cornell.passes")
    rtref("sampler", "--scene", "unused", "--passes", p, "--width", 24, "--height", 20, "--num-dims", 40, "--blue-noise", 1, "--out", os.path.join(out, "sampler_bn.u32"))
    rtref("sampler", "--scene", "unused", "--passes", p, "--width", 24, "--height", 20, "--num-dims", 40, "--blue-noise", 0, "--out", os.path.join(out, "sampler_nobn.u32"))
    # the generated scene sources are only needed to rebuild the fixtures
    total = sum(os.path.getsize(os.path.join(out, f)) for f in os.listdir(out) if os.path.isfile(os.path.join(out, f)))
    print("golden fixtures written to %s (%.1f KiB)" % (out, total / 1024.0))


if __name__ == "__main__":
    main()
