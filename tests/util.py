"""Shared helpers for the parity tests."""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_EXACT = os.path.join(ROOT, "oracle", "_ref", "exact")
ORACLE_AVX = os.path.join(ROOT, "oracle", "_ref", "avx")

RENDERER_IDS = {"mis": (1, 0), "pt": (0, 0), "dbgtri": (2, 1), "dbgdepth": (2, 2),
                "dbgbase": (2, 8), "dbgnorm": (2, 4), "dbgrough": (2, 10), "dbguv": (2, 7)}   # (RtbRenderer, RtbDebugMode)
GOLDEN_RENDERS = [
    ("c1_sphere_area", "mis"),
    ("cornell", "mis"), ("cornell", "pt"), ("cornell", "dbgtri"), ("cornell", "dbgdepth"),
    ("cornell_dof", "mis"),
    ("materials", "mis"), ("materials", "pt"),
    ("mesh_small", "mis"), ("mesh_small", "pt"), ("mesh_small", "dbgtri"), ("mesh_small", "dbgdepth"),
    ("textures", "mis"), ("textures", "pt"), ("textures", "dbgbase"), ("textures", "dbgnorm"), ("textures", "dbgrough"), ("textures", "dbguv"),
]


def golden(name):
    return os.path.join(GOLDEN, name)


def load_sum(path, w, h):
    return np.fromfile(path, dtype=np.float32).reshape(h, w, 3)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_identical(a, b, what):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    same = bits(a) == bits(b)
    if not same.all():
        bad = np.argwhere(~same)
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
        raise AssertionError("%s: %d of %d floats differ bitwise (first at %s: got %r want %r, max rel %.3e)"
                             % (what, bad.shape[0], same.size, bad[0].tolist(), a[tuple(bad[0])], b[tuple(bad[0])], float(rel.max())))


def oracle_available(variant="exact"):
    root = ORACLE_EXACT if variant == "exact" else ORACLE_AVX
    return os.path.exists(os.path.join(root, "rtref")) and os.path.exists(os.path.join(root, "libCore.so"))


def run_rtref(args, variant="exact", seed="12345", timeout=1800):
    """Run the oracle harness (the real reference) - test infrastructure only."""
    root = ORACLE_EXACT if variant == "exact" else ORACLE_AVX
    run_dir = os.path.join(root, "run")
    os.makedirs(run_dir, exist_ok=True)
    env = dict(os.environ, RT_FIXED_SEED=seed)
    r = subprocess.run([os.path.join(root, "rtref")] + [str(a) for a in args], cwd=run_dir, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert r.returncode == 0, "rtref failed: %s\n%s" % (args, r.stderr[-2000:])
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def reference_scene_manifest():
    """tests/golden/refscenes: the reference's own Data/TestScenes rendered by the reference (make_golden.py)."""
    p = os.path.join(GOLDEN, "refscenes", "manifest.json")
    return json.load(open(p)) if os.path.exists(p) else {}


REFERENCE_SCENE_RENDERS = [(n, r) for n, e in sorted(reference_scene_manifest().items()) if e.get("loads") for r in ("mis", "pt")]


def read_progress(path):
    """rtref --dump-progress: one record per Render(): RenderingProgress, the block list after the pass and, after even
    passes, ComputeBlockError of each block of the PREVIOUS list (what UpdateBlocksList compared with its thresholds)."""
    import numpy as np
    raw = np.fromfile(path, dtype=np.uint32)
    out, i = [], 0
    while i < len(raw):
        rec = {"passesFinished": int(raw[i]), "activePixels": int(raw[i + 1]), "activeBlocks": int(raw[i + 2]),
               "converged": raw[i + 3:i + 4].view(np.float32)[0], "averageError": raw[i + 4:i + 5].view(np.float32)[0]}
        n = int(raw[i + 5]); i += 6
        rec["blocks"] = raw[i:i + 4 * n].reshape(n, 4).copy(); i += 4 * n
        m = int(raw[i]); i += 1
        rec["previousErrors"] = raw[i:i + m].view(np.float32).copy(); i += m
        out.append(rec)
    return out
