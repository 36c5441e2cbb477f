#!/usr/bin/env python3
"""Mutation fuzzing of the host-side file readers under ASan + UBSan (development aid).

    python tools/loader_fuzz/run.py [iterations per input]

Inputs: the BMP textures, scene JSON files and the OBJ mesh of tests/golden/scenes, plus every *.exr / *.dds found in
the directories given by LOADER_FUZZ_EXTRA (colon separated; e.g. the temporary directories the EXR / DDS tests write).
Mutations: byte flips, truncation, overwritten header words; JSON gets structural token insertions and deletions.
Any non-zero exit (sanitizer report, abort, uncaught exception) is printed and the offending file kept under /tmp.
Round 1: 575 corrupted EXR, 480 BMP / DDS and 340 JSON / OBJ inputs, no findings."""
import glob
import os
import random
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SCENES = os.path.join(ROOT, "tests", "golden", "scenes")


def main():
    iterations = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    work = tempfile.mkdtemp(prefix="loader_fuzz_")
    exe = os.path.join(work, "fuzz_harness")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-ffp-contract=off", "-mfma",
                    "-w", "-o", exe, os.path.join(ROOT, "tools", "loader_fuzz", "fuzz_harness.cpp"), "-lz"], check=True)
    rng = random.Random(1)
    runs = findings = 0

    def run(args, blob, keep_as):
        nonlocal runs, findings
        r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        runs += 1
        if r.returncode != 0:
            findings += 1
            path = os.path.join(work, "finding%d_%s" % (findings, keep_as))
            open(path, "wb").write(blob)
            print("FINDING", path, r.stderr[-1500:])

    bitmaps = glob.glob(os.path.join(SCENES, "TEXTURES", "*.bmp"))
    for d in filter(None, os.environ.get("LOADER_FUZZ_EXTRA", "").split(":")):
        bitmaps += glob.glob(os.path.join(d, "**", "*.exr"), recursive=True) + glob.glob(os.path.join(d, "**", "*.dds"), recursive=True)
    for f in bitmaps:
        data = open(f, "rb").read()
        if len(data) > 400000:
            continue
        for t in range(iterations):
            d = bytearray(data)
            if t % 3 == 0:
                for _ in range(rng.randint(1, 8)):
                    d[rng.randrange(len(d))] = rng.randrange(256)
            elif t % 3 == 1:
                d = d[:rng.randrange(1, len(d))]
            else:
                i = rng.randrange(0, max(1, min(len(d) - 4, 400)))
                d[i:i + 4] = bytes(rng.randrange(256) for _ in range(4))
            p = os.path.join(work, "t.bin")
            open(p, "wb").write(d)
            run(["bitmap", p], d, os.path.basename(f))

    scn = os.path.join(work, "scenes") + "/"
    shutil.copytree(SCENES, scn)
    tokens = [b"{", b"}", b"[", b"]", b'"', b",", b":", b"-", b"1e999", b"null", b"true", b"0", b"-0.0", b'"mesh"', b'"sphere"', b'"mix"', b"99999999999"]
    for f in sorted(glob.glob(os.path.join(SCENES, "*.json"))):
        data = open(f, "rb").read()
        for t in range(iterations):
            d = bytearray(data)
            for _ in range(rng.randint(1, 4)):
                i, m = rng.randrange(len(d)), rng.randrange(4)
                if m == 0:
                    d[i:i + rng.randint(1, 20)] = b""
                elif m == 1:
                    d[i:i] = rng.choice(tokens)
                elif m == 2:
                    d[i] = rng.randrange(32, 127)
                else:
                    d = d[:i]
                if not d:
                    d = bytearray(b"{")
            open(scn + "t.json", "wb").write(d)
            run(["scene", scn + "t.json", scn], d, os.path.basename(f))
    obj = open(os.path.join(SCENES, "MODELS", "mesh_small.obj"), "rb").read().split(b"\n")
    for t in range(2 * iterations):
        lines = list(obj)
        for _ in range(rng.randint(1, 6)):
            i, m = rng.randrange(len(lines)), rng.randrange(5)
            lines[i] = [b"f 1 2 999999", b"f -1 -2 -3", b"v nan inf 1e40", lines[i][:rng.randrange(len(lines[i]) + 1)], b"f 1/2/3 4//5 6/7"][m]
        blob = b"\n".join(lines[:rng.randrange(10, len(lines))] if t % 3 == 0 else lines)
        open(scn + "MODELS/mesh_small.obj", "wb").write(blob)
        run(["scene", scn + "mesh_small.json", scn], blob, "mesh_small.obj")
    print("runs %d, findings %d (work directory %s)" % (runs, findings, work))
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
