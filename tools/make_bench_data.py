#!/usr/bin/env python3
"""Development helper (authoring container): (re)create the git-ignored bench_data/ that tools/quick_bench.py and
tools/ab.sh use on the GPU box:

    bench_data/scenes/...            the synthetic sponza-class scene (same generator as bench.py)
    bench_data/sponza_class.rtb      that scene flattened by the product's own host loader
    bench_data/sponza_480.sum/.passes   2 passes at 480x270 rendered by the exact-math reference (oracle/_ref/exact)
"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "scenes"))
import make_scenes  # noqa: E402
from raytracer_b200 import host_api  # noqa: E402
from tests import util  # noqa: E402


def main():
    bd = os.path.join(ROOT, "bench_data")
    scn = os.path.join(bd, "scenes")
    sj = os.path.join(scn, "sponza_class.json")
    if not os.path.exists(sj):
        make_scenes.write_all(scn, sponza=True)
    rtb = os.path.join(bd, "sponza_class.rtb")
    if not os.path.exists(rtb):
        host_api.Scene(sj, scn + "/").save_flat(rtb, 1920, 1080)
    sump = os.path.join(bd, "sponza_480.sum")
    if not os.path.exists(sump) and util.oracle_available("exact"):
        util.run_rtref(["render", "--scene", sj, "--data", scn + "/", "--width", 480, "--height", 270, "--passes", 2, "--dims", 128,
                        "--renderer", "Path Tracer MIS", "--depth", 8, "--strategy", "all", "--dump-sum", sump,
                        "--dump-passes", os.path.join(bd, "sponza_480.passes")])
    print("bench_data ready:", sorted(os.listdir(bd)))


if __name__ == "__main__":
    main()
