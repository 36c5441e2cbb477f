#!/usr/bin/env python3
"""Per-CUDA-line view of an ncu source page (ncu -i rep --page source --csv --print-source cuda,sass > file): samples,
warp instructions, lanes and the top stall reasons of the hottest source lines of each captured kernel.
Usage: tools/ncu_lines.py file.csv [top-n]"""
import csv
import sys
from collections import defaultdict

csv.field_size_limit(1 << 30)


def main():
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kernels, names, cur, fpath, hdr = [], [], None, None, None
    seen_files = set()
    for r in csv.reader(open(sys.argv[1])):
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
            continue
        if r[0] == "Function Name":
            # one (file, kernel) section per source file a launch touched: a new launch starts when the kernel changes or a
            # file comes round again
            fn = r[1].split("(")[0].replace("void ", "").strip()
            if cur is None or fn != names[-1] or fpath in seen_files:
                cur = defaultdict(lambda: defaultdict(float)); kernels.append(cur); names.append(fn); seen_files = set()
            seen_files.add(fpath)
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        # the rows that carry a line number hold that line's totals; the SASS rows below them repeat the parts
        if hdr is None or len(r) < len(hdr) - 2 or not r[0].isdigit():
            continue
        # (ncu does not escape quotes inside the source text: find the "-","-" of the two SASS columns instead of counting)
        k = next((i for i in range(2, len(r) - 1) if r[i] == "-" and r[i + 1] == "-"), None)
        if k is None:
            continue
        r = [r[0], ",".join(r[1:k])] + r[k:]
        r = r + ["0"] * (len(hdr) - len(r))
        key = (fpath.split("/")[-1], int(r[0]), r[1].strip()[:90])
        d = cur[key]
        for name in ("# Samples", "Instructions Executed", "Thread Instructions Executed"):
            d[name] += float(r[hdr.index(name)].replace("-", "0") or 0)
        for i, name in enumerate(hdr):
            if name.startswith("stall_") and "Not Issued" not in name:
                d[name] += float((r[i] if r[i] != "-" else "0") or 0)
    for ki, k in enumerate(kernels):
        ts = sum(d["# Samples"] for d in k.values()); ti = sum(d["Instructions Executed"] for d in k.values())
        print("=== [%d] %s: %.0f samples, %.1f M warp instructions" % (ki, names[ki], ts, ti / 1e6))
        tot = defaultdict(float)
        for d in k.values():
            for n, v in d.items():
                if n.startswith("stall_"): tot[n] += v
        print("    stalls overall: " + ", ".join("%s %.1f%%" % (n[6:], 100 * v / max(ts, 1)) for n, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
        for key, d in sorted(k.items(), key=lambda x: -x[1]["# Samples"])[:top]:
            st = sorted(((n[6:], v) for n, v in d.items() if n.startswith("stall_")), key=lambda x: -x[1])[:3]
            print("  %5.1f%% smp %5.1f%% ins %4.1f ln  %-16s:%-4d %-90s | %s" % (
                100 * d["# Samples"] / max(ts, 1), 100 * d["Instructions Executed"] / max(ti, 1), d["Thread Instructions Executed"] / max(d["Instructions Executed"], 1),
                key[0], key[1], key[2], " ".join("%s %.0f%%" % (n, 100 * v / max(d["# Samples"], 1)) for n, v in st)))


if __name__ == "__main__":
    main()
