#!/usr/bin/env python3
"""Aggregate the SASS source page of an .ncu-rep (ncu -i rep --page source --csv --print-source sass > file) into
contiguous regions of similar execution count: where the warp instructions go and how many lanes are active there.
Usage: tools/ncu_source_regions.py file.csv [section-index]"""
import csv
import sys


def sections(path):
    rows = list(csv.reader(open(path)))
    out, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            out.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) > 10:
            cur["rows"].append(r)
    return out


def main():
    secs = sections(sys.argv[1])
    which = [int(sys.argv[2])] if len(sys.argv) > 2 else range(len(secs))
    for si in which:
        s = secs[si]
        h = s["hdr"]
        ia, it, isrc, isamp = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Source"), h.index("# Samples")
        rows = s["rows"]
        ti = sum(int(r[ia]) for r in rows); tt = sum(int(r[it]) for r in rows); ts = sum(int(r[isamp]) for r in rows)
        print("=== [%d] %s: %d SASS lines, %.1fM warp instr, %.2f lanes/instr" % (si, s["name"], len(rows), ti / 1e6, tt / max(ti, 1)))
        # regions: split where the execution count changes by more than 25 %
        start = 0
        def flush(a, b):
            i = sum(int(r[ia]) for r in rows[a:b]); t = sum(int(r[it]) for r in rows[a:b]); sm = sum(int(r[isamp]) for r in rows[a:b])
            if i > 0.01 * ti or sm > 0.01 * ts:
                print("  sass %4d-%4d  %5.1f%% of instr  %5.1f%% of samples  %5.1f lanes   first: %s" % (
                    a, b - 1, 100.0 * i / ti, 100.0 * sm / max(ts, 1), t / max(i, 1), rows[a][isrc].strip()[:50]))
        for k in range(1, len(rows) + 1):
            if k == len(rows) or abs(int(rows[k][ia]) - int(rows[k - 1][ia])) > 0.25 * max(int(rows[k][ia]), int(rows[k - 1][ia]), 1):
                flush(start, k)
                start = k


if __name__ == "__main__":
    main()
