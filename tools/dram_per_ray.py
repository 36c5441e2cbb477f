#!/usr/bin/env python3
"""profiles/traverse_dram_bytes.json (bench.py's roofline.traffic) from the DRAM pass of tools/profile.sh:
   tools/dram_per_ray.py gpurun_out/dram_<tag>.csv gpurun_out/dram_<tag>.log <tag>
The csv holds dram__bytes_read/write.sum of the 9 + 9 traversal launches of ONE 8-pass 1080p wavefront (ncu, one GPU);
the log is tools/quick_bench.py's output of the same run, whose `profiled` line has the rays those launches carried."""
import ast, csv, json, os, re, sys

csvp, logp, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(csvp) if l.startswith('"'))]
hdr = rows[0]; k, m, v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot, launches = {}, {}
for r in rows[1:]:
    if not r[m].startswith("dram__bytes"): continue
    name = "k_traverse" if "k_traverse" in r[k] else "k_shadow"
    tot[name] = tot.get(name, 0.0) + float(r[v].replace(",", ""))
    if r[m] == "dram__bytes_read.sum": launches[name] = launches.get(name, 0) + 1
prof = [ast.literal_eval(l.split(":", 1)[1].strip()) for l in open(logp) if l.startswith("profiled")][0]
rays = {"k_traverse": prof["traverseRays"], "k_shadow": prof["shadowRays"]}
out = {"source": "profiles/%s_dram_per_launch.csv: dram__bytes_read.sum + dram__bytes_write.sum summed over the %d + %d traversal launches of one 8-pass 1080p wavefront (ncu, N=1), divided by the rays those launches carried" % (tag, launches.get("k_traverse", 0), launches.get("k_shadow", 0))}
for name in tot:
    per_ray = tot[name] / rays[name]
    for alias in ([name] if name == "k_traverse" else ["k_shadow", "k_shadow_wide"]):
        out[alias + "_dram_bytes_per_ray"] = per_ray
    out[name + "_dram_bytes_per_launch"] = tot[name] / launches[name]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traverse_dram_bytes.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
