#!/usr/bin/env bash
# round 2, GPU call AL (1 GPU): the kernels outside the pass loop at 1080p (front buffer, bloom, block errors), ncu durations
set -u
mkdir -p gpurun_out
python tools/aux_probe.py 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_front|k_blur|k_pixel_error|k_row_error|k_block_error|k_unshard" --csv --log-file gpurun_out/aux_r2al.csv python tools/aux_probe.py > gpurun_out/aux_r2al.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/aux_r2al.csv') if l.startswith('"'))]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value"); u=h.index("Metric Unit")
tot=collections.OrderedDict(); n=collections.Counter()
for r in rows[1:]:
    name=r[k].split('(')[0]; val=float(r[v].replace(',','')); us = val*1e-3 if r[u] in ('ns','nsecond') else (val if r[u] in ('us','usecond') else val*1e3)
    tot[name]=tot.get(name,0)+us; n[name]+=1
for name in tot: print("%-16s %3d launches, %9.1f us in total, %8.1f us per launch" % (name, n[name], tot[name], tot[name]/n[name]))
PY
