#!/usr/bin/env bash
# round 2, GPU call AM (1 GPU): k_blur with larger load batches per thread (the kernel is one dependent chain per line)
set -u
mkdir -p gpurun_out
for v in blur64 blur128; do
  echo "=== $v"
  RTB200_LIB=$PWD/bench_data/variants/$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bloom" 2>&1 | tail -1
  RTB200_LIB=$PWD/bench_data/variants/$v.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_blur" --csv --log-file gpurun_out/blur_$v.csv python tools/aux_probe.py > gpurun_out/blur_$v.log 2>&1
  grep "read_front bloom=0.1" gpurun_out/blur_$v.log
  python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/blur_$v.csv') if l.startswith('"'))]
h=rows[0]; v=h.index("Metric Value"); u=h.index("Metric Unit")
vals=[float(r[v].replace(',',''))*(1e-3 if r[u] in ('ns','nsecond') else 1.0) for r in rows[1:]]
print("k_blur: %d launches, %.1f us per launch" % (len(vals), sum(vals)/len(vals)))
PY
done
