#!/usr/bin/env bash
# round 2, GPU call Y (1 GPU): k_shade at 5 CTAs per SM: tile sizes, binning on / off; one `--set full` capture of k_shade
# (depth 0 and depth 1 of a 4-pass 1080p wavefront) with source, for the per-line stall picture
set -u
mkdir -p gpurun_out
qb() { timeout 150 python tools/quick_bench.py 1920 1080 16 2>&1 | grep -E "parity|profiled|Mray" | sed -e 's/counters.*//' | cut -c1-330 | tail -2; }
echo "=== default (5 CTAs/SM)"; qb
for v in t512 t256; do echo "=== $v"; RTB200_LIB=$PWD/bench_data/variants/$v.so qb; done
echo "=== no binning"; RTB_SORT_SHADE=0 qb
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 18 -c 2 -o gpurun_out/prof_r2y \
    python tools/quick_bench.py 1920 1080 4 > gpurun_out/ncu_r2y.log 2>&1
tail -2 gpurun_out/ncu_r2y.log; ls -la gpurun_out | tail -4
