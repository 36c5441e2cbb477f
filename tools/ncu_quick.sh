#!/usr/bin/env bash
# Run under gpurun: one `--set full` capture of depth 0 / depth 1 of k_traverse and k_shadow.   tools/ncu_quick.sh <tag>
set -u
TAG="${1:-q}"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_traverse|k_shadow" -s 36 -c 4 -o gpurun_out/prof_$TAG \
    python tools/quick_bench.py 1920 1080 4 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
