#!/usr/bin/env bash
# Development helper run under gpurun: parity tests, quick bench, and (optionally) an ncu capture of the hot kernels.
#   tools/gpu_probe.sh <tag> [ncu]
set -u
TAG="${1:-probe}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python tools/quick_bench.py 1920 1080 8 2>&1 | tail -3
if [ "${2:-}" = "ncu" ]; then
    ncu --set full --clock-control none --import-source on -k regex:"k_traverse|k_shadow" -s 38 -c 4 -o gpurun_out/prof_$TAG python tools/quick_bench.py 1920 1080 1 > gpurun_out/ncu_$TAG.log 2>&1
fi
