#!/usr/bin/env bash
# Development helper run under gpurun: A/B of kernel variants with tools/quick_bench.py.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== default"; timeout 120 python tools/quick_bench.py 1920 1080 16 2>&1 | tail -3
for v in bench_data/variants/*.so; do echo "=== $v"; RTB200_LIB=$v timeout 120 python tools/quick_bench.py 1920 1080 16 2>&1 | tail -2; done
