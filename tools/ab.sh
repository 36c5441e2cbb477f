#!/usr/bin/env bash
# Development helper run under gpurun: A/B of kernel variants with tools/quick_bench.py.
set -u
mkdir -p gpurun_out
echo "=== default"; python tools/quick_bench.py 1920 1080 16 2>&1 | tail -2
echo "=== no carveout hint"; RTB_NO_CARVEOUT=1 python tools/quick_bench.py 1920 1080 16 2>&1 | tail -2
echo "=== batch 16M paths"; RTB_BATCH_PATHS=16000000 QB_BATCH=8 python tools/quick_bench.py 1920 1080 16 2>&1 | tail -2
echo "=== batch 32M paths"; RTB_BATCH_PATHS=32000000 QB_BATCH=16 python tools/quick_bench.py 1920 1080 16 2>&1 | tail -2
echo "=== batch 4M paths"; RTB_BATCH_PATHS=4000000 python tools/quick_bench.py 1920 1080 16 2>&1 | tail -2
for v in bench_data/variants/*.so; do echo "=== $v"; RTB200_LIB=$v python tools/quick_bench.py 1920 1080 16 2>&1 | tail -2; done
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
