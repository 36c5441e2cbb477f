#!/usr/bin/env bash
# round 2, GPU call Z (1 GPU): k_shade after the ncu line view - zero-numerator bypass in normalize3, one atomic instruction
# for the three queue counters, raster gather only for paths that need it, key chains side by side, tile 512
set -u
mkdir -p gpurun_out
qb() { timeout 150 python tools/quick_bench.py 1920 1080 16 2>&1 | grep -E "parity|profiled|Mray" | sed -e 's/counters.*//' | cut -c1-330 | tail -3; }
echo "=== default"; qb
for v in bench_data/variants/*.so; do echo "=== $v"; RTB200_LIB=$PWD/$v qb | tail -2; done
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2z_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR|^real" gpurun_out/r2z_pytest.log | tail -8
