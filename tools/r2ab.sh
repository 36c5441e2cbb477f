#!/usr/bin/env bash
# round 2, GPU call AB (1 GPU): A/B of whatever sits in bench_data/variants against the default build
set -u
mkdir -p gpurun_out
qb() { timeout 150 python tools/quick_bench.py 1920 1080 16 2>&1 | grep -E "parity|profiled|Mray" | sed -e 's/counters.*//' -e "s/'traverseLaunches.*//" | cut -c1-200 | tail -3; }
echo "=== default"; qb
for v in bench_data/variants/*.so; do echo "=== $v"; RTB200_LIB=$PWD/$v qb; done
echo "=== default again"; qb | tail -2
