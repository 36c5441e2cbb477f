#!/usr/bin/env bash
# round 2, GPU call X (1 GPU): k_shade per integrator (split), L2 prefetch of the tile's path state, queue prefetch in the
# persistent kernels, k_shade register budgets / tile sizes after the split; then the whole GPU suite, timed
set -u
mkdir -p gpurun_out
qb() { timeout 150 python tools/quick_bench.py 1920 1080 16 2>&1 | grep -E "parity|profiled|Mray" | sed -e 's/counters.*//' | cut -c1-330; }
echo "=== default (split)"; qb
for v in base split pf_shade pf_queue pf_queue2k pf_both split_mb5 split_mb4 tile512 tile2048; do echo "=== $v"; RTB200_LIB=$PWD/bench_data/variants/$v.so qb | tail -2; done
echo "=== default again"; qb | tail -2
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2x_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR|^real" gpurun_out/r2x_pytest.log | tail -12
