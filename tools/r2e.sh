#!/usr/bin/env bash
# round 2, GPU call E: VCM on the device (tests), full suite, A/B of the any-hit kernels with repetitions
set -u
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_vcm.py tests/test_furnace.py -m gpu -q -x ) > gpurun_out/r2e_vcm.log 2>&1; grep -E "passed|failed|Error|error|assert" gpurun_out/r2e_vcm.log | tail -12
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2e_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/r2e_pytest.log | tail -3
qb() { timeout 200 python tools/quick_bench.py 1920 1080 16 2>&1 | grep -E "profiled" | sed -e 's/raygenMs.*traverseMs/traverseMs/' -e "s/'traverseLaunches.*//" | cut -c1-330 | tail -2 | tr '\n' ' '; echo; }
for rep in 1 2; do
echo "=== default (wide kernel 384x2, no refill, hot 0)"; qb
echo "=== simple kernel, binary tree"; RTB_SHADOW_KERNEL=simple RTB_NO_WIDE=1 qb
for v in bench_data/variants/*.so; do echo "=== $v"; RTB200_LIB=$v qb; done
done
