#!/usr/bin/env bash
# round 2, GPU call D: triangle pre-test, optimal collapse, refilling wide kernel
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2d_pytest.log 2>&1; grep -E "passed|failed" gpurun_out/r2d_pytest.log | tail -2
qb() { timeout 200 python tools/quick_bench.py 1920 1080 16 2>&1 | grep -E "parity|profiled|Mray" | sed -e 's/raygenMs.*traverseMs/traverseMs/' -e 's/counters.*//' -e "s/'traverseLaunches.*//" | cut -c1-330 | tail -3; }
echo "=== default (wide kernel 512x2, refill 22)"; qb
echo "=== hot=0"; RTB_HOT_NODES=0 qb
echo "=== simple kernel, binary tree"; RTB_SHADOW_KERNEL=simple RTB_NO_WIDE=1 qb
for v in bench_data/variants/*.so; do echo "=== $v"; RTB200_LIB=$v qb; done
echo "=== ncu"
ncu --set full --clock-control none --import-source on -k regex:"k_shadow_wide|k_traverse" -s 36 -c 4 -o gpurun_out/prof_r2d \
    python tools/quick_bench.py 1920 1080 4 > gpurun_out/ncu_r2d.log 2>&1
tail -2 gpurun_out/ncu_r2d.log
python tools/ncu_summary.py gpurun_out/prof_r2d.ncu-rep 2>&1 | grep -E "^---|time_duration|thread_inst_executed_per|issue_active|inst_executed.sum|hit_rate|long_scoreboard"
