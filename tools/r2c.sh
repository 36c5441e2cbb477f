#!/usr/bin/env bash
# round 2, GPU call C: parity with the fixed 8-wide any-hit kernel, A/B of its launch shapes, k_tail spill test, ncu capture
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
qb() { timeout 200 python tools/quick_bench.py 1920 1080 16 2>&1 | grep -E "parity|profiled|Mray" | sed -e 's/raygenMs.*traverseMs/traverseMs/' -e 's/counters.*//' | cut -c1-330 | tail -3; }
echo "=== default (wide kernel 512x2)"; qb
echo "=== hot=0"; RTB_HOT_NODES=0 qb
echo "=== hot=256"; RTB_HOT_NODES=256 qb
echo "=== simple kernel + wide tree"; RTB_SHADOW_KERNEL=simple qb
echo "=== simple kernel, binary tree (round 1 walk)"; RTB_SHADOW_KERNEL=simple RTB_NO_WIDE=1 qb
for v in bench_data/variants/sw_*.so; do echo "=== $v"; RTB200_LIB=$v qb; done
echo "=== leaf tris 2"; RTB_WIDE_LEAF_TRIS=2 qb
echo "=== leaf tris 4"; RTB_WIDE_LEAF_TRIS=4 qb
echo "=== slots: k_tail with spills / without, under overlap"
sp() { python tools/slot_probe.py "$@" 2>&1 | head -3 | cut -c1-40 | tr '\n' ';'; echo; }
echo "slots4 default (k_tail off with slots):"; RTB_SLOTS=4 sp mesh_small
echo "slots4 + k_tail (spilling build):"; RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 sp mesh_small
echo "slots4 + k_tail (spill-free build, 250 regs):"; RTB200_LIB=bench_data/variants/tail2.so RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 sp mesh_small
echo "slots4 + k_tail (spill-free build) again:"; RTB200_LIB=bench_data/variants/tail2.so RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 sp mesh_small
echo "=== ncu"
ncu --set full --clock-control none --import-source on -k regex:"k_shadow_wide" -s 18 -c 2 -o gpurun_out/prof_r2c \
    python tools/quick_bench.py 1920 1080 4 > gpurun_out/ncu_r2c.log 2>&1
tail -2 gpurun_out/ncu_r2c.log
python tools/ncu_summary.py gpurun_out/prof_r2c.ncu-rep 2>&1 | head -60
