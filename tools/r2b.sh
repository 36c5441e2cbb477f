#!/usr/bin/env bash
# round 2, GPU call B: parity with the 8-wide any-hit kernel, A/B of its launch shapes, overlap-bug bisect
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
qb() { timeout 200 python tools/quick_bench.py 1920 1080 16 2>&1 | grep -E "parity|profiled|Mray" | sed -e 's/raygenMs.*traverseMs/traverseMs/' | cut -c1-330; }
echo "=== default (wide kernel 512x2)"; qb
echo "=== hot=0"; RTB_HOT_NODES=0 qb
echo "=== simple kernel + wide tree"; RTB_SHADOW_KERNEL=simple qb
echo "=== simple kernel, binary tree (round 1 walk)"; RTB_SHADOW_KERNEL=simple RTB_NO_WIDE=1 qb
for v in bench_data/variants/*.so; do echo "=== $v"; RTB200_LIB=$v qb; done
echo "=== leaf tris 2"; RTB_WIDE_LEAF_TRIS=2 qb
echo "=== leaf tris 4"; RTB_WIDE_LEAF_TRIS=4 qb
echo "=== slots bisect"
sp() { python tools/slot_probe.py "$@" 2>&1 | head -3 | cut -c1-40 | tr '\n' ';'; echo; }
echo "slots4 mesh_small mis:"; RTB_SLOTS=4 sp mesh_small
echo "slots4 mesh_small mis simple kernel:"; RTB_SHADOW_KERNEL=simple RTB_SLOTS=4 sp mesh_small
echo "slots4 mesh_small mis simple kernel nowide:"; RTB_NO_WIDE=1 RTB_SHADOW_KERNEL=simple RTB_SLOTS=4 sp mesh_small
echo "slots4 mesh_small mis notail:"; RTB_TAIL_PATHS=0 RTB_SLOTS=4 sp mesh_small
echo "slots4 mesh_small mis notail simple nowide:"; RTB_TAIL_PATHS=0 RTB_NO_WIDE=1 RTB_SHADOW_KERNEL=simple RTB_SLOTS=4 sp mesh_small
echo "slots4 mesh_small pt:"; RTB_SLOTS=4 sp mesh_small 0 0
echo "slots4 mesh_small pt notail:"; RTB_TAIL_PATHS=0 RTB_SLOTS=4 sp mesh_small 0 0
echo "slots4 mesh_small dbg:"; RTB_SLOTS=4 sp mesh_small 2 1
echo "slots4 serialize:"; RTB_SERIALIZE_SLOTS=1 RTB_SLOTS=4 sp mesh_small
echo "slots4 maxconn1:"; CUDA_DEVICE_MAX_CONNECTIONS=1 RTB_SLOTS=4 sp mesh_small
echo "slots2:"; RTB_SLOTS=2 sp mesh_small
echo "=== racecheck"
RTB_SLOTS=4 timeout 400 compute-sanitizer --tool racecheck python tools/slot_probe.py mesh_small 2>&1 | grep -v "^bad pix\|^a \[" | tail -15 | cut -c1-250
