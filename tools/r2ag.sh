#!/usr/bin/env bash
# round 2, GPU call AG (1 GPU): section 4c - which ingredient of k_tail's closest-hit walk is sensitive to overlap?
set -u
mkdir -p gpurun_out /tmp/sp
P="python tools/slot_probe2.py"
sc=mesh_small
show() { tail -1 | sed -e "s/.*numRays': \([0-9]*\).*numShadowBoxTests': \([0-9]*\).*/numRays \1 first-hit checksum \2/"; }
for v in tail_probe tp_ldg128 tp_plain tp_O1; do
  V=$PWD/bench_data/variants/$v.so
  echo "=== $v"
  echo -n "serial: "; RTB200_LIB=$V $P $sc /tmp/sp/r_${v}_serial.npy | show
  for rep in 1 2; do
    echo -n "4 slots + k_tail: "; RTB200_LIB=$V RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/r_${v}_$rep.npy | show
    python tools/slot_probe2.py --cmp /tmp/sp/r_${v}_serial_0.npy /tmp/sp/r_${v}_${rep}_1.npy
  done
done
echo "=== default library, simple traversal kernels (k_traverse / k_shadow), 4 slots, no k_tail"
echo -n "serial: "; RTB_TRAV_KERNEL=simple RTB_SHADOW_KERNEL=simple $P $sc /tmp/sp/r_simple_serial.npy | show
for rep in 1 2; do
  echo -n "4 slots: "; RTB_TRAV_KERNEL=simple RTB_SHADOW_KERNEL=simple RTB_SLOTS=4 $P $sc /tmp/sp/r_simple_$rep.npy | show
  python tools/slot_probe2.py --cmp /tmp/sp/r_simple_serial_0.npy /tmp/sp/r_simple_${rep}_1.npy
done
