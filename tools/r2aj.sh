#!/usr/bin/env bash
# round 2, GPU call AJ (1 GPU): section 4c - k_tail at ptxas -O3 with single optimiser switches off
set -u
mkdir -p gpurun_out /tmp/sp
P="python tools/slot_probe2.py"
sc=mesh_small
show() { tail -1 | sed -e "s/.*numRays': \([0-9]*\).*numShadowBoxTests': \([0-9]*\).*/numRays \1 first-hit checksum \2/"; }
for v in t_noexp t_nobbmerge t_noconsts; do
  V=$PWD/bench_data/variants/$v.so
  echo "=== $v"
  echo -n "serial: "; RTB200_LIB=$V $P $sc /tmp/sp/r_${v}_serial.npy | show
  for rep in 1 2; do
    echo -n "4 slots: "; RTB200_LIB=$V RTB_SLOTS=4 $P $sc /tmp/sp/r_${v}_$rep.npy | show
    for k in 0 1 2; do python tools/slot_probe2.py --cmp /tmp/sp/r_${v}_serial_0.npy /tmp/sp/r_${v}_${rep}_$k.npy; done | sed 's/.*npy: //' | tr '\n' ';'; echo
  done
done
