#!/usr/bin/env bash
# round 2, GPU call AF (1 GPU): section 4c - one L1 / shared-memory split for EVERY kernel of the library (RTB_CARVEOUT now covers
# the persistent kernels too): does k_tail still find different hits under overlap?
set -u
mkdir -p gpurun_out /tmp/sp
P="python tools/slot_probe2.py"
V=$PWD/bench_data/variants/tail_probe.so
sc=mesh_small
show() { tail -1 | sed -e "s/.*numShadowBoxTests': \([0-9]*\).*/first-hit checksum \1/"; }
echo "--- serial";                            RTB200_LIB=$V $P $sc /tmp/sp/q_serial.npy | show
for c in none 100 75 50 25 0; do
  for rep in 1 2; do
    if [ $c = none ]; then echo -n "--- 4 slots + k_tail, default splits: "; RTB200_LIB=$V RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/q_${c}_$rep.npy | show
    else echo -n "--- 4 slots + k_tail, RTB_CARVEOUT=$c: "; RTB200_LIB=$V RTB_CARVEOUT=$c RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/q_${c}_$rep.npy | show; fi
    python tools/slot_probe2.py --cmp /tmp/sp/q_serial_0.npy /tmp/sp/q_${c}_${rep}_1.npy
  done
done
