#!/usr/bin/env bash
# round 2, GPU call AE (1 GPU): section 4c - checksums of what k_tail reads and finds, serial against overlapped
set -u
mkdir -p gpurun_out /tmp/sp
P="python tools/slot_probe2.py"
V=$PWD/bench_data/variants/tail_probe.so
for sc in mesh_small cornell; do
echo "=== $sc  (numRayBoxTests = paths k_tail processed, numRayTriangleTests = checksum of their input state, numShadowTriangleTests = of their pass colour on entry, numShadowBoxTests = of their first hits)"
echo "--- serial";                 RTB200_LIB=$V $P $sc /tmp/sp/p_serial.npy | tail -1 | cut -c1-400
echo "--- serial again";           RTB200_LIB=$V $P $sc /tmp/sp/p_serial2.npy | tail -1 | cut -c1-400
echo "--- 4 slots + k_tail";       RTB200_LIB=$V RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/p_s4t.npy | tail -1 | cut -c1-400
echo "--- 4 slots + k_tail again"; RTB200_LIB=$V RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/p_s4t2.npy | tail -1 | cut -c1-400
echo "--- 4 slots + k_tail, slots serialised by events"; RTB200_LIB=$V RTB_SERIALIZE_SLOTS=1 RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/p_s4ts.npy | tail -1 | cut -c1-400
python tools/slot_probe2.py --cmp /tmp/sp/p_serial_0.npy /tmp/sp/p_s4t_0.npy
done
