#!/usr/bin/env bash
# round 2, GPU call AD (1 GPU): section 4c - is k_tail the victim or the aggressor of the overlap problem?
set -u
mkdir -p gpurun_out /tmp/sp
P="python tools/slot_probe2.py"
sc=mesh_small
echo "--- serial reference (1 slot, k_tail on)";                 $P $sc /tmp/sp/serial.npy | tail -1
echo "--- 4 slots, no k_tail";                     RTB_SLOTS=4 $P $sc /tmp/sp/s4.npy | tail -1
echo "--- 4 slots + k_tail";   RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/s4t.npy | tail -1
echo "--- 4 slots + k_tail, no graph"; RTB_GRAPH=0 RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/s4tng.npy | tail -1
echo "--- 2 slots + k_tail";   RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=2 $P $sc /tmp/sp/s2t.npy | tail -1
V=$PWD/bench_data/variants/tail_discard.so
echo "--- discard build, serial";                  RTB200_LIB=$V $P $sc /tmp/sp/d_serial.npy | tail -1
echo "--- discard build, 4 slots + k_tail"; RTB200_LIB=$V RTB_TAIL_WITH_SLOTS=1 RTB_SLOTS=4 $P $sc /tmp/sp/d_s4t.npy | tail -1
for r in 0 1 2; do
python tools/slot_probe2.py --cmp /tmp/sp/serial_0.npy /tmp/sp/s4_$r.npy
python tools/slot_probe2.py --cmp /tmp/sp/serial_0.npy /tmp/sp/s4t_$r.npy
python tools/slot_probe2.py --cmp /tmp/sp/serial_0.npy /tmp/sp/s4tng_$r.npy
python tools/slot_probe2.py --cmp /tmp/sp/serial_0.npy /tmp/sp/s2t_$r.npy
python tools/slot_probe2.py --cmp /tmp/sp/d_serial_0.npy /tmp/sp/d_s4t_$r.npy
done
python tools/slot_probe2.py --cmp /tmp/sp/serial_0.npy /tmp/sp/serial_2.npy
python tools/slot_probe2.py --cmp /tmp/sp/d_serial_0.npy /tmp/sp/d_serial_2.npy
echo "--- initcheck, 1 slot, k_tail on (uninitialised global memory read by any kernel)"
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python tools/slot_probe2.py $sc /tmp/sp/ic.npy 2>&1 | grep -v "^$" | head -60
