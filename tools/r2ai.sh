#!/usr/bin/env bash
# round 2, GPU call AI (1 GPU): k_tail in a translation unit of its own at ptxas -O1: overlap probe (against the same split build
# with the tail at -O3 as the control), the whole GPU suite, what the -O1 tail costs (bench, 1/8-shard probe)
set -u
mkdir -p gpurun_out /tmp/sp
P="python tools/slot_probe2.py"
show() { tail -1 | sed -e "s/.*numRays': \([0-9]*\).*numShadowBoxTests': \([0-9]*\).*/numRays \1 first-hit checksum \2/"; }
for sc in mesh_small textures; do
for v in probe_tailO1 probe_tailO3; do
  V=$PWD/bench_data/variants/$v.so
  echo "=== $sc $v"
  echo -n "serial: "; RTB200_LIB=$V $P $sc /tmp/sp/r_${v}_serial.npy | show
  for rep in 1 2 3 4; do
    echo -n "4 slots (k_tail in the chain): "; RTB200_LIB=$V RTB_SLOTS=4 $P $sc /tmp/sp/r_${v}_$rep.npy | show
    for k in 0 1 2; do python tools/slot_probe2.py --cmp /tmp/sp/r_${v}_serial_0.npy /tmp/sp/r_${v}_${rep}_$k.npy; done | sed 's/.*npy: //' | tr '\n' ';'; echo
  done
done
done
( time timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2ai_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR|^real" gpurun_out/r2ai_pytest.log | tail -8
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']
print('$1: value %.1f e2e %.1f ms/step %.3f | %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d.get('e2e_frame_check')))"; }
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | summ "bench 20 steps"
RTB_SLOTS=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | summ "bench 20 steps, 2 slots"
python tools/shard_probe.py 2>&1 | tail -12
