#!/usr/bin/env bash
# round 2, GPU call L (1 GPU): new defaults (k_shadow_wide deferred + per-axis slack + PRMT fix + division-free sure accept,
# vote exit 6 in the binary walks): whole GPU suite, bench A/B of the leaf bias / refill / vote variants, ncu
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2l_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2l_pytest.log | tail -12
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; k=r['k_traverse']
print('$1: value %.1f e2e %.1f ms/step %.3f | %s %.3f ms/launch frac %.3f share %.3f | traverse %.3f ms/launch frac %.3f | shade share %.3f | %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel'][:14], r['ms_per_launch'], r['frac'], r['share_of_pass'], k['ms_per_launch'], k['frac'], r['k_shade_share_of_pass'], d.get('e2e_frame_check')))"; }
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
timeout 300 $B 2>/dev/null | summ "default (wide b4 r20, vote 6)"
RTB_SHADOW_KERNEL=simple timeout 300 $B 2>/dev/null | summ "binary any-hit, vote 6"
for n in b1 b2 b3 b2r16 b2r24 tv0 tv8; do v=$PWD/bench_data/variants/$n; LD_LIBRARY_PATH=$v RTB200_LIB=$v/librtb200.so timeout 300 $B 2>/dev/null | summ "$n"; done
ncu --set full --clock-control none --import-source on -k regex:"k_shadow_wide|k_traverse|k_shade" -s 54 -c 6 -o gpurun_out/prof_r2l \
    python tools/quick_bench.py 1920 1080 4 > gpurun_out/ncu_r2l.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2l.ncu-rep 2>&1 | grep -E "^---|time_duration|thread_inst_executed_per|issue_active|inst_executed.sum|registers"
grep -E "profiled" gpurun_out/ncu_r2l.log | tail -1
