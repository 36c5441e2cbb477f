#!/usr/bin/env bash
# round 2, GPU call K (1 GPU): k_shadow_wide with per-axis slack + one queue entry per visit; variants loaded for real
# (LD_LIBRARY_PATH so that librtb200_host.so binds to the variant as well); vote-exit variants of the binary walks
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "any_hit or (config3 and wide) or single_light" ) > gpurun_out/r2k_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2k_pytest.log | tail -8
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; k=r['k_traverse']
print('$1: value %.1f e2e %.1f ms/step %.3f | shadow %.3f ms/launch frac %.3f share %.3f | traverse %.3f ms/launch | shade share %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['ms_per_launch'], r['frac'], r['share_of_pass'], k['ms_per_launch'], r['k_shade_share_of_pass']))"; }
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
timeout 300 $B 2>/dev/null | summ "binary"
RTB_SHADOW_KERNEL=wide timeout 300 $B 2>/dev/null | summ "wide default (defer, b4, r20, 384x2)"
for n in b2 b8 b64 r1 r12 r28 t256x3 t512x2; do v=$PWD/bench_data/variants/$n; LD_LIBRARY_PATH=$v RTB200_LIB=$v/librtb200.so RTB_SHADOW_KERNEL=wide timeout 300 $B 2>/dev/null | summ "wide $n"; done
for n in tv3 tv6 tv12; do v=$PWD/bench_data/variants/$n; LD_LIBRARY_PATH=$v RTB200_LIB=$v/librtb200.so timeout 300 $B 2>/dev/null | summ "binary $n"; done
RTB_SHADOW_KERNEL=wide ncu --set full --clock-control none --import-source on -k regex:"k_shadow_wide" -s 18 -c 2 -o gpurun_out/prof_r2k \
    python tools/quick_bench.py 1920 1080 4 > gpurun_out/ncu_r2k.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2k.ncu-rep 2>&1 | grep -E "^---|time_duration|thread_inst_executed_per|issue_active|inst_executed.sum|hit_rate|long_scoreboard|warps_active"
grep -E "profiled" gpurun_out/ncu_r2k.log | tail -1
