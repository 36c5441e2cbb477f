#!/usr/bin/env bash
# round 2, GPU call Q (1 GPU): one-byte occlusion flag; k_shadow_wide at 416 / 448 / 512 threads per CTA
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "any_hit or config3 or sum_buffer or batched" ) > gpurun_out/r2q_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2q_pytest.log | tail -12
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; k=[v for v in r.values() if isinstance(v, dict) and 'ms_per_launch' in v][0]
print('$1: value %.1f e2e %.1f ms/step %.3f | %s %.3f ms/launch share %.3f | %s %.3f ms/launch share %.3f | shade share %.3f | %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel'][:14], r['ms_per_launch'], r['share_of_pass'], k['kernel'][:14], k['ms_per_launch'], k['share_of_pass'], r['k_shade_share_of_pass'], d.get('e2e_frame_check')))"; }
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
timeout 300 $B 2>/dev/null | summ "default 384x2"
for n in t416 t448 t512; do v=$PWD/bench_data/variants/$n; LD_LIBRARY_PATH=$v RTB200_LIB=$v/librtb200.so timeout 300 $B 2>/dev/null | summ "$n"; done
