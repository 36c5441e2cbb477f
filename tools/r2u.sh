#!/usr/bin/env bash
# round 2, GPU call U (1 GPU): what the bit-exact compile flags cost (fast-math build, timing only); CTA shapes / thresholds after the last changes
set -u
mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; k=[v for v in r.values() if isinstance(v, dict) and 'ms_per_launch' in v][0]
print('$1: value %.1f e2e %.1f ms/step %.3f | %s %.3f ms/launch share %.3f | %s %.3f ms/launch share %.3f | shade share %.3f (%.3f ms/step) | %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel'][:14], r['ms_per_launch'], r['share_of_pass'], k['kernel'][:14], k['ms_per_launch'], k['share_of_pass'], r['k_shade_share_of_pass'], r['k_shade_share_of_pass'] * d['ms_per_step'], d.get('e2e_frame_check')))"; }
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
timeout 300 $B 2>/dev/null | summ "default"
for n in fastmath t320 t256 r12 r20 b3 sm5 sm4; do v=$PWD/bench_data/variants/$n; LD_LIBRARY_PATH=$v RTB200_LIB=$v/librtb200.so timeout 300 $B 2>/dev/null | summ "$n"; done
