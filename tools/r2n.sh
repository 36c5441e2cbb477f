#!/usr/bin/env bash
# round 2, GPU call N (1 GPU): persistent closest-hit kernel (k_traverse_p) vs k_traverse
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "any_hit" ) > gpurun_out/r2n_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2n_pytest.log | tail -12
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; k=[v for v in r.values() if isinstance(v, dict) and 'ms_per_launch' in v][0]
print('$1: value %.1f e2e %.1f ms/step %.3f | %s %.3f ms/launch share %.3f | %s %.3f ms/launch share %.3f | shade share %.3f | %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['kernel'][:14], r['ms_per_launch'], r['share_of_pass'], k['kernel'][:14], k['ms_per_launch'], k['share_of_pass'], r['k_shade_share_of_pass'], d.get('e2e_frame_check')))"; }
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
timeout 300 $B 2>/dev/null | summ "default"
RTB_TRAV_KERNEL=persistent timeout 300 $B 2>/dev/null | summ "persistent traverse r12"
for n in tp4 tp8 tp20; do v=$PWD/bench_data/variants/$n; RTB_TRAV_KERNEL=persistent LD_LIBRARY_PATH=$v RTB200_LIB=$v/librtb200.so timeout 300 $B 2>/dev/null | summ "persistent traverse $n"; done
