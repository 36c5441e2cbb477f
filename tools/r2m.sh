#!/usr/bin/env bash
# round 2, GPU call M (1 GPU): PRMT selector fix in the 8-wide node test, b2/r16 defaults, CTA-local material binning in k_shade
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2m_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2m_pytest.log | tail -12
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; k=[v for v in r.values() if isinstance(v, dict) and 'ms_per_launch' in v][0]
print('$1: value %.1f e2e %.1f %s ms/step %.3f | %s %.3f ms/launch frac %.3f share %.3f | %s %.3f ms/launch frac %.3f | shade share %.3f | %s' % (d['value'], d['e2e']['value'], ['%.3f' % x for x in d['e2e'].get('repeats_s', [])], d['ms_per_step'], r['kernel'][:14], r['ms_per_launch'], r['frac'], r['share_of_pass'], k['kernel'][:14], k['ms_per_launch'], k['frac'], r['k_shade_share_of_pass'], d.get('e2e_frame_check')))"; }
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
timeout 300 $B 2>/dev/null | summ "default"
RTB_SORT_SHADE=0 timeout 300 $B 2>/dev/null | summ "no material binning"
timeout 300 $B 2>/dev/null | summ "default again"
timeout 600 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; summ "driver-style (20 steps)" < gpurun_out/r2m_bench.json
