#!/usr/bin/env bash
# round 2, GPU call H (1 GPU): suite; bench with the restored binary any-hit kernel vs k_shadow_wide; small-shard graph A/B
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2h_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2h_pytest.log | tail -12
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); r=d['roofline']; k=r['k_traverse']
print('$1: value %.1f e2e %.1f ms/step %.3f launches %d | shadow %.3f ms/launch frac %.3f share %.3f | traverse %.3f ms/launch frac %.3f | shade share %.3f | check %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], r['ms_per_launch'], r['frac'], r['share_of_pass'], k['ms_per_launch'], k['frac'], r['k_shade_share_of_pass'], d.get('e2e_frame_check')))"; }
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench1.json 2> gpurun_out/r2h_bench1.err; summ "binary 20" < gpurun_out/r2h_bench1.json; tail -2 gpurun_out/r2h_bench1.err
timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2>/dev/null | summ "binary 64"
RTB_SHADOW_KERNEL=wide timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2>/dev/null | summ "wide 64"
RTB_GRAPH=0 timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2>/dev/null | summ "binary 64 nograph"
for g in 1 0; do RTB_GRAPH=$g timeout 200 python tools/shard_probe.py 8 20 2>&1 | tail -1; done
for g in 1 0; do RTB_GRAPH=$g RTB_MAX_BATCH=10 timeout 200 python tools/shard_probe.py 8 20 2>&1 | tail -1; done
RTB_TAIL_PATHS=0 timeout 200 python tools/shard_probe.py 8 20 2>&1 | tail -1
