#!/usr/bin/env bash
# round 2, GPU call AK (1 GPU): final state - GPU suite, smoke, bench line
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2ak_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR|^real" gpurun_out/r2ak_pytest.log | tail -8
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2ak_bench.json') if l.startswith('{')][-1]); r=d['roofline']
print('bench: value %.1f e2e %.1f ms/step %.3f steps %d launches %d | frac %.3f share %.3f | trav frac %.3f | shade share %.3f | %s | clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['steps'], d['gpu_launches'], r['frac'], r['share_of_pass'], r['k_traverse']['frac'], r['k_shade_share_of_pass'], d.get('e2e_frame_check'), d['clocks']))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
