#!/usr/bin/env bash
# round 2, GPU call G (1 GPU): full suite (graphs, adapter, slots), bench c3 + reference arm, c5, graph on/off A/B
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r2g_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2g_pytest.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench1.json 2> gpurun_out/r2g_bench1.err; tail -c 2500 gpurun_out/r2g_bench1.json; tail -3 gpurun_out/r2g_bench1.err
timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_bench64.json 2> gpurun_out/r2g_bench64.err; tail -c 300 gpurun_out/r2g_bench64.json | head -c 300; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2g_bench64.json') if l.startswith('{')][-1]); print('\n64 steps: value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
RTB_GRAPH=0 timeout 300 python bench.py --steps 64 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('no graph 64 steps: value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err; tail -c 500 gpurun_out/r2g_ref.json
timeout 300 python bench.py --config c5 --steps 16 --warmup 3 > gpurun_out/r2g_c5.json 2> gpurun_out/r2g_c5.err; tail -c 300 gpurun_out/r2g_c5.json; tail -3 gpurun_out/r2g_c5.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
