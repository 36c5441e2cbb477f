#!/usr/bin/env bash
# round 2, GPU call AQ (1 GPU): VCM kernels as step loops with lane refill on persistent grids
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_vcm.py tests/test_furnace.py -m gpu -q -x 2>&1 | tail -2
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$1: value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"; }
for g in 0 3 6; do RTB_VCM_GRID=$g timeout 120 python bench.py --config c5 --steps 16 --warmup 3 2>/dev/null | summ "c5, RTB_VCM_GRID=$g"; done
