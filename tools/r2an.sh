#!/usr/bin/env bash
# round 2, GPU call AN (1 GPU): wavefront size (RTB_BATCH_PATHS) on the final kernels
set -u
summ() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$1: value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"; }
B="python bench.py --steps 64 --warmup 5 --no-cpu-baseline"
timeout 200 $B 2>/dev/null | summ "16 M paths (default)"
RTB_BATCH_PATHS=33554432 timeout 200 $B 2>/dev/null | summ "32 M paths"
