#!/usr/bin/env bash
# Development helper run under gpurun: effect of the number of wavefront slots (RTB_SLOTS) on the bench numbers.
set -u
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2 3; do python tools/slot_probe.py mesh_small 2>/dev/null | cut -c1-100 | head -4 | tr '\n' ';'; echo; done
for i in 1 2 3; do python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -1; done
for n in ${SLOT_LIST:-1 4}; do
  echo "=== RTB_SLOTS=$n"
  RTB_SLOTS=$n python bench.py --steps 64 --warmup 4 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']; print('value %.1f Mray/s  e2e %.1f  ms/step %.3f launches %d  k_shadow %.3f ms/launch frac %.3f  k_traverse %.3f ms/launch frac %.3f shade share %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], r['ms_per_launch'], r['frac'], r['k_traverse']['ms_per_launch'], r['k_traverse']['frac'], r['k_shade_share_of_pass']))
    elif l.strip(): print(l.rstrip()[:300])
"
done
