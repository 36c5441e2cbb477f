#!/usr/bin/env bash
# Development helper run under gpurun: correctness under overlapping wavefronts + speed for every library in bench_data/variants.
set -u
for v in bench_data/variants/*.so; do
  echo "=== $v"
  export RTB200_LIB=$v
  for i in 1 2 3; do python tools/slot_probe.py mesh_small 2>/dev/null | cut -c1-60 | head -4 | tr '\n' ';'; echo; done
  for i in 1 2; do python tools/slot_probe.py textures 2>/dev/null | cut -c1-60 | head -4 | tr '\n' ';'; echo; done
  python tools/quick_bench.py 1920 1080 16 2>&1 | tail -2 | cut -c1-330
done
