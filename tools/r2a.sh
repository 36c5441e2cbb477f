#!/usr/bin/env bash
# round 2, GPU call A: full gpu suite (with the new full-size live-reference tests), bench + reference arm, and the
# overlapping-wavefront probe with the L1 / shared split pinned or not
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.json
( time python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; tail -c 400 gpurun_out/r2a_ref.json; tail -3 gpurun_out/r2a_ref.err
for carve in none 100 50 0; do
  for sc in mesh_small materials; do
    for i in 1 2; do
      if [ $carve = none ]; then RTB_SLOTS=4 python tools/slot_probe.py $sc 2>&1 | head -3 | cut -c1-60 | tr '\n' ';'
      else RTB_CARVEOUT=$carve RTB_SLOTS=4 python tools/slot_probe.py $sc 2>&1 | head -3 | cut -c1-60 | tr '\n' ';'; fi
      echo " [carve=$carve scene=$sc]"
    done
  done
done 2>&1 | tee gpurun_out/r2a_slots.log
