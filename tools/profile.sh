#!/usr/bin/env bash
# Run under gpurun: bench line, ncu launch list of the same command, one `--set full` capture of the traversal kernels.
#   tools/profile.sh <tag>
set -u
TAG="${1:-r1}"
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_$TAG.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_$TAG.log | tail -8
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json
# launch list (times are cold-cache and serialised: only the SHARES are comparable with the bench)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# full capture: first wavefront of the 1080p batch (depth 0 and depth 1 of the closest-hit, shading and any-hit kernels); the
# 480x270 parity render in front launches 2 passes x 9 depths x 3 kernels = 54 matching kernels
ncu --set full --clock-control none --import-source on -k regex:"k_traverse|k_shadow|k_shade" -s 54 -c 6 -o gpurun_out/prof_$TAG \
    python tools/quick_bench.py 1920 1080 4 > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -8
# DRAM traffic of every k_traverse / k_shadow launch of ONE 8-pass wavefront (9 depths): bench.py's roofline.traffic
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_traverse|k_shadow" -s 36 -c 18 \
    --csv --log-file gpurun_out/dram_$TAG.csv python tools/quick_bench.py 1920 1080 8 > gpurun_out/dram_$TAG.log 2>&1
python tools/dram_per_ray.py gpurun_out/dram_$TAG.csv gpurun_out/dram_$TAG.log $TAG > gpurun_out/dram_per_ray_$TAG.json
cp profiles/traverse_dram_bytes.json gpurun_out/traverse_dram_bytes_$TAG.json
