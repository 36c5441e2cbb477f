#!/usr/bin/env bash
# round 2, GPU call AA (1 GPU): `--set full` capture with source of the depth-1 launches (bounce rays) of the three hot kernels
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_traverse_p|k_shadow_wide|k_shade" -s 57 -c 3 -o gpurun_out/prof_r2aa \
    python tools/quick_bench.py 1920 1080 4 > gpurun_out/ncu_r2aa.log 2>&1
tail -2 gpurun_out/ncu_r2aa.log; ls -la gpurun_out | tail -3
