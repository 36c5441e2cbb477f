#!/usr/bin/env bash
# round 2, GPU call F (2 GPUs): full suite incl. multi-device tests, bench N=1 / N=2 (c3), c5 N=1 / N=2
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2f_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2f_pytest.log | tail -8
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench1.json 2> gpurun_out/r2f_bench1.err; tail -c 1500 gpurun_out/r2f_bench1.json; tail -3 gpurun_out/r2f_bench1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err; tail -c 900 gpurun_out/r2f_bench2.json; tail -5 gpurun_out/r2f_bench2.err
python bench.py --config c5 --steps 16 --warmup 3 > gpurun_out/r2f_c5_1.json 2> gpurun_out/r2f_c5_1.err; tail -c 1200 gpurun_out/r2f_c5_1.json; tail -5 gpurun_out/r2f_c5_1.err
$TR bench.py --config c5 --gpus 2 --steps 16 --warmup 3 > gpurun_out/r2f_c5_2.json 2> gpurun_out/r2f_c5_2.err; tail -c 600 gpurun_out/r2f_c5_2.json; tail -5 gpurun_out/r2f_c5_2.err
