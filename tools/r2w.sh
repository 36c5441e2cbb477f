#!/usr/bin/env bash
# round 2, GPU call W (1 GPU): Light Tracer renderer on the device
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_vcm.py -m gpu -q ) > gpurun_out/r2w_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/r2w_pytest.log | tail -14
