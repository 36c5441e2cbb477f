#!/usr/bin/env bash
# round 2, GPU call AC (N GPUs): multi-GPU tests and the driver-style scaling run on the final kernels
set -u
mkdir -p gpurun_out
N=${1:-2}
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "multi or scattered or shard" ) > gpurun_out/r2ac_pytest_$N.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2ac_pytest_$N.log | tail -8
for K in 20 64; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps $K --warmup 5 > gpurun_out/r2ac_scale${K}_$N.json 2> gpurun_out/r2ac_scale${K}_$N.err
python -c "
import json
f='gpurun_out/r2ac_scale${K}_$N.json'
d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, d['n_gpus'], d['steps'], 'value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['e2e'].get('repeats_s'), d.get('e2e_frame_check'))
" || tail -5 gpurun_out/r2ac_scale${K}_$N.err
done
