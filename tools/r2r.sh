#!/usr/bin/env bash
# round 2, GPU call R (8 GPUs): driver-style scaling run with the new kernels (20 steps, 5 warm-up)
set -u
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2r_scale_$N.json 2> gpurun_out/r2r_scale_$N.err
tail -c 2500 gpurun_out/r2r_scale_$N.json; tail -3 gpurun_out/r2r_scale_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 64 --warmup 5 > gpurun_out/r2r_scale64_$N.json 2> gpurun_out/r2r_scale64_$N.err
python -c "
import json
for f in ('gpurun_out/r2r_scale_$N.json','gpurun_out/r2r_scale64_$N.json'):
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, d['n_gpus'], d['steps'], 'value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d.get('e2e_frame_check'))
"
