#!/bin/bash
# round 2, N GPUs (gpurun --gpus N): mgpu_check on N ranks + weak / strong bench lines.
# A call on N GPUs is charged N x its box time: every command below carries a SHORT timeout of its own (a hang in the first one once
# burnt 40 GPU-minutes: the outer limit had been clamped below the inner `timeout 600`).
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/mg${N}_check.log 2>&1; echo "mgpu_check rc=$?"; grep "mgpu" gpurun_out/mg${N}_check.log | cut -c1-260
P=29600
run() {  # name, extra env, extra args
  P=$((P+1))
  env $2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline $3 > gpurun_out/mg${N}_$1.json 2> gpurun_out/mg${N}_$1.err
  python -c "
import json
try:
    b=json.loads([l for l in open('gpurun_out/mg${N}_$1.json') if l.startswith('{')][-1]); print('$1', b['n_gpus'], b['scaling'], 'value', round(b['value']), 'e2e', round(b['e2e']['value']), 'pageable', round(b['e2e_pageable']['value']), 'ms', round(b['ms_per_step'],2), {k:round(b['roofline'][k],4) for k in ('frac','avg_launch_us')}, b['bed']['particles_total'], b['bed']['ghost_rows_rank0'])
except Exception as e: print('$1 no json', e)
"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/mg${N}_$1.err | tail -2 | cut -c1-300
}
run weak "SEDI_X=1" ""
run strong "SEDI_X=1" "--scaling strong"
