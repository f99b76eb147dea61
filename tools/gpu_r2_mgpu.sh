#!/bin/bash
# round 2, N GPUs (gpurun --gpus N): multi-GPU parity under pytest, then weak / strong scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 > gpurun_out/mg_pytest.log 2>&1; echo "mgpu pytest rc=$?"; tail -12 gpurun_out/mg_pytest.log | cut -c1-400
P=29600
for mode in weak strong; do
  P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 3 --scaling $mode --no-cpu-baseline > gpurun_out/mg${N}_bench_$mode.json 2> gpurun_out/mg${N}_bench_$mode.err; echo "bench $mode rc=$?"
  python -c "
import json
try:
    b=json.load(open('gpurun_out/mg${N}_bench_$mode.json')); print('$mode', b['n_gpus'], 'value', round(b['value']), 'e2e', round(b['e2e']['value']), 'pageable', round(b['e2e_pageable']['value']), 'ms', round(b['ms_per_step'],2), {k:b['roofline'][k] for k in ('frac','avg_launch_us')}, b['bed']['particles_total'], b['bed']['ghost_rows_rank0'], b['bed']['halo'])
except Exception as e: print('no json', e)
"; tail -3 gpurun_out/mg${N}_bench_$mode.err | cut -c1-300
done
SEDI_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mg${N}_bench_nograph.json 2> gpurun_out/mg${N}_bench_nograph.err; python -c "
import json
b=json.load(open('gpurun_out/mg${N}_bench_nograph.json')); print('weak no-graph', round(b['value']), round(b['ms_per_step'],2))"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mg_bench_n1.json 2>/dev/null; python -c "
import json
b=json.load(open('gpurun_out/mg_bench_n1.json')); print('N=1', round(b['value']), 'e2e', round(b['e2e']['value']), round(b['ms_per_step'],2), b['roofline']['frac'])"
