#!/bin/bash
# round 2, N GPUs (gpurun --gpus N): multi-GPU parity + feature checks, then weak / strong scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 900 > gpurun_out/mg_pytest.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/mg_pytest.log | cut -c1-200
P=29600
run() {  # name, extra env, extra args
  P=$((P+1))
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline $3 > gpurun_out/mg${N}_$1.json 2> gpurun_out/mg${N}_$1.err
  python -c "
import json
try:
    b=json.loads([l for l in open('gpurun_out/mg${N}_$1.json') if l.startswith('{')][-1]); print('$1', b['n_gpus'], b['scaling'], 'value', round(b['value']), 'e2e', round(b['e2e']['value']), 'pageable', round(b['e2e_pageable']['value']), 'ms', round(b['ms_per_step'],2), {k:round(b['roofline'][k],4) for k in ('frac','avg_launch_us')}, b['bed']['particles_total'], b['bed']['ghost_rows_rank0'])
except Exception as e: print('$1 no json', e)
"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/mg${N}_$1.err | tail -2 | cut -c1-300
}
run weak "SEDI_X=1" ""
run weak_unfused "SEDI_HALO_FUSED=0" ""
run weak_nograph "SEDI_GRAPH=0" ""
run strong "SEDI_X=1" "--scaling strong"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mg_bench_n1.json 2>/dev/null; python -c "
import json
b=json.loads([l for l in open('gpurun_out/mg_bench_n1.json') if l.startswith('{')][-1]); print('N=1', round(b['value']), 'e2e', round(b['e2e']['value']), round(b['ms_per_step'],2), b['roofline']['frac'], b['roofline']['avg_launch_us'])"
