#!/bin/bash
# round 2, 2 GPUs: whole GPU suite (multi-GPU parity + feature checks included), then configs[3] / configs[4] on 2 GPUs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/mf_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/mf_pytest.log | cut -c1-300
P=29700
for c in 3 4; do
  P=$((P+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --config $c --steps 5 --warmup 3 --ramp 3 --no-cpu-baseline > gpurun_out/mf_bench_cfg$c.json 2> gpurun_out/mf_bench_cfg$c.err; echo "bench cfg$c rc=$?"
  python -c "
import json
try:
    b=json.loads([l for l in open('gpurun_out/mf_bench_cfg$c.json') if l.startswith('{')][-1]); print('cfg$c', b['n_gpus'], 'value', round(b['value']), 'e2e', round(b['e2e']['value']), 'ms', round(b['ms_per_step'],2), {k:b['roofline'][k] for k in ('frac','avg_launch_us')}, b['bed']['particles_total'], b['bed']['ghost_rows_rank0'])
except Exception as e: print('no json', e)
"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/mf_bench_cfg$c.err | tail -4 | cut -c1-300
done
