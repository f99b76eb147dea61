#!/bin/bash
# round 2, 2 GPUs: where the multi-GPU sub-step time goes (SEDI_PROF_SPLIT: sub-step kernel / ghost exchange + barrier), border rows clustered or not, push fused or not
N=${1:-2}
mkdir -p gpurun_out
P=29700
run() {  # name, extra env, extra args
  P=$((P+1))
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline $3 > gpurun_out/sg${N}_$1.json 2> gpurun_out/sg${N}_$1.err
  python -c "
import json
try:
    b=json.loads([l for l in open('gpurun_out/sg${N}_$1.json') if l.startswith('{')][-1]); print('$1', b['n_gpus'], b['scaling'], 'value', round(b['value']), 'e2e', round(b['e2e']['value']), 'ms', round(b['ms_per_step'],2), {k:round(b['roofline'][k],4) for k in ('frac','avg_launch_us')}, b['bed']['ghost_rows_rank0'], b['gpu_launches'])
except Exception as e: print('$1 no json', e)
"; grep "prof split" gpurun_out/sg${N}_$1.err | head -4; grep -v "^\*\*\*\|OMP_NUM\|prof split" gpurun_out/sg${N}_$1.err | tail -2 | cut -c1-300
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/sg_check.log 2>&1; echo "mgpu_check rc=$?"; grep -c "\-> OK" gpurun_out/sg_check.log; grep "FAIL" gpurun_out/sg_check.log | head
run cluster "SEDI_X=1" ""
run nocluster "SEDI_BORDER_CLUSTER_OFF=1" ""
run cluster_split "SEDI_PROF_SPLIT=1" ""
run nocluster_split "SEDI_BORDER_CLUSTER_OFF=1 SEDI_PROF_SPLIT=1" ""
run unfused_split "SEDI_HALO_FUSED=0 SEDI_PROF_SPLIT=1" ""
SEDI_PROF_SPLIT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/sg_bench_n1.json 2> gpurun_out/sg_bench_n1.err; grep "prof split" gpurun_out/sg_bench_n1.err | head -3
