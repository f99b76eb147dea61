#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py > gpurun_out/mg2_check.log 2>&1; echo "mgpu_check rc=$?"; tail -8 gpurun_out/mg2_check.log
SEDI_KSTEP_PATH=rows timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/mgpu_check.py > gpurun_out/mg2_check_rows.log 2>&1; echo "mgpu_check rows rc=$?"; tail -5 gpurun_out/mg2_check_rows.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/mg2_bench.json 2> gpurun_out/mg2_bench.err; cat gpurun_out/mg2_bench.json | cut -c1-400
