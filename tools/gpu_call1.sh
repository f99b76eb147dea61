#!/bin/bash
# round-1 session-3 GPU call: parity suite on the row-block kernel, then kernel variant sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 1500 python tools/kstep_sweep.py --envs 'rows=' --out gpurun_out/c1_sweep_a.json > gpurun_out/c1_sweep_a.log 2>&1
timeout 900 python tools/kstep_sweep.py --libs sedifoam_b200/libsedi_b200.so --envs 'ell=SEDI_KSTEP_PATH=ell;t4x4=SEDI_BIN_TILE=4x4;t8x8=SEDI_BIN_TILE=8x8;t8x4=SEDI_BIN_TILE=8x4;ell_t8x8=SEDI_KSTEP_PATH=ell,SEDI_BIN_TILE=8x8' --out gpurun_out/c1_sweep_b.json > gpurun_out/c1_sweep_b.log 2>&1
timeout 600 python tools/kstep_sweep.py --libs build_variants/r128.so --envs 't8x8=SEDI_BIN_TILE=8x8;t16x8=SEDI_BIN_TILE=16x8;t6x6=SEDI_BIN_TILE=6x6' --out gpurun_out/c1_sweep_c.json > gpurun_out/c1_sweep_c.log 2>&1
cat gpurun_out/c1_sweep_a.log gpurun_out/c1_sweep_b.log gpurun_out/c1_sweep_c.log | grep '^{' | cut -c1-250
