#!/bin/bash
mkdir -p gpurun_out
export SEDI_KSTEP_PATH=ell
timeout 900 python tools/kstep_sweep.py --envs 'ell=' --out gpurun_out/c3_sweep.json > gpurun_out/c3_sweep.log 2>&1
grep '^{' gpurun_out/c3_sweep.log | cut -c1-200
