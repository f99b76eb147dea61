#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kstep_sweep.py --nocouple --envs 'rows=;ell=SEDI_KSTEP_PATH=ell' --out gpurun_out/c2_sweep.json > gpurun_out/c2_sweep.log 2>&1
grep '^{' gpurun_out/c2_sweep.log | cut -c1-260
for v in rows ell; do
  if [ $v = ell ]; then export SEDI_KSTEP_PATH=ell; K=k_step; else unset SEDI_KSTEP_PATH; K=k_step_rows; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 120 -c 1 -f -o gpurun_out/prof_c2_$v python tools/kstep_sweep.py --one --steps 1 --warm 1 > gpurun_out/c2_ncu_$v.log 2>&1
  python tools/ncu_summary.py gpurun_out/prof_c2_$v.ncu-rep > gpurun_out/c2_ncu_$v.txt 2>&1
done
head -40 gpurun_out/c2_ncu_rows.txt
