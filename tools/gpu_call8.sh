#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kstep_sweep.py --envs 'ell=' --out gpurun_out/c8_sweep.json > gpurun_out/c8_sweep.log 2>&1
grep '^{' gpurun_out/c8_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['lib'], r['env'], round(r.get('kstep_us', -1), 1), r.get('error', '')[:200])"
for v in i2_g4; do
SEDI_B200_LIB=build_variants/$v.so timeout 300 ncu --metrics l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,lts__t_bytes.sum --clock-control none -k regex:k_step --launch-skip 120 -c 1 python tools/kstep_sweep.py --one --steps 1 --warm 1 2>&1 | grep -A12 "k_step" | head -16
done
