#!/bin/bash
mkdir -p gpurun_out
for v in d1 d2 d3 d4; do
echo "== $v"
SEDI_B200_LIB=build_variants/$v.so timeout 300 ncu --metrics l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_step --launch-skip 20 -c 1 python tools/kstep_sweep.py --one --steps 1 --warm 0 --substeps 40 2>&1 | grep -A14 "void k_step" | grep -v "^ *-\|Section\|Warning\|Metric Name"
done
