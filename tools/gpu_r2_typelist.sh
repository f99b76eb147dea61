#!/bin/bash
# round 2: the type-list configurations (configs[3] cohesive, configs[4] lubrication) + parity tests
mkdir -p gpurun_out
for c in 3 4; do SEDI_BENCH_TRACE=100 timeout 240 python bench.py --config $c --steps 10 --warmup 3 --ramp 5 --no-cpu-baseline > gpurun_out/t_bench_cfg$c.json 2> gpurun_out/t_bench_cfg$c.err; echo "bench cfg$c rc=$?"; python -c "
import json,sys
try:
    b=json.load(open('gpurun_out/t_bench_cfg$c.json')); print(b['value'], {k:b['roofline'][k] for k in ('frac','avg_launch_us')}, {k:b['bed'][k] for k in ('particles_total','pairs_per_particle','touching_pairs_per_particle','ell_width','neighbor_rebuilds_in_timed_region')}, b['ms_per_step'])
except Exception as e: print('no json', e)
"; tail -2 gpurun_out/t_bench_cfg$c.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 60 -c 1 -f -o gpurun_out/prof_t_cfg4 python tools/kstep_sweep.py --one --config 4 --bed random --steps 1 --warm 0 --substeps 80 > gpurun_out/t_ncu4.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_t_cfg4.ncu-rep > gpurun_out/t_cfg4_ncu_full.txt 2>&1; head -24 gpurun_out/t_cfg4_ncu_full.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_pytest.log
