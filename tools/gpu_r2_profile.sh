#!/bin/bash
# round 2: bench lines (both arms) + launch list + full ncu capture of the default kernel, for profiles/r02_*
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p2_smoke.log 2>&1; tail -1 gpurun_out/p2_smoke.log
timeout 900 python bench.py > gpurun_out/p2_bench.json 2> gpurun_out/p2_bench.err; cat gpurun_out/p2_bench.json; tail -2 gpurun_out/p2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p2_bench_ref.json 2> gpurun_out/p2_bench_ref.err; cat gpurun_out/p2_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p2_launches.csv python bench.py --steps 1 --warmup 1 --ramp 0 --no-cpu-baseline > gpurun_out/p2_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 120 -c 1 -f -o gpurun_out/prof_p2_kstep python tools/kstep_sweep.py --one --steps 1 --warm 1 > gpurun_out/p2_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_p2_kstep.ncu-rep > gpurun_out/p2_kstep_ncu_full.txt 2>&1
head -24 gpurun_out/p2_kstep_ncu_full.txt
