#!/bin/bash
# what the driver runs at round end, plus the ncu evidence for profiles/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/re_pytest.log 2>&1; tail -2 gpurun_out/re_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/re_smoke.log 2>&1; tail -1 gpurun_out/re_smoke.log
timeout 900 python bench.py > gpurun_out/re_bench.json 2> gpurun_out/re_bench.err; cat gpurun_out/re_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/re_bench_ref.json 2> gpurun_out/re_bench_ref.err; cat gpurun_out/re_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/re_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/re_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 120 -c 1 -f -o gpurun_out/prof_re_kstep python tools/kstep_sweep.py --one --steps 1 --warm 1 > gpurun_out/re_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_re_kstep.ncu-rep > gpurun_out/re_kstep_ncu_full.txt 2>&1
head -22 gpurun_out/re_kstep_ncu_full.txt
