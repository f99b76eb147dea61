#!/bin/bash
# round 2: what the driver runs at round end (pytest -m gpu, smoke, both bench arms) + launch list + full ncu capture + the other configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/f_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; cut -c1-400 gpurun_out/f_bench.json; tail -2 gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; cut -c1-300 gpurun_out/f_bench_ref.json
for c in 1 3 4; do timeout 400 python bench.py --config $c --steps 20 --warmup 3 --ramp 10 > gpurun_out/f_bench_cfg$c.json 2> gpurun_out/f_bench_cfg$c.err; echo "bench cfg$c rc=$?"; python -c "
import json
try:
    b=json.loads([l for l in open('gpurun_out/f_bench_cfg$c.json') if l.startswith('{')][-1]); print(round(b['value']), round(b['e2e']['value']), {k:b['roofline'][k] for k in ('frac','avg_launch_us','kernel')}, b['ms_per_step'], b['cpu_baseline'])
except Exception as e: print('no json', e)
"; tail -2 gpurun_out/f_bench_cfg$c.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 1 --ramp 0 --no-cpu-baseline > gpurun_out/f_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 120 -c 1 -f -o gpurun_out/prof_f_kstep python tools/kstep_sweep.py --one --steps 1 --warm 1 > gpurun_out/f_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_f_kstep.ncu-rep > gpurun_out/f_kstep_ncu_full.txt 2>&1
head -22 gpurun_out/f_kstep_ncu_full.txt
