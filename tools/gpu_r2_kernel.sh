#!/bin/bash
# round 2: settle the benchmark column, then sweep + bench + ncu + parity tests of the warp-queue kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
if [ "$SEDI_RESETTLE" = "1" ]; then SEDI_KSTEP_PATH=ell timeout 900 python tools/make_settled_column.py --steps 600000 --chunk 50000 > gpurun_out/settle.log 2>&1; echo "settle rc=$?"; tail -4 gpurun_out/settle.log; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c1_smoke.log
timeout 900 python tools/kstep_sweep.py --bed settled --envs 'sell=' --out gpurun_out/sweep_settled.json > gpurun_out/sweep_settled.log 2>&1
timeout 400 python tools/kstep_sweep.py --bed lattice --libs sedifoam_b200/libsedi_b200.so --envs 'sell=;ell=SEDI_KSTEP_PATH=ell' --out gpurun_out/sweep_lattice.json > gpurun_out/sweep_lattice.log 2>&1
grep -h '^{' gpurun_out/sweep_settled.log gpurun_out/sweep_lattice.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r.get('bed'), r['lib'], r['env'], round(r.get('kstep_us', -1), 1), round(r.get('GBps_alg', 0)), r.get('pairs_per_particle'), r.get('touching_pairs_per_particle'), r.get('state_sha'), r.get('error', '')[:300])"
timeout 900 python bench.py --steps 50 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?"; cat gpurun_out/c1_bench.json; tail -3 gpurun_out/c1_bench.err
for c in 1 3 4; do SEDI_BENCH_TRACE=100 timeout 240 python bench.py --config $c --steps 10 --warmup 3 --ramp 5 --no-cpu-baseline > gpurun_out/c1_bench_cfg$c.json 2> gpurun_out/c1_bench_cfg$c.err; echo "bench cfg$c rc=$?"; cut -c1-300 gpurun_out/c1_bench_cfg$c.json; python -c "
import json,sys
try:
    b=json.load(open('gpurun_out/c1_bench_cfg$c.json')); print({k:b['roofline'][k] for k in ('frac','avg_launch_us','kernel')}, {k:b['bed'][k] for k in ('particles_total','pairs_per_particle','touching_pairs_per_particle','ell_width')}, b['ms_per_step'])
except Exception as e: print('no json', e)
"; tail -2 gpurun_out/c1_bench_cfg$c.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 1 --warmup 1 --ramp 0 --no-cpu-baseline > gpurun_out/c1_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 120 -c 1 -f -o gpurun_out/prof_c1_wq python tools/kstep_sweep.py --one --steps 1 --warm 1 > gpurun_out/c1_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_c1_wq.ncu-rep > gpurun_out/c1_wq_ncu_full.txt 2>&1; head -30 gpurun_out/c1_wq_ncu_full.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/c1_pytest.log
