#!/bin/bash
# round 2: quick sweep of k_step_sell build variants on the settled bed (+ configs 3 / 4 + parity tests of the in-tree build)
mkdir -p gpurun_out
timeout 900 python tools/kstep_sweep.py --bed settled --envs 'sell=' --out gpurun_out/sweep2_settled.json > gpurun_out/sweep2_settled.log 2>&1
grep -h '^{' gpurun_out/sweep2_settled.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r.get('bed'), r['lib'], r['env'], round(r.get('kstep_us', -1), 1), round(r.get('GBps_alg', 0)), r.get('pairs_per_particle'), r.get('touching_pairs_per_particle'), r.get('state_sha'), r.get('error', '')[:300])"
for c in 3 4; do timeout 240 python bench.py --config $c --steps 10 --warmup 3 --ramp 5 --no-cpu-baseline > gpurun_out/s2_bench_cfg$c.json 2> gpurun_out/s2_bench_cfg$c.err; echo "bench cfg$c rc=$?"; python -c "
import json
try:
    b=json.loads([l for l in open('gpurun_out/s2_bench_cfg$c.json') if l.startswith('{')][-1]); print(round(b['value']), {k:b['roofline'][k] for k in ('frac','avg_launch_us')}, b['ms_per_step'])
except Exception as e: print('no json', e)
"; done
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s2_pytest.log
