#!/bin/bash
# round 2: quick sweep of k_step_sell build variants on the settled bed (+ parity tests of the in-tree build)
mkdir -p gpurun_out
timeout 900 python tools/kstep_sweep.py --bed settled --envs 'sell=' --out gpurun_out/sweep2_settled.json > gpurun_out/sweep2_settled.log 2>&1
grep -h '^{' gpurun_out/sweep2_settled.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r.get('bed'), r['lib'], r['env'], round(r.get('kstep_us', -1), 1), round(r.get('GBps_alg', 0)), r.get('pairs_per_particle'), r.get('touching_pairs_per_particle'), r.get('state_sha'), r.get('error', '')[:300])"
SEDI_BENCH_VERBOSE=1 SEDI_BENCH_TRACE=40 timeout 200 python bench.py --config 4 --size 0.4 --steps 3 --warmup 1 --ramp 2 --no-cpu-baseline > gpurun_out/s2_cfg4_small.json 2> gpurun_out/s2_cfg4_small.err; echo "cfg4 small rc=$?"; tail -25 gpurun_out/s2_cfg4_small.err | cut -c1-200; cut -c1-300 gpurun_out/s2_cfg4_small.json
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s2_pytest.log
