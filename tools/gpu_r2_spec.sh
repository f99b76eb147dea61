#!/bin/bash
# round 2: sweep of k_step_sell build variants (build_variants/*.so, see tools/build_variant.sh) on the settled bed
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1200 python tools/kstep_sweep.py --bed settled --nocouple --steps 10 --warm 5 --envs 'sell=' --out gpurun_out/spec_sweep.json > gpurun_out/spec_sweep.log 2>&1
grep -h '^{' gpurun_out/spec_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['lib'].split('/')[-1], r['env'], round(r.get('kstep_us', -1), 1), round(r.get('GBps_alg', 0)), r.get('rebuilds'), r.get('state_sha'), r.get('error', '')[:300])"
if [ "$SEDI_SWEEP_TESTS" = "1" ]; then timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/spec_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/spec_pytest.log | cut -c1-300; fi
