#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kstep_sweep.py --envs 'ell=' --out gpurun_out/sweep.json > gpurun_out/sweep.log 2>&1
grep '^{' gpurun_out/sweep.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['lib'], r['env'], round(r.get('kstep_us', -1), 1), r.get('error', '')[:200])"
