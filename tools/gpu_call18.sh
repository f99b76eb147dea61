#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
timeout 300 python tools/kstep_sweep.py --libs sedifoam_b200/libsedi_b200.so --envs 'ell=' 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['lib'], r['env'], round(r.get('kstep_us', -1), 1), r.get('error', '')[:200])"
