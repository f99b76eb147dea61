#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/c4_pytest.log 2>&1; tail -4 gpurun_out/c4_pytest.log
SEDI_KSTEP_PATH=rows timeout 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/c4_pytest_rows.log 2>&1; tail -2 gpurun_out/c4_pytest_rows.log
timeout 900 python tools/kstep_sweep.py --envs 'ell=' --out gpurun_out/c4_sweep.json > gpurun_out/c4_sweep.log 2>&1
timeout 300 python tools/kstep_sweep.py --libs sedifoam_b200/libsedi_b200.so --envs 'rows=SEDI_KSTEP_PATH=rows' --out gpurun_out/c4_sweep_rows.json >> gpurun_out/c4_sweep.log 2>&1
grep '^{' gpurun_out/c4_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['lib'], r['env'], round(r.get('kstep_us', -1), 1), r.get('error', '')[:200])"
