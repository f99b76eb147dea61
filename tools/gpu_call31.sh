#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
SEDI_GRAPH=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | tail -2
timeout 300 python tools/kstep_sweep.py --libs sedifoam_b200/libsedi_b200.so --steps 5 --warm 2 --envs 'graph=;nograph=SEDI_GRAPH=0' 2>&1 | grep '^{' | cut -c1-110
python bench.py --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print(r['value'], r['ms_per_step'], r['roofline']['avg_launch_us'], r['roofline']['kernel_share_of_step'], r['config']['neighbor_rebuilds_in_timed_region'], r['e2e']['value'], r['gpu_launches'])"
