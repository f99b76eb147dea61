#!/bin/bash
# compute-sanitizer over the hot path (run on a GPU box: `gpurun -- bash tools/sanitize.sh`, or `--gpus 2` for the halo kernels).
#   memcheck  : out-of-bounds / misaligned accesses of every kernel the smoke invocation and a small bench launch
#   racecheck : shared-memory hazards (k_step_sell's list-word stash, k_step_wq's queue / panel, k_window_sort)
#   2 ranks   : k_halo_push / k_halo_signal_wait and the ctrl[0] early-exit protocol under memcheck (tests/mgpu_check.py)
# The sanitizer slows kernels ~50x: small beds only.
set -u
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
export SEDI_COLUMN=column_256x4.npz
$CS --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/san_memcheck_smoke.log
$CS --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 gpurun_out/san_racecheck_smoke.log
for path in wq ell; do
  SEDI_KSTEP_PATH=$path $CS --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_racecheck_$path.log 2>&1; echo "racecheck $path rc=$?"; tail -2 gpurun_out/san_racecheck_$path.log
done
$CS --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cohesive_opt1 or lubricate_poly or settled_random or zcylinder" > gpurun_out/san_memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"; tail -3 gpurun_out/san_memcheck_parity.log
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  $CS --tool memcheck --target-processes all --error-exitcode 3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/mgpu_check.py > gpurun_out/san_memcheck_mgpu.log 2>&1; echo "memcheck 2-rank rc=$?"; tail -6 gpurun_out/san_memcheck_mgpu.log
fi
