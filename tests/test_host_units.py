"""-m "not gpu": the __host__ __device__ helpers of the CUDA sources executed on the host (tests/host/host_units.cu, compiled with nvcc;
no GPU involved): the per-cell row sort of the fixed-order scatter (sedi_couple.cuh: fcell_shell_sort)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_host_device_helpers_on_the_host(tmp_path):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc) and not shutil.which("nvcc"):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "host_units")
    r = subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, os.path.join(HERE, "host", "host_units.cu")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "shell sort ok" in r.stdout, r.stdout[-2000:]
