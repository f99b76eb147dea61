"""CPU: the C-ABI library loads and exports every symbol include/sedi_b200.h declares (no compute calls: no GPU here),
and the host-side script parser (the LAMMPS input-script plug-in API, style names of interfaceToLammps/style_user.h)
accepts the commands of the shipped in.lammps files."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import sedifoam_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sb.build_library()
    return sb.load_library()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sedi_b200.h")).read()
    names = set(re.findall(r"\b((?:lammps|sedi)_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 17 + 20
    for n in names:
        assert hasattr(lib, n), n
    assert names == set(sb.EXPORTED_SYMBOLS)
    # the 17 functions of the reference's interfaceToLammps/library.h:29-63
    ref = ["open", "close", "file", "command", "sync", "get_global_n", "get_initial_np", "get_initial_info", "get_local_n",
           "get_local_domain", "get_local_info", "put_local_info", "step", "set_timestep", "get_timestep", "create_particle",
           "delete_particle"]
    for r in ref:
        assert "lammps_" + r in names


def test_header_compiles_as_c():
    src = '#include "sedi_b200.h"\nint main(void){return sedi_abi_version == 0;}\n'
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"], input=src.encode(),
                       capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_host_side_queries_need_no_gpu(lib):
    """script parsing and the pre-run queries are host work; anything that computes aborts without a device"""
    assert lib.sedi_abi_version() == 1
    assert lib.sedi_device_count() >= 0
    from sedifoam_b200 import cases
    e = sb.Lammps()
    case = cases.fluidized_bed(dims=(3, 4, 5))
    cases.apply(case, e)
    n = 60
    assert e.get_global_n() == n and e.get_local_n() == n
    assert e.get_timestep() == 2.0e-6
    e.set_timestep(1.0e-6)
    assert e.get_timestep() == 1.0e-6
    info = e.get_initial_info()
    assert np.array_equal(info["tag"], case["tag"]) and np.array_equal(info["x"], case["x"])
    assert np.allclose(info["rho"], 2650.0, rtol=1e-11)
    dom = e.get_local_domain()
    assert np.allclose(dom[0::2], case["box_lo"]) and np.allclose(dom[1::2], case["box_hi"])
    e.close()


def test_compute_without_gpu_fails_loudly():
    """no CPU fallback: a compute entry point on a box without a CUDA device aborts with a message"""
    if sb.load_library().sedi_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import sedifoam_b200 as sb\nfrom sedifoam_b200 import cases\n"
            "e = sb.Lammps(); cases.apply(cases.fluidized_bed(dims=(3,3,3)), e); e.step(1)\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True)
    assert r.returncode != 0
    assert b"no CPU fallback" in r.stderr


def test_shipped_input_scripts_parse(tmp_path):
    """every command of the reference's shipped in.lammps files is accepted (fixture: the xiaocase3 script text,
    tests/golden/xiaocase3_in.lammps is a data fixture copied from cases/auto-testing/test-cases/xiaocase3)"""
    data = tmp_path / "IC_uniform.in"
    data.write_text("LAMMPS data file\n\n1 atoms\n1 atom types\n\n0 0.004 xlo xhi\n0 0.004 ylo yhi\n0 0.0005 zlo zhi\n\nAtoms\n\n"
                    "1 1 0.000083 2000 0.002 0.0019 0.00025\n")
    script = open(os.path.join(ROOT, "tests", "golden", "xiaocase3_in.lammps")).read().replace("IC_uniform.in", str(data))
    e = sb.Lammps()
    for ln in script.splitlines():
        e.command(ln)
    assert e.get_local_n() == 1
    info = e.get_initial_info()
    assert info["diam"][0] == 0.000083 and abs(info["rho"][0] - 2000) < 1e-6
    assert e.get_timestep() == 2e-7
    e.close()
