"""The Foam-side C++ driver (tests/host/foam_side_driver.cpp) uses the boundary exactly like softParticleCloud does:
`new LAMMPS(0,NULL,comm)`, `lmp_->input->one(line)`, the library.h functions with the LAMMPS* handle, `delete lmp_`
(reference: lammpsFoam/softParticleCloud.C:57-206, :838-922), compiled against include/lammps_shim/ and linked to
libsedi_b200.so; plus the C++ mirror of enhancedCloud's public interface (include/sedi_cloud.hpp)."""
import os
import subprocess

import numpy as np
import pytest

import sedifoam_b200 as sb
from sedifoam_b200 import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    sb.build_library()
    out = str(tmp_path_factory.mktemp("host") / "foam_side_driver")
    libdir = os.path.dirname(sb.library_path())
    subprocess.run(["g++", "-std=c++11", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "include", "lammps_shim"),
                    os.path.join(ROOT, "tests", "host", "foam_side_driver.cpp"), "-o", out, "-L", libdir, "-lsedi_b200",
                    "-Wl,-rpath," + libdir], check=True)
    return out


def _run(driver, args, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run([driver] + [str(a) for a in args], capture_output=True, text=True, env=e, timeout=170)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_foam_side_compiles_and_reads_the_script(driver, tmp_path):
    """host-only part of initLammps: script fed line by line, atom table served back (no GPU needed)"""
    case = cases.fluidized_bed(dims=(4, 5, 3))
    script = cases.write_lammps_files(case, str(tmp_path))
    out = _run(driver, [script, 0, 1, "info"])
    rows = np.array([[float(v) for v in ln.split()[1:]] for ln in out.splitlines() if ln.startswith("I ")])
    assert len(rows) == 60 and "nGlobal 60 nLocal 60 dt 1.9999999999999999e-06" in out
    assert np.array_equal(rows[:, 0].astype(int), case["tag"])
    assert np.array_equal(rows[:, 4:7], case["x"]) and np.array_equal(rows[:, 2], case["diam"])


@pytest.mark.gpu
def test_put_step_get_sequence_matches_oracle(driver, tmp_path, oracle_mod):
    case = cases.fluidized_bed(dims=(8, 9, 8), vjit=0.0)   # a read_data file carries no velocities
    script = cases.write_lammps_files(case, str(tmp_path))
    acc = (0.5, 3.0, -0.25)
    out = _run(driver, [script, 3, 50, "host"] + list(acc))
    rows = np.array([[float(v) for v in ln.split()[1:]] for ln in out.splitlines() if ln.startswith("P ")])
    o = np.argsort(rows[:, 0])
    rows = rows[o]
    ora = oracle_mod.Oracle("port")
    cases.apply(case, ora)
    ora.setup()
    m = case["rho"] * np.pi / 6.0 * case["diam"] ** 3
    for _ in range(3):
        ora.put_fdrag(m[:, None] * np.asarray(acc)[None, :], case["tag"])
        ora.run(50)
    a = ora.atoms()
    L = np.abs(case["box_hi"] - case["box_lo"]).max()
    assert np.abs(rows[:, 1:4] - a["x"]).max() / L < 1e-6
    assert np.abs(rows[:, 4:7] - a["v"]).max() / np.abs(a["v"]).max() < 1e-6


@pytest.mark.gpu
def test_enhanced_cloud_mirror(driver, tmp_path):
    """evolve() + calcTcFields() through include/sedi_cloud.hpp against the same sequence through the ctypes binding"""
    case = cases.fluidized_bed(dims=(8, 9, 8), vjit=0.0)
    script = cases.write_lammps_files(case, str(tmp_path))
    env = {"MESH_X0": case["mesh_lo"][0], "MESH_Y0": case["mesh_lo"][1], "MESH_Z0": case["mesh_lo"][2],
           "MESH_X1": case["mesh_hi"][0], "MESH_Y1": case["mesh_hi"][1], "MESH_Z1": case["mesh_hi"][2],
           "MESH_NX": case["mesh_n"][0], "MESH_NY": case["mesh_n"][1], "MESH_NZ": case["mesh_n"][2]}
    env = {k: repr(float(v)) if "N" not in k[5:] else str(int(v)) for k, v in env.items()}
    Uf = (0.0, 0.05, 0.0)
    out = _run(driver, [script, 2, 50, "cloud"] + list(Uf), env)
    rows = np.array([[float(v) for v in ln.split()[1:]] for ln in out.splitlines() if ln.startswith("C ")])
    e = sb.Lammps()
    cases.apply(case, e)
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.coupling_config(sb.DRAG_ERGUN_WENYU, sb.FORCE_DRAG | sb.FORCE_PGRAD, 1e-6, 1000.0, (0, -9.8, 0), 50 * 2e-6)
    e.setup()
    g, Ue = e.scatter_alpha_u()
    C = len(g)
    Ufc = np.tile(Uf, (C, 1)); gradp = np.tile([0.0, -9800.0, 0.0], (C, 1))
    for _ in range(2):
        e.put_cell_fields(Ufc, None, gradp)
        e.compute_fluid_force(); e.sedi_step(50); e.locate()
        g, Ue = e.scatter_alpha_u()
        A, Om = e.calc_tc()
    assert rows.shape == (C, 5)
    assert np.abs(rows[:, 1] - g).max() <= 1e-12 * np.abs(g).max()
    assert np.abs(rows[:, 2:5] - A).max() <= 1e-9 * np.abs(A).max()
    assert g.sum() > 0 and np.abs(A).max() > 0
