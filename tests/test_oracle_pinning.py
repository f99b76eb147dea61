"""Pins the CPU oracle ("port" restatement) to the reference's OWN unmodified plug-in sources.

oracle/_ref/libsedi_ref.so is built by oracle/Makefile from /root/reference/interfaceToLammps/*.cpp (compiled by path
against oracle/stubs/, nothing copied).  Both backends run inside the same restated LAMMPS time loop
(oracle/oracle_driver.hpp); every force law of the hot path must agree BIT FOR BIT, step after step.
Skipped where the prebuilt reference objects are absent (they are git-ignored but travel to the GPU box).
"""
import numpy as np
import pytest

from sedifoam_b200 import cases
import util
from util import make_oracle

SCENARIOS = {
    "hertz_bed_walls": lambda: cases.fluidized_bed(dims=(6, 8, 6)),
    "hertz_column_periodic": lambda: cases.sediment_column(dims=(6, 12, 6), phi=0.45, jitter_frac=0.08),
    "cohesive_opt1": lambda: cases.cohesive_shear_bed(dims=(6, 6, 6), opt=1),
    "cohesive_opt0": lambda: cases.cohesive_shear_bed(dims=(6, 6, 6), opt=0),
    "lubricate_poly": lambda: cases.poly_lubricated(dims=(6, 6, 6)),
    # in-tree Hooke walls (fix_wall_granFix.cpp:356-554) under the stock Hooke pairs, moving walls, the cylinder wall
    "hooke_history_walls": lambda: util.hooke_history_bed(dims=(6, 8, 6)),
    "hooke_walls": lambda: util.hooke_bed(dims=(6, 8, 6)),
    "wiggle_wall_hertz": lambda: util.wiggle_wall_bed(dims=(6, 8, 6)),
    "wiggle_wall_hooke": lambda: util.wiggle_wall_bed(dims=(6, 8, 6), style="hooke"),
    "shear_wall": lambda: util.shear_wall_bed(dims=(6, 8, 6)),
    "zcylinder": lambda: util.zcylinder_bed(dims=(7, 7, 6)),
    "zcylinder_rotating": lambda: util.zcylinder_bed(dims=(7, 7, 6), shear="x"),
    "zcylinder_shear_z": lambda: util.zcylinder_bed(dims=(7, 7, 6), shear="z"),
    "settled_random": lambda: util.settled_random_bed(columns=(1, 1)),
}


def _drive(o, case, nsteps=60, fdrag=True):
    n = len(case["tag"])
    rng = np.random.default_rng(7)
    o.setup()
    if fdrag:
        o.put_fdrag(rng.normal(scale=1e-7, size=(n, 3)), case["tag"][::-1].copy())
    o.run(nsteps)
    return o.atoms()


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_port_matches_reference_objects_bitwise(oracle_mod, name):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libsedi_ref.so not built (needs /root/reference)")
    case = SCENARIOS[name]()
    a = _drive(make_oracle(oracle_mod, case, "port"), case)
    b = _drive(make_oracle(oracle_mod, case, "reference"), case)
    for k in ("x", "v", "omega", "f", "torque"):
        assert np.array_equal(a[k], b[k]), "%s differs between port and reference objects in %s" % (k, name)
    assert np.abs(a["f"]).max() > 0.0


def test_history_and_lists_match_reference(oracle_mod):
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libsedi_ref.so not built (needs /root/reference)")
    case = cases.fluidized_bed(dims=(6, 8, 6))
    res = []
    for kind in ("port", "reference"):
        o = make_oracle(oracle_mod, case, kind)
        o.run(120)
        res.append((o.pairs("gran", history=True), o.wall_shear(1), o.stat("nbuilds"), o.stat("pair_evals")))
    (pa, wa, ba, ea), (pb, wb, bb, eb) = res
    assert ba == bb and ea == eb and ba >= 1
    for u, v in zip(pa, pb):
        assert np.array_equal(u, v)
    assert np.array_equal(wa, wb)
    assert pa[2].sum() > 0  # some contacts are touching


def _jd_inputs(n=20000, seed=5):
    """covers every branch of both closures: Re below / above 1000, beta below / above 0.8 and 0.85, alpha = 0 and
    alpha -> 1 (beta clipped at ROOTVSMALL), Ur = 0 (Re clipped at ROOTVSMALL)"""
    rng = np.random.default_rng(seed)
    Ur = np.concatenate([10.0 ** rng.uniform(-8, 1.5, n), [0.0, 0.0, 1e-300, 50.0]])
    alpha = np.concatenate([rng.uniform(0.0, 0.75, n), [0.0, 1.0, 0.2, 0.15]])
    alpha[: n // 10] = rng.uniform(0.0, 0.2, n // 10)
    pd = np.concatenate([10.0 ** rng.uniform(-5, -2, n), [5e-4, 5e-4, 5e-4, 5e-3]])
    return Ur, alpha, pd


@pytest.mark.parametrize("model", [0, 1])
def test_drag_closure_port_matches_reference_objects_bitwise(oracle_mod, model):
    """pins the fluid-side restatement (ora_foam_jd_*) to the reference's own ErgunWenYu.C:86-145 / SyamlalOBrien.C:85-144"""
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libsedi_ref.so not built (needs /root/reference)")
    Ur, alpha, pd = _jd_inputs()
    for nuf, rhof in ((1.0e-6, 1000.0), (1.5e-5, 1.2)):
        a = oracle_mod.jd(model, Ur, alpha, pd, nuf, rhof, kind="port")
        b = oracle_mod.jd(model, Ur, alpha, pd, nuf, rhof, kind="reference")
        assert np.array_equal(a, b, equal_nan=True)
        Re = np.maximum((1 - alpha) * Ur * pd / nuf, 1e-150)
        assert (Re > 1000).any() and (Re < 1000).any() and ((1 - alpha) <= 0.8).any() and ((1 - alpha) > 0.85).any()
