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
