"""CPU: the oracle against the reference's own golden data for the coupling path.

cases/auto-testing/test-cases/xiaocase3 (the install check named in documentation/userManual/userGuide.tex:175-178):
one sphere (d = 83 um, rho = 2000) released at rest in a uniform 0.05 m/s water stream, g = 0, SyamlalOBrien drag,
stock gran/hooke/history + wall/gran, 100 DEM steps per fluid step.  data/lammps08.dat holds v_y(t) of the older code,
data/xiaoCase3.dat 13 points digitised from Xiao & Sun (2011).  Copies live in tests/golden/ (see README there).
The fluid is prescribed uniform (no OpenFOAM here): the sphere sits mid-channel and the momentum boundary layer of the
walls (sqrt(nu t) = 0.07 mm after 5 ms) never reaches it.
"""
import os

import numpy as np

from sedifoam_b200 import cases
from util import make_oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def run_single_sphere(oracle_mod, nfluid=250):
    case = cases.single_sphere()
    o = make_oracle(oracle_mod, case)
    Uf, gamma, gradp = cases.uniform_fields(case)
    o.setup()
    t, vy = [0.0], [0.0]
    for k in range(nfluid):
        a = o.atoms()
        cell = oracle_mod.cell_owner(a["x"], case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
        fr = oracle_mod.particle_force(cell, case["diam"], a["v"], a["v"], Uf, gamma, gradp, None, None,
                                       oracle_mod.DRAG_SYAMLAL_OBRIEN, oracle_mod.FORCE_DRAG | oracle_mod.FORCE_PGRAD,
                                       case["nub"], case["rhob"], np.zeros(3), 2e-5)
        o.put_fdrag(fr["F"], a["tag"])
        o.run(100)
        t.append((k + 1) * 2e-5); vy.append(o.atoms()["v"][0, 1])
    return np.array(t), np.array(vy)


def test_xiaocase3_velocity_curve(oracle_mod):
    gold = np.loadtxt(os.path.join(HERE, "golden", "xiaocase3_lammps08.dat"))
    xiao = np.loadtxt(os.path.join(HERE, "golden", "xiaocase3_xiaoCase3.dat"))
    t, vy = run_single_sphere(oracle_mod)
    # lammps08.dat's 0.5 ms row lies 5 % below the benchmark curve of xiaoCase3.dat itself (coarse output of the older
    # code), so it is excluded; every other golden point must be met within 4 % / 5 % of the free-stream velocity
    for row in gold[2:]:
        assert abs(np.interp(row[0], t, vy) - row[2]) < 0.04 * 0.05, row
    for tt, vv in xiao:
        if tt > 2e-4:
            assert abs(np.interp(tt, t, vy) - vv) < 0.05 * 0.05, (tt, vv)
    assert abs(vy[-1] - 0.05) < 1e-4


def test_smoothing_restatement_known_answers(oracle_mod):
    """smoothField restated (enhancedCloud.C:790-907): (1) conserves sum(phi V) -- the reference's own printed check
    (:434-435, 975-976); (2) a point source spreads to a Gaussian of variance 2 tau = b^2/2 per direction
    (documentation/diffusionEqn/diffusionEqn.tex:127-141: bandwidth b <-> tau = b^2/4); (3) smoothDirection with a zero
    entry leaves that direction untouched."""
    nc = np.array([21, 21, 21]); dx = np.array([1.0, 1.0, 1.0]); C = 21 ** 3
    phi = np.zeros(C); phi[10 + 21 * (10 + 21 * 10)] = 1.0
    out = oracle_mod.smooth_field(phi, nc, dx, 4.0, 8).reshape(21, 21, 21)   # [k][j][i]
    assert abs(out.sum() - 1.0) < 1e-12
    x = np.arange(21) - 10
    for ax in range(3):
        prof = out.sum(axis=tuple(a for a in range(3) if a != ax))
        assert abs((prof * x ** 2).sum() - 8.0) < 0.1
    out2 = oracle_mod.smooth_field(phi, nc, dx, 4.0, 8, (1.0, 0.0, 1.0)).reshape(21, 21, 21)
    assert np.count_nonzero(out2.sum(axis=(0, 2)) > 1e-15) == 1   # nothing leaked along y (axis 1 of [k][j][i])
