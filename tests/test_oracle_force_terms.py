"""CPU: closed-form checks of the restated history / lubrication / inlet branches of updateDragOnParticles
(lammpsFoam/enhancedCloud.C:197-257, g1n :1372-1384, softParticleCloud::pointInRegion softParticleCloud.C:1354-1415).
The reference ships no golden data for these branches (they are off in every shipped case): parity unpinned, known answers only."""
import numpy as np


def _one(oracle_mod, flags, x, U, UOld, Uf, UfOld, d=1e-3, nub=1e-6, rhob=1000.0, dT=1e-4, k=1, S=None, n0=None, inlet=(0, 0, 0),
         box=(0,) * 9, opt=0, ecc=(0, 0, 0)):
    S = np.zeros((1, 3)) if S is None else S
    n0 = np.zeros(1) if n0 is None else n0
    F = np.zeros((1, 3))
    m = np.array([2650.0 * np.pi / 6 * d ** 3])
    oracle_mod.particle_force_extra(np.zeros(1, np.int32), np.array([x], float), np.array([d]), m, np.array([U], float), np.array([UOld], float),
                                    np.array([Uf], float), np.array([UfOld], float), flags, nub, rhob, dT, k, S, n0, F, inlet, box, opt, ecc)
    return F[0], S[0], n0[0], m[0]


def test_wall_lubrication_closed_form(oracle_mod):
    d = 1e-3
    for gap, active in ((0.05 * d, True), (0.2 * d, False), (0.00005 * d, False)):
        F, _, _, _ = _one(oracle_mod, 64, (0, 0.5 * d + gap, 0), (0, -0.03, 0), (0, -0.03, 0), (0, 0, 0), (0, 0, 0), d=d)
        want = 6 * 3.1416 * 1e-6 * 1000.0 * 0.03 / gap * d * d / 4.0 if active else 0.0
        assert np.isclose(F[1], want, rtol=1e-9, atol=0) and F[0] == 0 and F[2] == 0   # 1e-9: gap = y - d/2 cancels six digits


def test_inlet_force_replaces_and_regions(oracle_mod):
    box = (0, 1, 0, 1, 0, 1, 0, 0, 0)
    F, _, _, m = _one(oracle_mod, 128, (0.5, 0.5, 0.5), (0.1, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), inlet=(0.3, 0, 0.1), box=box, opt=1)
    assert np.allclose(F, m * (np.array([0.3, 0, 0.1]) - np.array([0.1, 0, 0])) / 1e-4, rtol=1e-14)
    F, _, _, _ = _one(oracle_mod, 128, (1.5, 0.5, 0.5), (0.1, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), inlet=(0.3, 0, 0.1), box=box, opt=1)
    assert np.all(F == 0)
    # hollow cylinder along x from (0,0,0) to (1,0,0), radii 0.2 .. 0.5
    cyl = (0, 1, 0, 0, 0, 0, 0.2, 0.5, 0)
    for pt, inside in (((0.5, 0.3, 0), True), ((0.5, 0.1, 0), False), ((0.5, 0.6, 0), False), ((1.2, 0.3, 0), False)):
        F, _, _, _ = _one(oracle_mod, 128, pt, (0, 0, 0), (0, 0, 0), (0, 0, 0), (0, 0, 0), inlet=(1, 0, 0), box=cyl, opt=2)
        assert (F[0] != 0) == inside


def test_history_force_short_time_regime(oracle_mod):
    """constant acceleration from rest: while t < tau_h the sum grows linearly and FH = g1n(n) * n * Cb a / sqrt(dT)"""
    d, nub, rhob, dT, a = 1e-3, 1e-6, 1000.0, 1e-4, 2.0
    S = np.zeros((1, 3)); n0 = np.zeros(1)
    Cb = -1.5 * d * d * rhob * (3.1416 * nub) ** 0.5
    for k in range(1, 6):
        U = (a * dT * k, 0, 0); Uo = (a * dT * (k - 1), 0, 0)
        F, Sk, n0k, _ = _one(oracle_mod, 32, (0, 0, 0), U, Uo, (0.05, 0, 0), (0.05, 0, 0), d=d, nub=nub, rhob=rhob, dT=dT, k=k, S=S, n0=n0)
        g = 0.9279 if k < 1 else 0.9279 * (2 * k - 1) / k * k ** (-k / (2 * k - 1)) + 0.001531
        assert n0k == 0.0
        assert np.isclose(Sk[0], k * Cb * a / np.sqrt(dT), rtol=1e-10)
        assert np.isclose(F[0], g * Sk[0] * dT, rtol=1e-12)


def test_history_force_window_regime_moves_n0(oracle_mod):
    d, nub, dT = 1e-4, 1e-6, 1e-3      # tau_d = 0.01, large Re -> tau_h ~ 0.087^2 * 0.01 << dT * k
    S = np.array([[1.0, 0.0, 0.0]]); n0 = np.zeros(1)
    F, Sk, n0k, _ = _one(oracle_mod, 32, (0, 0, 0), (1.0, 0, 0), (1.0, 0, 0), (0, 0, 0), (0, 0, 0), d=d, nub=nub, dT=dT, k=50, S=S, n0=n0)
    ReP = 1.0 * d / nub
    tau_h = d * d / nub * (0.632 / ReP + 0.087) ** 2
    dn = tau_h / dT
    assert np.isclose(n0k, 50 - dn, rtol=1e-12)
    assert np.isclose(Sk[0], (dn - 1) / dn * 1.0, rtol=1e-12)      # tau_h == tau_h_old, no new increment (dupdt = 0)
    assert np.isclose(F[0], 0.9279 * Sk[0] * dT, rtol=1e-12)       # dn < 1 -> g1n = 0.9279


def test_blockmesh_grading_and_labels(oracle_mod):
    """host helper for graded / stacked blockMesh blocks: geometric face spacing and block-wise labels"""
    from sedifoam_b200 import cases
    f = cases.blockmesh_divide(0.0, 1.0, 10, 10.0)
    w = np.diff(f)
    assert np.isclose(w[-1] / w[0], 10.0) and np.allclose(w[1:] / w[:-1], 10.0 ** (1 / 9)) and f[0] == 0.0 and f[-1] == 1.0
    assert np.allclose(cases.blockmesh_divide(0.0, 2.0, 4, 1.0), [0, 0.5, 1.0, 1.5, 2.0])
    xf, yf, zf, label = cases.blockmesh_stacked((0, 2, 2, 1.0), [(0, 1, 2, 1.0), (1, 3, 1, 1.0)], (0, 1, 2, 1.0))
    assert sorted(label) == list(range(12))
    # block 0 holds labels 0..7 (2 x 2 x 2), block 1 labels 8..11 (2 x 1 x 2): tensor cell (i=1, j=2, k=1) is block 1's (1, 0, 1)
    assert label[1 + 2 * (2 + 3 * 1)] == 8 + 1 + 2 * (0 + 1 * 1)
    pts = np.array([[0.5, 0.25, 0.25], [1.5, 2.0, 0.75], [0.5, 3.0, 0.5]])
    assert list(oracle_mod.cell_owner_rect(pts, xf, yf, zf, label)) == [0, 8 + 1 + 2 * 1, -1]
