"""CPU: analytic known answers for the north-star kernels the reference ships no golden data for (SURVEY.md 8c:
a1 gran/hertzFix/history, a4 fix cohesive, a5 lubricate/poly, a9 ErgunWenYu).  Both checkers are held to them: the
reference's own sources compiled by path (kind "reference", when /root/reference was present at build time) and the port."""
import numpy as np
import pytest

KINDS = ["port", "reference"]


def _sim(oracle_mod, kind, lo, hi, periodic, x, d, rho, v, script):
    if kind == "reference" and not oracle_mod.have_reference():
        pytest.skip("oracle/_ref not built on this machine")
    o = oracle_mod.Oracle(kind)
    o.command("atom_style sphere")
    o.command("boundary " + " ".join(periodic))
    o.command("newton off")
    o.command("communicate single vel yes")
    o.set_box(np.asarray(lo, float), np.asarray(hi, float), 1)
    n = len(x)
    o.add_atoms(np.arange(1, n + 1), np.ones(n, np.int32), np.asarray(d, float), np.full(n, rho), np.asarray(x, float), np.asarray(v, float))
    o.commands(script)
    return o


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("e", [0.9, 0.6, 0.3])
def test_hertzfix_head_on_collision_restitution(oracle_mod, kind, e):
    """pair_gran_hertzFix_history.cpp:192-200: the 'Fix' damping  2 sqrt(5/6) beta sqrt(S_n m*) v_n  with
    beta = -ln e / sqrt(ln^2 e + pi^2)  is the constant-restitution closure of the Hertz contact: two equal spheres meeting
    head on must separate with e times the approach speed, whatever that speed is (gamman IS the restitution coefficient)."""
    d, rho, v0 = 1.0e-3, 2500.0, 0.05
    script = f"""
neighbor 0.0005 bin
neigh_modify delay 0
pair_style gran/hertzFix/history 1.0e7 NULL {e} NULL 0.0 1
pair_coeff * *
timestep 2.0e-8
fix 1 all nve/sphere
"""
    for speed in (v0, 4 * v0):
        x = [[0.0045, 0.005, 0.005], [0.0045 + 1.02 * d, 0.005, 0.005]]
        v = [[speed, 0, 0], [-speed, 0, 0]]
        o = _sim(oracle_mod, kind, (0, 0, 0), (0.01, 0.01, 0.01), ("f", "f", "f"), x, [d, d], rho, v, script)
        o.setup()
        touched = False
        for _ in range(400):
            o.run(200)
            a = o.atoms()
            gap = a["x"][1, 0] - a["x"][0, 0] - d
            touched = touched or gap < 0
            if touched and gap > 0:
                break
        assert touched and gap > 0
        rebound = 0.5 * (a["v"][1, 0] - a["v"][0, 0])
        assert abs(rebound / speed - e) < 0.02 * e, (kind, e, speed, rebound / speed)
        assert abs(a["v"][0, 0] + a["v"][1, 0]) < 1e-12 * speed + 1e-15      # momentum: equal and opposite, exactly


@pytest.mark.parametrize("kind", KINDS)
def test_cohesive_opt1_closed_form(oracle_mod, kind):
    """fix_cohesive.cpp:217-250 (opt 1, unretarded van der Waals between spheres):
    F = A (2R)^6 / 6 / s^2 / (r + 2R)^2 / r^3 along the line of centres, attractive, clamped at s = smin."""
    d, rho = 5.0e-5, 2650.0
    ah, lam, smin, smax = 1.0e-20, 1.0e-7, 4.0e-10, 1.0e-6
    script = f"""
neighbor 2.0e-6 bin
neigh_modify delay 0
pair_style gran/hertzFix/history 1.0e7 NULL 0.9 NULL 0.4 1
pair_coeff * *
timestep 1.0e-9
fix 1 all nve/sphere
fix 2 all cohesive {ah} {lam} {smin} {smax} 1
"""
    for s in (5.0e-8, 5.0e-9, 2.0e-10):
        x = [[2.0e-4, 2.0e-4, 2.0e-4], [2.0e-4 + d + s, 2.0e-4, 2.0e-4]]
        o = _sim(oracle_mod, kind, (0, 0, 0), (4e-4, 4e-4, 4e-4), ("f", "f", "f"), x, [d, d], rho, np.zeros((2, 3)), script)
        o.setup()
        o.run(1)      # FixCohe::setup() lacks the int argument (fix_cohesive.h:33): the force first appears in a step
        f = o.atoms()["f"]
        r = d + s
        se = max(s, smin)
        re = d + se if s < smin else r
        want = ah * d ** 6 / 6.0 / se ** 2 / (re + d) ** 2 / re ** 3 if s >= smin else \
            ah * d ** 6 / 6.0 / smin ** 2 / (smin + 2 * d) ** 2 / (smin + d) ** 3
        assert f[0, 0] > 0 and f[1, 0] < 0                                   # attraction
        assert abs(f[0, 0] / want - 1.0) < 1e-9 and abs(f[0, 0] + f[1, 0]) < 1e-12 * want, (kind, s, f[0, 0], want)
        assert np.abs(f[:, 1:]).max() == 0.0


@pytest.mark.parametrize("kind", KINDS)
def test_lubrication_squeeze_closed_form(oracle_mod, kind):
    """pair_lubricate_poly.cpp:300-330 with flaglog 0: only the squeeze term  6 pi mu r_i beta0^2/(1+beta0)^2 / h * v_n
    (h = gap / r_i, beta0 = r_j / r_i), opposing the approach; flagfld 0 removes the isotropic FLD drag."""
    mu = 1.0e-3
    di, dj, rho = 4.0e-4, 6.0e-4, 2650.0
    gap, vn = 2.0e-5, 0.01
    script = f"""
neighbor 1.0e-4 bin
neigh_modify delay 0
pair_style lubricate/poly {mu} 0 0 1.0e-7 {1.5 * dj} 1 0
pair_coeff * *
timestep 1.0e-9
fix 1 all nve/sphere
"""
    L = 4.0e-3
    c = 0.5 * L
    x = [[c, c, c], [c + 0.5 * (di + dj) + gap, c, c]]
    v = [[vn, 0, 0], [-vn, 0, 0]]
    o = _sim(oracle_mod, kind, (0, 0, 0), (L, L, L), ("p", "p", "p"), x, [di, dj], rho, v, script)
    o.setup()
    f = o.atoms()["f"]
    ri, rj = 0.5 * di, 0.5 * dj
    for k, (ra, rb) in enumerate(((ri, rj), (rj, ri))):
        b0 = rb / ra
        h = gap / ra
        want = 6.0 * np.pi * mu * ra * (b0 * b0 / (1 + b0) ** 2 / h) * (2 * vn)
        sign = -1.0 if k == 0 else 1.0                                       # resists the approach
        assert abs(f[k, 0] / (sign * want) - 1.0) < 1e-9, (kind, k, f[k, 0], sign * want)
    assert np.abs(f[:, 1:]).max() < 1e-25


def test_ergun_wenyu_dilute_limit_and_terminal_velocity(oracle_mod):
    """ErgunWenYu.C:104-132: for alpha -> 0 the closure is the Wen-Yu branch with beta^-2.65 -> 1, i.e. the standard
    sphere drag  F = (pi/8) C_d rho_f d^2 U^2,  C_d = 24 (1 + 0.15 Re^0.687) / Re  (0.44 above Re = 1000); the terminal
    velocity of a single settling sphere follows from the force balance with the reduced weight."""
    d, nu, rhof, rhos, g = 5.0e-4, 1.0e-6, 1000.0, 2650.0, 9.81
    U = np.array([1e-4, 1e-3, 1e-2, 5e-2, 0.2, 3.0])
    K = oracle_mod.jd(oracle_mod.DRAG_ERGUN_WENYU, U, np.zeros_like(U), np.full_like(U, d), nu, rhof)
    Re = U * d / nu
    Cd = np.where(Re > 1000.0, 0.44, 24.0 * (1 + 0.15 * Re ** 0.687) / Re)
    F = K * (np.pi / 6 * d ** 3) * U                                         # (1 - alpha) = 1
    assert np.allclose(F, np.pi / 8 * Cd * rhof * d * d * U * U, rtol=1e-12)
    # terminal velocity by bisection on the restated closure vs the classical iterative formula
    W = (rhos - rhof) * g * np.pi / 6 * d ** 3
    lo, hi = 1e-6, 1.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        Fm = oracle_mod.jd(oracle_mod.DRAG_ERGUN_WENYU, np.array([mid]), np.zeros(1), np.array([d]), nu, rhof)[0] * (np.pi / 6 * d ** 3) * mid
        lo, hi = (mid, hi) if Fm < W else (lo, mid)
    ut = 0.5 * (lo + hi)
    Ret = ut * d / nu
    assert abs(np.pi / 8 * 24.0 * (1 + 0.15 * Ret ** 0.687) / Ret * rhof * d * d * ut * ut / W - 1.0) < 1e-9
    assert 0.05 < ut < 0.09                                                  # 0.5 mm quartz sand in water: about 7 cm/s
