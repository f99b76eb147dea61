"""-m gpu: BASELINE.json's full sizes (1e5 / 1e6 particles).

Parity against the reference's own compiled objects at size (configs[1]: 1e5 particles x 1000 sub-steps, configs[2]:
1e6 particles x 100 sub-steps, configs[3]: 1e6 cohesive x 30, configs[4]: 1.25e6 lubricated x 12; about 20-60 s of single-core CPU
each), then size-independent properties of the hot path:

  * the directed neighbour list is symmetric (every pair sits in both partners' rows) and agrees with the
    brute-force count of overlapping lattice neighbours;
  * Newton's third law: in a periodic box without body forces the total momentum is conserved to round-off;
  * the run is bitwise reproducible (no atomics on the force path);
  * scatter is conservative: sum(gamma V) = sum(V_p);
  * the boundary round trip (get_local_info -> put_local_info) is the identity on tags.
"""
import numpy as np
import pytest

import sedifoam_b200 as sb
from sedifoam_b200 import cases
from util import make_engine

pytestmark = pytest.mark.gpu


def _fullsize_parity(oracle_mod, case, nsteps, with_fdrag=True):
    from util import directed_from_oracle, engine_rows, make_oracle, rel_err, sort_rows
    o = make_oracle(oracle_mod, case)
    e = make_engine(case)
    o.setup(); e.setup()
    if with_fdrag:
        fd = cases.bench_fluid_force(case)
        o.put_fdrag(fd, case["tag"]); e.put_local_info(fd, case["tag"])
    # list at the first build: bit-exact pair set
    ro, _, _ = directed_from_oracle(o, "gran"); rg, _, _ = engine_rows(e, "gran")
    (ro,) = sort_rows(ro); (rg,) = sort_rows(rg)
    assert ro.shape == rg.shape and np.array_equal(ro, rg)
    half = nsteps // 2
    o.run(half); e.step(half); o.run(nsteps - half); e.step(nsteps - half)
    a, b = o.atoms(), e.atoms()
    L = np.abs(case["box_hi"] - case["box_lo"]).max()
    assert np.array_equal(a["tag"], b["tag"])
    assert rel_err(b["x"], a["x"], scale=L) < 1.0e-6      # north_star: 1e-6 relative after a fixed step count
    assert rel_err(b["v"], a["v"]) < 1.0e-6
    assert rel_err(b["omega"], a["omega"]) < 1.0e-6
    assert e.stat("nbuilds") == o.stat("nbuilds") and e.stat("pair_evals") == o.stat("pair_evals")
    ro, to, so = directed_from_oracle(o, "gran", history=True); rg, tg, sg = engine_rows(e, "gran")
    ro, to, so = sort_rows(ro, to, so); rg, tg, sg = sort_rows(rg, tg, sg)
    assert np.array_equal(ro, rg) and np.array_equal(to, tg)    # the pair set and the touching set after the run
    if np.abs(so).max() > 0:
        assert rel_err(sg, so) < 1.0e-5
    return o.kind, int(to.sum())


def test_config1_1e5_column_parity_with_reference_objects(oracle_mod):
    """configs[1]: 1e5 spheres sedimenting in the periodic column, 1000 DEM sub-steps, against the reference objects"""
    case = cases.random_column()
    assert len(case["tag"]) == 100000
    kind, _ = _fullsize_parity(oracle_mod, case, 1000)
    assert kind == ("reference" if oracle_mod.have_reference() else "port")


def test_config2_1e6_settled_bed_parity_with_reference_objects(oracle_mod):
    """configs[2]: the benchmark's own 1e6-particle settled random bed, 100 DEM sub-steps, against the reference objects"""
    case = cases.settled_bed()
    assert len(case["tag"]) == 1000000
    kind, touching = _fullsize_parity(oracle_mod, case, 100)
    assert touching > 2 * 2.0e6      # directed touching entries: more than two touching pairs per particle


def test_config3_1e6_cohesive_bed_parity_with_reference_objects(oracle_mod):
    """configs[3]: the 1e6-particle cohesive settled bed under the sheared lid (gran/hertzFix/history + fix cohesive over its own
    half list), 30 DEM sub-steps, against the reference objects (fix_cohesive.cpp:138-263)"""
    case = cases.settled_cohesive_bed()
    assert len(case["tag"]) == 1000000
    _fullsize_parity(oracle_mod, case, 30)


def test_config4_1p25e6_polydisperse_lubricated_parity_with_reference_objects(oracle_mod):
    """configs[4]: 1.25e6 polydisperse spheres at phi 0.55, hybrid/overlay gran/hertzFix/history + lubricate/poly over the full list
    (about 50 entries per row in two row segments), 12 DEM sub-steps, against the reference objects
    (pair_lubricate_poly.cpp:233-403)"""
    case = cases.random_poly_lubricated()
    assert len(case["tag"]) == 1250000
    _fullsize_parity(oracle_mod, case, 12)


def _pairs_symmetric(e):
    p = e.pairs()
    sel = p["gran"].astype(bool)
    a = p["ti"][sel].astype(np.int64); b = p["tj"][sel].astype(np.int64); img = p["img"][sel].astype(np.int64)
    key = (a << 32) | b
    rev = (b << 32) | a
    # periodic images: the reverse entry carries the mirrored image code 26 - img
    k1 = np.sort(key * 32 + img); k2 = np.sort(rev * 32 + (26 - img))
    return np.array_equal(k1, k2), len(a)


def test_bed_1e6_list_symmetry_determinism_conservation():
    case = cases.fluidized_bed(dims=(100, 100, 100))
    n = len(case["tag"])
    assert n == 1000000
    res = []
    for rep in range(2):
        e = make_engine(case)
        e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
        e.setup()
        if rep == 0:
            ok, m = _pairs_symmetric(e)
            assert ok and m == 2 * e.stat("gran_pairs")
            # every lattice site overlaps its (up to) 6 neighbours: 3 n - faces pairs
            assert e.stat("gran_pairs") == 3 * 100 * 100 * 99
            g, Ue = e.scatter_alpha_u()
            V = np.prod((case["mesh_hi"] - case["mesh_lo"]) / case["mesh_n"])
            assert abs(g.sum() * V / (n * np.pi / 6 * 5e-4 ** 3) - 1.0) < 1e-11
            loc = e.get_local_info()
            e.put_local_info(np.zeros((n, 3)), loc["tag"])
            loc2 = e.get_local_info()
            assert np.array_equal(np.sort(loc["tag"]), np.arange(1, n + 1)) and np.array_equal(loc["tag"], loc2["tag"])
        e.step(60)
        st = e.atoms()
        assert np.isfinite(st["x"]).all() and np.isfinite(st["v"]).all()
        res.append((st["x"].copy(), st["v"].copy(), st["omega"].copy()))
        e.close()
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)


def test_periodic_1e5_momentum_conservation():
    """configs[1] geometry made fully periodic and force-free: sum(m v) must not move (pair forces are bitwise
    antisymmetric by construction, so only the summation round-off of each particle's own row remains)"""
    case = cases.sediment_column(dims=(36, 103, 27), phi=0.52, jitter_frac=0.02)
    assert len(case["tag"]) == 100116
    case["periodic"] = ("p", "p", "p")
    case["script"] = "\n".join(ln for ln in case["script"].splitlines() if "gravity" not in ln and "wall/granFix" not in ln)
    rng = np.random.default_rng(2)
    case["v"] = rng.normal(scale=0.05, size=case["x"].shape)
    case["v"] -= case["v"].mean(axis=0)
    e = make_engine(case)
    e.setup()
    m = case["rho"] * np.pi / 6 * case["diam"] ** 3
    e.step(400)
    st = e.atoms()
    p1 = (st["rmass"][:, None] * st["v"]).sum(axis=0)
    scale = (m[:, None] * np.abs(case["v"])).sum()
    assert e.stat("pair_evals") > 0 and e.stat("nbuilds") >= 1
    assert np.abs(p1).max() < 1e-11 * scale
    ok, _ = _pairs_symmetric(e)
    assert ok


def test_cohesive_and_lubrication_full_lists_run_at_scale():
    """configs[3] / configs[4] kernels at 2.5e5 particles: finite state, symmetric lists, type list populated"""
    c3 = cases.cohesive_shear_bed(dims=(64, 60, 64))
    e = make_engine(c3)
    e.step(50)
    st = e.atoms()
    assert np.isfinite(st["v"]).all() and e.stat("type_entries") >= e.stat("gran_entries") > 0
    e.close()
    c4 = cases.poly_lubricated(dims=(60, 60, 60))
    e = make_engine(c4)
    e.step(20)
    st = e.atoms()
    assert np.isfinite(st["v"]).all() and e.stat("type_entries") > 0
    ok, _ = _pairs_symmetric(e)
    assert ok
