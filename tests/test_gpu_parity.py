"""-m gpu: the CUDA engine (through the C-ABI of include/sedi_b200.h) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): neighbour and cell-owner indices bit-exact; particle positions / velocities within
1e-6 relative after a fixed step count.  Relative = max abs difference / max abs value of the quantity.
"""
import numpy as np
import pytest

from sedifoam_b200 import cases
import util
from util import directed_from_oracle, engine_rows, make_engine, make_oracle, rel_err, sort_rows

pytestmark = pytest.mark.gpu

TOL = 1.0e-6  # north_star tolerance for floating-point state after a fixed step count

SCENARIOS = {
    "hertz_bed_walls": (lambda: cases.fluidized_bed(dims=(12, 14, 12)), 400),
    "hertz_column_periodic": (lambda: cases.sediment_column(dims=(10, 24, 10), phi=0.45, jitter_frac=0.08), 400),
    "hooke_history_bed": (lambda: util.hooke_history_bed(dims=(10, 12, 10)), 300),
    "hooke_bed": (lambda: util.hooke_bed(dims=(8, 10, 8)), 300),
    # moving walls and the cylinder wall of fix wall/granFix (fix_wall_granFix.cpp:254-264, :309-322)
    "wiggle_wall_hertz": (lambda: util.wiggle_wall_bed(dims=(8, 10, 8)), 300),
    "wiggle_wall_hooke": (lambda: util.wiggle_wall_bed(dims=(8, 10, 8), style="hooke"), 300),
    "shear_wall": (lambda: util.shear_wall_bed(dims=(8, 10, 8)), 300),
    "zcylinder": (lambda: util.zcylinder_bed(dims=(9, 9, 8)), 300),
    "zcylinder_rotating": (lambda: util.zcylinder_bed(dims=(9, 9, 8), shear="x"), 300),
    # ragged rows with a force-carrying contact network (the bench's kind of bed): settled random packing
    "settled_random": (lambda: util.settled_random_bed(columns=(2, 2)), 400),
    "cohesive_opt1": (lambda: cases.cohesive_shear_bed(dims=(10, 8, 10), opt=1), 300),
    "cohesive_opt0": (lambda: cases.cohesive_shear_bed(dims=(10, 8, 10), opt=0), 300),
    "lubricate_poly": (lambda: cases.poly_lubricated(dims=(10, 10, 10)), 200),
    # the dense polydisperse packing of configs[4] (phi 0.55, full list of ~50 entries per row, two row segments)
    "lubricate_poly_dense": (lambda: cases.random_poly_lubricated(tiles=(1, 1, 1), tile_n=1500), 200),
    "frozen_floor": (lambda: _frozen(cases.sediment_column(dims=(8, 12, 8), phi=0.50, jitter_frac=0.02)), 300),
    # skin = d: rows of ~32 neighbours, mostly not touching (the skin of cases/example-cases/transport-bedload/in.lammps:12);
    # exercises the look-ahead ring refill of k_step and the growth of the ELL capacity
    "hertz_large_skin": (lambda: cases.fluidized_bed(dims=(9, 10, 9), skin_frac=1.0, vjit=0.05), 300),
    # ragged rows: polydisperse radii, random positions from a dilute lattice with large jitter, periodic box
    "hertz_polydisperse": (lambda: _poly(cases.sediment_column(dims=(9, 14, 9), phi=0.25, jitter_frac=0.3)), 400),
}


def _frozen(case):
    # bottom lattice layer becomes a frozen rough floor: type 2, group + fix freeze, nve on the rest
    y = case["x"][:, 1]
    low = y < y.min() + 0.5 * case["diam"][0]
    case["type"] = np.where(low, 2, 1).astype(np.int32)
    case["ntypes"] = 2
    case["script"] = case["script"].replace("fix 1 all nve/sphere", "group bed type 2\ngroup mobile subtract all bed\nfix 1 mobile nve/sphere")
    case["script"] += "fix fz bed freeze\n"
    return case


def _poly(case):
    rng = np.random.default_rng(23)
    n = len(case["tag"])
    case["diam"] = case["diam"] * rng.uniform(0.6, 1.0, size=n)
    case["v"] = rng.normal(scale=0.2, size=(n, 3))
    return case


def _fdrag(case):
    rng = np.random.default_rng(11)
    n = len(case["tag"])
    m = case["rho"] * np.pi / 6.0 * case["diam"] ** 3
    return rng.normal(scale=2.0, size=(n, 3)) * m[:, None], rng.permutation(case["tag"])


def _pair(oracle_mod, name):
    mk, nsteps = SCENARIOS[name]
    case = mk()
    o = make_oracle(oracle_mod, case)
    e = make_engine(case)
    return case, o, e, nsteps


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_neighbour_lists_bit_exact_at_build(oracle_mod, name):
    case, o, e, _ = _pair(oracle_mod, name)
    o.setup(); e.setup()
    ro, _, _ = directed_from_oracle(o, "gran")
    re_, _, _ = engine_rows(e, "gran")
    (ro,) = sort_rows(ro); (re_,) = sort_rows(re_)
    assert ro.shape == re_.shape and np.array_equal(ro, re_)
    assert len(ro) > 0 or "lubricate" in case["script"]   # the dilute lubrication case has no granular neighbours
    assert e.stat("gran_pairs") == o.stat("gran_pairs")
    if "cohesive" in case["script"]:
        rt, _, _ = directed_from_oracle(o, "half")
    elif "lubricate" in case["script"]:
        rt, _, _ = directed_from_oracle(o, "full")
    else:
        return
    rg, _, _ = engine_rows(e, "type")
    (rt,) = sort_rows(rt); (rg,) = sort_rows(rg)
    assert rt.shape == rg.shape and np.array_equal(rt, rg)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_setup_forces(oracle_mod, name):
    case, o, e, _ = _pair(oracle_mod, name)
    fd, tags = _fdrag(case)
    o.setup(); e.setup()
    a, b = o.atoms(), e.atoms()
    assert np.array_equal(a["tag"], b["tag"])
    assert np.array_equal(a["x"], b["x"])  # nothing has moved; periodic wrap identical
    # sums of ~50 lubrication terms with cancellation (dense packing): a few ulp per term (reciprocal / rsqrt forms)
    assert rel_err(b["f"], a["f"]) < (1e-11 if name == "lubricate_poly_dense" else 1e-12)
    assert rel_err(b["torque"], a["torque"], scale=max(np.abs(a["torque"]).max(), 1e-300)) < 1e-10 or np.abs(a["torque"]).max() == 0


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_trajectory_parity(oracle_mod, name):
    case, o, e, nsteps = _pair(oracle_mod, name)
    fd, tags = _fdrag(case)
    o.setup(); e.setup()
    o.put_fdrag(fd, tags); e.put_local_info(fd, tags)
    half = nsteps // 2
    o.run(half); e.step(half)
    o.run(nsteps - half); e.step(nsteps - half)
    a, b = o.atoms(), e.atoms()
    L = np.abs(case["box_hi"] - case["box_lo"]).max()
    assert np.isfinite(b["x"]).all() and np.isfinite(b["v"]).all()
    assert rel_err(b["x"], a["x"], scale=L) < TOL
    assert rel_err(b["v"], a["v"]) < TOL
    if np.abs(a["omega"]).max() > 0:
        assert rel_err(b["omega"], a["omega"]) < TOL
    assert rel_err(b["f"], a["f"]) < 1e-5
    # list bookkeeping is integer work: identical rebuild count, identical pair-evaluation count, identical pair set
    assert e.stat("nbuilds") == o.stat("nbuilds")
    assert e.stat("steps") == nsteps
    assert e.stat("pair_evals") == o.stat("pair_evals")
    ro, to, so = directed_from_oracle(o, "gran", history="hooke " not in case["script"])
    rg, tg, sg = engine_rows(e, "gran")
    ro, to, so = sort_rows(ro, to, so); rg, tg, sg = sort_rows(rg, tg, sg)
    assert np.array_equal(ro, rg)
    if "lubricate" in case["script"]:
        rt, _, _ = directed_from_oracle(o, "full")
        rq, _, _ = engine_rows(e, "type")
        (rt,) = sort_rows(rt); (rq,) = sort_rows(rq)
        assert len(rt) > 0 and np.array_equal(rt, rq)
    if to is not None and len(to):
        assert np.array_equal(to, tg)
        if np.abs(so).max() > 0:
            assert rel_err(sg, so) < 1e-5


def test_checker_is_the_reference_objects(oracle_mod):
    """on a box that has oracle/_ref/libsedi_ref.so (it travels with the snapshot) every parity test of this suite
    compares the CUDA path with the reference's own compiled sources, not with the port"""
    if not oracle_mod.have_reference():
        pytest.skip("oracle/_ref/libsedi_ref.so absent: the checker is the port (pinned to the reference objects by tests/test_oracle_pinning.py)")
    case, o, e, _ = _pair(oracle_mod, "hertz_bed_walls")
    assert o.kind == "reference" and o.lib.ora_has_ref() == 1


@pytest.mark.parametrize("name", ["hertz_bed_walls", "wiggle_wall_hertz", "wiggle_wall_hooke", "shear_wall", "zcylinder_rotating", "settled_random"])
def test_wall_history_parity(oracle_mod, name):
    case, o, e, nsteps = _pair(oracle_mod, name)
    o.run(nsteps); e.step(nsteps)
    nw = sum(1 for ln in case["script"].splitlines() if "wall/gran" in ln)
    for w in range(nw):
        a, b = o.wall_shear(w), e.wall_shear(w)
        if np.abs(a).max() > 0:
            assert rel_err(b, a) < 1e-5
        assert np.array_equal(np.abs(a).sum(axis=1) > 0, np.abs(b).sum(axis=1) > 0)


def test_run_is_deterministic_and_chunk_independent(oracle_mod, monkeypatch):
    """same inputs -> bitwise identical state, whatever the launch chunking (no atomics on the force path)"""
    case = cases.fluidized_bed(dims=(12, 14, 12))
    res = []
    for chunk in ("16", "16", "5"):
        monkeypatch.setenv("SEDI_CHUNK", chunk)
        e = make_engine(case)
        e.step(150); e.step(37)
        res.append(e.atoms())
    for k in ("x", "v", "omega", "f", "torque"):
        assert np.array_equal(res[0][k], res[1][k])
        assert np.array_equal(res[0][k], res[2][k])


def test_step_split_invariance(oracle_mod):
    """run 100 == run 40 + run 60 (lammps_step is called once per coupling sub-cycle, softParticleCloud.C:886)"""
    case = cases.sediment_column(dims=(10, 16, 10), phi=0.45, jitter_frac=0.08)
    e1 = make_engine(case); e1.step(100)
    e2 = make_engine(case); e2.step(40); e2.step(60)
    a, b = e1.atoms(), e2.atoms()
    for k in ("x", "v", "omega"):
        assert np.array_equal(a[k], b[k])


def test_edge_cases_empty_single_and_lost_particle(oracle_mod):
    """empty system, one particle, and a particle that leaves a non-periodic box (it stays owned, LAMMPS `lost ignore`
    is not modelled: the reference inputs abort on lost atoms, thermo_modify lost error)"""
    case = cases.fluidized_bed(dims=(2, 2, 2))
    for k in ("tag", "type", "diam", "rho"):
        case[k] = case[k][:0]
    case["x"] = case["x"][:0]; case["v"] = case["v"][:0]
    e = make_engine(case)
    e.step(5)
    assert e.get_local_n() == 0 and e.stat("pair_evals") == 0
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    g, Ue = e.scatter_alpha_u()
    assert g.sum() == 0.0
    e.close()
    one = cases.fluidized_bed(dims=(1, 1, 1), vjit=0.0)
    one["script"] = "\n".join(ln for ln in one["script"].splitlines() if "wall/granFix" not in ln)
    one["v"][:] = (3.0, 0.0, 0.0)     # flies out of the box in x
    o = make_oracle(oracle_mod, one); e = make_engine(one)
    o.run(500); e.step(500)
    a, b = o.atoms(), e.atoms()
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["v"], b["v"])
    assert b["x"][0, 0] > one["box_hi"][0]
