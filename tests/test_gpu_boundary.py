"""-m gpu: the drop-in boundary (interfaceToLammps/library.h) and the coupling kernels against the oracle."""
import numpy as np
import pytest

from sedifoam_b200 import cases
from sedifoam_b200 import DRAG_ERGUN_WENYU, DRAG_SYAMLAL_OBRIEN, FORCE_DRAG, FORCE_PGRAD, FORCE_BUOY, FORCE_ADDEDMASS, FORCE_LIFT
from util import make_engine, make_oracle, rel_err

pytestmark = pytest.mark.gpu


def test_initial_info_and_local_info(oracle_mod):
    case = cases.fluidized_bed(dims=(6, 7, 6))
    e = make_engine(case)
    n = len(case["tag"])
    assert e.get_global_n() == n and e.get_local_n() == n
    assert e.get_initial_np(1)[0] == n
    info = e.get_initial_info()
    o = np.argsort(info["tag"])
    assert np.array_equal(info["tag"][o], case["tag"])
    assert np.array_equal(info["x"][o], case["x"])
    assert np.array_equal(info["diam"][o], case["diam"])
    # density comes back through the reference's truncated-pi literal (library.cpp:200): close to, not equal to, rho
    assert np.all(np.abs(info["rho"][o] / case["rho"] - 1.0) < 1e-11) and not np.array_equal(info["rho"][o], case["rho"])
    assert np.allclose(e.get_local_domain(), np.stack([case["box_lo"], case["box_hi"]], axis=1).ravel())
    e.setup()
    info2 = e.get_initial_info()   # now served from the device
    o2 = np.argsort(info2["tag"])
    for k in ("x", "v", "diam", "rho", "type"):
        assert np.array_equal(info2[k][o2], info[k][o])
    rng = np.random.default_rng(3)
    foam = rng.integers(0, 4, size=n).astype(np.int32)
    perm = rng.permutation(n)
    e.put_local_info(np.zeros((n, 3)), case["tag"][perm], foam_cpu=foam[perm])
    loc = e.get_local_info()
    o3 = np.argsort(loc["tag"])
    assert np.array_equal(loc["tag"][o3], case["tag"])
    assert np.array_equal(loc["foamCpuId"][o3], foam)
    assert np.all(loc["lmpCpuId"] == 0)
    assert np.array_equal(loc["x"][o3], case["x"])


def test_timestep_accessors():
    case = cases.fluidized_bed(dims=(4, 4, 4))
    e = make_engine(case)
    assert e.get_timestep() == 2.0e-6
    e.set_timestep(1.0e-6)
    assert e.get_timestep() == 1.0e-6


def test_put_local_info_drives_fdrag(oracle_mod):
    """a free particle under a constant fluid force: v = F/m * t exactly as the oracle integrates it"""
    case = cases.sediment_column(dims=(4, 4, 4), phi=0.05, jitter_frac=0.0)
    case["script"] = case["script"].replace("fix 2 all gravity 9.8", "fix 2 all gravity 0.0")
    o = make_oracle(oracle_mod, case); e = make_engine(case)
    n = len(case["tag"])
    F = np.tile([1e-7, -2e-7, 3e-7], (n, 1)) * np.arange(1, n + 1)[:, None]
    o.setup(); e.setup()
    perm = np.random.default_rng(5).permutation(n)
    o.put_fdrag(F[perm], case["tag"][perm]); e.put_local_info(F[perm], case["tag"][perm])
    o.run(50); e.step(50)
    a, b = o.atoms(), e.atoms()
    assert np.array_equal(a["v"], b["v"]) and np.array_equal(a["x"], b["x"])   # no pair forces: bitwise
    m = case["rho"] * np.pi / 6 * case["diam"] ** 3
    # the setup force evaluation preceded the put, so the first half kick saw no fluid force: 49.5 steps of F/m
    assert np.allclose(b["v"], F / m[:, None] * 49.5 * 2e-6, rtol=1e-9)


def test_added_mass_carrier_rho(oracle_mod):
    """fix fdrag <carrier_rho>: LAMMPS-side added-mass term (fix_fluid_drag.cpp:144-163)"""
    case = cases.sediment_column(dims=(5, 6, 5), phi=0.45, jitter_frac=0.08)
    case["script"] = case["script"].replace("fix 3 all fdrag", "fix 3 all fdrag 1000")
    o = make_oracle(oracle_mod, case); e = make_engine(case)
    o.run(200); e.step(200)
    a, b = o.atoms(), e.atoms()
    assert rel_err(b["v"], a["v"]) < 1e-6 and rel_err(b["x"], a["x"]) < 1e-6


def _fields(case, rng):
    C = int(np.prod(case["mesh_n"]))
    Uf = rng.normal(scale=0.05, size=(C, 3)); gradp = rng.normal(scale=100.0, size=(C, 3))
    DDtU = rng.normal(scale=5.0, size=(C, 3)); curlU = rng.normal(scale=20.0, size=(C, 3))
    gamma = rng.uniform(0.0, 0.6, size=C)
    return Uf, gamma, gradp, DDtU, curlU


@pytest.mark.parametrize("model", [DRAG_ERGUN_WENYU, DRAG_SYAMLAL_OBRIEN])
@pytest.mark.parametrize("flags", [FORCE_DRAG | FORCE_PGRAD, FORCE_DRAG | FORCE_PGRAD | FORCE_BUOY | FORCE_ADDEDMASS | FORCE_LIFT])
def test_coupling_gather_force_scatter(oracle_mod, model, flags):
    case = cases.fluidized_bed(dims=(12, 14, 12), vjit=0.05)
    rng = np.random.default_rng(17)
    e = make_engine(case)
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.coupling_config(model, flags, case["nub"], case["rhob"], case["g"], 2.0e-4)
    Uf, gamma, gradp, DDtU, curlU = _fields(case, rng)
    e.step(30)   # also records UOld for the added-mass term
    st0 = e.atoms()
    e.step(20)
    st = e.atoms()
    e.put_cell_fields(Uf, gamma, gradp, DDtU, curlU)
    e.enable_diag(True)
    e.compute_fluid_force()
    dg = e.coupling_diag()
    # cell owner: bit exact
    cell = oracle_mod.cell_owner(st["x"], case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    assert np.array_equal(dg["cell"], cell)
    d = 2.0 * st["radius"]
    ref = oracle_mod.particle_force(cell, d, st["v"], st0["v"] if False else _uold(e, st0, st), Uf, gamma, gradp, DDtU, curlU, model, flags,
                                    case["nub"], case["rhob"], np.asarray(case["g"], float), 2.0e-4)
    assert rel_err(dg["Uri"], ref["Uri"]) < 1e-14
    assert rel_err(dg["alpha"], ref["alpha"]) == 0
    assert rel_err(dg["Jd"], ref["Jd"]) < 1e-12      # pow() differs from glibc in the last ulps
    if oracle_mod.have_reference():   # the reference's own ErgunWenYu.C / SyamlalOBrien.C (stub-compiled into oracle/_ref)
        jd_ref = oracle_mod.jd(model, dg["magUri"], dg["alpha"], d, case["nub"], case["rhob"], kind="reference")
        assert rel_err(dg["Jd"], jd_ref) < 1e-12
    assert rel_err(dg["F"], ref["F"]) < 1e-12
    # scatter 1 (particleToEulerianField) and scatter 2 (calcTcFields)
    C = len(gamma)
    cellV = np.full(C, np.prod((case["mesh_hi"] - case["mesh_lo"]) / case["mesh_n"]))
    g_ref, Ue_ref = oracle_mod.particle_to_eulerian(cell, d, st["v"], cellV)
    e.put_cell_fields(gamma=gamma)
    A_ref, Om_ref = oracle_mod.calc_tc(cell, d, st["v"], Uf, gamma, cellV, model, case["nub"], case["rhob"])
    A, Om = e.calc_tc()
    assert rel_err(A, A_ref) < 1e-11 and np.all(Om == 0) and np.all(Om_ref == 0)
    g, Ue = e.scatter_alpha_u()
    assert rel_err(g, g_ref) < 1e-12 and rel_err(Ue, Ue_ref) < 1e-11
    # conservation: sum gamma V = sum Vp  (the reference's built-in invariant, enhancedCloud.C:964-976)
    assert abs((g * cellV).sum() / (np.pi / 6 * (d[cell >= 0] ** 3).sum()) - 1.0) < 1e-12


def _uold(e, st0, st):
    # UOld is the velocity before the last lammps_step call
    return st0["v"]


def test_single_sphere_golden_curve():
    """configs[0] = shipped cases/auto-testing/test-cases/xiaocase3: v_y(t) against data/lammps08.dat and xiaoCase3.dat"""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.loadtxt(os.path.join(here, "golden", "xiaocase3_lammps08.dat"))
    xiao = np.loadtxt(os.path.join(here, "golden", "xiaocase3_xiaoCase3.dat"))
    case = cases.single_sphere()
    e = make_engine(case)
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.coupling_config(DRAG_SYAMLAL_OBRIEN, FORCE_DRAG | FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 2e-5)
    Uf, gamma, gradp = cases.uniform_fields(case)
    e.put_cell_fields(Uf, gamma, gradp)
    t, vy = [0.0], [0.0]
    for k in range(250):
        e.compute_fluid_force()
        e.sedi_step(100)
        t.append((k + 1) * 2e-5); vy.append(e.atoms()["v"][0, 1])
    t = np.array(t); vy = np.array(vy)
    for row in gold[2:]:   # see tests/test_oracle_golden.py for the tolerances
        assert abs(np.interp(row[0], t, vy) - row[2]) < 0.04 * 0.05
    for tt, vv in xiao:
        if tt > 2e-4:
            assert abs(np.interp(tt, t, vy) - vv) < 0.05 * 0.05


def test_smoothing_matches_restated_solver(oracle_mod):
    """diffusion smoothing: PCG on the GPU against the sparse direct solve of the restatement;
    conservation of sum(phi V) is the reference's own invariant (enhancedCloud.C:434-435, 975-976)"""
    case = cases.fluidized_bed(dims=(10, 12, 9))
    e = make_engine(case)
    nc = np.array([7, 9, 5], np.int32)
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], nc)
    dx = (case["mesh_hi"] - case["mesh_lo"]) / nc
    rng = np.random.default_rng(4)
    C = int(np.prod(nc))
    for b, steps, D in ((2.0 * dx[0], 3, (1.0, 1.0, 1.0)), (4.0 * dx[1], 2, (1.0, 0.0, 2.5))):
        e.smooth_config(b, steps, D)
        phi = rng.uniform(0.0, 1.0, size=C); vec = rng.normal(size=(C, 3))
        ref_s = oracle_mod.smooth_field(phi, nc, dx, b, steps, D)
        ref_v = oracle_mod.smooth_field(vec, nc, dx, b, steps, D)
        got_s = e.smooth_field(phi); got_v = e.smooth_field(vec)
        assert rel_err(got_s, ref_s) < 1e-10 and rel_err(got_v, ref_v) < 1e-10
        assert abs(got_s.sum() - phi.sum()) < 1e-10 * phi.sum()
        assert np.abs(got_v.sum(axis=0) - vec.sum(axis=0)).max() < 1e-9 * np.abs(vec).sum()
        assert got_s.std() < phi.std() and 0 < e.smooth_last_iters() < 200


def test_scatter_with_smoothing_flags(oracle_mod):
    """alphaSmooth / UpSmooth / dragSmooth wired into particleToEulerianField and calcTcFields in the reference's order"""
    case = cases.fluidized_bed(dims=(12, 14, 12), vjit=0.05)
    e = make_engine(case)
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.coupling_config(DRAG_ERGUN_WENYU, FORCE_DRAG | FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 2e-4)
    nc = case["mesh_n"]; dx = (case["mesh_hi"] - case["mesh_lo"]) / nc
    b, steps = 2.0e-3, 2
    st = e.atoms(); e.setup(); st = e.atoms()
    cell = oracle_mod.cell_owner(st["x"], case["mesh_lo"], case["mesh_hi"], nc)
    d = 2.0 * st["radius"]; C = int(np.prod(nc)); cellV = np.full(C, np.prod(dx))
    # reference order (enhancedCloud.C:932-962): /V, smooth gamma, smooth Ue, then Ue /= gamma
    g0, _ = oracle_mod.particle_to_eulerian(cell, d, st["v"], cellV)
    mom = np.zeros((C, 3)); vol = np.pi / 6 * d ** 3
    np.add.at(mom, cell[cell >= 0], (vol[:, None] * st["v"])[cell >= 0])
    mom /= cellV[:, None]
    g_ref = oracle_mod.smooth_field(g0, nc, dx, b, steps)
    m_ref = oracle_mod.smooth_field(mom, nc, dx, b, steps)
    Ue_ref = np.where(g_ref[:, None] > 1e-150, m_ref / np.maximum(g_ref[:, None], 1e-300), m_ref)
    e.smooth_config(b, steps, None, 2 | 8)
    g, Ue = e.scatter_alpha_u()
    assert rel_err(g, g_ref) < 1e-9 and rel_err(Ue, Ue_ref) < 1e-8
    assert abs(g.sum() / g0.sum() - 1.0) < 1e-10
    # dragSmooth: Asrc (1-gamma) -> smooth -> / (1-gamma)
    Uf = np.tile([0.0, 0.05, 0.0], (C, 1))
    e.put_cell_fields(Uf, g_ref, None)
    A0, _ = oracle_mod.calc_tc(cell, d, st["v"], Uf, g_ref, cellV, DRAG_ERGUN_WENYU, case["nub"], case["rhob"])
    A_ref = oracle_mod.smooth_field(A0 * (1 - g_ref[:, None]), nc, dx, b, steps) / (1 - g_ref[:, None])
    e.smooth_config(b, steps, None, 4)
    A, _ = e.calc_tc()
    assert rel_err(A, A_ref) < 1e-8


def test_create_and_delete_particles(oracle_mod):
    """lammps_create_particle / lammps_delete_particle (library.cpp:406-621): counts, tags, masses with the reference's
    pi literal, and the run continues with the new population"""
    case = cases.fluidized_bed(dims=(5, 5, 5), vjit=0.0)
    e = make_engine(case)
    e.command("group active type 1")
    e.step(20)
    n0 = e.get_local_n()
    top = case["box_hi"].copy()
    pos = np.array([[0.5 * top[0], 0.9 * top[1], 0.5 * top[2]], [0.25 * top[0], 0.9 * top[1], 0.25 * top[2]]])
    e.create_particle(pos, [1001.0, 1002.0], 4.0e-4, 2500.0, 1, (0.0, -0.1, 0.0))
    assert e.get_local_n() == n0 + 2 and e.get_global_n() == n0 + 2
    e.step(20)
    st = e.atoms()
    k = int(np.flatnonzero(st["tag"] == 1001)[0])
    r = 2.0e-4
    assert st["radius"][k] == r and st["rmass"][k] == 4.0 * 3.14159265358917323846 / 3.0 * r * r * r * 2500.0
    assert st["v"][k, 1] < -0.1           # it has been falling under gravity
    e.delete_particle([1002, int(case["tag"][0])])
    assert e.get_local_n() == n0
    e.step(20)
    st = e.atoms()
    assert 1002 not in st["tag"] and case["tag"][0] not in st["tag"] and 1001 in st["tag"]
    assert np.isfinite(st["x"]).all()


def test_four_spheres_rho_golden_dump_gpu(oracle_mod):
    """second shipped golden case, multiParticlesCollideRho (tolerances: tests/test_oracle_golden.py::check_four_spheres_rho)"""
    from test_oracle_golden import check_four_spheres_rho
    case = cases.four_spheres_collide("rho")
    e = make_engine(case)
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.coupling_config(DRAG_SYAMLAL_OBRIEN, FORCE_DRAG | FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 1e-3)
    Uf, _, gradp = cases.uniform_fields(case)
    e.put_cell_fields(Uf, None, gradp)
    e.setup()
    rows = []
    for k in range(400):
        e.scatter_alpha_u(device_only=True)
        e.compute_fluid_force()
        e.sedi_step(50)
        if (k + 1) % 20 == 0:
            rows.append(e.atoms())
    check_four_spheres_rho(rows)


def test_four_spheres_golden_dump_gpu(oracle_mod):
    """shipped case multiParticlesCollideDia through the device-resident coupling API, against the reference's own
    dump rows (tolerances and their reason: tests/test_oracle_golden.py::check_four_spheres)"""
    from test_oracle_golden import check_four_spheres
    import os
    case = cases.four_spheres_collide()
    e = make_engine(case)
    # the case's own dump command (in.lammps:31): the golden files are per-particle extracts of this file
    e.command("dump id all custom 1000 four_spheres.bubblemd id type diameter mass x y z vx vy vz")
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.coupling_config(DRAG_SYAMLAL_OBRIEN, FORCE_DRAG | FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 1e-3)
    Uf, _, gradp = cases.uniform_fields(case)
    e.put_cell_fields(Uf, None, gradp)
    e.setup()
    rows = []
    for k in range(400):
        e.scatter_alpha_u(device_only=True)   # gamma of the current positions feeds alpha_p
        e.compute_fluid_force()
        e.sedi_step(50)
        if (k + 1) % 20 == 0:
            rows.append(e.atoms())
    check_four_spheres(rows)
    e.close()
    # ---- the dump file itself: EXTERNAL LAMMPS `dump custom` text layout, one snapshot per 1000 DEM steps incl. step 0
    path = os.path.join(os.environ["SEDI_DUMP_DIR"], "four_spheres.bubblemd")
    lines = open(path).read().split("\n")
    snaps = [i for i, l in enumerate(lines) if l == "ITEM: TIMESTEP"]
    assert len(snaps) == 21 and [int(lines[i + 1]) for i in snaps] == list(range(0, 20001, 1000))
    s0 = snaps[0]
    assert lines[s0 + 2] == "ITEM: NUMBER OF ATOMS" and lines[s0 + 3] == "4"
    assert lines[s0 + 4] == "ITEM: BOX BOUNDS ff ff ff" and lines[s0 + 5:s0 + 8] == ["0 0.2", "0 0.1", "0 0.1"]
    assert lines[s0 + 8] == "ITEM: ATOMS id type diameter mass x y z vx vy vz "
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multiParticlesCollideDia")
    for p in range(4):   # step-0 rows are identical TEXT to the reference's dump rows (same "%d %d %g ... " formatting)
        assert lines[s0 + 9 + p] == open(os.path.join(gold, "p%d.dat" % (p + 1))).readline().rstrip("\n")
    for n in range(20):  # later snapshots carry the same state the API returns, at %g precision
        got = np.array([[float(t) for t in lines[snaps[n + 1] + 9 + p].split()] for p in range(4)])
        assert np.allclose(got[:, 4:7], rows[n]["x"], rtol=6e-6, atol=0) and np.allclose(got[:, 7:10], rows[n]["v"], rtol=6e-6, atol=1e-12)


@pytest.mark.parametrize("subcycles", [1, 2])
def test_coupling_history_lubrication_inlet_terms(oracle_mod, subcycles):
    """the last three branches of updateDragOnParticles (enhancedCloud.C:197-257): reduced-order Basset history force
    carried over four coupling steps (per-particle sumDeltaFb / n0 state, both regimes of the window), lubrication
    against the y = 0 wall, and inlet forcing inside a box region -- against the restatement, in the reference's order"""
    from sedifoam_b200 import FORCE_HISTORY, FORCE_WALL_LUB, FORCE_INLET
    case = cases.fluidized_bed(dims=(10, 8, 9), vjit=0.05)
    d0 = float(case["diam"][0])
    case["x"] = case["x"].copy()
    case["x"][:, 1] += 0.05 * d0                       # bottom layer: wall gap 0.049 d, inside the lubrication window
    e = make_engine(case)
    rng = np.random.default_rng(23)
    flags = FORCE_DRAG | FORCE_PGRAD | FORCE_HISTORY | FORCE_WALL_LUB | FORCE_INLET
    deltaT = 2.0e-3                                     # tau_h ~ 3 ms here: steps 3 and 4 run in the moving-window regime
    nub, rhob = case["nub"], case["rhob"]
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.coupling_config(DRAG_ERGUN_WENYU, flags, nub, rhob, case["g"], deltaT)
    ext = case["box_hi"] - case["box_lo"]
    box = [case["box_lo"][0], case["box_lo"][0] + 0.3 * ext[0], case["box_lo"][1], case["box_hi"][1], case["box_lo"][2], case["box_hi"][2], 0, 0, 0]
    inlet = (0.0, 0.01, 0.002)
    e.coupling_inlet(inlet, box, 1)
    e.enable_diag(True)
    n = len(case["tag"])
    S = np.zeros((n, 3)); n0 = np.zeros(n)
    e.setup()
    v_prev = np.zeros((n, 3))                           # softParticle starts with UOld = 0 (softParticle.C:58)
    Uf_prev = None
    hit = {"lub": 0, "inlet": 0, "window": 0}
    for k in range(4 * subcycles):
        # evolve(): one force evaluation per sub-cycle, all sub-cycles of a fluid step see the same runTime().timeIndex()
        tix = 1 + k // subcycles
        if k % subcycles == 0:
            Uf, gamma, gradp, DDtU, curlU = _fields(case, rng)
            e.put_cell_fields(Uf, gamma, gradp)
            e.coupling_time_index(tix)
        st = e.atoms()
        e.compute_fluid_force()
        dg = e.coupling_diag()
        cell = oracle_mod.cell_owner(st["x"], case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
        assert np.array_equal(dg["cell"], cell)
        dia = 2.0 * st["radius"]
        ref = oracle_mod.particle_force(cell, dia, st["v"], v_prev, Uf, gamma, gradp, None, None, DRAG_ERGUN_WENYU,
                                        flags & 31, nub, rhob, np.asarray(case["g"], float), deltaT)
        F = np.ascontiguousarray(ref["F"])
        n0_before = n0.copy()
        oracle_mod.particle_force_extra(cell, st["x"], dia, st["rmass"], st["v"], v_prev, Uf, Uf if Uf_prev is None else Uf_prev,
                                        flags, nub, rhob, deltaT, tix, S, n0, F, inlet, box, 1)
        assert rel_err(dg["F"], F) < 1e-11, (k, rel_err(dg["F"], F))
        Sg, n0g = e.history_state()
        assert rel_err(Sg, S) < 1e-11 and np.abs(n0g - n0).max() < 1e-9
        gap = st["x"][:, 1] - 0.5 * dia
        hit["lub"] += int(((gap < 0.1 * dia) & (gap > 1e-4 * dia)).sum())
        hit["inlet"] += int((st["x"][:, 0] < box[1]).sum())
        hit["window"] += int((n0 != n0_before).sum())
        v_prev = st["v"]
        if k % subcycles == subcycles - 1:
            Uf_prev = Uf
        e.sedi_step(10)
    assert hit["lub"] > 0 and hit["inlet"] > 0 and hit["window"] > 0   # every branch was exercised


def test_cell_owner_graded_stacked_blocks(oracle_mod):
    """SURVEY 8a15 on the mesh of cases/example-cases/BL24-TH1 (three blocks stacked in y, the lowest graded 0.1, cells
    numbered block by block): owner labels bit-exact against the restated interval search, cell volumes through the
    scatter's built-in invariant sum(gamma V) = sum(Vp) (enhancedCloud.C:964-976)"""
    xf, yf, zf, label = cases.blockmesh_stacked((0.0, 0.016, 16, 1.0), [(0.0, 0.008, 20, 0.1), (0.008, 0.012, 40, 1.0), (0.012, 0.016, 40, 1.0)],
                                                (0.0, 0.008, 4, 1.0))
    assert len(yf) == 101 and np.all(np.diff(yf) > 0) and abs((yf[20] - yf[19]) / (yf[1] - yf[0]) - 0.1) < 1e-12
    case = cases.sediment_column(dims=(20, 18, 10), d=5.0e-4, phi=0.35)     # 10 x 9 x 5 mm column inside the 16 x 16 x 8 mm mesh
    e = make_engine(case)
    e.mesh_rectilinear(xf, yf, zf, label)
    assert e.mesh_ncells() == 16 * 100 * 4
    e.coupling_config(DRAG_ERGUN_WENYU, FORCE_DRAG | FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 2e-4)
    e.step(40)
    e.enable_diag(True)
    C = e.mesh_ncells()
    e.put_cell_fields(np.zeros((C, 3)), np.zeros(C), np.zeros((C, 3)))
    e.compute_fluid_force()
    st = e.atoms()
    cell = oracle_mod.cell_owner_rect(st["x"], xf, yf, zf, label)
    assert np.array_equal(e.coupling_diag()["cell"], cell) and (cell >= 0).all()
    assert len(np.unique(cell)) > 300                                     # spread over graded and uniform blocks
    # volumes: tensor cell (i, j, k) -> label
    V = np.zeros(C)
    nx, ny = 16, 100
    for k in range(4):
        for j in range(ny):
            V[label[np.arange(nx) + nx * (j + ny * k)]] = np.diff(xf) * (yf[j + 1] - yf[j]) * (zf[k + 1] - zf[k])
    d = 2.0 * st["radius"]
    g_ref, Ue_ref = oracle_mod.particle_to_eulerian(cell, d, st["v"], V)
    g, Ue = e.scatter_alpha_u()
    assert rel_err(g, g_ref) < 1e-12 and rel_err(Ue, Ue_ref) < 1e-11
    assert abs((g * V).sum() / (np.pi / 6 * (d ** 3).sum()) - 1.0) < 1e-12
    # points outside the mesh are unowned; a point exactly on an interior face belongs to the upper cell
    probe = np.array([[0.004, yf[20], 0.002], [0.004, -1e-9, 0.002], [0.016, 0.001, 0.002]])
    assert list(oracle_mod.cell_owner_rect(probe, xf, yf, zf, label)) == [int(label[4 + 16 * (20 + 100 * 1)]), -1, -1]


def test_smoothing_on_graded_stacked_mesh(oracle_mod):
    """diffusion smoothing on the BL24-TH1 mesh layout (graded + stacked blocks, block-wise labels, the case's own
    smoothDirection diagonal 4 / 2 / 4): volume-weighted PCG on the GPU against the direct solve of the volume-integrated
    finite-volume system; conservation of sum(phi V)"""
    xf, yf, zf, label = cases.blockmesh_stacked((0.0, 0.016, 16, 1.0), [(0.0, 0.008, 20, 0.1), (0.008, 0.012, 40, 1.0), (0.012, 0.016, 40, 1.0)],
                                                (0.0, 0.008, 4, 1.0))
    case = cases.sediment_column(dims=(6, 6, 5), d=5.0e-4, phi=0.35)
    e = make_engine(case)
    e.mesh_rectilinear(xf, yf, zf, label)
    C = e.mesh_ncells()
    rng = np.random.default_rng(5)
    phi = rng.uniform(0.0, 0.6, size=C)
    vec = rng.normal(size=(C, 3))
    V = np.zeros(C)
    nx, ny = 16, 100
    for k in range(4):
        for j in range(ny):
            V[label[np.arange(nx) + nx * (j + ny * k)]] = np.diff(xf) * (yf[j + 1] - yf[j]) * (zf[k + 1] - zf[k])
    b, steps, D = 5.0e-4, 3, (4.0, 2.0, 4.0)       # cases/example-cases/BL24-TH1/constant/cloudProperties:35-36, :60
    e.smooth_config(b, steps, D)
    got_s = e.smooth_field(phi); got_v = e.smooth_field(vec)
    ref_s = oracle_mod.smooth_field_rect(phi, xf, yf, zf, label, b, steps, D)
    ref_v = oracle_mod.smooth_field_rect(vec, xf, yf, zf, label, b, steps, D)
    assert rel_err(got_s, ref_s) < 1e-9 and rel_err(got_v, ref_v) < 1e-9
    assert abs((got_s * V).sum() / (phi * V).sum() - 1.0) < 1e-11
    assert got_s.std() < phi.std() and 0 < e.smooth_last_iters() < 400


def test_create_delete_keep_contact_and_wall_history(oracle_mod):
    """library.cpp:406-621 edits the atom table in place: the other particles keep their contact history
    (FixShearHistory), wall history and per-atom fix arrays.  A particle injected far above the bed and deleted again
    must leave the bed's trajectory where an undisturbed run puts it (shear springs are loaded: a run that dropped the
    history at the injection ends measurably elsewhere)."""
    import os
    case = cases.fluidized_bed(dims=(8, 8, 8), vjit=0.2)          # lively bed: tangential springs get loaded
    top = case["box_hi"]
    pos = np.array([[0.5 * top[0], 0.97 * top[1], 0.5 * top[2]]])

    def run(disturb, reset=False):
        if reset:
            os.environ["SEDI_INJECT_RESET"] = "1"
        try:
            e = make_engine(case)
            e.command("group active type 1")
            m = float(case["rho"][0] * np.pi / 6.0 * case["diam"][0] ** 3)
            e.put_local_info(np.tile([0.0, 0.2 * 9.8 * m, 0.0], (len(case["tag"]), 1)), case["tag"])
            e.step(150)
            if disturb:
                e.create_particle(pos, [5001.0], 4.0e-4, 2500.0, 1, (0.0, 0.0, 0.0))
                assert e.get_local_n() == len(case["tag"]) + 1
            e.step(60)
            if disturb:
                e.delete_particle([5001])
                assert e.get_local_n() == len(case["tag"])
            e.step(150)
            st = e.atoms()
            pr = e.pairs()
            ws = e.wall_shear(1)
            e.close()
            return st, pr, ws
        finally:
            os.environ.pop("SEDI_INJECT_RESET", None)

    ref, pref, wref = run(False)
    got, pgot, wgot = run(True)
    L = float(np.abs(case["box_hi"] - case["box_lo"]).max())
    vmax = np.abs(ref["v"]).max()
    assert np.array_equal(got["tag"], ref["tag"])
    ex, ev = np.abs(got["x"] - ref["x"]).max() / L, np.abs(got["v"] - ref["v"]).max() / vmax
    assert ex < 1e-12 and ev < 1e-12, (ex, ev)      # measured: bitwise identical
    assert rel_err(wgot, wref) < 1e-6 and np.abs(wref).max() > 0
    # loaded springs exist, and they are the same springs
    key = lambda p: np.lexsort((p["tj"], p["ti"]))
    a, b = key(pref), key(pgot)
    assert np.array_equal(pref["ti"][a], pgot["ti"][b]) and np.array_equal(pref["tj"][a], pgot["tj"][b])
    assert np.abs(pref["shear"]).max() > 0 and rel_err(pgot["shear"][b], pref["shear"][a]) < 1e-6
    # control: the old behaviour (history dropped at the edit) does not pass this bar
    lost, _, _ = run(True, reset=True)
    assert np.abs(lost["v"] - ref["v"]).max() / vmax > 1e-5


def test_restart_resumes_exactly(oracle_mod, tmp_path):
    """`write_restart` / `read_restart` / `restart N file` (per-atom state of FixWallGranFix::pack_restart
    fix_wall_granFix.cpp:750-777, fix fdrag's arrays, FixShearHistory): a run resumed from the file ends bitwise where the
    uninterrupted run ends -- contact history, wall history and the stored forces all travel."""
    import os
    from sedifoam_b200 import Lammps
    case = cases.fluidized_bed(dims=(8, 8, 8), vjit=0.2)
    old = os.environ.get("SEDI_DUMP_DIR")
    os.environ["SEDI_DUMP_DIR"] = str(tmp_path)
    try:
        a = make_engine(case)
        m = float(case["rho"][0] * np.pi / 6.0 * case["diam"][0] ** 3)
        a.put_local_info(np.tile([0.0, 0.2 * 9.8 * m, 0.0], (len(case["tag"]), 1)), case["tag"])
        a.command("restart 100 bed.*.rst")
        a.step(150)
        a.command("restart 0")
        a.command("write_restart bed_150.rst")
        a.step(150)
        sa, pa, wa = a.atoms(), a.pairs(), a.wall_shear(1)
        a.close()
        assert os.path.exists(tmp_path / "bed.100.rst") and not os.path.exists(tmp_path / "bed.200.rst")
        for fname, nmore in (("bed_150.rst", 150), ("bed.100.rst", 200)):
            b = Lammps()
            b.command("atom_style sphere")
            b.command("read_restart " + fname)
            b.commands(case["script"])
            b.step(nmore)
            sb, pb, wb = b.atoms(), b.pairs(), b.wall_shear(1)
            b.close()
            assert np.array_equal(sb["tag"], sa["tag"])
            assert np.array_equal(sb["x"], sa["x"]) and np.array_equal(sb["v"], sa["v"]) and np.array_equal(sb["omega"], sa["omega"]), fname
            key = lambda p: np.lexsort((p["tj"], p["ti"]))
            ka, kb = key(pa), key(pb)
            assert np.array_equal(pa["ti"][ka], pb["ti"][kb]) and np.array_equal(pa["shear"][ka], pb["shear"][kb]) and np.abs(pa["shear"]).max() > 0
            assert np.array_equal(wa, wb) and np.abs(wa).max() > 0
    finally:
        if old is None:
            os.environ.pop("SEDI_DUMP_DIR", None)
        else:
            os.environ["SEDI_DUMP_DIR"] = old


def _read_foam_field(path):
    """entries of an OpenFOAM ASCII IOField file written by sedi_write_lagrangian"""
    txt = open(path).read()
    assert "FoamFile" in txt and "format      ascii;" in txt
    body = txt[txt.index("// * * *"):]
    lines = [ln.strip() for ln in body.splitlines()[1:] if ln.strip() and not ln.startswith("//")]
    n = int(lines[0])
    assert lines[1] == "(" and lines[2 + n] == ")"
    return n, lines[2:2 + n], txt


def test_lagrangian_fields_openfoam_layout(oracle_mod, tmp_path):
    """softParticle::writeFields (softParticleIO.C:157-197): positions with the owner cell, d, tag, lmpCpuId, type, U, ensembleU
    (+ density, n0 that readFields :113-152 requires) as OpenFOAM ASCII IOFields; values equal the device state.  No
    OpenFOAM here to read them back: the layout follows the IOField / Cloud positions text format (parity unpinned)."""
    case = cases.fluidized_bed(dims=(6, 7, 5), vjit=0.05)
    e = make_engine(case)
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.step(40)
    e.locate()
    e.write_lagrangian(tmp_path, "0.1/lagrangian/cloud")
    st = e.atoms()
    n = len(st["tag"])
    cell = oracle_mod.cell_owner(st["x"], case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    m, rows, txt = _read_foam_field(tmp_path / "positions")
    assert m == n and "class       Cloud<softParticle>;" in txt and 'location    "0.1/lagrangian/cloud";' in txt
    got = np.array([[float(v) for v in r.replace("(", "").replace(")", "").split()] for r in rows])
    assert np.allclose(got[:, :3], st["x"], rtol=1e-14, atol=0) and np.array_equal(got[:, 3].astype(int), cell)
    for name, cls, want in (("d", "scalarField", 2.0 * st["radius"]), ("tag", "labelField", st["tag"]), ("type", "labelField", st["type"]),
                            ("lmpCpuId", "labelField", np.zeros(n)), ("n0", "scalarField", np.zeros(n)),
                            ("density", "scalarField", 3.0 * st["rmass"] / (4.0 * 3.14159265358917323846 * st["radius"] ** 3))):
        m, rows, txt = _read_foam_field(tmp_path / name)
        assert m == n and ("class       %s;" % cls) in txt
        assert np.allclose(np.array([float(r) for r in rows]), want, rtol=1e-14, atol=0), name
    for name, want in (("U", st["v"]), ("ensembleU", np.zeros((n, 3)))):
        m, rows, txt = _read_foam_field(tmp_path / name)
        assert m == n and "class       vectorField;" in txt
        got = np.array([[float(v) for v in r.strip("()").split()] for r in rows])
        assert np.allclose(got, want, rtol=1e-14, atol=0), name


def test_particle_outside_the_mesh_is_ignored_by_the_coupling(oracle_mod):
    """the reference deletes a particle that hits a non-processor patch from the Foam cloud while it lives on in LAMMPS
    (softParticle.C:177-184) -- after which its own lammps_put_local_info indexes out of range.  Decision
    here: such a particle gets owner cell -1, no fluid force, and takes no part in the scatters; the DEM keeps it."""
    case = cases.fluidized_bed(dims=(6, 7, 5), vjit=0.0)
    hi = case["mesh_hi"].copy()
    hi[1] = case["x"][:, 1].max() - 1.6 * case["diam"][0]        # the mesh ends below the two top layers
    nc = case["mesh_n"].copy(); nc[1] = max(1, nc[1] // 2)
    e = make_engine(case)
    e.mesh_box(case["mesh_lo"], hi, nc)
    e.coupling_config(DRAG_ERGUN_WENYU, FORCE_DRAG | FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 2.0e-4)
    C = int(np.prod(nc))
    e.put_cell_fields(np.tile([0.0, 0.3, 0.0], (C, 1)), np.full(C, 0.4), np.tile([0.0, -9800.0, 0.0], (C, 1)))
    e.setup()
    e.enable_diag(True)
    e.compute_fluid_force()
    dg = e.coupling_diag()
    st = e.atoms()
    out = st["x"][:, 1] >= hi[1]
    assert out.sum() == 2 * 6 * 5 and np.array_equal(dg["cell"] < 0, out)
    assert np.all(dg["F"][out] == 0.0) and np.all(np.abs(dg["F"][~out, 1]) > 0.0)
    g, Ue = e.scatter_alpha_u()
    V = np.prod((hi - case["mesh_lo"]) / nc)
    assert abs((g * V).sum() / (np.pi / 6 * (2 * st["radius"][~out]) ** 3).sum() - 1.0) < 1e-12     # only the inside particles are scattered
    e.step(20)
    assert e.get_local_n() == len(case["tag"])                                                  # the DEM keeps every particle


def test_reference_printouts_and_timers(oracle_mod):
    """what lammpsFoam.C / writeCPUTime.H read besides the fields (enhancedCloud.H:206-249): averageInfo, the two
    conservation printouts ("total F before / after", "total U solid before / after", enhancedCloud.C:395-435, 936-976)
    and the timers.  Sums against numpy on the same state; smoothing conserves them (the reference's built-in invariant)."""
    case = cases.fluidized_bed(dims=(8, 9, 7), vjit=0.1)
    e = make_engine(case)
    e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    e.coupling_config(DRAG_ERGUN_WENYU, FORCE_DRAG | FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 2.0e-4)
    e.smooth_config(2.0 * case["diam"][0] * 2, 2)
    rng = np.random.default_rng(3)
    Uf, gamma, gradp, _, _ = _fields(case, rng)
    e.put_cell_fields(Uf, gamma, gradp)
    e.enable_conservation_sums(True)
    e.step(30)
    st = e.atoms()
    vol = np.pi / 6.0 * (2.0 * st["radius"]) ** 3
    ai = e.average_info()
    assert abs(ai["totalVolume"] / vol.sum() - 1.0) < 1e-13
    assert np.allclose(ai["totalVel"], (vol[:, None] * st["v"]).sum(axis=0), rtol=1e-11, atol=1e-25)
    assert np.allclose(ai["averageVel"], ai["totalVel"] / ai["totalVolume"], rtol=1e-13)
    g, Ue = e.scatter_alpha_u()
    A, Om = e.calc_tc()
    s = e.conservation_sums()
    scaleU = np.abs(vol[:, None] * st["v"]).sum()
    assert np.abs(s["U_before"] - (vol[:, None] * st["v"]).sum(axis=0)).max() < 1e-11 * scaleU
    assert np.abs(s["U_after"] - s["U_before"]).max() < 1e-9 * scaleU          # diffusion smoothing conserves sum(gamma Ue V)
    V = np.prod((case["mesh_hi"] - case["mesh_lo"]) / case["mesh_n"])
    assert np.abs(s["U_after"] - (Ue * (g * V)[:, None]).sum(axis=0)).max() < 1e-11 * scaleU
    scaleF = np.abs(A * V * (1 - g)[:, None]).sum()
    assert np.abs(s["F_after"] - (A * V * (1 - g)[:, None]).sum(axis=0)).max() < 1e-11 * scaleF
    assert np.abs(s["F_after"] - s["F_before"]).max() < 1e-9 * scaleF and scaleF > 0
    t = e.timers()
    assert t["cpuTimeSplit"][4] > 0.0 and t["particleMoveTime"] > 0.0 and t["diffusionTimeCount"][1] > 0.0
    assert np.all(t["cpuTimeSplit"][:3] == 0.0)                                 # no all-to-alls here


def test_scatter_is_bitwise_reproducible_and_order_independent():
    """gamma / Ue / Asrc are summed per cell in a fixed order (no floating-point atomics): two runs give identical bits, through
    neighbour rebuilds that re-order the rows, and a whole coupled loop (force -> sub-steps -> scatter) is reproducible"""
    case = cases.settled_bed(columns=(2, 2), column="column_256x4.npz")
    Uf0, gamma0, gradp0 = cases.uniform_fields(case)
    out = []
    for rep in range(2):
        e = make_engine(case)
        e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
        e.coupling_config(DRAG_ERGUN_WENYU, FORCE_DRAG | FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 2.0e-4)
        e.put_cell_fields(Uf0, gamma0, gradp0)
        e.setup()
        res = []
        for k in range(3):
            g, Ue = e.scatter_alpha_u()
            e.put_cell_fields(Uf0, g, gradp0)
            e.compute_fluid_force()
            e.sedi_step(40)
            if k == 1:
                e.force_rebuild()
            A, _ = e.calc_tc()
            res.append((g.copy(), Ue.copy(), A.copy()))
        st = e.atoms()
        o = np.argsort(st["tag"])
        res.append((st["x"][o].copy(), st["v"][o].copy(), st["omega"][o].copy()))
        out.append(res)
        e.close()
    for a, b in zip(out[0], out[1]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    assert np.abs(out[0][0][2]).max() > 0
