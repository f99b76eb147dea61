"""Multi-GPU parity check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py
Every rank feeds the SAME global case to its engine (each keeps the particles of its own brick, like LAMMPS' read_data),
runs DEM sub-steps with NCCL ghost exchange and migration, and rank 0 compares the gathered state with the CPU oracle
run on the whole box: positions / velocities within 1e-6 relative, identical rebuild counts, identical global pair count.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sedifoam_b200 as sb  # noqa: E402
from sedifoam_b200 import cases  # noqa: E402

SCEN = {
    "bed_walls": (lambda: cases.fluidized_bed(dims=(16, 12, 14), vjit=0.05), 300, None),
    "column_periodic": (lambda: cases.sediment_column(dims=(14, 20, 12), phi=0.45, jitter_frac=0.08), 400, None),
    "column_periodic_x2": (lambda: cases.sediment_column(dims=(14, 20, 12), phi=0.45, jitter_frac=0.08), 300, "x"),
    "cohesive": (lambda: cases.cohesive_shear_bed(dims=(14, 8, 12), opt=1), 200, None),
    "settled_random": (lambda: cases.settled_bed(columns=(3, 2), column="column_256x4.npz"), 300, None),
}


def _gather(eng, world, keys=("x", "v", "omega")):
    st = eng.atoms()
    parts = [None] * world
    dist.all_gather_object(parts, {k: st[k] for k in ("tag",) + tuple(keys)})
    tag = np.concatenate([p["tag"] for p in parts])
    o = np.argsort(tag)
    return {k: np.concatenate([p[k] for p in parts])[o] for k in ("tag",) + tuple(keys)}


ERRS = []


def _close(a, b, tol=1e-8):
    if not np.array_equal(a["tag"], b["tag"]):
        ERRS.append("tags differ (%d vs %d)" % (len(a["tag"]), len(b["tag"])))
        return False
    e = [float(np.abs(a[k] - b[k]).max() / max(np.abs(b[k]).max(), 1e-300)) for k in ("x", "v", "omega")]
    ERRS.append("x %.1e v %.1e w %.1e" % tuple(e))
    return all(v <= tol for v in e)


def fd_bed(case):
    return cases.bench_fluid_force(case)


def features(rank, world, lr):
    """multi-GPU features without a CPU oracle of their own: the several-GPU engine against the SAME engine on one GPU
    (which the single-GPU suite checks against the reference objects): restart written as per-rank parts and read back,
    particle injection / deletion keeping contact and wall history, and the history-force state migrating with its
    particle.  Tolerance 1e-8 relative (ghost partners of a bin arrive in no fixed order, so sums differ in the last bits, and a bed at
    rest amplifies that in its tiny velocities: the multi-GPU engine is within 3e-11 of the oracle there)."""
    import tempfile
    ok = True
    tmp = [tempfile.mkdtemp(prefix="sedi_mg_") if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    os.environ["SEDI_DUMP_DIR"] = tmp[0]

    def engines(case, grid=None):
        multi = sb.Lammps(device=lr)
        cases.apply(case, multi)
        uid = [sb.Lammps.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        multi.comm_init(rank, world, uid[0], grid)
        single = sb.Lammps(device=lr)
        cases.apply(case, single)
        return multi, single

    # ---- restart: run, write parts, continue; read the parts into fresh engines, continue; equal end states
    case = cases.settled_bed(columns=(3, 2), column="column_256x4.npz")
    m, s1 = engines(case)
    fd = cases.bench_fluid_force(case)
    for e in (m, s1):
        e.setup()
    loc = m.get_local_info(); m.put_local_info(fd[loc["tag"] - 1], loc["tag"]); s1.put_local_info(fd, case["tag"])
    m.step(120); s1.step(120)
    m.command("write_restart mg.rst"); s1.command("write_restart sg.rst")
    m.step(80); s1.step(80)
    a_multi, a_single = _gather(m, world), s1.atoms()
    good = _close(a_multi, a_single)
    m.close(); s1.close()
    def fresh(fname, multi):   # a new engine that knows nothing but the restart file and the script
        e = sb.Lammps(device=lr)
        e.command("atom_style sphere")
        e.command("read_restart " + fname)      # box, boundary, atoms and per-atom state come from the file
        if multi:
            uid = [sb.Lammps.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            e.comm_init(rank, world, uid[0], None)
        e.commands(case["script"])
        return e

    m2, s2 = fresh("mg.rst", True), fresh("sg.rst", False)
    m2.step(80); s2.step(80)
    b_multi, b_single = _gather(m2, world), s2.atoms()
    good = good and _close(b_multi, a_multi) and _close(b_single, a_single) and _close(b_multi, b_single)   # (a list rebuilt at another time orders a row differently: last-bit differences)
    # a restart written by 2 GPUs read by 1 GPU
    s3 = fresh("mg.rst", False)
    s3.step(80)
    good = good and _close(s3.atoms(), a_single)
    if rank == 0:
        print("[mgpu restart] parts written / read on %d GPUs, resumed runs agree (%s) -> %s" % (world, "; ".join(ERRS), "OK" if good else "FAIL"), flush=True)
    del ERRS[:]
    ok = ok and good
    for e in (m2, s2, s3):
        e.close()

    # ---- injection / deletion with history kept: an intruder far above the bed comes and goes, the bed must not notice
    m, s1 = engines(case)
    u, _ = engines(case)
    for e in (m, s1, u):
        e.setup()
    for e in (m, u):
        loc = e.get_local_info(); e.put_local_info(fd[loc["tag"] - 1], loc["tag"])
    s1.put_local_info(fd, case["tag"])
    for e in (m, s1, u):
        e.step(100)
    top = float(case["box_hi"][1]) - 4.0 * float(case["diam"][0])
    newtag = int(case["tag"].max()) + 1
    pos = np.array([[0.5 * case["box_hi"][0], top, 0.5 * case["box_hi"][2]]])
    for e in (m, s1):
        e.create_particle(pos, [newtag], float(case["diam"][0]), 2650.0, 1, [0.0, 0.0, 0.0])
        e.step(30)
        e.delete_particle([newtag])
        e.step(70)
    u.step(100)
    gm, gu, gs = _gather(m, world), _gather(u, world), s1.atoms()
    good = _close(gm, gu) and _close(gm, gs) and len(gm["tag"]) == len(case["tag"])
    if rank == 0:
        print("[mgpu inject] create + delete on %d GPUs leaves the bed where the undisturbed run puts it (%s) -> %s" % (world, "; ".join(ERRS), "OK" if good else "FAIL"), flush=True)
    del ERRS[:]
    ok = ok and good
    for e in (m, s1, u):
        e.close()

    # ---- history force (reduced-order Basset term): per-particle state migrates between bricks
    case = cases.sediment_column(dims=(14, 20, 12), phi=0.45, jitter_frac=0.08)
    rng = np.random.default_rng(5)
    case["v"] = rng.normal(scale=0.15, size=case["x"].shape)     # particles cross the brick faces
    m, s1 = engines(case, (world, 1, 1))
    C = int(np.prod(case["mesh_n"]))
    for e in (m, s1):
        e.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
        e.coupling_config(sb.DRAG_ERGUN_WENYU, sb.FORCE_DRAG | sb.FORCE_PGRAD | sb.FORCE_HISTORY, case["nub"], case["rhob"], case["g"], 4.0e-4)
        e.setup()
    arrivals = 0
    for k in range(6):
        Uf = rng.normal(scale=0.05, size=(C, 3)); gam = rng.uniform(0.2, 0.5, size=C); gp = rng.normal(scale=50.0, size=(C, 3))
        bufs = [Uf, gam, gp]
        dist.broadcast_object_list(bufs, src=0)
        for e in (m, s1):
            e.put_cell_fields(bufs[0], bufs[1], bufs[2])
            e.coupling_time_index(1 + k)
            e.compute_fluid_force()
            e.sedi_step(200)
        arrivals += m.comm_stat("arrivals")
    good = _close(_gather(m, world), s1.atoms())
    hs_m = m.history_state(); hs_s = s1.history_state()
    parts = [None] * world
    dist.all_gather_object(parts, dict(tag=m.atoms()["tag"], S=hs_m[0], n0=hs_m[1]))
    tag = np.concatenate([p["tag"] for p in parts]); o = np.argsort(tag)
    S = np.concatenate([p["S"] for p in parts])[o]; n0 = np.concatenate([p["n0"] for p in parts])[o]
    eS = float(np.abs(S - hs_s[0]).max() / max(np.abs(hs_s[0]).max(), 1e-300)); en0 = float(np.abs(n0 - hs_s[1]).max())
    ERRS.append("sumDeltaFb %.1e n0 %.1e max|S| %.2e" % (eS, en0, np.abs(hs_s[0]).max()))
    good = good and eS <= 1e-8 and en0 < 1e-6 and np.abs(hs_s[0]).max() > 0
    tot = torch.tensor([arrivals], device="cuda"); dist.all_reduce(tot)
    good = good and int(tot.item()) > 0
    if rank == 0:
        print("[mgpu history force] state follows %d migrated particles, forces and state equal the single-GPU run (%s) -> %s" % (int(tot.item()), "; ".join(ERRS), "OK" if good else "FAIL"), flush=True)
    del ERRS[:]
    ok = ok and good
    m.close(); s1.close()

    # ---- bootstrap without a host call: lammps_open + script + the pre-run queries + lammps_step, as softParticleCloud::initLammps
    # drives the reference (softParticleCloud.C:57-206); the bricks and the NCCL communicator come up by themselves
    case = cases.settled_bed(columns=(3, 2), column="column_256x4.npz")
    os.environ["SEDI_AUTO_COMM"] = "1"
    a = sb.Lammps(device=lr)
    cases.apply(case, a)
    nglobal = a.get_global_n()
    npr = a.get_initial_np(world)
    info = a.get_initial_info()
    dom = a.get_local_domain()
    a.step(0)
    os.environ.pop("SEDI_AUTO_COMM")
    loc = a.get_local_info(); a.put_local_info(fd_bed(case)[loc["tag"] - 1], loc["tag"])
    a.step(150)
    ref = sb.Lammps(device=lr)
    cases.apply(case, ref)
    ref.step(0); ref.put_local_info(fd_bed(case), case["tag"]); ref.step(150)
    inside = np.all((info["x"][:, 0] >= dom[0] - 1e-12) & (info["x"][:, 0] < dom[1] + 1e-12))
    good = nglobal == len(case["tag"]) and int(npr.sum()) == nglobal and npr[rank] == len(info["tag"]) and a.comm_stat("links") > 0 and bool(inside)
    good = _close(_gather(a, world), ref.atoms()) and good
    flag = torch.tensor([1 if good else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    good = int(flag.item()) == 1
    if rank == 0:
        print("[mgpu bootstrap] lammps_open + script + initLammps queries on %d ranks, no sedi_comm_init: np = %s (%s) -> %s" % (world, npr.tolist(), "; ".join(ERRS), "OK" if good else "FAIL"), flush=True)
    del ERRS[:]
    ok = ok and good
    a.close(); ref.close()
    return ok


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ok = True
    for name, (mk, nsteps, split) in SCEN.items():
        case = mk()
        n = len(case["tag"])
        eng = sb.Lammps(device=lr)
        cases.apply(case, eng)
        uid = [sb.Lammps.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        grid = None
        if split == "x":
            grid = (world, 1, 1)
        eng.comm_init(rank, world, uid[0], grid)
        rng = np.random.default_rng(11)
        m = case["rho"] * np.pi / 6.0 * case["diam"] ** 3
        fd = rng.normal(scale=2.0, size=(n, 3)) * m[:, None]
        eng.setup()
        loc = eng.get_local_info()
        mine = loc["tag"] - 1
        eng.put_local_info(fd[mine], loc["tag"])
        eng.step(nsteps // 2)
        loc = eng.get_local_info()   # ownership may have changed
        eng.put_local_info(fd[loc["tag"] - 1], loc["tag"])
        eng.step(nsteps - nsteps // 2)
        st = eng.atoms()
        stats = dict(nlocal=eng.get_local_n(), nbuilds=eng.stat("nbuilds"), pair_evals=eng.stat("pair_evals"), ghosts=eng.comm_stat("ghost_rows"),
                     links=eng.comm_stat("links"), halo=eng.comm_stat("halo_calls"), p2p=eng.comm_stat("p2p"), gran_entries=eng.stat("gran_entries"))
        gathered = [None] * world
        dist.all_gather_object(gathered, dict(tag=st["tag"], x=st["x"], v=st["v"], omega=st["omega"], stats=stats, dom=eng.get_local_domain(),
                                              nglobal=eng.get_global_n()))
        if rank == 0:
            from oracle import pyoracle
            o = pyoracle.Oracle("reference" if pyoracle.have_reference() else "port")
            cases.apply(case, o)
            o.setup()
            o.put_fdrag(fd, case["tag"])
            o.run(nsteps // 2); o.put_fdrag(fd, case["tag"]); o.run(nsteps - nsteps // 2)
            a = o.atoms()
            tag = np.concatenate([g["tag"] for g in gathered]); x = np.concatenate([g["x"] for g in gathered]); v = np.concatenate([g["v"] for g in gathered])
            w = np.concatenate([g["omega"] for g in gathered])
            oo = np.argsort(tag)
            L = np.abs(case["box_hi"] - case["box_lo"]).max()
            good = len(tag) == n and np.array_equal(tag[oo], case["tag"])
            ex = ev = ew = float("nan")
            if good:
                # periodic wrapping happens at rebuilds; compare modulo the period
                dx = x[oo] - a["x"]
                for d in range(3):
                    if case["periodic"][d] == "p":
                        prd = case["box_hi"][d] - case["box_lo"][d]
                        dx[:, d] -= prd * np.round(dx[:, d] / prd)
                ex = np.abs(dx).max() / L
                ev = np.abs(v[oo] - a["v"]).max() / np.abs(a["v"]).max()
                ew = np.abs(w[oo] - a["omega"]).max() / max(np.abs(a["omega"]).max(), 1e-300)
                good = ex < 1e-6 and ev < 1e-6 and ew < 1e-6
            nb = [g["stats"]["nbuilds"] for g in gathered]
            pe = sum(g["stats"]["pair_evals"] for g in gathered)
            good = good and all(b == o.stat("nbuilds") for b in nb) and all(g["nglobal"] == n for g in gathered)
            print("[mgpu %s] ranks=%d n=%d nlocal=%s ghosts=%s links=%s rebuilds=%s (oracle %d) halo_calls=%d p2p=%d err x=%.2e v=%.2e w=%.2e pair_evals=%d (oracle %d) -> %s"
                  % (name, world, n, [g["stats"]["nlocal"] for g in gathered], [g["stats"]["ghosts"] for g in gathered],
                     [g["stats"]["links"] for g in gathered], nb, o.stat("nbuilds"), gathered[0]["stats"]["halo"], gathered[0]["stats"]["p2p"], ex, ev, ew, pe,
                     o.stat("pair_evals"), "OK" if good else "FAIL"), flush=True)
            ok = ok and good
        eng.close()
        dist.barrier()
    ok = features(rank, world, lr) and ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
