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
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.stdout.flush()
    os._exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
