#!/usr/bin/env python
"""Generator of the benchmark's settled random bed (sedifoam_b200/data/column_*.npz).

A column of random tiles (sedifoam_b200/packing.py: overlap-removal packing at phi = 0.58) stacked on a granular floor,
periodic in x and z, is run to rest under gravity and the benchmark's fluid force (0.3 of the weight, upwards) with the
DEM itself -- gran/hertzFix/history, the bench's own kn / e / mu / dt.  The result (positions, residual velocities and
spins, by tag) is periodic in x and z, so cases.settled_bed() builds beds of any size by repeating it.

    gpurun -- python tools/make_settled_column.py                 # CUDA engine (minutes of DEM time in seconds)
    python tools/make_settled_column.py --oracle --tile-n 64 ...  # the CPU oracle (reference objects): small columns only

The CUDA engine is used as a tool here (it is parity-checked against the reference objects by tests/); the output is
synthetic input data for both arms of bench.py and for the tests, not a result."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tile-n", type=int, default=5000)
    ap.add_argument("--height", type=int, default=8)
    ap.add_argument("--steps", type=int, default=600000, help="upper bound of DEM steps")
    ap.add_argument("--chunk", type=int, default=20000)
    ap.add_argument("--vtol", type=float, default=2.0e-4, help="stop when the largest particle speed is below this (m/s)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--oracle", action="store_true")
    args = ap.parse_args()
    from sedifoam_b200 import cases
    case = cases.settling_column(tile_n=args.tile_n, height=args.height)
    n = len(case["tag"])
    if args.oracle:
        from oracle import pyoracle
        sim = pyoracle.Oracle("reference" if pyoracle.have_reference() else "port")
    else:
        import sedifoam_b200 as sb
        sim = sb.Lammps(device=0)
    cases.apply(case, sim)
    sim.setup() if args.oracle else sim.step(0)
    fd = cases.bench_fluid_force(case)
    if args.oracle:
        sim.put_fdrag(fd, case["tag"])
    else:
        sim.put_local_info(fd, case["tag"])
    t0 = time.time()
    done = 0
    log = []
    while done < args.steps:
        sim.run(args.chunk) if args.oracle else sim.step(args.chunk)
        done += args.chunk
        a = sim.atoms()
        vmax = float(np.sqrt((a["v"] ** 2).sum(1)).max())
        p = sim.pairs("gran", history=True) if args.oracle else None
        if args.oracle:
            touching = int(p[2].sum())
        else:
            pr = sim.pairs()
            touching = int(pr["touch"][pr["gran"].astype(bool)].sum()) // 2
        rec = dict(steps=done, vmax=vmax, ytop=float(a["x"][:, 1].max()), ymean=float(a["x"][:, 1].mean()), touching_pairs_per_particle=touching / n,
                   rebuilds=int(sim.stat("nbuilds")), wall_s=time.time() - t0)
        log.append(rec)
        print(json.dumps(rec), flush=True)
        if vmax < args.vtol:
            break
    a = sim.atoms()   # sorted by tag
    Lx = float(case["box_hi"][0] - case["box_lo"][0]); Lz = float(case["box_hi"][2] - case["box_lo"][2])
    x = a["x"].copy()
    x[:, 0] = np.mod(x[:, 0], Lx); x[:, 2] = np.mod(x[:, 2], Lz)
    x[x[:, 0] >= Lx, 0] = 0.0; x[x[:, 2] >= Lz, 2] = 0.0
    out = args.out or os.path.join(ROOT, "sedifoam_b200", "data", "column_%dx%d.npz" % (args.tile_n, args.height))
    meta = dict(tile_n=args.tile_n, height=args.height, steps=done, vmax=log[-1]["vmax"], touching_pairs_per_particle=log[-1]["touching_pairs_per_particle"],
                engine="oracle" if args.oracle else "cuda", d=float(case["diam"][0]), kn=1.0e7, e=0.9, mu=0.4, dt=case["dt"], fluid_force="0.3 m g up")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, x=x, v=a["v"].astype(np.float32), omega=a["omega"].astype(np.float32), Lx=Lx, Lz=Lz, d=float(case["diam"][0]),
                        meta=json.dumps(meta))
    print("wrote", out, json.dumps(meta), flush=True)
    mirror = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(mirror) and not args.oracle:
        import shutil
        shutil.copy(out, os.path.join(mirror, os.path.basename(out)))


if __name__ == "__main__":
    main()
