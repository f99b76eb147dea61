#!/usr/bin/env python
"""Generator of the benchmark's settled random bed (sedifoam_b200/data/column_*.npz).

A column of random tiles (sedifoam_b200/packing.py: overlap-removal packing at phi = 0.58) stacked on a granular floor,
periodic in x and z, is run to rest under gravity and the benchmark's fluid force (ErgunWenYu drag of the prescribed
upflow + pressure-gradient force, together 0.45 of the weight, evaluated by the bench's own coupled loop) with the DEM
itself -- gran/hertzFix/history, the bench's own kn / e / mu / dt.  The result (positions, residual velocities and
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
    ap.add_argument("--force", type=float, default=0.447, help="--oracle: constant fluid force as a fraction of the weight")
    args = ap.parse_args()
    import sedifoam_b200 as sb
    from sedifoam_b200 import cases
    case = cases.settling_column(tile_n=args.tile_n, height=args.height)
    n = len(case["tag"])
    if args.oracle:
        from oracle import pyoracle
        sim = pyoracle.Oracle("reference" if pyoracle.have_reference() else "port")
    else:
        sim = sb.Lammps(device=0)
    cases.apply(case, sim)
    case["fluid_force_over_weight"] = args.force
    if args.oracle:   # constant fluid force (the CPU oracle has no coupling loop)
        sim.setup()
        sim.put_fdrag(cases.bench_fluid_force(case), case["tag"])
    else:             # the bench's own coupled loop: scatter -> ErgunWenYu + pressure-gradient force -> 100 DEM sub-steps
        sim.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
        sim.coupling_config(sb.DRAG_ERGUN_WENYU, sb.FORCE_DRAG | sb.FORCE_PGRAD, case["nub"], case["rhob"], case["g"], 100 * case["dt"])
        Uf, gamma, gradp = cases.uniform_fields(case)
        sim.put_cell_fields(Uf, gamma, gradp)
        sim.step(0)
    t0 = time.time()
    done = 0
    log = []
    while done < args.steps:
        if args.oracle:
            sim.run(args.chunk)
        else:
            for _ in range(args.chunk // 100):
                sim.scatter_alpha_u(device_only=True)
                sim.compute_fluid_force()
                sim.sedi_step(100)
        done += args.chunk
        a = sim.atoms()
        speed = np.sqrt((a["v"] ** 2).sum(1))
        vmax = float(speed.max())
        p = sim.pairs("gran", history=True) if args.oracle else None
        if args.oracle:
            touching = int(p[2].sum())
        else:
            pr = sim.pairs()
            touching = int(pr["touch"][pr["gran"].astype(bool)].sum()) // 2
        rec = dict(steps=done, vmax=vmax, vmedian=float(np.median(speed)), v99=float(np.quantile(speed, 0.99)), ytop=float(a["x"][:, 1].max()), ymean=float(a["x"][:, 1].mean()), touching_pairs_per_particle=touching / n,
                   rebuilds=int(sim.stat("nbuilds")), wall_s=time.time() - t0)
        log.append(rec)
        print(json.dumps(rec), flush=True)
        if rec["v99"] < args.vtol:   # rattlers keep bouncing inside their cages: judge by the 99th percentile
            break
    a = sim.atoms()   # sorted by tag
    Lx = float(case["box_hi"][0] - case["box_lo"][0]); Lz = float(case["box_hi"][2] - case["box_lo"][2])
    x = a["x"].copy()
    x[:, 0] = np.mod(x[:, 0], Lx); x[:, 2] = np.mod(x[:, 2], Lz)
    x[x[:, 0] >= Lx, 0] = 0.0; x[x[:, 2] >= Lz, 2] = 0.0
    out = args.out or os.path.join(ROOT, "sedifoam_b200", "data", "column_%dx%d.npz" % (args.tile_n, args.height))
    fow = args.force
    if not args.oracle:   # the force the coupled loop actually applies, as a fraction of the weight
        sim.scatter_alpha_u(device_only=True)
        sim.enable_diag(True)
        sim.compute_fluid_force()
        dg = sim.coupling_diag()
        m = case["rho"] * np.pi / 6.0 * case["diam"] ** 3
        fow = float(np.mean(dg["F"][:, 1] / (m * 9.8)))
    meta = dict(tile_n=args.tile_n, height=args.height, steps=done, vmax=log[-1]["vmax"], v99=log[-1]["v99"], fluid_force_over_weight=fow,
                Uf=list(case["Uf"]), touching_pairs_per_particle=log[-1]["touching_pairs_per_particle"],
                engine="oracle" if args.oracle else "cuda", d=float(case["diam"][0]), kn=1.0e7, e=0.9, mu=0.4, dt=case["dt"])
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
