#!/usr/bin/env python
"""Time the fused DEM sub-step kernel of several library builds / runtime switches on the bench bed (one process per
variant, no torch) and check that every variant ends in the same state bit for bit.

  python tools/kstep_sweep.py                       # all libraries under build_variants/ plus the in-tree one
  python tools/kstep_sweep.py --one                 # worker: current environment, prints one JSON line

Variants are shared libraries (SEDI_B200_LIB) crossed with environment switches given as name=ENV1=v1,ENV2=v2."""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(dims, steps, warm, substeps, couple=True, bed="settled", cfg=2, size=1.0):
    import numpy as np
    import sedifoam_b200 as sb
    from sedifoam_b200 import cases
    import bench
    case = cases.fluidized_bed(dims=dims) if (bed == "lattice" and cfg == 2) else bench.build_case(cfg, bed, 1, 0, "weak", size=size)
    eng = sb.Lammps(device=0)
    cases.apply(case, eng)
    eng.mesh_box(case["mesh_lo"], case["mesh_hi"], case["mesh_n"])
    eng.coupling_config(sb.DRAG_ERGUN_WENYU, sb.FORCE_DRAG | sb.FORCE_PGRAD, case["nub"], case["rhob"], case["g"], substeps * case["dt"])
    Uf, gamma, gradp = cases.uniform_fields(case)
    eng.put_cell_fields(Uf, gamma, gradp)
    eng.setup()
    eng.scatter_alpha_u(device_only=True)

    def step():
        if couple:   # the scatter uses FP64 atomics: with it in the loop the state is not bitwise reproducible
            eng.compute_fluid_force()
        eng.sedi_step(substeps)
        if couple:
            eng.scatter_alpha_u(device_only=True)
            eng.calc_tc(device_only=True)

    if not couple:
        eng.compute_fluid_force()

    for _ in range(warm):
        step()
    eng.reset_stats()
    eng.profile(True)
    eng.synchronize()
    eng.timer_start()
    for _ in range(steps):
        step()
    ms = eng.timer_stop_ms()
    ksteps, kms = eng.get_profile()
    a = eng.atoms()
    order = np.argsort(a["tag"])
    h = hashlib.sha256()
    for k in ("x", "v", "omega"):
        h.update(np.ascontiguousarray(a[k][order]).tobytes())
    n = len(a["tag"])
    pairs = eng.stat("gran_pairs")
    ls = eng.list_stats()
    out = {"bed": bed, "cfg": cfg, "n": n, "touching_pairs_per_particle": ls["touching_pairs_per_particle"], "ell_width": ls["ell_width"], "kstep_us": 1e3 * kms / max(1, ksteps), "ms_per_step": ms / steps, "rebuilds": eng.stat("nbuilds"), "launches_timed": ksteps,
           "pairs_per_particle": pairs / n, "GBps_alg": (188.0 * n + 56.0 * pairs) / (kms / max(1, ksteps) * 1e-3) / 1e9 if kms > 0 else 0.0,
           "state_sha": h.hexdigest()[:16]}
    print("SWEEP " + json.dumps(out), flush=True)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", action="store_true")
    ap.add_argument("--dims", default="100x100x100")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--substeps", type=int, default=100)
    ap.add_argument("--libs", default=None, help="comma list of .so paths (default: in-tree + build_variants/*.so)")
    ap.add_argument("--envs", default="default=", help="semicolon list name=ENV1=v1,ENV2=v2")
    ap.add_argument("--out", default=None)
    ap.add_argument("--nocouple", action="store_true", help="DEM sub-steps only (bitwise reproducible state hash)")
    ap.add_argument("--bed", default="settled", choices=["settled", "random", "lattice"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--size", type=float, default=1.0)
    args = ap.parse_args()
    dims = tuple(int(v) for v in args.dims.split("x"))
    if args.one:
        one(dims, args.steps, args.warm, args.substeps, couple=not args.nocouple, bed=args.bed, cfg=args.config, size=args.size)
        return 0
    libs = args.libs.split(",") if args.libs else None
    if libs is None:
        libs = [os.path.join(ROOT, "sedifoam_b200", "libsedi_b200.so")]
        vd = os.path.join(ROOT, "build_variants")
        if os.path.isdir(vd):
            libs += sorted(os.path.join(vd, f) for f in os.listdir(vd) if f.endswith(".so"))
    envs = []
    for item in args.envs.split(";"):
        name, _, rest = item.partition("=")
        kv = dict(p.split("=", 1) for p in rest.split(",") if "=" in p)
        envs.append((name, kv))
    rows = []
    for lib in libs:
        for name, kv in envs:
            env = dict(os.environ)
            env["SEDI_B200_LIB"] = lib
            env.update(kv)
            cmd = [sys.executable, os.path.abspath(__file__), "--one", "--dims", args.dims, "--steps", str(args.steps), "--warm", str(args.warm),
                   "--substeps", str(args.substeps), "--bed", args.bed, "--config", str(args.config), "--size", str(args.size)] + (["--nocouple"] if args.nocouple else [])
            try:
                r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
                line = [l for l in r.stdout.splitlines() if l.startswith("SWEEP ")]
                res = json.loads(line[-1][6:]) if line else {"error": (r.stderr or r.stdout)[-400:]}
            except subprocess.TimeoutExpired:
                res = {"error": "timeout"}
            res.update({"lib": os.path.basename(lib), "env": name})
            rows.append(res)
            print(json.dumps(res), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)
    return 0


if __name__ == "__main__":
    rc = main()
    sys.stdout.flush()
    os._exit(rc)
