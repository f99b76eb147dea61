"""ctypes binding of libsedi_b200.so (see include/sedi_b200.h for the authoritative declarations)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DRAG_ERGUN_WENYU, DRAG_SYAMLAL_OBRIEN = 0, 1
FORCE_DRAG, FORCE_PGRAD, FORCE_BUOY, FORCE_ADDEDMASS, FORCE_LIFT = 1, 2, 4, 8, 16
FORCE_HISTORY, FORCE_WALL_LUB, FORCE_INLET = 32, 64, 128

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

# every symbol include/sedi_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
EXPORTED_SYMBOLS = [
    "lammps_open", "lammps_close", "lammps_file", "lammps_command", "lammps_sync", "lammps_get_global_n",
    "lammps_get_initial_np", "lammps_get_initial_info", "lammps_get_local_n", "lammps_get_local_domain",
    "lammps_get_local_info", "lammps_put_local_info", "lammps_step", "lammps_set_timestep", "lammps_get_timestep",
    "lammps_create_particle", "lammps_delete_particle",
    "sedi_abi_version", "sedi_config_json", "sedi_device_count", "sedi_set_device", "sedi_set_box", "sedi_add_atoms", "sedi_set_omega",
    "sedi_get_state", "sedi_get_pairs", "sedi_get_wall_shear", "sedi_get_row_stats", "sedi_force_rebuild", "sedi_get_stat", "sedi_reset_stats",
    "sedi_synchronize", "sedi_stream", "sedi_last_step_ms", "sedi_timer_start", "sedi_timer_stop_ms", "sedi_profile",
    "sedi_get_profile", "sedi_mesh_box", "sedi_mesh_rectilinear", "sedi_mesh_ncells", "sedi_coupling_config",
    "sedi_put_cell_fields", "sedi_coupling_time_index", "sedi_coupling_inlet", "sedi_get_history_state", "sedi_locate", "sedi_compute_fluid_force", "sedi_scatter_alpha_u", "sedi_calc_tc",
    "sedi_enable_diag", "sedi_get_coupling_diag", "sedi_step", "sedi_enable_conservation_sums", "sedi_get_conservation_sums",
    "sedi_average_info", "sedi_get_timers", "sedi_write_lagrangian", "sedi_comm_init", "sedi_comm_unique_id", "sedi_comm_rank", "sedi_comm_stat",
    "sedi_decomp_grid", "sedi_decomp_owner", "sedi_decomp_links", "sedi_smooth_config", "sedi_smooth_uf", "sedi_smooth_field",
    "sedi_smooth_last_iters",
]


def library_path():
    return os.environ.get("SEDI_B200_LIB", os.path.join(_HERE, "libsedi_b200.so"))


def build_library(verbose=False):
    """Compile csrc/sedi_engine.cu for sm_100a into sedifoam_b200/libsedi_b200.so (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc", "sedi_engine.cu")
    out = library_path()
    deps = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    deps.append(os.path.join(_HERE, "..", "include", "sedi_b200.h"))
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", out, src]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return out


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def load_library():
    """Load the CUDA library.  A missing library is an error (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError("libsedi_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the particle hot path)")
    lib = C.CDLL(path)
    vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
    sig = {
        "lammps_open": (None, [i, vp, i, C.POINTER(vp)]),
        "lammps_close": (None, [vp]),
        "lammps_file": (None, [vp, C.c_char_p]),
        "lammps_command": (C.c_char_p, [vp, C.c_char_p]),
        "lammps_sync": (None, [vp]),
        "lammps_get_global_n": (i, [vp]),
        "lammps_get_initial_np": (None, [vp, vp]),
        "lammps_get_initial_info": (None, [vp] + [vp] * 7),
        "lammps_get_local_n": (i, [vp]),
        "lammps_get_local_domain": (None, [vp, vp]),
        "lammps_get_local_info": (None, [vp] + [vp] * 5),
        "lammps_put_local_info": (None, [vp, i, vp, vp, vp, vp]),
        "lammps_step": (None, [vp, i]),
        "lammps_set_timestep": (None, [vp, d]),
        "lammps_get_timestep": (d, [vp]),
        "lammps_create_particle": (None, [vp, i, vp, vp, d, d, i, vp]),
        "lammps_delete_particle": (None, [vp, vp, i]),
        "sedi_abi_version": (i, []),
        "sedi_device_count": (i, []),
        "sedi_config_json": (i, [vp, C.c_char_p, i]),
        "sedi_set_device": (None, [vp, i]),
        "sedi_set_box": (None, [vp, vp, vp, i]),
        "sedi_add_atoms": (None, [vp, i] + [vp] * 6),
        "sedi_set_omega": (None, [vp, i, vp, vp]),
        "sedi_get_state": (None, [vp] + [vp] * 10),
        "sedi_get_pairs": (ll, [vp, vp, vp, vp, vp, vp, ll]),
        "sedi_get_wall_shear": (None, [vp, i, vp]),
        "sedi_get_row_stats": (None, [vp, vp, vp]),
        "sedi_force_rebuild": (None, [vp]),
        "sedi_get_stat": (ll, [vp, i]),
        "sedi_reset_stats": (None, [vp]),
        "sedi_synchronize": (None, [vp]),
        "sedi_stream": (vp, [vp]),
        "sedi_last_step_ms": (d, [vp]),
        "sedi_timer_start": (None, [vp]),
        "sedi_timer_stop_ms": (d, [vp]),
        "sedi_profile": (None, [vp, i]),
        "sedi_get_profile": (ll, [vp, C.POINTER(C.c_double)]),
        "sedi_mesh_box": (None, [vp, vp, vp, vp]),
        "sedi_mesh_rectilinear": (None, [vp, vp, vp, vp, vp, vp]),
        "sedi_mesh_ncells": (i, [vp]),
        "sedi_coupling_config": (None, [vp, i, i, d, d, vp, d]),
        "sedi_put_cell_fields": (None, [vp] + [vp] * 5),
        "sedi_coupling_time_index": (None, [vp, i]),
        "sedi_coupling_inlet": (None, [vp, vp, vp, i, vp]),
        "sedi_get_history_state": (None, [vp, vp, vp]),
        "sedi_locate": (None, [vp]),
        "sedi_compute_fluid_force": (None, [vp]),
        "sedi_scatter_alpha_u": (None, [vp, vp, vp]),
        "sedi_calc_tc": (None, [vp, vp, vp]),
        "sedi_enable_diag": (None, [vp, i]),
        "sedi_smooth_config": (None, [vp, d, i, vp, i]),
        "sedi_smooth_uf": (None, [vp]),
        "sedi_smooth_field": (None, [vp, vp, i]),
        "sedi_smooth_last_iters": (i, [vp]),
        "sedi_get_coupling_diag": (None, [vp] + [vp] * 6),
        "sedi_step": (None, [vp, i]),
        "sedi_enable_conservation_sums": (None, [vp, i]),
        "sedi_get_conservation_sums": (None, [vp, vp, vp, vp, vp]),
        "sedi_average_info": (None, [vp, vp, vp, vp]),
        "sedi_get_timers": (None, [vp, vp, vp, vp]),
        "sedi_write_lagrangian": (None, [vp, C.c_char_p, C.c_char_p]),
        "sedi_comm_init": (i, [vp, i, i, vp, i, vp]),
        "sedi_comm_unique_id": (i, [vp, i]),
        "sedi_comm_rank": (i, [vp]),
        "sedi_comm_stat": (ll, [vp, i]),
        "sedi_decomp_grid": (None, [i, vp, vp]),
        "sedi_decomp_owner": (i, [vp, vp, vp, vp]),
        "sedi_decomp_links": (i, [i, vp, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


_STATS = {"nbuilds": 0, "pair_evals": 1, "steps": 2, "gran_entries": 3, "type_entries": 4, "ell_cap": 5, "launches": 6,
          "nlocal": 7, "gran_pairs": 8, "nghost": 9, "pair_evals_unique": 10}


class Lammps:
    """One engine instance == the `void *` handle of the reference's library.h (lammps_open ... lammps_close).

    Method names follow the C functions without the prefix; array arguments are caller-allocated numpy arrays
    exactly like the caller-allocated `double *` of the reference (softParticleCloud.C:139-147, 908-912).
    """

    def __init__(self, device=None):
        self.lib = load_library()
        h = C.c_void_p()
        self.lib.lammps_open(0, None, 0, C.byref(h))
        self.h = h
        if device is not None:
            self.lib.sedi_set_device(self.h, int(device))

    def close(self):
        if getattr(self, "h", None):
            self.lib.lammps_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- library.h -------------------------------------------------------------------------------------------
    def command(self, line):
        self.lib.lammps_command(self.h, line.encode())

    def commands(self, text):
        for ln in text.strip().splitlines():
            self.command(ln)

    def config(self):
        """what the script parser understood (dict): pair / fix / group / dump settings"""
        import json
        buf = C.create_string_buffer(1 << 16)
        n = self.lib.sedi_config_json(self.h, buf, len(buf))
        if n < 0:
            buf = C.create_string_buffer(-n + 16)
            n = self.lib.sedi_config_json(self.h, buf, len(buf))
        return json.loads(buf.value.decode())

    def file(self, path):
        self.lib.lammps_file(self.h, str(path).encode())

    def sync(self):
        self.lib.lammps_sync(self.h)

    def get_global_n(self):
        return self.lib.lammps_get_global_n(self.h)

    def get_initial_np(self, nprocs=1):
        np_ = np.zeros(nprocs, np.int32)
        self.lib.lammps_get_initial_np(self.h, _vp(np_))
        return np_

    def get_initial_info(self):
        n = self.get_local_n()
        x = np.zeros((n, 3)); v = np.zeros((n, 3)); d = np.zeros(n); rho = np.zeros(n)
        tag = np.zeros(n, np.int32); cpu = np.zeros(n, np.int32); typ = np.zeros(n, np.int32)
        self.lib.lammps_get_initial_info(self.h, _vp(x), _vp(v), _vp(d), _vp(rho), _vp(tag), _vp(cpu), _vp(typ))
        return dict(x=x, v=v, diam=d, rho=rho, tag=tag, lmpCpuId=cpu, type=typ)

    def get_local_n(self):
        return self.lib.lammps_get_local_n(self.h)

    def get_local_domain(self):
        dom = np.zeros(6)
        self.lib.lammps_get_local_domain(self.h, _vp(dom))
        return dom

    def get_local_info(self, x=None, v=None, foam=None, lmp=None, tag=None):
        n = self.get_local_n()
        x = np.zeros((n, 3)) if x is None else x
        v = np.zeros((n, 3)) if v is None else v
        foam = np.zeros(n, np.int32) if foam is None else foam
        lmp = np.zeros(n, np.int32) if lmp is None else lmp
        tag = np.zeros(n, np.int32) if tag is None else tag
        self.lib.lammps_get_local_info(self.h, _vp(x), _vp(v), _vp(foam), _vp(lmp), _vp(tag))
        return dict(x=x, v=v, foamCpuId=foam, lmpCpuId=lmp, tag=tag)

    def put_local_info(self, fdrag, tags, foam_cpu=None, DuDt=None):
        n = len(tags)
        fdrag = _f64(fdrag, (n, 3))
        tags = _i32(tags)
        foam_cpu = np.zeros(n, np.int32) if foam_cpu is None else _i32(foam_cpu)
        DuDt = None if DuDt is None else _f64(DuDt, (n, 3))   # ignored by the reference too (library.cpp:314-367)
        self.lib.lammps_put_local_info(self.h, n, _vp(fdrag), _vp(DuDt), _vp(foam_cpu), _vp(tags))

    def step(self, n):
        self.lib.lammps_step(self.h, int(n))

    def set_timestep(self, dt):
        self.lib.lammps_set_timestep(self.h, float(dt))

    def get_timestep(self):
        return self.lib.lammps_get_timestep(self.h)

    def create_particle(self, pos, tags, diameter, rho, typ, vel):
        pos = _f64(pos).reshape(-1, 3)
        t = _f64(tags)
        vel = _f64(vel)
        self.lib.lammps_create_particle(self.h, len(t), _vp(pos), _vp(t), float(diameter), float(rho), int(typ), _vp(vel))

    def delete_particle(self, tags):
        t = _i32(tags)
        self.lib.lammps_delete_particle(self.h, _vp(t), len(t))

    # ---- sedi_* ----------------------------------------------------------------------------------------------
    def set_box(self, lo, hi, ntypes=1):
        lo = _f64(lo); hi = _f64(hi)
        self.lib.sedi_set_box(self.h, _vp(lo), _vp(hi), ntypes)

    def add_atoms(self, tag, typ, diam, rho, x, v=None):
        n = len(tag)
        tag = _i32(tag); typ = _i32(typ); diam = _f64(diam); rho = _f64(rho); x = _f64(x, (n, 3))
        v = None if v is None else _f64(v, (n, 3))
        self.lib.sedi_add_atoms(self.h, n, _vp(tag), _vp(typ), _vp(diam), _vp(rho), _vp(x), _vp(v))

    def set_omega(self, tags, omega):
        tags = _i32(tags); omega = _f64(omega, (len(tags), 3))
        self.lib.sedi_set_omega(self.h, len(tags), _vp(tags), _vp(omega))

    def setup(self):
        self.step(0)

    def force_rebuild(self):
        self.lib.sedi_force_rebuild(self.h)

    def stat(self, name):
        return int(self.lib.sedi_get_stat(self.h, _STATS[name]))

    def reset_stats(self):
        self.lib.sedi_reset_stats(self.h)

    def synchronize(self):
        self.lib.sedi_synchronize(self.h)

    def last_step_ms(self):
        return float(self.lib.sedi_last_step_ms(self.h))

    def timer_start(self):
        self.lib.sedi_timer_start(self.h)

    def timer_stop_ms(self):
        return float(self.lib.sedi_timer_stop_ms(self.h))

    def profile(self, on=True):
        self.lib.sedi_profile(self.h, 1 if on else 0)

    def get_profile(self):
        ms = C.c_double(0.0)
        n = int(self.lib.sedi_get_profile(self.h, C.byref(ms)))
        return n, float(ms.value)

    def atoms(self):
        """dict of owned-atom arrays sorted by tag (identity across the boundary is the tag)."""
        n = self.get_local_n()
        out = {k: np.zeros((n, 3)) for k in ("x", "v", "omega", "f", "torque")}
        radius = np.zeros(n); rmass = np.zeros(n)
        tag = np.zeros(n, np.int32); typ = np.zeros(n, np.int32); mask = np.zeros(n, np.int32)
        self.lib.sedi_get_state(self.h, _vp(out["x"]), _vp(out["v"]), _vp(out["omega"]), _vp(out["f"]), _vp(out["torque"]),
                                _vp(radius), _vp(rmass), _vp(tag), _vp(typ), _vp(mask))
        o = np.argsort(tag, kind="stable")
        res = {k: a[o] for k, a in out.items()}
        res.update(tag=tag[o], radius=radius[o], rmass=rmass[o], type=typ[o], mask=mask[o], order=o)
        return res

    def pairs(self):
        """directed neighbour rows: tag_i, tag_j, meta(flags|image), touch, shear"""
        m = int(self.lib.sedi_get_pairs(self.h, None, None, None, None, None, 0))
        ti = np.zeros(m, np.int32); tj = np.zeros(m, np.int32); meta = np.zeros(m, np.uint32)
        touch = np.zeros(m, np.int32); shear = np.zeros((m, 3))
        if m:
            self.lib.sedi_get_pairs(self.h, _vp(ti), _vp(tj), _vp(meta), _vp(touch), _vp(shear), m)
        return dict(ti=ti, tj=tj, gran=(meta >> 30) & 1, type=(meta >> 31) & 1, img=(meta >> 25) & 31, touch=touch, shear=shear)

    def list_stats(self):
        """shape of the neighbour list of this rank: pairs and overlapping pairs per particle, row-length histogram,
        ELL width and fill (what a benchmark line says about its bed)"""
        n = self.get_local_n()
        nn = np.zeros(n, np.int32); nt = np.zeros(n, np.int32)
        if n:
            self.lib.sedi_get_row_stats(self.h, _vp(nn), _vp(nt))
        cap = self.stat("ell_cap")
        tot = float(nn.sum())
        return {"pairs_per_particle": self.stat("gran_pairs") / max(1, n), "directed_entries_per_row": tot / max(1, n),
                "touching_pairs_per_particle": 0.5 * float(nt.sum()) / max(1, n),
                "touching_fraction": float(nt.sum()) / max(1.0, float(self.stat("gran_entries"))),
                "row_length_hist": np.bincount(nn).tolist() if n else [], "ell_width": cap, "ell_fill": tot / max(1.0, float(cap) * n)}

    def wall_shear(self, wall):
        """wall history [n][3] in device row order together with the row tags"""
        n = self.get_local_n()
        out = np.zeros((n, 3))
        self.lib.sedi_get_wall_shear(self.h, wall, _vp(out))
        tag = np.zeros(n, np.int32)
        self.lib.sedi_get_state(self.h, None, None, None, None, None, None, None, _vp(tag), None, None)
        o = np.argsort(tag, kind="stable")
        return out[o]

    # ---- coupling ---------------------------------------------------------------------------------------------
    def mesh_box(self, lo, hi, ncell):
        lo = _f64(lo); hi = _f64(hi); nc = _i32(ncell)
        self.lib.sedi_mesh_box(self.h, _vp(lo), _vp(hi), _vp(nc))

    def mesh_rectilinear(self, xf, yf, zf, label=None):
        xf, yf, zf = _f64(xf), _f64(yf), _f64(zf)
        nc = np.array([len(xf) - 1, len(yf) - 1, len(zf) - 1], np.int32)
        lab = None if label is None else np.ascontiguousarray(label, np.int32)
        self.lib.sedi_mesh_rectilinear(self.h, _vp(nc), _vp(xf), _vp(yf), _vp(zf), _vp(lab))

    def mesh_ncells(self):
        return self.lib.sedi_mesh_ncells(self.h)

    def coupling_config(self, drag_model, force_flags, nub, rhob, g=(0, 0, 0), deltaT=1.0):
        g = _f64(g)
        self.lib.sedi_coupling_config(self.h, drag_model, force_flags, nub, rhob, _vp(g), deltaT)

    def coupling_time_index(self, time_index):
        self.lib.sedi_coupling_time_index(self.h, int(time_index))

    def coupling_inlet(self, inlet_force, inlet_box, region_option=1, eccentricity=(0, 0, 0)):
        box = np.zeros(9); box[:len(inlet_box)] = inlet_box
        f = _f64(inlet_force); e = _f64(eccentricity)
        self.lib.sedi_coupling_inlet(self.h, _vp(f), _vp(box), int(region_option), _vp(e))

    def history_state(self):
        """(sumDeltaFb[n][3], n0[n]) sorted by tag"""
        n = self.get_local_n()
        s = np.zeros((n, 3)); n0 = np.zeros(n)
        self.lib.sedi_get_history_state(self.h, _vp(s), _vp(n0))
        tag = np.zeros(n, np.int32)
        self.lib.sedi_get_state(self.h, None, None, None, None, None, None, None, _vp(tag), None, None)
        o = np.argsort(tag, kind="stable")
        return s[o], n0[o]

    def put_cell_fields(self, Uf=None, gamma=None, gradp=None, DDtU=None, curlU=None):
        a = [None if f is None else _f64(f) for f in (Uf, gamma, gradp, DDtU, curlU)]
        self.lib.sedi_put_cell_fields(self.h, *[_vp(f) for f in a])

    def locate(self):
        self.lib.sedi_locate(self.h)

    def compute_fluid_force(self):
        self.lib.sedi_compute_fluid_force(self.h)

    def scatter_alpha_u(self, gamma=None, Ue=None, device_only=False):
        if device_only:   # results stay in HBM (gamma feeds the next sedi_compute_fluid_force)
            self.lib.sedi_scatter_alpha_u(self.h, None, None)
            return None, None
        Cn = self.mesh_ncells()
        gamma = np.zeros(Cn) if gamma is None else gamma
        Ue = np.zeros((Cn, 3)) if Ue is None else Ue
        self.lib.sedi_scatter_alpha_u(self.h, _vp(gamma), _vp(Ue))
        return gamma, Ue

    def calc_tc(self, Asrc=None, Omega=None, device_only=False):
        if device_only:
            self.lib.sedi_calc_tc(self.h, None, None)
            return None, None
        Cn = self.mesh_ncells()
        Asrc = np.zeros((Cn, 3)) if Asrc is None else Asrc
        Omega = np.zeros(Cn) if Omega is None else Omega
        self.lib.sedi_calc_tc(self.h, _vp(Asrc), _vp(Omega))
        return Asrc, Omega

    def smooth_config(self, bandwidth, steps, Ddiag=None, flags=15):
        D = None if Ddiag is None else _f64(Ddiag)
        self.lib.sedi_smooth_config(self.h, float(bandwidth), int(steps), _vp(D), int(flags))

    def smooth_uf(self):
        self.lib.sedi_smooth_uf(self.h)

    def smooth_field(self, field):
        f = np.ascontiguousarray(field, np.float64).copy()
        ncomp = 1 if f.ndim == 1 else f.shape[1]
        self.lib.sedi_smooth_field(self.h, _vp(f), ncomp)
        return f

    def smooth_last_iters(self):
        return int(self.lib.sedi_smooth_last_iters(self.h))

    def enable_diag(self, on=True):
        self.lib.sedi_enable_diag(self.h, 1 if on else 0)

    def coupling_diag(self):
        """per-particle coupling diagnostics sorted by tag"""
        n = self.get_local_n()
        cell = np.zeros(n, np.int32); Uri = np.zeros((n, 3)); mag = np.zeros(n); al = np.zeros(n); Jd = np.zeros(n); F = np.zeros((n, 3))
        self.lib.sedi_get_coupling_diag(self.h, _vp(cell), _vp(Uri), _vp(mag), _vp(al), _vp(Jd), _vp(F))
        tag = np.zeros(n, np.int32)
        self.lib.sedi_get_state(self.h, None, None, None, None, None, None, None, _vp(tag), None, None)
        o = np.argsort(tag, kind="stable")
        return dict(cell=cell[o], Uri=Uri[o], magUri=mag[o], alpha=al[o], Jd=Jd[o], F=F[o], tag=tag[o])

    def sedi_step(self, n):
        self.lib.sedi_step(self.h, int(n))

    def write_lagrangian(self, directory, location=None):
        """OpenFOAM lagrangian field files of the cloud (softParticleIO.C:157-197) into `directory`"""
        self.lib.sedi_write_lagrangian(self.h, str(directory).encode(), None if location is None else location.encode())

    def enable_conservation_sums(self, on=True):
        self.lib.sedi_enable_conservation_sums(self.h, 1 if on else 0)

    def conservation_sums(self):
        """the reference's printed invariants: dict(F_before, F_after, U_before, U_after), vectors of 3"""
        a = [np.zeros(3) for _ in range(4)]
        self.lib.sedi_get_conservation_sums(self.h, *[_vp(v) for v in a])
        return dict(F_before=a[0], F_after=a[1], U_before=a[2], U_after=a[3])

    def average_info(self):
        vol = np.zeros(1); tv = np.zeros(3); av = np.zeros(3)
        self.lib.sedi_average_info(self.h, _vp(vol), _vp(tv), _vp(av))
        return dict(totalVolume=float(vol[0]), totalVel=tv, averageVel=av)

    def timers(self):
        d = np.zeros(2); m = np.zeros(1); c = np.zeros(6)
        self.lib.sedi_get_timers(self.h, _vp(d), _vp(m), _vp(c))
        return dict(diffusionTimeCount=d, particleMoveTime=float(m[0]), cpuTimeSplit=c)

    # ---- multi-GPU ---------------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        """ncclUniqueId bytes (call on rank 0, broadcast to the others)"""
        buf = (C.c_char * 256)()
        nb = load_library().sedi_comm_unique_id(C.cast(buf, C.c_void_p), 256)
        if nb <= 0:
            raise RuntimeError("NCCL is not available to libsedi_b200.so")
        return bytes(buf[:nb])

    def comm_init(self, rank, nranks, uid, procgrid=None):
        pg = None if procgrid is None else _i32(procgrid)
        b = C.create_string_buffer(uid, len(uid))
        self.lib.sedi_comm_init(self.h, int(rank), int(nranks), C.cast(b, C.c_void_p), len(uid), _vp(pg))

    def comm_stat(self, name):
        return int(self.lib.sedi_comm_stat(self.h, {"halo_calls": 0, "send_rows": 1, "ghost_rows": 2, "links": 3, "arrivals": 4, "p2p": 5}[name]))


def decomp_grid(nranks, boxlen):
    g = np.zeros(3, np.int32)
    bl = _f64(boxlen)
    load_library().sedi_decomp_grid(int(nranks), _vp(bl), _vp(g))
    return g


def decomp_owner(x, boxlo, boxhi, grid):
    lib = load_library()
    x = _f64(x).reshape(-1, 3); lo = _f64(boxlo); hi = _f64(boxhi); g = _i32(grid)
    return np.array([lib.sedi_decomp_owner(_vp(x[k]), _vp(lo), _vp(hi), _vp(g)) for k in range(len(x))], np.int32)


def decomp_links(rank, grid, periodic, prd):
    peers = np.zeros(26, np.int32); offs = np.zeros((26, 3), np.int32); sh = np.zeros((26, 3))
    g = _i32(grid); per = _i32(periodic); p = _f64(prd)
    n = load_library().sedi_decomp_links(int(rank), _vp(g), _vp(per), _vp(p), _vp(peers), _vp(offs), _vp(sh))
    return peers[:n], offs[:n], sh[:n]
