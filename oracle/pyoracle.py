"""oracle/pyoracle.py -- ctypes wrapper of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
kind="port"      -> oracle/liboracle.so (our restatement; travels to the GPU box)
kind="reference" -> oracle/_ref/libsedi_ref.so (the reference's own plug-in sources, stub-compiled)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(quiet=True):
    """(Re)build the checker with oracle/Makefile.  _ref is only built where /root/reference exists."""
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _load(kind):
    if kind in _LIBS:
        return _LIBS[kind]
    path = os.path.join(_HERE, "liboracle.so" if kind == "port" else os.path.join("_ref", "libsedi_ref.so"))
    if not os.path.exists(path):
        if kind == "port":
            build()
        else:
            raise FileNotFoundError(path)
    lib = C.CDLL(path)
    lib.ora_create.restype = C.c_void_p
    lib.ora_create.argtypes = [C.c_int]
    lib.ora_destroy.argtypes = [C.c_void_p]
    lib.ora_command.argtypes = [C.c_void_p, C.c_char_p]
    lib.ora_set_box.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
    lib.ora_add_atoms.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_void_p]
    lib.ora_set_omega.argtypes = [C.c_void_p, _dp]
    lib.ora_nlocal.argtypes = [C.c_void_p]
    lib.ora_nghost.argtypes = [C.c_void_p]
    lib.ora_get_atoms.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    lib.ora_get_radius_mass.argtypes = [C.c_void_p, _dp, _dp]
    lib.ora_put_fdrag.argtypes = [C.c_void_p, C.c_int, _dp, C.c_void_p, _ip]
    lib.ora_set_timestep.argtypes = [C.c_void_p, C.c_double]
    lib.ora_run.argtypes = [C.c_void_p, C.c_longlong]
    lib.ora_setup.argtypes = [C.c_void_p]
    lib.ora_reneighbor.argtypes = [C.c_void_p]
    lib.ora_stat.restype = C.c_longlong
    lib.ora_stat.argtypes = [C.c_void_p, C.c_int]
    lib.ora_get_pairs.restype = C.c_longlong
    lib.ora_get_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    lib.ora_get_wall_shear.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ora_foam_jd_ergun_wenyu.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    lib.ora_foam_jd_syamlal_obrien.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    lib.ora_foam_particle_force.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_double, C.c_double, _dp, C.c_double,
                                            _dp, _dp, _dp, _dp, _dp, _dp]
    lib.ora_foam_particle_force_extra.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double,
                                                  C.c_double, C.c_int, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp]
    lib.ora_foam_cell_owner_rect.argtypes = [C.c_int, _dp, _ip, _dp, _dp, _dp, C.c_void_p, _ip]
    lib.ora_foam_cell_owner.argtypes = [C.c_int, _dp, _dp, _dp, _ip, _ip]
    lib.ora_foam_particle_to_eulerian.argtypes = [C.c_int, _ip, _dp, _dp, C.c_int, _dp, _dp, _dp]
    lib.ora_foam_calc_tc.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, _dp, _dp]
    if kind == "reference" and hasattr(lib, "ora_ref_jd_ergun_wenyu"):
        lib.ora_ref_jd_ergun_wenyu.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
        lib.ora_ref_jd_syamlal_obrien.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    _LIBS[kind] = lib
    return lib


def have_reference():
    return os.path.exists(os.path.join(_HERE, "_ref", "libsedi_ref.so"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU DEM time loop driven by LAMMPS script lines -- the checker for the CUDA engine."""

    def __init__(self, kind="port"):
        self.kind = kind
        self.lib = _load(kind)
        self.h = self.lib.ora_create(1 if kind == "reference" else 0)
        if not self.h:
            raise RuntimeError("oracle backend %r unavailable" % kind)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ora_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def command(self, line):
        self.lib.ora_command(self.h, line.encode())

    def commands(self, text):
        for ln in text.strip().splitlines():
            self.command(ln)

    def set_box(self, lo, hi, ntypes=1):
        self.lib.ora_set_box(self.h, np.ascontiguousarray(lo, np.float64), np.ascontiguousarray(hi, np.float64), ntypes)

    def add_atoms(self, tag, typ, diam, rho, x, v=None):
        n = len(tag)
        x = np.ascontiguousarray(x, np.float64).reshape(n, 3)
        v = None if v is None else np.ascontiguousarray(v, np.float64).reshape(n, 3)
        self.lib.ora_add_atoms(self.h, n, np.ascontiguousarray(tag, np.int32), np.ascontiguousarray(typ, np.int32),
                               np.ascontiguousarray(diam, np.float64), np.ascontiguousarray(rho, np.float64), x, _ptr(v))

    def set_omega(self, omega):
        self.lib.ora_set_omega(self.h, np.ascontiguousarray(omega, np.float64))

    @property
    def nlocal(self):
        return self.lib.ora_nlocal(self.h)

    def setup(self):
        self.lib.ora_setup(self.h)

    def run(self, n):
        self.lib.ora_run(self.h, int(n))

    def reneighbor(self):
        self.lib.ora_reneighbor(self.h)

    def set_timestep(self, dt):
        self.lib.ora_set_timestep(self.h, float(dt))

    def put_fdrag(self, fdrag, tags, foam_cpu=None):
        n = len(tags)
        fc = None if foam_cpu is None else np.ascontiguousarray(foam_cpu, np.int32)
        self.lib.ora_put_fdrag(self.h, n, np.ascontiguousarray(fdrag, np.float64).reshape(n, 3), _ptr(fc),
                               np.ascontiguousarray(tags, np.int32))

    def atoms(self):
        """dict of owned-atom arrays, sorted by tag (identity across the boundary is the tag)."""
        n = self.nlocal
        out = {k: np.zeros((n, 3)) for k in ("x", "v", "omega", "f", "torque")}
        tag = np.zeros(n, np.int32)
        self.lib.ora_get_atoms(self.h, _ptr(out["x"]), _ptr(out["v"]), _ptr(out["omega"]), _ptr(out["f"]),
                               _ptr(out["torque"]), _ptr(tag))
        o = np.argsort(tag, kind="stable")
        res = {k: a[o] for k, a in out.items()}
        res["tag"] = tag[o]
        return res

    def stat(self, name):
        return int(self.lib.ora_stat(self.h, {"nbuilds": 0, "pair_evals": 1, "steps": 2, "gran_pairs": 3,
                                              "half_pairs": 4, "full_pairs": 5}[name]))

    def pairs(self, which="gran", history=False):
        w = {"gran": 0, "half": 1, "full": 2}[which]
        m = int(self.lib.ora_get_pairs(self.h, w, None, None, None, None, 0, None))
        ti = np.zeros(m, np.int32); tj = np.zeros(m, np.int32)
        touch = np.zeros(m, np.int32) if history else None
        shear = np.zeros((m, 3)) if history else None
        ghost = np.zeros(m, np.int32)
        self.lib.ora_get_pairs(self.h, w, _ptr(ti), _ptr(tj), _ptr(touch), _ptr(shear), m, _ptr(ghost))
        self.last_ghost = ghost
        return (ti, tj, touch, shear) if history else (ti, tj)

    def wall_shear(self, wall):
        out = np.zeros((self.nlocal, 3))
        self.lib.ora_get_wall_shear(self.h, wall, out)
        return out


# ---- OpenFOAM-side coupling restatement -------------------------------------------------------------
FORCE_DRAG, FORCE_PGRAD, FORCE_BUOY, FORCE_ADDEDMASS, FORCE_LIFT = 1, 2, 4, 8, 16
DRAG_ERGUN_WENYU, DRAG_SYAMLAL_OBRIEN = 0, 1


def jd(model, Ur, alpha, pd, nuf, rhof, kind="port"):
    """Jd closure.  kind="port": the restatement (ora_foam_jd_*); kind="reference": the reference's own
    lammpsFoam/dragModels/{ErgunWenYu,SyamlalOBrien}.C compiled by path against oracle/stubs_foam/ (oracle/_ref)."""
    lib = _load(kind)
    Ur = np.ascontiguousarray(Ur, np.float64)
    out = np.zeros_like(Ur)
    if kind == "reference":
        fn = lib.ora_ref_jd_ergun_wenyu if model == DRAG_ERGUN_WENYU else lib.ora_ref_jd_syamlal_obrien
    else:
        fn = lib.ora_foam_jd_ergun_wenyu if model == DRAG_ERGUN_WENYU else lib.ora_foam_jd_syamlal_obrien
    fn(len(Ur), Ur, np.ascontiguousarray(alpha, np.float64), np.ascontiguousarray(pd, np.float64), nuf, rhof, out)
    return out


def particle_force(cell, d, U, UOld, Uf, gamma, gradp, DDtU, curlU, model, flags, nub, rhob, g, deltaT, kind="port"):
    lib = _load(kind)
    n = len(cell)
    c = lambda a: np.ascontiguousarray(a, np.float64)
    Uri = np.zeros((n, 3)); mag = np.zeros(n); al = np.zeros(n); Jd = np.zeros(n); F = np.zeros((n, 3)); DuDt = np.zeros((n, 3))
    DDtU = None if DDtU is None else c(DDtU)
    curlU = None if curlU is None else c(curlU)
    lib.ora_foam_particle_force(n, np.ascontiguousarray(cell, np.int32), c(d), c(U), c(UOld), c(Uf), c(gamma), c(gradp),
                                _ptr(DDtU), _ptr(curlU), model, flags, nub, rhob, c(g), deltaT, Uri, mag, al, Jd, F, DuDt)
    return dict(Uri=Uri, magUri=mag, alpha=al, Jd=Jd, F=F, DuDt=DuDt)


def particle_force_extra(cell, x, d, mass, U, UOld, Uf, UfOld, flags, nub, rhob, deltaT, time_index, sumDeltaFb, n0, F,
                         inlet_force=(0, 0, 0), inlet_box=(0,) * 9, region_option=0, ecc=(0, 0, 0), kind="port"):
    """history (32) / wall-lubrication (64) / inlet (128) branches of updateDragOnParticles applied to F in place;
    sumDeltaFb and n0 are the per-particle history state, updated in place"""
    lib = _load(kind)
    c = lambda a: np.ascontiguousarray(a, np.float64)
    assert F.flags.c_contiguous and sumDeltaFb.flags.c_contiguous and n0.flags.c_contiguous
    lib.ora_foam_particle_force_extra(len(cell), np.ascontiguousarray(cell, np.int32), c(x), c(d), c(mass), c(U), c(UOld), c(Uf), c(UfOld),
                                      flags, nub, rhob, deltaT, time_index, sumDeltaFb, n0, c(inlet_force), c(inlet_box), region_option,
                                      c(ecc), F)
    return F


def cell_owner(x, lo, hi, ncell, kind="port"):
    lib = _load(kind)
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros(len(x), np.int32)
    lib.ora_foam_cell_owner(len(x), x, np.ascontiguousarray(lo, np.float64), np.ascontiguousarray(hi, np.float64),
                            np.ascontiguousarray(ncell, np.int32), out)
    return out


def cell_owner_rect(x, xf, yf, zf, label=None, kind="port"):
    lib = _load(kind)
    x = np.ascontiguousarray(x, np.float64)
    c = lambda a: np.ascontiguousarray(a, np.float64)
    nc = np.array([len(xf) - 1, len(yf) - 1, len(zf) - 1], np.int32)
    lab = None if label is None else np.ascontiguousarray(label, np.int32)
    out = np.zeros(len(x), np.int32)
    lib.ora_foam_cell_owner_rect(len(x), x, nc, c(xf), c(yf), c(zf), _ptr(lab), out)
    return out


def particle_to_eulerian(cell, d, U, cellV, kind="port"):
    lib = _load(kind)
    Cn = len(cellV)
    gamma = np.zeros(Cn); Ue = np.zeros((Cn, 3))
    lib.ora_foam_particle_to_eulerian(len(cell), np.ascontiguousarray(cell, np.int32), np.ascontiguousarray(d, np.float64),
                                      np.ascontiguousarray(U, np.float64), Cn, np.ascontiguousarray(cellV, np.float64), gamma, Ue)
    return gamma, Ue


def calc_tc(cell, d, U, Uf, gamma, cellV, model, nub, rhob, kind="port"):
    lib = _load(kind)
    Cn = len(cellV)
    Asrc = np.zeros((Cn, 3)); Omega = np.zeros(Cn)
    c = lambda a: np.ascontiguousarray(a, np.float64)
    lib.ora_foam_calc_tc(len(cell), np.ascontiguousarray(cell, np.int32), c(d), c(U), c(Uf), c(gamma), Cn, c(cellV), model,
                         nub, rhob, Asrc, Omega)
    return Asrc, Omega


def smooth_field(phi, ncell, dx, bandwidth, steps, Ddiag=(1.0, 1.0, 1.0)):
    """enhancedCloud::smoothField restated (lammpsFoam/enhancedCloud.C:790-907, :564-568):
    `steps` implicit-Euler steps of d_tau = (b^2/4)/steps of d(phi)/d(tau) = div(D grad phi), zeroGradient walls, on the
    uniform box mesh; each step one sparse direct solve (scipy) -- the converged answer of OpenFOAM's PCG (tol 1e-10).
    PARITY UNPINNED at bit level (no OpenFOAM here); pinned by conservation of sum(phi V) and the Gaussian limit."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    nx, ny, nz = [int(v) for v in ncell]
    C = nx * ny * nz
    dtau = (bandwidth * bandwidth / 4.0) / (steps + 1.0e-150)
    w = [dtau * Ddiag[k] / (dx[k] * dx[k]) for k in range(3)]

    def lap1(n):
        e = np.ones(n)
        L = sp.diags([e[:-1], -2 * e, e[:-1]], [-1, 0, 1], format="lil")
        L[0, 0] = -1.0; L[n - 1, n - 1] = -1.0      # zeroGradient: the wall face carries no flux
        if n == 1:
            L[0, 0] = 0.0
        return sp.csr_matrix(L)
    Ix, Iy, Iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    # cell = i + nx (j + ny k): x fastest
    Lx = sp.kron(Iz, sp.kron(Iy, lap1(nx))); Ly = sp.kron(Iz, sp.kron(lap1(ny), Ix)); Lz = sp.kron(lap1(nz), sp.kron(Iy, Ix))
    A = sp.identity(C) - (w[0] * Lx + w[1] * Ly + w[2] * Lz)
    lu = spl.splu(sp.csc_matrix(A))
    out = np.array(phi, np.float64, copy=True)
    flat = out.reshape(C, -1)
    for _ in range(int(steps)):
        for k in range(flat.shape[1]):
            flat[:, k] = lu.solve(flat[:, k])
    return out


def smooth_field_rect(phi, xf, yf, zf, label, bandwidth, steps, Ddiag=(1.0, 1.0, 1.0)):
    """smoothField on a rectilinear (graded / stacked-block) mesh: the volume-integrated finite-volume system of
    fvm::ddt - fvm::laplacian on an orthogonal mesh,  V_c (phi_c - phi_c^n)/d_tau = sum_f D_nn A_f (phi_nb - phi_c)/delta_f,
    zeroGradient walls, solved directly.  phi is indexed by the host cell label.  PARITY UNPINNED (no OpenFOAM here)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    h = [np.diff(np.asarray(f, np.float64)) for f in (xf, yf, zf)]
    nx, ny, nz = len(h[0]), len(h[1]), len(h[2])
    C = nx * ny * nz
    lab = np.arange(C) if label is None else np.asarray(label)
    dtau = (bandwidth * bandwidth / 4.0) / (steps + 1.0e-150)
    idx = np.arange(C).reshape(nz, ny, nx)              # tensor index t = i + nx (j + ny k)
    HX, HY, HZ = np.meshgrid(h[2], h[1], h[0], indexing="ij")[::-1]   # each [nz][ny][nx]: hx, hy, hz of the cell
    V = (HX * HY * HZ).ravel()
    rows, cols, vals = [], [], []
    diag = V.copy()
    for axis, (H, D) in enumerate(((HX, Ddiag[0]), (HY, Ddiag[1]), (HZ, Ddiag[2]))):
        ax = 2 - axis                                     # array axis of this direction
        sl_lo = [slice(None)] * 3; sl_hi = [slice(None)] * 3
        sl_lo[ax] = slice(0, -1); sl_hi[ax] = slice(1, None)
        a = idx[tuple(sl_lo)].ravel(); b = idx[tuple(sl_hi)].ravel()
        area = (V / H.ravel())                            # face area = V / h along the normal
        coef = dtau * D * area[a] / (0.5 * (H.ravel()[a] + H.ravel()[b]))
        rows += [a, b]; cols += [b, a]; vals += [-coef, -coef]
        np.add.at(diag, a, coef); np.add.at(diag, b, coef)
    rows.append(np.arange(C)); cols.append(np.arange(C)); vals.append(diag)
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(C, C))
    lu = spl.splu(A)
    out = np.array(phi, np.float64, copy=True)
    flat = out.reshape(C, -1)
    for _ in range(int(steps)):
        for k in range(flat.shape[1]):
            t = flat[lab, k]                              # tensor order
            flat[lab, k] = lu.solve(V * t)
    return out
