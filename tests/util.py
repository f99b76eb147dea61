"""helpers shared by the parity tests"""
import numpy as np

from sedifoam_b200 import cases


def make_oracle(pyoracle, case, kind="port"):
    o = pyoracle.Oracle(kind)
    cases.apply(case, o)
    return o


def make_engine(case):
    from sedifoam_b200 import Lammps
    e = Lammps()
    cases.apply(case, e)
    return e


def directed_from_oracle(o, which="gran", history=False):
    """oracle half list -> directed rows (ti, tj, is_image), the form the CUDA engine stores.
    owned-owned pairs appear once in the oracle (expanded to both directions here); owned-ghost (periodic image)
    pairs are stored by both owners already."""
    res = o.pairs(which, history=history)
    ti, tj = res[0], res[1]
    g = o.last_ghost.astype(bool)
    if which == "full":
        rows = np.stack([ti, tj, g.astype(np.int32)], axis=1)
        return rows, None, None
    a = np.concatenate([ti, tj[~g]])
    b = np.concatenate([tj, ti[~g]])
    im = np.concatenate([g.astype(np.int32), np.zeros((~g).sum(), np.int32)])
    rows = np.stack([a, b, im], axis=1)
    if history:
        touch, shear = res[2], res[3]
        t = np.concatenate([touch, touch[~g]])
        s = np.concatenate([shear, -shear[~g]])
        return rows, t, s
    return rows, None, None


def sort_rows(rows, *extra):
    o = np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))
    return (rows[o],) + tuple(None if e is None else e[o] for e in extra)


def engine_rows(e, flag="gran"):
    p = e.pairs()
    sel = p[flag].astype(bool)
    rows = np.stack([p["ti"][sel], p["tj"][sel], (p["img"][sel] != 13).astype(np.int32)], axis=1)
    return rows, p["touch"][sel], p["shear"][sel]


def rel_err(a, b, scale=None):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    s = np.maximum(np.abs(b).max() if scale is None else scale, 1e-300)
    return np.abs(a - b).max() / s if a.size else 0.0


def small_bed(dims=(8, 10, 8), **kw):
    return cases.fluidized_bed(dims=dims, **kw)
