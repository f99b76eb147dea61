"""helpers shared by the parity tests"""
import numpy as np

from sedifoam_b200 import cases


def default_kind(pyoracle):
    """the checker of the parity tests: the reference's OWN plug-in sources (oracle/_ref, compiled by path from
    /root/reference; the prebuilt library travels to the GPU box) wherever that library exists, else the port that
    tests/test_oracle_pinning.py pins to it bit for bit"""
    return "reference" if pyoracle.have_reference() else "port"


def make_oracle(pyoracle, case, kind=None):
    o = pyoracle.Oracle(kind or default_kind(pyoracle))
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


# ---- scenarios shared by the pinning test (port == reference objects on the CPU) and the GPU parity tests ----------------
def hooke_history_bed(dims=(10, 12, 10)):
    """stock gran/hooke/history pair (EXTERNAL, restated) + the in-tree Hooke-history wall (fix_wall_granFix.cpp:441-554)"""
    case = cases.fluidized_bed(dims=dims)
    case["script"] = case["script"].replace("gran/hertzFix/history 10000000 NULL 0.9 NULL 0.4 1",
                                            "gran/hooke/history 2000.0 NULL 50.0 NULL 0.4 1")
    case["script"] = case["script"].replace("wall/granFix 10000000 NULL 0.9 NULL 0.4 1", "wall/gran 2000.0 NULL 50.0 NULL 0.4 1")
    assert "hooke" in case["script"] and "hertz" not in case["script"]
    return case


def hooke_bed(dims=(8, 10, 8)):
    """history-free gran/hooke pair + the in-tree Hooke wall (fix_wall_granFix.cpp:356-437)"""
    case = hooke_history_bed(dims)
    case["script"] = case["script"].replace("gran/hooke/history", "gran/hooke")
    return case


def wiggle_wall_bed(dims=(8, 10, 8), style="hertz"):
    """the floor oscillates (wall/granFix ... wiggle y A T, fix_wall_granFix.cpp:254-262): wall position and velocity
    change every sub-step"""
    case = cases.fluidized_bed(dims=dims) if style == "hertz" else hooke_history_bed(dims)
    d = case["diam"][0]
    out = []
    for ln in case["script"].splitlines():
        if ln.startswith("fix yw "):
            ln += " wiggle y %.9g %.9g" % (0.02 * d, 4.0e-4)
        out.append(ln)
    case["script"] = "\n".join(out) + "\n"
    return case


def shear_wall_bed(dims=(8, 10, 8)):
    """the z walls translate along x (wall/granFix ... shear x v, :263)"""
    case = cases.fluidized_bed(dims=dims)
    out = []
    for ln in case["script"].splitlines():
        if ln.startswith("fix zw "):
            ln += " shear x 0.05"
        out.append(ln)
    case["script"] = "\n".join(out) + "\n"
    return case


def zcylinder_bed(dims=(8, 8, 8), shear=None):
    """bed inside a vertical cylinder wall (wall/granFix ... zcylinder R [shear x|z v], :309-322): the lattice is centred
    on the axis, the cylinder touches the outermost particles; `shear x` is the rotating cylinder of the reference"""
    case = cases.fluidized_bed(dims=dims)
    d = case["diam"][0]
    lo, hi = case["box_lo"].copy(), case["box_hi"].copy()
    c = 0.5 * (case["x"].min(axis=0) + case["x"].max(axis=0))
    shift = np.array([c[0], c[1], 0.0])
    case["x"] = case["x"] - shift
    rad = np.sqrt(case["x"][:, 0] ** 2 + case["x"][:, 1] ** 2)
    far = rad > rad.max() - 0.75 * d          # drop the lattice corners: the cylinder then touches a ring of particles
    keep = ~far
    for k in ("tag", "type", "diam", "rho"):
        case[k] = case[k][keep]
    case["x"] = np.ascontiguousarray(case["x"][keep]); case["v"] = np.ascontiguousarray(case["v"][keep])
    case["tag"] = np.arange(1, keep.sum() + 1, dtype=np.int32)
    rad = rad[keep]
    R = rad.max() + 0.5 * d - 2.0e-3 * d
    case["box_lo"] = np.array([-R - d, -R - d, lo[2]]); case["box_hi"] = np.array([R + d, R + d, hi[2]])
    case["mesh_lo"], case["mesh_hi"] = case["box_lo"].copy(), case["box_hi"].copy()
    out = []
    for ln in case["script"].splitlines():
        if ln.startswith("fix xw ") or ln.startswith("fix yw "):
            continue
        out.append(ln)
    kn = "10000000 NULL 0.9 NULL 0.4 1"
    out.append("fix cw all wall/granFix %s zcylinder %.17g%s" % (kn, R, "" if shear is None else " shear %s 0.05" % shear))
    case["script"] = "\n".join(out) + "\n"
    case["script"] = case["script"].replace("gravity 9.8 vector 0 -1 0", "gravity 9.8 vector 0 0 -1")
    return case


def settled_random_bed(columns=(2, 2)):
    """ragged rows with a force-carrying contact network: the small settled random column (sedifoam_b200/data/
    column_256x4.npz, tools/make_settled_column.py --oracle) repeated in x and z"""
    return cases.settled_bed(columns=columns, column="column_256x4.npz")
