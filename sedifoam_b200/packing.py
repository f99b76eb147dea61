"""Seeded synthetic random sphere packings (BASELINE.json: "synthetic random packings of the named N").

A packing is built from a *periodic random tile*: n spheres dropped uniformly at random into a periodic cube sized for
the target solid fraction, then relaxed by repeated pairwise overlap removal (every overlapping pair is pushed apart
along its line of centres) until the largest overlap is below `tol` diameters.  Near the random-loose-packing density
(phi ~ 0.58) the result is a disordered, mechanically plausible bed: about 2.4 touching pairs per particle with
overlaps of order 1e-3 d -- the compression a 100-particle-deep bed of the benchmark's soft spheres (kn = 1e7) has under
its own weight -- and ragged neighbour rows (4.7 list pairs per particle at skin 0.25 d, 2..12 per row).  A dilute tile
(phi = 0.3) relaxes to zero overlap.  The tile is periodic, so any box is filled by repeating it and cutting with
planes; cut faces are flat and meet a granular wall placed one radius outside.

Tiles are cached as .npz next to this file (committed for the benchmark sizes; regenerated from the seed if absent).
Pure numpy / scipy, used by tests/, bench.py (both arms) and tools/ -- never by the CUDA library.
"""
import hashlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def relax_tile(n, phi, seed, dlo=1.0, dhi=1.0, tol=1.0e-3, margin=0.0, push=0.6, maxit=6000):
    """n spheres with diameters U[dlo, dhi] (dlo == dhi: monodisperse) in a periodic cube at solid fraction phi.
    Returns (x[n,3], d[n], L).  margin > 0 relaxes slightly inflated spheres so the true spheres end with a gap."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    d = rng.uniform(dlo, dhi, n) if dhi > dlo else np.full(n, float(dlo))
    r = 0.5 * d * (1.0 + margin)
    L = float((np.pi / 6.0 * np.sum(d ** 3) / phi) ** (1.0 / 3.0))
    x = rng.uniform(0.0, L, (n, 3))
    reach = float(d.max() * (1.0 + margin))
    dmean = float(d.mean())
    for _ in range(maxit):
        x %= L
        x[x >= L] = 0.0
        pairs = cKDTree(x, boxsize=L).query_pairs(reach, output_type="ndarray")
        i, j = pairs[:, 0], pairs[:, 1]
        dx = x[i] - x[j]
        dx -= L * np.round(dx / L)
        dist = np.sqrt((dx ** 2).sum(1))
        ov = r[i] + r[j] - dist
        m = ov > 0.0
        if not m.any() or ov[m].max() < tol * dmean:
            break
        i, j, dx, dist, ov = i[m], j[m], dx[m], np.maximum(dist[m], 1e-12), ov[m]
        s = (push * 0.5 * ov / dist)[:, None] * dx
        for k in range(3):
            x[:, k] += np.bincount(i, s[:, k], n) - np.bincount(j, s[:, k], n)
    else:
        raise RuntimeError("relax_tile: no convergence (phi too high for overlap removal)")
    x %= L
    x[x >= L] = 0.0
    return x, d, L


def tile(n=4096, phi=0.58, seed=20261017, dlo=1.0, dhi=1.0, tol=1.0e-3, margin=0.0):
    """cached relax_tile (unit mean-diameter scale: multiply by the physical diameter)"""
    key = "n%d_phi%.4f_s%d_d%.3f_%.3f_t%.0e_m%.3f" % (n, phi, seed, dlo, dhi, tol, margin)
    if key in _CACHE:
        return _CACHE[key]
    path = os.path.join(_HERE, "data", "tile_" + hashlib.sha1(key.encode()).hexdigest()[:12] + ".npz")
    if os.path.exists(path):
        z = np.load(path)
        if str(z["key"]) == key:
            _CACHE[key] = (z["x"], z["d"], float(z["L"]))
            return _CACHE[key]
    x, d, L = relax_tile(n, phi, seed, dlo, dhi, tol, margin)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + ".tmp%d.npz" % os.getpid()
        np.savez_compressed(tmp, x=x, d=d, L=L, key=key)
        os.replace(tmp, path)
    except OSError:
        pass
    _CACHE[key] = (x, d, L)
    return _CACHE[key]


def fill(tile_xdl, lo_t, hi_t, grid=None):
    """all tile images whose centre lies in [lo_t, hi_t) (box given in units of the tile edge, fractional allowed).
    Returns positions (same length unit as the tile), diameters, and a 0-based global id that does not depend on the
    block asked for: id = image index in the `grid` of tile images (default: the images needed for hi_t) * n + particle
    index -- bricks of one bed generated on different ranks agree on the tags."""
    x0, d0, L = tile_xdl
    lo_t = np.asarray(lo_t, np.float64); hi_t = np.asarray(hi_t, np.float64)
    i0 = np.floor(lo_t + 1e-12).astype(int); i1 = np.ceil(hi_t - 1e-12).astype(int)
    g = i1 if grid is None else np.ceil(np.asarray(grid, np.float64) - 1e-12).astype(int)
    xs, ds, ids = [], [], []
    n0 = len(d0)
    for ix in range(i0[0], i1[0]):
        for iy in range(i0[1], i1[1]):
            for iz in range(i0[2], i1[2]):
                x = x0 + L * np.array([ix, iy, iz], np.float64)
                m = np.all((x >= lo_t * L) & (x < hi_t * L), axis=1)
                xs.append(x[m]); ds.append(d0[m])
                img = (np.int64(ix) * g[1] + iy) * g[2] + iz
                ids.append(img * n0 + np.nonzero(m)[0])
    if not xs:
        return np.zeros((0, 3)), np.zeros(0), np.zeros(0, np.int64)
    return np.concatenate(xs), np.concatenate(ds), np.concatenate(ids)


def row_stats(x, d, skin, box=None):
    """list pairs and touching pairs per particle, row-length histogram of the directed list (non-periodic estimate
    unless box = periodic edge lengths is given)"""
    from scipy.spatial import cKDTree
    n = len(d)
    t = cKDTree(x, boxsize=box)
    pr = t.query_pairs(float(d.max()) + skin, output_type="ndarray")
    dx = x[pr[:, 0]] - x[pr[:, 1]]
    if box is not None:
        dx -= np.asarray(box) * np.round(dx / np.asarray(box))
    dist = np.sqrt((dx ** 2).sum(1))
    rs = 0.5 * (d[pr[:, 0]] + d[pr[:, 1]])
    lst = dist <= rs + skin
    rows = np.bincount(pr[lst, 0], minlength=n) + np.bincount(pr[lst, 1], minlength=n)
    return dict(list_pairs_per_particle=float(lst.sum()) / n, touching_pairs_per_particle=float((dist < rs).sum()) / n,
                row_hist=np.bincount(rows).tolist(), row_max=int(rows.max()), row_mean=float(rows.mean()))
