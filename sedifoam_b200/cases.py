"""Synthetic particle cases (seeded) shared by tests/ and bench.py -- the configurations of BASELINE.json / SURVEY.md 8(d).

A case is a plain dict: box, atoms (tag, type, diam, rho, x, v), a LAMMPS script (the plug-in API of the reference:
`pair_style gran/hertzFix/history`, `fix fdrag`, `fix cohesive`, `fix wall/granFix`, `pair_style lubricate/poly` ...),
and the fluid-side description (uniform blockMesh box, cell fields, cloudProperties switches).  `apply(case, sim)`
feeds it to either the CUDA engine (sedifoam_b200.Lammps) or the CPU oracle (oracle.pyoracle.Oracle): both expose
set_box / add_atoms / commands.  SI numbers pass through LAMMPS `lj` units as in every shipped in.lammps.
"""
import numpy as np

from . import packing

SEED = 20261017
TILE_N = 5000      # particles per periodic random tile: 5 x 8 x 5 tiles = 1e6, 2 x 5 x 2 = 1e5, 5 x 10 x 5 = 1.25e6


def lattice(dims, spacing, origin, jitter, rng, block=None):
    """simple-cubic lattice sites; block = (i0, i1, j0, j1, k0, k1) restricts to a sub-lattice of the global one
    (multi-GPU: every rank generates only its own brick).  Returns positions and global 1-based lattice tags."""
    nx, ny, nz = dims
    i0, i1, j0, j1, k0, k1 = (0, nx, 0, ny, 0, nz) if block is None else block
    ix, iy, iz = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), np.arange(k0, k1), indexing="ij")
    tags = (ix.ravel().astype(np.int64) * ny + iy.ravel()) * nz + iz.ravel() + 1
    x = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    x = np.asarray(origin, np.float64) + (x + 0.5) * np.asarray(spacing, np.float64)
    if jitter:
        x += rng.uniform(-jitter, jitter, size=x.shape)
    return x, tags.astype(np.int32)


def _base(x, d, rho, box_lo, box_hi, periodic, script, mesh_cell, typ=None, v=None, ntypes=1, extra=None, tags=None):
    n = len(x)
    d = np.full(n, d, np.float64) if np.isscalar(d) else np.asarray(d, np.float64)
    case = dict(
        box_lo=np.asarray(box_lo, np.float64), box_hi=np.asarray(box_hi, np.float64), periodic=periodic, ntypes=ntypes,
        tag=np.arange(1, n + 1, dtype=np.int32) if tags is None else tags, type=np.ones(n, np.int32) if typ is None else np.asarray(typ, np.int32),
        diam=d, rho=np.full(n, rho, np.float64), x=np.ascontiguousarray(x), v=np.zeros((n, 3)) if v is None else v,
        script=script,
    )
    lo, hi = case["box_lo"], case["box_hi"]
    nc = np.maximum(1, np.round((hi - lo) / mesh_cell)).astype(np.int32)
    case["mesh_lo"], case["mesh_hi"], case["mesh_n"] = lo.copy(), hi.copy(), nc
    case["nub"], case["rhob"] = 1.0e-6, 1000.0
    if extra:
        case.update(extra)
    return case


def apply(case, sim):
    """Feed a case to an engine/oracle object exposing set_box, add_atoms, commands."""
    sim.command("atom_style sphere")
    sim.command("boundary " + " ".join(case["periodic"]))
    sim.command("newton off")
    sim.command("communicate single vel yes")
    sim.set_box(case["box_lo"], case["box_hi"], case["ntypes"])
    sim.add_atoms(case["tag"], case["type"], case["diam"], case["rho"], case["x"], case["v"])
    sim.commands(case["script"])
    if case.get("omega") is not None:   # initial spins (settled beds), before the first setup on both sides
        if type(sim).__name__ == "Oracle":
            sim.set_omega(case["omega"])
        else:
            sim.set_omega(case["tag"], case["omega"])


def fluidized_bed(dims=(100, 100, 100), d=5.0e-4, rho=2650.0, overlap=2.0e-3, skin_frac=0.25, dt=2.0e-6, kn=1.0e7, e=0.9,
                  mu=0.4, head=0.5, seed=SEED, jitter_frac=1.0e-3, vjit=1.0e-3, block=None):
    """configs[2]: dense bed -- jittered simple-cubic lattice with every particle in (slightly pre-compressed,
    overlap*d) contact with its 6 lattice neighbours and the walls, solid fraction pi/6/(1-overlap)^3 = 0.527 -- in a
    box with granular walls on all sides and free head-room on top; Hertz-Mindlin pair + wall/granFix + gravity +
    fdrag.  The bed relaxes/expands during the run, which exercises history carry-over across neighbour rebuilds."""
    rng = np.random.default_rng(seed)
    a = d * (1.0 - overlap)
    ext = np.array(dims, np.float64) * a
    lo = np.zeros(3)
    hi = ext.copy()
    hi[1] = ext[1] * (1.0 + head)
    x, tags = lattice(dims, (a, a, a), lo, jitter_frac * d, rng, block)
    v = rng.uniform(-vjit, vjit, size=x.shape)
    script = f"""
neighbor {skin_frac * d:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
fix xw all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 xplane {lo[0]:.17g} {hi[0]:.17g}
fix yw all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane {lo[1]:.17g} {hi[1]:.17g}
fix zw all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 zplane {lo[2]:.17g} {hi[2]:.17g}
"""
    return _base(x, d, rho, lo, hi, ("f", "f", "f"), script, 4.0 * d, v=v, tags=tags,
                 extra=dict(name="fluidized_bed", Uf=(0.0, 0.02, 0.0), g=(0.0, -9.8, 0.0), dt=dt, substeps=100))


def sediment_column(dims=(36, 103, 27), d=5.0e-4, rho=2650.0, phi=0.30, skin_frac=0.25, dt=2.0e-6, kn=1.0e7, e=0.9, mu=0.4,
                    seed=SEED, jitter_frac=0.05):
    """configs[1]: monodisperse spheres sedimenting in a column periodic in x and z with a granular floor at y=0."""
    rng = np.random.default_rng(seed)
    a = d * (np.pi / 6.0 / phi) ** (1.0 / 3.0)
    ext = np.array(dims, np.float64) * a
    lo = np.zeros(3)
    hi = ext.copy()
    x, _ = lattice(dims, (a, a, a), lo, jitter_frac * d, rng)
    script = f"""
neighbor {skin_frac * d:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
fix yw all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane 0.0 NULL
"""
    return _base(x, d, rho, lo, hi, ("p", "f", "p"), script, 4.0 * d,
                 extra=dict(name="sediment_column", Uf=(0.0, 0.0, 0.0), g=(0.0, -9.8, 0.0), dt=dt, substeps=100))


def cohesive_shear_bed(dims=(40, 20, 40), d=5.0e-5, rho=2650.0, phi=0.55, dt=2.0e-8, kn=1.0e7, e=0.9, mu=0.4, opt=1,
                       vshear=0.01, seed=SEED, jitter_frac=1.0e-3):
    """configs[3]: cohesive silt bed, periodic in x/z, wall/granFix floor and a sheared wall/granFix lid, fix cohesive."""
    rng = np.random.default_rng(seed)
    a = d * (np.pi / 6.0 / phi) ** (1.0 / 3.0)
    ext = np.array(dims, np.float64) * a
    lo = np.zeros(3)
    hi = ext.copy()
    x, _ = lattice(dims, (a, a, a), lo, jitter_frac * d, rng)
    smax = 1.0e-6
    skin = max(0.25 * d, 2.0 * smax)
    script = f"""
neighbor {skin:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
fix 4 all cohesive 1e-20 1e-7 4e-10 {smax:.9g} {opt}
fix yb all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane 0.0 NULL
fix yt all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane NULL {hi[1]:.17g} shear x {vshear:.9g}
"""
    return _base(x, d, rho, lo, hi, ("p", "f", "p"), script, 4.0 * d,
                 extra=dict(name="cohesive_shear_bed", Uf=(0.05, 0.0, 0.0), g=(0.0, -9.8, 0.0), dt=dt, substeps=100))


def poly_lubricated(dims=(20, 20, 20), dmin=3.0e-4, dmax=7.0e-4, rho=2650.0, phi=0.18, dt=2.0e-6, kn=1.0e7, e=0.9, mu=0.4,
                    visc=1.0e-3, seed=SEED, vjit=0.01):
    """configs[4]: polydisperse periodic packing with hybrid/overlay gran/hertzFix/history + lubricate/poly (full list).
    cut_inner is 1.001*dmax: the reference switches lubrication off inside cut_inner (pair_lubricate_poly.cpp:294-297)
    and takes log(1/h) of the gap outside it (:308), so any overlapping pair must lie inside cut_inner (ri+rj <= dmax)
    or the reference itself produces NaN."""
    rng = np.random.default_rng(seed)
    n = int(np.prod(dims))
    diam = rng.uniform(dmin, dmax, size=n)
    vol = np.pi / 6.0 * np.sum(diam ** 3)
    a = (vol / phi / n) ** (1.0 / 3.0)
    ext = np.array(dims, np.float64) * a
    lo = np.zeros(3)
    hi = ext.copy()
    x, _ = lattice(dims, (a, a, a), lo, 0.02 * dmin, rng)
    v = rng.uniform(-vjit, vjit, size=x.shape)
    skin = 0.1 * dmin
    script = f"""
neighbor {skin:.9g} bin
neigh_modify delay 0
pair_style hybrid/overlay gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 lubricate/poly {visc:.9g} 1 1 {1.001 * dmax:.9g} {1.5 * dmax:.9g}
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 3 all fdrag
"""
    return _base(x, diam, rho, lo, hi, ("p", "p", "p"), script, 4.0 * dmax, v=v,
                 extra=dict(name="poly_lubricated", Uf=(0.0, 0.0, 0.0), g=(0.0, 0.0, 0.0), dt=dt, substeps=100))


def single_sphere():
    """configs[0]: the shipped install check cases/auto-testing/test-cases/xiaocase3 (1 sphere, d = 83 um, rho = 2000,
    uniform 0.05 m/s water flow, g = 0, stock gran/hooke/history + wall/gran, SyamlalOBrien drag, 100 DEM steps of
    2e-7 s per fluid step of 2e-5 s; in.lammps:3-32, IC_uniform.in, constant/transportProperties)."""
    x = np.array([[0.002, 0.0019, 0.00025]])
    script = """
neighbor 5.0e-4 bin
neigh_modify delay 0
pair_style gran/hooke/history 5000.0 NULL 11200 NULL 0.1 0
pair_coeff * *
timestep 2e-7
velocity all set 0.0 0.0 0.0 units box
fix 1 all nve/sphere
fix 2 all gravity 0.0 vector 0 -1 0
fix 3 all fdrag
fix xwall all wall/gran 5000.0 NULL 11200 NULL 0.1 0 xplane 0.00 0.004
fix ywall all wall/gran 5000.0 NULL 11200 NULL 0.1 0 yplane 0.00 0.004
fix zwall all wall/gran 5000.0 NULL 11200 NULL 0.1 0 zplane 0.00 0.0005
"""
    c = _base(x, 8.3e-5, 2000.0, (0, 0, 0), (0.004, 0.004, 0.0005), ("f", "f", "f"), script, 4.0e-4,
              extra=dict(name="single_sphere", Uf=(0.0, 0.05, 0.0), g=(0.0, 0.0, 0.0), dt=2e-7, substeps=100))
    c["mesh_n"] = np.array([10, 10, 1], np.int32)
    return c


def uniform_fields(case, gamma=None):
    """prescribed cell fields standing in for the OpenFOAM solver: uniform Uf, hydrostatic grad p, given gamma."""
    C = int(np.prod(case["mesh_n"]))
    Uf = np.tile(np.asarray(case["Uf"], np.float64), (C, 1))
    gradp = np.tile(case["rhob"] * np.asarray(case["g"], np.float64), (C, 1))
    gam = np.zeros(C) if gamma is None else gamma
    return Uf, gam, gradp


def write_lammps_files(case, directory):
    """the same case as LAMMPS files: a `read_data` file (id type diameter density x y z) and an
    in.lammps in the style of the shipped cases.  Returns the script path."""
    import os
    data = os.path.join(directory, "IC.in")
    lo, hi = case["box_lo"], case["box_hi"]
    with open(data, "w") as f:
        f.write("LAMMPS data file\n\n%d atoms\n%d atom types\n\n" % (len(case["tag"]), case["ntypes"]))
        for k, ax in enumerate("xyz"):
            f.write("%.17g %.17g %slo %shi\n" % (lo[k], hi[k], ax, ax))
        f.write("\nAtoms\n\n")
        for i in range(len(case["tag"])):
            f.write("%d %d %.17g %.17g %.17g %.17g %.17g\n" % (case["tag"][i], case["type"][i], case["diam"][i], case["rho"][i],
                                                              case["x"][i, 0], case["x"][i, 1], case["x"][i, 2]))
    script = os.path.join(directory, "in.lammps")
    with open(script, "w") as f:
        f.write("atom_style sphere\natom_modify map array\nboundary %s\nnewton off\ncommunicate single vel yes\n" % " ".join(case["periodic"]))
        f.write("read_data %s\n" % data)
        f.write(case["script"])
    return script


def four_spheres_collide(variant="dia"):
    """the shipped regression cases cases/auto-testing/test-cases/multiParticlesCollideDia and ...CollideRho (in.lammps,
    IC_uniform.in, constant/{cloudProperties,transportProperties,environmentalProperties}, blockMeshDict): four spheres
    of different diameter ("dia") or different density ("rho") in water, two of them overlapping at the start, stock
    gran/hooke/history + wall/gran, SyamlalOBrien, dt_DEM = 1e-5, dt_fluid = 1e-3, subCycles 2; golden dump every 1000
    DEM steps in data/origin/p{1..4}.dat."""
    if variant == "rho":
        x = np.array([[5e-2, 7.5e-2, 5e-2], [9e-2, 8.5e-2, 5e-2], [9.1e-2, 8.5e-2, 5e-2], [1.7e-1, 7.5e-2, 5e-2]])
        d = np.array([1.5e-3, 1.5e-3, 1.5e-3, 1.5e-3])
        rho = np.array([4650.0, 3650.0, 2650.0, 1650.0])
    else:
        x = np.array([[5e-2, 7.5e-2, 5e-2], [9e-2, 8.5e-2, 5e-2], [9.2e-2, 8.5e-2, 5e-2], [1.7e-1, 6.5e-2, 5e-2]])
        d = np.array([3.5e-3, 3.0e-3, 2.5e-3, 2.0e-3])
        rho = 2650.0
    script = """
neighbor 0.02 bin
neigh_modify delay 0
pair_style gran/hooke/history 4910.0 NULL 0 NULL 0.15 0
pair_coeff * *
timestep 1e-5
velocity all set 0.0 0.0 0.0 units box
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
fix xwall all wall/gran 4910.0 NULL 0 NULL 0 0 xplane 0.00 0.20
fix ywall all wall/gran 4910.0 NULL 0 NULL 0 0 yplane 0.00 0.10
fix zwall all wall/gran 4910.0 NULL 0 NULL 0 0 zplane 0.00 0.10
"""
    c = _base(x, d, rho, (0, 0, 0), (0.2, 0.1, 0.1), ("f", "f", "f"), script, 5e-3,
              extra=dict(name="four_spheres_collide", Uf=(0.0, 0.0, 0.0), g=(0.0, -9.8, 0.0), dt=1e-5, substeps=50))
    c["mesh_n"] = np.array([40, 20, 1], np.int32)
    return c


def blockmesh_divide(lo, hi, n, expansion=1.0):
    """face coordinates of one blockMesh edge: n cells from lo to hi, `simpleGrading` expansion ratio = last cell size /
    first cell size (EXTERNAL OpenFOAM blockMesh lineDivide: lambda_i = (1 - g^i) / (1 - g^n), g = ratio^(1/(n-1)),
    uniform when |g - 1| <= 1e-5; restated from the published algorithm -- a host with OpenFOAM passes mesh.points())."""
    lam = np.arange(n + 1, dtype=np.float64) / n
    if n > 1:
        g = float(expansion) ** (1.0 / (n - 1))
        if abs(g - 1.0) > 1e-5:
            lam = (1.0 - g ** np.arange(n + 1)) / (1.0 - g ** n)
    f = lo + lam * (hi - lo)
    f[0], f[-1] = lo, hi
    return f


def blockmesh_stacked(xspec, yblocks, zspec):
    """axis-aligned hex blocks stacked along y (cases/example-cases/BL24-TH1, transport-vortex-dune; one block = the
    graded single-block cases): xspec / zspec = (lo, hi, n, expansion), yblocks = [(lo, hi, n, expansion), ...].
    Returns xf, yf, zf and the blockMesh cell label of every tensor cell (cells are numbered block by block, x fastest)."""
    xf = blockmesh_divide(*xspec); zf = blockmesh_divide(*zspec)
    ys, offs, nys = [], [], []
    off = 0
    nx, nz = len(xf) - 1, len(zf) - 1
    for b, (lo, hi, n, e) in enumerate(yblocks):
        f = blockmesh_divide(lo, hi, n, e)
        ys.append(f if b == 0 else f[1:])
        offs.append(off); nys.append(n)
        off += nx * n * nz
    yf = np.concatenate(ys)
    ny = len(yf) - 1
    label = np.zeros(nx * ny * nz, np.int32)
    j0 = 0
    for b, n in enumerate(nys):
        for k in range(nz):
            for j in range(n):
                t = np.arange(nx) + nx * ((j0 + j) + ny * k)
                label[t] = offs[b] + np.arange(nx) + nx * (j + n * k)
        j0 += n
    return xf, yf, zf, label


# ---- random packings (BASELINE.json: "synthetic random packings of the named N"; generator: sedifoam_b200/packing.py) ----
GRAN_WALL = "fix {id} all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 {plane} {lo} {hi}"


def _fill_box(T, scale, tiles, lo, hi, brick):
    """particles of the tiled packing inside the whole box, or -- brick = (procgrid, rank) -- inside this rank's brick of
    the engine's decomposition (LAMMPS `processors`: uniform split of [lo, hi], x fastest) plus a thin margin; the
    engine keeps the atoms it owns (read_data semantics), tags are the same whoever generates a particle."""
    tiles = np.asarray(tiles, np.float64)
    blo, bhi = np.zeros(3), tiles.copy()
    if brick is not None:
        grid, rank = brick
        L = T[2] * scale
        c = (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))
        for k in range(3):
            w = (hi[k] - lo[k]) / grid[k]
            blo[k] = max(0.0, (lo[k] + c[k] * w) / L - 0.02) if c[k] > 0 else 0.0
            bhi[k] = min(tiles[k], (lo[k] + (c[k] + 1) * w) / L + 0.02) if c[k] < grid[k] - 1 else tiles[k]
    xt, dt_, ids = packing.fill(T, blo, bhi, grid=tiles)
    if len(ids) and ids.max() >= 2 ** 31 - 2:
        raise ValueError("too many particles for 32-bit tags")
    return xt * scale, dt_ * scale, (ids + 1).astype(np.int32)


def random_bed(tiles=(5, 8, 5), d=5.0e-4, rho=2650.0, phi=0.58, skin_frac=0.25, dt=2.0e-6, kn=1.0e7, e=0.9, mu=0.4, head=0.5,
               seed=SEED, vjit=1.0e-3, brick=None, tile_n=TILE_N):
    """configs[2]: 1e6-particle bed as a disordered packing -- `tiles` periodic random tiles of `tile_n` spheres at solid
    fraction `phi` (about 2.4 touching pairs per particle with overlaps of order 1e-3 d: a settled bed under its own
    weight), granular walls one radius outside the cut faces, free head-room on top; Hertz-Mindlin pair + wall/granFix +
    gravity + fdrag.  brick = (procgrid, rank) generates only that rank's brick (multi-GPU)."""
    T = packing.tile(tile_n, phi, seed)
    L = T[2] * d
    r = 0.5 * d
    lo = np.zeros(3) - r
    hi = np.asarray(tiles, np.float64) * L + r
    hi[1] = tiles[1] * L * (1.0 + head)
    x, _, tags = _fill_box(T, d, tiles, lo, hi, brick)
    rng = np.random.default_rng(seed + 1 + (0 if brick is None else brick[1]))
    v = rng.uniform(-vjit, vjit, size=x.shape)
    w = dict(kn=kn, e=e, mu=mu)
    script = f"""
neighbor {skin_frac * d:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
{GRAN_WALL.format(id="xw", plane="xplane", lo="%.17g" % lo[0], hi="%.17g" % hi[0], **w)}
{GRAN_WALL.format(id="yw", plane="yplane", lo="%.17g" % lo[1], hi="%.17g" % hi[1], **w)}
{GRAN_WALL.format(id="zw", plane="zplane", lo="%.17g" % lo[2], hi="%.17g" % hi[2], **w)}
"""
    return _base(x, d, rho, lo, hi, ("f", "f", "f"), script, 4.0 * d, v=v, tags=tags,
                 extra=dict(name="random_bed", Uf=(0.0, 0.02, 0.0), g=(0.0, -9.8, 0.0), dt=dt, substeps=100, packing="random",
                            solid_fraction=phi, tiles=tuple(tiles)))


def random_column(tiles=(2, 5, 2), d=5.0e-4, rho=2650.0, phi=0.30, skin_frac=0.25, dt=2.0e-6, kn=1.0e7, e=0.9, mu=0.4,
                  seed=SEED, tile_n=TILE_N, head=0.12, brick=None):
    """configs[1]: 1e5 monodisperse spheres at phi = 0.30 (random, overlap-free) sedimenting in a column periodic in x and
    z with a granular floor."""
    T = packing.tile(tile_n, phi, seed, margin=0.02)
    L = T[2] * d
    lo = np.array([0.0, -0.5 * d, 0.0])
    hi = np.asarray(tiles, np.float64) * L
    hi[1] *= (1.0 + head)
    x, _, tags = _fill_box(T, d, tiles, lo, hi, brick)
    script = f"""
neighbor {skin_frac * d:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
fix yw all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane {lo[1]:.17g} NULL
"""
    return _base(x, d, rho, lo, hi, ("p", "f", "p"), script, 4.0 * d, tags=tags,
                 extra=dict(name="random_column", Uf=(0.0, 0.0, 0.0), g=(0.0, -9.8, 0.0), dt=dt, substeps=100, packing="random",
                            solid_fraction=phi, tiles=tuple(tiles)))


def random_cohesive_bed(tiles=(5, 8, 5), d=5.0e-5, rho=2650.0, phi=0.58, dt=2.0e-8, kn=1.0e7, e=0.9, mu=0.4, opt=1, vshear=0.01,
                        seed=SEED, tile_n=TILE_N, brick=None):
    """configs[3]: 1e6-particle cohesive silt bed (d = 50 um) under shear: periodic in x / z, wall/granFix floor, sheared
    wall/granFix lid resting on the bed, fix cohesive (opt 1), Hertz-Mindlin pair."""
    T = packing.tile(tile_n, phi, seed)
    L = T[2] * d
    r = 0.5 * d
    lo = np.array([0.0, -r, 0.0])
    hi = np.asarray(tiles, np.float64) * L
    hi[1] = tiles[1] * L + r - 2.0e-3 * d      # the lid presses on the topmost particles
    x, _, tags = _fill_box(T, d, tiles, lo, hi, brick)
    smax = 1.0e-6
    skin = max(0.25 * d, 2.0 * smax)
    script = f"""
neighbor {skin:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
fix 4 all cohesive 1e-20 1e-7 4e-10 {smax:.9g} {opt}
fix yb all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane {lo[1]:.17g} NULL
fix yt all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane NULL {hi[1]:.17g} shear x {vshear:.9g}
"""
    return _base(x, d, rho, lo, hi, ("p", "f", "p"), script, 4.0 * d, tags=tags,
                 extra=dict(name="random_cohesive_bed", Uf=(0.05, 0.0, 0.0), g=(0.0, -9.8, 0.0), dt=dt, substeps=100,
                            packing="random", solid_fraction=phi, tiles=tuple(tiles)))


def random_poly_lubricated(tiles=(5, 10, 5), dmin=3.0e-4, dmax=7.0e-4, rho=2650.0, phi=0.55, dt=2.0e-6, kn=1.0e7, e=0.9, mu=0.4,
                           visc=1.0e-3, seed=SEED, vjit=0.01, brick=None, tile_n=TILE_N, flaglog=0):
    """configs[4]: polydisperse (d ~ U[dmin, dmax]) dense periodic packing at phi = 0.55 with hybrid/overlay
    gran/hertzFix/history + lubricate/poly (full list, cutoff 1.5 dmax).  5 x 10 x 5 tiles = 1.25e6 particles; the 1e7
    weak-scaling point is tiles = (10, 20, 10) on a 2 x 2 x 2 processor grid (brick = (grid, rank)).
    cut_inner = 1.001 dmax: the reference switches lubrication off inside cut_inner (pair_lubricate_poly.cpp:294-297) and
    takes log(1/h) of the gap outside it (:308), so every overlapping pair must lie inside cut_inner or the reference
    itself produces NaN.  flaglog = 0 (squeeze term only): with flaglog = 1 the reference's log forms give NEGATIVE
    resistances for gaps larger than the smaller radius (log(1/h) < 0), which a global cutoff of 1.5 dmax cannot avoid in
    a polydisperse packing -- the reference objects themselves blow up within 40 steps on this bed (spins of 1e10 rad/s,
    checked with oracle/_ref); the parity tests keep flaglog = 1 on a dilute packing where it is stable."""
    dm = 0.5 * (dmin + dmax)
    T = packing.tile(tile_n, phi, seed, dlo=dmin / dm, dhi=dmax / dm)
    L = T[2] * dm
    lo = np.zeros(3)
    hi = np.asarray(tiles, np.float64) * L
    x, diam, tags = _fill_box(T, dm, tiles, lo, hi, brick)
    rng = np.random.default_rng(seed + 1 + (0 if brick is None else brick[1]))
    v = rng.uniform(-vjit, vjit, size=x.shape)
    skin = 0.1 * dmin
    script = f"""
neighbor {skin:.9g} bin
neigh_modify delay 0
pair_style hybrid/overlay gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 lubricate/poly {visc:.9g} {flaglog} 1 {1.001 * dmax:.9g} {1.5 * dmax:.9g}
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 3 all fdrag
"""
    return _base(x, diam, rho, lo, hi, ("p", "p", "p"), script, 4.0 * dmax, v=v, tags=tags,
                 extra=dict(name="random_poly_lubricated", Uf=(0.0, 0.0, 0.0), g=(0.0, 0.0, 0.0), dt=dt, substeps=100,
                            packing="random", solid_fraction=phi, tiles=tuple(tiles)))


# upflow below minimum fluidisation: ErgunWenYu drag 0.07 of the weight + pressure-gradient (buoyancy) force 0.38 of it --
# the bed stays packed (0.02 m/s, the round-1 value, lifts it: drag + buoyancy = 1.05 of the weight)
BED_UF = (0.0, 0.002, 0.0)


# ---- settled beds: a random column relaxed under gravity by the DEM itself, then repeated in x and z -------------------------
def settling_column(tile_n=TILE_N, height=8, d=5.0e-4, rho=2650.0, phi=0.58, skin_frac=0.25, dt=2.0e-6, kn=1.0e7, e=0.9, mu=0.4,
                    seed=SEED, head=0.5):
    """the input of tools/make_settled_column.py: `height` random tiles stacked on a granular floor, periodic in x and z."""
    T = packing.tile(tile_n, phi, seed)
    L = T[2] * d
    tiles = (1, height, 1)
    lo = np.array([0.0, -0.5 * d, 0.0])
    hi = np.array([L, height * L * (1.0 + head), L])
    x, _, tags = _fill_box(T, d, tiles, lo, hi, None)
    script = f"""
neighbor {skin_frac * d:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
fix yw all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane {lo[1]:.17g} NULL
"""
    return _base(x, d, rho, lo, hi, ("p", "f", "p"), script, 4.0 * d, tags=tags,
                 extra=dict(name="settling_column", Uf=BED_UF, g=(0.0, -9.8, 0.0), dt=dt, substeps=100, packing="random",
                            solid_fraction=phi, tiles=tiles, tile_n=tile_n))


_COLUMNS = {}


def load_column(name):
    """a settled periodic column written by tools/make_settled_column.py: dict(x, v, omega, L = (Lx, Lz), d, meta)"""
    import json
    import os
    if name not in _COLUMNS:
        path = name if os.path.exists(name) else os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", name)
        z = np.load(path)
        _COLUMNS[name] = dict(x=z["x"].astype(np.float64), v=z["v"].astype(np.float64), omega=z["omega"].astype(np.float64),
                              L=(float(z["Lx"]), float(z["Lz"])), d=float(z["d"]), top=float(z["x"][:, 1].max()),
                              meta=str(z["meta"]), fow=float(json.loads(str(z["meta"])).get("fluid_force_over_weight", 0.3)))
    return _COLUMNS[name]


def settled_bed(columns=(5, 5), column=None, rho=2650.0, skin_frac=0.25, dt=2.0e-6, kn=1.0e7, e=0.9, mu=0.4,
                head=0.5, brick=None):
    """configs[2]: the 1e6-particle bed of the benchmark -- a random packing SETTLED under gravity and the bench's fluid
    force (column `column`: 5000-sphere random tiles stacked 8 high on a granular floor, periodic in x / z, run to rest
    with the DEM of this library; tools/make_settled_column.py), repeated `columns` (integers) times in x and z.  The bed
    stays periodic in x and z -- exactly the state it was settled in; side walls one radius outside the cut planes would
    give the packing 1 % of lateral room and let it slump --, rests on a wall/granFix floor and has free head-room below
    a wall/granFix lid.  Every particle sits in a force-carrying contact network (about 2.5 touching pairs per particle),
    rows are ragged.  Velocities / spins are the column's residual ones; the contact history starts from zero (friction
    re-mobilises within micro-slips).  brick = (procgrid, rank): only that rank's brick (multi-GPU); tags do not depend
    on who generates a particle."""
    import os
    column = column or os.environ.get("SEDI_COLUMN", "column_5000x8.npz")
    C = load_column(column)
    d = C["d"]
    r = 0.5 * d
    Lx, Lz = C["L"]
    gx, gz = int(round(columns[0])), int(round(columns[1]))
    if gx < 1 or gz < 1 or abs(gx - columns[0]) > 1e-9 or abs(gz - columns[1]) > 1e-9:
        raise ValueError("settled_bed: whole columns only (the bed is periodic in x and z)")
    lo = np.array([0.0, -r, 0.0])
    hi = np.array([gx * Lx, (C["top"] + r) * (1.0 + head), gz * Lz])
    xr = [0.0, gx * Lx]
    zr = [0.0, gz * Lz]
    if brick is not None:
        grid, rank = brick
        c = (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))
        if grid[1] != 1:
            raise ValueError("settled_bed bricks split x and z only")
        for k, rng_ in ((0, xr), (2, zr)):
            w = (hi[k] - lo[k]) / grid[k]
            m = 0.02 * (Lx if k == 0 else Lz)
            if c[k] > 0:
                rng_[0] = max(rng_[0], lo[k] + c[k] * w - m)
            if c[k] < grid[k] - 1:
                rng_[1] = min(rng_[1], lo[k] + (c[k] + 1) * w + m)
    n0 = len(C["x"])
    xs, vs, ws, ts = [], [], [], []
    for ix in range(int(np.floor(xr[0] / Lx + 1e-12)), int(np.ceil(xr[1] / Lx - 1e-12))):
        for iz in range(int(np.floor(zr[0] / Lz + 1e-12)), int(np.ceil(zr[1] / Lz - 1e-12))):
            x = C["x"] + np.array([ix * Lx, 0.0, iz * Lz])
            m = (x[:, 0] >= xr[0]) & (x[:, 0] < xr[1]) & (x[:, 2] >= zr[0]) & (x[:, 2] < zr[1])
            xs.append(x[m]); vs.append(C["v"][m]); ws.append(C["omega"][m])
            ts.append((ix * gz + iz) * n0 + np.nonzero(m)[0] + 1)
    x = np.concatenate(xs); v = np.concatenate(vs); om = np.concatenate(ws); tags = np.concatenate(ts).astype(np.int32)
    w = dict(kn=kn, e=e, mu=mu)
    script = f"""
neighbor {skin_frac * d:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
{GRAN_WALL.format(id="yw", plane="yplane", lo="%.17g" % lo[1], hi="%.17g" % hi[1], **w)}
"""
    return _base(x, d, rho, lo, hi, ("p", "f", "p"), script, 4.0 * d, v=v, tags=tags,
                 extra=dict(name="settled_bed", Uf=BED_UF, g=(0.0, -9.8, 0.0), dt=dt, substeps=100, packing="settled random",
                            omega=om, columns=(gx, gz), column=column, column_meta=C["meta"],
                            fluid_force_over_weight=C["fow"]))


def settled_cohesive_bed(columns=(5, 5), column=None, d=5.0e-5, rho=2650.0, dt=2.0e-8, kn=1.0e7, e=0.9, mu=0.4, opt=1, vshear=0.01, brick=None):
    """configs[3]: 1e6-particle cohesive silt bed (d = 50 um) under shear -- the settled random column scaled to the silt
    diameter, periodic in x / z, wall/granFix floor, a sheared wall/granFix lid pressing on the topmost particles,
    fix cohesive (opt 1) on every list pair, Hertz-Mindlin pair."""
    base = settled_bed(columns=columns, column=column, brick=brick)
    d0 = float(base["diam"][0])
    sc = d / d0
    x = base["x"] * sc
    r = 0.5 * d
    lo = base["box_lo"] * sc
    hi = base["box_hi"] * sc
    C = load_column(base["column"])
    hi[1] = C["top"] * sc + r - 1.0e-2 * d      # the lid presses on the topmost particles
    smax = 1.0e-6
    skin = max(0.25 * d, 2.0 * smax)
    script = f"""
neighbor {skin:.9g} bin
neigh_modify delay 0
pair_style gran/hertzFix/history {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1
pair_coeff * *
timestep {dt:.9g}
fix 1 all nve/sphere
fix 2 all gravity 9.8 vector 0 -1 0
fix 3 all fdrag
fix 4 all cohesive 1e-20 1e-7 4e-10 {smax:.9g} {opt}
fix yb all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane {lo[1]:.17g} NULL
fix yt all wall/granFix {kn:.9g} NULL {e:.9g} NULL {mu:.9g} 1 yplane NULL {hi[1]:.17g} shear x {vshear:.9g}
"""
    return _base(x, d, rho, lo, hi, ("p", "f", "p"), script, 4.0 * d, tags=base["tag"],
                 extra=dict(name="settled_cohesive_bed", Uf=(0.05, 0.0, 0.0), g=(0.0, -9.8, 0.0), dt=dt, substeps=100,
                            packing="settled random (scaled)", columns=base["columns"], column=base["column"], fluid_force_over_weight=0.3))


def bench_fluid_force(case):
    """the constant per-particle fluid force the benchmark's host side hands over: a fixed fraction of the weight, upwards
    -- for a settled bed the fraction the bed was settled with (mean ErgunWenYu drag + pressure-gradient force of the
    bench's prescribed fluid fields), so that the bed stays at rest whichever side computes the force"""
    m = case["rho"] * np.pi / 6.0 * case["diam"] ** 3
    f = float(case.get("fluid_force_over_weight", 0.3))
    return np.tile([0.0, f * 9.8, 0.0], (len(m), 1)) * m[:, None]
