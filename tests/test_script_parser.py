"""CPU: the LAMMPS input-script front end (sedifoam_b200/csrc/lmp_script.hpp) against HAND-WRITTEN expectations for every
shipped in.lammps that stays inside the hot path's command subset (tests/golden/in_lammps/*.in are the reference's case
inputs, cases/**/in.lammps).  The oracle reads scripts through the same header, so a parsing mistake would be invisible
to the parity tests; the expectations below were written from the script text, not from the parser.

Rules restated where the expectation is not literal: `kt = NULL -> 2/7 kn`, `gammat = NULL -> gamman / 2`, `dampflag 0 ->
gammat = 0` (interfaceToLammps/pair_gran_hertzFix_history.cpp:293-317 and the stock gran styles), numbers go through
atof ("1.91+e2" in the expMueller inputs therefore reads 1.91), groups get bits in order of definition after `all`."""
import os

import pytest

import sedifoam_b200 as sb

HERE = os.path.dirname(os.path.abspath(__file__))
NVE, GRAVITY, FDRAG, COHESIVE, WALL, FREEZE = 0, 1, 2, 3, 4, 5
X, Y, Z = 0, 1, 2
COLS = "id type diameter mass x y z vx vy vz"


def walls(kn, gn, xmu, planes):
    return [(name, WALL, 1, dict(kn=kn, kt=kn * 2.0 / 7.0, gamman=gn, gammat=0.0, xmu=xmu, wallstyle=st, lo=lo, hi=hi)) for name, st, lo, hi in planes]


def std_fixes(g=9.8, grp=1, gdir=(0, -1, 0), order=("nve", "gravity")):
    f = {"nve": (NVE, grp, {}), "gravity": (GRAVITY, grp, dict(g=g, gdir=list(gdir)))}
    out = [(str(k + 1), f[o][0], f[o][1], f[o][2]) for k, o in enumerate(order)]
    return out + [("3", FDRAG, grp, dict(carrier_rho=0.0))]


EXPECT = {
    "expMueller06": dict(periodic=[0, 0, 0], skin=5.0e-4, dt=4.0e-6, gran=(200.0, 1.91, 0.1), procgrid=[0, 0, 0],
                         fixes=std_fixes() + walls(200.0, 1.91, 0.1, [("xwall", X, 0.0, 0.044), ("ywall", Y, 0.0, 0.12), ("zwall", Z, 0.0, 0.01)]),
                         dump=(1000000, COLS)),
    "expMueller09": dict(periodic=[0, 0, 0], skin=5.0e-4, dt=4.0e-6, gran=(200.0, 1.91, 0.1), procgrid=[2, 1, 1],
                         fixes=std_fixes() + walls(200.0, 1.91, 0.1, [("xwall", X, 0.0, 0.044), ("ywall", Y, 0.0, 0.12), ("zwall", Z, 0.0, 0.01)]),
                         dump=(1000000, COLS)),
    "expWachem_PCM": dict(periodic=[0, 0, 0], skin=5.0e-4, dt=5.0e-6, gran=(10000.0, 15100.0, 0.3), procgrid=[0, 0, 0],
                          fixes=std_fixes() + walls(10000.0, 15100.0, 0.3, [("xwall", X, 0.0, 0.09), ("ywall", Y, 0.0, 0.5), ("zwall", Z, 0.0, 0.008)]),
                          dump=(10000, COLS)),
    "multiParticlesCollideDia": dict(periodic=[0, 0, 0], skin=0.02, dt=1.0e-5, gran=(4910.0, 0.0, 0.15), procgrid=[0, 0, 0],
                                     fixes=std_fixes() + walls(4910.0, 0.0, 0.0, [("xwall", X, 0.0, 0.2), ("ywall", Y, 0.0, 0.1), ("zwall", Z, 0.0, 0.1)]),
                                     dump=(1000, COLS)),
    "multiParticlesCollideRho": dict(periodic=[0, 0, 0], skin=0.02, dt=1.0e-5, gran=(4910.0, 0.0, 0.15), procgrid=[0, 0, 0],
                                     fixes=std_fixes() + walls(4910.0, 0.0, 0.0, [("xwall", X, 0.0, 0.2), ("ywall", Y, 0.0, 0.1), ("zwall", Z, 0.0, 0.1)]),
                                     dump=(1000, COLS)),
    "xiaocase1": dict(periodic=[0, 0, 0], skin=5.0e-4, dt=1.0e-5, gran=(4910.0, 8090.0, 0.15), procgrid=[2, 1, 1],
                      fixes=std_fixes() + walls(4910.0, 8090.0, 0.0, [("xwall", X, 0.0, 0.04), ("ywall", Y, 0.0, 0.2), ("zwall", Z, 0.0, 0.0075)]),
                      dump=(100000, COLS)),
    "xiaocase3": dict(periodic=[0, 0, 0], skin=5.0e-4, dt=2.0e-7, gran=(5000.0, 11200.0, 0.1), procgrid=[0, 0, 0],
                      fixes=std_fixes(g=0.0) + walls(5000.0, 11200.0, 0.1, [("xwall", X, 0.0, 0.004), ("ywall", Y, 0.0, 0.004), ("zwall", Z, 0.0, 0.0005)]),
                      dump=(1000, COLS)),
    "addDeleteParticles": dict(periodic=[0, 0, 0], skin=1.0e-4, dt=1.0e-5, gran=(800.0, 59.0, 0.1), procgrid=[0, 0, 0], groups=dict(all=1, bottom=2, active=4),
                               freeze_group_bit=2,
                               fixes=std_fixes(g=9.81, order=("gravity", "nve")) + [("4", FREEZE, 2, {})] +
                               walls(800.0, 59.0, 0.1, [("xwall", X, -0.5, 0.5), ("ywall", Y, -0.5, 0.5), ("zwall", Z, -0.5, 0.5)]),
                               dump=(10000, COLS)),
    "multiParticles": dict(periodic=[0, 0, 0], skin=0.02, dt=1.0e-5, gran=(4910.0, 0.0, 0.15), procgrid=[3, 1, 1],
                           fixes=std_fixes() + walls(4910.0, 0.0, 0.0, [("xwall", X, 0.0, 0.2), ("ywall", Y, 0.0, 0.2), ("zwall", Z, 0.0, 0.1)]),
                           dump=(1000, "id diameter type mass x y z vx vy vz")),
    "BL24-TH1": dict(periodic=[1, 0, 1], skin=5.0e-4, dt=5.0e-6, gran=(20.0, 7910.0, 0.4), procgrid=[0, 0, 0], groups=dict(all=1, bottom=2, active=4),
                     fixes=std_fixes(grp=4) + walls(20.0, 7910.0, 0.4, [("ywall", Y, -0.01, 0.014)]), dump=None),
    "jetFlow": dict(periodic=[0, 0, 0], skin=1.0e-3, dt=1.0e-6, gran=(200.0, 24300.0, 0.1), procgrid=[0, 0, 0], groups=dict(all=1, bottom=2, active=4),
                    fixes=std_fixes(g=9.81, grp=4, gdir=(0, 1, 0), order=("gravity", "nve")) +
                    walls(200.0, 24300.0, 0.1, [("ywall", Y, 0.0, 0.3), ("xwall", X, -0.05, 0.05), ("zwall", Z, -0.05, 0.05)]),
                    dump=(50000, COLS)),
    "transport-bedload": dict(periodic=[1, 0, 1], skin=5.0e-4, dt=2.0e-6, gran=(2000.0, 56000.0, 0.1), procgrid=[0, 0, 0],
                              groups=dict(all=1, bottom=2, active=4), freeze_group_bit=2,
                              fixes=std_fixes() + [("4", FREEZE, 2, {})] + walls(2000.0, 56000.0, 0.1, [("ywall", Y, 0.0, 0.04)]), dump=(100000, COLS)),
}


def _feed(name):
    e = sb.Lammps()
    for ln in open(os.path.join(HERE, "golden", "in_lammps", name + ".in")):
        if ln.split() and ln.split()[0] in ("read_data", "run"):   # data files are not shipped with the fixture; nothing is run here
            continue
        e.command(ln)
    return e.config()


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_shipped_script_is_understood_as_written(name):
    sb.build_library()
    ex = EXPECT[name]
    c = _feed(name)
    assert c["periodic"] == ex["periodic"] and c["newton_pair"] == 0 and c["neigh_modify_seen"] == 1
    assert c["skin"] == ex["skin"] and c["dt"] == ex["dt"] and c["procgrid"] == ex["procgrid"]
    kn, gn, xmu = ex["gran"]
    assert c["pair"] == 2                                    # gran/hooke/history in every shipped script
    assert c["gran"] == dict(kn=kn, kt=kn * 2.0 / 7.0, gamman=gn, gammat=0.0, xmu=xmu, dampflag=0)
    assert c["lub"]["enabled"] == 0
    assert c["groups"] == ex.get("groups", dict(all=1))
    assert c["freeze_group_bit"] == ex.get("freeze_group_bit", 0)
    assert [f["id"] for f in c["fixes"]] == [f[0] for f in ex["fixes"]]    # post_force order = script order
    nw = 0
    for got, (fid, kind, bit, extra) in zip(c["fixes"], ex["fixes"]):
        assert got["kind"] == kind and got["groupbit"] == bit, fid
        for k, v in extra.items():
            if k in ("kn", "kt", "gamman", "gammat", "xmu"):
                assert got["wall"][k] == v, (fid, k)
            else:
                assert got[k] == v, (fid, k)
        if kind == WALL:
            assert got["wall_index"] == nw and got["wiggle"] == 0 and got["wshear"] == 0
            nw += 1
    assert c["nwalls"] == nw
    if ex["dump"] is None:
        assert c["dumps"] == []
    else:
        assert len(c["dumps"]) == 1 and c["dumps"][0]["every"] == ex["dump"][0] and c["dumps"][0]["columns"].split() == ex["dump"][1].split()
        assert c["dumps"][0]["path"] in ("snapshot.bubblemd",) and c["dumps"][0]["groupbit"] == 1


def test_custom_styles_of_the_hot_path():
    """the north-star styles no shipped case uses: argument order and defaults from the reference sources --
    pair_gran_hertzFix_history.cpp:293-317, fix_cohesive.cpp:38-47, fix_wall_granFix.cpp:47-141, fix_fluid_drag.cpp:41-59,
    EXTERNAL pair lubricate/poly mu flaglog flagfld cutinner cutoff [flagHI flagVF]"""
    sb.build_library()
    e = sb.Lammps()
    for ln in """
boundary ff ff pp
newton off
neighbor 1.25e-4 bin
neigh_modify delay 0 every 1 check yes
pair_style hybrid/overlay gran/hertzFix/history 1e7 NULL 0.9 NULL 0.4 1 lubricate/poly 1e-3 1 0 7.0e-4 1.05e-3 1 0
pair_coeff * *
timestep 2e-6
group heavy type 2
fix a all nve/sphere
fix b heavy fdrag 1000.7
fix c all cohesive 1e-20 1e-7 4e-10 1e-6 1
fix d all wall/granFix 2e6 1e6 0.8 NULL 0.3 1 yplane 0.0 NULL shear x 0.05
fix e all wall/granFix 2e6 NULL 0.8 0.2 0.3 1 zcylinder 0.02 wiggle z 1e-4 0.01
""".strip().splitlines():
        e.command(ln)
    c = e.config()
    assert c["pair"] == 3 and c["gran"] == dict(kn=1e7, kt=1e7 * 2.0 / 7.0, gamman=0.9, gammat=0.45, xmu=0.4, dampflag=1)
    assert c["lub"] == dict(enabled=1, mu=1e-3, flaglog=1, flagfld=0, cut_inner=7.0e-4, cut_global=1.05e-3, flagHI=1, flagVF=0)
    f = {x["id"]: x for x in c["fixes"]}
    assert f["b"]["groupbit"] == 2 and f["b"]["carrier_rho"] == 1000.0          # atoi, fix_fluid_drag.cpp:53
    assert (f["c"]["ah"], f["c"]["lam"], f["c"]["smin"], f["c"]["smax"], f["c"]["opt"]) == (1e-20, 1e-7, 4e-10, 1e-6, 1)
    d = f["d"]
    assert d["wall"] == dict(kn=2e6, kt=1e6, gamman=0.8, gammat=0.4, xmu=0.3, dampflag=1)
    assert d["wallstyle"] == Y and d["lo"] == 0.0 and d["hi"] >= 1e20 and d["wshear"] == 1 and d["axis"] == X and d["vshear"] == 0.05 and d["wall_index"] == 0
    w = f["e"]
    assert w["wall"]["kt"] == 2e6 * 2.0 / 7.0 and w["wall"]["gammat"] == 0.2
    assert w["wallstyle"] == 3 and w["cylradius"] == 0.02 and w["wiggle"] == 1 and w["axis"] == Z and w["amplitude"] == 1e-4 and w["period"] == 0.01 and w["wall_index"] == 1


def test_neigh_modify_other_than_delay0_is_refused():
    import subprocess
    import sys
    code = "import sedifoam_b200 as sb; e = sb.Lammps(); e.command('neigh_modify delay 10')"
    r = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(HERE), capture_output=True, text=True)
    assert r.returncode != 0 and "neigh_modify" in r.stderr
