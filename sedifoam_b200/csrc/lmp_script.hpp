// lmp_script.hpp -- host-side mirror of the LAMMPS input-script plug-in API that sediFoam drives.
//
// sediFoam feeds ./in.lammps line by line into LAMMPS (reference: lammpsFoam/softParticleCloud.C:85-115,
// `lmp_->input->one(line)`), and its custom styles are registered by name in
// interfaceToLammps/style_user.h:43-107.  This header parses that command subset into a flat POD-ish
// `SimConfig` that the CUDA engine (sedi_engine.cu) uploads to the device.  The oracle (oracle/) includes
// this same header so that product and checker agree on what a script *means*; all arithmetic lives elsewhere.
//
// Command subset (every command seen in /root/reference/cases/**/in.lammps):
//   atom_style sphere | atom_modify | boundary | newton | communicate | processors | read_data | neighbor |
//   neigh_modify | pair_style {gran/hertzFix/history, gran/hooke/history, gran/hooke, lubricate/poly,
//   hybrid/overlay, none} | pair_coeff | timestep | velocity <grp> set | group <id> {type, id, subtract, union} |
//   fix {nve/sphere, gravity, fdrag, cohesive, wall/granFix, wall/gran, freeze} | run N [pre no post no] |
//   dump, thermo, thermo_style, thermo_modify, restart, units lj, dimension 3, region/create_box (box only)
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace sedi {

enum PairKind { PAIR_NONE = 0, PAIR_HOOKE = 1, PAIR_HOOKE_HISTORY = 2, PAIR_HERTZFIX_HISTORY = 3 };
enum FixKind { FIX_NVE_SPHERE = 0, FIX_GRAVITY = 1, FIX_FDRAG = 2, FIX_COHESIVE = 3, FIX_WALL_GRAN = 4, FIX_FREEZE = 5 };
enum WallStyle { XPLANE = 0, YPLANE = 1, ZPLANE = 2, ZCYLINDER = 3 };  // fix_wall_granFix.cpp:36

static const double WALL_BIG = 1.0e20;  // fix_wall_granFix.cpp:40

struct GranParams {  // pair_gran_hertzFix_history.cpp:293-317 (settings)
  double kn, kt, gamman, gammat, xmu;
  int dampflag;
};

struct LubParams {  // EXTERNAL PairLubricate::settings
  int enabled;
  double mu;
  int flaglog, flagfld;
  double cut_inner, cut_global;
  int flagHI, flagVF;
};

struct FixSpec {
  int kind;
  char id[32];
  int groupbit;
  // gravity: magnitude + normalised direction (EXTERNAL fix gravity ... vector)
  double g, gdir[3];
  // fdrag: fix_fluid_drag.cpp:41-59 (carrier_rho is parsed with atoi -- integer truncation is reference behaviour)
  double carrier_rho;
  // cohesive: fix_cohesive.cpp:38-47
  double ah, lam, smin, smax;
  int opt;
  // wall/granFix and stock wall/gran: fix_wall_granFix.cpp:44-168
  GranParams wall;
  int wallstyle;
  double lo, hi, cylradius;
  int wiggle, wshear, axis;
  double amplitude, period, vshear;
  long long time_origin;
  int wall_index;  // running index among wall fixes (selects the per-atom shear array)
};

struct AtomData {  // atom_style sphere, read_data line: id type diameter density x y z
  std::vector<int> tag, type;
  std::vector<double> x, v, omega;  // interleaved xyz
  std::vector<double> radius, rmass;
  size_t size() const { return tag.size(); }
};

struct Group {
  std::string name;
  int bit;
};

// `dump ID group custom N file field...` -- the golden-file format of the shipped cases
// (e.g. cases/auto-testing/test-cases/multiParticlesCollideDia/in.lammps:31).  Text layout = EXTERNAL LAMMPS
// DumpCustom: header items TIMESTEP / NUMBER OF ATOMS / BOX BOUNDS / ATOMS, one row per atom, "%d " for integer
// fields and "%g " for doubles.  Rows are written in ascending id (`dump_modify sort id`).
enum DumpField { DF_ID, DF_TYPE, DF_DIAMETER, DF_RADIUS, DF_MASS, DF_X, DF_Y, DF_Z, DF_VX, DF_VY, DF_VZ, DF_FX, DF_FY, DF_FZ,
                 DF_OMEGAX, DF_OMEGAY, DF_OMEGAZ, DF_TQX, DF_TQY, DF_TQZ };
struct DumpSpec {
  std::string id, path, columns;
  int groupbit;
  long long every, last_written;
  std::vector<int> fields;
  FILE *fp;
  DumpSpec() : groupbit(1), every(0), last_written(-1), fp(0) {}
};

struct SimConfig {
  std::vector<DumpSpec> dumps;
  std::string boundary_str[3];
  int periodic[3];
  double boxlo[3], boxhi[3];
  int have_box;
  int ntypes;
  double skin;
  double dt;
  int newton_pair;
  int pair;  // PairKind
  GranParams gran;
  LubParams lub;
  std::vector<FixSpec> fixes;
  std::vector<Group> groups;
  int nwalls;
  int procgrid[3];
  long long ntimestep;
  int freeze_group_bit;  // EXTERNAL: pair styles read the group bit of `fix freeze`
  int neigh_modify_seen; // LAMMPS' default is `delay 10`; the engine behaves as `delay 0` and says so when the command is absent
  SimConfig() {
    memset(periodic, 0, sizeof(periodic));
    for (int d = 0; d < 3; d++) { boxlo[d] = 0; boxhi[d] = 1; procgrid[d] = 0; }
    have_box = 0; ntypes = 1; skin = 0.3; dt = 0.005; newton_pair = 1; pair = PAIR_NONE;
    memset(&gran, 0, sizeof(gran)); memset(&lub, 0, sizeof(lub)); lub.flagHI = 1; lub.flagVF = 1;
    nwalls = 0; ntimestep = 0; freeze_group_bit = 0; neigh_modify_seen = 0;
    Group g; g.name = "all"; g.bit = 1; groups.push_back(g);
    for (int d = 0; d < 3; d++) boundary_str[d] = "pp";
  }
  int find_group(const std::string &n) const {
    for (size_t i = 0; i < groups.size(); i++) if (groups[i].name == n) return groups[i].bit;
    return 0;
  }
};

// The literal the reference uses for pi when converting mass<->density at the boundary
// (library.cpp:200, :460; fix_fluid_drag.cpp:147).  Digits differ from pi at the 13th place -- kept on purpose.
static const double SEDI_PI_LIBRARY = 3.14159265358917323846;
// MathConst::MY_PI (EXTERNAL math_const.h)
static const double SEDI_MY_PI = 3.14159265358979323846;

inline void fatal(const char *msg, const char *detail = "") {
  // reference convention: print and abort (library.cpp:380-383; LAMMPS error->all)
  fprintf(stderr, "ERROR: %s %s\n", msg, detail);
  fflush(stderr);
  abort();
}

inline std::vector<std::string> tokenize(const char *line) {
  std::vector<std::string> out;
  std::string cur;
  for (const char *p = line; *p; ++p) {
    if (*p == '#') break;
    if (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') {
      if (!cur.empty()) { out.push_back(cur); cur.clear(); }
    } else cur.push_back(*p);
  }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

struct ScriptAction {
  enum Kind { NONE, RUN, READ_DATA, WRITE_RESTART, READ_RESTART, RESTART_EVERY } kind;
  long long nsteps;
  std::string path, path2;
  ScriptAction() : kind(NONE), nsteps(0) {}
};

class Script {
 public:
  SimConfig cfg;
  AtomData atoms;
  // per-atom group mask, rebuilt by group commands (bit0 = all)
  std::vector<int> mask;

  static GranParams parse_gran(const std::vector<std::string> &a, size_t o, const char *what) {
    // identical rules to PairGranHertzFixHistory::settings (pair_gran_hertzFix_history.cpp:293-317)
    // and FixWallGranFix ctor (fix_wall_granFix.cpp:55-72); nktv2p = 1 in lj units.
    if (a.size() < o + 6) fatal("Illegal command (granular coefficients)", what);
    GranParams p;
    p.kn = atof(a[o].c_str());
    p.kt = (a[o + 1] == "NULL") ? p.kn * 2.0 / 7.0 : atof(a[o + 1].c_str());
    p.gamman = atof(a[o + 2].c_str());
    p.gammat = (a[o + 3] == "NULL") ? 0.5 * p.gamman : atof(a[o + 3].c_str());
    p.xmu = atof(a[o + 4].c_str());
    p.dampflag = atoi(a[o + 5].c_str());
    if (p.dampflag == 0) p.gammat = 0.0;
    if (p.kn < 0.0 || p.kt < 0.0 || p.gamman < 0.0 || p.gammat < 0.0 || p.xmu < 0.0 || p.xmu > 10000.0 ||
        p.dampflag < 0 || p.dampflag > 1)
      fatal("Illegal command (granular coefficients out of range)", what);
    return p;
  }

  void read_data(const std::string &path) {
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) fatal("Cannot open data file", path.c_str());
    char buf[1024];
    long natoms = 0;
    bool in_atoms = false;
    if (!fgets(buf, sizeof(buf), fp)) fatal("Empty data file", path.c_str());  // title line
    while (fgets(buf, sizeof(buf), fp)) {
      std::vector<std::string> t = tokenize(buf);
      if (t.empty()) continue;
      if (!in_atoms) {
        if (t.size() >= 2 && t[1] == "atoms") natoms = atol(t[0].c_str());
        else if (t.size() >= 3 && t[1] == "atom" && t[2] == "types") cfg.ntypes = atoi(t[0].c_str());
        else if (t.size() >= 4 && t[2] == "xlo") { cfg.boxlo[0] = atof(t[0].c_str()); cfg.boxhi[0] = atof(t[1].c_str()); cfg.have_box = 1; }
        else if (t.size() >= 4 && t[2] == "ylo") { cfg.boxlo[1] = atof(t[0].c_str()); cfg.boxhi[1] = atof(t[1].c_str()); }
        else if (t.size() >= 4 && t[2] == "zlo") { cfg.boxlo[2] = atof(t[0].c_str()); cfg.boxhi[2] = atof(t[1].c_str()); }
        else if (t[0] == "Atoms") in_atoms = true;
        continue;
      }
      if (t[0] == "Velocities") break;
      if (t.size() < 7) continue;
      double d = atof(t[2].c_str()), rho = atof(t[3].c_str());
      double xyz[3] = {atof(t[4].c_str()), atof(t[5].c_str()), atof(t[6].c_str())};
      add_atom(atoi(t[0].c_str()), atoi(t[1].c_str()), d, rho, xyz, 0);
    }
    fclose(fp);
    if ((long)atoms.size() != natoms) fatal("Did not assign all atoms correctly", path.c_str());
  }

  // radius = d/2 ; rmass = 4 pi/3 r^3 rho with MY_PI (EXTERNAL AtomVecSphere::data_atom)
  void add_atom(int tag, int type, double diameter, double density, const double *xyz, const double *vel) {
    double r = 0.5 * diameter;
    atoms.tag.push_back(tag);
    atoms.type.push_back(type);
    atoms.radius.push_back(r);
    atoms.rmass.push_back(4.0 * SEDI_MY_PI / 3.0 * r * r * r * density);
    for (int d = 0; d < 3; d++) {
      atoms.x.push_back(xyz[d]);
      atoms.v.push_back(vel ? vel[d] : 0.0);
      atoms.omega.push_back(0.0);
    }
    mask.push_back(1);
  }

  void apply_group(const std::vector<std::string> &a) {
    if (a.size() < 3) fatal("Illegal group command");
    int bit = cfg.find_group(a[1]);
    if (!bit) {
      if (cfg.groups.size() >= 31) fatal("Too many groups");
      Group g; g.name = a[1]; g.bit = 1 << (int)cfg.groups.size();
      cfg.groups.push_back(g); bit = g.bit;
    }
    size_t n = atoms.size();
    if (a[2] == "type" || a[2] == "id") {
      const std::vector<int> &key = (a[2] == "type") ? atoms.type : atoms.tag;
      for (size_t k = 3; k < a.size(); k++) {
        int lo, hi;
        size_t c = a[k].find(':');
        if (c == std::string::npos) lo = hi = atoi(a[k].c_str());
        else { lo = atoi(a[k].substr(0, c).c_str()); hi = atoi(a[k].substr(c + 1).c_str()); }
        for (size_t i = 0; i < n; i++) if (key[i] >= lo && key[i] <= hi) mask[i] |= bit;
      }
    } else if (a[2] == "subtract") {
      if (a.size() < 5) fatal("Illegal group subtract command");
      int b0 = cfg.find_group(a[3]);
      for (size_t i = 0; i < n; i++) {
        bool in = (mask[i] & b0) != 0;
        for (size_t k = 4; k < a.size(); k++) if (mask[i] & cfg.find_group(a[k])) in = false;
        if (in) mask[i] |= bit;
      }
    } else if (a[2] == "union") {
      for (size_t i = 0; i < n; i++)
        for (size_t k = 3; k < a.size(); k++) if (mask[i] & cfg.find_group(a[k])) mask[i] |= bit;
    } else fatal("Unsupported group style", a[2].c_str());
  }

  void apply_fix(const std::vector<std::string> &a) {
    if (a.size() < 4) fatal("Illegal fix command");
    FixSpec f;
    memset(&f, 0, sizeof(f));
    snprintf(f.id, sizeof(f.id), "%s", a[1].c_str());
    f.groupbit = cfg.find_group(a[2]);
    if (!f.groupbit) fatal("Could not find fix group ID", a[2].c_str());
    const std::string &st = a[3];
    if (st == "nve/sphere") f.kind = FIX_NVE_SPHERE;
    else if (st == "gravity") {
      // EXTERNAL fix gravity: `gravity g vector x y z` -> direction normalised
      f.kind = FIX_GRAVITY;
      if (a.size() < 9 || a[5] != "vector") fatal("Only `fix gravity g vector x y z` is supported");
      f.g = atof(a[4].c_str());
      double v[3] = {atof(a[6].c_str()), atof(a[7].c_str()), atof(a[8].c_str())};
      double len = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      for (int d = 0; d < 3; d++) f.gdir[d] = v[d] / len;
    } else if (st == "fdrag") {
      f.kind = FIX_FDRAG;  // fix_fluid_drag.cpp:53-58
      f.carrier_rho = (a.size() == 5) ? (double)atoi(a[4].c_str()) : 0.0;
    } else if (st == "cohesive") {
      f.kind = FIX_COHESIVE;  // fix_cohesive.cpp:41-47 (narg must be 8)
      if (a.size() != 9) fatal("Illegal fix cohesive command");
      f.ah = atof(a[4].c_str()); f.lam = atof(a[5].c_str()); f.smin = atof(a[6].c_str());
      f.smax = atof(a[7].c_str()); f.opt = atoi(a[8].c_str());
      if (f.opt != 0 && f.opt != 1) fatal("invalid option for cohesive force model");  // fix_cohesive.cpp:252
    } else if (st == "wall/granFix" || st == "wall/gran") {
      f.kind = FIX_WALL_GRAN;
      if (a.size() < 11) fatal("Illegal fix wall/granFix command");
      f.wall = parse_gran(a, 4, "fix wall/granFix");
      size_t i = 10;
      f.lo = -WALL_BIG; f.hi = WALL_BIG;
      if (a[i] == "xplane" || a[i] == "yplane" || a[i] == "zplane") {
        if (a.size() < i + 3) fatal("Illegal fix wall/granFix command");
        f.wallstyle = (a[i] == "xplane") ? XPLANE : (a[i] == "yplane") ? YPLANE : ZPLANE;
        if (a[i + 1] != "NULL") f.lo = atof(a[i + 1].c_str());
        if (a[i + 2] != "NULL") f.hi = atof(a[i + 2].c_str());
        i += 3;
      } else if (a[i] == "zcylinder") {
        if (a.size() < i + 2) fatal("Illegal fix wall/granFix command");
        f.wallstyle = ZCYLINDER; f.lo = f.hi = 0.0; f.cylradius = atof(a[i + 1].c_str());
        i += 2;
      } else fatal("Illegal fix wall/granFix wallstyle", a[i].c_str());
      while (i < a.size()) {
        if (a[i] == "wiggle") {
          if (i + 4 > a.size()) fatal("Illegal fix wall/granFix command");
          f.axis = (a[i + 1] == "x") ? 0 : (a[i + 1] == "y") ? 1 : 2;
          f.amplitude = atof(a[i + 2].c_str()); f.period = atof(a[i + 3].c_str()); f.wiggle = 1; i += 4;
        } else if (a[i] == "shear") {
          if (i + 3 > a.size()) fatal("Illegal fix wall/granFix command");
          f.axis = (a[i + 1] == "x") ? 0 : (a[i + 1] == "y") ? 1 : 2;
          f.vshear = atof(a[i + 2].c_str()); f.wshear = 1; i += 3;
        } else fatal("Illegal fix wall/granFix command");
      }
      if (f.wallstyle <= ZPLANE && cfg.periodic[f.wallstyle]) fatal("Cannot use wall in periodic dimension");
      if (f.wallstyle == ZCYLINDER && (cfg.periodic[0] || cfg.periodic[1])) fatal("Cannot use wall in periodic dimension");
      if (f.wiggle && f.wshear) fatal("Cannot wiggle and shear fix wall/granFix");
      if (f.wiggle && f.wallstyle == ZCYLINDER && f.axis != 2) fatal("Invalid wiggle direction for fix wall/granFix");
      if (f.wshear && f.wallstyle <= ZPLANE && f.axis == f.wallstyle) fatal("Invalid shear direction for fix wall/granFix");
      f.time_origin = cfg.ntimestep;  // fix_wall_granFix.cpp:167
      f.wall_index = cfg.nwalls++;
    } else if (st == "freeze") {
      f.kind = FIX_FREEZE;
      cfg.freeze_group_bit = f.groupbit;
    } else fatal("Unknown fix style", st.c_str());
    cfg.fixes.push_back(f);
  }

  void apply_pair_style(const std::vector<std::string> &a) {
    if (a.size() < 2) fatal("Illegal pair_style command");
    cfg.pair = PAIR_NONE; cfg.lub.enabled = 0;
    size_t i = 1;
    bool hybrid = (a[1] == "hybrid/overlay" || a[1] == "hybrid");
    if (hybrid) i = 2;
    while (i < a.size()) {
      const std::string &st = a[i];
      if (st == "none") { i++; }
      else if (st == "gran/hertzFix/history" || st == "gran/hooke/history" || st == "gran/hooke") {
        cfg.pair = (st == "gran/hertzFix/history") ? PAIR_HERTZFIX_HISTORY
                 : (st == "gran/hooke/history") ? PAIR_HOOKE_HISTORY : PAIR_HOOKE;
        cfg.gran = parse_gran(a, i + 1, "pair_style");
        i += 7;
      } else if (st == "lubricate/poly") {
        // pair_style lubricate/poly mu flaglog flagfld cutinner cutoff [flagHI flagVF]  (Appendix A10)
        if (a.size() < i + 6) fatal("Illegal pair_style lubricate/poly command");
        cfg.lub.enabled = 1;
        cfg.lub.mu = atof(a[i + 1].c_str()); cfg.lub.flaglog = atoi(a[i + 2].c_str());
        cfg.lub.flagfld = atoi(a[i + 3].c_str()); cfg.lub.cut_inner = atof(a[i + 4].c_str());
        cfg.lub.cut_global = atof(a[i + 5].c_str());
        cfg.lub.flagHI = cfg.lub.flagVF = 1;
        i += 6;
        if (hybrid) { /* optional flags cannot be told apart from the next sub-style name unless numeric */ }
        if (i + 1 < a.size() && isdigit((unsigned char)a[i][0]) && isdigit((unsigned char)a[i + 1][0])) {
          cfg.lub.flagHI = atoi(a[i].c_str()); cfg.lub.flagVF = atoi(a[i + 1].c_str()); i += 2;
        }
      } else fatal("Unknown pair style", st.c_str());
      if (!hybrid) break;
    }
  }

  void apply_dump(const std::vector<std::string> &a) {
    if (a.size() < 7) fatal("Illegal dump command");
    if (a[3] != "custom") fatal("Only `dump ID group custom N file fields...` is supported");
    DumpSpec d;
    d.id = a[1];
    d.groupbit = cfg.find_group(a[2]);
    if (!d.groupbit) fatal("dump: unknown group", a[2].c_str());
    d.every = atoll(a[4].c_str());
    if (d.every <= 0) fatal("Illegal dump command: N must be positive");
    d.path = a[5];
    static const struct { const char *name; int f; } names[] = {
        {"id", DF_ID}, {"tag", DF_ID}, {"type", DF_TYPE}, {"diameter", DF_DIAMETER}, {"radius", DF_RADIUS}, {"mass", DF_MASS},
        {"x", DF_X}, {"y", DF_Y}, {"z", DF_Z}, {"vx", DF_VX}, {"vy", DF_VY}, {"vz", DF_VZ}, {"fx", DF_FX}, {"fy", DF_FY}, {"fz", DF_FZ},
        {"omegax", DF_OMEGAX}, {"omegay", DF_OMEGAY}, {"omegaz", DF_OMEGAZ}, {"tqx", DF_TQX}, {"tqy", DF_TQY}, {"tqz", DF_TQZ}};
    for (size_t k = 6; k < a.size(); k++) {
      int f = -1;
      for (size_t m = 0; m < sizeof(names) / sizeof(names[0]); m++) if (a[k] == names[m].name) f = names[m].f;
      if (f < 0) fatal("dump custom: unsupported per-atom field", a[k].c_str());
      d.fields.push_back(f);
      d.columns += a[k] + " ";
    }
    for (size_t i = 0; i < cfg.dumps.size(); i++) if (cfg.dumps[i].id == d.id) fatal("Reuse of dump ID", d.id.c_str());
    cfg.dumps.push_back(d);
  }

  // Executes one script line.  Returns an action the engine must perform itself (run / nothing).
  ScriptAction one(const char *line) {
    ScriptAction act;
    std::vector<std::string> a = tokenize(line);
    if (a.empty()) return act;
    const std::string &c = a[0];
    if (c == "atom_style") { if (a.size() < 2 || a[1] != "sphere") fatal("Only atom_style sphere is supported"); }
    else if (c == "dump") apply_dump(a);
    else if (c == "undump") {
      for (size_t i = 0; i < cfg.dumps.size(); i++) if (a.size() > 1 && a[1] == cfg.dumps[i].id) {
        if (cfg.dumps[i].fp) fclose(cfg.dumps[i].fp);
        cfg.dumps.erase(cfg.dumps.begin() + i); break;
      }
    }
    else if (c == "atom_modify" || c == "communicate" || c == "comm_modify" || c == "thermo" ||
             c == "thermo_style" || c == "thermo_modify" || c == "dimension" || c == "echo" ||
             c == "log" || c == "dump_modify" || c == "compute") {
      // accepted, no effect on the hot path
    }
    else if (c == "neigh_modify") {
      // The engine tests the skin/2 displacement criterion after every sub-step and rebuilds at once: LAMMPS'
      // `delay 0 every 1 check yes` (every shipped in.lammps, e.g. xiaocase1/in.lammps:13).  Any other setting would
      // change when the lists are rebuilt and with it the pair sets and the history carry: refused, not ignored.
      for (size_t k = 1; k + 1 < a.size(); k += 2) {
        if (a[k] == "delay") { if (atoi(a[k + 1].c_str()) != 0) fatal("neigh_modify delay: only `delay 0` is supported (rebuild check after every step)"); }
        else if (a[k] == "every") { if (atoi(a[k + 1].c_str()) != 1) fatal("neigh_modify every: only `every 1` is supported"); }
        else if (a[k] == "check") { if (a[k + 1] != "yes") fatal("neigh_modify check: only `check yes` is supported"); }
        else if (a[k] == "one" || a[k] == "page" || a[k] == "binsize") { /* memory / bin tuning: no effect on the pair set */ }
        else fatal("neigh_modify keyword not supported:", a[k].c_str());
      }
      cfg.neigh_modify_seen = 1;
    }
    else if (c == "write_restart") { if (a.size() < 2) fatal("Illegal write_restart command"); act.kind = ScriptAction::WRITE_RESTART; act.path = a[1]; }
    else if (c == "read_restart") { if (a.size() < 2) fatal("Illegal read_restart command"); act.kind = ScriptAction::READ_RESTART; act.path = a[1]; }
    else if (c == "restart") {   // restart 0 | restart N file | restart N file1 file2
      if (a.size() < 2) fatal("Illegal restart command");
      act.kind = ScriptAction::RESTART_EVERY; act.nsteps = atoll(a[1].c_str());
      if (act.nsteps > 0 && a.size() < 3) fatal("Illegal restart command");
      if (a.size() > 2) act.path = a[2];
      if (a.size() > 3) act.path2 = a[3];
    }
    else if (c == "units") { if (a.size() > 1 && a[1] != "lj") fatal("Only lj units are supported (reference inputs set none)"); }
    else if (c == "boundary") {
      if (a.size() != 4) fatal("Illegal boundary command");
      for (int d = 0; d < 3; d++) { cfg.periodic[d] = (a[d + 1][0] == 'p'); cfg.boundary_str[d] = a[d + 1].size() == 1 ? a[d + 1] + a[d + 1] : a[d + 1]; }
    }
    else if (c == "newton") { cfg.newton_pair = (a.size() > 1 && a[1] == "on"); }
    else if (c == "processors") {
      if (a.size() < 4) fatal("Illegal processors command");
      for (int d = 0; d < 3; d++) cfg.procgrid[d] = (a[d + 1] == "*") ? 0 : atoi(a[d + 1].c_str());
    }
    else if (c == "read_data") { if (a.size() < 2) fatal("Illegal read_data command"); read_data(a[1]); act.kind = ScriptAction::READ_DATA; act.path = a[1]; }
    else if (c == "region") {
      // region ID block xlo xhi ylo yhi zlo zhi -- remembered for create_box
      if (a.size() >= 9 && a[2] == "block") { for (int d = 0; d < 3; d++) { region_lo[d] = atof(a[3 + 2 * d].c_str()); region_hi[d] = atof(a[4 + 2 * d].c_str()); } }
    }
    else if (c == "create_box") {
      if (a.size() < 3) fatal("Illegal create_box command");
      cfg.ntypes = atoi(a[1].c_str());
      for (int d = 0; d < 3; d++) { cfg.boxlo[d] = region_lo[d]; cfg.boxhi[d] = region_hi[d]; }
      cfg.have_box = 1;
    }
    else if (c == "neighbor") { if (a.size() < 3 || a[2] != "bin") fatal("Only `neighbor <skin> bin` is supported"); cfg.skin = atof(a[1].c_str()); }
    else if (c == "pair_style") apply_pair_style(a);
    else if (c == "pair_coeff") { /* granular styles take `* *`; lubricate cut-offs stay global */ }
    else if (c == "timestep") { if (a.size() < 2) fatal("Illegal timestep command"); cfg.dt = atof(a[1].c_str()); }
    else if (c == "velocity") {
      if (a.size() < 6 || a[2] != "set") fatal("Only `velocity <group> set vx vy vz` is supported");
      int bit = cfg.find_group(a[1]);
      for (size_t i = 0; i < atoms.size(); i++) if (mask[i] & bit)
        for (int d = 0; d < 3; d++) if (a[3 + d] != "NULL") atoms.v[3 * i + d] = atof(a[3 + d].c_str());
    }
    else if (c == "group") apply_group(a);
    else if (c == "fix") apply_fix(a);
    else if (c == "unfix") {
      for (size_t i = 0; i < cfg.fixes.size(); i++) if (a.size() > 1 && a[1] == cfg.fixes[i].id) { cfg.fixes.erase(cfg.fixes.begin() + i); break; }
    }
    else if (c == "run") { if (a.size() < 2) fatal("Illegal run command"); act.kind = ScriptAction::RUN; act.nsteps = atoll(a[1].c_str()); }
    else fatal("Unknown command", c.c_str());
    return act;
  }

  Script() { for (int d = 0; d < 3; d++) { region_lo[d] = 0; region_hi[d] = 1; } }

 private:
  double region_lo[3], region_hi[3];
};

}  // namespace sedi
