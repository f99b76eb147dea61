// oracle/ref_backend.cpp -- TEST INFRASTRUCTURE ONLY.  Built only where /root/reference exists, into
// oracle/_ref/libsedi_ref.so (git-ignored; travels to the GPU box as a prebuilt file).
//
// Drives the reference's OWN, UNMODIFIED plug-in sources -- compiled by path from
// /root/reference/interfaceToLammps/{pair_gran_hertzFix_history,fix_fluid_drag,fix_cohesive,fix_wall_granFix,
// pair_lubricate_poly}.cpp against the stub headers in oracle/stubs/ -- on the arrays of the oracle driver.
// No reference source is copied into this repository; this file only instantiates the reference classes.
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <unistd.h>
#include <vector>

#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "fix_cohesive.h"
#include "fix_fluid_drag.h"
#include "fix_wall_granFix.h"
#include "force.h"
#include "group.h"
#include "input.h"
#include "memory.h"
#include "modify.h"
#include "neigh_list.h"
#include "neighbor.h"
#include "pair_gran_hertzFix_history.h"
#include "pair_lubricate_poly.h"
#include "update.h"
#include "variable.h"

#include "oracle_backend.hpp"

namespace ora {

namespace {

struct Args {  // builds a char** argv from tokens
  std::vector<std::string> s;
  std::vector<char *> p;
  void add(const std::string &t) { s.push_back(t); }
  void add(double v) { char b[64]; snprintf(b, sizeof(b), "%.17g", v); s.push_back(b); }
  void addi(long long v) { char b[64]; snprintf(b, sizeof(b), "%lld", v); s.push_back(b); }
  char **argv() { p.clear(); for (size_t i = 0; i < s.size(); i++) p.push_back((char *)s[i].c_str()); return p.data(); }
  int argc() const { return (int)s.size(); }
};

// protected-member access for the wall fix (its shear array is protected, fix_wall_granFix.h:52-54)
class WallFixAccess : public LAMMPS_NS::FixWallGranFix {
 public:
  WallFixAccess(LAMMPS_NS::LAMMPS *l, int n, char **a) : FixWallGranFix(l, n, a) {}
  void set_shear(double **s) { LAMMPS_NS::Memory m; if (shear && shear != s && owns) { m.destroy(shear); } owns = false; shear = s; }
  void set_time_origin(int t) { time_origin = t; }
  bool owns = true;
};

void add_gran_args(Args &a, const sedi::GranParams &g) {
  // kt/gammat are passed explicitly (the NULL defaults were already resolved by the script parser with the
  // same rules); dampflag=1 so that the reference keeps the passed gammat instead of zeroing it again.
  a.add(g.kn); a.add(g.kt); a.add(g.gamman); a.add(g.gammat); a.add(g.xmu); a.addi(1);
}

}  // namespace

class RefBackend : public Backend {
 public:
  LAMMPS_NS::LAMMPS lmp;
  LAMMPS_NS::Atom atom;
  LAMMPS_NS::Update update;
  LAMMPS_NS::Force force;
  LAMMPS_NS::Neighbor neighbor;
  LAMMPS_NS::Comm comm;
  LAMMPS_NS::Memory memory;
  LAMMPS_NS::Error error;
  LAMMPS_NS::Domain domain;
  LAMMPS_NS::Modify modify;
  LAMMPS_NS::Input input;
  LAMMPS_NS::Variable variable;
  LAMMPS_NS::Group group;
  sedi::SimConfig cfg;
  LAMMPS_NS::PairGranHertzFixHistory *hertz;
  LAMMPS_NS::PairLubricatePoly *lub;
  std::vector<LAMMPS_NS::Fix *> fixes;  // parallel to cfg.fixes (NULL for styles not in the reference tree)
  Backend *port;                        // EXTERNAL pieces (stock gran/hooke/history pair) fall back to the port
  LAMMPS_NS::NeighList nl, nlh;
  double *cutsq_rows[64], *cutin_rows[64];
  std::vector<double> cutsq_data, cutin_data;

  RefBackend() : hertz(0), lub(0), port(make_port_backend()) {
    lmp.atom = &atom; lmp.update = &update; lmp.force = &force; lmp.neighbor = &neighbor; lmp.comm = &comm;
    lmp.memory = &memory; lmp.error = &error; lmp.domain = &domain; lmp.modify = &modify; lmp.input = &input;
    lmp.group = &group; input.variable = &variable;
  }

  const char *name() const { return "reference"; }

  void bind(const AtomView &av, const StepInfo &st) {
    atom.nlocal = av.nlocal; atom.nghost = av.nghost; atom.nmax = av.nlocal + av.nghost;
    atom.natoms = av.nlocal;
    atom.x = av.x; atom.v = av.v; atom.f = av.f; atom.omega = av.omega; atom.torque = av.torque;
    atom.radius = av.radius; atom.rmass = av.rmass; atom.type = av.type; atom.mask = av.mask; atom.tag = av.tag;
    update.dt = st.dt; update.ntimestep = st.ntimestep; update.setupflag = st.setupflag;
  }

  void bind_list(LAMMPS_NS::NeighList &l, LAMMPS_NS::NeighList *h, const NList &in) {
    l.inum = in.inum; l.ilist = in.ilist; l.numneigh = in.numneigh; l.firstneigh = in.firstneigh;
    if (h) { h->firstneigh = in.firsttouch; h->firstdouble = in.firstshear; l.listgranhistory = h; }
  }

  void init(const sedi::SimConfig &c, const AtomView &av, const BoxInfo &box, const StepInfo &st) {
    cfg = c;
    port->init(c, av, box, st);
    bind(av, st);
    atom.nmax = av.nlocal + av.nghost + 1;  // fix constructors allocate their own per-atom arrays; re-pointed at driver storage later
    force.newton_pair = cfg.newton_pair;
    for (int d = 0; d < 3; d++) { domain.boxlo[d] = box.lo[d]; domain.boxhi[d] = box.hi[d]; domain.prd[d] = box.hi[d] - box.lo[d]; }
    domain.xprd = domain.prd[0]; domain.yprd = domain.prd[1]; domain.zprd = domain.prd[2];
    domain.xperiodic = box.periodic[0]; domain.yperiodic = box.periodic[1]; domain.zperiodic = box.periodic[2];
    force.pair_style = cfg.pair == sedi::PAIR_HERTZFIX_HISTORY ? "gran/hertzFix/history"
                     : cfg.pair == sedi::PAIR_HOOKE_HISTORY ? "gran/hooke/history"
                     : cfg.pair == sedi::PAIR_HOOKE ? "gran/hooke" : "none";
    if (cfg.pair == sedi::PAIR_HERTZFIX_HISTORY) {
      hertz = new LAMMPS_NS::PairGranHertzFixHistory(&lmp);
      Args a; add_gran_args(a, cfg.gran);
      hertz->settings(a.argc(), a.argv());         // pair_gran_hertzFix_history.cpp:293-317
      hertz->gammat = cfg.gran.gammat;
      hertz->dt = st.dt_init;                      // EXTERNAL init_style: dt = update->dt
      hertz->freeze_group_bit = cfg.freeze_group_bit;
    }
    // fixes first (lubricate's init_style scans modify->fix for "deform"/"wall" styles)
    fixes.assign(cfg.fixes.size(), (LAMMPS_NS::Fix *)0);
    for (size_t k = 0; k < cfg.fixes.size(); k++) {
      const sedi::FixSpec &fx = cfg.fixes[k];
      Args a; a.add(fx.id); a.add("all");
      if (fx.kind == sedi::FIX_FDRAG) {
        a.add("fdrag"); a.addi((long long)fx.carrier_rho);
        fixes[k] = new LAMMPS_NS::FixFluidDrag(&lmp, a.argc(), a.argv());
      } else if (fx.kind == sedi::FIX_COHESIVE) {
        a.add("cohesive"); a.add(fx.ah); a.add(fx.lam); a.add(fx.smin); a.add(fx.smax); a.addi(fx.opt);
        int out = dup(1); fflush(stdout); int nul = open("/dev/null", O_WRONLY); dup2(nul, 1);   // ctor printf
        fixes[k] = new LAMMPS_NS::FixCohe(&lmp, a.argc(), a.argv());
        fflush(stdout); dup2(out, 1); close(out); close(nul);
      } else if (fx.kind == sedi::FIX_WALL_GRAN) {
        a.add("wall/granFix"); add_gran_args(a, fx.wall);
        const char *ws[4] = {"xplane", "yplane", "zplane", "zcylinder"};
        a.add(ws[fx.wallstyle]);
        if (fx.wallstyle == sedi::ZCYLINDER) a.add(fx.cylradius);
        else {
          if (fx.lo <= -sedi::WALL_BIG) a.add("NULL"); else a.add(fx.lo);
          if (fx.hi >= sedi::WALL_BIG) a.add("NULL"); else a.add(fx.hi);
        }
        const char *ax[3] = {"x", "y", "z"};
        if (fx.wiggle) { a.add("wiggle"); a.add(ax[fx.axis]); a.add(fx.amplitude); a.add(fx.period); }
        if (fx.wshear) { a.add("shear"); a.add(ax[fx.axis]); a.add(fx.vshear); }
        // the ctor rejects walls in periodic dimensions through domain->*periodic, which is bound above
        WallFixAccess *w = new WallFixAccess(&lmp, a.argc(), a.argv());
        w->set_time_origin((int)fx.time_origin);
        update.dt = st.dt_init;
        w->init();                                  // reads dt and the pair style, fix_wall_granFix.cpp:212-231
        update.dt = st.dt;
        fixes[k] = w;
      }
      if (fixes[k]) fixes[k]->groupbit = fx.groupbit;
    }
    modify.nfix = 0; modify.fix = 0;  // granular walls are not stock FixWall objects: keep lubricate off that path
    if (cfg.lub.enabled) {
      lub = new LAMMPS_NS::PairLubricatePoly(&lmp);
      lub->mu = cfg.lub.mu; lub->flaglog = cfg.lub.flaglog; lub->flagfld = cfg.lub.flagfld;
      lub->cut_inner_global = cfg.lub.cut_inner; lub->cut_global = cfg.lub.cut_global;
      lub->flagHI = cfg.lub.flagHI; lub->flagVF = cfg.lub.flagVF;
      const int nt = cfg.ntypes + 1;
      cutsq_data.assign((size_t)nt * nt, cfg.lub.cut_global * cfg.lub.cut_global);
      cutin_data.assign((size_t)nt * nt, cfg.lub.cut_inner);
      for (int t = 0; t < nt && t < 64; t++) { cutsq_rows[t] = &cutsq_data[(size_t)t * nt]; cutin_rows[t] = &cutin_data[(size_t)t * nt]; }
      lub->cutsq = cutsq_rows; lub->cut_inner = cutin_rows;
      lub->init_style();                           // pair_lubricate_poly.cpp:450-577 (R0, RT0, RS0, Ef = 0)
    }
  }

  void pair_granular(const AtomView &av, const NList &list, const StepInfo &st) {
    if (cfg.pair != sedi::PAIR_HERTZFIX_HISTORY) { port->pair_granular(av, list, st); return; }  // EXTERNAL stock styles
    bind(av, st);
    bind_list(nl, &nlh, list);
    hertz->list = &nl;
    hertz->compute(0, 0);
  }

  void pair_lubricate(const AtomView &av, const NList &full, const StepInfo &st) {
    bind(av, st);
    bind_list(nl, 0, full);
    lub->list = &nl;
    lub->compute(0, 0);
  }

  void fix_fdrag(int ifix, const AtomView &av, const FdragState &fs, const StepInfo &st) {
    bind(av, st);
    LAMMPS_NS::FixFluidDrag *fx = (LAMMPS_NS::FixFluidDrag *)fixes[ifix];
    fx->ffluiddrag = fs.ffluiddrag; fx->DuDt = fs.DuDt; fx->vOld = fs.vOld; fx->foamCpuId = fs.foamCpuId;
    fx->post_force(0);
  }

  void fix_cohesive(int ifix, const AtomView &av, const NList &half, const StepInfo &st) {
    bind(av, st);
    bind_list(nl, 0, half);
    LAMMPS_NS::FixCohe *fx = (LAMMPS_NS::FixCohe *)fixes[ifix];
    fx->init_list(0, &nl);
    const bool noisy = (cfg.fixes[ifix].opt == 0);   // post_force printf(" inum %i") every call, fix_cohesive.cpp:163
    int out = -1, nul = -1;
    if (noisy) { fflush(stdout); out = dup(1); nul = open("/dev/null", O_WRONLY); dup2(nul, 1); }
    fx->post_force(0);
    if (noisy) { fflush(stdout); dup2(out, 1); close(out); close(nul); }
  }

  void fix_wall(int ifix, const AtomView &av, double **shear, const StepInfo &st) {
    bind(av, st);
    WallFixAccess *w = (WallFixAccess *)fixes[ifix];
    w->set_shear(shear);
    w->post_force(0);
  }
};

Backend *make_ref_backend() { return new RefBackend(); }

}  // namespace ora
