// oracle/oracle_driver.hpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement of the EXTERNAL LAMMPS (lammps-1Feb14, not vendored in /root/reference) machinery that the
// reference plug-ins run inside: Verlet step order, fix nve/sphere, fix gravity, fix freeze, binned neighbour
// lists with `newton off`, ghost atoms for periodic images, and FixShearHistory's carry-over of contact
// history across rebuilds.  Follows SURVEY.md Appendix A1-A10 (the published behaviour of that release); the
// reference's own call sites that fix these semantics are cited inline.  Anchors: the script commands in
// cases/**/in.lammps, `lammps_step` = "run n pre no post no" (interfaceToLammps/library.cpp:372-386), and the
// Hooke golden dumps (tests/golden/) which exercise this whole loop.
#pragma once
#include <algorithm>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "oracle_backend.hpp"

namespace ora {

struct Arr2 {  // LAMMPS-style double** over one contiguous block
  std::vector<double> data;
  std::vector<double *> rows;
  void resize(size_t n) {
    data.resize(3 * n, 0.0);
    rows.resize(n ? n : 1);
    for (size_t i = 0; i < n; i++) rows[i] = &data[3 * i];
  }
  double **p() { return rows.data(); }
};

struct CSRList {
  std::vector<int> ilist, numneigh, offset, neigh, touch;
  std::vector<double> shear;
  std::vector<int *> firstneigh, firsttouch;
  std::vector<double *> firstshear;
  bool history;
  CSRList() : history(false) {}
  void finalize(int nlocal) {
    ilist.resize(nlocal); firstneigh.resize(nlocal ? nlocal : 1);
    if (history) {
      if (touch.size() != neigh.size()) touch.assign(neigh.size(), 0);
      if (shear.size() != 3 * neigh.size()) shear.assign(3 * neigh.size(), 0.0);
      firsttouch.resize(nlocal ? nlocal : 1); firstshear.resize(nlocal ? nlocal : 1);
    }
    for (int i = 0; i < nlocal; i++) {
      ilist[i] = i;
      firstneigh[i] = neigh.data() + offset[i];
      if (history) { firsttouch[i] = touch.data() + offset[i]; firstshear[i] = shear.data() + 3 * (size_t)offset[i]; }
    }
  }
  NList view() {
    NList l;
    l.inum = (int)ilist.size(); l.ilist = ilist.data(); l.numneigh = numneigh.data(); l.firstneigh = firstneigh.data();
    l.firsttouch = history ? firsttouch.data() : 0; l.firstshear = history ? firstshear.data() : 0;
    return l;
  }
};

class Sim {
 public:
  sedi::Script script;
  Backend *be;
  int nlocal, nghost;
  Arr2 x, v, f, omega, torque, xhold;
  std::vector<double> radius, rmass;
  std::vector<int> type, mask, tag;
  // ghosts: periodic images.  ghost k (local index nlocal+k) copies atom gsrc[k] shifted by gshift[3k..]
  std::vector<int> gsrc;
  std::vector<double> gshift;
  CSRList gran, half, full;
  // FixShearHistory per-atom partner store (Appendix A7)
  std::vector<std::vector<int> > partner;
  std::vector<std::vector<double> > shearpartner;
  // per-atom fix state
  Arr2 ffluiddrag, DuDt, vOld;
  std::vector<int> foamCpuId;
  std::vector<Arr2> wallshear;
  bool setup_done;
  double dt_init;
  long long nbuilds, npair_evals, nsteps_done;
  double cutneighmax;
  std::vector<double> cuttype;  // (ntypes+1)^2 force cut-offs per type pair

  explicit Sim(Backend *b) : be(b), nlocal(0), nghost(0), setup_done(false), dt_init(0), nbuilds(0), npair_evals(0),
                             nsteps_done(0), cutneighmax(0) {}
  ~Sim() { delete be; }

  sedi::SimConfig &cfg() { return script.cfg; }

  bool need_half() { for (size_t k = 0; k < cfg().fixes.size(); k++) if (cfg().fixes[k].kind == sedi::FIX_COHESIVE) return true; return false; }
  bool need_full() { return cfg().lub.enabled != 0; }
  bool need_gran() { return cfg().pair != sedi::PAIR_NONE; }

  void alloc(size_t nall) {
    x.resize(nall); v.resize(nall); f.resize(nall); omega.resize(nall); torque.resize(nall);
    radius.resize(nall); rmass.resize(nall); type.resize(nall); mask.resize(nall); tag.resize(nall);
  }

  // copy the script's atoms into the working arrays (done once, at the first run)
  void load_atoms() {
    const sedi::AtomData &a = script.atoms;
    nlocal = (int)a.size(); nghost = 0;
    alloc(nlocal);
    for (int i = 0; i < nlocal; i++) {
      for (int d = 0; d < 3; d++) { x.rows[i][d] = a.x[3 * i + d]; v.rows[i][d] = a.v[3 * i + d]; omega.rows[i][d] = a.omega[3 * i + d]; }
      radius[i] = a.radius[i]; rmass[i] = a.rmass[i]; type[i] = a.type[i]; tag[i] = a.tag[i]; mask[i] = script.mask[i];
    }
    xhold.resize(nlocal);
    ffluiddrag.resize(nlocal); DuDt.resize(nlocal); vOld.resize(nlocal); foamCpuId.assign(nlocal, 0);
    wallshear.resize(cfg().nwalls);
    for (int w = 0; w < cfg().nwalls; w++) wallshear[w].resize(nlocal);
    partner.assign(nlocal, std::vector<int>()); shearpartner.assign(nlocal, std::vector<double>());
  }

  AtomView view() {
    AtomView a;
    a.nlocal = nlocal; a.nghost = nghost;
    a.x = x.p(); a.v = v.p(); a.f = f.p(); a.omega = omega.p(); a.torque = torque.p();
    a.radius = radius.data(); a.rmass = rmass.data(); a.type = type.data(); a.mask = mask.data(); a.tag = tag.data();
    return a;
  }

  StepInfo stepinfo(int setupflag) {
    StepInfo s; s.dt = cfg().dt; s.dt_init = dt_init; s.ntimestep = cfg().ntimestep; s.setupflag = setupflag; return s;
  }

  // ---- cut-offs (EXTERNAL PairGranHookeHistory::init_style/init_one; PairLubricate::init_one; Neighbor::init)
  void compute_cutoffs() {
    const int nt = cfg().ntypes;
    std::vector<double> maxdyn(nt + 1, 0.0), maxfrz(nt + 1, 0.0);
    for (int i = 0; i < nlocal; i++) {
      if (mask[i] & cfg().freeze_group_bit) maxfrz[type[i]] = std::max(maxfrz[type[i]], radius[i]);
      else maxdyn[type[i]] = std::max(maxdyn[type[i]], radius[i]);
    }
    cuttype.assign((size_t)(nt + 1) * (nt + 1), 0.0);
    double cutmax = 0.0;
    for (int a = 1; a <= nt; a++)
      for (int b = 1; b <= nt; b++) {
        double c = 0.0;
        if (need_gran()) {
          c = maxdyn[a] + maxdyn[b];
          c = std::max(c, maxfrz[a] + maxdyn[b]);
          c = std::max(c, maxdyn[a] + maxfrz[b]);
        }
        if (cfg().lub.enabled) c = std::max(c, cfg().lub.cut_global);
        cuttype[(size_t)a * (nt + 1) + b] = c;
        cutmax = std::max(cutmax, c);
      }
    cutneighmax = cutmax + cfg().skin;
  }
  double cutneighsq(int ta, int tb) {
    double c = cuttype[(size_t)ta * (cfg().ntypes + 1) + tb] + cfg().skin;
    return c * c;
  }

  // ---- periodic wrap of owned atoms at reneighbouring (EXTERNAL Domain::pbc)
  void pbc() {
    for (int d = 0; d < 3; d++) {
      if (!cfg().periodic[d]) continue;
      const double lo = cfg().boxlo[d], hi = cfg().boxhi[d], prd = hi - lo;
      for (int i = 0; i < nlocal; i++) {
        double &c = x.rows[i][d];
        if (c < lo) c += prd;
        if (c >= hi) { c -= prd; c = std::max(c, lo); }
      }
    }
  }

  // ---- ghost atoms = periodic images within cutneighmax of a periodic face (EXTERNAL Comm::borders, single rank)
  void build_ghosts() {
    gsrc.clear(); gshift.clear();
    std::vector<double> gx;  // positions of ghosts created so far
    for (int d = 0; d < 3; d++) {
      if (!cfg().periodic[d]) continue;
      const double lo = cfg().boxlo[d], hi = cfg().boxhi[d], prd = hi - lo;
      const int ncur = nlocal + (int)gsrc.size();
      for (int k = 0; k < ncur; k++) {
        const double *p = (k < nlocal) ? x.rows[k] : &gx[3 * (size_t)(k - nlocal)];
        const double c = p[d];
        for (int side = 0; side < 2; side++) {
          bool take = side == 0 ? (c < lo + cutneighmax) : (c >= hi - cutneighmax);
          if (!take) continue;
          double sh[3] = {0, 0, 0};
          sh[d] = side == 0 ? prd : -prd;
          if (k >= nlocal) for (int e = 0; e < 3; e++) sh[e] += gshift[3 * (size_t)(k - nlocal) + e];
          const int root = (k < nlocal) ? k : gsrc[k - nlocal];
          gsrc.push_back(root);
          for (int e = 0; e < 3; e++) { gshift.push_back(sh[e]); gx.push_back(x.rows[root][e] + sh[e]); }
        }
      }
    }
    nghost = (int)gsrc.size();
    // grow arrays, keep owned data
    const size_t nall = (size_t)nlocal + nghost;
    std::vector<double> sx(x.data.begin(), x.data.begin() + 3 * (size_t)nlocal), sv(v.data.begin(), v.data.begin() + 3 * (size_t)nlocal),
        so(omega.data.begin(), omega.data.begin() + 3 * (size_t)nlocal), sf(f.data.begin(), f.data.begin() + 3 * (size_t)nlocal),
        st(torque.data.begin(), torque.data.begin() + 3 * (size_t)nlocal);
    alloc(nall);
    std::copy(sx.begin(), sx.end(), x.data.begin()); std::copy(sv.begin(), sv.end(), v.data.begin());
    std::copy(so.begin(), so.end(), omega.data.begin()); std::copy(sf.begin(), sf.end(), f.data.begin());
    std::copy(st.begin(), st.end(), torque.data.begin());
    for (int k = 0; k < nghost; k++) {
      const int g = nlocal + k, s = gsrc[k];
      radius[g] = radius[s]; rmass[g] = rmass[s]; type[g] = type[s]; mask[g] = mask[s]; tag[g] = tag[s];
    }
    forward_comm();
  }

  // per-step ghost refresh: x (shifted), v, omega -- `communicate single vel yes` (every shipped in.lammps)
  void forward_comm() {
    for (int k = 0; k < nghost; k++) {
      const int g = nlocal + k, s = gsrc[k];
      for (int d = 0; d < 3; d++) {
        x.rows[g][d] = x.rows[s][d] + gshift[3 * (size_t)k + d];
        v.rows[g][d] = v.rows[s][d];
        omega.rows[g][d] = omega.rows[s][d];
      }
    }
  }

  // ---- FixShearHistory::pre_exchange (Appendix A7): save touching pairs' shear keyed by partner tag
  void history_save() {
    if (!gran.history) return;
    for (int i = 0; i < nlocal; i++) { partner[i].clear(); shearpartner[i].clear(); }
    for (int i = 0; i < (int)gran.numneigh.size(); i++) {
      const int off = gran.offset[i];
      for (int jj = 0; jj < gran.numneigh[i]; jj++) {
        if (!gran.touch[off + jj]) continue;
        const int j = gran.neigh[off + jj];
        const double *s = &gran.shear[3 * (size_t)(off + jj)];
        partner[i].push_back(tag[j]);
        for (int d = 0; d < 3; d++) shearpartner[i].push_back(s[d]);
        if (j < nlocal) {
          partner[j].push_back(tag[i]);
          for (int d = 0; d < 3; d++) shearpartner[j].push_back(-s[d]);
        }
      }
    }
  }

  // ---- binned neighbour build, newton off (Appendix A6)
  void build_lists() {
    const int nall = nlocal + nghost;
    double blo[3], bhi[3];
    int nb[3];
    for (int d = 0; d < 3; d++) {
      blo[d] = cfg().boxlo[d] - (cfg().periodic[d] ? cutneighmax : 0.0);
      bhi[d] = cfg().boxhi[d] + (cfg().periodic[d] ? cutneighmax : 0.0);
      nb[d] = (int)floor((bhi[d] - blo[d]) / cutneighmax);
      if (nb[d] < 1) nb[d] = 1;
      if (nb[d] > 1024) nb[d] = 1024;
    }
    std::vector<int> bin(nall), cnt((size_t)nb[0] * nb[1] * nb[2] + 1, 0);
    std::vector<int> bc(3 * (size_t)nall);
    for (int i = 0; i < nall; i++) {
      int c[3];
      for (int d = 0; d < 3; d++) {
        c[d] = (int)floor((x.rows[i][d] - blo[d]) / (bhi[d] - blo[d]) * nb[d]);
        if (c[d] < 0) c[d] = 0;
        if (c[d] >= nb[d]) c[d] = nb[d] - 1;
        bc[3 * (size_t)i + d] = c[d];
      }
      bin[i] = c[0] + nb[0] * (c[1] + nb[1] * c[2]);
      cnt[bin[i] + 1]++;
    }
    for (size_t b = 1; b < cnt.size(); b++) cnt[b] += cnt[b - 1];
    std::vector<int> order(nall), fill(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < nall; i++) order[fill[bin[i]]++] = i;  // ascending local index inside a bin

    const bool hist = (cfg().pair == sedi::PAIR_HERTZFIX_HISTORY || cfg().pair == sedi::PAIR_HOOKE_HISTORY);
    const bool dog = need_gran(), doh = need_half(), dof = need_full();
    gran = CSRList(); half = CSRList(); full = CSRList();
    gran.history = hist && dog;
    gran.numneigh.assign(nlocal, 0); half.numneigh.assign(nlocal, 0); full.numneigh.assign(nlocal, 0);
    gran.offset.assign(nlocal, 0); half.offset.assign(nlocal, 0); full.offset.assign(nlocal, 0);
    std::vector<int> gtouch;
    std::vector<double> gshear;
    const double skin = cfg().skin;
    for (int i = 0; i < nlocal; i++) {
      gran.offset[i] = (int)gran.neigh.size(); half.offset[i] = (int)half.neigh.size(); full.offset[i] = (int)full.neigh.size();
      const double xi = x.rows[i][0], yi = x.rows[i][1], zi = x.rows[i][2], radi = radius[i];
      const int *ci = &bc[3 * (size_t)i];
      for (int bz = std::max(0, ci[2] - 1); bz <= std::min(nb[2] - 1, ci[2] + 1); bz++)
        for (int by = std::max(0, ci[1] - 1); by <= std::min(nb[1] - 1, ci[1] + 1); by++)
          for (int bx = std::max(0, ci[0] - 1); bx <= std::min(nb[0] - 1, ci[0] + 1); bx++) {
            const int b = bx + nb[0] * (by + nb[1] * bz);
            for (int k = cnt[b]; k < cnt[b + 1]; k++) {
              const int j = order[k];
              if (j == i) continue;
              const double delx = xi - x.rows[j][0], dely = yi - x.rows[j][1], delz = zi - x.rows[j][2];
              const double rsq = delx * delx + dely * dely + delz * delz;
              const double cnsq = (doh || dof) ? cutneighsq(type[i], type[j]) : 0.0;
              if (dof && rsq <= cnsq) full.neigh.push_back(j);
              if (j <= i) continue;  // own/own pairs once; own/ghost always (ghost index >= nlocal)
              if (doh && rsq <= cnsq) half.neigh.push_back(j);
              if (dog) {
                const double radsum = radi + radius[j];
                const double cutsq = (radsum + skin) * (radsum + skin);
                if (rsq <= cutsq) {
                  gran.neigh.push_back(j);
                  if (gran.history) {
                    int t = 0; double s[3] = {0, 0, 0};
                    if (rsq < radsum * radsum) {  // re-attach by partner tag
                      const std::vector<int> &pl = partner[i];
                      for (size_t m = 0; m < pl.size(); m++) if (pl[m] == tag[j]) { t = 1; for (int d = 0; d < 3; d++) s[d] = shearpartner[i][3 * m + d]; break; }
                    }
                    gtouch.push_back(t); for (int d = 0; d < 3; d++) gshear.push_back(s[d]);
                  }
                }
              }
            }
          }
      gran.numneigh[i] = (int)gran.neigh.size() - gran.offset[i];
      half.numneigh[i] = (int)half.neigh.size() - half.offset[i];
      full.numneigh[i] = (int)full.neigh.size() - full.offset[i];
    }
    if (gran.history) { gran.touch.swap(gtouch); gran.shear.swap(gshear); }
    gran.finalize(nlocal); half.finalize(nlocal); full.finalize(nlocal);
    for (int i = 0; i < nlocal; i++) for (int d = 0; d < 3; d++) xhold.rows[i][d] = x.rows[i][d];
    nbuilds++;
  }

  void reneighbor() {
    history_save();
    pbc();
    build_ghosts();
    build_lists();
  }

  // Neighbor::check_distance with `neigh_modify delay 0` (every 1 check yes): any owned atom moved > skin/2
  bool check_distance() {
    const double trig = 0.25 * cfg().skin * cfg().skin;
    for (int i = 0; i < nlocal; i++) {
      const double dx = x.rows[i][0] - xhold.rows[i][0], dy = x.rows[i][1] - xhold.rows[i][1], dz = x.rows[i][2] - xhold.rows[i][2];
      if (dx * dx + dy * dy + dz * dz > trig) return true;
    }
    return false;
  }

  // ---- force evaluation for the current positions (pair styles, then post_force fixes in script order)
  void compute_forces(int setupflag) {
    for (int i = 0; i < nlocal; i++) for (int d = 0; d < 3; d++) { f.rows[i][d] = 0.0; torque.rows[i][d] = 0.0; }
    AtomView av = view();
    StepInfo st = stepinfo(setupflag);
    if (need_gran()) { be->pair_granular(av, gran.view(), st); for (int i = 0; i < nlocal; i++) npair_evals += gran.numneigh[i]; }
    if (need_full()) be->pair_lubricate(av, full.view(), st);
    for (size_t k = 0; k < cfg().fixes.size(); k++) {
      const sedi::FixSpec &fx = cfg().fixes[k];
      switch (fx.kind) {
        case sedi::FIX_GRAVITY:  // EXTERNAL FixGravity::post_force: f += rmass * g * nhat
          for (int i = 0; i < nlocal; i++) if (mask[i] & fx.groupbit) {
            const double m = rmass[i];
            f.rows[i][0] += m * (fx.g * fx.gdir[0]); f.rows[i][1] += m * (fx.g * fx.gdir[1]); f.rows[i][2] += m * (fx.g * fx.gdir[2]);
          }
          break;
        case sedi::FIX_FDRAG: {
          FdragState fs; fs.ffluiddrag = ffluiddrag.p(); fs.DuDt = DuDt.p(); fs.vOld = vOld.p(); fs.foamCpuId = foamCpuId.data();
          be->fix_fdrag((int)k, av, fs, st);
          break;
        }
        case sedi::FIX_COHESIVE:
          // FixCohe declares setup() without the int argument (fix_cohesive.h:33), so LAMMPS' setup(vflag) never
          // reaches post_force: the cohesive force is absent from the setup evaluation.
          if (!setupflag) be->fix_cohesive((int)k, av, half.view(), st);
          break;
        case sedi::FIX_WALL_GRAN:
          be->fix_wall((int)k, av, wallshear[fx.wall_index].p(), st);
          break;
        case sedi::FIX_FREEZE:  // EXTERNAL FixFreeze::post_force
          for (int i = 0; i < nlocal; i++) if (mask[i] & fx.groupbit) for (int d = 0; d < 3; d++) { f.rows[i][d] = 0.0; torque.rows[i][d] = 0.0; }
          break;
        default: break;
      }
    }
  }

  // ---- EXTERNAL FixNVESphere (Appendix A4)
  void integrate(bool initial) {
    const double dtv = dt_init, dtf = 0.5 * dt_init, dtfrotate = dtf / 0.4;
    for (size_t k = 0; k < cfg().fixes.size(); k++) {
      const sedi::FixSpec &fx = cfg().fixes[k];
      if (fx.kind != sedi::FIX_NVE_SPHERE) continue;
      for (int i = 0; i < nlocal; i++) {
        if (!(mask[i] & fx.groupbit)) continue;
        const double dtfm = dtf / rmass[i];
        for (int d = 0; d < 3; d++) v.rows[i][d] += dtfm * f.rows[i][d];
        if (initial) for (int d = 0; d < 3; d++) x.rows[i][d] += dtv * v.rows[i][d];
        const double dtirotate = dtfrotate / (radius[i] * radius[i] * rmass[i]);
        for (int d = 0; d < 3; d++) omega.rows[i][d] += dtirotate * torque.rows[i][d];
      }
    }
  }

  // ---- Verlet::setup (first run of the session, even with `pre no`; softParticleCloud.C:189 lammps_step(0))
  void setup() {
    if (!nlocal && script.atoms.size()) load_atoms();
    dt_init = cfg().dt;
    compute_cutoffs();
    BoxInfo box;
    for (int d = 0; d < 3; d++) { box.lo[d] = cfg().boxlo[d]; box.hi[d] = cfg().boxhi[d]; box.periodic[d] = cfg().periodic[d]; }
    be->init(cfg(), view(), box, stepinfo(1));
    pbc(); build_ghosts(); build_lists();
    compute_forces(1);
    setup_done = true;
  }

  // ---- Verlet::run (Appendix A3)
  void run(long long n) {
    if (!setup_done) setup();
    for (long long s = 0; s < n; s++) {
      cfg().ntimestep++;
      integrate(true);
      if (check_distance()) reneighbor(); else forward_comm();
      compute_forces(0);
      integrate(false);
      nsteps_done++;
    }
  }

  void command(const char *line) {
    sedi::ScriptAction a = script.one(line);
    if (a.kind == sedi::ScriptAction::RUN) run(a.nsteps);
  }
};

}  // namespace ora
