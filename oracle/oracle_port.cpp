// oracle/oracle_port.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement ("port") of the arithmetic of sediFoam's LAMMPS-side plug-ins, written from the reference
// sources and citing the lines followed.  Operation ORDER follows the reference expression by expression so
// that, compiled with the same g++ -O2 -ffp-contract=off, results are bit-identical to the reference objects
// in oracle/_ref (tests/test_oracle_pinning.py asserts exactly that).
//
// Parity status: pinned against (1) the reference's own unmodified sources stub-compiled into
// oracle/_ref/libsedi_ref.so (bit-exact), (2) the reference's Hooke golden dumps
// cases/auto-testing/test-cases/multiParticlesCollide{Dia,Rho}/data/origin/p*.dat (tests/golden/).
#include <math.h>
#include <stdio.h>
#include <vector>
#include "oracle_backend.hpp"

namespace ora {

using sedi::FixSpec;
using sedi::GranParams;
using sedi::SimConfig;

namespace {

struct Vec3 { double a, b, c; };

// State that one sphere-sphere or sphere-wall contact evaluation needs.  The pair and wall laws of the
// reference are the same code with different "geometry inputs"; we evaluate both through contact_force().
struct ContactIn {
  double dx, dy, dz, rsq;  // branch vector from partner (or wall) to particle i
  double vr[3];            // relative translational velocity v_i - v_j (or v_i - v_wall)
  double wsum[3];          // radi*omega_i + radj*omega_j (wall: radius*omega_i) -- NOT yet divided by r
  double meff;
  double overlap_scale;    // pair: radi*radj/radsum applied as (radsum-r)*radi*radj/radsum ; see hz_arg()
  double rcontact;         // pair: radsum ; wall: radius
  double radi, radj;       // pair only
  bool wall;
  bool vn_divide;          // wall hertz_history divides by rsq instead of multiplying by 1/rsq
};

struct ContactOut {
  double fx, fy, fz;     // total force on i
  double t1, t2, t3;     // rinv * (d x fs): caller scales by -radius
};

// beta of the "Fix" damping model: gamman is a restitution coefficient e; beta = -ln e / sqrt(ln^2 e + pi^2).
// Reference writes log(gamman)/log(exp(1.0)) three times (pair_gran_hertzFix_history.cpp:195-196,
// fix_wall_granFix.cpp:602-603).  The expression is loop invariant; we evaluate the identical expression once.
double fix_beta(double gamman) {
  double lg = log(gamman) / log(exp(1.0));
  return -lg / sqrt(lg * lg + sedi::SEDI_MY_PI * sedi::SEDI_MY_PI);
}

// Hertz-Mindlin "Fix" law with shear history.
// pair: pair_gran_hertzFix_history.cpp:142-271 ; wall: fix_wall_granFix.cpp:558-679.
void hertzfix_contact(const ContactIn &c, const GranParams &p, double beta, double dt, bool shearupdate,
                      double *shear, ContactOut &o) {
  const double r = sqrt(c.rsq);
  const double rinv = 1.0 / r;
  const double rsqinv = 1.0 / c.rsq;

  // normal / tangential split of the relative velocity (pair :154-163, wall :581-590)
  const double vnnr = c.vr[0] * c.dx + c.vr[1] * c.dy + c.vr[2] * c.dz;
  double vn1, vn2, vn3;
  if (c.vn_divide) {
    vn1 = c.dx * vnnr / c.rsq; vn2 = c.dy * vnnr / c.rsq; vn3 = c.dz * vnnr / c.rsq;
  } else {
    vn1 = c.dx * vnnr * rsqinv; vn2 = c.dy * vnnr * rsqinv; vn3 = c.dz * vnnr * rsqinv;
  }
  const double vt1 = c.vr[0] - vn1, vt2 = c.vr[1] - vn2, vt3 = c.vr[2] - vn3;

  // rotational contribution (pair :167-169, wall :594-596)
  const double wr1 = c.wsum[0] * rinv, wr2 = c.wsum[1] * rinv, wr3 = c.wsum[2] * rinv;

  // sqrt argument: pair (radsum-r)*radi*radj/radsum (:192) ; wall (radius-r)*radius (:600)
  const double harg = c.wall ? (c.rcontact - r) * c.rcontact
                             : (c.rcontact - r) * c.radi * c.radj / c.rcontact;
  const double polyhertz = sqrt(harg);
  const double sn = 2.0 * 1.0 / 1.82 * p.kn * polyhertz;  // :192 / :600
  const double st = 8.0 * 1.0 / 8.84 * p.kn * polyhertz;  // :193 / :601

  const double damp = 2.0 * sqrt(5.0 / 6.0) * beta * vnnr * rsqinv;                           // :198 / :605
  const double ccel = polyhertz * 4.0 / 5.46 * p.kn * (c.rcontact - r) * rinv - sqrt(sn * c.meff) * damp;  // :200 / :607

  // relative tangential surface velocity (:204-206 / :611-613)
  const double vtr1 = vt1 - (c.dz * wr2 - c.dy * wr3);
  const double vtr2 = vt2 - (c.dx * wr3 - c.dz * wr1);
  const double vtr3 = vt3 - (c.dy * wr1 - c.dx * wr2);

  // shear history (:212-230 / :619-635)
  if (shearupdate) {
    shear[0] += vtr1 * dt; shear[1] += vtr2 * dt; shear[2] += vtr3 * dt;
  }
  const double shrmag = sqrt(shear[0] * shear[0] + shear[1] * shear[1] + shear[2] * shear[2]);
  double rsht = shear[0] * c.dx + shear[1] * c.dy + shear[2] * c.dz;
  rsht *= rsqinv;
  if (shearupdate) {
    shear[0] -= rsht * c.dx; shear[1] -= rsht * c.dy; shear[2] -= rsht * c.dz;
  }

  // tangential force: spring on the history + dashpot (:234-236 / :639-641)
  const double kts = -polyhertz * 8.0 / 8.84 * p.kt;
  const double ctd = sqrt(st * c.meff) * 2.0 * sqrt(5.0 / 6.0) * beta;
  double fs1 = kts * shear[0] - ctd * vtr1;
  double fs2 = kts * shear[1] - ctd * vtr2;
  double fs3 = kts * shear[2] - ctd * vtr3;

  // Coulomb cap with history rescale (:240-255 / :645-660)
  const double fs = sqrt(fs1 * fs1 + fs2 * fs2 + fs3 * fs3);
  const double fn = p.xmu * fabs(ccel * r);
  if (fs > fn) {
    if (shrmag != 0.0) {
      const double ratio = fn / fs;
      const double e1 = ctd * vtr1 / 8.84 * 8.0 / p.kt;
      const double e2 = ctd * vtr2 / 8.84 * 8.0 / p.kt;
      const double e3 = ctd * vtr3 / 8.84 * 8.0 / p.kt;
      shear[0] = ratio * (shear[0] + e1) - e1;
      shear[1] = ratio * (shear[1] + e2) - e2;
      shear[2] = ratio * (shear[2] + e3) - e3;
      fs1 *= ratio; fs2 *= ratio; fs3 *= ratio;
    } else fs1 = fs2 = fs3 = 0.0;
  }

  o.fx = c.dx * ccel + fs1; o.fy = c.dy * ccel + fs2; o.fz = c.dz * ccel + fs3;  // :259-261 / :664-666
  o.t1 = rinv * (c.dy * fs3 - c.dz * fs2);                                        // :266-268 / :672-674
  o.t2 = rinv * (c.dz * fs1 - c.dx * fs3);
  o.t3 = rinv * (c.dx * fs2 - c.dy * fs1);
}

// Linear spring-dashpot with shear history.
// wall: fix_wall_granFix.cpp:441-554 ; pair: EXTERNAL stock PairGranHookeHistory::compute,
// identical with radius-r -> radsum-r and meff from both masses.
void hooke_history_contact(const ContactIn &c, const GranParams &p, double dt, bool shearupdate, double *shear,
                           ContactOut &o) {
  const double r = sqrt(c.rsq);
  const double rinv = 1.0 / r;
  const double rsqinv = 1.0 / c.rsq;
  const double vnnr = c.vr[0] * c.dx + c.vr[1] * c.dy + c.vr[2] * c.dz;
  const double vn1 = c.dx * vnnr * rsqinv, vn2 = c.dy * vnnr * rsqinv, vn3 = c.dz * vnnr * rsqinv;
  const double vt1 = c.vr[0] - vn1, vt2 = c.vr[1] - vn2, vt3 = c.vr[2] - vn3;
  const double wr1 = c.wsum[0] * rinv, wr2 = c.wsum[1] * rinv, wr3 = c.wsum[2] * rinv;

  const double damp = c.meff * p.gamman * vnnr * rsqinv;           // :484
  const double ccel = p.kn * (c.rcontact - r) * rinv - damp;       // :485

  const double vtr1 = vt1 - (c.dz * wr2 - c.dy * wr3);
  const double vtr2 = vt2 - (c.dx * wr3 - c.dz * wr1);
  const double vtr3 = vt3 - (c.dy * wr1 - c.dx * wr2);

  if (shearupdate) {
    shear[0] += vtr1 * dt; shear[1] += vtr2 * dt; shear[2] += vtr3 * dt;
  }
  const double shrmag = sqrt(shear[0] * shear[0] + shear[1] * shear[1] + shear[2] * shear[2]);
  double rsht = shear[0] * c.dx + shear[1] * c.dy + shear[2] * c.dz;
  rsht = rsht * rsqinv;
  if (shearupdate) {
    shear[0] -= rsht * c.dx; shear[1] -= rsht * c.dy; shear[2] -= rsht * c.dz;
  }

  double fs1 = -(p.kt * shear[0] + c.meff * p.gammat * vtr1);     // :515-517
  double fs2 = -(p.kt * shear[1] + c.meff * p.gammat * vtr2);
  double fs3 = -(p.kt * shear[2] + c.meff * p.gammat * vtr3);

  const double fs = sqrt(fs1 * fs1 + fs2 * fs2 + fs3 * fs3);
  const double fn = p.xmu * fabs(ccel * r);
  if (fs > fn) {
    if (shrmag != 0.0) {
      const double ratio = fn / fs;                                // :524-536
      shear[0] = ratio * (shear[0] + c.meff * p.gammat * vtr1 / p.kt) - c.meff * p.gammat * vtr1 / p.kt;
      shear[1] = ratio * (shear[1] + c.meff * p.gammat * vtr2 / p.kt) - c.meff * p.gammat * vtr2 / p.kt;
      shear[2] = ratio * (shear[2] + c.meff * p.gammat * vtr3 / p.kt) - c.meff * p.gammat * vtr3 / p.kt;
      fs1 *= ratio; fs2 *= ratio; fs3 *= ratio;
    } else fs1 = fs2 = fs3 = 0.0;
  }
  o.fx = c.dx * ccel + fs1; o.fy = c.dy * ccel + fs2; o.fz = c.dz * ccel + fs3;
  o.t1 = rinv * (c.dy * fs3 - c.dz * fs2);
  o.t2 = rinv * (c.dz * fs1 - c.dx * fs3);
  o.t3 = rinv * (c.dx * fs2 - c.dy * fs1);
}

// History-free Hooke law: fix_wall_granFix.cpp:356-437 (wall); pair version is EXTERNAL stock gran/hooke.
void hooke_contact(const ContactIn &c, const GranParams &p, ContactOut &o) {
  const double r = sqrt(c.rsq);
  const double rinv = 1.0 / r;
  const double rsqinv = 1.0 / c.rsq;
  const double vnnr = c.vr[0] * c.dx + c.vr[1] * c.dy + c.vr[2] * c.dz;
  const double vn1 = c.dx * vnnr * rsqinv, vn2 = c.dy * vnnr * rsqinv, vn3 = c.dz * vnnr * rsqinv;
  const double vt1 = c.vr[0] - vn1, vt2 = c.vr[1] - vn2, vt3 = c.vr[2] - vn3;
  const double wr1 = c.wsum[0] * rinv, wr2 = c.wsum[1] * rinv, wr3 = c.wsum[2] * rinv;
  const double damp = c.meff * p.gamman * vnnr * rsqinv;
  const double ccel = p.kn * (c.rcontact - r) * rinv - damp;
  const double vtr1 = vt1 - (c.dz * wr2 - c.dy * wr3);
  const double vtr2 = vt2 - (c.dx * wr3 - c.dz * wr1);
  const double vtr3 = vt3 - (c.dy * wr1 - c.dx * wr2);
  double vrel = vtr1 * vtr1 + vtr2 * vtr2 + vtr3 * vtr3;
  vrel = sqrt(vrel);
  const double fn = p.xmu * fabs(ccel * r);       // :411-414
  const double fs = c.meff * p.gammat * vrel;
  double ft = 0.0;
  if (vrel != 0.0) ft = (fn < fs ? fn : fs) / vrel;
  const double fs1 = -ft * vtr1, fs2 = -ft * vtr2, fs3 = -ft * vtr3;
  o.fx = c.dx * ccel + fs1; o.fy = c.dy * ccel + fs2; o.fz = c.dz * ccel + fs3;
  o.t1 = rinv * (c.dy * fs3 - c.dz * fs2);
  o.t2 = rinv * (c.dz * fs1 - c.dx * fs3);
  o.t3 = rinv * (c.dx * fs2 - c.dy * fs1);
}

}  // namespace

class PortBackend : public Backend {
 public:
  SimConfig cfg;
  double beta_pair;
  std::vector<double> beta_wall;
  // lubricate/poly isotropic constants (pair_lubricate_poly.cpp:545-559)
  double R0, RT0, RS0;

  const char *name() const { return "port"; }

  void init(const SimConfig &c, const AtomView &av, const BoxInfo &box, const StepInfo &) {
    cfg = c;
    beta_pair = (cfg.pair == sedi::PAIR_HERTZFIX_HISTORY) ? fix_beta(cfg.gran.gamman) : 0.0;
    beta_wall.assign(cfg.fixes.size(), 0.0);
    for (size_t k = 0; k < cfg.fixes.size(); k++)
      if (cfg.fixes[k].kind == sedi::FIX_WALL_GRAN && cfg.pair == sedi::PAIR_HERTZFIX_HISTORY)
        beta_wall[k] = fix_beta(cfg.fixes[k].wall.gamman);
    R0 = RT0 = RS0 = 0.0;
    if (cfg.lub.enabled) {
      // PairLubricatePoly::init_style, pair_lubricate_poly.cpp:501-559.  Any fix whose style contains "wall"
      // sets flagwall (:496-513) and the reference then dereferences wallfix as a stock FixWall; granular walls
      // are not FixWall objects, so with walls present the reference reads garbage.  The oracle defines
      // vol_T = box volume in all cases and says so here.
      const double MY_PI = sedi::SEDI_MY_PI;
      double vol_T = (box.hi[0] - box.lo[0]) * (box.hi[1] - box.lo[1]) * (box.hi[2] - box.lo[2]);
      double volP = 0.0;
      for (int i = 0; i < av.nlocal; i++) volP += (4.0 / 3.0) * MY_PI * pow(av.radius[i], 3.0);  // :540-542
      double vol_f = volP / vol_T;
      if (!cfg.lub.flagVF) vol_f = 0;
      const double mu = cfg.lub.mu;
      if (cfg.lub.flaglog == 0) {
        R0 = 6 * MY_PI * mu * (1.0 + 2.16 * vol_f);
        RT0 = 8 * MY_PI * mu;
        RS0 = 20.0 / 3.0 * MY_PI * mu * (1.0 + 3.33 * vol_f + 2.80 * vol_f * vol_f);
      } else {
        R0 = 6 * MY_PI * mu * (1.0 + 2.725 * vol_f - 6.583 * vol_f * vol_f);
        RT0 = 8 * MY_PI * mu * (1.0 + 0.749 * vol_f - 2.469 * vol_f * vol_f);
        RS0 = 20.0 / 3.0 * MY_PI * mu * (1.0 + 3.64 * vol_f - 6.95 * vol_f * vol_f);
      }
    }
  }

  // PairGranHertzFixHistory::compute (pair_gran_hertzFix_history.cpp:45-287) and the stock Hooke variants.
  void pair_granular(const AtomView &av, const NList &list, const StepInfo &st) {
    const bool shearupdate = !st.setupflag;  // :65-66
    const int nlocal = av.nlocal;
    const int kind = cfg.pair;
    for (int ii = 0; ii < list.inum; ii++) {
      const int i = list.ilist[ii];
      const double xi = av.x[i][0], yi = av.x[i][1], zi = av.x[i][2];
      const double radi = av.radius[i];
      int *touch = list.firsttouch ? list.firsttouch[i] : 0;
      double *allshear = list.firstshear ? list.firstshear[i] : 0;
      const int *jlist = list.firstneigh[i];
      const int jnum = list.numneigh[i];
      for (int jj = 0; jj < jnum; jj++) {
        const int j = jlist[jj] & NEIGHMASK_;
        ContactIn c;
        c.dx = xi - av.x[j][0]; c.dy = yi - av.x[j][1]; c.dz = zi - av.x[j][2];
        c.rsq = c.dx * c.dx + c.dy * c.dy + c.dz * c.dz;
        const double radj = av.radius[j];
        const double radsum = radi + radj;
        if (c.rsq >= radsum * radsum) {  // :131-139
          if (touch) { touch[jj] = 0; allshear[3 * jj] = allshear[3 * jj + 1] = allshear[3 * jj + 2] = 0.0; }
          continue;
        }
        for (int d = 0; d < 3; d++) {
          c.vr[d] = av.v[i][d] - av.v[j][d];
          c.wsum[d] = radi * av.omega[i][d] + radj * av.omega[j][d];
        }
        const double mi = av.rmass[i], mj = av.rmass[j];
        c.meff = mi * mj / (mi + mj);                         // :187-189
        if (av.mask[i] & cfg.freeze_group_bit) c.meff = mj;
        if (av.mask[j] & cfg.freeze_group_bit) c.meff = mi;
        c.rcontact = radsum; c.radi = radi; c.radj = radj; c.wall = false; c.vn_divide = false;
        ContactOut o;
        double dummy[3] = {0, 0, 0};
        if (kind == sedi::PAIR_HERTZFIX_HISTORY) {
          touch[jj] = 1;
          hertzfix_contact(c, cfg.gran, beta_pair, st.dt_init, shearupdate, &allshear[3 * jj], o);
        } else if (kind == sedi::PAIR_HOOKE_HISTORY) {
          touch[jj] = 1;
          hooke_history_contact(c, cfg.gran, st.dt_init, shearupdate, &allshear[3 * jj], o);
        } else {
          (void)dummy;
          hooke_contact(c, cfg.gran, o);
        }
        av.f[i][0] += o.fx; av.f[i][1] += o.fy; av.f[i][2] += o.fz;
        av.torque[i][0] -= radi * o.t1; av.torque[i][1] -= radi * o.t2; av.torque[i][2] -= radi * o.t3;
        if (j < nlocal) {  // newton off: ghost partners get nothing (:273-280)
          av.f[j][0] -= o.fx; av.f[j][1] -= o.fy; av.f[j][2] -= o.fz;
          av.torque[j][0] -= radj * o.t1; av.torque[j][1] -= radj * o.t2; av.torque[j][2] -= radj * o.t3;
        }
      }
    }
  }

  // PairLubricatePoly::compute, pair_lubricate_poly.cpp:65-444, non-shearing branch (no fix deform => Ef = 0,
  // :574-576; the Ef terms of :263-282 are kept as explicit +/- 0 products so that rounding is identical).
  void pair_lubricate(const AtomView &av, const NList &full, const StepInfo &) {
    const double MY_PI = sedi::SEDI_MY_PI;
    const double vxmu2f = 1.0;  // lj units
    const double mu = cfg.lub.mu;
    const double cutsq = cfg.lub.cut_global * cfg.lub.cut_global;
    const double cut_inner = cfg.lub.cut_inner;
    const int flaglog = cfg.lub.flaglog;
    const double Ef = 0.0;
    for (int ii = 0; ii < full.inum; ii++) {
      const int i = full.ilist[ii];
      const double xi = av.x[i][0], yi = av.x[i][1], zi = av.x[i][2];
      const double radi = av.radius[i];
      const double wi[3] = {av.omega[i][0], av.omega[i][1], av.omega[i][2]};
      if (cfg.lub.flagfld) {  // :213-221
        av.f[i][0] -= vxmu2f * R0 * radi * av.v[i][0];
        av.f[i][1] -= vxmu2f * R0 * radi * av.v[i][1];
        av.f[i][2] -= vxmu2f * R0 * radi * av.v[i][2];
        const double radi3 = radi * radi * radi;
        av.torque[i][0] -= vxmu2f * RT0 * radi3 * wi[0];
        av.torque[i][1] -= vxmu2f * RT0 * radi3 * wi[1];
        av.torque[i][2] -= vxmu2f * RT0 * radi3 * wi[2];
      }
      if (!cfg.lub.flagHI) continue;
      const int *jlist = full.firstneigh[i];
      const int jnum = full.numneigh[i];
      for (int jj = 0; jj < jnum; jj++) {
        const int j = jlist[jj];
        const double delx = xi - av.x[j][0], dely = yi - av.x[j][1], delz = zi - av.x[j][2];
        const double rsq = delx * delx + dely * dely + delz * delz;
        const double radj = av.radius[j];
        if (!(rsq < cutsq)) continue;  // :241
        const double r = sqrt(rsq);
        const double wj[3] = {av.omega[j][0], av.omega[j][1], av.omega[j][2]};
        // closest-approach points (:252-257)
        const double xl[3] = {-delx / r * radi, -dely / r * radi, -delz / r * radi};
        const double jl[3] = {-delx / r * radj, -dely / r * radj, -delz / r * radj};
        // surface velocities incl. the (zero) strain-rate term (:263-282)
        double vi[3], vj[3];
        vi[0] = av.v[i][0] + (wi[1] * xl[2] - wi[2] * xl[1]) - (Ef * xl[0] + Ef * xl[1] + Ef * xl[2]);
        vi[1] = av.v[i][1] + (wi[2] * xl[0] - wi[0] * xl[2]) - (Ef * xl[0] + Ef * xl[1] + Ef * xl[2]);
        vi[2] = av.v[i][2] + (wi[0] * xl[1] - wi[1] * xl[0]) - (Ef * xl[0] + Ef * xl[1] + Ef * xl[2]);
        vj[0] = av.v[j][0] - (wj[1] * jl[2] - wj[2] * jl[1]) + (Ef * jl[0] + Ef * jl[1] + Ef * jl[2]);
        vj[1] = av.v[j][1] - (wj[2] * jl[0] - wj[0] * jl[2]) + (Ef * jl[0] + Ef * jl[1] + Ef * jl[2]);
        vj[2] = av.v[j][2] - (wj[0] * jl[1] - wj[1] * jl[0]) + (Ef * jl[0] + Ef * jl[1] + Ef * jl[2]);
        double h_sep = r - radi - radj;                       // :286
        if (r < cut_inner) h_sep = 100 * radi + 100 * radj;   // Rui's modification, :294-297
        h_sep = h_sep / radi;
        const double beta0 = radj / radi;
        const double beta1 = 1.0 + beta0;
        double a_sq, a_sh = 0.0, a_pu = 0.0;
        if (flaglog) {  // :307-323
          a_sq = beta0 * beta0 / beta1 / beta1 / h_sep +
                 (1.0 + 7.0 * beta0 + beta0 * beta0) / 5.0 / pow(beta1, 3.0) * log(1.0 / h_sep);
          a_sq += (1.0 + 18.0 * beta0 - 29.0 * beta0 * beta0 + 18.0 * pow(beta0, 3.0) + pow(beta0, 4.0)) / 21.0 /
                  pow(beta1, 4.0) * h_sep * log(1.0 / h_sep);
          a_sq *= 6.0 * MY_PI * mu * radi;
          a_sh = 4.0 * beta0 * (2.0 + beta0 + 2.0 * beta0 * beta0) / 15.0 / pow(beta1, 3.0) * log(1.0 / h_sep);
          a_sh += 4.0 * (16.0 - 45.0 * beta0 + 58.0 * beta0 * beta0 - 45.0 * pow(beta0, 3.0) + 16.0 * pow(beta0, 4.0)) /
                  375.0 / pow(beta1, 4.0) * h_sep * log(1.0 / h_sep);
          a_sh *= 6.0 * MY_PI * mu * radi;
          a_pu = beta0 * (4.0 + beta0) / 10.0 / beta1 / beta1 * log(1.0 / h_sep);
          a_pu += (32.0 - 33.0 * beta0 + 83.0 * beta0 * beta0 + 43.0 * pow(beta0, 3.0)) / 250.0 / pow(beta1, 3.0) *
                  h_sep * log(1.0 / h_sep);
          a_pu *= 8.0 * MY_PI * mu * pow(radi, 3.0);
        } else a_sq = 6.0 * MY_PI * mu * radi * (beta0 * beta0 / beta1 / beta1 / h_sep);  // :324
        const double vr1 = vi[0] - vj[0], vr2 = vi[1] - vj[1], vr3 = vi[2] - vj[2];
        const double vnnr = (vr1 * delx + vr2 * dely + vr3 * delz) / r;  // :335
        const double vn1 = vnnr * delx / r, vn2 = vnnr * dely / r, vn3 = vnnr * delz / r;
        const double vt1 = vr1 - vn1, vt2 = vr2 - vn2, vt3 = vr3 - vn3;
        double fx = a_sq * vn1, fy = a_sq * vn2, fz = a_sq * vn3;  // :348-350
        if (flaglog) { fx = fx + a_sh * vt1; fy = fy + a_sh * vt2; fz = fz + a_sh * vt3; }
        fx *= vxmu2f; fy *= vxmu2f; fz *= vxmu2f;
        av.f[i][0] -= fx; av.f[i][1] -= fy; av.f[i][2] -= fz;      // force on i only, :368-370
        if (flaglog) {  // :374-399
          double tx = xl[1] * fz - xl[2] * fy, ty = xl[2] * fx - xl[0] * fz, tz = xl[0] * fy - xl[1] * fx;
          av.torque[i][0] -= vxmu2f * tx; av.torque[i][1] -= vxmu2f * ty; av.torque[i][2] -= vxmu2f * tz;
          const double wdotn = ((wi[0] - wj[0]) * delx + (wi[1] - wj[1]) * dely + (wi[2] - wj[2]) * delz) / r;
          const double wt1 = (wi[0] - wj[0]) - wdotn * delx / r;
          const double wt2 = (wi[1] - wj[1]) - wdotn * dely / r;
          const double wt3 = (wi[2] - wj[2]) - wdotn * delz / r;
          tx = a_pu * wt1; ty = a_pu * wt2; tz = a_pu * wt3;
          av.torque[i][0] -= vxmu2f * tx; av.torque[i][1] -= vxmu2f * ty; av.torque[i][2] -= vxmu2f * tz;
        }
      }
    }
  }

  // FixFluidDrag::post_force, fix_fluid_drag.cpp:114-164
  void fix_fdrag(int ifix, const AtomView &av, const FdragState &fs, const StepInfo &st) {
    const FixSpec &fx = cfg.fixes[ifix];
    const double timeStep = st.dt;  // update->dt read live (:121)
    for (int i = 0; i < av.nlocal; i++) {
      if (!(av.mask[i] & fx.groupbit)) continue;
      const double r = av.radius[i];
      const double rho = 3.0 * av.rmass[i] / (4.0 * sedi::SEDI_PI_LIBRARY * r * r * r);  // :147
      for (int d = 0; d < 3; d++) {
        const double acc = ((av.v[i][d] - fs.vOld[i][d]) / timeStep);
        av.f[i][d] += fs.ffluiddrag[i][d] + fx.carrier_rho / rho * 0.5 * av.rmass[i] * (fs.DuDt[i][d] - acc);
      }
      for (int d = 0; d < 3; d++) fs.vOld[i][d] = av.v[i][d];
    }
  }

  // FixCohe::post_force, fix_cohesive.cpp:138-263.  Quirks kept: the outer loop runs ii < nlocal over ilist
  // (:165, :216) and neighbour indices are not masked (:176, :228).
  void fix_cohesive(int ifix, const AtomView &av, const NList &half, const StepInfo &) {
    const FixSpec &fx = cfg.fixes[ifix];
    const double ah = fx.ah, lam = fx.lam, smin = fx.smin, smax = fx.smax;
    const double PInv = 0.25 / atan(1.0);  // :153
    const int nlocal = av.nlocal;
    for (int ii = 0; ii < nlocal; ii++) {
      const int i = half.ilist[ii];
      if (!(av.mask[i] & fx.groupbit)) continue;
      const double xi = av.x[i][0], yi = av.x[i][1], zi = av.x[i][2];
      const double radi = av.radius[i];
      const int *jlist = half.firstneigh[i];
      const int jnum = half.numneigh[i];
      for (int jj = 0; jj < jnum; jj++) {
        const int j = jlist[jj];
        const double delx = xi - av.x[j][0], dely = yi - av.x[j][1], delz = zi - av.x[j][2];
        const double rsq = delx * delx + dely * dely + delz * delz;
        const double radsum = radi + av.radius[j];
        if (!(rsq < (radsum + smax) * (radsum + smax))) continue;
        const double r = sqrt(rsq);
        const double del = r - radsum;
        double ccel;
        if (fx.opt == 0) {  // retarded van der Waals, three branches (:187-195)
          if (del > lam * PInv)
            ccel = -ah * radsum * lam * (6.4988e-3 - 4.5316e-4 * lam / del + 1.1326e-5 * lam * lam / del / del) / del / del / del;
          else if (del > smin)
            ccel = -ah * (lam + 22.242 * del) * radsum * lam / 24.0 / (lam + 11.121 * del) / (lam + 11.121 * del) / del / del;
          else
            ccel = -ah * (lam + 22.242 * smin) * radsum * lam / 24.0 / (lam + 11.121 * smin) / (lam + 11.121 * smin) / smin / smin;
        } else {            // opt 1 (:239-244)
          if (del > smin)
            ccel = -ah * pow(radsum, 6) / 6.0 / del / del / (r + radsum) / (r + radsum) / r / r / r;
          else
            ccel = -ah * pow(radsum, 6) / 6.0 / smin / smin / (smin + 2.0 * radsum) / (smin + 2.0 * radsum) /
                   (smin + radsum) / (smin + radsum) / (smin + radsum);
        }
        const double rinv = 1 / r;
        const double cx = delx * ccel * rinv, cy = dely * ccel * rinv, cz = delz * ccel * rinv;
        av.f[i][0] += cx; av.f[i][1] += cy; av.f[i][2] += cz;
        if (cfg.newton_pair || j < nlocal) { av.f[j][0] -= cx; av.f[j][1] -= cy; av.f[j][2] -= cz; }
      }
    }
  }

  // FixWallGranFix::post_force, fix_wall_granFix.cpp:247-345
  void fix_wall(int ifix, const AtomView &av, double **shear, const StepInfo &st) {
    const FixSpec &fx = cfg.fixes[ifix];
    double wlo = fx.lo, whi = fx.hi;
    double vwall[3] = {0.0, 0.0, 0.0};
    if (fx.wiggle) {  // :254-262
      const double omega = 2.0 * sedi::SEDI_MY_PI / fx.period;
      const double arg = omega * (st.ntimestep - fx.time_origin) * st.dt_init;
      if (fx.wallstyle == fx.axis) {
        wlo = fx.lo + fx.amplitude - fx.amplitude * cos(arg);
        whi = fx.hi + fx.amplitude - fx.amplitude * cos(arg);
      }
      vwall[fx.axis] = fx.amplitude * omega * sin(arg);
    } else if (fx.wshear) vwall[fx.axis] = fx.vshear;
    const bool shearupdate = !st.setupflag;
    const int pairstyle = cfg.pair;
    for (int i = 0; i < av.nlocal; i++) {
      if (!(av.mask[i] & fx.groupbit)) continue;
      double dx = 0.0, dy = 0.0, dz = 0.0;
      const double radius = av.radius[i];
      if (fx.wallstyle <= sedi::ZPLANE) {  // :294-308
        const double xc = av.x[i][fx.wallstyle];
        const double del1 = xc - wlo, del2 = whi - xc;
        const double d = (del1 < del2) ? del1 : -del2;
        if (fx.wallstyle == sedi::XPLANE) dx = d; else if (fx.wallstyle == sedi::YPLANE) dy = d; else dz = d;
      } else {                             // zcylinder :309-322
        const double delxy = sqrt(av.x[i][0] * av.x[i][0] + av.x[i][1] * av.x[i][1]);
        const double delr = fx.cylradius - delxy;
        if (delr > radius) dz = fx.cylradius;
        else {
          dx = -delr / delxy * av.x[i][0];
          dy = -delr / delxy * av.x[i][1];
          if (fx.wshear && fx.axis != 2) {
            vwall[0] = fx.vshear * av.x[i][1] / delxy;
            vwall[1] = -fx.vshear * av.x[i][0] / delxy;
            vwall[2] = 0.0;
          }
        }
      }
      const double rsq = dx * dx + dy * dy + dz * dz;
      if (rsq > radius * radius) {  // :326-331
        if (pairstyle != sedi::PAIR_HOOKE) shear[i][0] = shear[i][1] = shear[i][2] = 0.0;
        continue;
      }
      ContactIn c;
      c.dx = dx; c.dy = dy; c.dz = dz; c.rsq = rsq;
      for (int d = 0; d < 3; d++) { c.vr[d] = av.v[i][d] - vwall[d]; c.wsum[d] = radius * av.omega[i][d]; }
      c.meff = av.rmass[i]; c.rcontact = radius; c.radi = c.radj = 0.0; c.wall = true;
      ContactOut o;
      if (pairstyle == sedi::PAIR_HERTZFIX_HISTORY) {
        c.vn_divide = true;
        hertzfix_contact(c, fx.wall, beta_wall[ifix], st.dt_init, shearupdate, shear[i], o);
      } else if (pairstyle == sedi::PAIR_HOOKE_HISTORY) {
        c.vn_divide = false;
        hooke_history_contact(c, fx.wall, st.dt_init, shearupdate, shear[i], o);
      } else {
        c.vn_divide = false;
        hooke_contact(c, fx.wall, o);
      }
      av.f[i][0] += o.fx; av.f[i][1] += o.fy; av.f[i][2] += o.fz;
      av.torque[i][0] -= radius * o.t1; av.torque[i][1] -= radius * o.t2; av.torque[i][2] -= radius * o.t3;
    }
  }

 private:
  static const int NEIGHMASK_ = 0x3FFFFFFF;
};

Backend *make_port_backend() { return new PortBackend(); }

}  // namespace ora
