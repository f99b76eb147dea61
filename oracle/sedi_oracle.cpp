// oracle/sedi_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// C API (ctypes-facing) of the CPU oracle:
//   ora_*      : DEM time loop (oracle_driver.hpp) with the port backend or, in oracle/_ref/libsedi_ref.so, the
//                reference's own objects (ref_backend.cpp).
//   ora_foam_* : restatement of the OpenFOAM-side coupling arithmetic of lammpsFoam/enhancedCloud.C,
//                dragModels/*, softParticle.H, which cannot be stub-compiled (needs fvMesh/volFields).
//                PARITY UNPINNED at bit level for these (no buildable reference); pinned by the shipped
//                xiaocase3 curve (tests/golden/) and closed-form known answers in tests/.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "oracle_driver.hpp"

using namespace ora;

static const double ROOTVSMALL = 1.0e-150;  // OpenFOAM double-precision constant
static const double FOAM_PI = 3.14159265358979323846;  // constant::mathematical::pi = M_PI

extern "C" {

void *ora_create(int use_ref) {
#ifdef SEDI_HAVE_REF
  return new Sim(use_ref ? make_ref_backend() : make_port_backend());
#else
  if (use_ref) return 0;
  return new Sim(make_port_backend());
#endif
}
int ora_has_ref(void) {
#ifdef SEDI_HAVE_REF
  return 1;
#else
  return 0;
#endif
}
void ora_destroy(void *p) { delete (Sim *)p; }
void ora_command(void *p, const char *line) { ((Sim *)p)->command(line); }

void ora_set_box(void *p, const double *lo, const double *hi, int ntypes) {
  Sim *s = (Sim *)p;
  for (int d = 0; d < 3; d++) { s->cfg().boxlo[d] = lo[d]; s->cfg().boxhi[d] = hi[d]; }
  s->cfg().have_box = 1; s->cfg().ntypes = ntypes;
}

// same per-atom inputs as a read_data "Atoms" line: id type diameter density x y z
void ora_add_atoms(void *p, int n, const int *tag, const int *type, const double *diam, const double *rho,
                   const double *x, const double *v) {
  Sim *s = (Sim *)p;
  for (int i = 0; i < n; i++) s->script.add_atom(tag[i], type[i], diam[i], rho[i], x + 3 * i, v ? v + 3 * i : 0);
}

void ora_set_omega(void *p, const double *omega) {  // test hook: non-zero initial spin
  Sim *s = (Sim *)p;
  if (s->nlocal) { for (int i = 0; i < s->nlocal; i++) for (int d = 0; d < 3; d++) s->omega.rows[i][d] = omega[3 * i + d]; }
  else s->script.atoms.omega.assign(omega, omega + 3 * s->script.atoms.size());
}

int ora_nlocal(void *p) { Sim *s = (Sim *)p; return s->nlocal ? s->nlocal : (int)s->script.atoms.size(); }
int ora_nghost(void *p) { return ((Sim *)p)->nghost; }

void ora_get_atoms(void *p, double *x, double *v, double *omega, double *f, double *torque, int *tag) {
  Sim *s = (Sim *)p;
  const size_t n3 = 3 * (size_t)s->nlocal;
  if (x) memcpy(x, s->x.data.data(), n3 * sizeof(double));
  if (v) memcpy(v, s->v.data.data(), n3 * sizeof(double));
  if (omega) memcpy(omega, s->omega.data.data(), n3 * sizeof(double));
  if (f) memcpy(f, s->f.data.data(), n3 * sizeof(double));
  if (torque) memcpy(torque, s->torque.data.data(), n3 * sizeof(double));
  if (tag) memcpy(tag, s->tag.data(), s->nlocal * sizeof(int));
}

void ora_get_radius_mass(void *p, double *radius, double *rmass) {
  Sim *s = (Sim *)p;
  memcpy(radius, s->radius.data(), s->nlocal * sizeof(double));
  memcpy(rmass, s->rmass.data(), s->nlocal * sizeof(double));
}

// lammps_put_local_info semantics (library.cpp:314-367): values matched to atoms by tag; DuDt is ignored there.
void ora_put_fdrag(void *p, int n, const double *fdrag, const int *foamCpuId, const int *tagIn) {
  Sim *s = (Sim *)p;
  if (!s->nlocal && s->script.atoms.size()) s->load_atoms();
  std::vector<std::pair<int, int> > a(n), b(n);
  for (int i = 0; i < n; i++) { a[i] = std::make_pair(s->tag[i], i); b[i] = std::make_pair(tagIn[i], i); }
  std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
  for (int j = 0; j < n; j++) {
    const int to = a[j].second, from = b[j].second;
    if (foamCpuId) s->foamCpuId[to] = foamCpuId[from];
    for (int d = 0; d < 3; d++) s->ffluiddrag.rows[to][d] = fdrag[3 * from + d];
  }
}

void ora_set_timestep(void *p, double dt) { ((Sim *)p)->cfg().dt = dt; }
void ora_run(void *p, long long n) { ((Sim *)p)->run(n); }
void ora_setup(void *p) { Sim *s = (Sim *)p; if (!s->setup_done) s->setup(); }
void ora_reneighbor(void *p) { ((Sim *)p)->reneighbor(); }

long long ora_stat(void *p, int which) {
  Sim *s = (Sim *)p;
  switch (which) {
    case 0: return s->nbuilds;
    case 1: return s->npair_evals;
    case 2: return s->nsteps_done;
    case 3: return (long long)s->gran.neigh.size();
    case 4: return (long long)s->half.neigh.size();
    case 5: return (long long)s->full.neigh.size();
    default: return -1;
  }
}

// Neighbour-list export as (tag_i, tag_j) rows (ghost partners report the tag of their source atom).
// which: 0 granular half list, 1 type-cutoff half list (fix cohesive), 2 full list (lubricate/poly).
long long ora_get_pairs(void *p, int which, int *ti, int *tj, int *touch, double *shear, long long cap, int *ghost) {
  Sim *s = (Sim *)p;
  CSRList &l = which == 0 ? s->gran : which == 1 ? s->half : s->full;
  long long m = 0;
  for (int i = 0; i < (int)l.numneigh.size(); i++)
    for (int jj = 0; jj < l.numneigh[i]; jj++) {
      const int k = l.offset[i] + jj;
      if (m < cap) {
        ti[m] = s->tag[i]; tj[m] = s->tag[l.neigh[k]];
        if (ghost) ghost[m] = (l.neigh[k] >= s->nlocal) ? 1 : 0;
        if (touch) touch[m] = (l.history ? l.touch[k] : 0);
        if (shear) for (int d = 0; d < 3; d++) shear[3 * m + d] = l.history ? l.shear[3 * (size_t)k + d] : 0.0;
      }
      m++;
    }
  return m;
}

void ora_get_wall_shear(void *p, int wall, double *out) {
  Sim *s = (Sim *)p;
  memcpy(out, s->wallshear[wall].data.data(), 3 * (size_t)s->nlocal * sizeof(double));
}

// ===================================================================================================
// OpenFOAM-side coupling arithmetic (restated; see header comment)
// ===================================================================================================

// ErgunWenYu::Jd, lammpsFoam/dragModels/ErgunWenYu/ErgunWenYu.C:86-145
void ora_foam_jd_ergun_wenyu(int n, const double *Ur, const double *alpha, const double *pd, double nuf, double rhof,
                             double *jd) {
  for (int i = 0; i < n; i++) {
    const double beta = fmax(1.0 - alpha[i], ROOTVSMALL);       // :104
    const double bp = pow(beta, -2.65);                         // :105
    const double Re = fmax(beta * Ur[i] * pd[i] / nuf, ROOTVSMALL);  // :106
    double Cds = 24.0 * (1.0 + 0.15 * pow(Re, 0.687)) / Re;     // :107
    if (Re > 1000.0) Cds = 0.44;                                // :109-115
    double K = 0.75 * Cds * rhof * Ur[i] * bp / pd[i];          // Wen & Yu :118
    if (beta <= 0.8)                                            // Ergun :122-132
      K = 150.0 * alpha[i] * nuf * rhof / ((beta * pd[i]) * (beta * pd[i])) + 1.75 * rhof * Ur[i] / (beta * pd[i]);
    jd[i] = K;
  }
}

// SyamlalOBrien::Jd, lammpsFoam/dragModels/SyamlalOBrien/SyamlalOBrien.C:85-144
void ora_foam_jd_syamlal_obrien(int n, const double *Ur, const double *alpha, const double *pd, double nuf, double rhof,
                                double *jd) {
  for (int i = 0; i < n; i++) {
    const double beta = fmax(1.0 - alpha[i], ROOTVSMALL);
    const double Ai = pow(beta, 4.14);
    double Bi = 0.8 * pow(beta, 1.28);
    if (beta > 0.85) Bi = pow(beta, 2.65);
    const double Re = fmax(Ur[i] * pd[i] / nuf, ROOTVSMALL);
    const double Vr = 0.5 * (Ai - 0.06 * Re + sqrt((0.06 * Re) * (0.06 * Re) + 0.12 * Re * (2.0 * Bi - Ai) + Ai * Ai));
    const double sq = 0.63 + 4.8 * sqrt(Vr / Re);
    const double Cds = sq * sq;
    jd[i] = 0.75 * Cds * rhof * Ur[i] / (pd[i] * (Vr * Vr));
  }
}

static void jd_dispatch(int model, int n, const double *Ur, const double *alpha, const double *pd, double nuf, double rhof, double *jd) {
  if (model == 0) ora_foam_jd_ergun_wenyu(n, Ur, alpha, pd, nuf, rhof, jd);
  else ora_foam_jd_syamlal_obrien(n, Ur, alpha, pd, nuf, rhof, jd);
}

// flags bit layout shared with include/sedi_b200.h (SEDI_FORCE_*)
enum { F_DRAG = 1, F_PGRAD = 2, F_BUOY = 4, F_ADDEDMASS = 8, F_LIFT = 16 };

// updateParticleUr (enhancedCloud.C:83-109) + updateParticleAlpha (:56-76) + Jd (:129) +
// updateDragOnParticles (:112-257; drag, pressure gradient, buoyancy, added mass, lift branches).
// cell < 0 (particle not located): Uri = 0, force = 0.  The reference's extra ++pIter in that branch (:101)
// walks off the particle list when the lost particle is the last one (undefined behaviour), so it is NOT restated.
void ora_foam_particle_force(int n, const int *cell, const double *d, const double *U, const double *UOld,
                             const double *Uf, const double *gamma, const double *gradp, const double *DDtU,
                             const double *curlU, int model, int flags, double nub, double rhob, const double *g,
                             double deltaT, double *Uri, double *magUri, double *alphap, double *Jd, double *pDrag,
                             double *pDuDt) {
  for (int i = 0; i < n; i++) {
    const int c = cell[i];
    if (c < 0) { for (int k = 0; k < 3; k++) Uri[3 * i + k] = 0.0; magUri[i] = 0.0; alphap[i] = 0.0; continue; }
    for (int k = 0; k < 3; k++) Uri[3 * i + k] = Uf[3 * c + k] - U[3 * i + k];                                  // :106
    magUri[i] = sqrt(Uri[3 * i] * Uri[3 * i] + Uri[3 * i + 1] * Uri[3 * i + 1] + Uri[3 * i + 2] * Uri[3 * i + 2]);  // :107
    alphap[i] = gamma[c];                                                                                       // :74
  }
  jd_dispatch(model, n, magUri, alphap, d, nub, rhob, Jd);
  for (int i = 0; i < n; i++) {
    const int c = cell[i];
    double F[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) pDuDt[3 * i + k] = 0.0;
    if (c < 0) { for (int k = 0; k < 3; k++) pDrag[3 * i + k] = 0.0; continue; }
    const double Vol = FOAM_PI * d[i] * d[i] * d[i] / 6.0;  // softParticle.H:270-273
    for (int k = 0; k < 3; k++) pDuDt[3 * i + k] = DDtU ? DDtU[3 * c + k] : 0.0;                               // :155
    if (flags & F_DRAG) for (int k = 0; k < 3; k++) F[k] += Jd[i] * (1.0 - alphap[i]) * Vol * Uri[3 * i + k];  // :157-162
    if (flags & F_PGRAD) for (int k = 0; k < 3; k++) F[k] += -gradp[3 * c + k] * Vol;                          // :163-168
    if (flags & F_BUOY) for (int k = 0; k < 3; k++) F[k] += -g[k] * rhob * Vol;                                // :169-173
    if (flags & F_ADDEDMASS) {                                                                                 // :175-188
      double acc[3], m2 = 0.0;
      for (int k = 0; k < 3; k++) { const double dupdt = (U[3 * i + k] - UOld[3 * i + k]) / deltaT; acc[k] = DDtU[3 * c + k] - dupdt; m2 += acc[k] * acc[k]; }
      const double m = sqrt(m2);
      if (m > 10) for (int k = 0; k < 3; k++) acc[k] = acc[k] / (m + ROOTVSMALL) * 10;
      for (int k = 0; k < 3; k++) F[k] += 0.5 * rhob * Vol * acc[k];
    }
    if (flags & F_LIFT) {                                                                                      // :189-196
      const double *w = &curlU[3 * c], *u = &Uri[3 * i];
      const double cr[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
      const double magw = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
      const double coef = 1.6 * rhob * sqrt(nub) * (d[i] * d[i]);
      for (int k = 0; k < 3; k++) F[k] += coef * cr[k] / sqrt(magw + ROOTVSMALL);
    }
    for (int k = 0; k < 3; k++) pDrag[3 * i + k] = F[k];
  }
}

// The remaining branches of updateDragOnParticles, applied ON TOP of ora_foam_particle_force's pDrag (the reference
// evaluates them in this order inside the same loop body, enhancedCloud.C:197-257):
//   history force   :197-234  reduced-order Basset model (Elghannay & Tafti 2016), per-particle state sumDeltaFb, n0
//   wall lubrication :235-248  y = 0 wall, active for 1e-4 d < gap < 0.1 d
//   inlet forcing   :249-257  REPLACES the force inside the inlet region (softParticleCloud::pointInRegion :1354-1415)
// flags: 32 history, 64 lubrication, 128 inlet.  region_option = addParticleOption (1 box, 2 hollow cylinder).
static double ora_g1n(double n) {  // enhancedCloud.C:1372-1384
  if (n < 1) return 0.9279;
  return 0.9279 * (2 * n - 1) / n * pow(n, -n / (2 * n - 1)) + 0.001531;
}
static bool ora_point_in_region(const double *pt, const double *box, int option, const double *ecc) {
  const double x1 = box[0], x2 = box[1], y1 = box[2], y2 = box[3], z1 = box[4], z2 = box[5], r1 = box[6], r2 = box[7];
  if (option == 1)
    return (pt[0] - x1) * (pt[0] - x2) < ROOTVSMALL && (pt[1] - y1) * (pt[1] - y2) < ROOTVSMALL && (pt[2] - z1) * (pt[2] - z2) < ROOTVSMALL;
  if (option == 2) {
    const double a[3] = {x2 - x1, y2 - y1, z2 - z1};
    const double h = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const double b[3] = {pt[0] - x1, pt[1] - y1, pt[2] - z1};
    const double dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    const double be[3] = {b[0] - ecc[0], b[1] - ecc[1], b[2] - ecc[2]};
    if (dot < 0.0 || dot > pow(h, 2)) return false;
    const double dsq = (b[0] * b[0] + b[1] * b[1] + b[2] * b[2]) - dot * dot / pow(h, 2);
    const double dsqE = (be[0] * be[0] + be[1] * be[1] + be[2] * be[2]) - dot * dot / pow(h, 2);
    return dsqE > r1 * r1 && dsq < r2 * r2;
  }
  return false;
}
void ora_foam_particle_force_extra(int n, const int *cell, const double *x, const double *d, const double *mass, const double *U,
                                   const double *UOld, const double *Uf, const double *UfOld, int flags, double nub, double rhob,
                                   double deltaT, int timeIndex, double *sumDeltaFb, double *n0, const double *inletForce,
                                   const double *inletBox, int region_option, const double *ecc, double *pDrag) {
  for (int i = 0; i < n; i++) {
    const int c = cell[i];
    if (c < 0) continue;
    double *F = &pDrag[3 * i];
    if (flags & 32) {
      const double tau_d = pow(d[i], 2) / nub;
      double Uri[3], UriOld[3], mU = 0, mUo = 0;
      for (int k = 0; k < 3; k++) { Uri[k] = Uf[3 * c + k] - U[3 * i + k]; UriOld[k] = UfOld[3 * c + k] - UOld[3 * i + k]; mU += Uri[k] * Uri[k]; mUo += UriOld[k] * UriOld[k]; }
      const double ReP = sqrt(mU) * d[i] / nub, RePOld = sqrt(mUo) * d[i] / nub;
      const double q = 0.632 / (ReP + ROOTVSMALL) + 0.087, qo = 0.632 / (RePOld + ROOTVSMALL) + 0.087;
      const double tau_h = tau_d * (q * q), tau_h_old = tau_d * (qo * qo);
      const double Cb = -1.5 * (d[i] * d[i]) * rhob * pow((3.1416 * nub), 0.5);
      const double nTotal = timeIndex;
      const double tau_t = deltaT * (nTotal - n0[i]);
      double FH[3];
      double *S = &sumDeltaFb[3 * i];
      double dfb[3];
      for (int k = 0; k < 3; k++) dfb[k] = Cb * ((U[3 * i + k] - UOld[3 * i + k]) / deltaT) / sqrt(deltaT);
      if (tau_t < tau_h) {
        const double dn = nTotal - n0[i];
        for (int k = 0; k < 3; k++) S[k] = S[k] + dfb[k];
        const double g = ora_g1n(dn);
        for (int k = 0; k < 3; k++) FH[k] = g * S[k];
      } else {
        for (int k = 0; k < 3; k++) S[k] = tau_h / tau_h_old * S[k];
        const double dn = tau_h / deltaT;
        for (int k = 0; k < 3; k++) S[k] = (dn - 1) / dn * S[k];
        n0[i] = nTotal - dn;
        for (int k = 0; k < 3; k++) S[k] = S[k] + dfb[k];
        const double g = ora_g1n(dn);
        for (int k = 0; k < 3; k++) FH[k] = g * S[k];
      }
      for (int k = 0; k < 3; k++) F[k] += FH[k] * deltaT;
    }
    if (flags & 64) {
      const double distMin = 0.0001 * d[i], distMax = 0.1 * d[i];
      const double distWall = x[3 * i + 1] - 0.5 * d[i];
      const double pVel = U[3 * i + 1];
      if (distWall < distMax && distWall > distMin) F[1] += 6 * 3.1416 * nub * rhob * (-pVel) / distWall * (d[i] * d[i]) / 4.0;
    }
    if ((flags & 128) && sqrt(inletForce[0] * inletForce[0] + inletForce[1] * inletForce[1] + inletForce[2] * inletForce[2]) > 0) {
      if (ora_point_in_region(&x[3 * i], inletBox, region_option, ecc))
        for (int k = 0; k < 3; k++) F[k] = mass[i] * (inletForce[k] - U[3 * i + k]) / deltaT;
    }
  }
}

// Cell owner on a rectilinear mesh (graded blocks / axis-aligned blocks stacked into a tensor-product grid, SURVEY
// 8a15): the face interval that contains the particle centre on each axis, mapped to the host's cell label.  Replaces
// the face-to-face tracking of softParticle::move (lammpsFoam/softParticle.C:102-151), whose result on such a mesh is
// the cell containing the end point.  Linear scan on purpose (the CUDA kernel bisects).
void ora_foam_cell_owner_rect(int n, const double *x, const int *nc, const double *xf, const double *yf, const double *zf,
                              const int *label, int *cell) {
  const double *f[3] = {xf, yf, zf};
  for (int p = 0; p < n; p++) {
    int idx[3]; bool in = true;
    for (int k = 0; k < 3; k++) {
      const double v = x[3 * p + k];
      idx[k] = -1;
      for (int i = 0; i < nc[k]; i++) if (v >= f[k][i] && v < f[k][i + 1]) { idx[k] = i; break; }
      if (idx[k] < 0) in = false;
    }
    if (!in) { cell[p] = -1; continue; }
    const int t = idx[0] + nc[0] * (idx[1] + nc[1] * idx[2]);
    cell[p] = label ? label[t] : t;
  }
}

// Cell owner on a single-block axis-aligned uniform blockMesh: cell = i + nx (j + ny k)  (SURVEY 8a15, Appendix B2).
// Points outside the block get -1 (the reference deletes such particles on the Foam side, softParticle.C:177-184).
void ora_foam_cell_owner(int n, const double *x, const double *lo, const double *hi, const int *ncell, int *cell) {
  for (int p = 0; p < n; p++) {
    int idx[3]; bool in = true;
    for (int k = 0; k < 3; k++) {
      const double dx = (hi[k] - lo[k]) / ncell[k];
      const double t = (x[3 * p + k] - lo[k]) / dx;
      idx[k] = (int)floor(t);
      if (t < 0.0 || idx[k] >= ncell[k]) in = false;
    }
    cell[p] = in ? idx[0] + ncell[0] * (idx[1] + ncell[1] * idx[2]) : -1;
  }
}

// particleToEulerianField without the optional diffusion smoothing (enhancedCloud.C:911-962):
// gamma = sum Vp / Vc ; Ue = sum Vp Up / Vc ; Ue /= gamma where gamma > ROOTVSMALL.
void ora_foam_particle_to_eulerian(int n, const int *cell, const double *d, const double *U, int C, const double *cellV,
                                   double *gamma, double *Ue) {
  for (int c = 0; c < C; c++) { gamma[c] = 0.0; Ue[3 * c] = Ue[3 * c + 1] = Ue[3 * c + 2] = 0.0; }
  for (int i = 0; i < n; i++) {
    const int c = cell[i];
    if (c < 0) continue;
    const double Vol = FOAM_PI * d[i] * d[i] * d[i] / 6.0;
    gamma[c] += Vol;
    for (int k = 0; k < 3; k++) Ue[3 * c + k] += Vol * U[3 * i + k];
  }
  for (int c = 0; c < C; c++) {
    gamma[c] /= cellV[c];
    for (int k = 0; k < 3; k++) Ue[3 * c + k] /= cellV[c];
    if (gamma[c] > ROOTVSMALL) for (int k = 0; k < 3; k++) Ue[3 * c + k] /= gamma[c];
  }
}

// calcTcFields without smoothing (enhancedCloud.C:316-416): alpha_p, Uri, Jd recomputed; omg = Vp Jd / Vc;
// Asrc[c] += omg (Up - Uf[c]); Omega is accumulated and then zeroed (:383,:391) => always 0;
// Asrc *= (1-gamma) ; [smooth] ; Asrc /= (1-gamma)  (:407-416, kept: the round trip is not an exact identity).
void ora_foam_calc_tc(int n, const int *cell, const double *d, const double *U, const double *Uf, const double *gamma,
                      int C, const double *cellV, int model, double nub, double rhob, double *Asrc, double *Omega) {
  std::vector<double> mag(n), al(n), jd(n);
  for (int i = 0; i < n; i++) {
    const int c = cell[i];
    if (c < 0) { mag[i] = 0.0; al[i] = 0.0; continue; }
    double u[3];
    for (int k = 0; k < 3; k++) u[k] = Uf[3 * c + k] - U[3 * i + k];
    mag[i] = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    al[i] = gamma[c];
  }
  jd_dispatch(model, n, mag.data(), al.data(), d, nub, rhob, jd.data());
  for (int c = 0; c < C; c++) { Omega[c] = 0.0; Asrc[3 * c] = Asrc[3 * c + 1] = Asrc[3 * c + 2] = 0.0; }
  for (int i = 0; i < n; i++) {
    const int c = cell[i];
    if (c < 0) continue;
    const double Vol = FOAM_PI * d[i] * d[i] * d[i] / 6.0;
    const double omg = Vol * jd[i] / cellV[c];
    for (int k = 0; k < 3; k++) Asrc[3 * c + k] += omg * (U[3 * i + k] - Uf[3 * c + k]);
  }
  for (int c = 0; c < C; c++) {
    for (int k = 0; k < 3; k++) { Asrc[3 * c + k] = Asrc[3 * c + k] * (1 - gamma[c]); Asrc[3 * c + k] /= (1 - gamma[c]); }
  }
}

}  // extern "C"
