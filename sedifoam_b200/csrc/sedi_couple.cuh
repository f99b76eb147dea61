// sedi_couple.cuh -- two-way fluid<->particle coupling kernels (the OpenFOAM-side half of the hot path).
//
// Reference arithmetic followed:
//   cell owner            softParticleCloud::setPositionVeloCpuId + Cloud::move on an axis-aligned blockMesh
//                         (lammpsFoam/softParticleCloud.C:547-578, softParticle.C:102-151)
//   gather + fluid force  enhancedCloud::updateParticleUr/Alpha/updateDragOnParticles (enhancedCloud.C:56-257)
//   drag closures         ErgunWenYu::Jd (dragModels/ErgunWenYu/ErgunWenYu.C:104-132),
//                         SyamlalOBrien::Jd (dragModels/SyamlalOBrien/SyamlalOBrien.C:106-143)
//   scatter 1             enhancedCloud::particleToEulerianField (enhancedCloud.C:911-962)
//   scatter 2             enhancedCloud::calcTcFields (enhancedCloud.C:316-416)
// Cell fields keep OpenFOAM's memory layout (Field<vector> = interleaved xyz doubles, Field<scalar> = doubles) so the
// host solver's internalField() storage can be passed straight through the C-ABI.
#pragma once
#include "sedi_device.cuh"

namespace sedi {

static __device__ __constant__ double kROOTVSMALL = 1.0e-150;

struct MeshBox {  // single-block uniform hex mesh: cell = i + nx (j + ny k)
  double lo[3], hi[3], dx[3];
  int nc[3];
  // rectilinear form (graded blocks, axis-aligned blocks stacked into one tensor-product grid): face coordinates per
  // axis, nc[d] + 1 values each, and the host solver's cell label of every tensor cell (null = i + nx (j + ny k))
  int rect;
  const double *face[3];
  const int *label;
};

// owner interval of x among ascending face coordinates: largest I with f[I] <= x, -1 outside [f[0], f[n])
__device__ __forceinline__ int face_interval(const double *f, int n, double x) {
  if (!(x >= f[0]) || !(x < f[n])) return -1;
  int lo = 0, hi = n;   // invariant: f[lo] <= x < f[hi]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (x >= f[mid]) lo = mid; else hi = mid; }
  return lo;
}

enum { SEDI_DRAG_ERGUN_WENYU = 0, SEDI_DRAG_SYAMLAL_OBRIEN = 1 };
enum { SEDI_FORCE_DRAG = 1, SEDI_FORCE_PGRAD = 2, SEDI_FORCE_BUOY = 4, SEDI_FORCE_ADDEDMASS = 8, SEDI_FORCE_LIFT = 16,
       SEDI_FORCE_HISTORY = 32, SEDI_FORCE_WALL_LUB = 64, SEDI_FORCE_INLET = 128 };

// enhancedCloud::g1n (enhancedCloud.C:1372-1384)
__device__ __forceinline__ double g1n(double n) {
  if (n < 1) return 0.9279;
  return 0.9279 * (2 * n - 1) / n * pow(n, -n / (2 * n - 1)) + 0.001531;
}
// softParticleCloud::pointInRegion (softParticleCloud.C:1354-1415): option 1 = box, option 2 = hollow cylinder
// between the axis points (x1,y1,z1) and (x2,y2,z2) with radii r1 < r2 and an eccentric inner hole
__device__ __forceinline__ bool point_in_region(double px, double py, double pz, const double *box, int option, const double *ecc) {
  const double x1 = box[0], x2 = box[1], y1 = box[2], y2 = box[3], z1 = box[4], z2 = box[5], r1 = box[6], r2 = box[7];
  if (option == 1)
    return (px - x1) * (px - x2) < kROOTVSMALL && (py - y1) * (py - y2) < kROOTVSMALL && (pz - z1) * (pz - z2) < kROOTVSMALL;
  if (option == 2) {
    const double a0 = x2 - x1, a1 = y2 - y1, a2 = z2 - z1;
    const double h = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
    const double b0 = px - x1, b1 = py - y1, b2 = pz - z1;
    const double dot = a0 * b0 + a1 * b1 + a2 * b2;
    const double e0 = b0 - ecc[0], e1 = b1 - ecc[1], e2 = b2 - ecc[2];
    if (dot < 0.0 || dot > pow(h, 2.0)) return false;
    const double dsq = (b0 * b0 + b1 * b1 + b2 * b2) - dot * dot / pow(h, 2.0);
    const double dsqE = (e0 * e0 + e1 * e1 + e2 * e2) - dot * dot / pow(h, 2.0);
    return dsqE > r1 * r1 && dsq < r2 * r2;
  }
  return false;
}

__device__ __forceinline__ double jd_closure(int model, double Ur, double alpha, double pd, double nuf, double rhof) {
  const double beta = fmax(1.0 - alpha, kROOTVSMALL);
  if (model == SEDI_DRAG_ERGUN_WENYU) {
    const double bp = pow(beta, -2.65);
    const double Re = fmax(beta * Ur * pd / nuf, kROOTVSMALL);
    double Cds = 24.0 * (1.0 + 0.15 * pow(Re, 0.687)) / Re;
    if (Re > 1000.0) Cds = 0.44;
    double K = 0.75 * Cds * rhof * Ur * bp / pd;
    if (beta <= 0.8) K = 150.0 * alpha * nuf * rhof / ((beta * pd) * (beta * pd)) + 1.75 * rhof * Ur / (beta * pd);
    return K;
  }
  const double Ai = pow(beta, 4.14);
  double Bi = 0.8 * pow(beta, 1.28);
  if (beta > 0.85) Bi = pow(beta, 2.65);
  const double Re = fmax(Ur * pd / nuf, kROOTVSMALL);
  const double Vr = 0.5 * (Ai - 0.06 * Re + sqrt((0.06 * Re) * (0.06 * Re) + 0.12 * Re * (2.0 * Bi - Ai) + Ai * Ai));
  const double sq = 0.63 + 4.8 * sqrt(Vr / Re);
  return 0.75 * (sq * sq) * rhof * Ur / (pd * (Vr * Vr));
}

__global__ void k_locate_cells(const D4 *posr, int n, MeshBox M, int *cell) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const D4 x = posr[p];
  if (M.rect) {
    const int i0 = face_interval(M.face[0], M.nc[0], x.x), i1 = face_interval(M.face[1], M.nc[1], x.y), i2 = face_interval(M.face[2], M.nc[2], x.z);
    int c = -1;
    if (i0 >= 0 && i1 >= 0 && i2 >= 0) { c = i0 + M.nc[0] * (i1 + M.nc[1] * i2); if (M.label) c = M.label[c]; }
    cell[p] = c;
    return;
  }
  const double t0 = (x.x - M.lo[0]) / M.dx[0], t1 = (x.y - M.lo[1]) / M.dx[1], t2 = (x.z - M.lo[2]) / M.dx[2];
  const int i0 = (int)floor(t0), i1 = (int)floor(t1), i2 = (int)floor(t2);
  const bool in = !(t0 < 0.0) && !(t1 < 0.0) && !(t2 < 0.0) && i0 < M.nc[0] && i1 < M.nc[1] && i2 < M.nc[2];
  cell[p] = in ? i0 + M.nc[0] * (i1 + M.nc[1] * i2) : -1;
}

struct ForceParams {
  int n, model, flags;
  const D4 *posr, *velm;
  const int *cell;
  const double *Uf, *gamma, *gradp, *DDtU, *curlU;  // cell fields (interleaved vectors)
  double *uold[3];                                   // particle velocity at the previous coupling step
  double *fdrag[3], *dudt[3];                        // outputs: fix fdrag's per-atom arrays
  double *Uri, *magUri, *alphap, *Jd;                // optional diagnostics (interleaved / scalar), may be null
  double nub, rhob, g[3], deltaT;
  // history force (enhancedCloud.C:197-234): previous fluid velocity field, per-particle sumDeltaFb / n0, fluid step index
  const double *UfOld;
  double *hsum[3], *hn0;
  int timeIndex;
  // inlet forcing (:249-257)
  double inletForce[3], inletBox[9], inletEcc[3];
  int inletOption;
};

__global__ void __launch_bounds__(256) k_particle_force(const __grid_constant__ ForceParams P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  const int c = P.cell[i];
  const D4 x = P.posr[i], v = P.velm[i];
  double F0 = 0.0, F1 = 0.0, F2 = 0.0, u0 = 0.0, u1 = 0.0, u2 = 0.0, mag = 0.0, al = 0.0, jd = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
  const double d = x.w * 2.0;  // library.cpp:194
  if (c >= 0) {
    u0 = P.Uf[3 * (size_t)c] - v.x; u1 = P.Uf[3 * (size_t)c + 1] - v.y; u2 = P.Uf[3 * (size_t)c + 2] - v.z;
    mag = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
    al = P.gamma[c];
  }
  // the closure is evaluated for every particle (unlocated ones with Ur = 0, alpha = 0), as Jd_ = drag_->Jd(magUri_) does
  jd = jd_closure(P.model, mag, al, d, P.nub, P.rhob);
  if (c >= 0) {
    const double Vol = 3.14159265358979323846 * d * d * d / 6.0;
    if (P.DDtU) { a0 = P.DDtU[3 * (size_t)c]; a1 = P.DDtU[3 * (size_t)c + 1]; a2 = P.DDtU[3 * (size_t)c + 2]; }
    if (P.flags & SEDI_FORCE_DRAG) { F0 += jd * (1.0 - al) * Vol * u0; F1 += jd * (1.0 - al) * Vol * u1; F2 += jd * (1.0 - al) * Vol * u2; }
    if (P.flags & SEDI_FORCE_PGRAD) { F0 += -P.gradp[3 * (size_t)c] * Vol; F1 += -P.gradp[3 * (size_t)c + 1] * Vol; F2 += -P.gradp[3 * (size_t)c + 2] * Vol; }
    if (P.flags & SEDI_FORCE_BUOY) { F0 += -P.g[0] * P.rhob * Vol; F1 += -P.g[1] * P.rhob * Vol; F2 += -P.g[2] * P.rhob * Vol; }
    if (P.flags & SEDI_FORCE_ADDEDMASS) {
      double c0 = a0 - (v.x - P.uold[0][i]) / P.deltaT, c1 = a1 - (v.y - P.uold[1][i]) / P.deltaT, c2 = a2 - (v.z - P.uold[2][i]) / P.deltaT;
      const double m = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
      if (m > 10) { c0 = c0 / (m + kROOTVSMALL) * 10; c1 = c1 / (m + kROOTVSMALL) * 10; c2 = c2 / (m + kROOTVSMALL) * 10; }
      F0 += 0.5 * P.rhob * Vol * c0; F1 += 0.5 * P.rhob * Vol * c1; F2 += 0.5 * P.rhob * Vol * c2;
    }
    if (P.flags & SEDI_FORCE_LIFT) {
      const double w0 = P.curlU[3 * (size_t)c], w1 = P.curlU[3 * (size_t)c + 1], w2 = P.curlU[3 * (size_t)c + 2];
      const double magw = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
      const double coef = 1.6 * P.rhob * sqrt(P.nub) * (d * d);
      const double s = sqrt(magw + kROOTVSMALL);
      F0 += coef * (u1 * w2 - u2 * w1) / s; F1 += coef * (u2 * w0 - u0 * w2) / s; F2 += coef * (u0 * w1 - u1 * w0) / s;
    }
  }
  if (c >= 0 && (P.flags & (SEDI_FORCE_HISTORY | SEDI_FORCE_WALL_LUB | SEDI_FORCE_INLET))) {
    if (P.flags & SEDI_FORCE_HISTORY) {  // reduced-order Basset history force, Elghannay & Tafti 2016 (:197-234)
      const double tau_d = pow(d, 2.0) / P.nub;
      const double o0 = P.uold[0][i], o1 = P.uold[1][i], o2 = P.uold[2][i];
      const double q0 = P.UfOld[3 * (size_t)c] - o0, q1 = P.UfOld[3 * (size_t)c + 1] - o1, q2 = P.UfOld[3 * (size_t)c + 2] - o2;
      const double ReP = mag * d / P.nub, RePOld = sqrt(q0 * q0 + q1 * q1 + q2 * q2) * d / P.nub;
      const double qa = 0.632 / (ReP + kROOTVSMALL) + 0.087, qb = 0.632 / (RePOld + kROOTVSMALL) + 0.087;
      const double tau_h = tau_d * (qa * qa), tau_h_old = tau_d * (qb * qb);
      const double Cb = -1.5 * (d * d) * P.rhob * pow((3.1416 * P.nub), 0.5);
      const double nTotal = P.timeIndex;
      double n0 = P.hn0[i];
      const double tau_t = P.deltaT * (nTotal - n0);
      const double sdt = sqrt(P.deltaT);
      const double b0 = Cb * ((v.x - o0) / P.deltaT) / sdt, b1 = Cb * ((v.y - o1) / P.deltaT) / sdt, b2 = Cb * ((v.z - o2) / P.deltaT) / sdt;
      double S0 = P.hsum[0][i], S1 = P.hsum[1][i], S2 = P.hsum[2][i], g;
      if (tau_t < tau_h) {
        const double dn = nTotal - n0;
        S0 = S0 + b0; S1 = S1 + b1; S2 = S2 + b2;
        g = g1n(dn);
      } else {
        S0 = tau_h / tau_h_old * S0; S1 = tau_h / tau_h_old * S1; S2 = tau_h / tau_h_old * S2;
        const double dn = tau_h / P.deltaT;
        S0 = (dn - 1) / dn * S0; S1 = (dn - 1) / dn * S1; S2 = (dn - 1) / dn * S2;
        n0 = nTotal - dn;
        S0 = S0 + b0; S1 = S1 + b1; S2 = S2 + b2;
        g = g1n(dn);
      }
      P.hsum[0][i] = S0; P.hsum[1][i] = S1; P.hsum[2][i] = S2; P.hn0[i] = n0;
      F0 += (g * S0) * P.deltaT; F1 += (g * S1) * P.deltaT; F2 += (g * S2) * P.deltaT;
    }
    if (P.flags & SEDI_FORCE_WALL_LUB) {  // lubrication against the y = 0 wall (:235-248)
      const double distMin = 0.0001 * d, distMax = 0.1 * d;
      const double distWall = x.y - 0.5 * d;
      if (distWall < distMax && distWall > distMin) F1 += 6 * 3.1416 * P.nub * P.rhob * (-v.y) / distWall * (d * d) / 4.0;
    }
    if ((P.flags & SEDI_FORCE_INLET) &&
        sqrt(P.inletForce[0] * P.inletForce[0] + P.inletForce[1] * P.inletForce[1] + P.inletForce[2] * P.inletForce[2]) > 0) {
      if (point_in_region(x.x, x.y, x.z, P.inletBox, P.inletOption, P.inletEcc)) {  // replaces the force (:253)
        F0 = v.w * (P.inletForce[0] - v.x) / P.deltaT; F1 = v.w * (P.inletForce[1] - v.y) / P.deltaT; F2 = v.w * (P.inletForce[2] - v.z) / P.deltaT;
      }
    }
  }
  P.fdrag[0][i] = F0; P.fdrag[1][i] = F1; P.fdrag[2][i] = F2;
  P.dudt[0][i] = a0; P.dudt[1][i] = a1; P.dudt[2][i] = a2;
  if (P.Uri) { P.Uri[3 * (size_t)i] = u0; P.Uri[3 * (size_t)i + 1] = u1; P.Uri[3 * (size_t)i + 2] = u2; }
  if (P.magUri) P.magUri[i] = mag;
  if (P.alphap) P.alphap[i] = al;
  if (P.Jd) P.Jd[i] = jd;
}

// remember U for the next coupling step's added-mass term (softParticleCloud.C:571-572: UOld = U; U = Vnew)
__global__ void k_save_uold(const D4 *velm, int n, double *u0, double *u1, double *u2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const D4 v = velm[i];
  u0[i] = v.x; u1[i] = v.y; u2[i] = v.z;
}

// ---- particle -> cell accumulation with a FIXED summation order (bitwise reproducible run to run, no floating-point atomics).
// After every cell-owner location the rows of each fluid cell are listed in ascending row order (count -> scan -> fill -> per-cell
// sort: integer work only); a warp then sums one cell: lane l adds the entries l, l + 32, ... of the cell's list in that order and the
// 32 partial sums are combined by a fixed shuffle tree.
__global__ void k_fcell_count(const int *cell, int n, int *count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int c = cell[i]; if (c >= 0) atomicAdd(&count[c], 1); }
}
__global__ void k_fcell_fill(const int *cell, int n, const int *start, int *fill, int *rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const int c = cell[i]; if (c >= 0) rows[start[c] + atomicAdd(&fill[c], 1)] = i; }
}
// ascending row index inside every cell: one thread per cell.  Shell sort (Ciura gaps, extended by 2.25x): the fill order is nearly
// sorted for the usual tens of particles per cell, and a coarse mesh with 1e4..1e5 particles in a cell stays O(n^1.3), not O(n^2)
__host__ __device__ inline void fcell_shell_sort(int *rows, int s, int e) {
  const int n = e - s;
  if (n < 2) return;
  const int gaps[14] = {1636, 701, 301, 132, 57, 23, 10, 4, 1, 0, 0, 0, 0, 0};
  int g0 = 1636;
  // gaps above the table for very large cells: 3681, 8282, ... (x 2.25), applied from the largest one below n downwards
  int big[24], nb = 0;
  while ((long long)g0 * 9 / 4 < n && nb < 24) { g0 = (int)((long long)g0 * 9 / 4); big[nb++] = g0; }
  for (int q = nb - 1; q >= -9; q--) {
    const int gap = q >= 0 ? big[q] : gaps[-q - 1];
    if (gap <= 0 || gap >= n) continue;
    for (int a = s + gap; a < e; a++) {
      const int ra = rows[a];
      int b = a - gap;
      while (b >= s && rows[b] > ra) { rows[b + gap] = rows[b]; b -= gap; }
      rows[b + gap] = ra;
    }
  }
}
__global__ void k_fcell_sort(const int *start, int C, int *rows) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  fcell_shell_sort(rows, start[c], start[c + 1]);
}
__device__ __forceinline__ double warp_tree_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;   // lane 0 holds the sum
}

// scatter 1: gamma = sum Vp, Ue = sum Vp Up over the cell's particles   (enhancedCloud.C:918-930)
__global__ void __launch_bounds__(256) k_scatter_alpha_u(const D4 *posr, const D4 *velm, const int *start, const int *rows, int C, double *gamma, double *Ue) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double vol = 0.0, m0 = 0.0, m1 = 0.0, m2 = 0.0;
  const int e = start[c + 1];
  for (int k = start[c] + lane; k < e; k += 32) {
    const int i = rows[k];
    const double d = posr[i].w * 2.0;
    const double vp = 3.14159265358979323846 * d * d * d / 6.0;
    const D4 v = velm[i];
    vol += vp; m0 += vp * v.x; m1 += vp * v.y; m2 += vp * v.z;
  }
  vol = warp_tree_sum(vol); m0 = warp_tree_sum(m0); m1 = warp_tree_sum(m1); m2 = warp_tree_sum(m2);
  if (lane == 0) { gamma[c] = vol; Ue[3 * (size_t)c] = m0; Ue[3 * (size_t)c + 1] = m1; Ue[3 * (size_t)c + 2] = m2; }
}
// gamma /= V ; Ue /= V ; Ue /= gamma where gamma > ROOTVSMALL   (:932-962, smoothing flags off)
__global__ void k_finalize_alpha_u(int C, const double *cellV, double *gamma, double *Ue) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double V = cellV[c];
  const double g = gamma[c] / V;
  gamma[c] = g;
  double u0 = Ue[3 * (size_t)c] / V, u1 = Ue[3 * (size_t)c + 1] / V, u2 = Ue[3 * (size_t)c + 2] / V;
  if (g > kROOTVSMALL) { u0 /= g; u1 /= g; u2 /= g; }
  Ue[3 * (size_t)c] = u0; Ue[3 * (size_t)c + 1] = u1; Ue[3 * (size_t)c + 2] = u2;
}

// scatter 2: Asrc[c] = sum Vp Jd / Vc (Up - Uf[c])   (enhancedCloud.C:356-386); Omega stays 0 (:391)
__global__ void __launch_bounds__(256) k_scatter_asrc(const D4 *posr, const D4 *velm, const int *start, const int *rows, int C, const double *Uf,
                                                      const double *gamma, const double *cellV, int model, double nub, double rhob,
                                                      double *Asrc) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  const int e = start[c + 1], b = start[c];
  if (b < e) {
    const double f0 = Uf[3 * (size_t)c], f1 = Uf[3 * (size_t)c + 1], f2 = Uf[3 * (size_t)c + 2];
    const double gam = gamma[c], Vc = cellV[c];
    for (int k = b + lane; k < e; k += 32) {
      const int i = rows[k];
      const D4 v = velm[i];
      const double d = posr[i].w * 2.0;
      const double u0 = f0 - v.x, u1 = f1 - v.y, u2 = f2 - v.z;
      const double mag = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
      const double jd = jd_closure(model, mag, gam, d, nub, rhob);
      const double Vol = 3.14159265358979323846 * d * d * d / 6.0;
      const double omg = Vol * jd / Vc;
      s0 += omg * (v.x - f0); s1 += omg * (v.y - f1); s2 += omg * (v.z - f2);
    }
  }
  s0 = warp_tree_sum(s0); s1 = warp_tree_sum(s1); s2 = warp_tree_sum(s2);
  if (lane == 0) { Asrc[3 * (size_t)c] = s0; Asrc[3 * (size_t)c + 1] = s1; Asrc[3 * (size_t)c + 2] = s2; }
}
// Asrc *= (1-gamma) ; [smooth] ; Asrc /= (1-gamma)   (:407-416)
__global__ void k_finalize_asrc(int C, const double *gamma, double *Asrc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double w = 1 - gamma[c];
  for (int k = 0; k < 3; k++) { double a = Asrc[3 * (size_t)c + k] * w; a /= w; Asrc[3 * (size_t)c + k] = a; }
}

// ---- the reference's built-in invariants (printed every step: enhancedCloud.C:395-435 "total F before / after",
// :936-976 "total U solid before / after") and averageInfo() (:1341-1370).  Deterministic two-stage sums.
// mode 0: sum f[c][k] ; 1: sum f[c][k] V[c] (1 - gamma[c]) ; 2: sum f[c][k] V[c] gamma[c]
__global__ void __launch_bounds__(256) k_field_sum_partial(int C, const double *f, const double *cellV, const double *gamma, int mode, double *partial) {
  __shared__ double sm[3][256];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int c = blockIdx.x * 256 + threadIdx.x; c < C; c += gridDim.x * 256) {
    double w = 1.0;
    if (mode == 1) w = cellV[c] * (1 - gamma[c]);
    else if (mode == 2) w = cellV[c] * gamma[c];
    a0 += f[3 * (size_t)c] * w; a1 += f[3 * (size_t)c + 1] * w; a2 += f[3 * (size_t)c + 2] * w;
  }
  sm[0][threadIdx.x] = a0; sm[1][threadIdx.x] = a1; sm[2][threadIdx.x] = a2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) for (int k = 0; k < 3; k++) sm[k][threadIdx.x] += sm[k][threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) for (int k = 0; k < 3; k++) partial[3 * blockIdx.x + k] = sm[k][0];
}
// averageInfo: sum Vp, sum Vp U over the owned particles
__global__ void __launch_bounds__(256) k_particle_sum_partial(int n, const D4 *posr, const D4 *velm, double *partial) {
  __shared__ double sm[4][256];
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const D4 p = posr[i], v = velm[i];
    const double d = 2.0 * p.w;
    const double vol = d * d * d * 3.14159265358979323846 / 6.0;   // softParticle::Vol(), softParticleI.H:270-273
    a[0] += vol; a[1] += v.x * vol; a[2] += v.y * vol; a[3] += v.z * vol;
  }
  for (int k = 0; k < 4; k++) sm[k][threadIdx.x] = a[k];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) for (int k = 0; k < 4; k++) sm[k][threadIdx.x] += sm[k][threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) for (int k = 0; k < 4; k++) partial[4 * blockIdx.x + k] = sm[k][0];
}
__global__ void __launch_bounds__(256) k_sum_final(const double *partial, int nb, int ncomp, double *out) {
  __shared__ double sm[256];
  for (int k = 0; k < ncomp; k++) {
    double acc = 0.0;
    for (int c = threadIdx.x; c < nb; c += 256) acc += partial[(size_t)ncomp * c + k];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out[k] = sm[0];
    __syncthreads();
  }
}

}  // namespace sedi
