// sedi_smooth.cuh -- diffusion-based coarse graining of the Eulerian particle fields on the GPU.
//
// Reference: enhancedCloud::smoothField (lammpsFoam/enhancedCloud.C:790-907) solves, for `diffusionSteps` pseudo-time
// steps of d_tau = (b^2/4)/diffusionSteps (b = diffusionBandWidth, :564-568),
//         (phi^{n+1} - phi^n) / d_tau = div( D grad phi^{n+1} ),   zeroGradient on every patch (:810),
// with D = smoothDirection (default identity, :578-583), by OpenFOAM's PCG/DIC to 1e-10 (cases/*/system/fvSolution).
// On the single-block uniform blockMesh this is the symmetric positive definite 7-point system
//         (1 + sum_faces w_f) phi_c - sum_faces w_f phi_nb = phi^n_c ,   w_f = d_tau D_nn / dx_n^2 ,
// solved here by Jacobi-preconditioned conjugate gradients with fixed-order (deterministic) reductions.  The solve is
// conservative: sum phi V is preserved (the reference prints exactly this check, enhancedCloud.C:434-435, 975-976).
// Only the diagonal of smoothDirection acts on an orthogonal mesh's implicit operator; off-diagonals are ignored.
#pragma once
#include "sedi_device.cuh"

namespace sedi {

// uniform mesh: w = d_tau * D_dd / dx_d^2.  Rectilinear mesh (rect != 0): finite-volume form per cell,
//   (1 + sum_f w_f) phi_c - sum_f w_f phi_nb = phi^n_c ,  w_f = dt_d / (h_c * 0.5 (h_c + h_nb)) ,  dt_d = d_tau * D_dd,
// with h the cell widths along the face normal (exactly fvm::laplacian's  A_f / (V_c delta_f)  on an orthogonal mesh).
// The row form is self-adjoint in the volume-weighted inner product, which is the one the CG loop then uses
// (equivalent to PCG on OpenFOAM's volume-integrated symmetric matrix).  Fields are stored under the host's cell
// labels; `label` maps the tensor index to them (null = identity), `diag` holds the matrix diagonal by label.
struct SmoothGrid {
  int nx, ny, nz; double wx, wy, wz;
  int rect; double dtx, dty, dtz; const double *hx, *hy, *hz; const int *label; const double *diag;
};

__device__ __forceinline__ int smooth_lab(const SmoothGrid &G, int t) { return G.label ? G.label[t] : t; }
// y = A x on a rectilinear mesh, one thread per tensor cell; also used (x == null) to build the diagonal
__global__ void k_smooth_apply_rect(SmoothGrid G, const double *x, double *y, int stride, int off, double *diag_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = G.nx * G.ny * G.nz;
  if (t >= C) return;
  const int i = t % G.nx, j = (t / G.nx) % G.ny, k = t / (G.nx * G.ny);
  const int c = smooth_lab(G, t);
  const double xc = x ? x[(size_t)c * stride + off] : 0.0;
  double acc = xc, dg = 1.0;
  const double hx = G.hx[i], hy = G.hy[j], hz = G.hz[k];
#define SEDI_FACE(cond, tn, h, hn, dt)                                                                 \
  if (cond) { const double w = (dt) / ((h) * (0.5 * ((h) + (hn)))); dg += w;                          \
              if (x) acc += w * (xc - x[(size_t)smooth_lab(G, tn) * stride + off]); }
  SEDI_FACE(i > 0, t - 1, hx, G.hx[i - 1], G.dtx)
  SEDI_FACE(i < G.nx - 1, t + 1, hx, G.hx[i + 1], G.dtx)
  SEDI_FACE(j > 0, t - G.nx, hy, G.hy[j - 1], G.dty)
  SEDI_FACE(j < G.ny - 1, t + G.nx, hy, G.hy[j + 1], G.dty)
  SEDI_FACE(k > 0, t - G.nx * G.ny, hz, G.hz[k - 1], G.dtz)
  SEDI_FACE(k < G.nz - 1, t + G.nx * G.ny, hz, G.hz[k + 1], G.dtz)
#undef SEDI_FACE
  if (x) y[c] = acc;
  if (diag_out) diag_out[c] = dg;
}

// y = A x  for one component of an interleaved field (stride = 1 scalar, 3 vector)
__global__ void k_smooth_apply(SmoothGrid G, const double *x, double *y, int stride, int off) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = G.nx * G.ny * G.nz;
  if (c >= C) return;
  const int i = c % G.nx, j = (c / G.nx) % G.ny, k = c / (G.nx * G.ny);
  const double xc = x[(size_t)c * stride + off];
  double acc = xc;
  if (i > 0) acc += G.wx * (xc - x[(size_t)(c - 1) * stride + off]);
  if (i < G.nx - 1) acc += G.wx * (xc - x[(size_t)(c + 1) * stride + off]);
  if (j > 0) acc += G.wy * (xc - x[(size_t)(c - G.nx) * stride + off]);
  if (j < G.ny - 1) acc += G.wy * (xc - x[(size_t)(c + G.nx) * stride + off]);
  if (k > 0) acc += G.wz * (xc - x[(size_t)(c - G.nx * G.ny) * stride + off]);
  if (k < G.nz - 1) acc += G.wz * (xc - x[(size_t)(c + G.nx * G.ny) * stride + off]);
  y[c] = acc;
}

__device__ __forceinline__ double smooth_diag(const SmoothGrid &G, int c) {
  if (G.rect) return G.diag[c];
  const int i = c % G.nx, j = (c / G.nx) % G.ny, k = c / (G.nx * G.ny);
  double d = 1.0;
  d += G.wx * ((i > 0) + (i < G.nx - 1)) + G.wy * ((j > 0) + (j < G.ny - 1)) + G.wz * ((k > 0) + (k < G.nz - 1));
  return d;
}

// deterministic dot product: per-block partial sums in a fixed tree, then one block adds the partials in order
__global__ void __launch_bounds__(256) k_dot_partial(const double *a, int sa, int oa, const double *b, int sb, int ob, int n, double *partial,
                                                     const double *w) {   // w: optional cell weights (volumes)
  __shared__ double sm[256];
  double acc = 0.0;
  if (w) { for (int c = blockIdx.x * 256 + threadIdx.x; c < n; c += gridDim.x * 256) acc += w[c] * (a[(size_t)c * sa + oa] * b[(size_t)c * sb + ob]); }
  else
  for (int c = blockIdx.x * 256 + threadIdx.x; c < n; c += gridDim.x * 256) acc += a[(size_t)c * sa + oa] * b[(size_t)c * sb + ob];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) k_dot_final(const double *partial, int nb, double *out) {
  __shared__ double sm[256];
  double acc = 0.0;
  for (int c = threadIdx.x; c < nb; c += 256) acc += partial[c];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) *out = sm[0];
}

// CG updates; the scalars live on the device so the loop needs no host round trip except the convergence check
// s[0] = rz, s[1] = pAp, s[2] = rz_new, s[3] = rr
__global__ void k_cg_init(SmoothGrid G, const double *x, int stride, int off, const double *Ax, double *r, double *z, double *p, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double rc = x[(size_t)c * stride + off] - Ax[c];   // b == x (initial guess = right-hand side)
  r[c] = rc;
  const double zc = rc / smooth_diag(G, c);
  z[c] = zc; p[c] = zc;
}
__global__ void k_cg_step1(double *x, int stride, int off, double *r, const double *p, const double *Ap, const double *s, SmoothGrid G, double *z, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double alpha = s[0] / s[1];
  x[(size_t)c * stride + off] += alpha * p[c];
  const double rc = r[c] - alpha * Ap[c];
  r[c] = rc;
  z[c] = rc / smooth_diag(G, c);
}
__global__ void k_cg_step2(double *p, const double *z, double *s, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double beta = s[2] / s[0];
  p[c] = z[c] + beta * p[c];
}
__global__ void k_cg_shift(double *s) { s[0] = s[2]; }

// field scaling used around the smoothing of Uf and Asrc: phi *= (1 - gamma) / phi /= (1 - gamma)  (:407-416, 675-690)
__global__ void k_scale_one_minus_gamma(int C, const double *gamma, double *f, int ncomp, int divide) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double w = 1 - gamma[c];
  for (int k = 0; k < ncomp; k++) { double a = f[(size_t)c * ncomp + k]; a = divide ? a / w : a * w; f[(size_t)c * ncomp + k] = a; }
}
// gamma /= V ; Ue /= V   and   Ue /= gamma where gamma > ROOTVSMALL : the two halves of k_finalize_alpha_u (:932-962)
__global__ void k_alpha_u_divV(int C, const double *cellV, double *gamma, double *Ue) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double V = cellV[c];
  gamma[c] = gamma[c] / V;
  for (int k = 0; k < 3; k++) Ue[3 * (size_t)c + k] = Ue[3 * (size_t)c + k] / V;
}
__global__ void k_alpha_u_divgamma(int C, const double *gamma, double *Ue) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double g = gamma[c];
  if (g > 1.0e-150) for (int k = 0; k < 3; k++) Ue[3 * (size_t)c + k] = Ue[3 * (size_t)c + k] / g;
}

}  // namespace sedi
