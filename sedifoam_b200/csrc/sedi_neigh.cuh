// sedi_neigh.cuh -- neighbour rebuild on the GPU: periodic wrap, cell binning (counting sort), physical
// re-ordering of every per-particle array into cell order, directed ELL neighbour list, and re-attachment of
// contact history by partner tag.
//
// Semantics restated (EXTERNAL LAMMPS lammps-1Feb14; requested by
// `neighbor <skin> bin` + `neigh_modify delay 0` in every shipped in.lammps, e.g.
// cases/auto-testing/test-cases/xiaocase1/in.lammps:12-13):
//   * granular list: pair kept when rsq <= (ri + rj + skin)^2          (flag NB_FLAG_GRAN)
//   * type list    : pair kept when rsq <= (cut[ti][tj] + skin)^2      (flag NB_FLAG_TYPE; fix cohesive's half
//                    list, fix_cohesive.cpp:72-83, and lubricate/poly's full list, pair_lubricate_poly.cpp:463-465)
//   * row layout   : granular entries (possibly also in the type list) fill slots [0, cap) in stencil order -- they own a
//                    history slot and a bit of the 64-bit touch mask, so cap <= 64 --; entries that are ONLY in the type
//                    list (beyond ri + rj + skin) follow in slots [cap, cap + tcap) without any such limit
//   * history      : a new pair that overlaps (rsq < (ri+rj)^2) inherits the shear stored for the same partner
//                    TAG before the rebuild, otherwise starts from zero.
// The predicates are evaluated with exactly the reference's operation order and no FMA contraction, so the pair
// SET is bit-exact against the CPU oracle; the list is directed (each pair appears in both rows).
#pragma once
#include "sedi_device.cuh"

namespace sedi {

struct BuildParams {
  int n, npad, cap;              // cap: slots of the granular segment [0, cap) of a row (entries with NB_FLAG_GRAN, history, touch bit)
  int tcap;                      // slots of the type-only segment [cap, cap + tcap): entries that are only in the type-cut-off list
  int *nt;                       // type-only entries per row
  int want_gran, want_type, ntypes;
  const D4 *posr, *omgt;
  const int *cellstart;
  int nb[3], periodic[3], tile[2];
  double lo[3], inv[3], prd[3];
  double skin;
  double cutneighsq[(MAX_TYPES + 1) * (MAX_TYPES + 1)];
  unsigned *nbr;
  int *nn;
  unsigned long long *tmask;
  D4 *shear;
  int have_old, npad_old;
  const int *oldidx;
  const unsigned *nbr_old;
  const int *nn_old;
  const unsigned long long *tmask_old;
  const D4 *shear_old;
  const D4 *omgt_old;
  int nlocal_rows;               // first ghost row (ghost partners are reached through gorder)
  const int *gcellstart, *gorder;  // ghost rows binned by cell (null on a single GPU)
  const int *crow;               // canonical (bin-ordered) position -> row, when the rows are sorted by work inside windows (null: identity)
  int n_old;                     // rows of the old arrays; oldidx >= n_old marks a particle that migrated in
  const int *arr_nh, *arr_tag;   // history carried by migrated particles
  const D4 *arr_shear;
  int *maxcount;                 // [0] longest granular segment found
  int *maxcount_t;               // [0] longest type-only segment found
  unsigned long long *npairs;    // directed entries: [0] granular, [1] type list, [2] granular periodic-image, [3] type periodic-image
};

__device__ __forceinline__ int bin_coord(double x, double lo, double inv, int nb) {
  int c = (int)floor((x - lo) * inv);
  return c < 0 ? 0 : (c >= nb ? nb - 1 : c);
}

// Domain::pbc for owned particles + bin id + per-cell histogram.  (per*, blo*) describe the GLOBAL box for the wrap;
// B describes the bins (which cover the local sub-domain plus its ghost shell on a multi-GPU run).
__global__ void k_wrap_bin(D4 *posr, const D4 *omgt, int n, BinParams B, double hi0, double hi1, double hi2, double prd0,
                           double prd1, double prd2, int *cellid, int *cellcount, int per0, int per1, int per2, double blo0,
                           double blo1, double blo2, const int *leave, int trash_cell, int do_wrap, int *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  D4 p = posr[i];
  const unsigned long long b = (unsigned long long)__double_as_longlong(omgt[i].w);
  const bool ghost = (bits_flags(b) & PFLAG_GHOST) != 0;
  if (trash_cell >= 0 && (ghost || (leave && leave[i]))) {  // stale ghost rows and migrated-away rows drop out of the sort
    cellid[i] = trash_cell;
    atomicAdd(&cellcount[trash_cell], 1);
    return;
  }
  if (!ghost && do_wrap) {
    bool ch = false;
    if (per0) { if (p.x < blo0) { p.x += prd0; ch = true; } if (p.x >= hi0) { p.x -= prd0; p.x = fmax(p.x, blo0); ch = true; } }
    if (per1) { if (p.y < blo1) { p.y += prd1; ch = true; } if (p.y >= hi1) { p.y -= prd1; p.y = fmax(p.y, blo1); ch = true; } }
    if (per2) { if (p.z < blo2) { p.z += prd2; ch = true; } if (p.z >= hi2) { p.z -= prd2; p.z = fmax(p.z, blo2); ch = true; } }
    if (ch) posr[i] = p;
    // still outside a periodic box after one wrap, or not a number: the particle moved more than a box length since the last
    // rebuild -- the run has blown up (LAMMPS: "Out of range atoms - cannot compute").  Reported, not binned into one cell.
    if (err && ((per0 && !(p.x >= blo0 && p.x < hi0)) || (per1 && !(p.y >= blo1 && p.y < hi1)) || (per2 && !(p.z >= blo2 && p.z < hi2)) ||
                !(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z))) atomicOr(err, 4);
  }
  const int cx = bin_coord(p.x, B.lo[0], B.inv[0], B.nb[0]);
  const int cy = bin_coord(p.y, B.lo[1], B.inv[1], B.nb[1]);
  const int cz = bin_coord(p.z, B.lo[2], B.inv[2], B.nb[2]);
  const int c = cell_index(cx, cy, cz, B.nb, B.tile);
  cellid[i] = c;
  atomicAdd(&cellcount[c], 1);
}

// ---- exclusive scan over cell counts: 3 small kernels, 4096 items per block
static const int SCAN_ITEMS = 4096;
__global__ void __launch_bounds__(1024) k_scan_local(const int *in, int *out, int n, int *blocksum) {
  __shared__ int sm[32];
  const int base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 4;
  int v[4], t = 0;
  for (int k = 0; k < 4; k++) { v[k] = (base + k < n) ? in[base + k] : 0; t += v[k]; }
  int inc = t;
  for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += u; }
  if ((threadIdx.x & 31) == 31) sm[threadIdx.x >> 5] = inc;
  __syncthreads();
  if (threadIdx.x < 32) {
    int w = sm[threadIdx.x], wi = w;
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += u; }
    sm[threadIdx.x] = wi - w;
    if (threadIdx.x == 31) blocksum[blockIdx.x] = wi;
  }
  __syncthreads();
  int ex = inc - t + sm[threadIdx.x >> 5];
  for (int k = 0; k < 4; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}
__global__ void __launch_bounds__(1024) k_scan_sums(int *blocksum, int nblocks) {  // single block, serial over chunks
  __shared__ int sm[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int idx = base + threadIdx.x;
    const int v = idx < nblocks ? blocksum[idx] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += u; }
    if ((threadIdx.x & 31) == 31) sm[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = sm[threadIdx.x], wi = w;
      for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, wi, o); if (threadIdx.x >= o) wi += u; }
      sm[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const int ex = inc - v + sm[threadIdx.x >> 5] + carry;
    if (idx < nblocks) blocksum[idx] = ex;
    __syncthreads();
    if (threadIdx.x == 1023) carry = ex + v;
    __syncthreads();
  }
}
__global__ void k_scan_add(int *out, int n, const int *blocksum, int total_slot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += blocksum[i / SCAN_ITEMS];
  (void)total_slot;
}

__global__ void k_bin_scatter(const int *cellid, int n, const int *cellstart, int *cellfill, int *order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cellid[i];
  order[cellstart[c] + atomicAdd(&cellfill[c], 1)] = i;
}

// make the order inside each cell deterministic (ascending tag): one thread per cell, insertion sort
__global__ void k_cell_sort(const int *cellstart, int ncells, int ntot, int *order, const D4 *omgt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const int s = cellstart[c], e = (c + 1 < ncells) ? cellstart[c + 1] : ntot;
  for (int a = s + 1; a < e; a++) {
    const int oa = order[a];
    const unsigned ta = (unsigned)bits_tag((unsigned long long)__double_as_longlong(omgt[oa].w));
    int b = a - 1;
    while (b >= s) {
      const int ob = order[b];
      const unsigned tb = (unsigned)bits_tag((unsigned long long)__double_as_longlong(omgt[ob].w));
      if (tb <= ta) break;
      order[b + 1] = ob; b--;
    }
    order[b + 1] = oa;
  }
}

struct PlaneList { int nplanes; const double *src[48]; double *dst[48]; };

// gather the three state quads into cell order; record xhold, the tag map and the old<->new index maps
__global__ void k_permute_quads(const int *order, int n, const D4 *posr_s, const D4 *velm_s, const D4 *omgt_s, D4 *posr_d,
                                D4 *velm_d, D4 *omgt_d, double *xh0, double *xh1, double *xh2, int *tag2idx, int maxtag,
                                const unsigned *wmask_s, unsigned *wmask_d, const int *foam_s, int *foam_d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = order[i];
  const D4 p = posr_s[o], v = velm_s[o], w = omgt_s[o];
  posr_d[i] = p; velm_d[i] = v; omgt_d[i] = w;
  xh0[i] = p.x; xh1[i] = p.y; xh2[i] = p.z;
  const unsigned long long b = (unsigned long long)__double_as_longlong(w.w);
  const int t = bits_tag(b);
  if (!(bits_flags(b) & PFLAG_GHOST) && t >= 0 && t <= maxtag) tag2idx[t] = i;
  if (wmask_s) wmask_d[i] = wmask_s[o];
  if (foam_s) foam_d[i] = foam_s[o];
}
__global__ void k_permute_planes(const int *order, int n, PlaneList L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = order[i];
  for (int k = 0; k < L.nplanes; k++) L.dst[k][i] = L.src[k][o];
}

// directed ELL neighbour list + history re-attachment.  One thread per particle, 27-cell stencil.
__global__ void __launch_bounds__(128) k_build_list(const __grid_constant__ BuildParams B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned ng = 0, nt = 0, ngi = 0, nti = 0;
  int cnt = 0, cntT = 0;
  if (i < B.n) {
    const D4 pi = B.posr[i];
    const unsigned long long bi = (unsigned long long)__double_as_longlong(B.omgt[i].w);
    if (!(bits_flags(bi) & PFLAG_GHOST)) {
      const int ti = bits_type(bi);
      const double radi = pi.w;
      const int cx = bin_coord(pi.x, B.lo[0], B.inv[0], B.nb[0]);
      const int cy = bin_coord(pi.y, B.lo[1], B.inv[1], B.nb[1]);
      const int cz = bin_coord(pi.z, B.lo[2], B.inv[2], B.nb[2]);
      const int orow_raw = B.have_old ? B.oldidx[i] : -1;
      const int arrk = (orow_raw >= B.n_old && B.arr_nh) ? orow_raw - B.n_old : -1;  // migrated in at this rebuild
      const int orow = (B.nn_old && orow_raw >= 0 && orow_raw < B.n_old) ? orow_raw : -1;
      const int nno = (orow >= 0) ? B.nn_old[orow] : 0;
      const unsigned long long tmo = (orow >= 0) ? B.tmask_old[orow] : 0ull;
      const int narrh = (arrk >= 0) ? B.arr_nh[arrk] : 0;
      unsigned long long tm = 0ull;
      auto visit = [&](const int j, const int img, const int ix, const int iy, const int iz) {
        if (j == i && img == NB_IMG_NONE) return;
        D4 pj = B.posr[j];
        if (img != NB_IMG_NONE) { pj.x = pj.x + ix * B.prd[0]; pj.y = pj.y + iy * B.prd[1]; pj.z = pj.z + iz * B.prd[2]; }
        const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
        const double rsq = delx * delx + dely * dely + delz * delz;
        const double radsum = radi + pj.w;
        unsigned flags = 0;
        if (B.want_gran) { const double cs = (radsum + B.skin) * (radsum + B.skin); if (rsq <= cs) flags |= NB_FLAG_GRAN; }
        if (B.want_type) {
          const int tj = bits_type((unsigned long long)__double_as_longlong(B.omgt[j].w));
          if (rsq <= B.cutneighsq[ti * (MAX_TYPES + 1) + tj]) flags |= NB_FLAG_TYPE;
        }
        if (!flags) return;
        // LAMMPS stores an owned-ghost pair in both owners' lists; a periodic image inside one GPU is the same thing
        const bool isimg = (img != NB_IMG_NONE) || (j >= B.nlocal_rows);
        if (flags & NB_FLAG_GRAN) { ng++; if (isimg) ngi++; }
        if (flags & NB_FLAG_TYPE) { nt++; if (isimg) nti++; }
        if (!(flags & NB_FLAG_GRAN)) {   // type-only entry (fix cohesive / lubricate/poly beyond the granular cut-off): no history
          if (cntT < B.tcap) B.nbr[(size_t)(B.cap + cntT) * B.npad + i] = (unsigned)j | ((unsigned)img << NB_IMG_SHIFT) | flags;
          cntT++;
          return;
        }
        if (cnt < B.cap) {
          const size_t slot = (size_t)cnt * B.npad + i;
          B.nbr[slot] = (unsigned)j | ((unsigned)img << NB_IMG_SHIFT) | flags;
          if ((flags & NB_FLAG_GRAN) && rsq < radsum * radsum) {
            const int tagj = bits_tag((unsigned long long)__double_as_longlong(B.omgt[j].w));
            if (tmo) {
              for (int so = 0; so < nno; so++) {
                if (!((tmo >> so) & 1ull)) continue;
                const size_t oslot = (size_t)so * B.npad_old + orow;
                const int jo = (int)(B.nbr_old[oslot] & NB_IDX_MASK);
                if (bits_tag((unsigned long long)__double_as_longlong(B.omgt_old[jo].w)) == tagj) {
                  B.shear[slot] = B.shear_old[oslot];
                  tm |= (1ull << cnt);
                  break;
                }
              }
            }
            for (int m = 0; m < narrh; m++) {
              if (B.arr_tag[arrk * 16 + m] == tagj) { B.shear[slot] = B.arr_shear[arrk * 16 + m]; tm |= (1ull << cnt); break; }
            }
          }
        }
        cnt++;
      };
      for (int dz = -1; dz <= 1; dz++) {
        int bz = cz + dz, iz = 0;
        if (bz < 0) { if (!B.periodic[2]) continue; bz += B.nb[2]; iz = -1; }
        else if (bz >= B.nb[2]) { if (!B.periodic[2]) continue; bz -= B.nb[2]; iz = 1; }
        for (int dy = -1; dy <= 1; dy++) {
          int by = cy + dy, iy = 0;
          if (by < 0) { if (!B.periodic[1]) continue; by += B.nb[1]; iy = -1; }
          else if (by >= B.nb[1]) { if (!B.periodic[1]) continue; by -= B.nb[1]; iy = 1; }
          for (int dx = -1; dx <= 1; dx++) {
            int bx = cx + dx, ix = 0;
            if (bx < 0) { if (!B.periodic[0]) continue; bx += B.nb[0]; ix = -1; }
            else if (bx >= B.nb[0]) { if (!B.periodic[0]) continue; bx -= B.nb[0]; ix = 1; }
            const int c = cell_index(bx, by, bz, B.nb, B.tile);
            const int img = (ix + 1) + 3 * (iy + 1) + 9 * (iz + 1);
            const int js = B.cellstart[c], je = B.cellstart[c + 1];
            if (B.crow) { for (int k = js; k < je; k++) visit(B.crow[k], img, ix, iy, iz); }
            else for (int j = js; j < je; j++) visit(j, img, ix, iy, iz);
            if (B.gcellstart) {
              const int gs = B.gcellstart[c], ge = B.gcellstart[c + 1];
              for (int k = gs; k < ge; k++) visit(B.nlocal_rows + B.gorder[k], img, ix, iy, iz);
            }
          }
        }
      }
      B.nn[i] = cnt < B.cap ? cnt : B.cap;
      B.nt[i] = cntT < B.tcap ? cntT : B.tcap;
      B.tmask[i] = tm;
    } else {
      B.nn[i] = 0; B.nt[i] = 0; B.tmask[i] = 0ull;
    }
  }
  const unsigned full = 0xffffffffu;
  const int wmax = __reduce_max_sync(full, cnt);
  const int wmaxt = __reduce_max_sync(full, cntT);
  const unsigned sg = __reduce_add_sync(full, ng), st = __reduce_add_sync(full, nt);
  const unsigned sgi = __reduce_add_sync(full, ngi), sti = __reduce_add_sync(full, nti);
  if ((threadIdx.x & 31) == 0) {
    if (wmax > 0) atomicMax(B.maxcount, wmax);
    if (wmaxt > 0) atomicMax(B.maxcount_t, wmaxt);
    if (sg) atomicAdd(&B.npairs[0], (unsigned long long)sg);
    if (st) atomicAdd(&B.npairs[1], (unsigned long long)st);
    if (sgi) atomicAdd(&B.npairs[2], (unsigned long long)sgi);
    if (sti) atomicAdd(&B.npairs[3], (unsigned long long)sti);
  }
}

// export the directed list as (tag_i, tag_j, flags|img) rows for the parity tests
__global__ void k_export_pairs(int n, int npad, const int *nn, const int *nt, int hcap, const unsigned *nbr, const D4 *omgt, const int *rowstart,
                               int *ti, int *tj, unsigned *meta, const unsigned long long *tmask, const D4 *shear, int *touch,
                               double *shear_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int tagi = bits_tag((unsigned long long)__double_as_longlong(omgt[i].w));
  const int base = rowstart[i];
  const unsigned long long tm = tmask[i];
  const int ng = nn[i], ntot = ng + nt[i];
  for (int k = 0; k < ntot; k++) {
    const int s = k < ng ? k : hcap + (k - ng);   // granular segment, then the type-only segment
    const size_t slot = (size_t)s * npad + i;
    const unsigned e = nbr[slot];
    const int j = (int)(e & NB_IDX_MASK);
    ti[base + k] = tagi;
    tj[base + k] = bits_tag((unsigned long long)__double_as_longlong(omgt[j].w));
    meta[base + k] = e & ~NB_IDX_MASK;
    const int t = (k < ng) ? (int)((tm >> s) & 1ull) : 0;
    touch[base + k] = t;
    D4 h = {0, 0, 0, 0};
    if (t) h = shear[slot];
    shear_out[3 * (size_t)(base + k)] = h.x; shear_out[3 * (size_t)(base + k) + 1] = h.y; shear_out[3 * (size_t)(base + k) + 2] = h.z;
  }
}

// wall history of one wall as [n][3]; rows whose touch bit is clear are zero (the reference zeroes them eagerly,
// fix_wall_granFix.cpp:326-331; the engine only clears the bit)
__global__ void k_wall_shear_export(const double *s0, const double *s1, const double *s2, const unsigned *wmask, int w, int n, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool t = (wmask[i] >> w) & 1u;
  out[3 * (size_t)i] = t ? s0[i] : 0.0; out[3 * (size_t)i + 1] = t ? s1[i] : 0.0; out[3 * (size_t)i + 2] = t ? s2[i] : 0.0;
}

}  // namespace sedi
