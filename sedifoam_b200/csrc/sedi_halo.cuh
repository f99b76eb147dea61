// sedi_halo.cuh -- device side of the multi-GPU brick decomposition: migration of owned particles at a neighbour
// rebuild, construction of the ghost (border) send lists, and the per-sub-step ghost refresh.
//
// Restates what LAMMPS' Comm does for sediFoam's inputs (`processors Px Py Pz`, `communicate single vel yes`,
// `newton off`): exchange() + borders() at every re-neighbouring, forward_comm() of
// x, v, omega of the border atoms every step; newton off => no reverse communication.  NVSwitch gives every GPU
// full bandwidth to every peer, so a brick talks to all (up to 26) neighbours directly instead of LAMMPS' three
// staged x/y/z sweeps; corner and edge ghosts are sent by their owner, nothing is forwarded.
#pragma once
#include "sedi_device.cuh"

namespace sedi {

static const int MIG_MAXH = 16;  // contact-history entries carried by a migrating particle

struct DecompDev {
  int grid[3], coord[3], periodic[3];
  double boxlo[3], boxhi[3], sublo[3], subhi[3], cutghost;
  int nlinks;
  int off[MAX_LINKS][3];
  double shift[MAX_LINKS][3];
  int linkof[27];  // (ox+1) + 3 (oy+1) + 9 (oz+1) -> link index, -1 if there is no such neighbour
};

// owner brick coordinate of a (wrapped) position: the ONE definition used by loading, migration and tests
__host__ __device__ inline int owner_coord(double x, double lo, double hi, int g) {
  if (g == 1) return 0;
  int c = (int)floor((x - lo) * g / (hi - lo));
  return c < 0 ? 0 : (c >= g ? g - 1 : c);
}

// ---- migration ------------------------------------------------------------------------------------------------------
// leave[i] = 1 + link for owned rows whose owner brick is no longer this one; per-link counters
__global__ void k_mig_classify(const D4 *posr, const D4 *omgt, int n, DecompDev D, int *leave, int *count, int cap, int *rows, int *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  leave[i] = 0;
  const unsigned long long b = (unsigned long long)__double_as_longlong(omgt[i].w);
  if (bits_flags(b) & PFLAG_GHOST) return;
  const D4 p = posr[i];
  const double x[3] = {p.x, p.y, p.z};
  int code = 0, mul = 1;
  bool stay = true;
  for (int d = 0; d < 3; d++) {
    int df = owner_coord(x[d], D.boxlo[d], D.boxhi[d], D.grid[d]) - D.coord[d];
    if (D.periodic[d] && D.grid[d] > 2) { if (df == D.grid[d] - 1) df = -1; else if (df == -(D.grid[d] - 1)) df = 1; }
    if (df < -1 || df > 1) { atomicOr(err, 2); df = df < 0 ? -1 : 1; }
    if (df != 0) stay = false;
    code += (df + 1) * mul; mul *= 3;
  }
  if (stay) return;
  const int L = D.linkof[code];
  if (L < 0) { atomicOr(err, 4); return; }
  const int slot = atomicAdd(&count[L], 1);
  if (slot >= cap) { atomicOr(err, 8); return; }
  rows[L * cap + slot] = i;
  leave[i] = 1 + L;
}

struct MigPlanes {
  int nwalls, npad, rec;           // rec = doubles per record
  int npl;                         // planes carried: 12, or 16 with the history-force state
  const D4 *posr, *velm, *omgt;
  const double *pl[16];            // fdrag, dudt, vold, uold (3 each) [+ sumDeltaFb xyz, n0: particleHistoryForce state, softParticle.H:104-107]
  const double *ws[MAX_WALLS * 3];
  const int *foam; const unsigned *wmask;
  const int *nn; const unsigned *nbr; const unsigned long long *tmask; const D4 *shear;  // old list (may be null)
  int *tag2idx; int maxtag;
};

// record layout (doubles): [0..11] quads, [12..12+npl) planes, then 3w wall shear, {foam|wmask}, nhist, MIG_MAXH x {tag, sx, sy, sz}
__global__ void k_mig_pack(int nlinks, int cap, const int *count, const int *rows, MigPlanes M, double *out, int *err) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int L = t / cap, k = t % cap;
  if (L >= nlinks || k >= count[L]) return;
  const int i = rows[L * cap + k];
  double *r = out + ((size_t)L * cap + k) * M.rec;
  const D4 p = M.posr[i], v = M.velm[i], w = M.omgt[i];
  r[0] = p.x; r[1] = p.y; r[2] = p.z; r[3] = p.w; r[4] = v.x; r[5] = v.y; r[6] = v.z; r[7] = v.w; r[8] = w.x; r[9] = w.y; r[10] = w.z; r[11] = w.w;
  for (int q = 0; q < M.npl; q++) r[12 + q] = M.pl[q][i];
  int o = 12 + M.npl;
  for (int q = 0; q < 3 * M.nwalls; q++) r[o++] = M.ws[q][i];
  long long iw = ((long long)(unsigned)M.foam[i]) | ((long long)M.wmask[i] << 32);
  r[o++] = __longlong_as_double(iw);
  int nh = 0;
  double *h = r + o + 1;
  if (M.nn) {
    const unsigned long long tm = M.tmask[i];
    const int nni = M.nn[i];
    for (int s = 0; s < nni; s++) {
      if (!((tm >> s) & 1ull)) continue;
      if (nh >= MIG_MAXH) { atomicOr(err, 16); break; }
      const size_t slot = (size_t)s * M.npad + i;
      const int j = (int)(M.nbr[slot] & NB_IDX_MASK);
      const int tj = bits_tag((unsigned long long)__double_as_longlong(M.omgt[j].w));
      const D4 sh = M.shear[slot];
      h[4 * nh] = __longlong_as_double((long long)tj); h[4 * nh + 1] = sh.x; h[4 * nh + 2] = sh.y; h[4 * nh + 3] = sh.z;
      nh++;
    }
  }
  r[o] = __longlong_as_double((long long)nh);
  const int tg = bits_tag((unsigned long long)__double_as_longlong(w.w));
  if (tg >= 0 && tg <= M.maxtag) M.tag2idx[tg] = -1;
}

struct MigDst {
  int nwalls, rec, npl;
  D4 *posr, *velm, *omgt;
  double *pl[16];
  double *ws[MAX_WALLS * 3];
  int *foam; unsigned *wmask; int *leave;
  int *arr_nh, *arr_tag; D4 *arr_shear;
};

__global__ void k_mig_unpack(int narr, const double *in, int row0, MigDst M) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= narr) return;
  const double *r = in + (size_t)k * M.rec;
  const int i = row0 + k;
  D4 p, v, w;
  p.x = r[0]; p.y = r[1]; p.z = r[2]; p.w = r[3]; v.x = r[4]; v.y = r[5]; v.z = r[6]; v.w = r[7]; w.x = r[8]; w.y = r[9]; w.z = r[10]; w.w = r[11];
  M.posr[i] = p; M.velm[i] = v; M.omgt[i] = w;
  for (int q = 0; q < M.npl; q++) M.pl[q][i] = r[12 + q];
  int o = 12 + M.npl;
  for (int q = 0; q < 3 * M.nwalls; q++) M.ws[q][i] = r[o++];
  const long long iw = __double_as_longlong(r[o++]);
  M.foam[i] = (int)(unsigned)(iw & 0xFFFFFFFFll); M.wmask[i] = (unsigned)(iw >> 32);
  M.leave[i] = 0;
  const int nh = (int)__double_as_longlong(r[o]);
  const double *h = r + o + 1;
  M.arr_nh[k] = nh;
  for (int m = 0; m < nh; m++) {
    M.arr_tag[k * MIG_MAXH + m] = (int)__double_as_longlong(h[4 * m]);
    D4 s; s.x = h[4 * m + 1]; s.y = h[4 * m + 2]; s.z = h[4 * m + 3]; s.w = 0.0;
    M.arr_shear[k * MIG_MAXH + m] = s;
  }
}

// ---- borders: deterministic (row-ordered) send lists for every link ------------------------------------------------
__device__ __forceinline__ unsigned border_mask(const D4 &p, const DecompDev &D) {
  const double x[3] = {p.x, p.y, p.z};
  int lo[3], hi[3];
  for (int d = 0; d < 3; d++) { lo[d] = x[d] < D.sublo[d] + D.cutghost; hi[d] = x[d] >= D.subhi[d] - D.cutghost; }
  unsigned m = 0;
  for (int L = 0; L < D.nlinks; L++) {
    bool in = true;
    for (int d = 0; d < 3; d++) { const int o = D.off[L][d]; if (o < 0) in = in && lo[d]; else if (o > 0) in = in && hi[d]; }
    if (in) m |= (1u << L);
  }
  return m;
}

__global__ void __launch_bounds__(256) k_border_count(const D4 *posr, int n, DecompDev D, int nblocks, int *blockcount) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned m = (i < n) ? border_mask(posr[i], D) : 0u;
  for (int L = 0; L < D.nlinks; L++) {
    const int c = __syncthreads_count((m >> L) & 1u);
    if (threadIdx.x == 0) blockcount[L * nblocks + blockIdx.x] = c;
  }
}

__global__ void __launch_bounds__(256) k_border_fill(const D4 *posr, int n, DecompDev D, int nblocks, const int *blockoff, int *sendrows) {
  __shared__ int wsum[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned m = (i < n) ? border_mask(posr[i], D) : 0u;
  for (int L = 0; L < D.nlinks; L++) {
    const bool f = (m >> L) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wsum[wid] = __popc(bal);
    __syncthreads();
    int base = blockoff[L * nblocks + blockIdx.x];
    for (int w = 0; w < wid; w++) base += wsum[w];
    if (f) sendrows[base + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
  }
}

struct HaloTable { int nlinks; int base[MAX_LINKS + 1]; double shift[MAX_LINKS][3]; };

// forward communication, sender side: border rows -> contiguous records {posr + shift, velm, omgt | GHOST}
__global__ void k_halo_pack(const D4 *posr, const D4 *velm, const D4 *omgt, const int *sendrows, HaloTable H, D4 *out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H.base[H.nlinks]) return;
  int L = 0;
  while (L + 1 < H.nlinks && e >= H.base[L + 1]) L++;
  const int i = sendrows[e];
  D4 p = posr[i], v = velm[i], w = omgt[i];
  p.x = p.x + H.shift[L][0]; p.y = p.y + H.shift[L][1]; p.z = p.z + H.shift[L][2];
  unsigned long long b = (unsigned long long)__double_as_longlong(w.w);
  b |= ((unsigned long long)PFLAG_GHOST) << 56;
  w.w = __longlong_as_double((long long)b);
  out[3 * (size_t)e] = p; out[3 * (size_t)e + 1] = v; out[3 * (size_t)e + 2] = w;
}

// ---- fused pack + send over NVLink peer memory -----------------------------------------------------------------------
// The ghost rows of every neighbour brick are mapped into this process (CUDA IPC); the border rows are written
// straight into them (posr + shift, velm, omgt | GHOST), no staging buffer, no NCCL call, no unpack kernel.
// row -> entries table of the fused ghost refresh (built once per rebuild from the per-link send lists)
__global__ void k_border_rowcount(const int *sendrows, int total, int *cnt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < total) atomicAdd(&cnt[sendrows[e]], 1);
}
__global__ void k_border_rowfill(const int *sendrows, HaloTable H, const int *bpos, int *fill, BorderEnt *bent) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H.base[H.nlinks]) return;
  int L = 0;
  while (L + 1 < H.nlinks && e >= H.base[L + 1]) L++;
  const int i = sendrows[e];
  BorderEnt b; b.link = L; b.pos = e - H.base[L];
  bent[bpos[i] + atomicAdd(&fill[i], 1)] = b;   // order inside a row is irrelevant: every entry has its own destination
}
__global__ void k_border_pack_cnt(int n, const int *cnt, unsigned char *bcnt, int *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = cnt[i];
  if (c > 255) atomicOr(err, 32);
  bcnt[i] = (unsigned char)c;
}
__global__ void k_halo_push(const D4 *posr, const D4 *velm, const D4 *omgt, const int *sendrows, const __grid_constant__ PushTable H) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H.base[H.nlinks]) return;
  int L = 0;
  while (L + 1 < H.nlinks && e >= H.base[L + 1]) L++;
  const int i = sendrows[e];
  D4 p = posr[i], v = velm[i], w = omgt[i];
  p.x = p.x + H.shift[L][0]; p.y = p.y + H.shift[L][1]; p.z = p.z + H.shift[L][2];
  unsigned long long b = (unsigned long long)__double_as_longlong(w.w);
  b |= ((unsigned long long)PFLAG_GHOST) << 56;
  w.w = __longlong_as_double((long long)b);
  const int r = H.rstart[L] + (e - H.base[L]);
  H.rposr[L][r] = p; H.rvelm[L][r] = v; H.romgt[L][r] = w;
  __threadfence_system();
}

// all-ranks barrier + rebuild-flag consensus through peer memory: rank r's slot [me] of every rank's signal array
// receives (epoch << 32 | my flag); then every rank waits until all its slots carry this epoch and takes the max flag.
struct SignalTable { int nranks, me; unsigned long long *rsig[64]; };
// The epoch is a device-resident counter (mysig[64], never written by a peer): all ranks run the same sequence of
// exchanges, so their counters agree, and the launch has no per-call argument -- a chunk of sub-steps including its halo
// kernels is one CUDA graph on several GPUs as well.
__global__ void k_halo_signal_wait(const __grid_constant__ SignalTable S, volatile unsigned long long *mysig, int *ctrl, int with_flag) {
  __shared__ int sflag[64];
  const int r = threadIdx.x;
  const unsigned epoch = (unsigned)mysig[64] + 1u;
  if (r < S.nranks) {
    const unsigned fl = with_flag ? (unsigned)ctrl[0] : 0u;
    // release: the border rows this rank wrote into its neighbours' memory (earlier kernels of this stream, ordered before this
    // thread by the grid boundary) become visible system-wide before the signal does -- fence cumulativity, PTX memory model
    __threadfence_system();
    *((volatile unsigned long long *)&S.rsig[r][S.me]) = ((unsigned long long)epoch << 32) | fl;
    unsigned long long v;
    do { v = mysig[r]; } while ((unsigned)(v >> 32) < epoch);
    sflag[r] = (int)(unsigned)(v & 0xffffffffull);
  }
  __syncthreads();
  if (r == 0) {
    mysig[64] = epoch;
    if (with_flag) {
      int m = 0;
      for (int k = 0; k < S.nranks; k++) m = max(m, sflag[k]);
      ctrl[0] = m;
    }
  }
  __threadfence_system();
}

// receiver side: records -> ghost rows [row0, row0 + nghost)
__global__ void k_halo_unpack(const D4 *in, int nghost, int row0, D4 *posr, D4 *velm, D4 *omgt) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  posr[row0 + g] = in[3 * (size_t)g]; velm[row0 + g] = in[3 * (size_t)g + 1]; omgt[row0 + g] = in[3 * (size_t)g + 2];
}

}  // namespace sedi

namespace sedi {
// Domain::pbc for owned rows (multi-GPU: done before the owner brick of a particle is evaluated)
__global__ void k_pbc_wrap(D4 *posr, const D4 *omgt, int n, int per0, int per1, int per2, double lo0, double lo1, double lo2, double hi0,
                           double hi1, double hi2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long b = (unsigned long long)__double_as_longlong(omgt[i].w);
  if (bits_flags(b) & PFLAG_GHOST) return;
  D4 p = posr[i];
  bool ch = false;
  if (per0) { const double prd = hi0 - lo0; if (p.x < lo0) { p.x += prd; ch = true; } if (p.x >= hi0) { p.x -= prd; p.x = fmax(p.x, lo0); ch = true; } }
  if (per1) { const double prd = hi1 - lo1; if (p.y < lo1) { p.y += prd; ch = true; } if (p.y >= hi1) { p.y -= prd; p.y = fmax(p.y, lo1); ch = true; } }
  if (per2) { const double prd = hi2 - lo2; if (p.z < lo2) { p.z += prd; ch = true; } if (p.z >= hi2) { p.z -= prd; p.z = fmax(p.z, lo2); ch = true; } }
  if (ch) posr[i] = p;
}
}  // namespace sedi
