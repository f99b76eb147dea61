// sedi_sell.cuh -- the sorted-row form of the fused DEM sub-step kernel (k_step_sell) and the row ordering it runs on.
//
// Same work per launch as k_step (sedi_step.cuh): force_clear -> pair->compute -> post_force fixes -> nve/sphere, same
// reference arithmetic (interfaceToLammps/pair_gran_hertzFix_history.cpp:120-285 and the fixes listed in
// sedi_step.cuh), same per-particle summation order (ascending list slot), hence the same results bit for bit.
//
// Why: on a random packing the rows are ragged (2..14 list entries, 0..9 of them overlapping).  The slot walk of k_step
// runs every warp for its longest row and executes the 290-instruction contact law whenever ANY lane overlaps at a
// slot (45 % lane utilisation there); the warp-queue kernel (sedi_wq.cuh) repairs the utilisation by redistributing
// contacts over the lanes through shared memory, but the L1 data pipe -- one wavefront per cycle per SM for shared
// memory AND for every scattered 32-byte gather -- then becomes the limiter (74 % busy, profiles/r02_*).  The cure used
// here is the one of sliced-ELLPACK sparse formats (SELL-C-sigma): keep one lane per particle and no exchange at all,
// and instead ORDER THE ROWS so that the 32 rows of a warp have the same amount of work:
//
//   ordering  at every neighbour rebuild the bin-ordered rows are stably re-sorted inside windows of SELL_WINDOW rows by
//             (overlapping entries, list entries) of the previous list (k_window_sort).  A window is a compact patch of
//             the bed, so the locality of the gathers is kept; the bins are reached through an index list (crow) during
//             the list build, so cells stay contiguous in the canonical order.  On the benchmark bed the longest lane of
//             a warp then has 5.2 overlapping entries against a mean of 4.96 (unsorted: 7.6);
//   phase 1   one lane per particle: list words requested with the particle's own state, partner positions gathered
//             eight at a time, distance TESTED only -> touch mask of the row;
//   phase 2   the lane walks its own overlapping entries in slot order: partner position / velocity / spin and the
//             history quad are gathered (prefetched one contact ahead), the contact law runs with every lane of the
//             warp busy, force / torque accumulate in registers.  No shared memory, no atomics, bitwise deterministic;
//   epilogue  step_epilogue<> (fixes in script order, final + initial integrate, skin/2 trigger).
#pragma once
#include "sedi_step.cuh"

namespace sedi {

#ifndef SEDI_SELL_THREADS
#define SEDI_SELL_THREADS 64
#endif
#ifndef SEDI_SELL_MINB
#define SEDI_SELL_MINB 8
#endif
#ifndef SEDI_SELL_PF
#define SEDI_SELL_PF 1     // prefetch the next overlapping entry's partner lines / history quad to L1
#endif
static const int SELL_WINDOW = 512;   // sigma of SELL-C-sigma: rows are sorted by work inside windows of this many rows

// ---- row ordering: stable sort of every window of bin-ordered rows by the work the row had under the previous list ----
// order[k]  : canonical (bin-ordered) position k -> old row            (input, from the counting sort)
// order2[r] : physical new row r -> old row                             (what the permutation kernels and the history re-attachment use)
// crow[k]   : canonical position k -> physical new row                  (how the list build reaches the rows of a bin)
__global__ void __launch_bounds__(SELL_WINDOW) k_window_sort(int n, const int *order, const unsigned long long *tmask_old, const int *nn_old,
                                                             int nrows_old, int *order2, int *crow) {
  __shared__ unsigned s[SELL_WINDOW];
  const int t = threadIdx.x;
  const int k = blockIdx.x * SELL_WINDOW + t;
  unsigned comp = 0xFFFFFFFFu;
  if (k < n) {
    const int o = order[k];
    unsigned key = 0u;
    if (tmask_old && o < nrows_old) {
      const unsigned a = (unsigned)__popcll(tmask_old[o]), b = (unsigned)nn_old[o];
      key = (a < 63u ? a : 63u) * 64u + (b < 63u ? b : 63u);
    }
    comp = ((4095u - key) << 10) | (unsigned)t;   // heavy rows first; ties keep the bin order (the composite is unique)
  }
  s[t] = comp;
  __syncthreads();
  for (int size = 2; size <= SELL_WINDOW; size <<= 1) {   // bitonic sort, ascending
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int p = t ^ stride;
      if (p > t) {
        const unsigned a = s[t], b = s[p];
        const bool up = ((t & size) == 0);
        if ((a > b) == up) { s[t] = b; s[p] = a; }
      }
      __syncthreads();
    }
  }
  const unsigned c = s[t];
  if (c != 0xFFFFFFFFu) {
    const int kc = blockIdx.x * SELL_WINDOW + (int)(c & 1023u);
    order2[k] = order[kc];
    crow[kc] = k;
  }
}

// ---- the kernel -------------------------------------------------------------------------------------------------------
template <int PAIR, bool PBC>
__global__ void __launch_bounds__(SEDI_SELL_THREADS, SEDI_SELL_MINB) k_step_sell(const __grid_constant__ StepParams P, const int seq) {
  constexpr bool HIST = (PAIR == PAIR_HERTZFIX_HISTORY || PAIR == PAIR_HOOKE_HISTORY);
  if (P.mode != MODE_SETUP) {
    const int fl = *(volatile int *)&P.ctrl[0];
    if (fl != 0 && fl < seq) return;  // an earlier step of this chunk asked for a neighbour rebuild: become a no-op
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && P.mode != MODE_SETUP) atomicAdd(&P.ctrl[1], 1);
  if (i >= P.n) return;

  // ---- own row, list words of the first 16 slots (twelve unconditionally: they depend on nothing; rows have >= 16 slots)
  __shared__ unsigned s_e[16][SEDI_SELL_THREADS];   // the row's list words, kept for phase 2 (one column per lane: conflict-free)
  const int tid = threadIdx.x;
  D4 pi = ldg_d4_stream(&P.posr_in[i]);
  D4 vi = ldg_d4_stream(&P.velm_in[i]);
  D4 wi = ldg_d4_stream(&P.omgt_in[i]);
  const int nni = ld_nc_s32(&P.nn[i]);
  const unsigned long long tm_old = HIST ? P.tmask[i] : 0ull;
  unsigned e16[16];
#pragma unroll
  for (int k = 0; k < 12; k++) e16[k] = ld_nc_u32(&P.nbr[(size_t)k * P.npad + i]);
#pragma unroll
  for (int k = 12; k < 16; k++) e16[k] = (k < nni) ? ld_nc_u32(&P.nbr[(size_t)k * P.npad + i]) : 0u;
  // per-particle inputs of the epilogue: start them towards L1 now, read them after the sweep (one request per 32-byte sector)
  if ((tid & 3) == 0) {
    if (P.has_fdrag) { prefetch_l1(&P.fdrag[0][i]); prefetch_l1(&P.fdrag[1][i]); prefetch_l1(&P.fdrag[2][i]); }
    if (P.mode == MODE_FUSED) { prefetch_l1(&P.xhold[0][i]); prefetch_l1(&P.xhold[1][i]); prefetch_l1(&P.xhold[2][i]); }
  }
  if ((tid & 7) == 0 && P.wmask) prefetch_l1(&P.wmask[i]);
  if (bits_flags((unsigned long long)__double_as_longlong(wi.w)) & PFLAG_GHOST) return;  // ghost rows are refreshed by the halo exchange
  const double radi = pi.w, mi = vi.w;
  const int maski = bits_mask((unsigned long long)__double_as_longlong(wi.w));
#pragma unroll
  for (int k = 0; k < 16; k++) { if (k >= nni) e16[k] = 0u; s_e[k][tid] = e16[k]; }
#pragma unroll
  for (int k = 8; k < 16; k++) if (k < nni) prefetch_l1(&P.posr_in[e16[k] & NB_IDX_MASK]);   // second gather batch: lines on their way while the first is tested

  // ---- phase 1: which list entries overlap (pair :131 `rsq >= radsum*radsum` -> no contact)
  unsigned long long touch = 0ull;
  auto test_entry = [&](const unsigned e, const D4 &pj_in, const int s) {
    if (!(e & NB_FLAG_GRAN)) return;
    D4 pj = pj_in;
    const int img = (int)((e >> NB_IMG_SHIFT) & 31u);
    if (PBC && img != NB_IMG_NONE) {  // periodic image = LAMMPS ghost: position is fl(x_j + shift)
      pj.x = pj.x + P.imgshift[img][0]; pj.y = pj.y + P.imgshift[img][1]; pj.z = pj.z + P.imgshift[img][2];
    }
    const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
    const double rsq = delx * delx + dely * dely + delz * delz;
    const double radsum = radi + pj.w;
    if (rsq < radsum * radsum) touch |= (1ull << s);
  };
#pragma unroll
  for (int b = 0; b < 16; b += 8) {
    if (b < nni) {
      D4 p8[8];
#pragma unroll
      for (int k = 0; k < 8; k++) p8[k] = ldg_d4(&P.posr_in[e16[b + k] & NB_IDX_MASK]);
#pragma unroll
      for (int k = 0; k < 8; k++) test_entry(e16[b + k], p8[k], b + k);
    }
  }
  for (int sb = 16; sb < nni; sb += 4) {   // long rows (large skin)
    unsigned e4[4];
    D4 p4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) e4[k] = (sb + k < nni) ? ld_nc_u32(&P.nbr[(size_t)(sb + k) * P.npad + i]) : 0u;
#pragma unroll
    for (int k = 0; k < 4; k++) p4[k] = ldg_d4(&P.posr_in[e4[k] & NB_IDX_MASK]);
#pragma unroll
    for (int k = 0; k < 4; k++) test_entry(e4[k], p4[k], sb + k);
  }

  // ---- phase 2: the overlapping entries, in slot order; the rows of a warp have (nearly) the same number of them
  const bool shearupdate = (P.mode != MODE_SETUP);
  HzCoef hc; hc.c_sn = P.c_sn; hc.c_ccel = P.c_ccel; hc.c_damp = P.c_damp; hc.c_kts = P.c_kts; hc.c_ctd = P.c_ctd; hc.c_ekt = P.c_ekt; hc.xmu = P.xmu;
  GranCoef gc; gc.kn = P.kn; gc.kt = P.kt; gc.gamman = P.gamman; gc.gammat = P.gammat; gc.xmu = P.xmu; gc.beta = P.beta;
  double fx = 0.0, fy = 0.0, fz = 0.0, tx = 0.0, ty = 0.0, tz = 0.0;   // pair accumulators (force_clear)
  unsigned long long m = touch;
  int s = 0;
  unsigned e = 0u;
  auto list_word = [&](const int sl) -> unsigned { return sl < 16 ? s_e[sl][tid] : ld_nc_u32(&P.nbr[(size_t)sl * P.npad + i]); };
  if (m) { s = __ffsll((long long)m) - 1; e = list_word(s); }
  while (m) {
    m &= m - 1;
    int sn = 0;
    unsigned en = 0u;
    if (m) { sn = __ffsll((long long)m) - 1; en = list_word(sn); }
    const size_t slot = (size_t)s * P.npad + i;
    const int j = (int)(e & NB_IDX_MASK);
    D4 pj = ldg_d4(&P.posr_in[j]);
    const D4 vj = ldg_d4(&P.velm_in[j]);
    const D4 wj = ldg_d4(&P.omgt_in[j]);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (HIST && ((tm_old >> s) & 1ull)) { const D4 h = ld_d4(&P.shear[slot]); s0 = h.x; s1 = h.y; s2 = h.z; }
#if SEDI_SELL_PF
    if (m) {   // the next contact's lines start towards L1 while this one is evaluated
      const int jn = (int)(en & NB_IDX_MASK);
      prefetch_l1(&P.posr_in[jn]); prefetch_l1(&P.velm_in[jn]); prefetch_l1(&P.omgt_in[jn]);
      if (HIST && ((tm_old >> sn) & 1ull)) prefetch_l1(&P.shear[(size_t)sn * P.npad + i]);
    }
#endif
    const int img = (int)((e >> NB_IMG_SHIFT) & 31u);
    if (PBC && img != NB_IMG_NONE) {
      pj.x = pj.x + P.imgshift[img][0]; pj.y = pj.y + P.imgshift[img][1]; pj.z = pj.z + P.imgshift[img][2];
    }
    const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
    const double rsq = delx * delx + dely * dely + delz * delz;
    const double radj = pj.w, mj = vj.w;
    const double radsum = radi + radj;
    const int maskj = bits_mask((unsigned long long)__double_as_longlong(wj.w));
    double meff = (PAIR == PAIR_HERTZFIX_HISTORY) ? div_nr(mi * mj, mi + mj) : (mi * mj) / (mi + mj);
    if (maski & P.freeze_groupbit) meff = mj;
    if (maskj & P.freeze_groupbit) meff = mi;
    const double vrx = vi.x - vj.x, vry = vi.y - vj.y, vrz = vi.z - vj.z;
    const double wsx = radi * wi.x + radj * wj.x, wsy = radi * wi.y + radj * wj.y, wsz = radi * wi.z + radj * wj.z;
    double fox, foy, foz, tox, toy, toz;
    if (PAIR == PAIR_HERTZFIX_HISTORY) {
      hertzfix_fast(delx, dely, delz, rsq, vrx, vry, vrz, wsx, wsy, wsz, meff, radsum, div_nr(radi * radj, radsum), hc, P.dtv, shearupdate,
                    s0, s1, s2, fox, foy, foz, tox, toy, toz);
    } else {
      V3 vr = {vrx, vry, vrz}, ws = {wsx, wsy, wsz}, sh = {s0, s1, s2}, fo, to;
      if (PAIR == PAIR_HOOKE_HISTORY) hooke_history_contact(delx, dely, delz, rsq, vr, ws, meff, radsum, gc, P.dtv, shearupdate, sh, fo, to);
      else hooke_contact(delx, dely, delz, rsq, vr, ws, meff, radsum, gc, fo, to);
      s0 = sh.x; s1 = sh.y; s2 = sh.z; fox = fo.x; foy = fo.y; foz = fo.z; tox = to.x; toy = to.y; toz = to.z;
    }
    if (HIST) { D4 h; h.x = s0; h.y = s1; h.z = s2; h.w = 0.0; st_d4(&P.shear[slot], h); }
    // reference: f[i] += F ; torque[i] -= radi * tor   (pair :259-271)
    fx += fox; fy += foy; fz += foz;
    tx -= radi * tox; ty -= radi * toy; tz -= radi * toz;
    s = sn; e = en;
  }
  if (HIST && touch != tm_old) P.tmask[i] = touch;
  double fd0 = 0.0, fd1 = 0.0, fd2 = 0.0, xh0 = 0.0, xh1 = 0.0, xh2 = 0.0;
  if (P.has_fdrag) { fd0 = ld_nc_f64(&P.fdrag[0][i]); fd1 = ld_nc_f64(&P.fdrag[1][i]); fd2 = ld_nc_f64(&P.fdrag[2][i]); }
  if (P.mode == MODE_FUSED) { xh0 = ld_nc_f64(&P.xhold[0][i]); xh1 = ld_nc_f64(&P.xhold[1][i]); xh2 = ld_nc_f64(&P.xhold[2][i]); }
  step_epilogue<PAIR, false>(P, i, seq, pi, vi, wi, fx, fy, fz, tx, ty, tz, 0.0, 0.0, 0.0, fd0, fd1, fd2, xh0, xh1, xh2, touch);
}

}  // namespace sedi
