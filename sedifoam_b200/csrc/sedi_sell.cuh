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
//   phase 1   one lane per particle: the row's sixteen list words are requested with the particle's own state and parked in shared
//             memory (one column per lane).  History styles: an entry that overlapped one sub-step ago goes to phase 2 UNTESTED (phase 2
//             repeats the reference's own test on the operands it gathers anyway); only the other entries are distance-tested here, four
//             partner positions in flight, and the first old contact's operands are requested in the same round trip;
//   phase 2   the lane walks its (presumably) overlapping entries in slot order.  prep() consumes the gathered partner position /
//             velocity / spin / history quad into 16 derived values, the NEXT entry's gathers are issued into the registers just
//             consumed, contact() runs the contact law on the derived values while those loads fly; force / torque accumulate in
//             registers.  No atomics, no inter-lane exchange, bitwise deterministic;
//   epilogue  step_epilogue<> (fixes in script order, final + initial integrate, skin/2 trigger, ghost refresh on several GPUs).
// Launch-uniform specialisations: M32 (rows of at most 16 granular slots) and StepParams::equal_spheres (one radius, one mass).
#pragma once
#include "sedi_step.cuh"

namespace sedi {

#ifndef SEDI_SELL_THREADS
#define SEDI_SELL_THREADS 64
#endif
#ifndef SEDI_SELL_MINB
#define SEDI_SELL_MINB 8
#endif
#ifndef SEDI_SELL_PFH
#define SEDI_SELL_PFH 1    // 1: the history quads of all slots that overlapped one sub-step ago are requested (L2) at the top of the kernel
#endif
#ifndef SEDI_SELL_SPEC
#define SEDI_SELL_SPEC 1   // 1 (history styles): phase 1 tests only the entries that did NOT overlap one sub-step ago; the entries that did go straight
                           // to phase 2, which repeats the overlap test with the operands it gathers anyway.  Halves the position gathers of phase 1,
                           // removes its second round trip, and the first contact's operands are requested together with the phase-1 positions
#endif
#ifndef SEDI_SELL_MONO
#define SEDI_SELL_MONO 1   // 1: equal-sphere pairs take meff = m / 2, reff = r / 2 directly
#endif
#ifndef SEDI_SELL_TRIM
#define SEDI_SELL_TRIM 1    // 1: all sixteen list words of the prologue are loaded unconditionally (rows have >= 16 slots)
#endif
#ifndef SEDI_SELL_P1B
#define SEDI_SELL_P1B 4   // partner positions in flight per phase-1 batch (SEDI_SELL_SPEC)
#endif
#ifndef SEDI_SELL_WINDOW
#define SEDI_SELL_WINDOW 1024
#endif
static const int SELL_WINDOW = SEDI_SELL_WINDOW;   // sigma of SELL-C-sigma: rows are sorted by work inside windows of this many rows (<= 1024)

// ---- row ordering: stable sort of every window of bin-ordered rows by the work the row had under the previous list ----
// order[k]  : canonical (bin-ordered) position k -> old row            (input, from the counting sort)
// order2[r] : physical new row r -> old row                             (what the permutation kernels and the history re-attachment use)
// crow[k]   : canonical position k -> physical new row                  (how the list build reaches the rows of a bin)
// several GPUs: rows whose particle lies within the ghost cut-off of a face shared with a neighbour brick (outside [lo, hi) in a split
// dimension; +-inf elsewhere) are the ones the sub-step kernel writes into the neighbour's memory.  They are sorted to the end of their
// window, i.e. into warps of their own, so that the other warps do not diverge into the push code of the epilogue
struct BorderBand { double lo[3], hi[3]; };
__global__ void __launch_bounds__(SELL_WINDOW) k_window_sort(int n, const int *order, const unsigned long long *tmask_old, const int *nn_old,
                                                             const int *nt_old, int nrows_old, int *order2, int *crow, const D4 *posr_old,
                                                             const BorderBand band) {
  __shared__ unsigned s[SELL_WINDOW];
  const int t = threadIdx.x;
  const int k = blockIdx.x * SELL_WINDOW + t;
  unsigned comp = 0xFFFFFFFFu;
  if (k < n) {
    const int o = order[k];
    unsigned key = 0u;
    if (tmask_old && o < nrows_old) {
      const unsigned a = (unsigned)__popcll(tmask_old[o]);
      if (nt_old) {   // type-cut-off list in use (fix cohesive / lubricate/poly): the row length dominates the work
        const unsigned b = ((unsigned)nn_old[o] + (unsigned)nt_old[o]) >> 1;
        key = (b < 63u ? b : 63u) * 64u + (a < 63u ? a : 63u);
      } else {
        const unsigned b = (unsigned)nn_old[o];
        key = (a < 63u ? a : 63u) * 64u + (b < 63u ? b : 63u);
      }
    }
    comp = ((4095u - key) << 10) | (unsigned)t;   // heavy rows first; ties keep the bin order (the composite is unique)
    if (posr_old) {
      const D4 p = posr_old[o];
      if (p.x < band.lo[0] || p.x >= band.hi[0] || p.y < band.lo[1] || p.y >= band.hi[1] || p.z < band.lo[2] || p.z >= band.hi[2]) comp |= (1u << 22);
    }
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

#ifndef SEDI_IMG_SEL
#define SEDI_IMG_SEL 0     // 1: periodic shift decoded arithmetically instead of an indexed (lane-divergent) constant-bank load
#endif
// periodic image = LAMMPS ghost: position is fl(x_j + shift), shift = (ix Lx, iy Ly, iz Lz), img = (ix+1) + 3 (iy+1) + 9 (iz+1)
__device__ __forceinline__ void apply_image(const StepParams &P, const int img, D4 &pj) {
#if SEDI_IMG_SEL
  const int cz = (img * 57) >> 9, r = img - 9 * cz, cy = (r * 11) >> 5, cx = r - 3 * cy;
  const double sx = cx == 0 ? -P.prd[0] : (cx == 2 ? P.prd[0] : 0.0);
  const double sy = cy == 0 ? -P.prd[1] : (cy == 2 ? P.prd[1] : 0.0);
  const double sz = cz == 0 ? -P.prd[2] : (cz == 2 ? P.prd[2] : 0.0);
  pj.x = pj.x + sx; pj.y = pj.y + sy; pj.z = pj.z + sz;
#else
  pj.x = pj.x + P.imgshift[img][0]; pj.y = pj.y + P.imgshift[img][1]; pj.z = pj.z + P.imgshift[img][2];
#endif
}

// contact / candidate masks of a row: 32-bit (M32) when every row of the list has at most 16 granular slots (launch-uniform; the usual
// case: a monodisperse row has <= 14 entries) -- all its list words are then the sixteen staged in shared memory --, 64-bit otherwise:
// half the bit-manipulation instructions of the pair sweep
template <bool M32> struct RowMask { typedef unsigned long long type; };
template <> struct RowMask<true> { typedef unsigned type; };
__device__ __forceinline__ int mask_ffs(unsigned m) { return __ffs((int)m) - 1; }
__device__ __forceinline__ int mask_ffs(unsigned long long m) { return __ffsll((long long)m) - 1; }

// ---- the kernel -------------------------------------------------------------------------------------------------------
// TYPELIST compiles the work of the type-cut-off list in (fix cohesive, pair lubricate/poly): a third walk over the row
// (both segments), see below -- doing it inside the distance-test walk was measured slower (configs[4]: 1478 instead of
// 1044 us per launch: eight partner positions plus the lubrication operands do not fit the registers); the plain
// granular instantiation carries none of it.
// one DEM sub-step of row i (one lane)
// TYPELIST: 0 none, 1 fix cohesive only, 2 pair lubricate/poly (and fix cohesive if present)
template <int PAIR, bool PBC, int TYPELIST, bool M32, class SkipF>
__device__ __forceinline__ void sell_row(const StepParams &P, const int seq, const int i, unsigned (*s_e)[SEDI_SELL_THREADS], SkipF skipf) {
  constexpr bool HIST = (PAIR == PAIR_HERTZFIX_HISTORY || PAIR == PAIR_HOOKE_HISTORY);
  typedef typename RowMask<M32>::type mask_t;
  constexpr int MBITS = M32 ? 32 : 64;
  const mask_t ONE = 1;
  // ---- own row, list words of the first 16 slots (twelve unconditionally: they depend on nothing; rows have >= 16 slots)
  const int tid = threadIdx.x;
  D4 pi = ldg_d4_stream(&P.posr_in[i]);
  D4 vi = ldg_d4_stream(&P.velm_in[i]);
  D4 wi = ldg_d4_stream(&P.omgt_in[i]);
  const int nni = ld_nc_s32(&P.nn[i]);
  const int nti = TYPELIST ? ld_nc_s32(&P.nt[i]) : 0;
  const mask_t tm_old = HIST ? (mask_t)P.tmask[i] : (mask_t)0;
  unsigned e16[16];
#pragma unroll
  for (int k = 0; k < 12; k++) e16[k] = ld_nc_u32(&P.nbr[(size_t)k * P.npad + i]);
#pragma unroll
  for (int k = 12; k < 16; k++) e16[k] = (SEDI_SELL_TRIM || k < nni) ? ld_nc_u32(&P.nbr[(size_t)k * P.npad + i]) : 0u;
  // per-particle inputs of the epilogue: start them towards L1 now, read them after the sweep (one request per 32-byte sector)
  if ((tid & 3) == 0) {
    if (P.has_fdrag) { prefetch_l1(&P.fdrag[0][i]); prefetch_l1(&P.fdrag[1][i]); prefetch_l1(&P.fdrag[2][i]); }
    if (P.mode == MODE_FUSED) { prefetch_l1(&P.xhold[0][i]); prefetch_l1(&P.xhold[1][i]); prefetch_l1(&P.xhold[2][i]); }
  }
  if ((tid & 7) == 0 && P.wmask) prefetch_l1(&P.wmask[i]);
  if ((tid & 31) == 0 && P.bcnt) prefetch_l1(&P.bcnt[i]);   // several GPUs: border-entry count of the row, read at the very end of the epilogue
#if SEDI_SELL_PFH == 1
  if (HIST) for (mask_t hm = tm_old; hm; hm &= hm - 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(&P.shear[(size_t)mask_ffs(hm) * P.npad + i]));
#elif SEDI_SELL_PFH == 2
  if (HIST) {   // unrolled and predicated for the first 16 slots (no bit scan, one pointer increment per slot), a loop for the rare longer rows
    const char *hp = (const char *)&P.shear[i];
    const size_t hst = (size_t)P.npad * sizeof(D4);
    const unsigned lo = (unsigned)tm_old;
#pragma unroll
    for (int k = 0; k < 16; k++) { if ((lo >> k) & 1u) asm volatile("prefetch.global.L2 [%0];" ::"l"(hp)); hp += hst; }
    for (mask_t hm = tm_old >> 16; hm; hm &= hm - 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(&P.shear[(size_t)(mask_ffs(hm) + 16) * P.npad + i]));
  }
#endif
  if (skipf()) return;   // no-op launch?  (decided here, behind the row's own loads, so that the flag's round trip is not exposed at the top of every warp)
  if (bits_flags((unsigned long long)__double_as_longlong(wi.w)) & PFLAG_GHOST) return;  // ghost rows are refreshed by the halo exchange
  const double radi = pi.w, mi = vi.w;
  const int maski = bits_mask((unsigned long long)__double_as_longlong(wi.w));
  // (the speculative flow never reads a list word at or beyond nni: no zero fill)
#pragma unroll
  for (int k = 0; k < 16; k++) { if (!(SEDI_SELL_TRIM && HIST && SEDI_SELL_SPEC) && k >= nni) e16[k] = 0u; s_e[k][tid] = e16[k]; }
  if (!(HIST && SEDI_SELL_SPEC)) {
#pragma unroll
    for (int k = 8; k < 16; k++) if (k < nni) prefetch_l1(&P.posr_in[e16[k] & NB_IDX_MASK]);   // second gather batch: lines on their way while the first is tested
  }

  // ---- the pieces of the pair sweep
  mask_t touch = 0;
  // distance test of one list entry (pair :131 `rsq >= radsum*radsum` -> no contact)
  auto test_entry = [&](const unsigned e, const D4 &pj_in, const int s, mask_t &mask) {
    if (!(e & NB_FLAG_GRAN)) return;
    D4 pj = pj_in;
    const int img = (int)((e >> NB_IMG_SHIFT) & 31u);
    if (PBC && img != NB_IMG_NONE) apply_image(P, img, pj);
    const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
    const double rsq = delx * delx + dely * dely + delz * delz;
    const double radsum = radi + pj.w;
    if (rsq < radsum * radsum) mask |= (ONE << s);
  };
  const bool shearupdate = (P.mode != MODE_SETUP);
  HzCoef hc; hc.c_sn = P.c_sn; hc.c_ccel = P.c_ccel; hc.c_damp = P.c_damp; hc.c_kts = P.c_kts; hc.c_ctd = P.c_ctd; hc.c_ekt = P.c_ekt; hc.xmu = P.xmu;
  GranCoef gc; gc.kn = P.kn; gc.kt = P.kt; gc.gamman = P.gamman; gc.gammat = P.gammat; gc.xmu = P.xmu; gc.beta = P.beta;
  double fx = 0.0, fy = 0.0, fz = 0.0, tx = 0.0, ty = 0.0, tz = 0.0;   // pair accumulators (force_clear)
  auto list_word = [&](const int sl) -> unsigned { return (M32 || sl < 16) ? s_e[sl][tid] : ld_nc_u32(&P.nbr[(size_t)sl * P.npad + i]); };   // M32: every row fits the staged sixteen
  // gathers of one overlapping entry: partner position / velocity / spin and the history quad of the slot
  struct Opnd { D4 pj, vj, wj; double h0, h1, h2; };
  auto gather = [&](const unsigned ew, const int sl, Opnd &o) {
    const int j = (int)(ew & NB_IDX_MASK);
    o.pj = ldg_d4(&P.posr_in[j]);
    o.vj = ldg_d4(&P.velm_in[j]);
    o.wj = ldg_d4(&P.omgt_in[j]);
    o.h0 = o.h1 = o.h2 = 0.0;
    if (HIST && ((tm_old >> sl) & ONE)) { const D4 h = ld_d4(&P.shear[(size_t)sl * P.npad + i]); o.h0 = h.x; o.h1 = h.y; o.h2 = h.z; }
  };
  // one entry in two stages.  prep(): everything that reads the gathered operands -- the reference's own overlap test (pair :131), relative
  // velocity, summed spin, effective mass / radius, old history -- into a set of derived values; false: the spheres do not overlap (an
  // entry that did one sub-step ago: its history is dropped with the bit).  contact(): contact law, history write-back, accumulation.
  // Between the two the operand registers are free, so the next entry's gathers are issued straight into them (phase 2 below).
  struct Drv { double delx, dely, delz, rsq, radsum, meff, reff, vrx, vry, vrz, wsx, wsy, wsz, s0, s1, s2; };
  auto prep = [&](const unsigned ew, const Opnd &o, Drv &d) -> bool {
    D4 pj = o.pj;
    const int img = (int)((ew >> NB_IMG_SHIFT) & 31u);
    if (PBC && img != NB_IMG_NONE) apply_image(P, img, pj);
    d.delx = pi.x - pj.x; d.dely = pi.y - pj.y; d.delz = pi.z - pj.z;
    d.rsq = d.delx * d.delx + d.dely * d.dely + d.delz * d.delz;
    const double radj = pj.w, mj = o.vj.w;
    d.radsum = radi + radj;
    const int maskj = bits_mask((unsigned long long)__double_as_longlong(o.wj.w));
    d.reff = 0.0;
    if (PAIR == PAIR_HERTZFIX_HISTORY) {
      // monodisperse system (launch-uniform, decided on the host for the whole system, so both directed evaluations of every
      // pair take the same branch): m m / (2 m) and r r / (2 r) without the two divisions
      if (SEDI_SELL_MONO && P.equal_spheres) { d.meff = 0.5 * mi; d.reff = 0.5 * radi; }
      else { d.meff = div_nr(mi * mj, mi + mj); d.reff = div_nr(radi * radj, d.radsum); }
    } else d.meff = (mi * mj) / (mi + mj);
    if (maski & P.freeze_groupbit) d.meff = mj;
    if (maskj & P.freeze_groupbit) d.meff = mi;
    d.vrx = vi.x - o.vj.x; d.vry = vi.y - o.vj.y; d.vrz = vi.z - o.vj.z;
    d.wsx = radi * wi.x + radj * o.wj.x; d.wsy = radi * wi.y + radj * o.wj.y; d.wsz = radi * wi.z + radj * o.wj.z;
    d.s0 = o.h0; d.s1 = o.h1; d.s2 = o.h2;
    return d.rsq < d.radsum * d.radsum;
  };
  auto contact = [&](const int sl, const Drv &d) {
    double s0 = d.s0, s1 = d.s1, s2 = d.s2, fox, foy, foz, tox, toy, toz;
    if (PAIR == PAIR_HERTZFIX_HISTORY) {
      hertzfix_fast(d.delx, d.dely, d.delz, d.rsq, d.vrx, d.vry, d.vrz, d.wsx, d.wsy, d.wsz, d.meff, d.radsum, d.reff, hc, P.dtv, shearupdate,
                    s0, s1, s2, fox, foy, foz, tox, toy, toz);
    } else {
      V3 vr = {d.vrx, d.vry, d.vrz}, ws = {d.wsx, d.wsy, d.wsz}, sh = {s0, s1, s2}, fo, to;
      if (PAIR == PAIR_HOOKE_HISTORY) hooke_history_contact(d.delx, d.dely, d.delz, d.rsq, vr, ws, d.meff, d.radsum, gc, P.dtv, shearupdate, sh, fo, to);
      else hooke_contact(d.delx, d.dely, d.delz, d.rsq, vr, ws, d.meff, d.radsum, gc, fo, to);
      s0 = sh.x; s1 = sh.y; s2 = sh.z; fox = fo.x; foy = fo.y; foz = fo.z; tox = to.x; toy = to.y; toz = to.z;
    }
    if (HIST) { D4 h; h.x = s0; h.y = s1; h.z = s2; h.w = 0.0; st_d4(&P.shear[(size_t)sl * P.npad + i], h); }
    // reference: f[i] += F ; torque[i] -= radi * tor   (pair :259-271)
    fx += fox; fy += foy; fz += foz;
    tx -= radi * tox; ty -= radi * toy; tz -= radi * toz;
  };

  mask_t m;   // entries handed to phase 2
  int s = 0;
  unsigned e = 0u;
  Opnd cur;
  bool have_cur = false;
  if (HIST && SEDI_SELL_SPEC) {
    // ---- phase 1, history styles: an entry that overlapped one sub-step ago (bit of tm_old) almost surely still does -- it goes to phase 2
    // untested (prep() repeats the reference's test on the operands it gathers anyway, so the result is the same set); only the other
    // entries of the row are tested here, four partner positions in flight at a time, and the operands of the first old contact are
    // requested in the same round trip.
    const mask_t valid = nni >= MBITS ? ~(mask_t)0 : ((ONE << nni) - ONE);
    const mask_t m_old = tm_old & valid;
    if (m_old) { s = mask_ffs(m_old); e = list_word(s); gather(e, s, cur); have_cur = true; }
    mask_t cand = valid & ~tm_old, m_new = 0;
    while (cand) {
      int sl[SEDI_SELL_P1B];
      unsigned ew[SEDI_SELL_P1B];
      D4 pp[SEDI_SELL_P1B];
#pragma unroll
      for (int k = 0; k < SEDI_SELL_P1B; k++) {
        sl[k] = 0; ew[k] = 0u;
        if (cand) { sl[k] = mask_ffs(cand); cand &= cand - 1; ew[k] = list_word(sl[k]); }
      }
#pragma unroll
      for (int k = 0; k < SEDI_SELL_P1B; k++) pp[k] = ldg_d4(&P.posr_in[ew[k] & NB_IDX_MASK]);
#pragma unroll
      for (int k = 0; k < SEDI_SELL_P1B; k++) test_entry(ew[k], pp[k], sl[k], m_new);
    }
    m = m_old | m_new;
    if (m) {
      const int sf = mask_ffs(m);
      if (!have_cur || sf != s) { s = sf; e = list_word(s); gather(e, s, cur); have_cur = true; }   // a new contact in front of the first old one (rare)
    }
  } else {
    // ---- phase 1: which list entries overlap
#pragma unroll
    for (int b = 0; b < 16; b += 8) {
      if (b < nni) {
        D4 p8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) p8[k] = ldg_d4(&P.posr_in[e16[b + k] & NB_IDX_MASK]);
#pragma unroll
        for (int k = 0; k < 8; k++) test_entry(e16[b + k], p8[k], b + k, touch);
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
      for (int k = 0; k < 4; k++) test_entry(e4[k], p4[k], sb + k, touch);
    }
    m = touch; touch = 0;
    if (m) { s = mask_ffs(m); e = list_word(s); gather(e, s, cur); }
  }

  // ---- phase 2: the (presumably) overlapping entries, in slot order; the rows of a warp have (nearly) the same number of them.
  // `cur` holds the gathered operands of entry (e, s).  They are consumed by prep(), the next entry's gathers are issued into the same
  // registers, and the contact law runs on the derived values while those loads are in flight (no second operand set, no copies).
  while (m) {
    Drv d;
    const bool ok = prep(e, cur, d);
    const int sc = s;
    m &= m - 1;
    if (m) { s = mask_ffs(m); e = list_word(s); gather(e, s, cur); }
    if (ok) { contact(sc, d); touch |= (ONE << sc); }
  }
  if (HIST && touch != tm_old) P.tmask[i] = (unsigned long long)touch;

  // ---- phase T (TYPELIST): fix cohesive / pair lubricate/poly over the entries of the type-cut-off list -- the granular segment
  // [0, nni) and the type-only segment [hcap, hcap + nti) --, four partner positions in flight at a time.  Rows of a warp have
  // (nearly) the same length, so the per-lane walk keeps the lanes busy.
  double lfx = 0.0, lfy = 0.0, lfz = 0.0, ltx = 0.0, lty = 0.0, ltz = 0.0;  // lubricate/poly
  double cfx = 0.0, cfy = 0.0, cfz = 0.0;                               // fix cohesive
  if (TYPELIST) {
    const CoheCoef co = cohesive_coef<(TYPELIST != 0)>(P);
    const int tagi = bits_tag((unsigned long long)__double_as_longlong(wi.w));
    const int ntot = nni + nti;
    const double inv_radi = 1.0 / radi;
    for (int sb = 0; sb < ntot; sb += 4) {
      unsigned e4[4];
      D4 p4[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int q = sb + k;
        e4[k] = (q < ntot) ? (q < nni ? list_word(q) : ld_nc_u32(&P.nbr[(size_t)(P.hcap + (q - nni)) * P.npad + i])) : 0u;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) p4[k] = ldg_d4(&P.posr_in[e4[k] & NB_IDX_MASK]);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned ew = e4[k];
        if (!(ew & NB_FLAG_TYPE)) continue;
        D4 pj = p4[k];
        const int img = (int)((ew >> NB_IMG_SHIFT) & 31u);
        if (PBC && img != NB_IMG_NONE) apply_image(P, img, pj);
        const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
        const double rsq = delx * delx + dely * dely + delz * delz;
        const double radj = pj.w;
        const double radsum = radi + radj;
        const int j = (int)(ew & NB_IDX_MASK);
        if (P.has_cohesive) cohesive_entry_fast(P, co, j, img, tagi, maski, radsum, rsq, delx, dely, delz, cfx, cfy, cfz);
        if (TYPELIST == 2 && P.lub_enabled && P.lub_flagHI && rsq < P.lub_cutsq) lubricate_entry_fast(P, j, pi, vi, wi, inv_radi, radj, rsq, delx, dely, delz, lfx, lfy, lfz, ltx, lty, ltz);
      }
    }
    if (TYPELIST == 2 && P.lub_enabled) {  // isotropic FLD terms (:213-221) are applied before the pair terms in the reference
      double ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
      if (P.lub_flagfld) {
        ax -= P.lub_R0 * radi * vi.x; ay -= P.lub_R0 * radi * vi.y; az -= P.lub_R0 * radi * vi.z;
        const double radi3 = radi * radi * radi;
        bx -= P.lub_RT0 * radi3 * wi.x; by -= P.lub_RT0 * radi3 * wi.y; bz -= P.lub_RT0 * radi3 * wi.z;
      }
      fx += ax + lfx; fy += ay + lfy; fz += az + lfz;
      tx += bx + ltx; ty += by + lty; tz += bz + ltz;
    }
  }
  double fd0 = 0.0, fd1 = 0.0, fd2 = 0.0, xh0 = 0.0, xh1 = 0.0, xh2 = 0.0;
  if (P.has_fdrag) { fd0 = ld_nc_f64(&P.fdrag[0][i]); fd1 = ld_nc_f64(&P.fdrag[1][i]); fd2 = ld_nc_f64(&P.fdrag[2][i]); }
  if (P.mode == MODE_FUSED) { xh0 = ld_nc_f64(&P.xhold[0][i]); xh1 = ld_nc_f64(&P.xhold[1][i]); xh2 = ld_nc_f64(&P.xhold[2][i]); }
  step_epilogue<PAIR, (TYPELIST != 0)>(P, i, seq, pi, vi, wi, fx, fy, fz, tx, ty, tz, cfx, cfy, cfz, fd0, fd1, fd2, xh0, xh1, xh2, (unsigned long long)touch);
}


template <int PAIR, bool PBC, int TYPELIST, bool M32>
__global__ void __launch_bounds__(SEDI_SELL_THREADS, TYPELIST == 2 ? 6 : SEDI_SELL_MINB) k_step_sell(const __grid_constant__ StepParams P, const int seq) {
  __shared__ unsigned s_e[16][SEDI_SELL_THREADS];   // the row's list words, kept for phase 2 (one column per lane: conflict-free)
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  // an earlier step of this chunk asked for a neighbour rebuild: this launch is a no-op (the flag is requested here, tested in sell_row)
  int fl = 0;
  if (P.mode != MODE_SETUP) fl = *(volatile int *)&P.ctrl[0];
  auto skipf = [&]() -> bool {
    const bool skip = (fl != 0 && fl < seq);
    if (i0 == 0 && P.mode != MODE_SETUP && !skip) atomicAdd(&P.ctrl[1], 1);   // executed sub-steps (an empty brick counts them too)
    return skip;
  };
  if (i0 < P.n) sell_row<PAIR, PBC, TYPELIST, M32>(P, seq, i0, s_e, skipf);
  else if (i0 == 0) skipf();
}

}  // namespace sedi
