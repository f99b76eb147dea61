// sedi_rows.cuh -- the row-block form of the fused DEM sub-step kernel (k_step_rows) and the list / history layout
// it runs on.
//
// Same work per launch as k_step (sedi_step.cuh): force_clear -> pair->compute -> post_force fixes -> nve/sphere,
// same reference arithmetic (interfaceToLammps/pair_gran_hertzFix_history.cpp:120-285 and the fixes listed in
// sedi_step.cuh), same results bit for bit (the contact law, the per-particle summation order and the epilogue
// are shared code).  What differs is how the pair sweep is mapped onto the machine:
//
//   * the directed neighbour list is stored ROW-CONTIGUOUS (CSR: the entries of particle i are
//     cnbr[off[i] .. off[i+1]), in the slot order of the ELL list it is compacted from), and the contact history
//     lives in three FP64 planes hx/hy/hz indexed by the same entry number: 24 B per directed entry instead of
//     the 32 B quad of the ELL layout, and no padding slots;
//   * a CTA owns a block of RT consecutive (bin-ordered) particle rows.  Thread 0 stages the block's
//     position / velocity / spin quads into shared memory with three TMA bulk copies (cp.async.bulk, completion on
//     an mbarrier) while the other threads fetch the row offsets and touch masks;
//   * the pair sweep runs ONE DIRECTED ENTRY PER THREAD over the block's contiguous entry range, RT entries per
//     round: list word, owner-row byte and history of consecutive lanes are consecutive addresses (coalesced both
//     ways), every lane of a round has a pair to evaluate (the ELL walk idles the lanes of short rows), and all
//     gathers of a thread are independent of each other;
//   * per-entry force / torque contributions go through a double-buffered shared-memory panel; after each round
//     the particle's own thread adds the contributions of its entries IN SLOT ORDER, so the floating-point sum is
//     the same as the sequential walk's and the run stays bitwise deterministic without atomics;
//   * the epilogue (fixes, integration, displacement check, write-back) is step_epilogue<>, one thread per row.
#pragma once
#include "sedi_step.cuh"

namespace sedi {

#ifndef SEDI_ROWS_THREADS
#define SEDI_ROWS_THREADS 64
#endif
#ifndef SEDI_ROWS_MINB
#define SEDI_ROWS_MINB 8
#endif
#ifndef SEDI_ROWS_TMA
#define SEDI_ROWS_TMA 1
#endif
#ifndef SEDI_ROWS_SPEC
#define SEDI_ROWS_SPEC 1   // a pair that touched in the previous sub-step requests velocity / spin / history together with the position
#endif
#ifndef SEDI_ROWS_PF
#define SEDI_ROWS_PF 1     // prefetch the next round's partner lines to L1 before the round barrier
#endif
#ifndef SEDI_ROWS_INBLK
#define SEDI_ROWS_INBLK 1  // a partner that lives in the CTA's own row block is read from the staged shared-memory copy
#endif
#ifndef SEDI_ROWS_LATE
#define SEDI_ROWS_LATE 1   // fluid force / xhold are read after the sweep (L2 prefetch up front) instead of held in registers
#endif

// ---- mbarrier / TMA bulk copy (PTX) ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// 1-D TMA: global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned ld_nc_u8(const unsigned char *p) { unsigned r; asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(r) : "l"(p)); return r; }
__device__ __forceinline__ double ld_f64_stream(const double *p) { double r; asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p)); return r; }

// ---- ELL -> row-contiguous compaction (after every neighbour rebuild) and the way back (before the history is needed
// in ELL form again: next rebuild's re-attachment, migration packing, sedi_get_pairs) ---------------------------------
__global__ void k_rows_fill(int n, int npad, const int *nn, const int *off, const unsigned *nbr, const D4 *shear,
                            const unsigned long long *tmask, unsigned *cnbr, unsigned char *crow, double *hx, double *hy, double *hz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = nn[i], o = off[i];
  const unsigned long long tm = tmask[i];
  for (int s = 0; s < c; s++) {
    const size_t slot = (size_t)s * npad + i;
    cnbr[o + s] = nbr[slot];
    crow[o + s] = (unsigned char)(i & 255);
    double a = 0.0, b = 0.0, d = 0.0;
    if ((tm >> s) & 1ull) { const D4 h = shear[slot]; a = h.x; b = h.y; d = h.z; }
    hx[o + s] = a; hy[o + s] = b; hz[o + s] = d;
  }
}
__global__ void k_rows_history_to_ell(int n, int npad, const int *nn, const int *off, const unsigned long long *tmask, const double *hx,
                                      const double *hy, const double *hz, D4 *shear) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = nn[i], o = off[i];
  const unsigned long long tm = tmask[i];
  for (int s = 0; s < c; s++) {
    if (!((tm >> s) & 1ull)) continue;
    D4 h; h.x = hx[o + s]; h.y = hy[o + s]; h.z = hz[o + s]; h.w = 0.0;
    shear[(size_t)s * npad + i] = h;
  }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
template <int PAIR>
__global__ void __launch_bounds__(SEDI_ROWS_THREADS, SEDI_ROWS_MINB) k_step_rows(const __grid_constant__ StepParams P, const int seq) {
  constexpr int RT = SEDI_ROWS_THREADS;
  constexpr bool HIST = (PAIR == PAIR_HERTZFIX_HISTORY || PAIR == PAIR_HOOKE_HISTORY);
  static_assert(RT <= 256 && (RT & (RT - 1)) == 0, "owner-row byte holds row & 255");
  if (P.mode != MODE_SETUP) {
    // uniform over the launch: ctrl[0] only ever holds the sequence number of a launch that has already run or of this one
    const int fl = *(volatile int *)&P.ctrl[0];
    if (fl != 0 && fl < seq) return;
  }
  __shared__ __align__(128) D4 s_pos[RT];
  __shared__ __align__(128) D4 s_vel[RT];
  __shared__ __align__(128) D4 s_omg[RT];
  __shared__ double s_part[2][6][RT];
  __shared__ unsigned long long s_tm[RT];
  __shared__ int s_off[RT + 1];
  __shared__ unsigned char s_flag[2][RT];
  __shared__ __align__(8) unsigned long long s_bar;

  const int t = threadIdx.x;
  const int r0 = blockIdx.x * RT;
  const int i = r0 + t;
  const bool own = i < P.n;
  if (i == 0 && P.mode != MODE_SETUP) atomicAdd(&P.ctrl[1], 1);

#if SEDI_ROWS_TMA
  if (t == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&s_bar, 3u * RT * (unsigned)sizeof(D4));
    tma_bulk_g2s(s_pos, &P.posr_in[r0], RT * (unsigned)sizeof(D4), &s_bar);   // rows r0 .. r0+RT-1 exist: arrays are padded to 128 rows
    tma_bulk_g2s(s_vel, &P.velm_in[r0], RT * (unsigned)sizeof(D4), &s_bar);
    tma_bulk_g2s(s_omg, &P.omgt_in[r0], RT * (unsigned)sizeof(D4), &s_bar);
  }
#else
  s_pos[t] = ldg_d4_stream(&P.posr_in[i]);
  s_vel[t] = ldg_d4_stream(&P.velm_in[i]);
  s_omg[t] = ldg_d4_stream(&P.omgt_in[i]);
#endif
  {
    const int lim = P.n;   // off[] has n + 1 entries
    s_off[t] = ld_nc_s32(&P.off[i < lim ? i : lim]);
    if (t == RT - 1) s_off[RT] = ld_nc_s32(&P.off[(r0 + RT) < lim ? (r0 + RT) : lim]);
    s_tm[t] = (HIST && own) ? P.tmask[i] : 0ull;
  }
  // streamed per-particle inputs of the epilogue: requested now, consumed after the sweep
  double fd0 = 0.0, fd1 = 0.0, fd2 = 0.0, xh0 = 0.0, xh1 = 0.0, xh2 = 0.0;
#if SEDI_ROWS_LATE
  if (own && (t & 3) == 0) {   // one request per 32-byte sector
    if (P.has_fdrag) { prefetch_l2(&P.fdrag[0][i]); prefetch_l2(&P.fdrag[1][i]); prefetch_l2(&P.fdrag[2][i]); }
    if (P.mode == MODE_FUSED) { prefetch_l2(&P.xhold[0][i]); prefetch_l2(&P.xhold[1][i]); prefetch_l2(&P.xhold[2][i]); }
  }
#else
  if (own) {
    if (P.has_fdrag) { fd0 = ld_nc_f64(&P.fdrag[0][i]); fd1 = ld_nc_f64(&P.fdrag[1][i]); fd2 = ld_nc_f64(&P.fdrag[2][i]); }
    if (P.mode == MODE_FUSED) { xh0 = ld_nc_f64(&P.xhold[0][i]); xh1 = ld_nc_f64(&P.xhold[1][i]); xh2 = ld_nc_f64(&P.xhold[2][i]); }
  }
#endif
  __syncthreads();
  const int e0 = s_off[0];
  const int ne = s_off[RT] - e0;
  const int my_a = s_off[t] - e0, my_b = s_off[t + 1] - e0;   // this particle's entries, relative to the block

  // first round's list word and owner byte go out before the wait on the staged rows
  unsigned w_nxt = 0u, r_nxt = 0u;
  if (t < ne) { w_nxt = ld_nc_u32(&P.cnbr[e0 + t]); r_nxt = ld_nc_u8(&P.crow[e0 + t]); }
#if SEDI_ROWS_TMA
  mbar_wait(&s_bar, 0);
#endif

  const bool shearupdate = (P.mode != MODE_SETUP);
  HzCoef hc; hc.c_sn = P.c_sn; hc.c_ccel = P.c_ccel; hc.c_damp = P.c_damp; hc.c_kts = P.c_kts; hc.c_ctd = P.c_ctd; hc.c_ekt = P.c_ekt; hc.xmu = P.xmu;
  GranCoef gc; gc.kn = P.kn; gc.kt = P.kt; gc.gamman = P.gamman; gc.gammat = P.gammat; gc.xmu = P.xmu; gc.beta = P.beta;

  double fx = 0.0, fy = 0.0, fz = 0.0, tx = 0.0, ty = 0.0, tz = 0.0;
  unsigned long long touch = 0ull;
  const double radi_own = s_pos[t].w;

  int buf = 0;
  for (int base = 0; base < ne; base += RT, buf ^= 1) {
    const int k = base + t;
    const unsigned e = w_nxt, rl = r_nxt;
    if (k + RT < ne) { w_nxt = ld_nc_u32(&P.cnbr[e0 + k + RT]); r_nxt = ld_nc_u8(&P.crow[e0 + k + RT]); }
    bool hit = false;
    if (k < ne && (e & NB_FLAG_GRAN)) {
      const int j = (int)(e & NB_IDX_MASK);
#if SEDI_ROWS_INBLK
      const bool inblk = (unsigned)(j - r0) < (unsigned)RT;
      D4 pj = inblk ? s_pos[j - r0] : ldg_d4(&P.posr_in[j]);
#else
      D4 pj = ldg_d4(&P.posr_in[j]);
#endif
      const int il = (int)((rl - (unsigned)r0) & 255u);
      const int s = (e0 + k) - s_off[il];
      const bool had = HIST && ((s_tm[il] >> s) & 1ull);
#if SEDI_ROWS_SPEC
      D4 vj_s, wj_s;
      double h0 = 0.0, h1 = 0.0, h2 = 0.0;
      if (had) {   // touched one sub-step ago: it still does, almost surely -- do not wait for the distance test
#if SEDI_ROWS_INBLK
        if (inblk) { vj_s = s_vel[j - r0]; wj_s = s_omg[j - r0]; }
        else
#endif
        { vj_s = ldg_d4(&P.velm_in[j]); wj_s = ldg_d4(&P.omgt_in[j]); }
        h0 = ld_f64_stream(&P.hx[e0 + k]); h1 = ld_f64_stream(&P.hy[e0 + k]); h2 = ld_f64_stream(&P.hz[e0 + k]);
      }
#endif
      const D4 pi = s_pos[il];
      const int img = (int)((e >> NB_IMG_SHIFT) & 31u);
      if (P.periodic_any && img != NB_IMG_NONE) {
        pj.x = pj.x + P.imgshift[img][0]; pj.y = pj.y + P.imgshift[img][1]; pj.z = pj.z + P.imgshift[img][2];
      }
      const double delx = pi.x - pj.x, dely = pi.y - pj.y, delz = pi.z - pj.z;
      const double rsq = delx * delx + dely * dely + delz * delz;
      const double radi = pi.w, radj = pj.w;
      const double radsum = radi + radj;
      if (rsq < radsum * radsum) {
        hit = true;
#if SEDI_ROWS_SPEC
        D4 vj, wj;
        if (had) { vj = vj_s; wj = wj_s; }
#if SEDI_ROWS_INBLK
        else if (inblk) { vj = s_vel[j - r0]; wj = s_omg[j - r0]; }
#endif
        else { vj = ldg_d4(&P.velm_in[j]); wj = ldg_d4(&P.omgt_in[j]); }
        double s0 = h0, s1 = h1, s2 = h2;
#else
#if SEDI_ROWS_INBLK
        const D4 vj = inblk ? s_vel[j - r0] : ldg_d4(&P.velm_in[j]);
        const D4 wj = inblk ? s_omg[j - r0] : ldg_d4(&P.omgt_in[j]);
#else
        const D4 vj = ldg_d4(&P.velm_in[j]);
        const D4 wj = ldg_d4(&P.omgt_in[j]);
#endif
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        if (had) { s0 = ld_f64_stream(&P.hx[e0 + k]); s1 = ld_f64_stream(&P.hy[e0 + k]); s2 = ld_f64_stream(&P.hz[e0 + k]); }
#endif
        const D4 vi = s_vel[il];
        const D4 wi = s_omg[il];
        const double mi = vi.w, mj = vj.w;
        const int maski = bits_mask((unsigned long long)__double_as_longlong(wi.w));
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
        if (HIST) { P.hx[e0 + k] = s0; P.hy[e0 + k] = s1; P.hz[e0 + k] = s2; }
        s_part[buf][0][t] = fox; s_part[buf][1][t] = foy; s_part[buf][2][t] = foz;
        s_part[buf][3][t] = tox; s_part[buf][4][t] = toy; s_part[buf][5][t] = toz;
      }
    }
    s_flag[buf][t] = hit ? 1 : 0;
#if SEDI_ROWS_PF
    if (k + RT < ne && (w_nxt & NB_FLAG_GRAN)) {   // the next round's list word has arrived by now
      const int jn = (int)(w_nxt & NB_IDX_MASK);
#if SEDI_ROWS_INBLK
      if ((unsigned)(jn - r0) >= (unsigned)RT)
#endif
      { prefetch_l1(&P.posr_in[jn]); prefetch_l1(&P.velm_in[jn]); prefetch_l1(&P.omgt_in[jn]); }
    }
#endif
    __syncthreads();
    // this particle's entries inside the round's window, in slot order (reference: f[i] += F ; torque[i] -= radi * tor, pair :259-271)
    const int lo = my_a > base ? my_a : base;
    const int hi = my_b < base + RT ? my_b : base + RT;
    for (int q = lo; q < hi; q++) {
      const int c = q - base;
      if (!s_flag[buf][c]) continue;
      touch |= (1ull << (q - my_a));
      fx += s_part[buf][0][c]; fy += s_part[buf][1][c]; fz += s_part[buf][2][c];
      tx -= radi_own * s_part[buf][3][c]; ty -= radi_own * s_part[buf][4][c]; tz -= radi_own * s_part[buf][5][c];
    }
    // no second barrier: the next round fills the other panel, and the barrier of that round orders the reuse of this one
  }
  if (!own) return;
  const D4 pi = s_pos[t], vi = s_vel[t], wi = s_omg[t];
  if (bits_flags((unsigned long long)__double_as_longlong(wi.w)) & PFLAG_GHOST) return;  // ghost rows are refreshed by the halo exchange
  if (HIST && touch != s_tm[t]) P.tmask[i] = touch;
#if SEDI_ROWS_LATE
  if (P.has_fdrag) { fd0 = ld_nc_f64(&P.fdrag[0][i]); fd1 = ld_nc_f64(&P.fdrag[1][i]); fd2 = ld_nc_f64(&P.fdrag[2][i]); }
  if (P.mode == MODE_FUSED) { xh0 = ld_nc_f64(&P.xhold[0][i]); xh1 = ld_nc_f64(&P.xhold[1][i]); xh2 = ld_nc_f64(&P.xhold[2][i]); }
#endif
  step_epilogue<PAIR, false>(P, i, seq, pi, vi, wi, fx, fy, fz, tx, ty, tz, 0.0, 0.0, 0.0, fd0, fd1, fd2, xh0, xh1, xh2, touch);
}

}  // namespace sedi
