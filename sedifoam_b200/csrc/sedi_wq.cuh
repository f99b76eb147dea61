// sedi_wq.cuh -- the warp-queue form of the fused DEM sub-step kernel (k_step_wq): the default pair sweep for the
// granular styles (gran/hertzFix/history, gran/hooke/history, gran/hooke) without a type-cut-off list.
//
// Same work per launch as k_step (sedi_step.cuh): force_clear -> pair->compute -> post_force fixes -> nve/sphere, same
// reference arithmetic (interfaceToLammps/pair_gran_hertzFix_history.cpp:120-285 and the fixes listed in
// sedi_step.cuh), same per-particle summation order, hence the same results bit for bit (contact law and epilogue are
// shared code).  What differs is the mapping of the pair sweep onto the machine, chosen for RAGGED rows (a random
// packing has 2..14 list entries per particle of which a third to a half overlap; the slot walk of k_step runs every
// warp for its longest row and executes the 290-instruction contact law whenever ANY lane overlaps at that slot):
//
//   phase 1  one lane per particle.  The first 16 list words of the row are requested together with the particle's own
//            state; partner positions are gathered eight at a time and only the distance is TESTED: the result is the
//            64-bit touch mask of the row;
//   queue    per list slot, a warp ballot of the touch bits gives every overlapping entry a position in one compact
//            queue of the warp's 32 rows, ordered BY SLOT, THEN LANE: consecutive queue entries belong to consecutive
//            rows and (rows being in bin order) mostly to neighbouring partners, so the gathers of a round coalesce
//            the way the ELL walk's do -- the L1 data pipe (wavefronts per gather), not HBM, is what an owner-ordered
//            queue saturates (profiles/r02_*);
//   phase 2  the warp evaluates the queue 32 entries per round, ONE OVERLAPPING CONTACT PER LANE whatever row it
//            belongs to (full lane utilisation of the expensive part).  The owner's position / velocity / spin come
//            from the warp's staged copy in shared memory, the partner's from L1/L2 (prefetched one round ahead), the
//            history quad is read and written in place in the ELL array;
//   reduce   force / torque of a round go through a double-buffered shared-memory panel; each owner adds its own
//            entries in ascending slot order -- the floating-point sum is the sequential walk's, the run stays bitwise
//            deterministic, there are no atomics;
//   epilogue step_epilogue<> (fixes in script order, final + initial integrate, skin/2 trigger), one lane per row.
// Only warp-level synchronisation is used (__syncwarp / ballots / shuffles); warps of a CTA are independent.
#pragma once
#include "sedi_step.cuh"

namespace sedi {

#ifndef SEDI_WQ_THREADS
#define SEDI_WQ_THREADS 128
#endif
#ifndef SEDI_WQ_MINB
#define SEDI_WQ_MINB 4
#endif
#ifndef SEDI_WQ_QCAP
#define SEDI_WQ_QCAP 16   // queue capacity: overlapping partners per row, averaged over the 32 rows of a warp
#endif
#ifndef SEDI_WQ_PF
#define SEDI_WQ_PF 1      // prefetch the next round's partner lines / history quad to L1 at the top of a round
#endif

static const int WQ_ERR_QUEUE = 2;   // ctrl[2] bit: a warp's contact queue overflowed

__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

template <int PAIR, bool PBC>
__global__ void __launch_bounds__(SEDI_WQ_THREADS, SEDI_WQ_MINB) k_step_wq(const __grid_constant__ StepParams P, const int seq) {
  constexpr int NW = SEDI_WQ_THREADS / 32;
  constexpr int QMAX = 32 * SEDI_WQ_QCAP;
  constexpr bool HIST = (PAIR == PAIR_HERTZFIX_HISTORY || PAIR == PAIR_HOOKE_HISTORY);
  if (P.mode != MODE_SETUP) {
    const int fl = *(volatile int *)&P.ctrl[0];
    if (fl != 0 && fl < seq) return;  // an earlier step of this chunk asked for a neighbour rebuild: become a no-op
  }
  __shared__ __align__(32) D4 s_pos[NW][32];
  __shared__ __align__(32) D4 s_vel[NW][32];
  __shared__ __align__(32) D4 s_omg[NW][32];
  __shared__ double s_part[NW][2][6][32];
  __shared__ unsigned long long s_tm[NW][32];
  __shared__ unsigned s_qe[NW][QMAX];          // queue: list word of the entry
  __shared__ unsigned short s_q[NW][QMAX];     // queue: (owner lane << 8) | slot
  __shared__ unsigned s_bal[NW][MAX_SLOTS];    // per slot: which lanes overlap at that slot
  __shared__ unsigned short s_off[NW][MAX_SLOTS];  // per slot: queue position of its first entry

  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row0 = blockIdx.x * SEDI_WQ_THREADS + w * 32;   // first row of this warp
  const int i = row0 + lane;
  if (i == 0 && P.mode != MODE_SETUP) atomicAdd(&P.ctrl[1], 1);
  if (row0 >= P.n) return;                                   // warp-uniform

  // ---- own row and the first 16 list words (arrays are padded to a multiple of 128 rows: loads past n are harmless)
  D4 pi = ldg_d4_stream(&P.posr_in[i]);
  D4 vi = ldg_d4_stream(&P.velm_in[i]);
  D4 wi = ldg_d4_stream(&P.omgt_in[i]);
  const int nn_raw = (i < P.n) ? ld_nc_s32(&P.nn[i]) : 0;
  unsigned e16[16];
#pragma unroll
  for (int k = 0; k < 16; k++) e16[k] = ld_nc_u32(&P.nbr[(size_t)k * P.npad + i]);   // hcap >= 12 rows exist; slots >= nn hold stale words, masked below
  const bool own = (i < P.n) && !(bits_flags((unsigned long long)__double_as_longlong(wi.w)) & PFLAG_GHOST);
  const int nni = own ? nn_raw : 0;
  const unsigned long long tm_old = (HIST && own) ? P.tmask[i] : 0ull;
  s_pos[w][lane] = pi; s_vel[w][lane] = vi; s_omg[w][lane] = wi; s_tm[w][lane] = tm_old;
  const double radi = pi.w;
  const int maxnn = __reduce_max_sync(full, nni);

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
    if (b < maxnn) {   // warp-uniform
      D4 p8[8];
#pragma unroll
      for (int k = 0; k < 8; k++) { if (b + k >= nni) e16[b + k] = 0u; p8[k] = ldg_d4(&P.posr_in[e16[b + k] & NB_IDX_MASK]); }
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

  // ---- queue: slot-major.  Entry (lane, s) sits at off[s] + (number of lower lanes that overlap at slot s).
  const unsigned lt = lanemask_lt();
  int total = 0;   // warp-uniform
#pragma unroll
  for (int k = 0; k < 16; k++) {
    if (k < maxnn) {
      const bool t = (touch >> k) & 1ull;
      const unsigned bal = __ballot_sync(full, t);
      if (lane == 0) { s_bal[w][k] = bal; s_off[w][k] = (unsigned short)total; }
      const int q = total + __popc(bal & lt);
      if (t && q < QMAX) { s_q[w][q] = (unsigned short)((lane << 8) | k); s_qe[w][q] = e16[k]; }
      total += __popc(bal);
    }
  }
  for (int k = 16; k < maxnn; k++) {
    const bool t = (touch >> k) & 1ull;
    const unsigned bal = __ballot_sync(full, t);
    if (lane == 0) { s_bal[w][k] = bal; s_off[w][k] = (unsigned short)total; }
    const int q = total + __popc(bal & lt);
    if (t && q < QMAX) { s_q[w][q] = (unsigned short)((lane << 8) | k); s_qe[w][q] = ld_nc_u32(&P.nbr[(size_t)k * P.npad + i]); }
    total += __popc(bal);
  }
  if (total > QMAX) {   // cannot happen for spheres of moderate size ratio; reported, never silently truncated
    if (lane == 0) atomicOr(&P.ctrl[2], WQ_ERR_QUEUE);
    total = 0; touch = 0ull;
  }
  __syncwarp();

  // ---- phase 2: one overlapping contact per lane and round
  const bool shearupdate = (P.mode != MODE_SETUP);
  HzCoef hc; hc.c_sn = P.c_sn; hc.c_ccel = P.c_ccel; hc.c_damp = P.c_damp; hc.c_kts = P.c_kts; hc.c_ctd = P.c_ctd; hc.c_ekt = P.c_ekt; hc.xmu = P.xmu;
  GranCoef gc; gc.kn = P.kn; gc.kt = P.kt; gc.gamman = P.gamman; gc.gammat = P.gammat; gc.xmu = P.xmu; gc.beta = P.beta;
  double fx = 0.0, fy = 0.0, fz = 0.0, tx = 0.0, ty = 0.0, tz = 0.0;   // pair accumulators (force_clear)
  unsigned long long rem = touch;   // own entries not yet added, lowest slot first
  int qnext = -1;                   // queue position of the lowest one
  if (rem) { const int k = __ffsll((long long)rem) - 1; qnext = (int)s_off[w][k] + __popc(s_bal[w][k] & lt); }
  int buf = 0;
  for (int base = 0; base < total; base += 32, buf ^= 1) {
    const int q = base + lane;
#if SEDI_WQ_PF
    if (q + 32 < total) {   // next round's partner lines and history quad start towards L1 now
      const unsigned en = s_qe[w][q + 32], qn = s_q[w][q + 32];
      const int jn = (int)(en & NB_IDX_MASK);
      prefetch_l1(&P.posr_in[jn]); prefetch_l1(&P.velm_in[jn]); prefetch_l1(&P.omgt_in[jn]);
      const int Ln = (int)(qn >> 8), sn = (int)(qn & 255u);
      if (HIST && ((s_tm[w][Ln] >> sn) & 1ull)) prefetch_l1(&P.shear[(size_t)sn * P.npad + (row0 + Ln)]);
    }
#endif
    if (q < total) {
      const unsigned ent = s_q[w][q], e = s_qe[w][q];
      const int L = (int)(ent >> 8), s = (int)(ent & 255u);
      const size_t slot = (size_t)s * P.npad + (row0 + L);
      const int j = (int)(e & NB_IDX_MASK);
      D4 pj = ldg_d4(&P.posr_in[j]);
      const D4 vj = ldg_d4(&P.velm_in[j]);
      const D4 wj = ldg_d4(&P.omgt_in[j]);
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      if (HIST && ((s_tm[w][L] >> s) & 1ull)) { const D4 h = ld_d4(&P.shear[slot]); s0 = h.x; s1 = h.y; s2 = h.z; }
      const D4 po = s_pos[w][L], vo = s_vel[w][L], wo = s_omg[w][L];
      const int img = (int)((e >> NB_IMG_SHIFT) & 31u);
      if (PBC && img != NB_IMG_NONE) {
        pj.x = pj.x + P.imgshift[img][0]; pj.y = pj.y + P.imgshift[img][1]; pj.z = pj.z + P.imgshift[img][2];
      }
      const double delx = po.x - pj.x, dely = po.y - pj.y, delz = po.z - pj.z;
      const double rsq = delx * delx + dely * dely + delz * delz;
      const double rado = po.w, radj = pj.w, mo = vo.w, mj = vj.w;
      const double radsum = rado + radj;
      const int masko = bits_mask((unsigned long long)__double_as_longlong(wo.w));
      const int maskj = bits_mask((unsigned long long)__double_as_longlong(wj.w));
      double meff = (PAIR == PAIR_HERTZFIX_HISTORY) ? div_nr(mo * mj, mo + mj) : (mo * mj) / (mo + mj);
      if (masko & P.freeze_groupbit) meff = mj;
      if (maskj & P.freeze_groupbit) meff = mo;
      const double vrx = vo.x - vj.x, vry = vo.y - vj.y, vrz = vo.z - vj.z;
      const double wsx = rado * wo.x + radj * wj.x, wsy = rado * wo.y + radj * wj.y, wsz = rado * wo.z + radj * wj.z;
      double fox, foy, foz, tox, toy, toz;
      if (PAIR == PAIR_HERTZFIX_HISTORY) {
        hertzfix_fast(delx, dely, delz, rsq, vrx, vry, vrz, wsx, wsy, wsz, meff, radsum, div_nr(rado * radj, radsum), hc, P.dtv, shearupdate,
                      s0, s1, s2, fox, foy, foz, tox, toy, toz);
      } else {
        V3 vr = {vrx, vry, vrz}, ws = {wsx, wsy, wsz}, sh = {s0, s1, s2}, fo, to;
        if (PAIR == PAIR_HOOKE_HISTORY) hooke_history_contact(delx, dely, delz, rsq, vr, ws, meff, radsum, gc, P.dtv, shearupdate, sh, fo, to);
        else hooke_contact(delx, dely, delz, rsq, vr, ws, meff, radsum, gc, fo, to);
        s0 = sh.x; s1 = sh.y; s2 = sh.z; fox = fo.x; foy = fo.y; foz = fo.z; tox = to.x; toy = to.y; toz = to.z;
      }
      if (HIST) { D4 h; h.x = s0; h.y = s1; h.z = s2; h.w = 0.0; st_d4(&P.shear[slot], h); }
      s_part[w][buf][0][lane] = fox; s_part[w][buf][1][lane] = foy; s_part[w][buf][2][lane] = foz;
      s_part[w][buf][3][lane] = tox; s_part[w][buf][4][lane] = toy; s_part[w][buf][5][lane] = toz;
    }
    __syncwarp();
    // this particle's entries inside the round, in slot order (reference: f[i] += F ; torque[i] -= radi * tor, pair :259-271)
    while (rem && qnext < base + 32) {
      const int c = qnext - base;
      fx += s_part[w][buf][0][c]; fy += s_part[w][buf][1][c]; fz += s_part[w][buf][2][c];
      tx -= radi * s_part[w][buf][3][c]; ty -= radi * s_part[w][buf][4][c]; tz -= radi * s_part[w][buf][5][c];
      rem &= rem - 1;
      if (rem) { const int k = __ffsll((long long)rem) - 1; qnext = (int)s_off[w][k] + __popc(s_bal[w][k] & lt); }
    }
    // no second barrier: the next round fills the other panel, and that round's barrier orders the reuse of this one
  }
  if (!own) return;
  if (HIST && touch != tm_old) P.tmask[i] = touch;
  double fd0 = 0.0, fd1 = 0.0, fd2 = 0.0, xh0 = 0.0, xh1 = 0.0, xh2 = 0.0;
  if (P.has_fdrag) { fd0 = ld_nc_f64(&P.fdrag[0][i]); fd1 = ld_nc_f64(&P.fdrag[1][i]); fd2 = ld_nc_f64(&P.fdrag[2][i]); }
  if (P.mode == MODE_FUSED) { xh0 = ld_nc_f64(&P.xhold[0][i]); xh1 = ld_nc_f64(&P.xhold[1][i]); xh2 = ld_nc_f64(&P.xhold[2][i]); }
  pi = s_pos[w][lane]; vi = s_vel[w][lane]; wi = s_omg[w][lane];
  step_epilogue<PAIR, false>(P, i, seq, pi, vi, wi, fx, fy, fz, tx, ty, tz, 0.0, 0.0, 0.0, fd0, fd1, fd2, xh0, xh1, xh2, touch);
}

}  // namespace sedi
