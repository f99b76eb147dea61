// sedi_device.cuh -- device-side data layout and kernel parameter blocks of the B200 particle engine.
//
// HBM layout (all FP64, see DESIGN.md "Data layout"):
//   posr[i] = {x, y, z, radius}            32 B, one 256-bit sector  -> one LDG.E.ENL2.256 per gathered partner
//   velm[i] = {vx, vy, vz, rmass}          32 B
//   omgt[i] = {wx, wy, wz, bits(tag|mask|type|flags)}  32 B
// The three quads are double-buffered (A/B): the fused step kernel reads A (own row + gathered partners) and
// writes B, so the whole DEM sub-step is ONE launch with no race between a particle's update and its
// partners' reads.  Streaming-only per-particle data (fluid force, vOld, xhold, f, torque, wall history) are
// SoA planes.  Neighbour list and contact history are ELL, slot-major: entry (i, s) at [s * npad + i], so a
// warp reading slot s of 32 consecutive particles touches contiguous memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sedi {

struct __align__(32) D4 { double x, y, z, w; };

// omgt.w bit layout
__host__ __device__ inline unsigned long long pack_bits(int tag, int mask, int type, int flags) {
  return ((unsigned long long)(unsigned)tag) | ((unsigned long long)(mask & 0xFFFF) << 32) |
         ((unsigned long long)(type & 0xFF) << 48) | ((unsigned long long)(flags & 0xFF) << 56);
}
__host__ __device__ inline int bits_tag(unsigned long long b) { return (int)(unsigned)(b & 0xFFFFFFFFull); }
__host__ __device__ inline int bits_mask(unsigned long long b) { return (int)((b >> 32) & 0xFFFF); }
__host__ __device__ inline int bits_type(unsigned long long b) { return (int)((b >> 48) & 0xFF); }
__host__ __device__ inline int bits_flags(unsigned long long b) { return (int)((b >> 56) & 0xFF); }
enum { PFLAG_GHOST = 1 };

// neighbour entry: [24:0] partner index, [29:25] periodic image code (0..26, 13 = none), [30] in the granular
// list (rsq <= (ri+rj+skin)^2), [31] in the type-cutoff list (rsq <= cutneighsq[ti][tj]) used by fix cohesive
// and pair lubricate/poly.
static const unsigned NB_IDX_MASK = 0x01FFFFFFu;
static const int NB_IMG_SHIFT = 25;
static const unsigned NB_FLAG_GRAN = 1u << 30;
static const unsigned NB_FLAG_TYPE = 1u << 31;
static const int NB_IMG_NONE = 13;
static const int MAX_SLOTS = 64;  // granular entries per row: the touch mask is one 64-bit word per particle (type-only entries are not limited)
static const int MAX_FIXES = 12;
static const int MAX_WALLS = 6;
static const int MAX_TYPES = 8;

struct FixDev {
  int kind, groupbit;
  int i0, i1, i2, i3;        // wall: wallstyle, wiggle, wshear, axis ; cohesive: opt
  long long time_origin;
  int wall_index, pad;
  double d[10];
  // gravity: d0..2 = g * nhat ; fdrag: d0 = carrier_rho ; cohesive: d0 ah, d1 lam, d2 smin, d3 smax
  // wall: d0 kn, d1 kt, d2 gamman, d3 gammat, d4 xmu, d5 lo, d6 hi, d7 cylradius, d8 beta, and
  //       wiggle: d[9] = amplitude (period in aux), shear: d[9] = vshear
  double aux;
};

// ---- ghost refresh fused into the sub-step kernel (multi-GPU, peer-memory path): the border rows of this brick are written
// straight into the neighbours' ghost rows by the thread that has just integrated them (step_epilogue)
static const int MAX_LINKS = 26;
struct PushTable {
  int nlinks;
  int base[MAX_LINKS + 1];
  int rstart[MAX_LINKS];       // first ghost row, in the peer's arrays, of the segment this link fills
  double shift[MAX_LINKS][3];
  D4 *rposr[MAX_LINKS], *rvelm[MAX_LINKS], *romgt[MAX_LINKS];   // the peer's arrays of the buffer being written
};
struct BorderEnt { int link, pos; };   // a border row's entry: link and position in that link's send list

struct StepParams {
  int n;        // rows in the particle arrays (owned + ghost)
  int npad;     // ELL leading dimension
  int nfix;
  int pair;     // PairKind
  int mode;     // StepMode
  int lub_enabled, lub_flaglog, lub_flagfld, lub_flagHI;
  int has_cohesive;
  int nve_groupbit;  // 0 = no integrator
  int freeze_groupbit;
  int periodic_any;
  int equal_spheres;  // all spheres of the system have one radius and one mass: meff = m / 2, reff = r / 2 (host-checked, same on every rank)
  long long ntimestep;
  const D4 *posr_in, *velm_in, *omgt_in;
  D4 *posr_out, *velm_out, *omgt_out;
  const int *nn;              // granular entries per row, slots [0, nn)
  const int *nt;              // type-only entries per row, slots [hcap, hcap + nt)
  int hcap;
  const unsigned *nbr;
  D4 *shear;                  // [slot * npad + i], .w unused
  unsigned long long *tmask;  // touching-slot mask per particle
  double *f[3], *tq[3];       // stored force/torque (written by SETUP/LAST, read by the initial-integrate kernel)
  double *fdrag[3], *dudt[3], *vold[3];
  double *xhold[3];
  double *wshear[MAX_WALLS][3];
  unsigned *wmask;            // per-particle wall-touch bits
  // fused ghost refresh (null on one GPU / NCCL halo): entries of row i are bent[bpos[i] .. bpos[i] + bcnt[i])
  const unsigned char *bcnt;
  const int *bpos;
  const BorderEnt *bent;
  const PushTable *push;      // device copy of the table for the buffer this launch writes
  int *ctrl;                  // [0] rebuild-needed flag, [1] steps completed, [2] error flags
  unsigned long long *counters;  // [0] directed pair visits, [1] directed touching pairs
  double dtv, dtf, dt_live, trigger_sq;
  double c_dtfm, c_dtirot;   // equal_spheres: dtf / m and (dtf / 0.4) / (r r m) of the one particle class
  double kn, kt, gamman, gammat, xmu, beta;
  double prd[3];
  double lub_mu, lub_cutsq, lub_cut_inner, lub_R0, lub_RT0;
  // host-folded constants of the Hertz-Mindlin "Fix" law (pair_gran_hertzFix_history.cpp:192-236)
  double c_sn;     // 2/1.82 * kn              : sn = c_sn * polyhertz
  double c_ccel;   // 4/5.46 * kn              : elastic part of ccel = polyhertz * c_ccel * (radsum - r) / r
  double c_damp;   // 2 sqrt(5/6) beta         : damp = c_damp * vnnr / rsq
  double c_kts;    // 8/8.84 * kt              : tangential spring = -polyhertz * c_kts * shear
  double c_ctd;    // sqrt(st/sn) 2 sqrt(5/6) beta : tangential dashpot = sqrt(sn meff) * c_ctd
  double c_ekt;    // 8/(8.84 kt)              : Coulomb rescale offset = ctd * vtr * c_ekt
  int has_fdrag, fdrag_added_mass, has_cyl_wall;
  double imgshift[27][3];  // periodic image code -> shift vector
  FixDev fix[MAX_FIXES];
};

enum StepMode {
  MODE_SETUP = 0,  // forces only, history frozen (update->setupflag), store f/torque
  MODE_FUSED = 1,  // forces + final_integrate(n) + initial_integrate(n+1), displacement check
  MODE_LAST = 2    // forces + final_integrate(n), store f/torque (end of a `run`)
};

struct BinParams {
  int n;
  double lo[3], inv[3];  // bin = floor((x - lo) * inv)
  int nb[3];
  int periodic[3];
  int tile[2];           // row order: x-major bins (tile = 1 x 1) or TX x TY bin tiles in the x-y plane, tile-major
};

// position of bin (bx, by, bz) in the sorted row order.  Plain: x-major.  Tiled: the bins of a TX x TY tile are
// contiguous (x fastest inside the tile), tiles x-major, then z -- a block of consecutive rows is then a compact
// x-y patch of the bed whose in-plane neighbours are mostly inside the same block.
__host__ __device__ inline int cell_index(int bx, int by, int bz, const int nb[3], const int tile[2]) {
  const int TX = tile[0], TY = tile[1];
  if (TX <= 1 && TY <= 1) return bx + nb[0] * (by + nb[1] * bz);
  const int ntx = (nb[0] + TX - 1) / TX, nty = (nb[1] + TY - 1) / TY;
  return (((bz * nty + by / TY) * ntx + bx / TX) * TY + by % TY) * TX + bx % TX;
}
__host__ __device__ inline long long cell_count(const int nb[3], const int tile[2]) {
  const int TX = tile[0] > 1 ? tile[0] : 1, TY = tile[1] > 1 ? tile[1] : 1;
  return (long long)(((nb[0] + TX - 1) / TX) * TX) * (((nb[1] + TY - 1) / TY) * TY) * nb[2];
}

}  // namespace sedi
