// sedi_comm.cuh -- spatial decomposition of the particle box over GPUs (one process per GPU) and the collectives the
// hot path needs.  Mirrors what LAMMPS' Comm does for sediFoam (`processors Px Py Pz`, `communicate single vel yes`,
// newton off => forward ghost communication only) with NCCL in place of MPI.
//
// The library does not link NCCL: sedi_comm_init() dlopen()s the libnccl.so.2 that is already in the process
// (torch's) and bootstraps its own communicator from a ncclUniqueId handed in through the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "lmp_script.hpp"
#include "sedi_halo.cuh"

namespace sedi {

class Engine;

// ---- pure host logic of the brick decomposition (also exported through the C-ABI for the CPU tests) -----------------
// LAMMPS-style processor grid: the factorisation of nranks with the smallest sub-domain surface
inline void decomp_auto_grid(int nranks, const double len[3], int grid[3]) {
  double best = 1e300;
  grid[0] = nranks; grid[1] = 1; grid[2] = 1;
  for (int gx = 1; gx <= nranks; gx++) {
    if (nranks % gx) continue;
    for (int gy = 1; gy <= nranks / gx; gy++) {
      if ((nranks / gx) % gy) continue;
      const int gz = nranks / gx / gy;
      const double sx = len[0] / gx, sy = len[1] / gy, sz = len[2] / gz;
      const double surf = sx * sy + sy * sz + sx * sz;
      if (surf < best * (1.0 - 1e-12)) { best = surf; grid[0] = gx; grid[1] = gy; grid[2] = gz; }
    }
  }
}
inline void decomp_coord(int rank, const int grid[3], int coord[3]) {
  coord[0] = rank % grid[0]; coord[1] = (rank / grid[0]) % grid[1]; coord[2] = rank / (grid[0] * grid[1]);
}
inline int decomp_rank(const int c[3], const int grid[3]) { return c[0] + grid[0] * (c[1] + grid[1] * c[2]); }
inline int decomp_owner(const double x[3], const double lo[3], const double hi[3], const int grid[3]) {
  int c[3];
  for (int d = 0; d < 3; d++) c[d] = owner_coord(x[d], lo[d], hi[d], grid[d]);
  return decomp_rank(c, grid);
}
struct LinkHost { int off[3]; int peer; double shift[3]; };
// neighbour links of one brick in canonical offset order (z slowest); a dimension that is not split has no links
inline std::vector<LinkHost> decomp_links(int rank, const int grid[3], const int periodic[3], const double prd[3]) {
  int coord[3];
  decomp_coord(rank, grid, coord);
  std::vector<LinkHost> out;
  for (int oz = -1; oz <= 1; oz++) for (int oy = -1; oy <= 1; oy++) for (int ox = -1; ox <= 1; ox++) {
    const int o[3] = {ox, oy, oz};
    if (!ox && !oy && !oz) continue;
    LinkHost l;
    int c[3];
    bool ok = true;
    for (int d = 0; d < 3 && ok; d++) {
      l.off[d] = o[d]; l.shift[d] = 0.0; c[d] = coord[d] + o[d];
      if (grid[d] == 1) { if (o[d]) ok = false; continue; }
      if (c[d] < 0) { if (!periodic[d]) ok = false; else { c[d] += grid[d]; l.shift[d] = prd[d]; } }
      else if (c[d] >= grid[d]) { if (!periodic[d]) ok = false; else { c[d] -= grid[d]; l.shift[d] = -prd[d]; } }
    }
    if (!ok) continue;
    l.peer = decomp_rank(c, grid);
    out.push_back(l);
  }
  return out;
}

template <class T> struct Buf;

struct Comm {
  int rank, nranks;
  int grid[3], coord[3];
  void *nccl_lib;
  void *nccl_comm;
  std::vector<LinkHost> links;
  std::vector<int> sendbase, sendcount, recvbase, recvcount;  // per link (send: my border rows; recv: my ghost rows)
  int total_send, total_recv;
  DecompDev dev;
  // device scratch (raw pointers managed in the implementation)
  int *d_small;            // 256 ints: counts etc.
  int *h_small;            // pinned mirror
  int *d_sendrows; size_t cap_sendrows;
  int *d_blockcnt, *d_blockoff, *d_blocksum; size_t cap_block;
  D4 *d_sendbuf, *d_recvbuf; size_t cap_sendbuf, cap_recvbuf;
  int *d_migrows; double *d_migsend, *d_migrecv; size_t cap_mig; int migcap, migrec;
  int *d_arr_nh, *d_arr_tag; D4 *d_arr_shear; size_t cap_arr;
  int *d_gcellid, *d_gcount, *d_gstart, *d_gfill, *d_gorder; size_t cap_g, cap_gcells;
  int narr_last;
  long long halo_calls;
  // NVLink peer-memory halo (CUDA IPC): the neighbours' particle arrays and every rank's signal array mapped here
  bool p2p;                       // true once the peer mappings are in place (SEDI_HALO=nccl keeps the NCCL path)
  unsigned long long *d_sig;      // my signal array [64]
  void *peer_base[64][7];         // per rank: posr[0], posr[1], velm[0], velm[1], omgt[0], omgt[1], sig (opened IPC mappings)
  void *exported[7];              // the local pointers the current mappings were made from
  std::vector<int> rstart;        // per link: first ghost row of my segment in the peer's arrays
  unsigned epoch;
  // ghost refresh fused into the sub-step kernel: row -> (link, position) entries and the push tables of both buffers, on the device
  unsigned char *d_bcnt; int *d_bpos, *d_bcount, *d_bfill; BorderEnt *d_bent; size_t cap_brow, cap_bent;
  PushTable *d_push[2];
  bool fused_push;                // SEDI_HALO_FUSED=0 keeps the separate k_halo_push launch
  void build_row_table(Engine &e);
  void make_push_table(PushTable &H, int buf) const;
  void setup_peer(Engine &e);
  void close_peer();
  Comm();

  double sublo(const SimConfig &c, int d, double shell) const {
    const double lo = c.boxlo[d], len = c.boxhi[d] - c.boxlo[d];
    if (grid[d] == 1) return lo;
    return lo + len * coord[d] / grid[d] - shell;
  }
  double subhi(const SimConfig &c, int d, double shell) const {
    const double lo = c.boxlo[d], len = c.boxhi[d] - c.boxlo[d];
    if (grid[d] == 1) return c.boxhi[d];
    return lo + len * (coord[d] + 1) / grid[d] + shell;
  }
  // periodic dimension handled by image codes inside one GPU (true) or by ghost rows from the peer (false)
  bool wraps(const SimConfig &c, int d) const { return c.periodic[d] && grid[d] == 1; }
  bool owns(const SimConfig &c, const double *x) const {
    for (int d = 0; d < 3; d++) if (owner_coord(x[d], c.boxlo[d], c.boxhi[d], grid[d]) != coord[d]) return false;
    return true;
  }

  int init(Engine &e, int rank_, int nranks_, const void *uid, int uid_bytes, const int *procgrid);
  static int unique_id(void *out, int cap);
  void destroy();
  void barrier();
  void allreduce_max_host(double *v, int n);
  void allreduce_sum_host(double *v, int n);
  long long allreduce_sum_ll(long long v);
  void allgather_int(int v, int *out);
  void allreduce_sum_dev(double *p, size_t n, cudaStream_t s);
  void setup_decomp(Engine &e);
  int migrate(Engine &e);                 // returns the number of particles that arrived
  void borders(Engine &e);                // send lists, ghost rows, ghost bins
  // per-sub-step ghost refresh (+ rebuild-flag consensus); pushed: the sub-step kernel has written the ghost rows itself
  void forward(Engine &e, int buf, bool with_flag, bool pushed = false);
};

}  // namespace sedi
