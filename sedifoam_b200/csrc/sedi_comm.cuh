// sedi_comm.cuh -- spatial decomposition of the particle box over GPUs (one process per GPU) and the collectives the
// hot path needs.  Mirrors what LAMMPS' Comm does for sediFoam (`processors Px Py Pz`, `communicate single vel yes`,
// newton off => forward ghost communication only) with NCCL in place of MPI.
//
// The library does not link NCCL: sedi_comm_init() dlopen()s the libnccl.so.2 that is already in the process
// (torch's) and bootstraps its own communicator from a ncclUniqueId handed in through the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include "lmp_script.hpp"

namespace sedi {

class Engine;

struct Comm {
  int rank, nranks;
  int grid[3], coord[3];
  void *nccl_lib;
  void *nccl_comm;
  Comm() : rank(0), nranks(1), nccl_lib(0), nccl_comm(0) { grid[0] = grid[1] = grid[2] = 1; coord[0] = coord[1] = coord[2] = 0; }

  // sub-domain of this rank (lammps_get_local_domain, library.cpp:222-240); `shell` widens it by the ghost cut-off
  // in dimensions that are split over ranks
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

  int init(Engine &e, int rank_, int nranks_, const void *uid, int uid_bytes, const int *procgrid);
  static int unique_id(void *out, int cap);
  void destroy();
  void barrier();
  void allreduce_max_host(double *v, int n);
  void allreduce_sum_host(double *v, int n);
  long long allreduce_sum_ll(long long v);
  void allgather_int(int v, int *out);
  void allreduce_sum_dev(double *p, size_t n, cudaStream_t s);
  void exchange_and_borders(Engine &e);
};

}  // namespace sedi
