// sedi_comm_impl.cuh -- implementation of sedi::Comm (included at the end of sedi_engine.cu, needs Engine).
#pragma once

namespace sedi {

inline int Comm::init(Engine &, int rank_, int nranks_, const void *, int, const int *) {
  if (nranks_ != 1) fatal("multi-GPU communicator not built into this library version");
  rank = rank_; nranks = nranks_;
  return 0;
}
inline int Comm::unique_id(void *, int) { return 0; }
inline void Comm::destroy() {}
inline void Comm::barrier() {}
inline void Comm::allreduce_max_host(double *, int) {}
inline void Comm::allreduce_sum_host(double *, int) {}
inline long long Comm::allreduce_sum_ll(long long v) { return v; }
inline void Comm::allgather_int(int v, int *out) { out[rank] = v; }
inline void Comm::allreduce_sum_dev(double *, size_t, cudaStream_t) {}
inline void Comm::exchange_and_borders(Engine &) {}

}  // namespace sedi
