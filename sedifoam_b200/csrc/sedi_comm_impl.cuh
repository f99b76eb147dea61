// sedi_comm_impl.cuh -- implementation of sedi::Comm (included at the end of sedi_engine.cu, needs Engine).
// NCCL is reached through dlopen/dlsym so that libsedi_b200.so has no link-time dependency on it: a single-GPU
// host never loads it, and under torchrun the process-wide libnccl.so.2 (torch's) is reused.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace sedi {

namespace nccl_dyn {
typedef ncclResult_t (*GetUniqueId_t)(ncclUniqueId *);
typedef ncclResult_t (*CommInitRank_t)(ncclComm_t *, int, ncclUniqueId, int);
typedef ncclResult_t (*CommDestroy_t)(ncclComm_t);
typedef ncclResult_t (*AllReduce_t)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*AllGather_t)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*Send_t)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*Recv_t)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
typedef ncclResult_t (*Group_t)();
typedef const char *(*ErrStr_t)(ncclResult_t);
static void *lib = 0;
static GetUniqueId_t GetUniqueId; static CommInitRank_t CommInitRank; static CommDestroy_t CommDestroy;
static AllReduce_t AllReduce; static AllGather_t AllGather; static Send_t Send; static Recv_t Recv;
static Group_t GroupStart, GroupEnd; static ErrStr_t ErrStr;
static bool load() {
  if (lib) return true;
  const char *names[] = {getenv("SEDI_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (int k = 0; k < 3 && !lib; k++) {
    if (!names[k]) continue;
    lib = dlopen(names[k], RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy already in the process (torch's), if any
    if (!lib) lib = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
  }
  if (!lib) return false;
#define SEDI_SYM(n) n = (n##_t)dlsym(lib, "nccl" #n); if (!n) return false
  SEDI_SYM(GetUniqueId); SEDI_SYM(CommInitRank); SEDI_SYM(CommDestroy); SEDI_SYM(AllReduce); SEDI_SYM(AllGather); SEDI_SYM(Send); SEDI_SYM(Recv);
#undef SEDI_SYM
  GroupStart = (Group_t)dlsym(lib, "ncclGroupStart"); GroupEnd = (Group_t)dlsym(lib, "ncclGroupEnd");
  ErrStr = (ErrStr_t)dlsym(lib, "ncclGetErrorString");
  return GroupStart && GroupEnd && ErrStr;
}
}  // namespace nccl_dyn

#define NK(call)                                                                                               \
  do {                                                                                                         \
    ncclResult_t r_ = (call);                                                                                  \
    if (r_ != ncclSuccess) {                                                                                   \
      fprintf(stderr, "ERROR: NCCL failure %s at %s:%d: %s\n", #call, __FILE__, __LINE__, nccl_dyn::ErrStr(r_)); \
      fflush(stderr);                                                                                          \
      abort();                                                                                                 \
    }                                                                                                          \
  } while (0)

template <class T>
static void grow_raw(T *&p, size_t &cap, size_t need) {
  if (need <= cap) return;
  if (p) CK(cudaFree(p));
  cap = need + need / 4 + 256;
  CK(cudaMalloc((void **)&p, cap * sizeof(T)));
}

inline Comm::Comm()
    : rank(0), nranks(1), nccl_lib(0), nccl_comm(0), total_send(0), total_recv(0), d_small(0), h_small(0), d_sendrows(0), cap_sendrows(0),
      d_blockcnt(0), d_blockoff(0), d_blocksum(0), cap_block(0), d_sendbuf(0), d_recvbuf(0), cap_sendbuf(0), cap_recvbuf(0), d_migrows(0),
      d_migsend(0), d_migrecv(0), cap_mig(0), migcap(0), migrec(0), d_arr_nh(0), d_arr_tag(0), d_arr_shear(0), cap_arr(0), d_gcellid(0),
      d_gcount(0), d_gstart(0), d_gfill(0), d_gorder(0), cap_g(0), cap_gcells(0), narr_last(0), halo_calls(0), p2p(false), d_sig(0), epoch(0),
      d_bcnt(0), d_bpos(0), d_bcount(0), d_bfill(0), d_bent(0), cap_brow(0), cap_bent(0), fused_push(true) {
  d_push[0] = d_push[1] = 0;
  if (const char *f = getenv("SEDI_HALO_FUSED")) fused_push = atoi(f) != 0;
  memset(peer_base, 0, sizeof(peer_base)); memset(exported, 0, sizeof(exported));
  grid[0] = grid[1] = grid[2] = 1; coord[0] = coord[1] = coord[2] = 0;
  memset(&dev, 0, sizeof(dev));
}

inline int Comm::unique_id(void *out, int cap) {
  if (!nccl_dyn::load() || cap < (int)sizeof(ncclUniqueId)) return 0;
  ncclUniqueId id;
  if (nccl_dyn::GetUniqueId(&id) != ncclSuccess) return 0;
  memcpy(out, &id, sizeof(id));
  return (int)sizeof(id);
}

inline int Comm::init(Engine &e, int rank_, int nranks_, const void *uid, int uid_bytes, const int *procgrid) {
  rank = rank_; nranks = nranks_;
  if (nranks == 1) return 0;
  if (!nccl_dyn::load()) fatal("sedi_comm_init: libnccl.so.2 not found (set SEDI_NCCL_LIB)");
  if (uid_bytes != (int)sizeof(ncclUniqueId)) fatal("sedi_comm_init: bad ncclUniqueId size");
  e.need_device();
  ncclUniqueId id;
  memcpy(&id, uid, sizeof(id));
  ncclComm_t c;
  NK(nccl_dyn::CommInitRank(&c, nranks, id, rank));
  nccl_comm = (void *)c;
  CK(cudaMalloc((void **)&d_small, 1024 * sizeof(int)));
  CK(cudaMallocHost((void **)&h_small, 1024 * sizeof(int)));
  // processor grid: argument > script `processors` > smallest-surface factorisation
  const SimConfig &cf = e.cfg();
  int g[3] = {0, 0, 0};
  if (procgrid && procgrid[0] * procgrid[1] * procgrid[2] == nranks) { g[0] = procgrid[0]; g[1] = procgrid[1]; g[2] = procgrid[2]; }
  else if (cf.procgrid[0] * cf.procgrid[1] * cf.procgrid[2] == nranks) { g[0] = cf.procgrid[0]; g[1] = cf.procgrid[1]; g[2] = cf.procgrid[2]; }
  else { const double len[3] = {cf.boxhi[0] - cf.boxlo[0], cf.boxhi[1] - cf.boxlo[1], cf.boxhi[2] - cf.boxlo[2]}; decomp_auto_grid(nranks, len, g); }
  for (int d = 0; d < 3; d++) grid[d] = g[d];
  decomp_coord(rank, grid, coord);
  e.loaded = false; e.setup_done = false;
  return 0;
}

inline void Comm::destroy() {
  if (nccl_comm && nranks > 1) barrier();   // nobody unmaps / frees while a peer may still push into this rank
  close_peer();
  if (nccl_comm && nranks > 1) barrier();
  if (d_sig) { cudaFree(d_sig); d_sig = 0; }
  if (nccl_comm) { nccl_dyn::CommDestroy((ncclComm_t)nccl_comm); nccl_comm = 0; }
  void *ptrs[] = {d_small, d_sendrows, d_blockcnt, d_blockoff, d_blocksum, d_sendbuf, d_recvbuf, d_migrows, d_migsend, d_migrecv, d_arr_nh,
                  d_arr_tag, d_arr_shear, d_gcellid, d_gcount, d_gstart, d_gfill, d_gorder, d_bcnt, d_bpos, d_bcount, d_bfill, d_bent, d_push[0], d_push[1]};
  for (size_t k = 0; k < sizeof(ptrs) / sizeof(ptrs[0]); k++) if (ptrs[k]) cudaFree(ptrs[k]);
  if (h_small) cudaFreeHost(h_small);
}

// ---- small host-visible collectives (setup / rebuild time only) ----------------------------------------------------
static inline Engine *g_comm_engine = 0;
inline void Comm::allreduce_max_host(double *v, int n) {
  if (nranks == 1) return;
  Engine &e = *g_comm_engine;
  double *d = (double *)(d_small + 512);  // 256 doubles of scratch
  if (n > 200) fatal("allreduce_max_host: too many values");
  CK(cudaMemcpyAsync(d, v, n * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  NK(nccl_dyn::AllReduce(d, d, n, ncclDouble, ncclMax, (ncclComm_t)nccl_comm, e.stream));
  CK(cudaMemcpyAsync(v, d, n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  CK(cudaStreamSynchronize(e.stream));
}
inline void Comm::allreduce_sum_host(double *v, int n) {
  if (nranks == 1) return;
  Engine &e = *g_comm_engine;
  double *d = (double *)(d_small + 512);
  if (n > 200) fatal("allreduce_sum_host: too many values");
  CK(cudaMemcpyAsync(d, v, n * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  NK(nccl_dyn::AllReduce(d, d, n, ncclDouble, ncclSum, (ncclComm_t)nccl_comm, e.stream));
  CK(cudaMemcpyAsync(v, d, n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  CK(cudaStreamSynchronize(e.stream));
}
inline long long Comm::allreduce_sum_ll(long long v) {
  if (nranks == 1) return v;
  double d = (double)v;
  allreduce_sum_host(&d, 1);
  return (long long)(d + 0.5);
}
inline void Comm::allgather_int(int v, int *out) {
  if (nranks == 1) { out[0] = v; return; }
  Engine &e = *g_comm_engine;
  if (nranks > 256) fatal("allgather_int: too many ranks");
  CK(cudaMemcpyAsync(d_small, &v, sizeof(int), cudaMemcpyHostToDevice, e.stream));
  NK(nccl_dyn::AllGather(d_small, d_small + 256, 1, ncclInt32, (ncclComm_t)nccl_comm, e.stream));
  CK(cudaMemcpyAsync(out, d_small + 256, nranks * sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  CK(cudaStreamSynchronize(e.stream));
}
inline void Comm::barrier() {
  if (nranks == 1) return;
  double z = 0.0;
  allreduce_sum_host(&z, 1);
}
inline void Comm::allreduce_sum_dev(double *p, size_t n, cudaStream_t s) {
  if (nranks == 1) return;
  NK(nccl_dyn::AllReduce(p, p, n, ncclDouble, ncclSum, (ncclComm_t)nccl_comm, s));
}

// ---- decomposition tables (after cut-offs are known) -----------------------------------------------------------------
inline void Comm::setup_decomp(Engine &e) {
  const SimConfig &c = e.cfg();
  const double prd[3] = {c.boxhi[0] - c.boxlo[0], c.boxhi[1] - c.boxlo[1], c.boxhi[2] - c.boxlo[2]};
  links = decomp_links(rank, grid, c.periodic, prd);
  if ((int)links.size() > MAX_LINKS) fatal("too many neighbour links");
  memset(&dev, 0, sizeof(dev));
  for (int d = 0; d < 3; d++) {
    dev.grid[d] = grid[d]; dev.coord[d] = coord[d]; dev.periodic[d] = c.periodic[d];
    dev.boxlo[d] = c.boxlo[d]; dev.boxhi[d] = c.boxhi[d]; dev.sublo[d] = sublo(c, d, 0.0); dev.subhi[d] = subhi(c, d, 0.0);
    if (grid[d] > 1 && (dev.subhi[d] - dev.sublo[d]) < e.cutneighmax) fatal("sub-domain thinner than the ghost cut-off: use fewer GPUs in this dimension");
  }
  dev.cutghost = e.cutneighmax;
  dev.nlinks = (int)links.size();
  for (int k = 0; k < 27; k++) dev.linkof[k] = -1;
  for (int L = 0; L < dev.nlinks; L++) {
    for (int d = 0; d < 3; d++) { dev.off[L][d] = links[L].off[d]; dev.shift[L][d] = links[L].shift[d]; }
    dev.linkof[(links[L].off[0] + 1) + 3 * (links[L].off[1] + 1) + 9 * (links[L].off[2] + 1)] = L;
  }
  sendbase.assign(dev.nlinks, 0); sendcount.assign(dev.nlinks, 0); recvbase.assign(dev.nlinks, 0); recvcount.assign(dev.nlinks, 0);
  // ghost bins share the geometry of the owned bins
  const long long nc = e.ncells_bin();
  if ((size_t)nc + 2 > cap_gcells) {
    if (d_gcount) { CK(cudaFree(d_gcount)); CK(cudaFree(d_gstart)); CK(cudaFree(d_gfill)); }
    cap_gcells = (size_t)nc + 2;
    CK(cudaMalloc((void **)&d_gcount, cap_gcells * sizeof(int)));
    CK(cudaMalloc((void **)&d_gstart, cap_gcells * sizeof(int)));
    CK(cudaMalloc((void **)&d_gfill, cap_gcells * sizeof(int)));
  }
  rstart.assign(dev.nlinks, 0);
  setup_peer(e);
}

inline void Comm::close_peer() {
  for (int r = 0; r < 64; r++) for (int k = 0; k < 7; k++) if (peer_base[r][k] && r != rank) { cudaIpcCloseMemHandle(peer_base[r][k]); peer_base[r][k] = 0; }
  p2p = false;
}

// map the neighbours' particle arrays and all signal arrays into this process (cudaIpc*), once per (re)load
inline void Comm::setup_peer(Engine &e) {
  const char *mode = getenv("SEDI_HALO");
  if (mode && !strcmp(mode, "nccl")) { p2p = false; return; }
  if (nranks > 64) { p2p = false; return; }
  if (!d_sig) { CK(cudaMalloc((void **)&d_sig, 65 * sizeof(unsigned long long))); CK(cudaMemset(d_sig, 0, 65 * sizeof(unsigned long long))); epoch = 0; }   // [64] = this rank's exchange counter
  void *mine[7] = {e.posr[0].p, e.posr[1].p, e.velm[0].p, e.velm[1].p, e.omgt[0].p, e.omgt[1].p, d_sig};
  bool same = p2p;
  for (int k = 0; k < 7; k++) if (mine[k] != exported[k]) same = false;
  // every rank must take the same branch: agree through a max-reduction
  double chg = same ? 0.0 : 1.0;
  allreduce_max_host(&chg, 1);
  if (chg == 0.0) return;
  close_peer();
  struct Pack { cudaIpcMemHandle_t h[7]; };
  Pack my;
  int ok = 1;
  for (int k = 0; k < 7; k++) if (cudaIpcGetMemHandle(&my.h[k], mine[k]) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  Pack *d_all;
  std::vector<Pack> all(nranks);
  CK(cudaMalloc((void **)&d_all, (size_t)(nranks + 1) * sizeof(Pack)));
  CK(cudaMemcpyAsync(d_all + nranks, &my, sizeof(Pack), cudaMemcpyHostToDevice, e.stream));
  NK(nccl_dyn::AllGather(d_all + nranks, d_all, sizeof(Pack), ncclChar, (ncclComm_t)nccl_comm, e.stream));
  CK(cudaMemcpyAsync(all.data(), d_all, (size_t)nranks * sizeof(Pack), cudaMemcpyDeviceToHost, e.stream));
  CK(cudaStreamSynchronize(e.stream));
  CK(cudaFree(d_all));
  std::vector<char> need(nranks, 0);
  for (size_t L = 0; L < links.size(); L++) need[links[L].peer] = 1;
  for (int r = 0; r < nranks && ok; r++) {
    if (r == rank) { for (int k = 0; k < 7; k++) peer_base[r][k] = mine[k]; continue; }
    for (int k = 0; k < 7 && ok; k++) {
      if (k < 6 && !need[r]) continue;   // particle arrays only of link peers; the signal array of everybody
      if (cudaIpcOpenMemHandle(&peer_base[r][k], all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); peer_base[r][k] = 0; ok = 0; }
    }
  }
  double okd = ok ? 1.0 : 0.0, neg = -okd;
  allreduce_max_host(&neg, 1);          // min over ranks
  if (neg != -1.0) {
    if (rank == 0) fprintf(stderr, "libsedi_b200: CUDA IPC peer mapping unavailable, ghost halo falls back to NCCL send/recv\n");
    close_peer();
    return;
  }
  for (int k = 0; k < 7; k++) exported[k] = mine[k];
  p2p = true;
}

static inline int link_with_offset(const std::vector<LinkHost> &links, int ox, int oy, int oz) {
  for (size_t L = 0; L < links.size(); L++) if (links[L].off[0] == ox && links[L].off[1] == oy && links[L].off[2] == oz) return (int)L;
  return -1;
}

// exchange one int per link with the matching link of the peer: out[L] -> peer, in[Lr] <- peer (canonical order)
static inline void exchange_counts(Comm &cm, Engine &e, const int *d_out, int *d_in) {
  NK(nccl_dyn::GroupStart());
  for (int oz = -1; oz <= 1; oz++) for (int oy = -1; oy <= 1; oy++) for (int ox = -1; ox <= 1; ox++) {
    if (!ox && !oy && !oz) continue;
    const int Ls = link_with_offset(cm.links, ox, oy, oz);
    if (Ls >= 0) NK(nccl_dyn::Send(d_out + Ls, 1, ncclInt32, cm.links[Ls].peer, (ncclComm_t)cm.nccl_comm, e.stream));
    const int Lr = link_with_offset(cm.links, -ox, -oy, -oz);
    if (Lr >= 0) NK(nccl_dyn::Recv(d_in + Lr, 1, ncclInt32, cm.links[Lr].peer, (ncclComm_t)cm.nccl_comm, e.stream));
  }
  NK(nccl_dyn::GroupEnd());
}

// ---- Comm::exchange(): owned particles that left this brick move to the neighbour that now owns them ---------------
inline int Comm::migrate(Engine &e) {
  const SimConfig &c = e.cfg();
  const int nl = e.nlocal, n_old = e.nlocal + e.nghost, T = 256;
  const int NL = dev.nlinks;
  if (nl) k_pbc_wrap<<<cdiv(nl, T), T, 0, e.stream>>>(e.posr[e.cur].p, e.omgt[e.cur].p, nl, c.periodic[0], c.periodic[1], c.periodic[2],
                                                      c.boxlo[0], c.boxlo[1], c.boxlo[2], c.boxhi[0], c.boxhi[1], c.boxhi[2]);
  const int npl = e.hist_alloc ? 16 : 12;   // the history-force state (enhancedCloud.C:197-234) migrates with its particle
  migrec = 12 + npl + 3 * c.nwalls + 2 + 4 * MIG_MAXH;
  const int want_cap = std::max(2048, nl / 40);
  if ((size_t)want_cap * MAX_LINKS > cap_mig) {
    if (d_migrows) { CK(cudaFree(d_migrows)); CK(cudaFree(d_migsend)); CK(cudaFree(d_migrecv)); }
    cap_mig = (size_t)want_cap * MAX_LINKS; migcap = want_cap;
    CK(cudaMalloc((void **)&d_migrows, cap_mig * sizeof(int)));
    CK(cudaMalloc((void **)&d_migsend, cap_mig * (size_t)(28 + 3 * MAX_WALLS + 2 + 4 * MIG_MAXH) * sizeof(double)));
    CK(cudaMalloc((void **)&d_migrecv, cap_mig * (size_t)(28 + 3 * MAX_WALLS + 2 + 4 * MIG_MAXH) * sizeof(double)));
  }
  int *d_cnt = d_small, *d_rcnt = d_small + 32, *d_err = d_small + 64;
  CK(cudaMemsetAsync(d_small, 0, 96 * sizeof(int), e.stream));
  if (nl) k_mig_classify<<<cdiv(nl, T), T, 0, e.stream>>>(e.posr[e.cur].p, e.omgt[e.cur].p, nl, dev, e.leave.p, d_cnt, migcap, d_migrows, d_err);
  exchange_counts(*this, e, d_cnt, d_rcnt);
  CK(cudaMemcpyAsync(h_small, d_small, 96 * sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  CK(cudaStreamSynchronize(e.stream));
  if (h_small[64]) { fprintf(stderr, "migration error flags 0x%x on rank %d\n", h_small[64], rank); fatal("particle moved further than one sub-domain or migration buffer overflow"); }
  std::vector<int> scnt(NL), rcnt(NL), roff(NL);
  int narr = 0, nleave = 0;
  for (int L = 0; L < NL; L++) { scnt[L] = h_small[L]; rcnt[L] = h_small[32 + L]; nleave += scnt[L]; }
  // receive order = canonical sender order
  for (int oz = -1; oz <= 1; oz++) for (int oy = -1; oy <= 1; oy++) for (int ox = -1; ox <= 1; ox++) {
    if (!ox && !oy && !oz) continue;
    const int Lr = link_with_offset(links, -ox, -oy, -oz);
    if (Lr >= 0) { roff[Lr] = narr; narr += rcnt[Lr]; }
  }
  long long moved = allreduce_sum_ll(nleave);
  narr_last = narr;
  if (moved == 0) return 0;
  if (n_old + narr > e.npad) fatal("row capacity exceeded by migrating particles (raise SEDI_ROW_SLACK)");
  if ((size_t)narr > cap_mig) fatal("migration receive buffer overflow");
  // pack
  Ell &Lo = e.ell[e.ecur];
  MigPlanes M;
  memset(&M, 0, sizeof(M));
  M.nwalls = c.nwalls; M.npad = Lo.npad; M.rec = migrec; M.npl = npl;
  M.posr = e.posr[e.cur].p; M.velm = e.velm[e.cur].p; M.omgt = e.omgt[e.cur].p;
  Plane2 *groups[] = {e.fdrag, e.dudt, e.vold, e.uold};
  for (int g = 0; g < 4; g++) for (int d = 0; d < 3; d++) M.pl[3 * g + d] = groups[g][d].get();
  if (e.hist_alloc) for (int d = 0; d < 4; d++) M.pl[12 + d] = e.hist[d].get();
  for (int w = 0; w < c.nwalls; w++) for (int d = 0; d < 3; d++) M.ws[3 * w + d] = e.wshear[w][d].get();
  M.foam = e.foam[e.icur].p; M.wmask = e.wmask[e.icur].p;
  if (Lo.valid) { M.nn = Lo.nn.p; M.nbr = Lo.nbr.p; M.tmask = Lo.tmask.p; M.shear = Lo.shear.p; }
  M.tag2idx = e.tag2idx.p; M.maxtag = e.maxtag;
  k_mig_pack<<<cdiv((long long)NL * migcap, T), T, 0, e.stream>>>(NL, migcap, d_cnt, d_migrows, M, d_migsend, d_err);
  // payload
  NK(nccl_dyn::GroupStart());
  for (int oz = -1; oz <= 1; oz++) for (int oy = -1; oy <= 1; oy++) for (int ox = -1; ox <= 1; ox++) {
    if (!ox && !oy && !oz) continue;
    const int Ls = link_with_offset(links, ox, oy, oz);
    if (Ls >= 0 && scnt[Ls]) NK(nccl_dyn::Send(d_migsend + (size_t)Ls * migcap * migrec, (size_t)scnt[Ls] * migrec, ncclDouble, links[Ls].peer, (ncclComm_t)nccl_comm, e.stream));
    const int Lr = link_with_offset(links, -ox, -oy, -oz);
    if (Lr >= 0 && rcnt[Lr]) NK(nccl_dyn::Recv(d_migrecv + (size_t)roff[Lr] * migrec, (size_t)rcnt[Lr] * migrec, ncclDouble, links[Lr].peer, (ncclComm_t)nccl_comm, e.stream));
  }
  NK(nccl_dyn::GroupEnd());
  if (narr) {
    if ((size_t)narr > cap_arr) {
      if (d_arr_nh) { CK(cudaFree(d_arr_nh)); CK(cudaFree(d_arr_tag)); CK(cudaFree(d_arr_shear)); }
      cap_arr = (size_t)narr + narr / 2 + 1024;
      CK(cudaMalloc((void **)&d_arr_nh, cap_arr * sizeof(int)));
      CK(cudaMalloc((void **)&d_arr_tag, cap_arr * MIG_MAXH * sizeof(int)));
      CK(cudaMalloc((void **)&d_arr_shear, cap_arr * MIG_MAXH * sizeof(D4)));
    }
    MigDst Dd;
    memset(&Dd, 0, sizeof(Dd));
    Dd.nwalls = c.nwalls; Dd.rec = migrec; Dd.npl = npl;
    Dd.posr = e.posr[e.cur].p; Dd.velm = e.velm[e.cur].p; Dd.omgt = e.omgt[e.cur].p;
    for (int g = 0; g < 4; g++) for (int d = 0; d < 3; d++) Dd.pl[3 * g + d] = groups[g][d].get();
    if (e.hist_alloc) for (int d = 0; d < 4; d++) Dd.pl[12 + d] = e.hist[d].get();
    for (int w = 0; w < c.nwalls; w++) for (int d = 0; d < 3; d++) Dd.ws[3 * w + d] = e.wshear[w][d].get();
    Dd.foam = e.foam[e.icur].p; Dd.wmask = e.wmask[e.icur].p; Dd.leave = e.leave.p;
    Dd.arr_nh = d_arr_nh; Dd.arr_tag = d_arr_tag; Dd.arr_shear = d_arr_shear;
    k_mig_unpack<<<cdiv(narr, T), T, 0, e.stream>>>(narr, d_migrecv, n_old, Dd);
  }
  CK(cudaMemcpyAsync(h_small, d_small, 96 * sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  CK(cudaStreamSynchronize(e.stream));
  if (h_small[64]) fatal("migrating particle carries more than 16 touching contacts");
  e.launches += 4;
  return narr;
}

// ---- Comm::borders(): deterministic send lists, ghost rows, ghost bins ------------------------------------------------
inline void Comm::borders(Engine &e) {
  const int nl = e.nlocal, T = 256, NL = dev.nlinks;
  const int nblocks = std::max(1, cdiv(nl, T));
  const size_t nb = (size_t)NL * nblocks + 1;
  if (nb > cap_block) {
    if (d_blockcnt) { CK(cudaFree(d_blockcnt)); CK(cudaFree(d_blockoff)); CK(cudaFree(d_blocksum)); }
    cap_block = nb + nb / 4 + 64;
    CK(cudaMalloc((void **)&d_blockcnt, cap_block * sizeof(int)));
    CK(cudaMalloc((void **)&d_blockoff, cap_block * sizeof(int)));
    CK(cudaMalloc((void **)&d_blocksum, (cap_block / SCAN_ITEMS + 2) * sizeof(int)));
  }
  CK(cudaMemsetAsync(d_blockcnt, 0, nb * sizeof(int), e.stream));
  k_border_count<<<nblocks, T, 0, e.stream>>>(e.posr[e.cur].p, nl, dev, nblocks, d_blockcnt);
  const int nscan = (int)nb, nblk = cdiv(nscan, SCAN_ITEMS);
  k_scan_local<<<nblk, 1024, 0, e.stream>>>(d_blockcnt, d_blockoff, nscan, d_blocksum);
  k_scan_sums<<<1, 1024, 0, e.stream>>>(d_blocksum, nblk);
  k_scan_add<<<cdiv(nscan, T), T, 0, e.stream>>>(d_blockoff, nscan, d_blocksum, 0);
  for (int L = 0; L <= NL; L++) CK(cudaMemcpyAsync(h_small + 128 + L, d_blockoff + (size_t)L * nblocks, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  CK(cudaStreamSynchronize(e.stream));
  total_send = h_small[128 + NL];
  for (int L = 0; L < NL; L++) { sendbase[L] = h_small[128 + L]; sendcount[L] = h_small[128 + L + 1] - h_small[128 + L]; }
  grow_raw(d_sendrows, cap_sendrows, (size_t)total_send + 1);
  grow_raw(d_sendbuf, cap_sendbuf, 3 * (size_t)total_send + 3);
  k_border_fill<<<nblocks, T, 0, e.stream>>>(e.posr[e.cur].p, nl, dev, nblocks, d_blockoff, d_sendrows);
  // counts -> ghost segments
  int *d_cnt = d_small, *d_rcnt = d_small + 32;
  for (int L = 0; L < NL; L++) h_small[L] = sendcount[L];
  CK(cudaMemcpyAsync(d_cnt, h_small, 32 * sizeof(int), cudaMemcpyHostToDevice, e.stream));
  exchange_counts(*this, e, d_cnt, d_rcnt);
  CK(cudaMemcpyAsync(h_small + 32, d_rcnt, 32 * sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  CK(cudaStreamSynchronize(e.stream));
  total_recv = 0;
  for (int oz = -1; oz <= 1; oz++) for (int oy = -1; oy <= 1; oy++) for (int ox = -1; ox <= 1; ox++) {
    if (!ox && !oy && !oz) continue;
    const int Lr = link_with_offset(links, -ox, -oy, -oz);
    if (Lr >= 0) { recvcount[Lr] = h_small[32 + Lr]; recvbase[Lr] = total_recv; total_recv += recvcount[Lr]; }
  }
  if (nl + total_recv > e.npad) fatal("row capacity exceeded by ghost particles (raise SEDI_ROW_SLACK)");
  if (nl + total_recv > (int)NB_IDX_MASK) fatal("too many rows for the 25-bit neighbour index");
  grow_raw(d_recvbuf, cap_recvbuf, 3 * (size_t)total_recv + 3);
  e.nghost = total_recv;
  if (p2p) {  // tell every peer where, in my arrays, the ghost segment it fills starts
    for (int L = 0; L < NL; L++) h_small[L] = nl + recvbase[L];
    CK(cudaMemcpyAsync(d_cnt, h_small, 32 * sizeof(int), cudaMemcpyHostToDevice, e.stream));
    exchange_counts(*this, e, d_cnt, d_rcnt);
    CK(cudaMemcpyAsync(h_small + 32, d_rcnt, 32 * sizeof(int), cudaMemcpyDeviceToHost, e.stream));
    CK(cudaStreamSynchronize(e.stream));
    for (int L = 0; L < NL; L++) rstart[L] = h_small[32 + L];
    if (fused_push) build_row_table(e);
  }
  forward(e, e.cur, false);
  // bin the ghost rows (index lists, no physical re-ordering)
  const long long nc = e.ncells_bin();
  const SimConfig &c = e.cfg();
  if ((size_t)total_recv + 1 > cap_g) {
    if (d_gcellid) { CK(cudaFree(d_gcellid)); CK(cudaFree(d_gorder)); }
    cap_g = (size_t)total_recv + total_recv / 4 + 1024;
    CK(cudaMalloc((void **)&d_gcellid, cap_g * sizeof(int)));
    CK(cudaMalloc((void **)&d_gorder, cap_g * sizeof(int)));
  }
  CK(cudaMemsetAsync(d_gcount, 0, (size_t)(nc + 2) * sizeof(int), e.stream));
  CK(cudaMemsetAsync(d_gfill, 0, (size_t)(nc + 2) * sizeof(int), e.stream));
  if (total_recv)
    k_wrap_bin<<<cdiv(total_recv, T), T, 0, e.stream>>>(e.posr[e.cur].p + nl, e.omgt[e.cur].p + nl, total_recv, e.bin, 0, 0, 0, 0, 0, 0, d_gcellid, d_gcount,
                                                        0, 0, 0, 0, 0, 0, (const int *)0, -1, 0, (int *)0);
  const int ns2 = (int)(nc + 2), nb2 = cdiv(ns2, SCAN_ITEMS);
  k_scan_local<<<nb2, 1024, 0, e.stream>>>(d_gcount, d_gstart, ns2, e.blocksum.p);
  k_scan_sums<<<1, 1024, 0, e.stream>>>(e.blocksum.p, nb2);
  k_scan_add<<<cdiv(ns2, T), T, 0, e.stream>>>(d_gstart, ns2, e.blocksum.p, 0);
  if (total_recv) {
    k_bin_scatter<<<cdiv(total_recv, T), T, 0, e.stream>>>(d_gcellid, total_recv, d_gstart, d_gfill, d_gorder);
    k_cell_sort<<<cdiv(nc, T), T, 0, e.stream>>>(d_gstart, (int)nc, total_recv, d_gorder, e.omgt[e.cur].p + nl);
  }
  e.launches += 10;
}

// ---- Comm::forward_comm(): ghost x, v, omega every sub-step, plus the rebuild-flag consensus ---------------------------
inline void Comm::make_push_table(PushTable &H, int buf) const {
  const int NL = dev.nlinks;
  memset(&H, 0, sizeof(H));
  H.nlinks = NL;
  for (int L = 0; L < NL; L++) {
    H.base[L] = sendbase[L]; H.rstart[L] = rstart[L];
    for (int d = 0; d < 3; d++) H.shift[L][d] = links[L].shift[d];
    const int pr = links[L].peer;
    H.rposr[L] = (D4 *)peer_base[pr][0 + buf]; H.rvelm[L] = (D4 *)peer_base[pr][2 + buf]; H.romgt[L] = (D4 *)peer_base[pr][4 + buf];
  }
  H.base[NL] = total_send;
}

// row -> (link, position) entries of the border rows, and the two push tables, for the ghost refresh fused into the sub-step kernel
inline void Comm::build_row_table(Engine &e) {
  const int nl = e.nlocal, T = 256, NL = dev.nlinks;
  if ((size_t)e.npad + 1 > cap_brow) {
    if (d_bcnt) { CK(cudaFree(d_bcnt)); CK(cudaFree(d_bpos)); CK(cudaFree(d_bcount)); CK(cudaFree(d_bfill)); }
    cap_brow = (size_t)e.npad + 1;
    CK(cudaMalloc((void **)&d_bcnt, cap_brow)); CK(cudaMalloc((void **)&d_bpos, cap_brow * sizeof(int)));
    CK(cudaMalloc((void **)&d_bcount, cap_brow * sizeof(int))); CK(cudaMalloc((void **)&d_bfill, cap_brow * sizeof(int)));
  }
  if ((size_t)total_send + 1 > cap_bent) {
    if (d_bent) CK(cudaFree(d_bent));
    cap_bent = (size_t)total_send + total_send / 4 + 1024;
    CK(cudaMalloc((void **)&d_bent, cap_bent * sizeof(BorderEnt)));
  }
  if (!d_push[0]) { CK(cudaMalloc((void **)&d_push[0], sizeof(PushTable))); CK(cudaMalloc((void **)&d_push[1], sizeof(PushTable))); }
  CK(cudaMemsetAsync(d_bcount, 0, cap_brow * sizeof(int), e.stream));
  CK(cudaMemsetAsync(d_bfill, 0, cap_brow * sizeof(int), e.stream));
  CK(cudaMemsetAsync(d_bcnt, 0, cap_brow, e.stream));
  if (total_send) k_border_rowcount<<<cdiv(total_send, T), T, 0, e.stream>>>(d_sendrows, total_send, d_bcount);
  const int nscan = nl + 1, nblk = cdiv(nscan, SCAN_ITEMS);
  e.blocksum.ensure((size_t)nblk + 1);
  k_scan_local<<<nblk, 1024, 0, e.stream>>>(d_bcount, d_bpos, nscan, e.blocksum.p);
  k_scan_sums<<<1, 1024, 0, e.stream>>>(e.blocksum.p, nblk);
  k_scan_add<<<cdiv(nscan, T), T, 0, e.stream>>>(d_bpos, nscan, e.blocksum.p, 0);
  HaloTable H;
  memset(&H, 0, sizeof(H));
  H.nlinks = NL;
  for (int L = 0; L < NL; L++) H.base[L] = sendbase[L];
  H.base[NL] = total_send;
  if (total_send) k_border_rowfill<<<cdiv(total_send, T), T, 0, e.stream>>>(d_sendrows, H, d_bpos, d_bfill, d_bent);
  if (nl) k_border_pack_cnt<<<cdiv(nl, T), T, 0, e.stream>>>(nl, d_bcount, d_bcnt, d_small + 64);
  for (int b = 0; b < 2; b++) {
    PushTable P;
    make_push_table(P, b);
    CK(cudaMemcpyAsync(d_push[b], &P, sizeof(PushTable), cudaMemcpyHostToDevice, e.stream));
    CK(cudaStreamSynchronize(e.stream));   // P is a stack object
  }
  e.launches += 6;
}

inline void Comm::forward(Engine &e, int buf, bool with_flag, bool pushed) {
  const int T = 256, NL = dev.nlinks;
  if (p2p) {
    halo_calls++;
    PushTable H;
    make_push_table(H, buf);
    if (total_send && !pushed) { k_halo_push<<<cdiv(total_send, T), T, 0, e.stream>>>(e.posr[buf].p, e.velm[buf].p, e.omgt[buf].p, d_sendrows, H); e.launches++; }
    SignalTable S;
    memset(&S, 0, sizeof(S));
    S.nranks = nranks; S.me = rank;
    for (int r = 0; r < nranks; r++) S.rsig[r] = (unsigned long long *)peer_base[r][6];
    k_halo_signal_wait<<<1, 64, 0, e.stream>>>(S, d_sig, e.ctrl.p, with_flag ? 1 : 0);
    e.launches++;
    return;
  }
  HaloTable H;
  memset(&H, 0, sizeof(H));
  H.nlinks = NL;
  for (int L = 0; L < NL; L++) { H.base[L] = sendbase[L]; for (int d = 0; d < 3; d++) H.shift[L][d] = links[L].shift[d]; }
  H.base[NL] = total_send;
  if (total_send) k_halo_pack<<<cdiv(total_send, T), T, 0, e.stream>>>(e.posr[buf].p, e.velm[buf].p, e.omgt[buf].p, d_sendrows, H, d_sendbuf);
  NK(nccl_dyn::GroupStart());
  for (int oz = -1; oz <= 1; oz++) for (int oy = -1; oy <= 1; oy++) for (int ox = -1; ox <= 1; ox++) {
    if (!ox && !oy && !oz) continue;
    const int Ls = link_with_offset(links, ox, oy, oz);
    if (Ls >= 0 && sendcount[Ls]) NK(nccl_dyn::Send(d_sendbuf + 3 * (size_t)sendbase[Ls], 12 * (size_t)sendcount[Ls], ncclDouble, links[Ls].peer, (ncclComm_t)nccl_comm, e.stream));
    const int Lr = link_with_offset(links, -ox, -oy, -oz);
    if (Lr >= 0 && recvcount[Lr]) NK(nccl_dyn::Recv(d_recvbuf + 3 * (size_t)recvbase[Lr], 12 * (size_t)recvcount[Lr], ncclDouble, links[Lr].peer, (ncclComm_t)nccl_comm, e.stream));
  }
  if (with_flag) NK(nccl_dyn::AllReduce(e.ctrl.p, e.ctrl.p, 1, ncclInt32, ncclMax, (ncclComm_t)nccl_comm, e.stream));
  NK(nccl_dyn::GroupEnd());
  if (total_recv) k_halo_unpack<<<cdiv(total_recv, T), T, 0, e.stream>>>(d_recvbuf, total_recv, e.nlocal, e.posr[buf].p, e.velm[buf].p, e.omgt[buf].p);
  e.launches += 2;
  halo_calls++;
}

}  // namespace sedi
