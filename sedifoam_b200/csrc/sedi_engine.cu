// sedi_engine.cu -- host-side engine of libsedi_b200.so: owns the device-resident particle state, drives the
// neighbour rebuild / fused DEM sub-step / coupling kernels on one CUDA stream, and exports the C-ABI declared in
// include/sedi_b200.h (the reference's interfaceToLammps/library.h boundary + the device-resident sedi_* API).
//
// What replaces what (reference file:line):
//   Engine::command        lmp_->input->one(line)                       lammpsFoam/softParticleCloud.C:85-115
//   Engine::run            lammps_step = "run n pre no post no"          interfaceToLammps/library.cpp:372-386
//                          + EXTERNAL Verlet::run / Neighbor / FixShearHistory
//   Engine::put_local      lammps_put_local_info                         interfaceToLammps/library.cpp:314-367
//   Engine::get_local      lammps_get_local_info                         interfaceToLammps/library.cpp:246-308
//   Engine::fluid_force... enhancedCloud::updateParticleUr/updateDragOnParticles/particleToEulerianField/calcTcFields
//                                                                        lammpsFoam/enhancedCloud.C:83-441, 911-980
// There is no CPU fallback: every compute entry point needs a CUDA device and aborts without one
// (reference error convention: print + abort, library.cpp:380-383).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <vector>
#include <limits>

#include "../../include/sedi_b200.h"
#include "../../include/lammps_shim/lammps.h"
#include "../../include/lammps_shim/input.h"
#include "lmp_script.hpp"
#include "sedi_device.cuh"
#include "sedi_neigh.cuh"
#include "sedi_step.cuh"
#include "sedi_wq.cuh"
#include "sedi_sell.cuh"
#include "sedi_couple.cuh"
#include "sedi_smooth.cuh"
#include "sedi_halo.cuh"
#include "sedi_comm.cuh"

#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) {                                                                                  \
      fprintf(stderr, "ERROR: CUDA failure %s at %s:%d: %s\n", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      fflush(stderr);                                                                                         \
      abort();                                                                                                \
    }                                                                                                         \
  } while (0)

namespace sedi {

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

template <class T>
struct Buf {  // grow-only device array
  T *p;
  size_t cap;
  Buf() : p(0), cap(0) {}
  void ensure(size_t n) {
    if (n <= cap) return;
    if (p) CK(cudaFree(p));
    size_t want = n + n / 8 + 64;
    CK(cudaMalloc((void **)&p, want * sizeof(T)));
    cap = want;
  }
  void ensure_keep(size_t n, size_t used, cudaStream_t s) {
    if (n <= cap) return;
    size_t want = n + n / 8 + 64;
    T *q;
    CK(cudaMalloc((void **)&q, want * sizeof(T)));
    if (p && used) CK(cudaMemcpyAsync(q, p, used * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (p) { CK(cudaStreamSynchronize(s)); CK(cudaFree(p)); }
    p = q; cap = want;
  }
  void release() { if (p) cudaFree(p); p = 0; cap = 0; }
};

template <class T>
struct Pinned {  // grow-only pinned host staging
  T *p;
  size_t cap;
  Pinned() : p(0), cap(0) {}
  void ensure(size_t n) {
    if (n <= cap) return;
    if (p) CK(cudaFreeHost(p));
    size_t want = n + n / 8 + 64;
    CK(cudaMallocHost((void **)&p, want * sizeof(T)));
    cap = want;
  }
  void release() { if (p) cudaFreeHost(p); p = 0; cap = 0; }
};

struct Plane2 {  // a per-particle plane with a permutation partner
  Buf<double> b[2];
  int cur;
  Plane2() : cur(0) {}
  double *get() { return b[cur].p; }
  double *alt() { return b[cur ^ 1].p; }
};

struct Ell {  // directed neighbour list + contact history, slot-major
  Buf<unsigned> nbr;
  Buf<int> nn, nt;        // granular entries [0, nn) and type-only entries [cap, cap + nt) of every row
  Buf<unsigned long long> tmask;
  Buf<D4> shear;
  int npad, cap, tcap;    // cap: granular slots (<= 64, each with a history quad); tcap: type-only slots
  bool valid;
  Ell() : npad(0), cap(0), tcap(0), valid(false) {}
};

// small pack / unpack kernels of the C-ABI boundary -------------------------------------------------------
__global__ void k_pack_local(const D4 *posr, const D4 *velm, const D4 *omgt, const int *foam, int n, double *x, double *v, int *tag,
                             int *foamid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const D4 p = posr[i], u = velm[i];
  x[3 * (size_t)i] = p.x; x[3 * (size_t)i + 1] = p.y; x[3 * (size_t)i + 2] = p.z;
  v[3 * (size_t)i] = u.x; v[3 * (size_t)i + 1] = u.y; v[3 * (size_t)i + 2] = u.z;
  tag[i] = bits_tag((unsigned long long)__double_as_longlong(omgt[i].w));
  foamid[i] = foam[i];
}

// lammps_put_local_info: the reference pairs the k-th smallest incoming tag with the k-th smallest local tag
// (library.cpp:343-366); for equal tag sets -- the only consistent use -- that is a lookup by tag.
__global__ void k_put_fdrag(int n, const double *fd, const int *tagin, const int *foamin, const int *tag2idx, int maxtag, double *f0,
                            double *f1, double *f2, int *foam, int *err) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int t = tagin[k];
  const int i = (t >= 0 && t <= maxtag) ? tag2idx[t] : -1;
  if (i < 0) { atomicOr(err, 1); return; }
  f0[i] = fd[3 * (size_t)k]; f1[i] = fd[3 * (size_t)k + 1]; f2[i] = fd[3 * (size_t)k + 2];
  if (foamin) foam[i] = foamin[k];
}

__global__ void k_unpack_state(const D4 *posr, const D4 *velm, const D4 *omgt, int n, double *x, double *v, double *w, double *radius,
                               double *rmass, int *tag, int *type, int *mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const D4 p = posr[i], u = velm[i], o = omgt[i];
  x[3 * (size_t)i] = p.x; x[3 * (size_t)i + 1] = p.y; x[3 * (size_t)i + 2] = p.z;
  v[3 * (size_t)i] = u.x; v[3 * (size_t)i + 1] = u.y; v[3 * (size_t)i + 2] = u.z;
  w[3 * (size_t)i] = o.x; w[3 * (size_t)i + 1] = o.y; w[3 * (size_t)i + 2] = o.z;
  radius[i] = p.w; rmass[i] = u.w;
  const unsigned long long b = (unsigned long long)__double_as_longlong(o.w);
  tag[i] = bits_tag(b); type[i] = bits_type(b); mask[i] = bits_mask(b);
}

__global__ void k_interleave3(const double *a, const double *b, const double *c, int n, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[3 * (size_t)i] = a[i]; out[3 * (size_t)i + 1] = b[i]; out[3 * (size_t)i + 2] = c[i];
}

__global__ void k_set_omega(int n, const int *tagin, const double *w, const int *tag2idx, int maxtag, D4 *omgt) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int t = tagin[k];
  if (t < 0 || t > maxtag) return;
  const int i = tag2idx[t];
  if (i < 0) return;
  D4 o = omgt[i];
  o.x = w[3 * (size_t)k]; o.y = w[3 * (size_t)k + 1]; o.z = w[3 * (size_t)k + 2];
  omgt[i] = o;
}

__global__ void k_fill_double(double *p, size_t n, double v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

class Engine;
static void g_comm_engine_set(Engine *e);
static const int NPLANES_BASE = 6;  // f, tq, fdrag, dudt, vold, uold (x3 each)

class Engine {
 public:
  Script script;
  int device;
  bool dev_ready, loaded, setup_done, params_dirty, cell_valid;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1, evk0, evk1, evt0, evt1, ev_get[2];
  bool prof_on;
  bool prof_split;                       // SEDI_PROF_SPLIT=1 (diagnostic, no graph): events around every sub-step kernel and every ghost exchange
  std::vector<cudaEvent_t> split_ev;
  double split_step_ms, split_halo_ms; long long split_n;
  double prof_ms;
  long long prof_steps;
  // particle rows
  int n, nlocal, nghost, npad, maxtag;   // n = nlocal + nghost rows; npad = capacity of the per-row arrays
  Buf<int> leave;                        // migration flags (multi-GPU)
  Buf<D4> posr[2], velm[2], omgt[2];
  int cur;
  Plane2 f[3], tq[3], fdrag[3], dudt[3], vold[3], uold[3], xhold[3];
  Plane2 wshear[MAX_WALLS][3];
  Plane2 hist[4];            // history-force state: sumDeltaFb xyz, n0 (allocated when SEDI_FORCE_HISTORY is configured)
  bool hist_alloc;
  Buf<double> UfOld;         // fluid velocity field of the previous coupling step (UfSmoothed_.oldTime())
  bool have_UfOld;
  int time_index;            // runTime().timeIndex() seen by fluid_force (set by the host once per fluid step)
  double inlet_force[3], inlet_box[9], inlet_ecc[3];
  int inlet_option;
  Buf<unsigned> wmask[2];
  Buf<int> foam[2];
  int icur;  // which of wmask/foam is live
  Ell ell[2];
  int ecur;
  // binning scratch
  Buf<int> cellid, cellcount, cellstart, cellfill, order, blocksum, tag2idx, rowstart;
  Buf<int> ctrl;                     // [0] rebuild flag, [1] steps executed, [2] errors, [3] max row length
  Buf<unsigned long long> counters;  // [0..3] list sizes written by the build, [4..5] optional k_step counters
  Pinned<int> h_ctrl;
  Pinned<unsigned long long> h_counters;
  BinParams bin;
  double cutneighmax;
  double eq_radius, eq_mass;
  int equal_spheres;   // every atom of the system has the same radius and mass (launch-uniform fast path of the pair law)
  double cutneighsq[(MAX_TYPES + 1) * (MAX_TYPES + 1)];
  double dt_init;
  bool setup_done_once;   // a full setup() has run at least once (init-time constants exist)
  double lub_R0, lub_RT0, lub_RS0;
  double beta_pair;
  StepParams base;
  // stats
  long long pair_evals_unique;
  long long nbuilds, pair_evals, steps_done, launches, list_gran_dir, list_type_dir, list_gran_img, list_type_img;
  int chunk;
  bool warned_neigh;
  int sm_count;
  MPI_Comm mpi_world;   // the communicator handed to lammps_open / new LAMMPS(0, NULL, comm) (library.h:29, softParticleCloud.C:60-62)
  bool comm_tried;
  bool use_wq;     // pair sweep on the warp-queue kernel (SEDI_KSTEP_PATH=wq)
  bool use_sell;   // pair sweep on the sorted-row kernel (default); SEDI_KSTEP_PATH=ell selects the streamed slot walk
  bool sell_sort;  // rows sorted by work inside windows at every rebuild (default with the sorted-row kernel; SEDI_SELL_SORT=0/1)
  Buf<int> order2, crow;
  double last_step_ms;
  bool count_in_kernel;
  // boundary staging
  Buf<double> d_stage_a, d_stage_b;
  Buf<int> d_stage_i, d_stage_j;
  Pinned<double> h_stage_a, h_stage_b;
  Pinned<int> h_stage_i, h_stage_j;
  // coupling
  MeshBox mesh;
  bool have_mesh;
  int ncells;
  Buf<int> cell;
  Buf<int> fc_count, fc_start, fc_fill, fc_rows, fc_blocksum;   // rows of every fluid cell, ascending (fixed-order scatter sums)
  Buf<double> Uf, gamma, gradp, DDtU, curlU, cellV, Ue, Asrc;
  bool have_DDtU, have_curlU, have_gradp, have_Uf;
  int drag_model, force_flags;
  double nub, rhob, gvec[3], deltaT;
  Buf<double> dg_Uri, dg_mag, dg_alpha, dg_Jd;
  bool want_diag;
  // diffusion smoothing (enhancedCloud.C:564-583, 790-907)
  double smooth_b, smooth_D[3];
  int smooth_steps, smooth_flags, smooth_iters_last;
  Buf<double> cg_r, cg_z, cg_p, cg_Ap, cg_partial, cg_s, cg_tmp;
  Pinned<double> h_cg;
  Comm comm;
  // the reference's timers (writeCPUTime.H:1-19) and conservation printouts, kept by the library so that a host's log stays comparable:
  // timers[0..1] diffusionTimeCount (field preparation / solves), [2] particleMoveTime (cell-owner location),
  // [3..8] cpuTimeSplit: assemble, transpose, flatten (no all-to-alls here: 0), foam->lammps, lammps, lammps->foam
  double timers[9];
  bool want_sums;
  double sums[12];   // Ftotal1, Ftotal2 (calcTcFields), Utotal1, Utotal2 (particleToEulerianField)
  Buf<double> sum_partial, sum_out;
  static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

  Engine()
      : device(0), dev_ready(false), loaded(false), setup_done(false), params_dirty(true), cell_valid(false), stream(0), n(0),
        nlocal(0), nghost(0), npad(0), maxtag(0), cur(0), icur(0), ecur(0), cutneighmax(0), eq_radius(0), eq_mass(0), equal_spheres(0), dt_init(0), setup_done_once(false), lub_R0(0), lub_RT0(0), lub_RS0(0),
        beta_pair(0), pair_evals_unique(0), nbuilds(0), pair_evals(0), steps_done(0), launches(0), list_gran_dir(0), list_type_dir(0), list_gran_img(0),
        list_type_img(0), chunk(50), last_step_ms(0), count_in_kernel(false), have_mesh(false), ncells(0), have_DDtU(false),
        have_curlU(false), have_gradp(false), have_Uf(false), drag_model(0), force_flags(SEDI_FORCE_DRAG | SEDI_FORCE_PGRAD), nub(1e-6), rhob(1000.0),
        deltaT(1.0), restart_every(0), restart_toggle(0), restart_pending(false), restart_carry(0), inject_pending(false), hist_alloc(false), have_UfOld(false), time_index(0), inlet_option(0), want_diag(false), smooth_b(0.0), smooth_steps(0), smooth_flags(0), smooth_iters_last(0), prof_on(false), prof_split(false), split_step_ms(0), split_halo_ms(0), split_n(0), prof_ms(0), prof_steps(0) {
    gvec[0] = gvec[1] = gvec[2] = 0.0;
    memset(timers, 0, sizeof(timers)); memset(sums, 0, sizeof(sums)); want_sums = false;
    memset(inlet_force, 0, sizeof(inlet_force)); memset(inlet_box, 0, sizeof(inlet_box)); memset(inlet_ecc, 0, sizeof(inlet_ecc));
    smooth_D[0] = smooth_D[1] = smooth_D[2] = 1.0;
    memset(&base, 0, sizeof(base));
    memset(&bin, 0, sizeof(bin));
    memset(&mesh, 0, sizeof(mesh));
    const char *e = getenv("SEDI_CHUNK");
    if (e && atoi(e) > 0) chunk = atoi(e);
    graph_on = true;
    e = getenv("SEDI_GRAPH");
    if (e && atoi(e) == 0) graph_on = false;
    e = getenv("SEDI_PROF_SPLIT");
    prof_split = (e && atoi(e) != 0);
    e = getenv("SEDI_KSTEP_PATH");
    warned_neigh = false; sm_count = 148; mpi_world = (MPI_Comm)0; comm_tried = false;
    use_wq = (e && !strcmp(e, "wq"));
    use_sell = !(e && (!strcmp(e, "ell") || !strcmp(e, "wq")));
    sell_sort = use_sell;
    if (const char *ss = getenv("SEDI_SELL_SORT")) sell_sort = atoi(ss) != 0;
    e = getenv("SEDI_DEVICE");
    if (e) device = atoi(e);
    else if ((e = getenv("LOCAL_RANK"))) device = atoi(e);
  }

  ~Engine() {
    if (dev_ready) { cudaSetDevice(device); destroy_graphs(); }
    for (size_t k = 0; k < script.cfg.dumps.size(); k++) if (script.cfg.dumps[k].fp) { fclose(script.cfg.dumps[k].fp); script.cfg.dumps[k].fp = 0; }
    if (!dev_ready) return;
    cudaSetDevice(device);
    if (prof_split && split_n) fprintf(stderr, "[sedi prof split] rank %d: %lld sub-steps, sub-step kernel %.2f us, ghost exchange / barrier %.2f us per sub-step\n", comm.rank, split_n, 1e3 * split_step_ms / split_n, 1e3 * split_halo_ms / split_n);
    for (size_t k = 0; k < split_ev.size(); k++) cudaEventDestroy(split_ev[k]);
    g_comm_engine_set(this);
    cudaStreamSynchronize(stream);
    // device memory is released with the process; explicit frees keep long-lived hosts clean
    for (int k = 0; k < 2; k++) { posr[k].release(); velm[k].release(); omgt[k].release(); wmask[k].release(); foam[k].release();
      ell[k].nbr.release(); ell[k].nn.release(); ell[k].nt.release(); ell[k].tmask.release(); ell[k].shear.release();
    }
    Plane2 *groups[] = {f, tq, fdrag, dudt, vold, uold, xhold};
    for (size_t g = 0; g < sizeof(groups) / sizeof(groups[0]); g++) for (int d = 0; d < 3; d++) { groups[g][d].b[0].release(); groups[g][d].b[1].release(); }
    for (int w = 0; w < MAX_WALLS; w++) for (int d = 0; d < 3; d++) { wshear[w][d].b[0].release(); wshear[w][d].b[1].release(); }
    cellid.release(); cellcount.release(); cellstart.release(); cellfill.release(); order.release(); blocksum.release();
    tag2idx.release(); rowstart.release(); ctrl.release(); counters.release(); h_ctrl.release(); h_counters.release();
    d_stage_a.release(); d_stage_b.release(); d_stage_i.release(); d_stage_j.release();
    h_stage_a.release(); h_stage_b.release(); h_stage_i.release(); h_stage_j.release();
    cell.release(); Uf.release(); gamma.release(); gradp.release(); DDtU.release(); curlU.release(); cellV.release(); Ue.release(); Asrc.release();
    dg_Uri.release(); dg_mag.release(); dg_alpha.release(); dg_Jd.release();
    cg_r.release(); cg_z.release(); cg_p.release(); cg_Ap.release(); cg_partial.release(); cg_s.release(); cg_tmp.release(); h_cg.release();
    comm.destroy();
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(evk0); cudaEventDestroy(evk1); cudaEventDestroy(ev_get[0]); cudaEventDestroy(ev_get[1]);
    cudaEventDestroy(evt0); cudaEventDestroy(evt1);
    cudaStreamDestroy(stream);
  }

  SimConfig &cfg() { return script.cfg; }

  // ---- device bring-up: no CPU fallback -------------------------------------------------------------------
  void need_device() {
    g_comm_engine_set(this);
    if (dev_ready) { CK(cudaSetDevice(device)); return; }
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt <= 0)
      fatal("libsedi_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback.",
            e != cudaSuccess ? cudaGetErrorString(e) : "");
    if (device >= cnt) device = device % cnt;
    CK(cudaSetDevice(device));
    { cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, device)); sm_count = pr.multiProcessorCount; }
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1)); CK(cudaEventCreate(&evk0)); CK(cudaEventCreate(&evk1));
    CK(cudaEventCreateWithFlags(&ev_get[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev_get[1], cudaEventDisableTiming));
    CK(cudaEventCreate(&evt0)); CK(cudaEventCreate(&evt1));
    ctrl.ensure(8); counters.ensure(8); h_ctrl.ensure(8); h_counters.ensure(8);
    CK(cudaMemsetAsync(ctrl.p, 0, 8 * sizeof(int), stream));
    CK(cudaMemsetAsync(counters.p, 0, 8 * sizeof(unsigned long long), stream));
    dev_ready = true;
  }

  // ---- script ------------------------------------------------------------------------------------------------
  void command(const char *line) {
    ScriptAction a = script.one(line);
    params_dirty = true;
    if (a.kind == ScriptAction::READ_DATA) loaded = false;
    if (a.kind == ScriptAction::RUN) run(a.nsteps);
    if (a.kind == ScriptAction::WRITE_RESTART) write_restart(a.path);
    if (a.kind == ScriptAction::READ_RESTART) read_restart(a.path);
    if (a.kind == ScriptAction::RESTART_EVERY) { restart_every = a.nsteps; restart_path[0] = a.path; restart_path[1] = a.path2; restart_toggle = 0; }
  }

  void file(const char *path) {
    FILE *fp = fopen(path, "r");
    if (!fp) fatal("Cannot open input script", path);
    char buf[2048];
    while (fgets(buf, sizeof(buf), fp)) command(buf);
    fclose(fp);
  }

  // ---- upload of the script's atoms (EXTERNAL read_data / AtomVecSphere::data_atom) -----------------------------
  void alloc_rows(int rows) {
    npad = ((rows + 127) / 128) * 128;
    if (npad < 128) npad = 128;
    leave.ensure(npad);
    for (int k = 0; k < 2; k++) { posr[k].ensure(npad); velm[k].ensure(npad); omgt[k].ensure(npad); wmask[k].ensure(npad); foam[k].ensure(npad); }
    Plane2 *groups[] = {f, tq, fdrag, dudt, vold, uold, xhold};
    for (size_t g = 0; g < sizeof(groups) / sizeof(groups[0]); g++)
      for (int d = 0; d < 3; d++) { groups[g][d].b[0].ensure(npad); groups[g][d].b[1].ensure(npad); }
    for (int w = 0; w < cfg().nwalls && w < MAX_WALLS; w++)
      for (int d = 0; d < 3; d++) { wshear[w][d].b[0].ensure(npad); wshear[w][d].b[1].ensure(npad); }
    cellid.ensure(npad); order.ensure(npad); cell.ensure(npad); rowstart.ensure(npad + 1);
  }

  void zero_plane(Plane2 &p) {
    CK(cudaMemsetAsync(p.b[0].p, 0, p.b[0].cap * sizeof(double), stream));
    CK(cudaMemsetAsync(p.b[1].p, 0, p.b[1].cap * sizeof(double), stream));
  }

  void load_atoms() {
    need_device();
    const AtomData &a = script.atoms;
    if (cfg().nwalls > MAX_WALLS) fatal("Too many wall fixes");
    if (cfg().ntypes > MAX_TYPES) fatal("Too many atom types for the device cut-off table");
    if ((int)cfg().fixes.size() > MAX_FIXES) fatal("Too many fixes");
    // every rank reads the whole atom table (as LAMMPS' read_data does) and keeps the atoms of its own brick
    std::vector<D4> hp, hv, hw;
    maxtag = 0;
    loaded_index.clear();
    const bool have_keep = keep_local.size() == a.size();
    for (size_t i = 0; i < a.size(); i++) {
      if (a.tag[i] < 0) fatal("Negative atom tag");
      maxtag = std::max(maxtag, a.tag[i]);
      if (comm.nranks > 1 && !(have_keep && keep_local[i])) {
        double xw[3] = {a.x[3 * i], a.x[3 * i + 1], a.x[3 * i + 2]};
        for (int d = 0; d < 3; d++) if (cfg().periodic[d]) {
          const double lo = cfg().boxlo[d], hi = cfg().boxhi[d], prd = hi - lo;
          if (xw[d] < lo) xw[d] += prd;
          if (xw[d] >= hi) { xw[d] -= prd; xw[d] = std::max(xw[d], lo); }
        }
        if (!comm.owns(cfg(), xw)) continue;
      }
      D4 p, v, w;
      p.x = a.x[3 * i]; p.y = a.x[3 * i + 1]; p.z = a.x[3 * i + 2]; p.w = a.radius[i];
      v.x = a.v[3 * i]; v.y = a.v[3 * i + 1]; v.z = a.v[3 * i + 2]; v.w = a.rmass[i];
      w.x = a.omega[3 * i]; w.y = a.omega[3 * i + 1]; w.z = a.omega[3 * i + 2];
      if (script.mask[i] > 0xFFFF) fatal("Too many groups for the 16-bit device mask");
      const unsigned long long b = pack_bits(a.tag[i], script.mask[i], a.type[i], 0);
      long long sb = (long long)b;
      memcpy(&w.w, &sb, 8);
      hp.push_back(p); hv.push_back(v); hw.push_back(w);
      loaded_index.push_back((int)i);
    }
    keep_local.clear();
    {
      double mt = (double)maxtag;
      comm.allreduce_max_host(&mt, 1);
      maxtag = (int)mt;
    }
    nlocal = (int)hp.size();
    nghost = 0;
    n = nlocal;
    if (n > (int)NB_IDX_MASK / 2) fatal("Too many particles per GPU for the 25-bit neighbour index");
    // capacity: owned rows + (multi-GPU) ghost shell and migration slack
    int rows = n;
    if (comm.nranks > 1) {
      const char *e = getenv("SEDI_ROW_SLACK");
      const double slack = e ? atof(e) : 0.6;
      rows = (int)(n * (1.0 + slack)) + 65536;
    }
    alloc_rows(rows);
    cur = 0; icur = 0; ecur = 0;
    drop_graphs();
    hist_alloc = false;   // history-force state restarts with the atom table (softParticle.C:63-64: n0 = 0, sumDeltaFb = 0)
    ell[0].valid = ell[1].valid = false;
    if (n) {
      CK(cudaMemcpyAsync(posr[0].p, hp.data(), n * sizeof(D4), cudaMemcpyHostToDevice, stream));
      CK(cudaMemcpyAsync(velm[0].p, hv.data(), n * sizeof(D4), cudaMemcpyHostToDevice, stream));
      CK(cudaMemcpyAsync(omgt[0].p, hw.data(), n * sizeof(D4), cudaMemcpyHostToDevice, stream));
    }
    Plane2 *groups[] = {f, tq, fdrag, dudt, vold, uold, xhold};
    for (size_t g = 0; g < sizeof(groups) / sizeof(groups[0]); g++) for (int d = 0; d < 3; d++) { groups[g][d].cur = 0; zero_plane(groups[g][d]); }
    for (int w = 0; w < cfg().nwalls; w++) for (int d = 0; d < 3; d++) { wshear[w][d].cur = 0; zero_plane(wshear[w][d]); }
    for (int k = 0; k < 2; k++) {
      CK(cudaMemsetAsync(wmask[k].p, 0, wmask[k].cap * sizeof(unsigned), stream));
      CK(cudaMemsetAsync(foam[k].p, 0, foam[k].cap * sizeof(int), stream));
    }
    tag2idx.ensure((size_t)maxtag + 2);
    CK(cudaMemsetAsync(tag2idx.p, 0xFF, tag2idx.cap * sizeof(int), stream));
    CK(cudaMemsetAsync(leave.p, 0, leave.cap * sizeof(int), stream));
    CK(cudaStreamSynchronize(stream));  // host vectors go out of scope
    loaded = true; setup_done = false; cell_valid = false;
  }

  // ---- cut-offs: EXTERNAL PairGranHookeHistory::init_one / PairLubricate::init_one / Neighbor::init -------------
  // (same rules as the oracle driver, evaluated on the host from the script's atoms)
  void compute_cutoffs() {
    const int nt = cfg().ntypes;
    const AtomData &a = script.atoms;
    std::vector<double> maxdyn(nt + 1, 0.0), maxfrz(nt + 1, 0.0);
    for (size_t i = 0; i < a.size(); i++) {
      const int t = a.type[i];
      if (t < 1 || t > nt) fatal("Atom type out of range");
      if (script.mask[i] & cfg().freeze_group_bit) maxfrz[t] = std::max(maxfrz[t], a.radius[i]);
      else maxdyn[t] = std::max(maxdyn[t], a.radius[i]);
    }
    comm.allreduce_max_host(maxdyn.data(), nt + 1);
    comm.allreduce_max_host(maxfrz.data(), nt + 1);
    double cutmax = 0.0;
    memset(cutneighsq, 0, sizeof(cutneighsq));
    for (int ta = 1; ta <= nt; ta++)
      for (int tb = 1; tb <= nt; tb++) {
        double c = 0.0;
        if (cfg().pair != PAIR_NONE) {
          c = maxdyn[ta] + maxdyn[tb];
          c = std::max(c, maxfrz[ta] + maxdyn[tb]);
          c = std::max(c, maxdyn[ta] + maxfrz[tb]);
        }
        if (cfg().lub.enabled) c = std::max(c, cfg().lub.cut_global);
        const double cn = c + cfg().skin;
        cutneighsq[ta * (MAX_TYPES + 1) + tb] = cn * cn;
        cutmax = std::max(cutmax, c);
      }
    cutneighmax = cutmax + cfg().skin;
    if (!(cutneighmax > 0.0)) cutneighmax = std::max(cfg().skin, 1e-30);
    // monodisperse system?  (same decision on every rank: the pair law's equal-sphere branch must be taken by both owners of a pair)
    {
      const double inf = std::numeric_limits<double>::infinity();
      double mm[4] = {-inf, -inf, -inf, -inf};   // max r, max -r, max m, max -m
      for (size_t i = 0; i < a.size(); i++) {
        mm[0] = std::max(mm[0], a.radius[i]); mm[1] = std::max(mm[1], -a.radius[i]);
        mm[2] = std::max(mm[2], a.rmass[i]); mm[3] = std::max(mm[3], -a.rmass[i]);
      }
      comm.allreduce_max_host(mm, 4);
      equal_spheres = (mm[0] == -mm[1] && mm[2] == -mm[3] && mm[0] > 0.0 && mm[2] > 0.0) ? 1 : 0;
      if (getenv("SEDI_NO_EQUAL_SPHERES")) equal_spheres = 0;
      eq_radius = mm[0]; eq_mass = mm[2];
    }
  }

  bool want_type_list() const {
    for (size_t k = 0; k < script.cfg.fixes.size(); k++) if (script.cfg.fixes[k].kind == FIX_COHESIVE) return true;
    return script.cfg.lub.enabled != 0;
  }

  // loop-invariant of the "Fix" damping model, evaluated with the reference's own expression
  // (pair_gran_hertzFix_history.cpp:195-196, fix_wall_granFix.cpp:602-603)
  static double fix_beta(double gamman) {
    const double lg = log(gamman) / log(exp(1.0));
    return -lg / sqrt(lg * lg + SEDI_MY_PI * SEDI_MY_PI);
  }

  void setup_bins() {
    // one bin >= the largest neighbour cut-off, 27-cell stencil; periodic dimensions are tiled exactly
    long long total = 1;
    double grow = 1.0;
    bin.tile[0] = bin.tile[1] = 1;
    if (const char *e = getenv("SEDI_BIN_TILE")) {   // "TXxTY": tile-major row order (see cell_index)
      int tx = 1, ty = 1;
      if (sscanf(e, "%dx%d", &tx, &ty) == 2 && tx >= 1 && ty >= 1 && tx <= 64 && ty <= 64) { bin.tile[0] = tx; bin.tile[1] = ty; }
    }
    for (int pass = 0; pass < 40; pass++) {
      total = 1;
      for (int d = 0; d < 3; d++) {
        const double lo = comm.sublo(cfg(), d, cutneighmax), hi = comm.subhi(cfg(), d, cutneighmax);
        const double len = hi - lo;
        int nb = (int)floor(len / (cutneighmax * grow));
        if (nb < 1) nb = 1;
        bin.nb[d] = nb; bin.lo[d] = lo; bin.inv[d] = nb / len;
        bin.periodic[d] = comm.wraps(cfg(), d) ? 1 : 0;
        total *= nb;
      }
      if (total <= (1ll << 26)) break;
      grow *= 1.3;
    }
    total = cell_count(bin.nb, bin.tile);
    bin.n = n;
    cellcount.ensure((size_t)total + 2); cellstart.ensure((size_t)total + 2); cellfill.ensure((size_t)total + 2);
    blocksum.ensure((size_t)cdiv(total + 2, SCAN_ITEMS) + 1);
    if (comm.nranks > 1) comm.setup_decomp(*this);
  }

  long long ncells_bin() const { return cell_count(bin.nb, bin.tile); }

  void build_base_params() {
    StepParams &P = base;
    memset(&P, 0, sizeof(P));
    const SimConfig &c = cfg();
    P.n = n; P.npad = ell[ecur].npad; P.nfix = (int)c.fixes.size(); P.pair = c.pair;
    P.lub_enabled = c.lub.enabled; P.lub_flaglog = c.lub.flaglog; P.lub_flagfld = c.lub.flagfld; P.lub_flagHI = c.lub.flagHI;
    P.freeze_groupbit = c.freeze_group_bit;
    P.equal_spheres = equal_spheres;
    P.periodic_any = c.periodic[0] | c.periodic[1] | c.periodic[2];
    P.dtv = dt_init; P.dtf = 0.5 * dt_init; P.dt_live = c.dt;
    P.trigger_sq = 0.25 * c.skin * c.skin;
    if (equal_spheres) {   // nve/sphere factors of the one particle class, with the reference's own expressions (EXTERNAL FixNVESphere)
      P.c_dtfm = P.dtf / eq_mass;
      P.c_dtirot = (P.dtf / 0.4) / (eq_radius * eq_radius * eq_mass);
    }
    P.kn = c.gran.kn; P.kt = c.gran.kt; P.gamman = c.gran.gamman; P.gammat = c.gran.gammat; P.xmu = c.gran.xmu;
    P.beta = (c.pair == PAIR_HERTZFIX_HISTORY) ? fix_beta(c.gran.gamman) : 0.0;
    {  // loop invariants of the Hertz-Mindlin "Fix" law, folded once (see hertzfix_fast)
      const double s56 = sqrt(5.0 / 6.0);
      P.c_sn = 2.0 * 1.0 / 1.82 * P.kn;
      P.c_ccel = 4.0 / 5.46 * P.kn;
      P.c_damp = 2.0 * s56 * P.beta;
      P.c_kts = 8.0 / 8.84 * P.kt;
      P.c_ctd = sqrt((8.0 / 8.84) / (2.0 / 1.82)) * (2.0 * s56 * P.beta);
      P.c_ekt = (P.kt != 0.0) ? 8.0 / (8.84 * P.kt) : 0.0;
    }
    for (int d = 0; d < 3; d++) P.prd[d] = c.boxhi[d] - c.boxlo[d];
    for (int img = 0; img < 27; img++) {
      const int ix = img % 3 - 1, iy = (img / 3) % 3 - 1, iz = img / 9 - 1;
      P.imgshift[img][0] = ix * P.prd[0]; P.imgshift[img][1] = iy * P.prd[1]; P.imgshift[img][2] = iz * P.prd[2];
    }
    P.lub_mu = c.lub.mu; P.lub_cutsq = c.lub.cut_global * c.lub.cut_global; P.lub_cut_inner = c.lub.cut_inner;
    P.lub_R0 = lub_R0; P.lub_RT0 = lub_RT0;
    P.ctrl = ctrl.p;
    P.counters = count_in_kernel ? counters.p + 4 : 0;
    for (size_t k = 0; k < c.fixes.size(); k++) {
      const FixSpec &s = c.fixes[k];
      FixDev &F = P.fix[k];
      F.kind = s.kind; F.groupbit = s.groupbit; F.time_origin = s.time_origin; F.wall_index = s.wall_index;
      switch (s.kind) {
        case FIX_NVE_SPHERE: P.nve_groupbit |= s.groupbit; break;
        case FIX_GRAVITY: for (int d = 0; d < 3; d++) F.d[d] = s.g * s.gdir[d]; break;
        case FIX_FDRAG: F.d[0] = s.carrier_rho; P.has_fdrag = 1; if (s.carrier_rho != 0.0) P.fdrag_added_mass = 1; break;
        case FIX_COHESIVE: F.d[0] = s.ah; F.d[1] = s.lam; F.d[2] = s.smin; F.d[3] = s.smax; F.i0 = s.opt; P.has_cohesive = 1; break;
        case FIX_WALL_GRAN:
          F.i0 = s.wallstyle; F.i1 = s.wiggle; F.i2 = s.wshear; F.i3 = s.axis;
          if (s.wallstyle > ZPLANE) P.has_cyl_wall = 1;
          F.d[0] = s.wall.kn; F.d[1] = s.wall.kt; F.d[2] = s.wall.gamman; F.d[3] = s.wall.gammat; F.d[4] = s.wall.xmu;
          F.d[5] = s.lo; F.d[6] = s.hi; F.d[7] = s.cylradius;
          F.d[8] = (c.pair == PAIR_HERTZFIX_HISTORY) ? fix_beta(s.wall.gamman) : 0.0;
          F.d[9] = s.wshear ? s.vshear : 0.0; F.aux = s.vshear;
          break;
        default: break;
      }
    }
    params_dirty = false;
    drop_graphs();
  }

  // several GPUs, peer-memory halo: the sub-step kernel writes the neighbours' ghost rows itself
  bool step_pushes() { return comm.nranks > 1 && comm.p2p && comm.fused_push; }

  // per-launch part of the parameter block
  void fill_launch(StepParams &P, int mode, int in, long long ntimestep) {
    P = base;
    P.mode = mode; P.ntimestep = ntimestep;
    P.n = nlocal;
    Ell &L = ell[ecur];
    P.npad = L.npad; P.nn = L.nn.p; P.nt = L.nt.p; P.hcap = L.cap; P.nbr = L.nbr.p; P.shear = L.shear.p; P.tmask = L.tmask.p;
    P.posr_in = posr[in].p; P.velm_in = velm[in].p; P.omgt_in = omgt[in].p;
    P.posr_out = posr[in ^ 1].p; P.velm_out = velm[in ^ 1].p; P.omgt_out = omgt[in ^ 1].p;
    for (int d = 0; d < 3; d++) {
      P.f[d] = f[d].get(); P.tq[d] = tq[d].get(); P.fdrag[d] = fdrag[d].get(); P.dudt[d] = dudt[d].get();
      P.vold[d] = vold[d].get(); P.xhold[d] = xhold[d].get();
      for (int w = 0; w < cfg().nwalls; w++) P.wshear[w][d] = wshear[w][d].get();
    }
    P.wmask = wmask[icur].p;
    if (comm.nranks > 1 && comm.p2p && comm.fused_push && mode == MODE_FUSED) {   // the kernel refreshes the neighbours' ghost rows itself
      P.bcnt = comm.d_bcnt; P.bpos = comm.d_bpos; P.bent = comm.d_bent; P.push = comm.d_push[in ^ 1];
    }
    const SimConfig &c = cfg();
    for (size_t k = 0; k < c.fixes.size(); k++) {
      const FixSpec &s = c.fixes[k];
      if (s.kind != FIX_WALL_GRAN || !s.wiggle) continue;
      FixDev &F = P.fix[k];  // fix_wall_granFix.cpp:254-262
      const double omega = 2.0 * SEDI_MY_PI / s.period;
      const double arg = omega * (ntimestep - s.time_origin) * dt_init;
      if (s.wallstyle == s.axis) {
        F.d[5] = s.lo + s.amplitude - s.amplitude * cos(arg);
        F.d[6] = s.hi + s.amplitude - s.amplitude * cos(arg);
      }
      F.d[9] = s.amplitude * omega * sin(arg);
    }
  }

  void launch_step(int mode, int in, long long ntimestep, int seq) {
    StepParams P;
    fill_launch(P, mode, in, ntimestep);
    const int KT = SEDI_KSTEP_THREADS;
    const int blocks = std::max(1, cdiv(nlocal, KT));   // an empty brick still counts its sub-steps (ctrl[1])
    const bool tl = P.has_cohesive || P.lub_enabled;
    const bool pbc = P.periodic_any != 0;
    if (use_sell && cfg().pair != PAIR_NONE) {   // sorted-row kernel: one lane per particle, rows of a warp carry equal work (sedi_sell.cuh)
      const int ST = SEDI_SELL_THREADS;
      const int sb = std::max(1, cdiv(nlocal, ST));
      // 32-bit row masks / all list words staged in shared memory when no row has more than 16 granular slots (SEDI_SELL_M32=0 forces the general form)
      static const bool allow32 = !(getenv("SEDI_SELL_M32") && atoi(getenv("SEDI_SELL_M32")) == 0);
      const bool m32 = allow32 && P.hcap <= 16;
#define SEDI_LAUNCH_SELL3(PK, PB, TL)                                                                         \
  do { if (m32) k_step_sell<PK, PB, TL, true><<<sb, ST, 0, stream>>>(P, seq); else k_step_sell<PK, PB, TL, false><<<sb, ST, 0, stream>>>(P, seq); } while (0)
#define SEDI_LAUNCH_SELL(PK)                                                                                  \
  do {                                                                                                        \
    if (P.lub_enabled) { if (pbc) SEDI_LAUNCH_SELL3(PK, true, 2); else SEDI_LAUNCH_SELL3(PK, false, 2); }    \
    else if (tl) { if (pbc) SEDI_LAUNCH_SELL3(PK, true, 1); else SEDI_LAUNCH_SELL3(PK, false, 1); }          \
    else { if (pbc) SEDI_LAUNCH_SELL3(PK, true, 0); else SEDI_LAUNCH_SELL3(PK, false, 0); }                  \
  } while (0)
      switch (cfg().pair) {
        case PAIR_HERTZFIX_HISTORY: SEDI_LAUNCH_SELL(PAIR_HERTZFIX_HISTORY); break;
        case PAIR_HOOKE_HISTORY: SEDI_LAUNCH_SELL(PAIR_HOOKE_HISTORY); break;
        default: SEDI_LAUNCH_SELL(PAIR_HOOKE); break;
      }
#undef SEDI_LAUNCH_SELL
#undef SEDI_LAUNCH_SELL3
      launches++;
      return;
    }
    if (use_wq && !tl && cfg().pair != PAIR_NONE) {   // warp-queue kernel: one overlapping contact per lane (sedi_wq.cuh)
      const int WT = SEDI_WQ_THREADS;
      const int wb = std::max(1, cdiv(nlocal, WT));
#define SEDI_LAUNCH_WQ(PK) do { if (pbc) k_step_wq<PK, true><<<wb, WT, 0, stream>>>(P, seq); else k_step_wq<PK, false><<<wb, WT, 0, stream>>>(P, seq); } while (0)
      switch (cfg().pair) {
        case PAIR_HERTZFIX_HISTORY: SEDI_LAUNCH_WQ(PAIR_HERTZFIX_HISTORY); break;
        case PAIR_HOOKE_HISTORY: SEDI_LAUNCH_WQ(PAIR_HOOKE_HISTORY); break;
        default: SEDI_LAUNCH_WQ(PAIR_HOOKE); break;
      }
#undef SEDI_LAUNCH_WQ
      launches++;
      return;
    }
#define SEDI_LAUNCH_KSTEP(PK)                                                                      \
  do {                                                                                             \
    if (tl) { if (pbc) k_step<PK, true, true><<<blocks, KT, 0, stream>>>(P, seq); else k_step<PK, true, false><<<blocks, KT, 0, stream>>>(P, seq); } \
    else { if (pbc) k_step<PK, false, true><<<blocks, KT, 0, stream>>>(P, seq); else k_step<PK, false, false><<<blocks, KT, 0, stream>>>(P, seq); } \
  } while (0)
    switch (cfg().pair) {
      case PAIR_HERTZFIX_HISTORY: SEDI_LAUNCH_KSTEP(PAIR_HERTZFIX_HISTORY); break;
      case PAIR_HOOKE_HISTORY: SEDI_LAUNCH_KSTEP(PAIR_HOOKE_HISTORY); break;
      case PAIR_HOOKE: SEDI_LAUNCH_KSTEP(PAIR_HOOKE); break;
      default: SEDI_LAUNCH_KSTEP(PAIR_NONE); break;
    }
#undef SEDI_LAUNCH_KSTEP
    launches++;
  }

  void launch_initial(int in, int seq) {
    StepParams P;
    fill_launch(P, MODE_FUSED, in, cfg().ntimestep);
    const int blocks = cdiv(nlocal, 256);
    if (!blocks) return;
    k_initial_integrate<<<blocks, 256, 0, stream>>>(P, seq);
    launches++;
  }

  // ---- neighbour rebuild (EXTERNAL Verlet: pre_exchange history save, pbc, exchange, borders, Neighbor::build) ---
  void rebuild(bool count = true) {
    need_device();
    drop_graphs();
    const SimConfig &c = cfg();
    const int T = 256;
    const int n_old = nlocal + nghost;  // rows that are valid in quads[cur] (ghost rows are dropped by the sort)
    int narr = 0;
    // leavers flagged in leave[], arrivals appended at rows >= n_old.  Not at the re-upload after an injection / deletion / restart:
    // its rows carry their contact history as arrival lists by upload order; a row that has strayed out of its brick (by less than
    // skin / 2) stays with its owner until the next regular rebuild, as between any two re-neighbourings
    if (comm.nranks > 1 && !inject_pending) narr = comm.migrate(*this);
    const int n_tmp = n_old + narr;
    const long long nc = ncells_bin();
    const int trash = (int)nc;                         // cell id of rows that leave the sort (ghosts, migrated-away)
    CK(cudaMemsetAsync(cellcount.p, 0, (size_t)(nc + 2) * sizeof(int), stream));
    CK(cudaMemsetAsync(cellfill.p, 0, (size_t)(nc + 2) * sizeof(int), stream));
    if (n_tmp) {
      k_wrap_bin<<<cdiv(n_tmp, T), T, 0, stream>>>(posr[cur].p, omgt[cur].p, n_tmp, bin, c.boxhi[0], c.boxhi[1], c.boxhi[2], c.boxhi[0] - c.boxlo[0],
                                                   c.boxhi[1] - c.boxlo[1], c.boxhi[2] - c.boxlo[2], cellid.p, cellcount.p,
                                                   c.periodic[0], c.periodic[1], c.periodic[2], c.boxlo[0], c.boxlo[1], c.boxlo[2],
                                                   comm.nranks > 1 ? leave.p : (const int *)0, trash, comm.nranks > 1 ? 0 : 1, ctrl.p + 2);
      CK(cudaMemcpyAsync(h_ctrl.p + 2, ctrl.p + 2, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      if (h_ctrl.p[2] & 4) fatal("Out of range atoms: a particle moved more than a box length (or became NaN) since the last neighbour rebuild -- the run is unstable");
    }
    const int nscan = (int)(nc + 2);
    const int nblk = cdiv(nscan, SCAN_ITEMS);
    k_scan_local<<<nblk, 1024, 0, stream>>>(cellcount.p, cellstart.p, nscan, blocksum.p);
    k_scan_sums<<<1, 1024, 0, stream>>>(blocksum.p, nblk);
    k_scan_add<<<cdiv(nscan, T), T, 0, stream>>>(cellstart.p, nscan, blocksum.p, 0);
    launches += 4;
    int nlocal_new = n_tmp;
    if (comm.nranks > 1) {  // kept rows = start of the trash cell
      CK(cudaMemcpyAsync(h_ctrl.p + 4, cellstart.p + nc, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      nlocal_new = h_ctrl.p[4];
    }
    if (n_tmp) {
      k_bin_scatter<<<cdiv(n_tmp, T), T, 0, stream>>>(cellid.p, n_tmp, cellstart.p, cellfill.p, order.p);
      k_cell_sort<<<cdiv(nc, T), T, 0, stream>>>(cellstart.p, (int)nc, nlocal_new, order.p, omgt[cur].p);
      launches += 2;
    }
    // rows sorted by work inside windows (SELL-C-sigma, sedi_sell.cuh): ord = new row -> old row, crow = bin position -> new row
    const int *ord = order.p;
    const bool sorted_rows = sell_sort && nlocal_new > 0;
    if (sorted_rows) {
      order2.ensure(npad); crow.ensure(npad);
      Ell &Lprev = ell[ecur];
      BorderBand band;
      for (int d = 0; d < 3; d++) {
        const double inf = std::numeric_limits<double>::infinity();
        const bool split = comm.nranks > 1 && comm.grid[d] > 1;
        band.lo[d] = split ? comm.sublo(cfg(), d, 0.0) + cutneighmax : -inf;
        band.hi[d] = split ? comm.subhi(cfg(), d, 0.0) - cutneighmax : inf;
      }
      k_window_sort<<<cdiv(nlocal_new, SELL_WINDOW), SELL_WINDOW, 0, stream>>>(nlocal_new, order.p, Lprev.valid ? Lprev.tmask.p : (const unsigned long long *)0,
                                                                               Lprev.valid ? Lprev.nn.p : (const int *)0,
                                                                               (Lprev.valid && want_type_list()) ? Lprev.nt.p : (const int *)0, nlocal, order2.p, crow.p,
                                                                               (step_pushes() && !getenv("SEDI_BORDER_CLUSTER_OFF")) ? posr[cur].p : (const D4 *)0, band);
      launches++;
      ord = order2.p;
    }
    if (nlocal_new) {
      // physical re-ordering into cell order: quads cur -> cur^1, planes get() -> alt()
      k_permute_quads<<<cdiv(nlocal_new, T), T, 0, stream>>>(ord, nlocal_new, posr[cur].p, velm[cur].p, omgt[cur].p, posr[cur ^ 1].p,
                                                             velm[cur ^ 1].p, omgt[cur ^ 1].p, xhold[0].get(), xhold[1].get(), xhold[2].get(),
                                                             tag2idx.p, maxtag, wmask[icur].p, wmask[icur ^ 1].p, foam[icur].p, foam[icur ^ 1].p);
      PlaneList L;
      L.nplanes = 0;
      Plane2 *groups[] = {f, tq, fdrag, dudt, vold, uold};
      for (int g = 0; g < NPLANES_BASE; g++) for (int d = 0; d < 3; d++) { L.src[L.nplanes] = groups[g][d].get(); L.dst[L.nplanes] = groups[g][d].alt(); L.nplanes++; }
      for (int w = 0; w < c.nwalls; w++) for (int d = 0; d < 3; d++) { L.src[L.nplanes] = wshear[w][d].get(); L.dst[L.nplanes] = wshear[w][d].alt(); L.nplanes++; }
      if (hist_alloc) for (int d = 0; d < 4; d++) { L.src[L.nplanes] = hist[d].get(); L.dst[L.nplanes] = hist[d].alt(); L.nplanes++; }
      k_permute_planes<<<cdiv(nlocal_new, T), T, 0, stream>>>(ord, nlocal_new, L);
      launches += 2;
    }
    {
      Plane2 *groups[] = {f, tq, fdrag, dudt, vold, uold};
      for (int g = 0; g < NPLANES_BASE; g++) for (int d = 0; d < 3; d++) groups[g][d].cur ^= 1;
      for (int w = 0; w < c.nwalls; w++) for (int d = 0; d < 3; d++) wshear[w][d].cur ^= 1;
      if (hist_alloc) for (int d = 0; d < 4; d++) hist[d].cur ^= 1;
    }
    const int oldq = cur;
    cur ^= 1; icur ^= 1;
    nlocal = nlocal_new; nghost = 0; n = nlocal;
    if (comm.nranks > 1) {
      CK(cudaMemsetAsync(leave.p, 0, (size_t)npad * sizeof(int), stream));
      comm.borders(*this);  // ghost rows [nlocal, nlocal + nghost) of the new buffers, binned by cell
      n = nlocal + nghost;
    }
    bin.n = n;
    // directed ELL list + history re-attachment
    Ell &Lo = ell[ecur], &Ln = ell[ecur ^ 1];
    const int npad_ell = std::max(128, ((nlocal + 127) / 128) * 128);
    BuildParams B;
    memset(&B, 0, sizeof(B));
    B.n = nlocal; B.want_gran = (c.pair != PAIR_NONE); B.want_type = want_type_list() ? 1 : 0; B.ntypes = c.ntypes;
    B.posr = posr[cur].p; B.omgt = omgt[cur].p; B.cellstart = cellstart.p;
    for (int d = 0; d < 3; d++) { B.nb[d] = bin.nb[d]; if (d < 2) B.tile[d] = bin.tile[d]; B.periodic[d] = bin.periodic[d]; B.lo[d] = bin.lo[d]; B.inv[d] = bin.inv[d]; B.prd[d] = c.boxhi[d] - c.boxlo[d]; }
    B.skin = c.skin;
    memcpy(B.cutneighsq, cutneighsq, sizeof(cutneighsq));
    B.have_old = (Lo.valid || narr > 0) ? 1 : 0; B.npad_old = Lo.npad; B.oldidx = ord;
    B.crow = sorted_rows ? crow.p : (const int *)0;
    B.nbr_old = Lo.nbr.p; B.nn_old = Lo.valid ? Lo.nn.p : (const int *)0; B.tmask_old = Lo.tmask.p; B.shear_old = Lo.shear.p; B.omgt_old = omgt[oldq].p;
    B.n_old = Lo.valid ? n_old : 0;
    B.nlocal_rows = nlocal;
    if (comm.nranks > 1) {
      B.gcellstart = comm.d_gstart; B.gorder = comm.d_gorder;
      if (narr > 0) { B.arr_nh = comm.d_arr_nh; B.arr_tag = comm.d_arr_tag; B.arr_shear = comm.d_arr_shear; if (!Lo.valid) B.n_old = n_old; }
    }
    if (inject_pending) {  // every row is an "arrival" that brings its contact history as (partner tag, shear) pairs
      B.have_old = 1; B.n_old = 0; B.nn_old = 0;
      B.arr_nh = inj_nh.p; B.arr_tag = inj_tag.p; B.arr_shear = inj_shear.p;
    }
    B.maxcount = ctrl.p + 3; B.maxcount_t = ctrl.p + 6; B.npairs = counters.p;
    if (Ln.cap < 16) Ln.cap = std::max(Lo.cap, 16);   // k_step_wq reads the first sixteen list words of every row unconditionally
    if (Ln.tcap < Lo.tcap) Ln.tcap = Lo.tcap;
    for (int attempt = 0; attempt < 3; attempt++) {
      Ln.npad = npad_ell;
      Ln.nn.ensure((size_t)npad_ell + 1); Ln.nt.ensure((size_t)npad_ell + 1); Ln.tmask.ensure(npad_ell);
      Ln.nbr.ensure((size_t)(Ln.cap + Ln.tcap) * npad_ell); Ln.shear.ensure((size_t)Ln.cap * npad_ell);
      B.npad = Ln.npad; B.cap = Ln.cap; B.tcap = Ln.tcap; B.nbr = Ln.nbr.p; B.nn = Ln.nn.p; B.nt = Ln.nt.p; B.tmask = Ln.tmask.p; B.shear = Ln.shear.p;
      CK(cudaMemsetAsync(ctrl.p + 3, 0, sizeof(int), stream));
      CK(cudaMemsetAsync(ctrl.p + 6, 0, sizeof(int), stream));
      CK(cudaMemsetAsync(counters.p, 0, 4 * sizeof(unsigned long long), stream));
      if (nlocal) k_build_list<<<cdiv(nlocal, 128), 128, 0, stream>>>(B);
      launches++;
      CK(cudaMemcpyAsync(h_ctrl.p, ctrl.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(h_ctrl.p + 6, ctrl.p + 6, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(h_counters.p, counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      CK(cudaGetLastError());
      int maxrow = h_ctrl.p[3], maxtype = h_ctrl.p[6];
      if (comm.nranks > 1) { double m[2] = {(double)maxrow, (double)maxtype}; comm.allreduce_max_host(m, 2); maxrow = (int)m[0]; maxtype = (int)m[1]; }  // same capacity decision on every rank
      if (maxrow <= Ln.cap && maxtype <= Ln.tcap) break;
      // the 64-bit contact-history mask limits the GRANULAR entries of a row (partners within ri + rj + skin); entries that are
      // only in the type-cut-off list (fix cohesive, lubricate/poly) have their own segment without such a limit
      if (maxrow > MAX_SLOTS) fatal("More than 64 granular neighbours (within ri + rj + skin) in one row: reduce the skin (the contact-history mask is 64 bits)");
      if (maxrow > Ln.cap) Ln.cap = std::max(16, std::min(MAX_SLOTS, ((maxrow + 2 + 3) / 4) * 4));
      if (maxtype > Ln.tcap) Ln.tcap = ((maxtype + 4 + 3) / 4) * 4;
      if (attempt == 2) fatal("Neighbour list capacity did not converge");
    }
    list_gran_dir = (long long)h_counters.p[0]; list_type_dir = (long long)h_counters.p[1];
    list_gran_img = (long long)h_counters.p[2]; list_type_img = (long long)h_counters.p[3];
    Ln.valid = true; Lo.valid = false;
    ecur ^= 1;
    if (count) nbuilds++;
    cell_valid = false;
  }

  // undirected granular list size as LAMMPS counts it (owned-ghost image pairs are stored by both owners)
  long long list_pairs_undirected() const { return (list_gran_dir - list_gran_img) / 2 + list_gran_img; }

  // ---- multi-GPU bootstrap without a host call (softParticleCloud.C:60-62 only hands an MPI_Comm to LAMMPS) ------
  // One rank <-> one GPU.  The NCCL unique id goes from rank 0 to the others
  //   * over the communicator itself when the library is built with -DSEDI_HAVE_MPI (MPI_Bcast), or
  //   * launcher environment + a rendezvous file when SEDI_AUTO_COMM=1 (RANK / WORLD_SIZE of torchrun, OMPI_COMM_WORLD_*,
  //     PMI_*, SLURM_*): "$SEDI_BOOTSTRAP_DIR/sedi_nccl_<MASTER_PORT or job id>_<n-th engine of the process>".
  // A host that calls sedi_comm_init itself (bench.py, the tests) is left alone.
  static int env_int(const char *const *names, int dflt) {
    for (int k = 0; names[k]; k++) if (const char *v = getenv(names[k])) return atoi(v);
    return dflt;
  }
  void auto_comm() {
    if (comm_tried || comm.nranks > 1) return;
    comm_tried = true;
    int world = 1, rank = 0;
    bool via_mpi = false;
#ifdef SEDI_HAVE_MPI
    { int init = 0; MPI_Initialized(&init); if (init) { MPI_Comm_size(mpi_world, &world); MPI_Comm_rank(mpi_world, &rank); via_mpi = world > 1; } }
#endif
    const char *ac = getenv("SEDI_AUTO_COMM");
    if (!via_mpi && ac && atoi(ac) != 0) {
      static const char *const ws[] = {"SEDI_WORLD_SIZE", "WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS", 0};
      static const char *const rk[] = {"SEDI_RANK", "RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID", 0};
      world = env_int(ws, 1); rank = env_int(rk, 0);
    }
    if (world <= 1) return;
    static int instance = 0;
    const int inst = instance++;
    char id[256];
    int nb = 0;
    if (rank == 0) { nb = Comm::unique_id(id, sizeof(id)); if (nb <= 0) fatal("multi-GPU bootstrap: NCCL is not available (libnccl.so.2)"); }
    if (via_mpi) {
#ifdef SEDI_HAVE_MPI
      MPI_Bcast(&nb, 1, MPI_INT, 0, mpi_world);
      MPI_Bcast(id, nb, MPI_BYTE, 0, mpi_world);
#endif
    } else {
      static const char *const job[] = {"SEDI_BOOTSTRAP_TOKEN", "MASTER_PORT", "SLURM_JOB_ID", "OMPI_MCA_ess_base_jobid", 0};
      const char *dir = getenv("SEDI_BOOTSTRAP_DIR");
      char path[1024], tmp[1100];
      snprintf(path, sizeof(path), "%s/sedi_nccl_%d_%d", dir ? dir : "/tmp", env_int(job, 0), inst);
      if (rank == 0) {
        snprintf(tmp, sizeof(tmp), "%s.tmp", path);
        FILE *fp = fopen(tmp, "wb");
        if (!fp) fatal("multi-GPU bootstrap: cannot write the rendezvous file", tmp);
        fwrite(&nb, sizeof(int), 1, fp); fwrite(id, 1, nb, fp); fclose(fp);
        if (rename(tmp, path) != 0) fatal("multi-GPU bootstrap: cannot publish the rendezvous file", path);
      } else {
        FILE *fp = 0;
        for (int k = 0; k < 12000 && !fp; k++) { fp = fopen(path, "rb"); if (!fp) usleep(10000); }
        if (!fp) fatal("multi-GPU bootstrap: rank 0 never published the NCCL id", path);
        if (fread(&nb, sizeof(int), 1, fp) != 1 || nb <= 0 || nb > (int)sizeof(id) || fread(id, 1, nb, fp) != (size_t)nb) fatal("multi-GPU bootstrap: bad rendezvous file", path);
        fclose(fp);
      }
      comm.init(*this, rank, world, id, nb, 0);
      comm.barrier();
      if (rank == 0) unlink(path);
      return;
    }
    comm.init(*this, rank, world, id, nb, 0);
  }

  // ---- Verlet::setup (first `run` of the session, even with `pre no`; softParticleCloud.C:189 lammps_step(0)) -----
  void setup(bool evaluate_forces = true) {
    auto_comm();
    if (restart_pending) { setup_from_restart(); return; }
    if (!loaded) load_atoms();
    need_device();
    if (!cfg().neigh_modify_seen && cfg().pair != PAIR_NONE && !warned_neigh) {
      fprintf(stderr, "WARNING: no neigh_modify command: LAMMPS' default is `delay 10 every 1 check yes`; libsedi_b200 checks the skin/2 "
                      "criterion after every step (`delay 0`, as every shipped in.lammps sets)\n");
      warned_neigh = true;
    }
    dt_init = cfg().dt;
    compute_cutoffs();
    if (cfg().lub.enabled) {  // PairLubricatePoly::init_style, pair_lubricate_poly.cpp:533-559 (vol_T = box volume)
      const AtomData &a = script.atoms;
      const double MY_PI = SEDI_MY_PI;
      const double vol_T = (cfg().boxhi[0] - cfg().boxlo[0]) * (cfg().boxhi[1] - cfg().boxlo[1]) * (cfg().boxhi[2] - cfg().boxlo[2]);
      double volP = 0.0;
      for (size_t i = 0; i < a.size(); i++) volP += (4.0 / 3.0) * MY_PI * pow(a.radius[i], 3.0);
      comm.allreduce_sum_host(&volP, 1);
      double vol_f = volP / vol_T;
      if (!cfg().lub.flagVF) vol_f = 0;
      const double mu = cfg().lub.mu;
      if (cfg().lub.flaglog == 0) {
        lub_R0 = 6 * MY_PI * mu * (1.0 + 2.16 * vol_f);
        lub_RT0 = 8 * MY_PI * mu;
        lub_RS0 = 20.0 / 3.0 * MY_PI * mu * (1.0 + 3.33 * vol_f + 2.80 * vol_f * vol_f);
      } else {
        lub_R0 = 6 * MY_PI * mu * (1.0 + 2.725 * vol_f - 6.583 * vol_f * vol_f);
        lub_RT0 = 8 * MY_PI * mu * (1.0 + 0.749 * vol_f - 2.469 * vol_f * vol_f);
        lub_RS0 = 20.0 / 3.0 * MY_PI * mu * (1.0 + 3.64 * vol_f - 6.95 * vol_f * vol_f);
      }
    }
    setup_bins();
    rebuild();
    build_base_params();
    if (evaluate_forces) {
      launch_step(MODE_SETUP, cur, cfg().ntimestep, 0);
      pair_evals += list_pairs_undirected();
    }
    // the rows are sorted by the work they had under the previous list; the first list has no predecessor, so the same
    // list is built once more now that its touch masks are known (same positions, same pair set: not a LAMMPS re-neighbouring)
    if (sell_sort && (comm.nranks > 1 || nlocal + nghost > 0)) {   // (a rebuild is collective on several GPUs: an empty brick takes part)
      const bool ip = inject_pending;   // the list just built carries the injected history: re-attach from it, not from the arrival lists again
      inject_pending = false;
      rebuild(false);
      inject_pending = ip;
    }
    setup_done = true;
    if (evaluate_forces) setup_done_once = true;
    if (!cfg().dumps.empty()) write_dumps();   // Output::setup writes the initial snapshot
  }

  // ---- `dump custom` (EXTERNAL LAMMPS DumpCustom text format) --------------------------------------
  // Called when HBM holds an end-of-step state (after MODE_SETUP or MODE_LAST): x(n), v(n), f(n), torque(n).
  void write_dumps() {
    SimConfig &c = cfg();
    bool due = false;
    for (size_t k = 0; k < c.dumps.size(); k++) if (c.ntimestep % c.dumps[k].every == 0 && c.dumps[k].last_written != c.ntimestep) due = true;
    if (!due || getenv("SEDI_NO_DUMP")) return;
    const int m = nlocal;
    std::vector<double> x(3 * (size_t)m), v(3 * (size_t)m), w(3 * (size_t)m), fo(3 * (size_t)m), to(3 * (size_t)m), r(m), ms(m);
    std::vector<int> tg(m), ty(m), mk(m);
    if (m) get_state(x.data(), v.data(), w.data(), fo.data(), to.data(), r.data(), ms.data(), tg.data(), ty.data(), mk.data());
    std::vector<int> ord(m);
    for (int i = 0; i < m; i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return tg[a] < tg[b]; });
    for (size_t k = 0; k < c.dumps.size(); k++) {
      DumpSpec &D = c.dumps[k];
      if (c.ntimestep % D.every != 0 || D.last_written == c.ntimestep) continue;
      if (!D.fp) {
        std::string path = D.path;
        const char *dir = getenv("SEDI_DUMP_DIR");
        if (dir && path.size() && path[0] != '/') path = std::string(dir) + "/" + path;
        if (comm.nranks > 1) { char suf[32]; snprintf(suf, sizeof(suf), ".%d", comm.rank); path += suf; }  // one file per GPU rank
        D.fp = fopen(path.c_str(), "w");
        if (!D.fp) fatal("Cannot open dump file", path.c_str());
      }
      long long cnt = 0;
      for (int i = 0; i < m; i++) if (mk[i] & D.groupbit) cnt++;
      fprintf(D.fp, "ITEM: TIMESTEP\n%lld\nITEM: NUMBER OF ATOMS\n%lld\n", c.ntimestep, cnt);
      fprintf(D.fp, "ITEM: BOX BOUNDS %s %s %s\n", c.boundary_str[0].c_str(), c.boundary_str[1].c_str(), c.boundary_str[2].c_str());
      for (int d = 0; d < 3; d++) fprintf(D.fp, "%g %g\n", c.boxlo[d], c.boxhi[d]);
      fprintf(D.fp, "ITEM: ATOMS %s\n", D.columns.c_str());
      for (int q = 0; q < m; q++) {
        const int i = ord[q];
        if (!(mk[i] & D.groupbit)) continue;
        for (size_t f = 0; f < D.fields.size(); f++) {
          switch (D.fields[f]) {
            case DF_ID: fprintf(D.fp, "%d ", tg[i]); break;
            case DF_TYPE: fprintf(D.fp, "%d ", ty[i]); break;
            case DF_DIAMETER: fprintf(D.fp, "%g ", 2.0 * r[i]); break;
            case DF_RADIUS: fprintf(D.fp, "%g ", r[i]); break;
            case DF_MASS: fprintf(D.fp, "%g ", ms[i]); break;
            case DF_X: case DF_Y: case DF_Z: fprintf(D.fp, "%g ", x[3 * (size_t)i + (D.fields[f] - DF_X)]); break;
            case DF_VX: case DF_VY: case DF_VZ: fprintf(D.fp, "%g ", v[3 * (size_t)i + (D.fields[f] - DF_VX)]); break;
            case DF_FX: case DF_FY: case DF_FZ: fprintf(D.fp, "%g ", fo[3 * (size_t)i + (D.fields[f] - DF_FX)]); break;
            case DF_OMEGAX: case DF_OMEGAY: case DF_OMEGAZ: fprintf(D.fp, "%g ", w[3 * (size_t)i + (D.fields[f] - DF_OMEGAX)]); break;
            default: fprintf(D.fp, "%g ", to[3 * (size_t)i + (D.fields[f] - DF_TQX)]); break;
          }
        }
        fputc('\n', D.fp);
      }
      fflush(D.fp);
      D.last_written = c.ntimestep;
    }
  }

  // Verlet::run with Output::write at the dump steps: the sub-step sequence is cut at every dump boundary so that the
  // snapshot is the reference's end-of-step state (the fused kernel otherwise holds x(n+1), v(n+1/2) between launches).
  void run(long long nsteps) {
    if (!setup_done) setup();
    if (nsteps <= 0) return;
    long long remaining = nsteps;
    double ms_total = 0.0;
    while (remaining > 0) {
      long long seg = remaining;
      if (!getenv("SEDI_NO_DUMP"))
        for (size_t k = 0; k < cfg().dumps.size(); k++) {
          const long long ev = cfg().dumps[k].every;
          const long long to_next = ev - (cfg().ntimestep % ev);
          if (to_next < seg) seg = to_next;
        }
      if (restart_every > 0) { const long long to_next = restart_every - (cfg().ntimestep % restart_every); if (to_next < seg) seg = to_next; }
      run_segment(seg);
      ms_total += last_step_ms;
      remaining -= seg;
      if (!cfg().dumps.empty()) write_dumps();
      if (restart_every > 0 && cfg().ntimestep % restart_every == 0) {
        std::string p = restart_path[restart_path[1].empty() ? 0 : restart_toggle];
        if (!restart_path[1].empty()) restart_toggle ^= 1;
        const size_t star = p.find('*');
        if (star != std::string::npos) { char b[32]; snprintf(b, sizeof(b), "%lld", cfg().ntimestep); p.replace(star, 1, b); }
        write_restart(p);
      }
    }
    last_step_ms = ms_total;
  }

  // ---- n DEM sub-steps ending in an end-of-step state, chunked so that the host only synchronises every `chunk` launches
  void run_segment(long long nsteps) {
    if (!setup_done) setup();
    need_device();
    if (params_dirty) build_base_params();
    if (nsteps <= 0) return;
    CK(cudaEventRecord(ev0, stream));
    long long remaining = nsteps;
    bool pending_initial = true;  // the next thing a step needs is its initial_integrate
    while (remaining > 0) {
      CK(cudaMemsetAsync(ctrl.p, 0, 2 * sizeof(int), stream));
      const int K = (int)std::min<long long>(remaining, chunk);
      int seq = 0, in = cur, nk = 0;
      const bool mg = comm.nranks > 1;
      if (pending_initial) { launch_initial(in, ++seq); in ^= 1; nk++; if (mg) comm.forward(*this, in, true); }
      if (prof_on) CK(cudaEventRecord(evk0, stream));
      bool wiggle = false;
      for (size_t k = 0; k < cfg().fixes.size(); k++) if (cfg().fixes[k].kind == FIX_WALL_GRAN && cfg().fixes[k].wiggle) wiggle = true;
      // one CUDA graph per chunk; on several GPUs it includes the halo push / signal kernels of every sub-step (peer-memory
      // path only: their launches carry no per-call argument).  Odd-sized remainders after a rebuild are launched directly.
      if (graph_on && !prof_split && (!mg || comm.p2p) && !wiggle && K == chunk && K > 1) {
        const int lastflag = (remaining == K) ? 1 : 0;
        StepGraph *g = 0;
        for (size_t k = 0; k < graphs.size(); k++) if (graphs[k].in == in && graphs[k].K == K && graphs[k].last == lastflag && graphs[k].seq0 == seq) g = &graphs[k];
        if (!g || g->stale) {
          const long long l0 = launches, h0 = comm.halo_calls;
          cudaGraph_t gr;
          CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
          int cin = in, cseq = seq;
          for (int s = 0; s < K; s++) {
            const bool lastk = (lastflag && s == K - 1);
            launch_step(lastk ? MODE_LAST : MODE_FUSED, cin, cfg().ntimestep + s + 1, ++cseq);
            cin ^= 1;
            if (mg && !lastk) comm.forward(*this, cin, true, step_pushes());
          }
          CK(cudaStreamEndCapture(stream, &gr));
          launches = l0; comm.halo_calls = h0;
          if (g) {  // same topology, new kernel arguments
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(g->exec, gr, &info) != cudaSuccess) {
              cudaGetLastError();
              CK(cudaGraphExecDestroy(g->exec));
              CK(cudaGraphInstantiate(&g->exec, gr, 0));
            }
            g->stale = false;
          } else {
            StepGraph ng; ng.in = in; ng.K = K; ng.last = lastflag; ng.seq0 = seq; ng.stale = false;
            CK(cudaGraphInstantiate(&ng.exec, gr, 0));
            if (graphs.size() >= 16) destroy_graphs();
            graphs.push_back(ng);
            g = &graphs.back();
          }
          CK(cudaGraphDestroy(gr));
        }
        CK(cudaGraphLaunch(g->exec, stream));
        launches += K; seq += K; in ^= (K & 1);
        if (mg) { const int nfw = K - (lastflag ? 1 : 0); launches += (step_pushes() ? 1 : 2) * nfw; comm.halo_calls += nfw; }
      } else
      for (int s = 0; s < K; s++) {
        const bool last = (remaining - s == 1);
        if (prof_split) { while ((int)split_ev.size() < 3 * chunk + 3) { cudaEvent_t ev; CK(cudaEventCreate(&ev)); split_ev.push_back(ev); } CK(cudaEventRecord(split_ev[3 * s], stream)); }
        launch_step(last ? MODE_LAST : MODE_FUSED, in, cfg().ntimestep + s + 1, ++seq);
        if (prof_split) CK(cudaEventRecord(split_ev[3 * s + 1], stream));
        in ^= 1;
        if (mg && !last) comm.forward(*this, in, true, step_pushes());  // ghost x, v, omega of the new positions (written by the step kernel itself on the peer-memory path) + rebuild consensus
        if (prof_split) CK(cudaEventRecord(split_ev[3 * s + 2], stream));
      }
      if (prof_on) CK(cudaEventRecord(evk1, stream));
      CK(cudaMemcpyAsync(h_ctrl.p, ctrl.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      CK(cudaGetLastError());
      const int flag = h_ctrl.p[0], done = h_ctrl.p[1];
      if (prof_split) {
        for (int s = 0; s < K && 3 * s + 2 < (int)split_ev.size(); s++) {
          float a = 0.f, b = 0.f;
          if (cudaEventElapsedTime(&a, split_ev[3 * s], split_ev[3 * s + 1]) == cudaSuccess && cudaEventElapsedTime(&b, split_ev[3 * s + 1], split_ev[3 * s + 2]) == cudaSuccess) {
            split_step_ms += a; split_halo_ms += b; split_n++;
          } else cudaGetLastError();
        }
      }
      if (prof_on) {  // the K k_step launches are back to back on the stream: their summed duration / executed count
        float kms = 0.f;
        CK(cudaEventElapsedTime(&kms, evk0, evk1));
        prof_ms += kms; prof_steps += done;
      }
      if (h_ctrl.p[2] & WQ_ERR_QUEUE) fatal("Contact queue overflow: more than 24 overlapping partners per particle on average in one warp of rows");
      if (h_ctrl.p[2]) fatal("Device-side error flag raised during the DEM step");
      if ((nk + done) & 1) cur ^= 1;
      pair_evals += (long long)done * list_pairs_undirected();
      pair_evals_unique += (long long)done * list_gran_dir;  // halves: every undirected pair has two directed entries system-wide
      steps_done += done;
      remaining -= done;
      cfg().ntimestep += done;
      pending_initial = false;
      if (flag != 0) {
        if (remaining <= 0) fatal("internal: rebuild requested after the last sub-step");
        rebuild();   // positions of step ntimestep+1 are already integrated; its forces come next
      } else if (done != K) fatal("internal: sub-step chunk ended early without a rebuild request");
    }
    CK(cudaEventRecord(ev1, stream));
    CK(cudaEventSynchronize(ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ev0, ev1));
    last_step_ms = ms;
    cell_valid = false;
  }

  // ---- C-ABI data movement ---------------------------------------------------------------------------------------
  // The unchanged host (softParticleCloud.C:908-912, 959-963) hands over pageable `new double[]` arrays: they are staged through
  // page-locked buffers.  The host-side copy between the two is split over a few threads (one core moves ~10 GB/s, the DMA 25+),
  // and on the way back every array is copied out while the next one is still arriving.
  static void par_memcpy(void *dst, const void *src, size_t bytes) {
    const size_t MINB = (size_t)4 << 20;
    int nt = (int)std::min<size_t>(4, bytes / MINB);
    if (const char *e = getenv("SEDI_COPY_THREADS")) nt = std::max(1, std::min(16, atoi(e)));
    if (nt <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
    for (int t = 1; t < nt; t++) {
      const size_t o = per * t;
      if (o >= bytes) break;
      const size_t len = std::min(per, bytes - o);
      th.emplace_back([=]() { memcpy((char *)dst + o, (const char *)src + o, len); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (size_t t = 0; t < th.size(); t++) th[t].join();
  }
  static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  }
  void put_local(int nin, const double *fd, const int *foamin, const int *tagin) {
    if (!loaded) load_atoms();
    need_device();
    if (nin != nlocal) {  // the reference only warns and then indexes out of range (library.cpp:335-341): fatal here
      fprintf(stderr, "Incoming drag not consistent with local particle number.\nIncoming drag is: %5d, local particle number is: %5d.\n", nin, nlocal);
      fatal("lammps_put_local_info: particle count mismatch");
    }
    if (!nin) return;
    if (!setup_done) setup();  // tag2idx is filled by the first build
    d_stage_a.ensure(3 * (size_t)nin); d_stage_i.ensure(nin); d_stage_j.ensure(nin);
    // caller arrays that are already page-locked go straight to the copy engine; pageable ones are staged
    const void *src_a = fd, *src_i = tagin, *src_j = foamin;
    if (!is_pinned(fd)) { h_stage_a.ensure(3 * (size_t)nin); par_memcpy(h_stage_a.p, fd, 3 * (size_t)nin * sizeof(double)); src_a = h_stage_a.p; }
    if (!is_pinned(tagin)) { h_stage_i.ensure(nin); memcpy(h_stage_i.p, tagin, nin * sizeof(int)); src_i = h_stage_i.p; }
    if (foamin && !is_pinned(foamin)) { h_stage_j.ensure(nin); memcpy(h_stage_j.p, foamin, nin * sizeof(int)); src_j = h_stage_j.p; }
    CK(cudaMemcpyAsync(d_stage_a.p, src_a, 3 * (size_t)nin * sizeof(double), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_stage_i.p, src_i, nin * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (foamin) CK(cudaMemcpyAsync(d_stage_j.p, src_j, nin * sizeof(int), cudaMemcpyHostToDevice, stream));
    k_put_fdrag<<<cdiv(nin, 256), 256, 0, stream>>>(nin, d_stage_a.p, d_stage_i.p, foamin ? d_stage_j.p : 0, tag2idx.p, maxtag,
                                                    fdrag[0].get(), fdrag[1].get(), fdrag[2].get(), foam[icur].p, ctrl.p + 2);
    launches++;
    CK(cudaMemcpyAsync(h_ctrl.p, ctrl.p, 3 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (h_ctrl.p[2]) fatal("lammps_put_local_info: incoming tag not owned by this rank");
  }

  void get_local(double *x, double *v, int *foamid, int *lmpid, int *tag) {
    if (!loaded) load_atoms();
    need_device();
    const int m = nlocal;
    if (!m) return;
    d_stage_a.ensure(3 * (size_t)m); d_stage_b.ensure(3 * (size_t)m); d_stage_i.ensure(m); d_stage_j.ensure(m);
    k_pack_local<<<cdiv(m, 256), 256, 0, stream>>>(posr[cur].p, velm[cur].p, omgt[cur].p, foam[icur].p, m, d_stage_a.p, d_stage_b.p,
                                                   d_stage_i.p, d_stage_j.p);
    launches++;
    // page-locked caller arrays receive the DMA directly; pageable ones go through the pinned staging buffers
    const bool px = x && is_pinned(x), pv = v && is_pinned(v), pt = tag && is_pinned(tag), pf = foamid && is_pinned(foamid);
    if (x && !px) h_stage_a.ensure(3 * (size_t)m);
    if (v && !pv) h_stage_b.ensure(3 * (size_t)m);
    if (tag && !pt) h_stage_i.ensure(m);
    if (foamid && !pf) h_stage_j.ensure(m);
    if (x) { CK(cudaMemcpyAsync(px ? (void *)x : (void *)h_stage_a.p, d_stage_a.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream)); CK(cudaEventRecord(ev_get[0], stream)); }
    if (v) { CK(cudaMemcpyAsync(pv ? (void *)v : (void *)h_stage_b.p, d_stage_b.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream)); CK(cudaEventRecord(ev_get[1], stream)); }
    if (tag) CK(cudaMemcpyAsync(pt ? (void *)tag : (void *)h_stage_i.p, d_stage_i.p, m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (foamid) CK(cudaMemcpyAsync(pf ? (void *)foamid : (void *)h_stage_j.p, d_stage_j.p, m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (lmpid) for (int i = 0; i < m; i++) lmpid[i] = comm.rank;   // (while the first array arrives)
    if (x && !px) { CK(cudaEventSynchronize(ev_get[0])); par_memcpy(x, h_stage_a.p, 3 * (size_t)m * sizeof(double)); }   // v is still arriving
    if (v && !pv) { CK(cudaEventSynchronize(ev_get[1])); par_memcpy(v, h_stage_b.p, 3 * (size_t)m * sizeof(double)); }
    CK(cudaStreamSynchronize(stream));
    if (tag && !pt) memcpy(tag, h_stage_i.p, m * sizeof(int));
    if (foamid && !pf) memcpy(foamid, h_stage_j.p, m * sizeof(int));
  }

  // full state in device row order (owned rows first `nlocal` rows are owned; ghosts follow)
  void get_state(double *x, double *v, double *w, double *fo, double *to, double *radius, double *rmass, int *tag, int *type, int *mask) {
    if (!loaded) load_atoms();
    need_device();
    const int m = nlocal;
    if (!m) return;
    Buf<double> dx, dv, dw, dr, dm, df;
    Buf<int> dt, dy, dk;
    dx.ensure(3 * (size_t)m); dv.ensure(3 * (size_t)m); dw.ensure(3 * (size_t)m); dr.ensure(m); dm.ensure(m); df.ensure(3 * (size_t)m);
    dt.ensure(m); dy.ensure(m); dk.ensure(m);
    k_unpack_state<<<cdiv(m, 256), 256, 0, stream>>>(posr[cur].p, velm[cur].p, omgt[cur].p, m, dx.p, dv.p, dw.p, dr.p, dm.p, dt.p, dy.p, dk.p);
    if (x) CK(cudaMemcpyAsync(x, dx.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (v) CK(cudaMemcpyAsync(v, dv.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (w) CK(cudaMemcpyAsync(w, dw.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (radius) CK(cudaMemcpyAsync(radius, dr.p, m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (rmass) CK(cudaMemcpyAsync(rmass, dm.p, m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (tag) CK(cudaMemcpyAsync(tag, dt.p, m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (type) CK(cudaMemcpyAsync(type, dy.p, m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (mask) CK(cudaMemcpyAsync(mask, dk.p, m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (fo) {
      k_interleave3<<<cdiv(m, 256), 256, 0, stream>>>(f[0].get(), f[1].get(), f[2].get(), m, df.p);
      CK(cudaMemcpyAsync(fo, df.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
    }
    if (to) {
      k_interleave3<<<cdiv(m, 256), 256, 0, stream>>>(tq[0].get(), tq[1].get(), tq[2].get(), m, df.p);
      CK(cudaMemcpyAsync(to, df.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
    }
    dx.release(); dv.release(); dw.release(); dr.release(); dm.release(); df.release(); dt.release(); dy.release(); dk.release();
  }

  long long get_pairs(int *ti, int *tj, unsigned *meta, int *touch, double *shear, long long capacity) {
    if (!setup_done) setup();
    need_device();
    Ell &L = ell[ecur];
    const int n = nlocal;
    std::vector<int> hn(n), ht(n), rs(n + 1, 0);
    if (n) CK(cudaMemcpyAsync(hn.data(), L.nn.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (n) CK(cudaMemcpyAsync(ht.data(), L.nt.p, n * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    for (int i = 0; i < n; i++) rs[i + 1] = rs[i] + hn[i] + ht[i];
    const long long total = rs[n];
    if (capacity < total || !total) return total;
    Buf<int> dti, dtj, dto;
    Buf<unsigned> dme;
    Buf<double> dsh;
    dti.ensure(total); dtj.ensure(total); dto.ensure(total); dme.ensure(total); dsh.ensure(3 * (size_t)total);
    CK(cudaMemcpyAsync(rowstart.p, rs.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    k_export_pairs<<<cdiv(n, 128), 128, 0, stream>>>(n, L.npad, L.nn.p, L.nt.p, L.cap, L.nbr.p, omgt[cur].p, rowstart.p, dti.p, dtj.p, dme.p, L.tmask.p,
                                                     L.shear.p, dto.p, dsh.p);
    if (ti) CK(cudaMemcpyAsync(ti, dti.p, total * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (tj) CK(cudaMemcpyAsync(tj, dtj.p, total * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (meta) CK(cudaMemcpyAsync(meta, dme.p, total * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    if (touch) CK(cudaMemcpyAsync(touch, dto.p, total * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (shear) CK(cudaMemcpyAsync(shear, dsh.p, 3 * (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    dti.release(); dtj.release(); dto.release(); dme.release(); dsh.release();
    return total;
  }

  // per-row list statistics: entries of every owned row and how many of them overlapped in the last sub-step
  void get_row_stats(int *nn_out, int *ntouch_out) {
    if (!setup_done) setup();
    need_device();
    const int m = nlocal;
    if (!m) return;
    Ell &L = ell[ecur];
    std::vector<unsigned long long> tm(m);
    std::vector<int> ht(m);
    if (nn_out) CK(cudaMemcpyAsync(nn_out, L.nn.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(ht.data(), L.nt.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(tm.data(), L.tmask.p, (size_t)m * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (nn_out) for (int i = 0; i < m; i++) nn_out[i] += ht[i];
    if (ntouch_out) for (int i = 0; i < m; i++) ntouch_out[i] = __builtin_popcountll(tm[i]);
  }

  void get_wall_shear(int w, double *out) {
    need_device();
    if (w < 0 || w >= cfg().nwalls) fatal("sedi_get_wall_shear: no such wall");
    const int m = nlocal;
    if (!m) return;
    Buf<double> d;
    d.ensure(3 * (size_t)m);
    // rows whose wall-touch bit is clear hold stale values; the reference zeroes them eagerly (:326-331)
    k_wall_shear_export<<<cdiv(m, 256), 256, 0, stream>>>(wshear[w][0].get(), wshear[w][1].get(), wshear[w][2].get(), wmask[icur].p, w, m, d.p);
    CK(cudaMemcpyAsync(out, d.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    d.release();
  }

  // history-force state in device row order (rows match sedi_get_state)
  void get_history_state(double *sum, double *n0) {
    need_device();
    const int m = nlocal;
    if (!m) return;
    if (!hist_alloc) { if (sum) memset(sum, 0, 3 * (size_t)m * sizeof(double)); if (n0) memset(n0, 0, (size_t)m * sizeof(double)); return; }
    Buf<double> d;
    d.ensure(3 * (size_t)m);
    if (sum) {
      k_interleave3<<<cdiv(m, 256), 256, 0, stream>>>(hist[0].get(), hist[1].get(), hist[2].get(), m, d.p);
      CK(cudaMemcpyAsync(sum, d.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    if (n0) CK(cudaMemcpyAsync(n0, hist[3].get(), (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    d.release();
  }

  void set_omega(int m, const int *tag, const double *w) {
    if (!loaded) {   // before the first run the atoms still live in the script: initial spins of the data set
      AtomData &a = script.atoms;
      std::vector<std::pair<int, int> > idx(a.size());
      for (size_t i = 0; i < a.size(); i++) idx[i] = std::make_pair(a.tag[i], (int)i);
      std::sort(idx.begin(), idx.end());
      for (int k = 0; k < m; k++) {
        std::vector<std::pair<int, int> >::iterator it = std::lower_bound(idx.begin(), idx.end(), std::make_pair(tag[k], -1));
        if (it == idx.end() || it->first != tag[k]) continue;
        for (int d = 0; d < 3; d++) a.omega[3 * (size_t)it->second + d] = w[3 * (size_t)k + d];
      }
      return;
    }
    if (!setup_done) setup();
    Buf<int> dt; Buf<double> dw;
    dt.ensure(m); dw.ensure(3 * (size_t)m);
    CK(cudaMemcpyAsync(dt.p, tag, m * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(dw.p, w, 3 * (size_t)m * sizeof(double), cudaMemcpyHostToDevice, stream));
    k_set_omega<<<cdiv(m, 256), 256, 0, stream>>>(m, dt.p, dw.p, tag2idx.p, maxtag, omgt[cur].p);
    CK(cudaStreamSynchronize(stream));
    dt.release(); dw.release();
  }

  // ---- coupling (device-resident mirror of enhancedCloud) -----------------------------------------------------------
  void mesh_box(const double *lo, const double *hi, const int *nc) {
    need_device();
    long long C = 1;
    for (int d = 0; d < 3; d++) { mesh.lo[d] = lo[d]; mesh.hi[d] = hi[d]; mesh.nc[d] = nc[d]; mesh.dx[d] = (hi[d] - lo[d]) / nc[d]; C *= nc[d]; }
    mesh.rect = 0; mesh.label = 0; mesh.face[0] = mesh.face[1] = mesh.face[2] = 0;
    if (C <= 0 || C > 0x7fffffff) fatal("sedi_mesh_box: bad cell count");
    ncells = (int)C;
    Uf.ensure(3 * (size_t)C); gamma.ensure(C); gradp.ensure(3 * (size_t)C); DDtU.ensure(3 * (size_t)C); curlU.ensure(3 * (size_t)C);
    cellV.ensure(C); Ue.ensure(3 * (size_t)C); Asrc.ensure(3 * (size_t)C);
    CK(cudaMemsetAsync(Uf.p, 0, 3 * (size_t)C * sizeof(double), stream));
    CK(cudaMemsetAsync(gamma.p, 0, (size_t)C * sizeof(double), stream));
    CK(cudaMemsetAsync(gradp.p, 0, 3 * (size_t)C * sizeof(double), stream));
    const double V = mesh.dx[0] * mesh.dx[1] * mesh.dx[2];
    k_fill_double<<<cdiv(C, 256), 256, 0, stream>>>(cellV.p, (size_t)C, V);
    have_mesh = true; have_DDtU = have_curlU = false; have_gradp = true; cell_valid = false;
  }

  // Rectilinear mesh (graded blocks, e.g. `simpleGrading (1 10 1)` of cases/example-cases/transport-bedload,
  // and axis-aligned blocks stacked into one tensor-product grid, cases/example-cases/BL24-TH1): face coordinates
  // per axis as the host mesh has them, plus the host's cell label of every tensor cell (NULL = i + nx (j + ny k)).
  Buf<double> mesh_faces[3], mesh_width[3], cg_diag;
  // CUDA graphs of the k_step launches of one chunk (single GPU, no wiggling wall): one graph launch per chunk, so the
  // sub-step sequence does not depend on the host keeping up with 50 launches.  The kernel arguments baked into a graph
  // (buffer parity, sequence numbers, list / plane pointers) stay valid until the next rebuild or parameter change.
  struct StepGraph { int in, K, last, seq0; bool stale; cudaGraphExec_t exec; };
  std::vector<StepGraph> graphs;
  bool graph_on;
  void destroy_graphs() { for (size_t k = 0; k < graphs.size(); k++) cudaGraphExecDestroy(graphs[k].exec); graphs.clear(); }
  // pointers / parameters changed (rebuild, new script line): the executable graphs are kept and re-parameterised in
  // place on their next use (cudaGraphExecUpdate), which is much cheaper than instantiating them again
  void drop_graphs() { for (size_t k = 0; k < graphs.size(); k++) graphs[k].stale = true; }
  // checkpoint / resume (`restart N file`, `write_restart`, `read_restart`): own binary format, see write_restart
  long long restart_every;
  std::string restart_path[2];
  int restart_toggle;
  bool restart_pending;
  struct RowCarryBox;             // defined with RowCarry below
  RowCarryBox *restart_carry;
  std::vector<std::string> restart_wall_ids;
  // particle injection / deletion: state of the surviving particles carried across the re-upload
  std::vector<char> keep_local;     // per atom of script.atoms: 1 = came from this rank's device rows (stays here whatever the brick says: the
                                    // next migration moves it), 0 = new / read from a file (kept only by the brick that owns its position)
  std::vector<int> loaded_index;    // device row (upload order) -> index in script.atoms
  bool inject_pending;
  Buf<int> inj_nh, inj_tag;
  Buf<D4> inj_shear;
  Buf<int> mesh_label;
  void mesh_rectilinear(const int *nc, const double *xf, const double *yf, const double *zf, const int *label) {
    need_device();
    const double *f[3] = {xf, yf, zf};
    long long C = 1;
    for (int d = 0; d < 3; d++) {
      if (nc[d] < 1) fatal("sedi_mesh_rectilinear: bad cell count");
      for (int k = 0; k < nc[d]; k++) if (!(f[d][k + 1] > f[d][k])) fatal("sedi_mesh_rectilinear: face coordinates must ascend");
      C *= nc[d];
    }
    if (C > 0x7fffffff) fatal("sedi_mesh_rectilinear: bad cell count");
    const int ncu[3] = {nc[0], nc[1], nc[2]};
    double lo[3], hi[3];
    for (int d = 0; d < 3; d++) { lo[d] = f[d][0]; hi[d] = f[d][nc[d]]; }
    mesh_box(lo, hi, ncu);   // allocates the cell fields; dx / cellV are overwritten below
    std::vector<double> V((size_t)C);
    std::vector<char> seen((size_t)C, 0);
    for (int k = 0; k < nc[2]; k++) for (int j = 0; j < nc[1]; j++) for (int i = 0; i < nc[0]; i++) {
      const long long t = i + (long long)nc[0] * (j + (long long)nc[1] * k);
      const long long c = label ? label[t] : t;
      if (c < 0 || c >= C || seen[c]) fatal("sedi_mesh_rectilinear: cell labels must be a permutation of 0..C-1");
      seen[c] = 1;
      V[c] = (xf[i + 1] - xf[i]) * (yf[j + 1] - yf[j]) * (zf[k + 1] - zf[k]);
    }
    CK(cudaMemcpyAsync(cellV.p, V.data(), (size_t)C * sizeof(double), cudaMemcpyHostToDevice, stream));
    for (int d = 0; d < 3; d++) {
      mesh_faces[d].ensure((size_t)nc[d] + 1);
      CK(cudaMemcpyAsync(mesh_faces[d].p, f[d], ((size_t)nc[d] + 1) * sizeof(double), cudaMemcpyHostToDevice, stream));
      mesh.face[d] = mesh_faces[d].p;
      std::vector<double> h((size_t)nc[d]);
      for (int k = 0; k < nc[d]; k++) h[k] = f[d][k + 1] - f[d][k];
      mesh_width[d].ensure((size_t)nc[d]);
      CK(cudaMemcpy(mesh_width[d].p, h.data(), (size_t)nc[d] * sizeof(double), cudaMemcpyHostToDevice));
    }
    mesh.label = 0;
    if (label) { mesh_label.ensure((size_t)C); CK(cudaMemcpyAsync(mesh_label.p, label, (size_t)C * sizeof(int), cudaMemcpyHostToDevice, stream)); mesh.label = mesh_label.p; }
    CK(cudaStreamSynchronize(stream));
    mesh.rect = 1;
  }

  void put_cell_fields(const double *hUf, const double *hgamma, const double *hgradp, const double *hDDtU, const double *hcurlU) {
    if (!have_mesh) fatal("sedi_put_cell_fields: call sedi_mesh_box first");
    need_device();
    const size_t C = ncells;
    if (hUf && (force_flags & SEDI_FORCE_HISTORY)) {  // the field about to be replaced becomes UfSmoothed_.oldTime()
      UfOld.ensure(3 * C);
      if (have_Uf) { CK(cudaMemcpyAsync(UfOld.p, Uf.p, 3 * C * sizeof(double), cudaMemcpyDeviceToDevice, stream)); have_UfOld = true; }
    }
    if (hUf) { CK(cudaMemcpyAsync(Uf.p, hUf, 3 * C * sizeof(double), cudaMemcpyHostToDevice, stream)); have_Uf = true; }
    if (hgamma) CK(cudaMemcpyAsync(gamma.p, hgamma, C * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (hgradp) CK(cudaMemcpyAsync(gradp.p, hgradp, 3 * C * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (hDDtU) { CK(cudaMemcpyAsync(DDtU.p, hDDtU, 3 * C * sizeof(double), cudaMemcpyHostToDevice, stream)); have_DDtU = true; }
    if (hcurlU) { CK(cudaMemcpyAsync(curlU.p, hcurlU, 3 * C * sizeof(double), cudaMemcpyHostToDevice, stream)); have_curlU = true; }
    CK(cudaStreamSynchronize(stream));  // caller may reuse / free its arrays (pageable copies are staged by the driver)
  }

  void locate() {
    if (!have_mesh) fatal("sedi_locate: call sedi_mesh_box first");
    if (!loaded) load_atoms();
    need_device();
    const double t0 = now_s();
    if (nlocal) k_locate_cells<<<cdiv(nlocal, 256), 256, 0, stream>>>(posr[cur].p, nlocal, mesh, cell.p);
    launches++;
    {  // rows of every fluid cell in ascending order: the scatter kernels sum a cell's particles in this fixed order
      const int C = (int)ncells, nscan = C + 1, nblk = cdiv(nscan, SCAN_ITEMS);
      fc_count.ensure((size_t)C + 2); fc_start.ensure((size_t)C + 2); fc_fill.ensure((size_t)C + 2); fc_rows.ensure(npad);
      fc_blocksum.ensure((size_t)nblk + 1);
      CK(cudaMemsetAsync(fc_count.p, 0, ((size_t)C + 2) * sizeof(int), stream));
      CK(cudaMemsetAsync(fc_fill.p, 0, ((size_t)C + 2) * sizeof(int), stream));
      if (nlocal) k_fcell_count<<<cdiv(nlocal, 256), 256, 0, stream>>>(cell.p, nlocal, fc_count.p);
      k_scan_local<<<nblk, 1024, 0, stream>>>(fc_count.p, fc_start.p, nscan, fc_blocksum.p);
      k_scan_sums<<<1, 1024, 0, stream>>>(fc_blocksum.p, nblk);
      k_scan_add<<<cdiv(nscan, 256), 256, 0, stream>>>(fc_start.p, nscan, fc_blocksum.p, 0);
      if (nlocal) {
        k_fcell_fill<<<cdiv(nlocal, 256), 256, 0, stream>>>(cell.p, nlocal, fc_start.p, fc_fill.p, fc_rows.p);
        k_fcell_sort<<<cdiv(C, 128), 128, 0, stream>>>(fc_start.p, C, fc_rows.p);
      }
      launches += 6;
    }
    cell_valid = true;
    if (want_sums) { CK(cudaStreamSynchronize(stream)); }   // timing mode: the host's log attributes the time to this call
    timers[2] += now_s() - t0;
  }

  void fluid_force() {
    if (!setup_done) setup();
    if (!cell_valid) locate();
    if ((force_flags & SEDI_FORCE_ADDEDMASS) && !have_DDtU) fatal("added-mass force needs DDtU");
    if ((force_flags & SEDI_FORCE_LIFT) && !have_curlU) fatal("lift force needs curlU");
    if (force_flags & SEDI_FORCE_HISTORY) {
      if (!hist_alloc) {  // softParticle starts with n0 = 0, sumDeltaFb = 0 (softParticle.C:63-64)
        for (int d = 0; d < 4; d++) for (int b = 0; b < 2; b++) { hist[d].b[b].ensure(npad); CK(cudaMemsetAsync(hist[d].b[b].p, 0, (size_t)npad * sizeof(double), stream)); }
        hist_alloc = true;
      }
      if (!have_UfOld) {  // first step: oldTime() == current field
        UfOld.ensure(3 * (size_t)ncells);
        CK(cudaMemcpyAsync(UfOld.p, Uf.p, 3 * (size_t)ncells * sizeof(double), cudaMemcpyDeviceToDevice, stream));
        have_UfOld = true;
      }
    }
    ForceParams P;
    memset(&P, 0, sizeof(P));
    P.n = nlocal; P.model = drag_model; P.flags = force_flags;
    P.posr = posr[cur].p; P.velm = velm[cur].p; P.cell = cell.p;
    P.Uf = Uf.p; P.gamma = gamma.p; P.gradp = gradp.p; P.DDtU = have_DDtU ? DDtU.p : 0; P.curlU = have_curlU ? curlU.p : 0;
    for (int d = 0; d < 3; d++) { P.uold[d] = uold[d].get(); P.fdrag[d] = fdrag[d].get(); P.dudt[d] = dudt[d].get(); }
    if (want_diag) {
      dg_Uri.ensure(3 * (size_t)npad); dg_mag.ensure(npad); dg_alpha.ensure(npad); dg_Jd.ensure(npad);
      P.Uri = dg_Uri.p; P.magUri = dg_mag.p; P.alphap = dg_alpha.p; P.Jd = dg_Jd.p;
    }
    P.nub = nub; P.rhob = rhob; P.deltaT = deltaT;
    for (int d = 0; d < 3; d++) P.g[d] = gvec[d];
    P.UfOld = UfOld.p; P.timeIndex = time_index;
    if (hist_alloc) { for (int d = 0; d < 3; d++) P.hsum[d] = hist[d].get(); P.hn0 = hist[3].get(); }
    for (int d = 0; d < 3; d++) { P.inletForce[d] = inlet_force[d]; P.inletEcc[d] = inlet_ecc[d]; }
    for (int d = 0; d < 9; d++) P.inletBox[d] = inlet_box[d];
    P.inletOption = inlet_option;
    // no auto-increment: evolve() evaluates the force once per sub-cycle and every evaluation of one fluid step sees the
    // same runTime().timeIndex() (enhancedCloud.C:197-234); the host sets it with sedi_coupling_time_index
    if (nlocal) k_particle_force<<<cdiv(nlocal, 256), 256, 0, stream>>>(P);
    launches++;
  }

  // ---- smoothField: `diffusionSteps` implicit-Euler diffusion steps, each one SPD 7-point solve by Jacobi-PCG ------------
  bool smoothing_on(int flag) const { return (smooth_flags & flag) && smooth_b > 0.0 && smooth_steps > 0; }
  void dot_to(const double *a, int sa, int oa, const double *b, int sb, int ob, int nC, double *out) {
    const int nb = std::min(1024, cdiv(nC, 256));
    k_dot_partial<<<nb, 256, 0, stream>>>(a, sa, oa, b, sb, ob, nC, cg_partial.p, mesh.rect ? cellV.p : (const double *)0);
    k_dot_final<<<1, 256, 0, stream>>>(cg_partial.p, nb, out);
    launches += 2;
  }
  void smooth_component(double *field, int stride, int off) {
    const int C = ncells, T = 256;
    cg_r.ensure(C); cg_z.ensure(C); cg_p.ensure(C); cg_Ap.ensure(C); cg_partial.ensure(1024); cg_s.ensure(8); h_cg.ensure(8);
    SmoothGrid G;
    G.nx = mesh.nc[0]; G.ny = mesh.nc[1]; G.nz = mesh.nc[2];
    const double dtau = (smooth_b * smooth_b / 4) / (smooth_steps + 1.0e-150);   // enhancedCloud.C:564-565
    G.wx = dtau * smooth_D[0] / (mesh.dx[0] * mesh.dx[0]); G.wy = dtau * smooth_D[1] / (mesh.dx[1] * mesh.dx[1]); G.wz = dtau * smooth_D[2] / (mesh.dx[2] * mesh.dx[2]);
    const int nblk = cdiv(C, T);
    G.rect = mesh.rect; G.hx = G.hy = G.hz = 0; G.label = 0; G.diag = 0; G.dtx = G.dty = G.dtz = 0.0;
    if (mesh.rect) {  // finite-volume coefficients from the cell widths; the diagonal is tabulated once per call
      G.dtx = dtau * smooth_D[0]; G.dty = dtau * smooth_D[1]; G.dtz = dtau * smooth_D[2];
      G.hx = mesh_width[0].p; G.hy = mesh_width[1].p; G.hz = mesh_width[2].p; G.label = mesh.label;
      cg_diag.ensure(C);
      k_smooth_apply_rect<<<nblk, T, 0, stream>>>(G, (const double *)0, (double *)0, 1, 0, cg_diag.p);
      G.diag = cg_diag.p;
      launches++;
    }
    auto apply = [&](const double *x, double *y, int st, int of) {
      if (mesh.rect) k_smooth_apply_rect<<<nblk, T, 0, stream>>>(G, x, y, st, of, (double *)0);
      else k_smooth_apply<<<nblk, T, 0, stream>>>(G, x, y, st, of);
    };
    for (int step = 0; step < smooth_steps; step++) {
      // A x = b with b = field (also the initial guess); s = {rz, pAp, rz_new, rr, bb}
      dot_to(field, stride, off, field, stride, off, C, cg_s.p + 4);
      apply(field, cg_Ap.p, stride, off);
      k_cg_init<<<nblk, T, 0, stream>>>(G, field, stride, off, cg_Ap.p, cg_r.p, cg_z.p, cg_p.p, C);
      dot_to(cg_r.p, 1, 0, cg_z.p, 1, 0, C, cg_s.p + 0);
      dot_to(cg_r.p, 1, 0, cg_r.p, 1, 0, C, cg_s.p + 3);
      launches += 2;
      int it = 0;
      for (; it < 500; it++) {
        if ((it & 3) == 0) {  // convergence check: ||r|| <= 1e-12 ||b||
          CK(cudaMemcpyAsync(h_cg.p, cg_s.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, stream));
          CK(cudaStreamSynchronize(stream));
          if (!(h_cg.p[3] > 1.0e-24 * h_cg.p[4]) || h_cg.p[0] == 0.0) break;
        }
        apply(cg_p.p, cg_Ap.p, 1, 0);
        dot_to(cg_p.p, 1, 0, cg_Ap.p, 1, 0, C, cg_s.p + 1);
        k_cg_step1<<<nblk, T, 0, stream>>>(field, stride, off, cg_r.p, cg_p.p, cg_Ap.p, cg_s.p, G, cg_z.p, C);
        dot_to(cg_r.p, 1, 0, cg_z.p, 1, 0, C, cg_s.p + 2);
        dot_to(cg_r.p, 1, 0, cg_r.p, 1, 0, C, cg_s.p + 3);
        k_cg_step2<<<nblk, T, 0, stream>>>(cg_p.p, cg_z.p, cg_s.p, C);
        k_cg_shift<<<1, 1, 0, stream>>>(cg_s.p);
        launches += 4;
      }
      smooth_iters_last = it;
    }
  }
  void smooth_device_field(double *f, int ncomp) {
    const double t0 = now_s();
    for (int k = 0; k < ncomp; k++) smooth_component(f, ncomp, k);
    CK(cudaStreamSynchronize(stream));
    timers[1] += now_s() - t0;
  }
  // UfSmoothed = Uf (1-gamma) -> smooth -> / (1-gamma)   (enhancedCloud.C:675-690)
  void smooth_uf() {
    if (!have_mesh) fatal("sedi_smooth_uf: call sedi_mesh_box first");
    need_device();
    if (!smoothing_on(1)) return;
    k_scale_one_minus_gamma<<<cdiv(ncells, 256), 256, 0, stream>>>(ncells, gamma.p, Uf.p, 3, 0);
    smooth_device_field(Uf.p, 3);
    k_scale_one_minus_gamma<<<cdiv(ncells, 256), 256, 0, stream>>>(ncells, gamma.p, Uf.p, 3, 1);
    launches += 2;
  }
  void smooth_host_field(double *h, int ncomp) {  // test / host hook: smooth a host array in place
    if (!have_mesh) fatal("sedi_smooth_field: call sedi_mesh_box first");
    need_device();
    cg_tmp.ensure((size_t)ncells * ncomp);
    CK(cudaMemcpyAsync(cg_tmp.p, h, (size_t)ncells * ncomp * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (smooth_b > 0.0 && smooth_steps > 0) smooth_device_field(cg_tmp.p, ncomp);
    CK(cudaMemcpyAsync(h, cg_tmp.p, (size_t)ncells * ncomp * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }

  // sum over cells of a [C][3] field with the reference's weights (see k_field_sum_partial); all ranks hold the full field
  void field_sum(const double *f, int mode, double *out3) {
    const int C = ncells, nb = std::min(256, cdiv(C, 256));
    sum_partial.ensure(4 * 256); sum_out.ensure(4);
    k_field_sum_partial<<<nb, 256, 0, stream>>>(C, f, cellV.p, gamma.p, mode, sum_partial.p);
    k_sum_final<<<1, 256, 0, stream>>>(sum_partial.p, nb, 3, sum_out.p);
    CK(cudaMemcpyAsync(out3, sum_out.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    launches += 2;
  }
  // enhancedCloud::averageInfo (enhancedCloud.C:1341-1370): total particle volume, sum Vp U, volume-averaged velocity
  void average_info(double *totalVolume, double *totalVel, double *averageVel) {
    if (!setup_done) setup();
    need_device();
    double h[4] = {0.0, 0.0, 0.0, 0.0};
    if (nlocal) {
      const int nb = std::min(256, cdiv(nlocal, 256));
      sum_partial.ensure(4 * 256); sum_out.ensure(4);
      k_particle_sum_partial<<<nb, 256, 0, stream>>>(nlocal, posr[cur].p, velm[cur].p, sum_partial.p);
      k_sum_final<<<1, 256, 0, stream>>>(sum_partial.p, nb, 4, sum_out.p);
      CK(cudaMemcpyAsync(h, sum_out.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      launches += 2;
    }
    comm.allreduce_sum_host(h, 4);
    if (totalVolume) *totalVolume = h[0];
    for (int d = 0; d < 3; d++) { if (totalVel) totalVel[d] = h[1 + d]; if (averageVel) averageVel[d] = h[1 + d] / (h[0] + 1.0e-150); }
  }

  void scatter_alpha_u(double *hgamma, double *hUe) {
    if (!setup_done) setup();
    if (!cell_valid) locate();
    const size_t C = ncells;
    k_scatter_alpha_u<<<cdiv(32 * C, 256), 256, 0, stream>>>(posr[cur].p, velm[cur].p, fc_start.p, fc_rows.p, (int)C, gamma.p, Ue.p);
    if (comm.nranks > 1) { comm.allreduce_sum_dev(gamma.p, C, stream); comm.allreduce_sum_dev(Ue.p, 3 * C, stream); }
    if (want_sums) field_sum(Ue.p, 0, sums + 6);   // Utotal1 = sum of Vp Up per cell, before the division by V (:936-941)
    if (smoothing_on(8) || smoothing_on(2)) {  // enhancedCloud.C:932-962 with alphaSmooth / UpSmooth
      k_alpha_u_divV<<<cdiv(C, 256), 256, 0, stream>>>((int)C, cellV.p, gamma.p, Ue.p);
      if (smoothing_on(8)) smooth_device_field(gamma.p, 1);
      if (smoothing_on(2)) smooth_device_field(Ue.p, 3);
      k_alpha_u_divgamma<<<cdiv(C, 256), 256, 0, stream>>>((int)C, gamma.p, Ue.p);
    } else {
      k_finalize_alpha_u<<<cdiv(C, 256), 256, 0, stream>>>((int)C, cellV.p, gamma.p, Ue.p);
    }
    launches += 2;
    if (want_sums) field_sum(Ue.p, 2, sums + 9);   // Utotal2 = sum Ue V gamma after smoothing (:966-970)
    if (hgamma) CK(cudaMemcpyAsync(hgamma, gamma.p, C * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (hUe) CK(cudaMemcpyAsync(hUe, Ue.p, 3 * C * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }

  void calc_tc(double *hAsrc, double *hOmega) {
    if (!setup_done) setup();
    if (!cell_valid) locate();
    const size_t C = ncells;
    k_scatter_asrc<<<cdiv(32 * C, 256), 256, 0, stream>>>(posr[cur].p, velm[cur].p, fc_start.p, fc_rows.p, (int)C, Uf.p, gamma.p, cellV.p, drag_model, nub, rhob, Asrc.p);
    if (comm.nranks > 1) comm.allreduce_sum_dev(Asrc.p, 3 * C, stream);
    if (want_sums) field_sum(Asrc.p, 1, sums + 0);   // Ftotal1 = sum Asrc V (1 - gamma) before smoothing (:395-403)
    if (smoothing_on(4)) {  // Asrc (1-gamma) -> smooth -> / (1-gamma)   (enhancedCloud.C:407-416, dragSmooth)
      k_scale_one_minus_gamma<<<cdiv(C, 256), 256, 0, stream>>>((int)C, gamma.p, Asrc.p, 3, 0);
      smooth_device_field(Asrc.p, 3);
      k_scale_one_minus_gamma<<<cdiv(C, 256), 256, 0, stream>>>((int)C, gamma.p, Asrc.p, 3, 1);
    } else {
      k_finalize_asrc<<<cdiv(C, 256), 256, 0, stream>>>((int)C, gamma.p, Asrc.p);
    }
    launches += 2;
    if (want_sums) field_sum(Asrc.p, 1, sums + 3);   // Ftotal2 after smoothing (:421-429)
    if (hAsrc) CK(cudaMemcpyAsync(hAsrc, Asrc.p, 3 * C * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (hOmega) memset(hOmega, 0, C * sizeof(double));  // enhancedCloud.C:391
  }

  void get_coupling_diag(int *hcell, double *Uri, double *mag, double *al, double *Jd, double *F) {
    need_device();
    const int m = nlocal;
    if (!m) return;
    if (hcell) CK(cudaMemcpyAsync(hcell, cell.p, m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    if (want_diag && dg_Uri.p) {
      if (Uri) CK(cudaMemcpyAsync(Uri, dg_Uri.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
      if (mag) CK(cudaMemcpyAsync(mag, dg_mag.p, m * sizeof(double), cudaMemcpyDeviceToHost, stream));
      if (al) CK(cudaMemcpyAsync(al, dg_alpha.p, m * sizeof(double), cudaMemcpyDeviceToHost, stream));
      if (Jd) CK(cudaMemcpyAsync(Jd, dg_Jd.p, m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    }
    if (F) {
      Buf<double> d;
      d.ensure(3 * (size_t)m);
      k_interleave3<<<cdiv(m, 256), 256, 0, stream>>>(fdrag[0].get(), fdrag[1].get(), fdrag[2].get(), m, d.p);
      CK(cudaMemcpyAsync(F, d.p, 3 * (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      d.release();
    }
    CK(cudaStreamSynchronize(stream));
  }

  // UOld = U before the particles are advanced (softParticleCloud.C:571-572); only the added-mass force reads it
  void save_uold_if_ready() {
    if (!loaded || !setup_done || !(force_flags & (SEDI_FORCE_ADDEDMASS | SEDI_FORCE_HISTORY)) || !nlocal) return;
    need_device();
    k_save_uold<<<cdiv(nlocal, 256), 256, 0, stream>>>(velm[cur].p, nlocal, uold[0].get(), uold[1].get(), uold[2].get());
    launches++;
  }

  // ---- particle injection / deletion (library.cpp:406-621).  Host round trip: the device state is
  // folded back into the script's atom table, edited there, and uploaded again; the next run re-does setup.  Contact
  // and wall history of surviving particles restart from zero (the reference keeps them) -- documented in DESIGN.md.
  // Everything a surviving particle owns besides its state quads, by device row (= order of script.atoms after
  // sync_host_atoms): per-atom fix arrays (fix_fluid_drag.cpp:211-224, fix_wall_granFix.cpp:726-732), stored force and
  // torque, wall-touch bits, Foam rank, history-force state, and the contact history as (partner tag, shear) lists --
  // what LAMMPS' AtomVec::copy + Fix::copy_arrays keep when library.cpp:406-621 edits the atom table.
  struct RowCarry {
    int n, nplanes;
    std::vector<std::vector<double> > planes;   // fdrag, dudt, vold, uold, f, tq (18), wall shear (3 per wall), history force (4)
    std::vector<unsigned> wmask;
    std::vector<int> foam;
    std::vector<int> nh, htag;                  // contact history: count and partner tags, MIG_MAXH per row
    std::vector<double> hshear;                 // 3 per entry
    RowCarry() : n(0), nplanes(0) {}
  };
  std::vector<double *> carry_plane_ptrs() {
    std::vector<double *> p;
    Plane2 *groups[] = {fdrag, dudt, vold, uold, f, tq};
    for (int g = 0; g < 6; g++) for (int d = 0; d < 3; d++) p.push_back(groups[g][d].get());
    for (int w = 0; w < cfg().nwalls; w++) for (int d = 0; d < 3; d++) p.push_back(wshear[w][d].get());
    if (hist_alloc) for (int d = 0; d < 4; d++) p.push_back(hist[d].get());
    return p;
  }
  bool carry_enabled() const { return setup_done && !getenv("SEDI_INJECT_RESET"); }
  void save_rows(RowCarry &R) {
    const int m = nlocal;
    R.n = m;
    std::vector<double *> ptrs = carry_plane_ptrs();
    R.nplanes = (int)ptrs.size();
    R.planes.assign(ptrs.size(), std::vector<double>((size_t)m));
    for (size_t k = 0; k < ptrs.size(); k++) if (m) CK(cudaMemcpyAsync(R.planes[k].data(), ptrs[k], (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    R.wmask.assign(m, 0u); R.foam.assign(m, 0);
    if (m) {
      CK(cudaMemcpyAsync(R.wmask.data(), wmask[icur].p, (size_t)m * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(R.foam.data(), foam[icur].p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaStreamSynchronize(stream));
    // contact history by tag (directed rows, as stored)
    R.nh.assign(m, 0); R.htag.assign((size_t)m * MIG_MAXH, 0); R.hshear.assign((size_t)m * MIG_MAXH * 3, 0.0);
    const long long tot = get_pairs(0, 0, 0, 0, 0, 0);
    if (tot > 0) {
      std::vector<int> ti(tot), tj(tot), touch(tot);
      std::vector<unsigned> meta(tot);
      std::vector<double> sh(3 * (size_t)tot);
      get_pairs(ti.data(), tj.data(), meta.data(), touch.data(), sh.data(), tot);
      std::vector<int> tg(m);
      get_state(0, 0, 0, 0, 0, 0, 0, tg.data(), 0, 0);
      std::vector<int> row_of(maxtag + 2, -1);
      for (int i = 0; i < m; i++) if (tg[i] >= 0 && tg[i] <= maxtag) row_of[tg[i]] = i;
      for (long long e = 0; e < tot; e++) {
        if (!touch[e]) continue;
        const int i = row_of[ti[e]];
        if (i < 0) continue;
        if (R.nh[i] >= MIG_MAXH) fatal("more than 16 touching partners on one particle: contact history cannot be carried across injection / deletion");
        const size_t s = (size_t)i * MIG_MAXH + R.nh[i]++;
        R.htag[s] = tj[e];
        for (int d = 0; d < 3; d++) R.hshear[3 * s + d] = sh[3 * (size_t)e + d];
      }
    }
  }
  // keep[t] = old row (index into R) of the t-th atom of the new table, -1 for an injected particle.  The device rows are
  // the atoms this rank kept at the upload (loaded_index: row -> table index).
  void restore_rows(const RowCarry &R, const std::vector<int> &keep) {
    const int m = nlocal;
    if (!m) return;
    if ((int)loaded_index.size() != m) fatal("internal: restore_rows without an upload map");
    if (R.nplanes > 18 + 3 * cfg().nwalls && !hist_alloc) {  // history-force planes existed before the re-upload
      for (int d = 0; d < 4; d++) for (int b = 0; b < 2; b++) { hist[d].b[b].ensure(npad); CK(cudaMemsetAsync(hist[d].b[b].p, 0, (size_t)npad * sizeof(double), stream)); hist[d].cur = 0; }
      hist_alloc = true;
    }
    std::vector<int> old(m);
    for (int i = 0; i < m; i++) old[i] = keep[loaded_index[i]];
    std::vector<double *> ptrs = carry_plane_ptrs();
    std::vector<double> tmp((size_t)m);
    for (size_t k = 0; k < ptrs.size() && k < R.planes.size(); k++) {
      for (int i = 0; i < m; i++) tmp[i] = old[i] >= 0 ? R.planes[k][old[i]] : 0.0;
      CK(cudaMemcpy(ptrs[k], tmp.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
    }
    std::vector<unsigned> wm(m); std::vector<int> fo(m), nh(m), ht((size_t)m * MIG_MAXH, 0);
    std::vector<D4> hs((size_t)m * MIG_MAXH);
    memset(hs.data(), 0, hs.size() * sizeof(D4));
    for (int i = 0; i < m; i++) {
      const int o = old[i];
      wm[i] = o >= 0 ? R.wmask[o] : 0u; fo[i] = o >= 0 ? R.foam[o] : 0; nh[i] = o >= 0 ? R.nh[o] : 0;
      for (int q = 0; q < nh[i]; q++) {
        const size_t so = (size_t)o * MIG_MAXH + q, sn = (size_t)i * MIG_MAXH + q;
        ht[sn] = R.htag[so];
        hs[sn].x = R.hshear[3 * so]; hs[sn].y = R.hshear[3 * so + 1]; hs[sn].z = R.hshear[3 * so + 2];
      }
    }
    CK(cudaMemcpy(wmask[icur].p, wm.data(), (size_t)m * sizeof(unsigned), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(foam[icur].p, fo.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice));
    inj_nh.ensure((size_t)m); inj_tag.ensure((size_t)m * MIG_MAXH); inj_shear.ensure((size_t)m * MIG_MAXH);
    CK(cudaMemcpy(inj_nh.p, nh.data(), (size_t)m * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(inj_tag.p, ht.data(), ht.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(inj_shear.p, hs.data(), hs.size() * sizeof(D4), cudaMemcpyHostToDevice));
  }
  // re-upload the edited atom table and bring the survivors' state back.  The reference edits the table in place and
  // forces a re-neighbouring at the next step (library.cpp:486-490, :606-611); here the table is rebuilt at once.
  void reinject(const RowCarry &R, const std::vector<int> &keep) {
    const long long step = cfg().ntimestep;
    load_atoms();
    restore_rows(R, keep);
    inject_pending = true;
    // bins + list with the carried contact history.  No force evaluation: the reference does not run Verlet::setup
    // here (`run N pre no`), the next half-kick uses the stored f(n) -- which does not know the injected particles
    // yet -- and an extra evaluation would apply the Coulomb rescale of the shear springs once more
    // (pair_gran_hertzFix_history.cpp:244-252 is not guarded by shearupdate).
    const bool had_setup = setup_done_once;
    const double dt0 = dt_init, r0 = lub_R0, rt0 = lub_RT0, rs0 = lub_RS0;
    setup(false);
    // the reference never re-initialises at an edit (library.cpp:406-621): init-time constants keep their values
    if (had_setup) { dt_init = dt0; lub_R0 = r0; lub_RT0 = rt0; lub_RS0 = rs0; build_base_params(); }
    inject_pending = false;
    cfg().ntimestep = step;
    CK(cudaStreamSynchronize(stream));
  }

  // ---- checkpoint / resume ---------------------------------------------------------------------------------------
  // The reference restarts its LAMMPS side with LAMMPS' own `restart` / `read_restart` (commented out in the shipped
  // inputs, e.g. multiParticlesCollideDia/in.lammps:37); the per-atom state that must survive is what
  // FixWallGranFix::pack_restart (fix_wall_granFix.cpp:750-777), FixFluidDrag's per-atom arrays and FixShearHistory
  // carry.  LAMMPS' binary layout is not reproduced: the file is this library's own (magic "SEDIRST1", little endian):
  // header, atom table, the RowCarry planes, wall-touch bits, Foam rank, contact history as (partner tag, shear) lists.
  // Resume is exact: no force evaluation at the restart (stored f(n), torque(n) are used by the next half-kick).
  struct RowCarryBox { RowCarry R; };
  static void wr(FILE *fp, const void *p, size_t n) { if (n && fwrite(p, 1, n, fp) != n) fatal("write_restart: short write"); }
  static void rd(FILE *fp, void *p, size_t n) { if (n && fread(p, 1, n, fp) != n) fatal("read_restart: file is truncated"); }
  std::string resolve_path(const std::string &path) const {
    const char *dir = getenv("SEDI_DUMP_DIR");
    if (dir && path.size() && path[0] != '/') return std::string(dir) + "/" + path;
    return path;
  }
  // Several GPUs: every rank writes its own rows to "<file>.<rank>" (same layout); read_restart reads all parts on every
  // rank and each brick keeps what it owns, so the number of GPUs may change between write and read.
  void write_restart(const std::string &path_in) {
    std::string path = path_in;
    if (comm.nranks > 1) { char suf[32]; snprintf(suf, sizeof(suf), ".%d", comm.rank); path += suf; }
    if (!setup_done) setup();
    RowCarry R;
    save_rows(R);
    sync_host_atoms();
    const AtomData &a = script.atoms;
    const SimConfig &c = cfg();
    FILE *fp = fopen(resolve_path(path).c_str(), "wb");
    if (!fp) fatal("Cannot open restart file", path.c_str());
    const char magic[8] = {'S', 'E', 'D', 'I', 'R', 'S', 'T', '1'};
    wr(fp, magic, 8);
    const long long step = c.ntimestep; wr(fp, &step, 8);
    wr(fp, &c.dt, 8); wr(fp, &dt_init, 8);
    wr(fp, &c.ntypes, 4); wr(fp, c.periodic, 12); wr(fp, c.boxlo, 24); wr(fp, c.boxhi, 24);
    const int n = (int)a.size(), nw = c.nwalls, hh = (hist_alloc ? 1 : 0) | (comm.nranks << 8);   // bits 8.. : number of parts
    wr(fp, &n, 4); wr(fp, &nw, 4); wr(fp, &hh, 4); wr(fp, &time_index, 4);
    for (size_t k = 0; k < c.fixes.size(); k++) if (c.fixes[k].kind == FIX_WALL_GRAN) {
      const int len = (int)strlen(c.fixes[k].id), wi = c.fixes[k].wall_index;
      wr(fp, &wi, 4); wr(fp, &len, 4); wr(fp, c.fixes[k].id, len);
    }
    wr(fp, a.tag.data(), 4 * (size_t)n); wr(fp, a.type.data(), 4 * (size_t)n); wr(fp, script.mask.data(), 4 * (size_t)n);
    wr(fp, a.x.data(), 24 * (size_t)n); wr(fp, a.v.data(), 24 * (size_t)n); wr(fp, a.omega.data(), 24 * (size_t)n);
    wr(fp, a.radius.data(), 8 * (size_t)n); wr(fp, a.rmass.data(), 8 * (size_t)n);
    const int np = R.nplanes; wr(fp, &np, 4);
    for (int k = 0; k < np; k++) wr(fp, R.planes[k].data(), 8 * (size_t)n);
    wr(fp, R.wmask.data(), 4 * (size_t)n); wr(fp, R.foam.data(), 4 * (size_t)n);
    wr(fp, R.nh.data(), 4 * (size_t)n); wr(fp, R.htag.data(), 4 * (size_t)n * MIG_MAXH); wr(fp, R.hshear.data(), 24 * (size_t)n * MIG_MAXH);
    fclose(fp);
  }
  void read_restart(const std::string &path) {
    // one file, or the parts "<file>.0" ... "<file>.<P-1>" a multi-GPU run wrote (P is in the header of every part)
    std::string first = resolve_path(path);
    FILE *probe = fopen(first.c_str(), "rb");
    bool parts = false;
    if (!probe) { first = resolve_path(path) + ".0"; probe = fopen(first.c_str(), "rb"); parts = true; }
    if (!probe) fatal("Cannot open restart file", path.c_str());
    fclose(probe);
    SimConfig &c = cfg();
    AtomData &a = script.atoms;
    a = AtomData();
    script.mask.clear();
    if (!restart_carry) restart_carry = new RowCarryBox();
    RowCarry &R = restart_carry->R;
    R = RowCarry();
    int nparts = 1, np_all = -1;
    for (int part = 0; part < nparts; part++) {
      std::string fn = resolve_path(path);
      if (parts) { char suf[32]; snprintf(suf, sizeof(suf), ".%d", part); fn += suf; }
      FILE *fp = fopen(fn.c_str(), "rb");
      if (!fp) fatal("Cannot open restart file", fn.c_str());
      char magic[8]; rd(fp, magic, 8);
      if (memcmp(magic, "SEDIRST1", 8)) fatal("read_restart: not a libsedi_b200 restart file", fn.c_str());
      long long step; rd(fp, &step, 8); c.ntimestep = step;
      double dt_file, dt0; rd(fp, &dt_file, 8); rd(fp, &dt0, 8); c.dt = dt_file;
      rd(fp, &c.ntypes, 4); rd(fp, c.periodic, 12); rd(fp, c.boxlo, 24); rd(fp, c.boxhi, 24); c.have_box = 1;
      for (int d = 0; d < 3; d++) c.boundary_str[d] = c.periodic[d] ? "pp" : "ff";
      int n, nw, hh, tix; rd(fp, &n, 4); rd(fp, &nw, 4); rd(fp, &hh, 4); rd(fp, &tix, 4);
      if (n < 0 || nw < 0 || nw > MAX_WALLS) fatal("read_restart: corrupt header");
      if (parts) nparts = std::max(1, hh >> 8);
      hh &= 1;
      time_index = tix;
      restart_wall_ids.assign(nw, std::string());
      for (int k = 0; k < nw; k++) {
        int wi, len; rd(fp, &wi, 4); rd(fp, &len, 4);
        if (wi < 0 || wi >= nw || len < 0 || len > 4096) fatal("read_restart: corrupt wall table");
        std::string id(len, ' '); rd(fp, &id[0], len); restart_wall_ids[wi] = id;
      }
      const size_t n0 = a.tag.size(), nn = n0 + (size_t)n;
      a.tag.resize(nn); a.type.resize(nn); script.mask.resize(nn); a.x.resize(3 * nn); a.v.resize(3 * nn); a.omega.resize(3 * nn);
      a.radius.resize(nn); a.rmass.resize(nn);
      rd(fp, a.tag.data() + n0, 4 * (size_t)n); rd(fp, a.type.data() + n0, 4 * (size_t)n); rd(fp, script.mask.data() + n0, 4 * (size_t)n);
      rd(fp, a.x.data() + 3 * n0, 24 * (size_t)n); rd(fp, a.v.data() + 3 * n0, 24 * (size_t)n); rd(fp, a.omega.data() + 3 * n0, 24 * (size_t)n);
      rd(fp, a.radius.data() + n0, 8 * (size_t)n); rd(fp, a.rmass.data() + n0, 8 * (size_t)n);
      int np; rd(fp, &np, 4);
      if (np != 18 + 3 * nw + 4 * hh) fatal("read_restart: corrupt plane table");
      if (np_all < 0) { np_all = np; R.nplanes = np; R.planes.assign(np, std::vector<double>()); }
      if (np != np_all) fatal("read_restart: the parts disagree on the per-atom planes");
      for (int k = 0; k < np; k++) { R.planes[k].resize(nn); rd(fp, R.planes[k].data() + n0, 8 * (size_t)n); }
      R.wmask.resize(nn); R.foam.resize(nn); R.nh.resize(nn); R.htag.resize(nn * MIG_MAXH); R.hshear.resize(nn * MIG_MAXH * 3);
      rd(fp, R.wmask.data() + n0, 4 * (size_t)n); rd(fp, R.foam.data() + n0, 4 * (size_t)n);
      rd(fp, R.nh.data() + n0, 4 * (size_t)n); rd(fp, R.htag.data() + n0 * MIG_MAXH, 4 * (size_t)n * MIG_MAXH);
      rd(fp, R.hshear.data() + n0 * MIG_MAXH * 3, 24 * (size_t)n * MIG_MAXH);
      fclose(fp);
      R.n = (int)nn;
    }
    keep_local.clear();   // every brick keeps the atoms it owns
    loaded = false; setup_done = false; restart_pending = true;
  }
  // first setup after read_restart: the fixes are known now, so the stored per-atom state can be matched to them
  // (LAMMPS matches restart_peratom data by fix id; wall history is matched by the wall fix's id here)
  void setup_from_restart() {
    restart_pending = false;
    RowCarry &F = restart_carry->R;
    const SimConfig &c = cfg();
    const int n = F.n, nw_file = (int)restart_wall_ids.size(), hh = (F.nplanes - 18 - 3 * nw_file) / 4;
    RowCarry R;
    R.n = n;
    R.wmask.assign(n, 0u); R.foam = F.foam; R.nh = F.nh; R.htag = F.htag; R.hshear = F.hshear;
    for (int k = 0; k < 18; k++) R.planes.push_back(F.planes[k]);
    std::vector<int> file_wall(c.nwalls, -1);
    for (size_t k = 0; k < c.fixes.size(); k++) if (c.fixes[k].kind == FIX_WALL_GRAN)
      for (int w = 0; w < nw_file; w++) if (restart_wall_ids[w] == std::string(c.fixes[k].id)) file_wall[c.fixes[k].wall_index] = w;
    for (int w = 0; w < c.nwalls; w++) for (int d = 0; d < 3; d++)
      R.planes.push_back(file_wall[w] >= 0 ? F.planes[18 + 3 * file_wall[w] + d] : std::vector<double>((size_t)n, 0.0));
    for (int i = 0; i < n; i++) for (int w = 0; w < c.nwalls; w++) if (file_wall[w] >= 0 && ((F.wmask[i] >> file_wall[w]) & 1u)) R.wmask[i] |= (1u << w);
    if (hh) for (int d = 0; d < 4; d++) R.planes.push_back(F.planes[18 + 3 * nw_file + d]);
    R.nplanes = (int)R.planes.size();
    std::vector<int> keep(n);
    for (int i = 0; i < n; i++) keep[i] = i;
    hist_alloc = false;
    reinject(R, keep);
    delete restart_carry; restart_carry = 0;
  }

  // ---- OpenFOAM lagrangian fields (softParticle::writeFields, softParticleIO.C:157-197): positions (with the owner cell),
  // d, tag, lmpCpuId, type, U, ensembleU in OpenFOAM's ASCII IOField layout under <dir>/ (the host passes
  // "<time>/lagrangian/<cloudName>").  density and n0 are written too: readFields (:113-152) demands them although the
  // reference's writeFields forgets to write them.  Rows in ascending tag; particles outside the mesh keep cell -1.
  void write_lagrangian(const char *dir, const char *location) {
    if (!setup_done) setup();
    if (have_mesh && !cell_valid) locate();
    const int m = nlocal;
    std::vector<double> x(3 * (size_t)m), v(3 * (size_t)m), r(m), ms(m), n0(m, 0.0);
    std::vector<int> tg(m), ty(m), cl(m, -1);
    if (m) {
      get_state(x.data(), v.data(), 0, 0, 0, r.data(), ms.data(), tg.data(), ty.data(), 0);
      if (have_mesh) { CK(cudaMemcpyAsync(cl.data(), cell.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, stream)); CK(cudaStreamSynchronize(stream)); }
      if (hist_alloc) get_history_state(0, n0.data());
    }
    std::vector<int> ord(m);
    for (int i = 0; i < m; i++) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) { return tg[a] < tg[b]; });
    const std::string base = std::string(dir) + (comm.nranks > 1 ? "/processor" + std::to_string(comm.rank) : std::string());
    auto open_field = [&](const char *name, const char *cls) -> FILE * {
      const std::string path = std::string(dir) + "/" + name + (comm.nranks > 1 ? "." + std::to_string(comm.rank) : std::string());
      FILE *fp = fopen(path.c_str(), "w");
      if (!fp) fatal("Cannot open lagrangian field file", path.c_str());
      fprintf(fp, "/*--------------------------------*- C++ -*----------------------------------*\\\n"
                  "| =========                 |                                                 |\n"
                  "| \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox           |\n"
                  "\\*---------------------------------------------------------------------------*/\n"
                  "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       %s;\n    location    \"%s\";\n    object      %s;\n}\n"
                  "// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n%d\n(\n", cls, location, name, m);
      return fp;
    };
    auto close_field = [&](FILE *fp) { fprintf(fp, ")\n\n\n// ************************************************************************* //\n"); fclose(fp); };
    (void)base;
    FILE *fp = open_field("positions", "Cloud<softParticle>");
    for (int q = 0; q < m; q++) { const int i = ord[q]; fprintf(fp, "(%.15g %.15g %.15g) %d\n", x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2], cl[i]); }
    close_field(fp);
    fp = open_field("d", "scalarField");
    for (int q = 0; q < m; q++) fprintf(fp, "%.15g\n", 2.0 * r[ord[q]]);
    close_field(fp);
    fp = open_field("density", "scalarField");
    for (int q = 0; q < m; q++) { const int i = ord[q]; fprintf(fp, "%.15g\n", 3.0 * ms[i] / (4.0 * SEDI_PI_LIBRARY * r[i] * r[i] * r[i])); }   // library.cpp:200
    close_field(fp);
    fp = open_field("n0", "scalarField");
    for (int q = 0; q < m; q++) fprintf(fp, "%.15g\n", n0[ord[q]]);
    close_field(fp);
    fp = open_field("tag", "labelField");
    for (int q = 0; q < m; q++) fprintf(fp, "%d\n", tg[ord[q]]);
    close_field(fp);
    fp = open_field("lmpCpuId", "labelField");
    for (int q = 0; q < m; q++) fprintf(fp, "%d\n", comm.rank);
    close_field(fp);
    fp = open_field("type", "labelField");
    for (int q = 0; q < m; q++) fprintf(fp, "%d\n", ty[ord[q]]);
    close_field(fp);
    fp = open_field("U", "vectorField");
    for (int q = 0; q < m; q++) { const int i = ord[q]; fprintf(fp, "(%.15g %.15g %.15g)\n", v[3 * (size_t)i], v[3 * (size_t)i + 1], v[3 * (size_t)i + 2]); }
    close_field(fp);
    fp = open_field("ensembleU", "vectorField");   // never assigned by the reference: stays (0 0 0) (softParticle.C:56)
    for (int q = 0; q < m; q++) fprintf(fp, "(0 0 0)\n");
    close_field(fp);
  }

  void sync_host_atoms() {
    if (!loaded) return;
    if (!nlocal) { script.atoms = AtomData(); script.mask.clear(); return; }   // an empty brick owns no atoms
    const int m = nlocal;
    std::vector<double> x(3 * (size_t)m), v(3 * (size_t)m), w(3 * (size_t)m), r(m), ms(m);
    std::vector<int> tg(m), ty(m), mk(m);
    get_state(x.data(), v.data(), w.data(), 0, 0, r.data(), ms.data(), tg.data(), ty.data(), mk.data());
    AtomData &a = script.atoms;
    a.tag = tg; a.type = ty; a.x = x; a.v = v; a.omega = w; a.radius = r; a.rmass = ms;
    script.mask = mk;
  }
  void create_particles(int np, const double *pos, const double *tagd, double diameter, double rho, int type, const double *vel) {
    const bool carry = carry_enabled() && (comm.nranks > 1 || nlocal > 0);   // (collective on several GPUs: an empty brick takes part in the re-upload)
    RowCarry R;
    if (carry) save_rows(R);
    sync_host_atoms();
    std::vector<int> keep;
    for (size_t i = 0; i < script.atoms.size(); i++) keep.push_back((int)i);
    std::vector<char> kl(script.atoms.size(), carry ? 1 : 0);   // rows that live on this GPU stay here; new particles go to the brick that owns them
    const int active = cfg().find_group("active");  // library.cpp:447-450: mask = 1 | bit("active")
    for (int m = 0; m < np; m++) {
      script.add_atom((int)tagd[m], type, diameter, rho, pos + 3 * (size_t)m, vel);
      // the reference uses its truncated pi literal for the mass of injected particles (library.cpp:460)
      const double rad = 0.5 * diameter;
      script.atoms.rmass.back() = 4.0 * SEDI_PI_LIBRARY / 3.0 * rad * rad * rad * rho;
      script.mask.back() = 1 | active;
      keep.push_back(-1);
      kl.push_back(0);
    }
    loaded = false; setup_done = false;
    keep_local = kl;
    if (carry) reinject(R, keep);
  }
  void delete_particles(const int *list, int nd) {  // list holds atom tags (library.cpp:507-621)
    const bool carry = carry_enabled() && (comm.nranks > 1 || nlocal > 0);   // (collective on several GPUs: an empty brick takes part in the re-upload)
    RowCarry R;
    if (carry) save_rows(R);
    sync_host_atoms();
    std::vector<int> keep;
    std::vector<int> del(list, list + nd);
    std::sort(del.begin(), del.end());
    AtomData &a = script.atoms, b;
    std::vector<int> mk;
    for (size_t i = 0; i < a.size(); i++) {
      if (std::binary_search(del.begin(), del.end(), a.tag[i])) continue;
      b.tag.push_back(a.tag[i]); b.type.push_back(a.type[i]); b.radius.push_back(a.radius[i]); b.rmass.push_back(a.rmass[i]);
      for (int d = 0; d < 3; d++) { b.x.push_back(a.x[3 * i + d]); b.v.push_back(a.v[3 * i + d]); b.omega.push_back(a.omega[3 * i + d]); }
      mk.push_back(script.mask[i]);
      keep.push_back((int)i);
    }
    script.atoms = b; script.mask = mk;
    loaded = false; setup_done = false;
    keep_local.assign(b.size(), carry ? 1 : 0);
    if (carry) reinject(R, keep);
  }
};

}  // namespace sedi

#include "sedi_comm_impl.cuh"
namespace sedi { static void g_comm_engine_set(Engine *e) { g_comm_engine = e; } }

// =====================================================================================================================
// C-ABI
// =====================================================================================================================
using sedi::Engine;

// The handle that crosses the boundary is a LAMMPS_NS::LAMMPS* exactly as in the reference (softParticleCloud.H:71
// holds `LAMMPS* lmp_` and passes it as the `void *` of library.h): include/lammps_shim/lammps.h.
namespace LAMMPS_NS {
LAMMPS::LAMMPS(int, char **, MPI_Comm communicator) : input(new Input(this)), engine(new Engine()), world(communicator) {
  ((Engine *)engine)->mpi_world = communicator;
}
LAMMPS::~LAMMPS() { delete (Engine *)engine; delete input; }
char *Input::one(const char *line) { ((Engine *)lmp->engine)->command(line); return NULL; }
void Input::file(const char *path) { ((Engine *)lmp->engine)->file(path); }
}  // namespace LAMMPS_NS

static inline Engine *E(void *p) {
  if (!p) sedi::fatal("NULL LAMMPS handle passed to libsedi_b200");
  return (Engine *)((LAMMPS_NS::LAMMPS *)p)->engine;
}
// the pre-run queries of softParticleCloud::initLammps (softParticleCloud.C:119-163) are per-rank in the reference: on
// several ranks the bricks must exist (and own their atoms) before they are answered
static inline Engine *EQ(void *p) {
  Engine *e = E(p);
  e->auto_comm();
  if (e->comm.nranks > 1 && !e->loaded) e->load_atoms();
  return e;
}

extern "C" {

void lammps_open(int argc, char **argv, MPI_Comm comm, void **ptr) { *ptr = (void *)new LAMMPS_NS::LAMMPS(argc, argv, comm); }
void lammps_close(void *ptr) { delete (LAMMPS_NS::LAMMPS *)ptr; }
void lammps_file(void *ptr, char *path) { E(ptr)->file(path); }
char *lammps_command(void *ptr, char *line) { E(ptr)->command(line); return NULL; }
void lammps_sync(void *ptr) { Engine *e = E(ptr); if (e->dev_ready) { CK(cudaSetDevice(e->device)); CK(cudaStreamSynchronize(e->stream)); } e->comm.barrier(); }
int lammps_get_global_n(void *ptr) { Engine *e = EQ(ptr); long long m = e->loaded ? e->nlocal : (long long)e->script.atoms.size(); return (int)e->comm.allreduce_sum_ll(m); }
void lammps_get_initial_np(void *ptr, int *np) {
  Engine *e = EQ(ptr);
  const int m = e->loaded ? e->nlocal : (int)e->script.atoms.size();
  e->comm.allgather_int(m, np);
}
void lammps_get_initial_info(void *ptr, double *coords, double *velos, double *diam, double *rho, int *tag, int *lmpCpuId,
                             int *type) {
  Engine *e = EQ(ptr);
  if (!e->loaded) {  // before the first run the atoms still live in the script (read_data order)
    const sedi::AtomData &a = e->script.atoms;
    for (size_t i = 0; i < a.size(); i++) {
      for (int d = 0; d < 3; d++) { coords[3 * i + d] = a.x[3 * i + d]; velos[3 * i + d] = a.v[3 * i + d]; }
      const double r = a.radius[i];
      diam[i] = r * 2.0;
      rho[i] = 3.0 * a.rmass[i] / (4.0 * sedi::SEDI_PI_LIBRARY * r * r * r);  // library.cpp:200
      type[i] = a.type[i]; tag[i] = a.tag[i]; lmpCpuId[i] = e->comm.rank;
    }
    return;
  }
  const int m = e->nlocal;
  std::vector<double> r(m), ms(m);
  e->get_state(coords, velos, 0, 0, 0, r.data(), ms.data(), tag, type, 0);
  for (int i = 0; i < m; i++) {
    diam[i] = r[i] * 2.0;
    rho[i] = 3.0 * ms[i] / (4.0 * sedi::SEDI_PI_LIBRARY * r[i] * r[i] * r[i]);
    lmpCpuId[i] = e->comm.rank;
  }
}
int lammps_get_local_n(void *ptr) { Engine *e = EQ(ptr); return e->loaded ? e->nlocal : (int)e->script.atoms.size(); }
void lammps_get_local_domain(void *ptr, double *dom) {
  Engine *e = EQ(ptr);
  for (int d = 0; d < 3; d++) { dom[2 * d] = e->comm.sublo(e->cfg(), d, 0.0); dom[2 * d + 1] = e->comm.subhi(e->cfg(), d, 0.0); }
}
void lammps_get_local_info(void *ptr, double *coords, double *velos, int *foamCpuId, int *lmpCpuId, int *tag) {
  Engine *e = E(ptr);
  const double t0 = Engine::now_s();
  e->get_local(coords, velos, foamCpuId, lmpCpuId, tag);
  e->timers[8] += Engine::now_s() - t0;   // cpuTimeSplit[5]: lammps -> foam
}
void lammps_put_local_info(void *ptr, int nLocalIn, double *fdrag, double *DuDt, int *foamCpuIdIn, int *tagIn) {
  (void)DuDt;  // ignored by the reference as well (library.cpp:314-367 never reads it)
  Engine *e = E(ptr);
  const double t0 = Engine::now_s();
  e->put_local(nLocalIn, fdrag, foamCpuIdIn, tagIn);
  e->timers[6] += Engine::now_s() - t0;   // cpuTimeSplit[3]: foam -> lammps
}
void lammps_step(void *ptr, int n) {
  Engine *e = E(ptr);
  const double t0 = Engine::now_s();
  e->save_uold_if_ready(); e->run(n);
  e->timers[7] += Engine::now_s() - t0;   // cpuTimeSplit[4]: lammps
}
void lammps_set_timestep(void *ptr, double dt) { Engine *e = E(ptr); e->cfg().dt = dt; e->params_dirty = true; }
double lammps_get_timestep(void *ptr) { return E(ptr)->cfg().dt; }
void lammps_create_particle(void *ptr, int npAdd, double *position, double *tag, double diameter, double rho, int type,
                            double *vel) {
  E(ptr)->create_particles(npAdd, position, tag, diameter, rho, type, vel);
}
void lammps_delete_particle(void *ptr, int *deleteList, int nDelete) { E(ptr)->delete_particles(deleteList, nDelete); }

int sedi_abi_version(void) { return 1; }
/* what the script parser understood, as JSON (host only, no device needed): the parser check of tests/test_abi.py compares it with
 * hand-written expectations for every shipped in.lammps */
int sedi_config_json(void *ptr, char *buf, int cap) {
  const sedi::SimConfig &c = E(ptr)->cfg();
  std::string o = "{";
  char t[512];
  snprintf(t, sizeof(t), "\"periodic\": [%d, %d, %d], \"skin\": %.17g, \"dt\": %.17g, \"newton_pair\": %d, \"pair\": %d, \"ntypes\": %d, \"nwalls\": %d, "
           "\"freeze_group_bit\": %d, \"neigh_modify_seen\": %d, \"procgrid\": [%d, %d, %d], ", c.periodic[0], c.periodic[1], c.periodic[2], c.skin, c.dt, c.newton_pair, c.pair, c.ntypes, c.nwalls,
           c.freeze_group_bit, c.neigh_modify_seen, c.procgrid[0], c.procgrid[1], c.procgrid[2]);
  o += t;
  snprintf(t, sizeof(t), "\"gran\": {\"kn\": %.17g, \"kt\": %.17g, \"gamman\": %.17g, \"gammat\": %.17g, \"xmu\": %.17g, \"dampflag\": %d}, ", c.gran.kn, c.gran.kt,
           c.gran.gamman, c.gran.gammat, c.gran.xmu, c.gran.dampflag);
  o += t;
  snprintf(t, sizeof(t), "\"lub\": {\"enabled\": %d, \"mu\": %.17g, \"flaglog\": %d, \"flagfld\": %d, \"cut_inner\": %.17g, \"cut_global\": %.17g, \"flagHI\": %d, \"flagVF\": %d}, ",
           c.lub.enabled, c.lub.mu, c.lub.flaglog, c.lub.flagfld, c.lub.cut_inner, c.lub.cut_global, c.lub.flagHI, c.lub.flagVF);
  o += t;
  o += "\"groups\": {";
  for (size_t k = 0; k < c.groups.size(); k++) { snprintf(t, sizeof(t), "%s\"%s\": %d", k ? ", " : "", c.groups[k].name.c_str(), c.groups[k].bit); o += t; }
  o += "}, \"fixes\": [";
  for (size_t k = 0; k < c.fixes.size(); k++) {
    const sedi::FixSpec &f = c.fixes[k];
    snprintf(t, sizeof(t), "%s{\"id\": \"%s\", \"kind\": %d, \"groupbit\": %d, \"g\": %.17g, \"gdir\": [%.17g, %.17g, %.17g], \"carrier_rho\": %.17g, ", k ? ", " : "", f.id, f.kind,
             f.groupbit, f.g, f.gdir[0], f.gdir[1], f.gdir[2], f.carrier_rho);
    o += t;
    snprintf(t, sizeof(t), "\"ah\": %.17g, \"lam\": %.17g, \"smin\": %.17g, \"smax\": %.17g, \"opt\": %d, ", f.ah, f.lam, f.smin, f.smax, f.opt);
    o += t;
    snprintf(t, sizeof(t), "\"wall\": {\"kn\": %.17g, \"kt\": %.17g, \"gamman\": %.17g, \"gammat\": %.17g, \"xmu\": %.17g, \"dampflag\": %d}, \"wallstyle\": %d, \"lo\": %.17g, "
             "\"hi\": %.17g, \"cylradius\": %.17g, \"wiggle\": %d, \"wshear\": %d, \"axis\": %d, \"amplitude\": %.17g, \"period\": %.17g, \"vshear\": %.17g, \"wall_index\": %d}",
             f.wall.kn, f.wall.kt, f.wall.gamman, f.wall.gammat, f.wall.xmu, f.wall.dampflag, f.wallstyle, f.lo, f.hi, f.cylradius, f.wiggle, f.wshear, f.axis, f.amplitude,
             f.period, f.vshear, f.wall_index);
    o += t;
  }
  o += "], \"dumps\": [";
  for (size_t k = 0; k < c.dumps.size(); k++) {
    snprintf(t, sizeof(t), "%s{\"id\": \"%s\", \"every\": %lld, \"path\": \"%s\", \"columns\": \"%s\", \"groupbit\": %d}", k ? ", " : "", c.dumps[k].id.c_str(), c.dumps[k].every,
             c.dumps[k].path.c_str(), c.dumps[k].columns.c_str(), c.dumps[k].groupbit);
    o += t;
  }
  o += "]}";
  if ((int)o.size() + 1 > cap) return -(int)o.size() - 1;
  memcpy(buf, o.c_str(), o.size() + 1);
  return (int)o.size();
}
int sedi_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
  return cnt;
}
void sedi_set_device(void *ptr, int dev) { Engine *e = E(ptr); if (e->dev_ready) sedi::fatal("sedi_set_device must precede the first compute call"); e->device = dev; }
void sedi_set_box(void *ptr, const double *lo, const double *hi, int ntypes) {
  Engine *e = E(ptr);
  for (int d = 0; d < 3; d++) { e->cfg().boxlo[d] = lo[d]; e->cfg().boxhi[d] = hi[d]; }
  e->cfg().have_box = 1; e->cfg().ntypes = ntypes; e->params_dirty = true;
}
void sedi_add_atoms(void *ptr, int n, const int *tag, const int *type, const double *diameter, const double *density, const double *x,
                    const double *v) {
  Engine *e = E(ptr);
  for (int i = 0; i < n; i++) e->script.add_atom(tag[i], type[i], diameter[i], density[i], x + 3 * (size_t)i, v ? v + 3 * (size_t)i : 0);
  e->loaded = false; e->setup_done = false;
}
void sedi_set_omega(void *ptr, int n, const int *tag, const double *omega) { E(ptr)->set_omega(n, tag, omega); }
void sedi_get_state(void *ptr, double *x, double *v, double *omega, double *f, double *torque, double *radius, double *rmass, int *tag,
                    int *type, int *mask) {
  E(ptr)->get_state(x, v, omega, f, torque, radius, rmass, tag, type, mask);
}
long long sedi_get_pairs(void *ptr, int *tag_i, int *tag_j, unsigned *meta, int *touch, double *shear, long long cap) {
  return E(ptr)->get_pairs(tag_i, tag_j, meta, touch, shear, cap);
}
void sedi_get_wall_shear(void *ptr, int wall, double *shear) { E(ptr)->get_wall_shear(wall, shear); }
void sedi_get_row_stats(void *ptr, int *entries, int *touching) { E(ptr)->get_row_stats(entries, touching); }
void sedi_force_rebuild(void *ptr) { Engine *e = E(ptr); if (!e->setup_done) e->setup(); else e->rebuild(); }
long long sedi_get_stat(void *ptr, int which) {
  Engine *e = E(ptr);
  switch (which) {
    case 0: return e->nbuilds;
    case 1: return e->pair_evals;
    case 2: return e->steps_done;
    case 3: return e->list_gran_dir;
    case 4: return e->list_type_dir;
    case 5: return e->ell[e->ecur].cap + e->ell[e->ecur].tcap;
    case 6: return e->launches;
    case 7: return e->nlocal;
    case 8: return e->list_pairs_undirected();
    case 9: return e->n - e->nlocal;
    case 10: return e->pair_evals_unique / 2;
    default: return -1;
  }
}
void sedi_reset_stats(void *ptr) { Engine *e = E(ptr); e->nbuilds = e->pair_evals = e->steps_done = e->launches = e->pair_evals_unique = 0; }
void sedi_synchronize(void *ptr) { Engine *e = E(ptr); if (e->dev_ready) { CK(cudaSetDevice(e->device)); CK(cudaStreamSynchronize(e->stream)); } }
void *sedi_stream(void *ptr) { Engine *e = E(ptr); e->need_device(); return (void *)e->stream; }
double sedi_last_step_ms(void *ptr) { return E(ptr)->last_step_ms; }
void sedi_timer_start(void *ptr) { Engine *e = E(ptr); e->need_device(); CK(cudaEventRecord(e->evt0, e->stream)); }
double sedi_timer_stop_ms(void *ptr) {
  Engine *e = E(ptr);
  e->need_device();
  CK(cudaEventRecord(e->evt1, e->stream));
  CK(cudaEventSynchronize(e->evt1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e->evt0, e->evt1));
  return ms;
}
void sedi_profile(void *ptr, int on) { Engine *e = E(ptr); e->prof_on = (on != 0); e->prof_ms = 0; e->prof_steps = 0; }
long long sedi_get_profile(void *ptr, double *kernel_ms) { Engine *e = E(ptr); if (kernel_ms) *kernel_ms = e->prof_ms; return e->prof_steps; }

void sedi_mesh_box(void *ptr, const double *lo, const double *hi, const int *ncell) { E(ptr)->mesh_box(lo, hi, ncell); }
void sedi_mesh_rectilinear(void *ptr, const int *ncell, const double *xfaces, const double *yfaces, const double *zfaces, const int *cell_label) {
  E(ptr)->mesh_rectilinear(ncell, xfaces, yfaces, zfaces, cell_label);
}
int sedi_mesh_ncells(void *ptr) { return E(ptr)->ncells; }
void sedi_coupling_config(void *ptr, int drag_model, int force_flags, double nub, double rhob, const double *g, double deltaT) {
  Engine *e = E(ptr);
  if (drag_model != SEDI_DRAG_ERGUN_WENYU_ID && drag_model != SEDI_DRAG_SYAMLAL_OBRIEN_ID) sedi::fatal("Unknown dragModel id");
  e->drag_model = drag_model; e->force_flags = force_flags; e->nub = nub; e->rhob = rhob; e->deltaT = deltaT;
  for (int d = 0; d < 3; d++) e->gvec[d] = g ? g[d] : 0.0;
}
void sedi_coupling_time_index(void *ptr, int time_index) { E(ptr)->time_index = time_index; }
void sedi_coupling_inlet(void *ptr, const double *inlet_force, const double *inlet_box, int region_option, const double *eccentricity) {
  Engine *e = E(ptr);
  for (int d = 0; d < 3; d++) { e->inlet_force[d] = inlet_force ? inlet_force[d] : 0.0; e->inlet_ecc[d] = eccentricity ? eccentricity[d] : 0.0; }
  for (int d = 0; d < 9; d++) e->inlet_box[d] = inlet_box ? inlet_box[d] : 0.0;
  e->inlet_option = region_option;
}
void sedi_get_history_state(void *ptr, double *sumDeltaFb, double *n0) { E(ptr)->get_history_state(sumDeltaFb, n0); }
void sedi_put_cell_fields(void *ptr, const double *Uf, const double *gamma, const double *gradp, const double *DDtU, const double *curlU) {
  E(ptr)->put_cell_fields(Uf, gamma, gradp, DDtU, curlU);
}
void sedi_locate(void *ptr) { Engine *e = E(ptr); if (!e->setup_done) e->setup(); e->locate(); }
void sedi_compute_fluid_force(void *ptr) { E(ptr)->fluid_force(); }
void sedi_scatter_alpha_u(void *ptr, double *gamma, double *Ue) { E(ptr)->scatter_alpha_u(gamma, Ue); }
void sedi_calc_tc(void *ptr, double *Asrc, double *Omega) { E(ptr)->calc_tc(Asrc, Omega); }
void sedi_smooth_config(void *ptr, double bandwidth, int steps, const double *Ddiag, int flags) {
  Engine *e = E(ptr);
  e->smooth_b = bandwidth; e->smooth_steps = steps; e->smooth_flags = flags;
  for (int d = 0; d < 3; d++) e->smooth_D[d] = Ddiag ? Ddiag[d] : 1.0;
}
void sedi_smooth_uf(void *ptr) { E(ptr)->smooth_uf(); }
void sedi_smooth_field(void *ptr, double *field, int ncomp) { E(ptr)->smooth_host_field(field, ncomp); }
int sedi_smooth_last_iters(void *ptr) { return E(ptr)->smooth_iters_last; }
void sedi_enable_diag(void *ptr, int on) { E(ptr)->want_diag = (on != 0); }
void sedi_get_coupling_diag(void *ptr, int *cell, double *Uri, double *magUri, double *alphap, double *Jd, double *F) {
  E(ptr)->get_coupling_diag(cell, Uri, magUri, alphap, Jd, F);
}
void sedi_step(void *ptr, int n) {
  Engine *e = E(ptr);
  const double t0 = Engine::now_s();
  e->save_uold_if_ready(); e->run(n);
  e->timers[7] += Engine::now_s() - t0;
}
void sedi_write_lagrangian(void *ptr, const char *dir, const char *location) { E(ptr)->write_lagrangian(dir, location ? location : ""); }
void sedi_enable_conservation_sums(void *ptr, int on) { E(ptr)->want_sums = (on != 0); }
void sedi_get_conservation_sums(void *ptr, double *Ftotal_before, double *Ftotal_after, double *Utotal_before, double *Utotal_after) {
  Engine *e = E(ptr);
  for (int d = 0; d < 3; d++) {
    if (Ftotal_before) Ftotal_before[d] = e->sums[d];
    if (Ftotal_after) Ftotal_after[d] = e->sums[3 + d];
    if (Utotal_before) Utotal_before[d] = e->sums[6 + d];
    if (Utotal_after) Utotal_after[d] = e->sums[9 + d];
  }
}
void sedi_average_info(void *ptr, double *totalVolume, double *totalVel, double *averageVel) { E(ptr)->average_info(totalVolume, totalVel, averageVel); }
void sedi_get_timers(void *ptr, double *diffusionTimeCount, double *particleMoveTime, double *cpuTimeSplit) {
  Engine *e = E(ptr);
  if (diffusionTimeCount) { diffusionTimeCount[0] = e->timers[0]; diffusionTimeCount[1] = e->timers[1]; }
  if (particleMoveTime) *particleMoveTime = e->timers[2];
  if (cpuTimeSplit) for (int k = 0; k < 6; k++) cpuTimeSplit[k] = e->timers[3 + k];
}

int sedi_comm_init(void *ptr, int rank, int nranks, const void *nccl_unique_id, int id_bytes, const int *procgrid) {
  return E(ptr)->comm.init(*E(ptr), rank, nranks, nccl_unique_id, id_bytes, procgrid);
}
int sedi_comm_unique_id(void *out, int cap) { return sedi::Comm::unique_id(out, cap); }
int sedi_comm_rank(void *ptr) { return E(ptr)->comm.rank; }
long long sedi_comm_stat(void *ptr, int which) {
  Engine *e = E(ptr);
  switch (which) { case 0: return e->comm.halo_calls; case 1: return e->comm.total_send; case 2: return e->comm.total_recv; case 3: return (long long)e->comm.links.size(); case 4: return e->comm.narr_last; case 5: return e->comm.p2p ? 1 : 0; default: return -1; }
}
/* pure host logic of the brick decomposition (no GPU, no NCCL): used by the CPU tests */
void sedi_decomp_grid(int nranks, const double *boxlen, int *grid) { sedi::decomp_auto_grid(nranks, boxlen, grid); }
int sedi_decomp_owner(const double *x, const double *boxlo, const double *boxhi, const int *grid) { return sedi::decomp_owner(x, boxlo, boxhi, grid); }
int sedi_decomp_links(int rank, const int *grid, const int *periodic, const double *prd, int *peers, int *offsets, double *shifts) {
  std::vector<sedi::LinkHost> l = sedi::decomp_links(rank, grid, periodic, prd);
  for (size_t k = 0; k < l.size(); k++) { peers[k] = l[k].peer; for (int d = 0; d < 3; d++) { offsets[3 * k + d] = l[k].off[d]; shifts[3 * k + d] = l[k].shift[d]; } }
  return (int)l.size();
}

}  // extern "C"
