/* sedi_b200.h -- C-ABI of libsedi_b200.so, the B200-native particle hot path of sediFoam.
 *
 * Two groups of entry points, all `extern "C"`, plain pointers and sizes only:
 *
 * (1) The drop-in boundary: the 17 `lammps_*` functions of the reference's own C interface between the OpenFOAM
 *     side and LAMMPS (/root/reference/interfaceToLammps/library.h:29-63).  Same names, same argument meaning,
 *     caller-allocated arrays, identity across the boundary = atom tag.  `softParticleCloud`
 *     (lammpsFoam/softParticleCloud.C:119-163, :838-922) links against these unchanged.  MPI_Comm is an `int`
 *     here when no MPI is present (define SEDI_HAVE_MPI before including to use <mpi.h>).
 *
 * (2) `sedi_*`: the device-resident API used by the host-side mirror of `enhancedCloud` (host/) so that per-particle
 *     data never crosses PCIe: fluid cell fields in, Eulerian cell fields out.
 *
 * Errors follow the reference convention (library.cpp:380-383): print to stderr and abort().  There is no CPU
 * fallback: every compute entry point requires a CUDA device and aborts without one.
 */
#ifndef SEDI_B200_H
#define SEDI_B200_H

#ifdef SEDI_HAVE_MPI
#include <mpi.h>
#else
typedef int MPI_Comm; /* single-process stand-in; reference: library.h:19 includes mpi.h */
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (1) drop-in boundary: interfaceToLammps/library.h ------------------------------------------------ */
void lammps_open(int argc, char **argv, MPI_Comm comm, void **ptr);            /* library.h:29, library.cpp:44-49  */
void lammps_close(void *ptr);                                                  /* library.h:30, library.cpp:55-59  */
void lammps_file(void *ptr, char *path);                                       /* library.h:31, library.cpp:65-69  */
char *lammps_command(void *ptr, char *line);                                   /* library.h:32, library.cpp:75-79  */
void lammps_sync(void *ptr);                                                   /* library.h:34, library.cpp:82-87  */
int lammps_get_global_n(void *ptr);                                            /* library.h:35, library.cpp:96-101 */
void lammps_get_initial_np(void *ptr, int *np);                                /* library.h:38, library.cpp:109-135 */
void lammps_get_initial_info(void *ptr, double *coords, double *velos, double *diam, double *rho, int *tag,
                             int *lmpCpuId, int *type);                        /* library.h:40-42, library.cpp:142-206 */
int lammps_get_local_n(void *ptr);                                             /* library.h:45, library.cpp:210-217 */
void lammps_get_local_domain(void *ptr, double *domain);                       /* library.h:48, library.cpp:222-240 */
void lammps_get_local_info(void *ptr, double *coords, double *velos, int *foamCpuId, int *lmpCpuId,
                           int *tag);                                          /* library.h:51-52, library.cpp:246-308 */
void lammps_put_local_info(void *ptr, int nLocalIn, double *fdrag, double *DuDt, int *foamCpuIdIn,
                           int *tagIn);                                        /* library.h:55-56, library.cpp:314-367 */
void lammps_step(void *ptr, int n);                                            /* library.h:58, library.cpp:372-386 */
void lammps_set_timestep(void *ptr, double dt);                                /* library.h:59, library.cpp:398-402 */
double lammps_get_timestep(void *ptr);                                         /* library.h:60, library.cpp:390-394 */
void lammps_create_particle(void *ptr, int npAdd, double *position, double *tag, double diameter, double rho,
                            int type, double *vel);                            /* library.h:61-62, library.cpp:406-503 */
void lammps_delete_particle(void *ptr, int *deleteList, int nDelete);          /* library.h:63, library.cpp:507-621 */

/* ---- (2) device-resident engine API -------------------------------------------------------------------- */
int sedi_abi_version(void);
int sedi_device_count(void); /* number of CUDA devices visible (0 on a CPU-only box; never aborts) */
/* the parsed script as JSON (host only): pair / fix / group / dump settings as the engine understood them; returns the
 * length written, or -(needed size) when cap is too small */
int sedi_config_json(void *ptr, char *buf, int cap);
void sedi_set_device(void *ptr, int dev); /* before the first compute call; default: $SEDI_DEVICE, $LOCAL_RANK, 0 */

/* programmatic equivalents of `read_data` (box header + Atoms section: id type diameter density x y z) */
void sedi_set_box(void *ptr, const double *lo, const double *hi, int ntypes);
void sedi_add_atoms(void *ptr, int n, const int *tag, const int *type, const double *diameter, const double *density,
                    const double *x, const double *v /* may be NULL */);
void sedi_set_omega(void *ptr, int n, const int *tag, const double *omega);

/* full state download for tests / checkpoints, rows in device order; any pointer may be NULL */
void sedi_get_state(void *ptr, double *x, double *v, double *omega, double *f, double *torque, double *radius,
                    double *rmass, int *tag, int *type, int *mask);
/* directed neighbour list as (tag_i, tag_j, meta) rows; returns the row count; call with cap = 0 to size.
 * meta: bit30 granular list, bit31 type-cutoff list, bits 25..29 periodic image code (13 = none). */
long long sedi_get_pairs(void *ptr, int *tag_i, int *tag_j, unsigned *meta, int *touch, double *shear, long long cap);
void sedi_get_wall_shear(void *ptr, int wall, double *shear /* [n][3], device order */);
/* list statistics per owned row (device order): directed list entries, and how many of them overlapped in the last
 * sub-step (contact-history mask); either pointer may be NULL */
void sedi_get_row_stats(void *ptr, int *entries, int *touching);
void sedi_force_rebuild(void *ptr);
/* which: 0 neighbour rebuilds, 1 undirected granular pair evaluations, 2 DEM steps, 3 directed granular entries,
 *        4 directed type-list entries, 5 ELL row capacity, 6 kernel launches issued, 7 local particle count,
 *        8 undirected granular list size as LAMMPS counts it (owned-ghost pairs stored by both owners), 9 ghost rows,
 *        10 pair evaluations counted once per undirected pair system-wide (directed entries / 2, summed over steps) */
long long sedi_get_stat(void *ptr, int which);
void sedi_reset_stats(void *ptr);
void sedi_synchronize(void *ptr);
void *sedi_stream(void *ptr); /* the cudaStream_t every kernel of this engine is launched on */
/* last event-timed DEM kernel time of lammps_step / sedi_step in milliseconds (CUDA events on the engine stream) */
double sedi_last_step_ms(void *ptr);
/* CUDA-event stopwatch on the engine stream (bench.py's timed region) */
void sedi_timer_start(void *ptr);
double sedi_timer_stop_ms(void *ptr);
/* per-kernel timing of the fused DEM sub-step kernel: CUDA events bracket each back-to-back group of k_step
 * launches; sedi_get_profile returns the number of executed launches and their summed duration */
void sedi_profile(void *ptr, int on);
long long sedi_get_profile(void *ptr, double *kernel_ms);

/* --- coupling (mirror of enhancedCloud, device resident).  Mesh = single-block uniform blockMesh. */
void sedi_mesh_box(void *ptr, const double *lo, const double *hi, const int *ncell);
/* Rectilinear mesh: graded blocks (`simpleGrading`) and axis-aligned blocks stacked into one tensor-product grid.
 * xfaces/yfaces/zfaces hold ncell[d] + 1 ascending face coordinates (as in the host's mesh.points()); cell_label[i + nx
 * (j + ny k)] is the host solver's label of that cell (blockMesh numbers cells block by block), NULL = the tensor index.
 * Cell owner = the face interval containing the particle centre (replaces softParticle::move tracking,
 * lammpsFoam/softParticle.C:102-151).  Diffusion smoothing uses the finite-volume Laplacian of the cell widths. */
void sedi_mesh_rectilinear(void *ptr, const int *ncell, const double *xfaces, const double *yfaces, const double *zfaces,
                           const int *cell_label);
int sedi_mesh_ncells(void *ptr);
/* drag model / force switches: names of constant/cloudProperties (enhancedCloud.C:586-598) */
#define SEDI_DRAG_ERGUN_WENYU_ID 0
#define SEDI_DRAG_SYAMLAL_OBRIEN_ID 1
#define SEDI_FORCE_DRAG_BIT 1
#define SEDI_FORCE_PGRAD_BIT 2
#define SEDI_FORCE_BUOY_BIT 4
#define SEDI_FORCE_ADDEDMASS_BIT 8
#define SEDI_FORCE_LIFT_BIT 16
#define SEDI_FORCE_HISTORY_BIT 32   /* particleHistoryForce, enhancedCloud.C:197-234 (the state migrates with its particle) */
#define SEDI_FORCE_WALL_LUB_BIT 64  /* lubricationForce against the y = 0 wall, enhancedCloud.C:235-248 */
#define SEDI_FORCE_INLET_BIT 128    /* inletForce inside inletBox, enhancedCloud.C:249-257 */
void sedi_coupling_config(void *ptr, int drag_model, int force_flags, double nub, double rhob, const double *g,
                          double deltaT);
/* host cell fields -> device (Uf, gradp, DDtU, curlU are [C][3]; gamma is [C]); NULL = leave unchanged / absent */
void sedi_put_cell_fields(void *ptr, const double *Uf, const double *gamma, const double *gradp, const double *DDtU,
                          const double *curlU);
/* runTime().timeIndex() seen by sedi_compute_fluid_force (particleHistoryForce only, enhancedCloud.C:197-234).  The host
 * sets it once per fluid time step, BEFORE the force evaluations of that step: with subCycles > 1 evolve() evaluates the
 * force several times per fluid step and every evaluation sees the same index.  The library never advances it itself. */
void sedi_coupling_time_index(void *ptr, int time_index);
/* inletForce vector, inletBox (x1 x2 y1 y2 z1 z2 r1 r2 -), addParticleOption (1 box, 2 hollow cylinder) and
 * addParticleBoxEccentricity of constant/cloudProperties (softParticleCloud.C:471, :1354-1415) */
void sedi_coupling_inlet(void *ptr, const double *inlet_force, const double *inlet_box, int region_option,
                         const double *eccentricity);
/* history-force state per owned particle in sedi_get_state row order: sumDeltaFb[n][3], n0[n] (softParticle.H:104-107) */
void sedi_get_history_state(void *ptr, double *sumDeltaFb, double *n0);
/* locate particles in cells (cell owner index, int32, -1 = outside) */
void sedi_locate(void *ptr);
/* updateParticleUr + updateParticleAlpha + Jd + updateDragOnParticles: writes fix fdrag's per-atom force on device */
void sedi_compute_fluid_force(void *ptr);
/* particleToEulerianField: gamma[C], Ue[C][3] (host pointers, may be NULL to keep results on the device) */
void sedi_scatter_alpha_u(void *ptr, double *gamma, double *Ue);
/* calcTcFields: Asrc[C][3] ; Omega[C] is identically zero in the reference (enhancedCloud.C:391) */
void sedi_calc_tc(void *ptr, double *Asrc, double *Omega);
/* diffusion smoothing of the Eulerian particle fields (enhancedCloud::smoothField, enhancedCloud.C:790-907):
 * bandwidth = diffusionBandWidth, steps = diffusionSteps, Ddiag = diagonal of smoothDirection (NULL = identity),
 * flags = which fields are smoothed inside the sedi_* calls (names of constant/cloudProperties, :573-576) */
#define SEDI_SMOOTH_UF_BIT 1
#define SEDI_SMOOTH_UP_BIT 2
#define SEDI_SMOOTH_DRAG_BIT 4
#define SEDI_SMOOTH_ALPHA_BIT 8
void sedi_smooth_config(void *ptr, double bandwidth, int steps, const double *Ddiag, int flags);
void sedi_smooth_uf(void *ptr);                                  /* Uf (1-gamma) -> smooth -> / (1-gamma), :675-690 */
void sedi_smooth_field(void *ptr, double *field, int ncomp);    /* smooth a host cell field in place (ncomp 1 or 3) */
int sedi_smooth_last_iters(void *ptr);                           /* PCG iterations of the last solve */
void sedi_enable_diag(void *ptr, int on); /* keep Uri/|Uri|/alpha/Jd per particle at the next sedi_compute_fluid_force */
/* diagnostics of the last sedi_compute_fluid_force, device order: cell[n], Uri[n][3], magUri[n], alpha[n], Jd[n],
 * F[n][3]; any pointer may be NULL */
void sedi_get_coupling_diag(void *ptr, int *cell, double *Uri, double *magUri, double *alphap, double *Jd, double *F);
/* same as lammps_step but never touches host particle arrays */
void sedi_step(void *ptr, int n);
/* OpenFOAM lagrangian fields of the cloud (softParticle::writeFields, lammpsFoam/softParticleIO.C:157-197): positions (with
 * owner cell), d, tag, lmpCpuId, type, U, ensembleU -- plus density and n0, which readFields (:113-152) requires -- as ASCII
 * IOField files in `dir`; `location` is the FoamFile header's location string, e.g. "0.1/lagrangian/cloud" (may be NULL).
 * On several GPUs every rank writes "<field>.<rank>". */
void sedi_write_lagrangian(void *ptr, const char *dir, const char *location);
/* the reference's built-in invariants, printed by it every step: "total F before / after" of calcTcFields
 * (enhancedCloud.C:395-435: sum_c Asrc V (1 - gamma) before and after smoothing) and "total U solid before / after" of
 * particleToEulerianField (:936-976: sum_p Vp Up, and sum_c Ue V gamma after smoothing).  Off by default (four small
 * reductions and host round trips per coupling step); vectors of 3, any pointer may be NULL. */
void sedi_enable_conservation_sums(void *ptr, int on);
void sedi_get_conservation_sums(void *ptr, double *Ftotal_before, double *Ftotal_after, double *Utotal_before, double *Utotal_after);
/* enhancedCloud::averageInfo (enhancedCloud.C:1341-1370): total particle volume, sum Vp Up [3], volume-averaged velocity [3] */
void sedi_average_info(void *ptr, double *totalVolume, double *totalVel, double *averageVel);
/* the reference's timers (writeCPUTime.H:1-19), accumulated wall-clock seconds: diffusionTimeCount[2] (enhancedCloud.H:228),
 * particleMoveTime (:234: cell-owner location), cpuTimeSplit[6] (softParticleCloud.H:351-354: assemble, transpose, flatten,
 * foam->lammps, lammps, lammps->foam; the first three are the all-to-alls this library does not have: 0) */
void sedi_get_timers(void *ptr, double *diffusionTimeCount, double *particleMoveTime, double *cpuTimeSplit);

/* --- multi-GPU: one process per GPU, brick decomposition of the particle box (LAMMPS `processors Px Py Pz`).
 * A host that links with -DSEDI_HAVE_MPI needs none of these calls: the library bootstraps NCCL from the MPI_Comm handed to
 * lammps_open / new LAMMPS (rank 0 creates the id, MPI_Bcast distributes it); SEDI_AUTO_COMM=1 does the same from the launcher
 * environment (RANK / WORLD_SIZE, OMPI_*, PMI_*, SLURM_*) through a rendezvous file.  A host that bootstraps itself: the NCCL
 * unique id is produced by rank 0 (sedi_comm_unique_id) and distributed by the host (torch.distributed in bench.py);
 * procgrid may be NULL (taken from the script / factorised). */
int sedi_comm_unique_id(void *out, int cap); /* returns the id size in bytes, 0 when NCCL is unavailable */
int sedi_comm_init(void *ptr, int rank, int nranks, const void *nccl_unique_id, int id_bytes, const int *procgrid);
int sedi_comm_rank(void *ptr);
/* which: 0 ghost refreshes issued, 1 border rows sent per refresh, 2 ghost rows, 3 neighbour links, 4 arrivals at the last rebuild */
long long sedi_comm_stat(void *ptr, int which);
/* host-only decomposition logic: LAMMPS-style processor grid, owner rank of a position, neighbour links of a brick
 * (peers[26], offsets[26][3], shifts[26][3]; returns the link count) */
void sedi_decomp_grid(int nranks, const double *boxlen, int *grid);
int sedi_decomp_owner(const double *x, const double *boxlo, const double *boxhi, const int *grid);
int sedi_decomp_links(int rank, const int *grid, const int *periodic, const double *prd, int *peers, int *offsets, double *shifts);

#ifdef __cplusplus
}
#endif
#endif /* SEDI_B200_H */
