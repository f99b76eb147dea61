// oracle/oracle_backend.hpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Interface between the oracle's CPU time loop (oracle_driver.hpp, a restatement of the EXTERNAL LAMMPS
// Verlet/Neighbor/FixShearHistory semantics listed in SURVEY.md Appendix A) and the force kernels.
// Two implementations exist:
//   * PortBackend (oracle_port.cpp): our own restatement of the reference arithmetic, each function citing the
//     reference file:line it follows.  Travels to the GPU box; this is the checker the -m gpu tests use.
//   * RefBackend  (ref_backend.cpp): the reference's *own* unmodified sources compiled by path from
//     /root/reference/interfaceToLammps against oracle/stubs/, built into oracle/_ref/libsedi_ref.so.
//     Used to pin the port bit-for-bit, and as the "reference" CPU baseline.
#pragma once
#include "../sedifoam_b200/csrc/lmp_script.hpp"

namespace ora {

// LAMMPS memory layout: double** rows over one contiguous [n][3] block, owned atoms first then ghosts.
struct AtomView {
  int nlocal, nghost;
  double **x, **v, **f, **omega, **torque;
  double *radius, *rmass;
  int *type, *mask, *tag;
};

// LAMMPS NeighList (+ listgranhistory for the granular list).
struct NList {
  int inum;
  int *ilist, *numneigh;
  int **firstneigh;
  int **firsttouch;     // granular history only
  double **firstshear;  // granular history only, 3 doubles per neighbour
};

struct StepInfo {
  double dt;       // update->dt (live)
  double dt_init;  // value cached by init_style()/init() at the first run
  long long ntimestep;
  int setupflag;
};

struct BoxInfo {
  double lo[3], hi[3];
  int periodic[3];
};

// per-atom state of `fix fdrag` (fix_fluid_drag.h:30-33)
struct FdragState {
  double **ffluiddrag, **DuDt, **vOld;
  int *foamCpuId;
};

class Backend {
 public:
  virtual ~Backend() {}
  virtual const char *name() const = 0;
  // called once at the first run (LAMMPS init()): styles read their settings, lubricate computes R0/RT0/RS0
  virtual void init(const sedi::SimConfig &cfg, const AtomView &av, const BoxInfo &box, const StepInfo &st) = 0;
  virtual void pair_granular(const AtomView &av, const NList &list, const StepInfo &st) = 0;
  virtual void pair_lubricate(const AtomView &av, const NList &full, const StepInfo &st) = 0;
  virtual void fix_fdrag(int ifix, const AtomView &av, const FdragState &fs, const StepInfo &st) = 0;
  virtual void fix_cohesive(int ifix, const AtomView &av, const NList &half, const StepInfo &st) = 0;
  virtual void fix_wall(int ifix, const AtomView &av, double **shear, const StepInfo &st) = 0;
};

Backend *make_port_backend();
#ifdef SEDI_HAVE_REF
Backend *make_ref_backend();
#endif

}  // namespace ora
