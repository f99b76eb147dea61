// oracle/stubs: Fix base class. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_FIX_H
#define SEDI_STUB_FIX_H
#include "pointers.h"
#include <string.h>
namespace LAMMPS_NS {
class NeighList;
class Fix : protected Pointers {
 public:
  char *id, *style; int igroup, groupbit;
  int force_reneighbor; bigint next_reneighbor;
  int restart_peratom, create_attribute, local_flag, size_local_rows, size_local_cols;
  double **array_local; int xflag;
  Fix(LAMMPS *l, int narg, char **arg) : Pointers(l), igroup(0), groupbit(1), force_reneighbor(0),
    next_reneighbor(0), restart_peratom(0), create_attribute(0), local_flag(0), size_local_rows(0),
    size_local_cols(0), array_local(0), xflag(0) {
    id = strdup(narg > 0 ? arg[0] : ""); style = strdup(narg > 2 ? arg[2] : "");
  }
  virtual ~Fix() {}
  virtual int setmask() = 0;
  virtual void init() {}
  virtual void init_list(int, NeighList *) {}
  virtual void setup(int) {}
  virtual void post_force(int) {}
  virtual void post_force_respa(int, int, int) {}
  virtual void min_post_force(int) {}
  virtual void *extract(const char *, int &) { return 0; }
  virtual void set_arrays(int) {}
  virtual void grow_arrays(int) {}
  virtual void copy_arrays(int, int, int) {}
  virtual int pack_exchange(int, double *) { return 0; }
  virtual int unpack_exchange(int, double *) { return 0; }
  virtual int pack_restart(int, double *) { return 0; }
  virtual void unpack_restart(int, int) {}
  virtual int size_restart(int) { return 0; }
  virtual int maxsize_restart() { return 0; }
  virtual void reset_dt() {}
  virtual double memory_usage() { return 0.0; }
  void set_groupbit(int b) { groupbit = b; }
};
namespace FixConst {
  static const int INITIAL_INTEGRATE = 1<<0, POST_INTEGRATE = 1<<1, PRE_EXCHANGE = 1<<2, PRE_NEIGHBOR = 1<<3,
    PRE_FORCE = 1<<4, POST_FORCE = 1<<5, FINAL_INTEGRATE = 1<<6, END_OF_STEP = 1<<7, THERMO_ENERGY = 1<<8,
    INITIAL_INTEGRATE_RESPA = 1<<9, POST_INTEGRATE_RESPA = 1<<10, PRE_FORCE_RESPA = 1<<11,
    POST_FORCE_RESPA = 1<<12, FINAL_INTEGRATE_RESPA = 1<<13, MIN_PRE_EXCHANGE = 1<<14,
    MIN_PRE_FORCE = 1<<15, MIN_POST_FORCE = 1<<16, MIN_ENERGY = 1<<17, POST_RUN = 1<<18;
}
using namespace FixConst;
}
#endif
