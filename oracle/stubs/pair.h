// oracle/stubs: Pair base class. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_PAIR_H
#define SEDI_STUB_PAIR_H
#include "pointers.h"
namespace LAMMPS_NS {
class NeighList;
class Pair : protected Pointers {
 public:
  int evflag, vflag_fdotr, vflag_either, eflag_either, no_virial_fdotr_compute;
  NeighList *list; double *svector; double **cutsq;
  Pair(LAMMPS *l) : Pointers(l), evflag(0), vflag_fdotr(0), vflag_either(0), eflag_either(0),
    no_virial_fdotr_compute(0), list(0), svector(0), cutsq(0) {}
  virtual ~Pair() {}
  virtual void compute(int, int) = 0;
  virtual void settings(int, char **) = 0;
  virtual void init_style() {}
  virtual double single(int, int, int, int, double, double, double, double &) { return 0.0; }
  void ev_setup(int, int) { evflag = 0; }
  void ev_tally_xyz(int, int, int, int, double, double, double, double, double, double, double, double) {}
  void v_tally_tensor(int, int, int, int, double, double, double, double, double, double) {}
};
}
#endif
