// oracle/stubs: state of stock PairGranHookeHistory that gran/hertzFix/history inherits. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_PAIR_GRAN_HOOKE_HISTORY_H
#define SEDI_STUB_PAIR_GRAN_HOOKE_HISTORY_H
#include "pair.h"
namespace LAMMPS_NS {
class Fix;
class PairGranHookeHistory : public Pair {
 public:
  int computeflag;
  double kn, kt, gamman, gammat, xmu; int dampflag; double dt; int freeze_group_bit;
  int neighprev; Fix *fix_rigid; double *mass_rigid; int nmax;
  PairGranHookeHistory(LAMMPS *l) : Pair(l), computeflag(0), kn(0), kt(0), gamman(0), gammat(0), xmu(0),
    dampflag(0), dt(0), freeze_group_bit(0), neighprev(0), fix_rigid(0), mass_rigid(0), nmax(0) {
    svector = new double[4];
  }
  virtual void compute(int, int) {}
  virtual void settings(int, char **) {}
};
}
#endif
