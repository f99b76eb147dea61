// oracle/stubs: FixWall (only a cast target in pair_lubricate_poly.cpp). TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_FIX_WALL_H
#define SEDI_STUB_FIX_WALL_H
#include "fix.h"
namespace LAMMPS_NS {
class FixWall : public Fix {
 public:
  int nwall; int wallwhich[6]; double coord0[6]; int xstyle[6]; int xindex[6]; char *xstr[6];
  FixWall(LAMMPS *l, int n, char **a) : Fix(l, n, a), nwall(0) {}
  int setmask() { return 0; }
};
}
#endif
