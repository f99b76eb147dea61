// oracle/stubs: FixDeform (only a cast target). TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_FIX_DEFORM_H
#define SEDI_STUB_FIX_DEFORM_H
#include "fix.h"
namespace LAMMPS_NS {
class FixDeform : public Fix {
 public:
  int remapflag;
  FixDeform(LAMMPS *l, int n, char **a) : Fix(l, n, a), remapflag(0) {}
  int setmask() { return 0; }
};
}
#endif
