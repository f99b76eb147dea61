// oracle/stubs: Domain. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_DOMAIN_H
#define SEDI_STUB_DOMAIN_H
#include "pointers.h"
namespace LAMMPS_NS {
class Domain {
 public:
  int xperiodic, yperiodic, zperiodic; double xprd, yprd, zprd; double prd[3]; double h_rate[6], h_ratelo[3];
  double boxlo[3], boxhi[3], sublo[3], subhi[3];
  Domain() : xperiodic(0), yperiodic(0), zperiodic(0), xprd(1), yprd(1), zprd(1) {
    for (int i = 0; i < 3; i++) { prd[i] = 1; h_ratelo[i] = 0; boxlo[i] = sublo[i] = 0; boxhi[i] = subhi[i] = 1; }
    for (int i = 0; i < 6; i++) h_rate[i] = 0;
  }
  void x2lamda(double *x, double *l) { for (int i = 0; i < 3; i++) l[i] = (x[i] - boxlo[i]) / prd[i]; }
};
}
#endif
