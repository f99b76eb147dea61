// oracle/stubs: Update. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_UPDATE_H
#define SEDI_STUB_UPDATE_H
#include "pointers.h"
namespace LAMMPS_NS {
class Update {
 public:
  double dt; bigint ntimestep; int setupflag; char *integrate_style; void *integrate;
  Update() : dt(0), ntimestep(0), setupflag(0), integrate_style((char *)"verlet"), integrate(0) {}
};
}
#endif
