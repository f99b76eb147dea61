// oracle/stubs: per-atom arrays touched by the reference plug-ins. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_ATOM_H
#define SEDI_STUB_ATOM_H
#include "pointers.h"
namespace LAMMPS_NS {
class Atom {
 public:
  double **x, **v, **f, **omega, **torque, **angmom, **extra;
  double *radius, *rmass, *mass;
  int *type, *mask, *tag;
  int nlocal, nghost, nmax;
  bigint natoms;
  int sphere_flag;
  Atom() : x(0), v(0), f(0), omega(0), torque(0), angmom(0), extra(0), radius(0), rmass(0), mass(0),
           type(0), mask(0), tag(0), nlocal(0), nghost(0), nmax(0), natoms(0), sphere_flag(1) {}
  void add_callback(int) {}
  void delete_callback(const char *, int) {}
};
}
#endif
