// oracle/stubs: LAMMPS Pointers base + LAMMPS root object. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_POINTERS_H
#define SEDI_STUB_POINTERS_H
#include "lmptype.h"
#include "mpi.h"
#include <stdio.h>
namespace LAMMPS_NS {
class Atom; class Update; class Force; class Neighbor; class Comm; class Memory; class Error;
class Domain; class Modify; class Input; class Group;
class LAMMPS {
 public:
  Atom *atom; Update *update; Force *force; Neighbor *neighbor; Comm *comm; Memory *memory;
  Error *error; Domain *domain; Modify *modify; Input *input; Group *group; MPI_Comm world;
  LAMMPS() : atom(0), update(0), force(0), neighbor(0), comm(0), memory(0), error(0), domain(0),
             modify(0), input(0), group(0), world(0) {}
};
class Pointers {
 public:
  Pointers(LAMMPS *p) : lmp(p), memory(p->memory), error(p->error), atom(p->atom), update(p->update),
    force(p->force), neighbor(p->neighbor), comm(p->comm), domain(p->domain), modify(p->modify),
    input(p->input), group(p->group), world(p->world) {}
  virtual ~Pointers() {}
 protected:
  LAMMPS *lmp;
  Memory *&memory; Error *&error; Atom *&atom; Update *&update; Force *&force; Neighbor *&neighbor;
  Comm *&comm; Domain *&domain; Modify *&modify; Input *&input; Group *&group; MPI_Comm &world;
};
}
#endif
