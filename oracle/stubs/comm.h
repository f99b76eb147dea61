// oracle/stubs: Comm. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_COMM_H
#define SEDI_STUB_COMM_H
#include "pointers.h"
namespace LAMMPS_NS {
class Pair;
class Comm { public: int me, ghost_velocity; Comm() : me(0), ghost_velocity(1) {} void forward_comm_pair(Pair *) {} };
}
#endif
