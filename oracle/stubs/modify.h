// oracle/stubs: Modify. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_MODIFY_H
#define SEDI_STUB_MODIFY_H
#include "pointers.h"
namespace LAMMPS_NS {
class Fix;
class Modify { public: int nfix; Fix **fix; Modify() : nfix(0), fix(0) {} };
}
#endif
