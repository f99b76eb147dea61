// oracle/stubs: Input. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_INPUT_H
#define SEDI_STUB_INPUT_H
#include "variable.h"
namespace LAMMPS_NS { class Input { public: Variable *variable; Input() : variable(0) {} }; }
#endif
