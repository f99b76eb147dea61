// oracle/stubs: Variable. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_VARIABLE_H
#define SEDI_STUB_VARIABLE_H
namespace LAMMPS_NS { class Variable { public: int find(char *) { return -1; } double compute_equal(int) { return 0.0; } }; }
#endif
