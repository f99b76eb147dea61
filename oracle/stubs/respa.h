// oracle/stubs: Respa (never instantiated; only cast targets). TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_RESPA_H
#define SEDI_STUB_RESPA_H
namespace LAMMPS_NS {
class Respa { public: int nlevels; void copy_flevel_f(int) {} void copy_f_flevel(int) {} };
}
#endif
