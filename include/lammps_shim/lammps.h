/* lammps.h -- drop-in stand-in for the LAMMPS class header that sediFoam's Foam side includes
 * (/root/reference/lammpsFoam/include/LammpsCollection.H:8).  softParticleCloud uses exactly three things of the
 * C++ class (softParticleCloud.C:62, :106, :357): `new LAMMPS(0, NULL, comm)`, `lmp_->input->one(line)` and
 * `delete lmp_`; everything else goes through the C functions of library.h with the LAMMPS* as the `void *` handle.
 * libsedi_b200.so defines this class, so the Foam side compiles and links unchanged. */
#ifndef SEDI_SHIM_LAMMPS_H
#define SEDI_SHIM_LAMMPS_H
#include "../sedi_b200.h" /* MPI_Comm (real <mpi.h> when SEDI_HAVE_MPI is defined) */

namespace LAMMPS_NS {

class Input;

class LAMMPS {
 public:
  Input *input;   /* lmp_->input->one(line) */
  void *engine;   /* the B200 particle engine behind this instance */
  MPI_Comm world;
  LAMMPS(int narg, char **arg, MPI_Comm communicator);
  ~LAMMPS();

 private:
  LAMMPS(const LAMMPS &);
  LAMMPS &operator=(const LAMMPS &);
};

}  // namespace LAMMPS_NS
#endif
