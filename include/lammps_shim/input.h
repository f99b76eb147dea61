/* input.h -- stand-in for LAMMPS' Input class: one(line) executes a single input-script command
 * (reference use: /root/reference/lammpsFoam/softParticleCloud.C:106). */
#ifndef SEDI_SHIM_INPUT_H
#define SEDI_SHIM_INPUT_H
#include "lammps.h"

namespace LAMMPS_NS {

class Input {
 public:
  explicit Input(LAMMPS *l) : lmp(l) {}
  char *one(const char *line); /* returns NULL, like a command that is not `run`-style output */
  void file(const char *path);

 private:
  LAMMPS *lmp;
};

}  // namespace LAMMPS_NS
#endif
