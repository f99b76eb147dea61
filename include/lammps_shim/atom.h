/* atom.h -- included by LammpsCollection.H:10 but no member of Atom is touched by the Foam side. */
#ifndef SEDI_SHIM_ATOM_H
#define SEDI_SHIM_ATOM_H
#include "lammps.h"
#endif
