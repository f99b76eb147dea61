/* library.h -- the C interface of interfaceToLammps/library.h:29-63, served by libsedi_b200.so. */
#ifndef SEDI_SHIM_LIBRARY_H
#define SEDI_SHIM_LIBRARY_H
#include "../sedi_b200.h"
#endif
