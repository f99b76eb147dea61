// oracle/stubs: Error -> print + abort. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_ERROR_H
#define SEDI_STUB_ERROR_H
#include "pointers.h"
#include <stdio.h>
#include <stdlib.h>
namespace LAMMPS_NS {
class Error {
 public:
  void all(const char *f, int l, const char *m) { fprintf(stderr, "ERROR: %s (%s:%d)\n", m, f, l); abort(); }
  void one(const char *f, int l, const char *m) { all(f, l, m); }
  void warning(const char *, int, const char *m, int = 1) { fprintf(stderr, "WARNING: %s\n", m); }
};
}
#endif
