// oracle/stubs: Force. lj units => all conversion factors 1. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_FORCE_H
#define SEDI_STUB_FORCE_H
#include "pointers.h"
#include <stdlib.h>
#include <string.h>
namespace LAMMPS_NS {
class Pair;
class Force {
 public:
  double nktv2p, vxmu2f; int newton_pair; Pair *pair; const char *pair_style;
  Force() : nktv2p(1.0), vxmu2f(1.0), newton_pair(0), pair(0), pair_style("") {}
  double numeric(const char *, int, char *s) { return atof(s); }
  int inumeric(const char *, int, char *s) { return atoi(s); }
  Pair *pair_match(const char *word, int exact) {
    if (exact && strcmp(pair_style, word) == 0) return (Pair *)1;
    if (!exact && strstr(pair_style, word)) return (Pair *)1;
    return 0;
  }
};
}
#endif
