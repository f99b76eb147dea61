// oracle/stubs: NeighList. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_NEIGH_LIST_H
#define SEDI_STUB_NEIGH_LIST_H
#include "pointers.h"
namespace LAMMPS_NS {
class NeighList {
 public:
  int index, inum; int *ilist, *numneigh; int **firstneigh; double **firstdouble; NeighList *listgranhistory;
  NeighList() : index(0), inum(0), ilist(0), numneigh(0), firstneigh(0), firstdouble(0), listgranhistory(0) {}
};
}
#endif
