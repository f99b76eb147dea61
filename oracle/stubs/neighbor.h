// oracle/stubs: Neighbor (requests are recorded, lists are supplied by the oracle harness). TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_NEIGHBOR_H
#define SEDI_STUB_NEIGHBOR_H
#include "pointers.h"
#include "neigh_request.h"
namespace LAMMPS_NS {
class Neighbor {
 public:
  int ago; int nrequest; NeighRequest *requests[16]; void *requestor[16];
  Neighbor() : ago(0), nrequest(0) {}
  int request(void *who) { requests[nrequest] = new NeighRequest(); requestor[nrequest] = who; return nrequest++; }
  void build_one(int) {}
};
}
#endif
