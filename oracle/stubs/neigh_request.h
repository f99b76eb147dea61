// oracle/stubs: NeighRequest. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_NEIGH_REQUEST_H
#define SEDI_STUB_NEIGH_REQUEST_H
namespace LAMMPS_NS {
class NeighRequest { public: int pair, fix, half, full, gran, granhistory, dnum; NeighRequest() : pair(1), fix(0), half(1), full(0), gran(0), granhistory(0), dnum(0) {} };
}
#endif
