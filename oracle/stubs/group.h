// oracle/stubs: Group. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_GROUP_H
#define SEDI_STUB_GROUP_H
namespace LAMMPS_NS { class Group { public: int bitmask[32]; int find(const char *) { return 0; } }; }
#endif
