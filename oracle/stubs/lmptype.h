// oracle/stubs: minimal stand-in for LAMMPS lmptype.h (lammps-1Feb14, EXTERNAL, not vendored in the reference).
// TEST INFRASTRUCTURE ONLY -- declares just what the five reference plug-in sources touch.
#ifndef SEDI_STUB_LMPTYPE_H
#define SEDI_STUB_LMPTYPE_H
#include <stdint.h>
#include <limits.h>
namespace LAMMPS_NS {
typedef int tagint;
typedef int64_t bigint;
typedef int imageint;
#define NEIGHMASK 0x3FFFFFFF
#define MAXSMALLINT INT_MAX
#define BIGINT_FORMAT "%ld"
}
#ifndef FLERR
#define FLERR __FILE__,__LINE__
#endif
#ifndef MAX
#define MAX(A,B) ((A) > (B) ? (A) : (B))
#endif
#ifndef MIN
#define MIN(A,B) ((A) < (B) ? (A) : (B))
#endif
#endif
