// oracle/stubs: single-process MPI stand-in (MPI is absent in this image). TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_MPI_H
#define SEDI_STUB_MPI_H
#include <string.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_SUM 0
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{ memcpy(r, s, (size_t)n * (size_t)t); return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Abort(MPI_Comm, int) { return 0; }
#endif
