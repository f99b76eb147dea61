// oracle/stubs: Memory create/grow/destroy for contiguous 1-D and 2-D arrays. TEST INFRASTRUCTURE ONLY.
#ifndef SEDI_STUB_MEMORY_H
#define SEDI_STUB_MEMORY_H
#include "pointers.h"
#include <stdlib.h>
namespace LAMMPS_NS {
class Memory {
 public:
  template <typename T> T *create(T *&a, int n, const char *) { a = (T *)malloc(sizeof(T) * (size_t)(n > 0 ? n : 1)); return a; }
  template <typename T> T *grow(T *&a, int n, const char *) { a = (T *)realloc(a, sizeof(T) * (size_t)(n > 0 ? n : 1)); return a; }
  template <typename T> void destroy(T *&a) { free(a); a = 0; }
  template <typename T> T **create(T **&a, int n1, int n2, const char *) {
    if (n1 < 1) n1 = 1;
    T *d = (T *)malloc(sizeof(T) * (size_t)n1 * n2);
    a = (T **)malloc(sizeof(T *) * (size_t)n1);
    for (int i = 0; i < n1; i++) a[i] = d + (size_t)i * n2;
    return a;
  }
  template <typename T> T **grow(T **&a, int n1, int n2, const char *s) {
    if (!a) return create(a, n1, n2, s);
    if (n1 < 1) n1 = 1;
    T *d = (T *)realloc(a[0], sizeof(T) * (size_t)n1 * n2);
    a = (T **)realloc(a, sizeof(T *) * (size_t)n1);
    for (int i = 0; i < n1; i++) a[i] = d + (size_t)i * n2;
    return a;
  }
  template <typename T> void destroy(T **&a) { if (a) { free(a[0]); free(a); } a = 0; }
};
}
#endif
