// host_units.cu -- host-side execution of the __host__ __device__ helpers of the CUDA sources (no GPU needed): compiled with
// nvcc by tests/test_host_units.py.  The same functions run inside the kernels.
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>
#include <math.h>
#include <cuda_runtime.h>
#include "../../sedifoam_b200/csrc/sedi_device.cuh"
#include "../../sedifoam_b200/csrc/lmp_script.hpp"
#include "../../sedifoam_b200/csrc/sedi_couple.cuh"
int main() {
  std::mt19937 rng(7);
  int bad = 0;
  for (int n : {0, 1, 2, 3, 5, 17, 100, 1000, 1635, 1636, 1637, 5000, 40000, 300000}) {
    for (int mode = 0; mode < 3; mode++) {
      std::vector<int> v(n + 10, -7);
      for (int i = 0; i < n; i++) v[5 + i] = i * 3 + 1;
      if (mode == 0) std::shuffle(v.begin() + 5, v.begin() + 5 + n, rng);
      else if (mode == 1) std::reverse(v.begin() + 5, v.begin() + 5 + n);
      else for (int i = 0; i + 8 < n; i += 7) std::swap(v[5 + i], v[5 + i + 3]);
      std::vector<int> ref(v.begin() + 5, v.begin() + 5 + n);
      std::sort(ref.begin(), ref.end());
      sedi::fcell_shell_sort(v.data(), 5, 5 + n);
      bool ok = std::equal(ref.begin(), ref.end(), v.begin() + 5);
      for (int i = 0; i < 5; i++) ok = ok && v[i] == -7 && v[5 + n + i] == -7;
      if (!ok) { printf("FAIL n=%d mode=%d\n", n, mode); bad++; }
    }
  }
  printf(bad ? "shell sort: %d failures\n" : "shell sort ok\n", bad);
  return bad;
}
