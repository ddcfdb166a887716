// Host build of the fp64 box-filter line function (hiarch_b200/csrc/hia_box64_core.cuh), driven
// exactly like hia_used_filter_f64 drives it on the device: slow-axis pass, transpose, slow-axis
// pass on the transposed array, transpose back. Compiled with g++ -ffp-contract=off by
// tests/test_box64_core.py.
#include <stdint.h>
#include <vector>
#include "../../hiarch_b200/csrc/hia_box64_core.cuh"

extern "C" int box64_host(const double* a, int rows, int cols, int size, double* out) {
  using namespace hia;
  const long long n = (long long)rows * cols;
  std::vector<double> B(n), T(n);
  for (int c = 0; c < cols; ++c) box64_line(a, B.data(), rows, cols, c, size);       // axis 0
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) T[(long long)c * rows + r] = B[(long long)r * cols + c];
  for (int r = 0; r < rows; ++r) box64_line(T.data(), B.data(), cols, rows, r, size);  // axis 1 on the transpose
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) out[(long long)r * cols + c] = B[(long long)c * rows + r];
  return 0;
}
