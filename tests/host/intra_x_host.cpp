// Host build of the per-(row, centre) arithmetic of the prefix-sum intra-x kernels
// (hiarch_b200/csrc/hia_intra_x_core.cuh), compiled with g++ by tests/test_intra_x_core.py.
// Same loop nest as the device path: row prefix sums, then every centre adds its row contribution.
#include <stdint.h>
#include <vector>
#include "../../hiarch_b200/csrc/hia_intra_x_core.cuh"

extern "C" int intra_x_prefix_host(const float* m, int n, int64_t ld, int k_len, double* out) {
  using namespace hia;
  std::vector<IxPrefix> P(n + 1);
  std::vector<double> num(n, 0.0), den(n, 0.0);
  std::vector<IxCentre> cc(n);
  for (int c = 0; c < n; ++c) cc[c] = ix_centre(c, n, k_len);
  for (int r = 0; r < n; ++r) {
    P[0].s0 = P[0].s1 = 0.0;
    for (int q = 0; q < n; ++q) {
      const double v = (double)m[(int64_t)r * ld + q];
      P[q + 1].s0 = P[q].s0 + v;
      P[q + 1].s1 = P[q].s1 + (double)q * v;
    }
    for (int c = 0; c < n; ++c) ix_row_contrib(cc[c], n, k_len, r, P.data(), num[c], den[c]);
  }
  for (int c = 0; c < n; ++c) out[c] = num[c] / den[c];
  return 0;
}
