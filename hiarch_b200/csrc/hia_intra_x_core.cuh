// a13 intra-x response through row prefix sums: the per-(row, centre) arithmetic, shared by the
// device kernels (hia_gf.cu) and a host build of the very same functions (tests/test_intra_x_core.py
// compiles this header with g++ and checks it against the oracle without a GPU).
//
// Reference: get_intra_x_kernal / get_intra_x,
// publish/confor/src/S1_preprocess/S1_2_find_anchors/S2_2_find_by_intra_x.py:25-50, :116-126.
// For a centre c the kernel weight of a lower-triangle cell (row y, column x <= y) is
//     w_c(x, y) = max(0, 1 - |k x - y + n| / (k_len sqrt(1 + k^2))),   k = (c - n) / c   (c >= 1)
//     w_0(x, y) = max(0, 1 - x / k_len)                                 (c == 0, vertical line)
// and np.tril(K).T + np.tril(K, -1) gives the upper cell (x, y) the same weight. Hence
//     sum(M * K_c) = sum_r [ sum_{q <= r} M[r][q] w_c(q, r)  +  sum_{q > r} M[r][q] w_c(r, q) ].
// Along a row of M both terms are TENT functions of the column q (two linear pieces each), so
// with the row prefix sums P0[j] = sum_{q < j} M[r][q] and P1[j] = sum_{q < j} q M[r][q] one
// (row, centre) pair costs <= 8 prefix look-ups instead of ~2 k_len cells: O(n^2) for all centres
// instead of the reference's O(n^3) (and the O(n^2 k_len) of the direct band evaluation).
// The weight is continuous in (x, y), so a boundary cell attributed to the neighbouring piece or
// dropped at weight ~1e-16 changes nothing above the fp64 rounding level.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HIA_HD __host__ __device__ __forceinline__
#define HIA_ALIGN16 __align__(16)
#else
#define HIA_HD inline
#define HIA_ALIGN16 __attribute__((aligned(16)))
#endif

namespace hia {

struct HIA_ALIGN16 IxPrefix { double s0, s1; };   // (sum of values, sum of column * value) over columns < j

struct IxCentre {          // per-centre constants, hoisted out of the row loop
  int vertical;            // c == 0
  double a;                // -k = (n - c) / c  > 0
  double inv_a;            // 1 / a
  double inv_D;            // 1 / (k_len sqrt(1 + k^2))
  double D;                // k_len sqrt(1 + k^2): half-width of the tent across columns of the upper part
  double h;                // D / a: half-width across columns of the lower part
  double inv_klen;
};

HIA_HD IxCentre ix_centre(int c, int n, int k_len) {
  IxCentre cc;
  cc.vertical = (c == 0);
  cc.inv_klen = 1.0 / (double)k_len;
  if (cc.vertical) { cc.a = cc.inv_a = cc.inv_D = cc.D = cc.h = 0.0; return cc; }
  const double k = ((double)c - (double)n) / (double)c;
  const double nrm = sqrt(1.0 + k * k);
  cc.a = -k;
  cc.inv_a = 1.0 / cc.a;
  cc.D = (double)k_len * nrm;
  cc.inv_D = 1.0 / cc.D;
  cc.h = cc.D * cc.inv_a;
  return cc;
}

// clamp a real bound to [lo, hi + 1] BEFORE the int conversion (x0, h reach ~n^2)
HIA_HD int ix_clamp_int(double v, int lo, int hi) {
  if (!(v > (double)lo)) return lo;
  if (v > (double)hi + 1.0) return hi + 1;
  return (int)v;
}

// sum_{q = lo}^{hi} (alpha + beta q) m[q] from the prefix sums. Branch-free: an empty piece
// (lo > hi) reads P[hi + 1] twice and contributes exactly 0, so the look-ups of all pieces of a
// (row, centre) pair are independent loads the hardware can overlap. Requires 0 <= hi + 1 <= n and,
// for a non-empty piece, 0 <= lo.
HIA_HD double ix_piece(const IxPrefix* P, int lo, int hi, double alpha, double beta) {
  lo = lo > hi ? hi + 1 : lo;
  const IxPrefix p1 = P[hi + 1], p0 = P[lo];
  return alpha * (p1.s0 - p0.s0) + beta * (p1.s1 - p0.s1);
}
// sum_{q = lo}^{hi} (alpha + beta q)
HIA_HD double ix_linsum(int lo, int hi, double alpha, double beta) {
  const int c = hi - lo + 1;
  const double cnt = c > 0 ? (double)c : 0.0;
  return alpha * cnt + beta * 0.5 * ((double)lo + (double)hi) * cnt;
}

// Contribution of row r of M (prefix sums P[0..n]) to centre c: num += sum(M[r][:] * K_c[r][:]),
// den += the kernel mass attributed to row r (lower cells: 2 w off the diagonal, w on it).
HIA_HD void ix_row_contrib(const IxCentre& cc, int n, int k_len, int r, const IxPrefix* P,
                           double& num, double& den) {
  if (cc.vertical) {
    // lower part: w = 1 - q / k_len for q < k_len
    const int hi = r < k_len - 1 ? r : k_len - 1;
    const double beta = -cc.inv_klen;
    num += ix_piece(P, 0, hi, 1.0, beta);
    den += 2.0 * ix_linsum(0, hi, 1.0, beta);
    if (r <= hi) den -= 1.0 + beta * (double)r;
    // upper part: constant weight 1 - r / k_len over q in (r, n)
    if (r < k_len) num += ix_piece(P, r + 1, n - 1, 1.0 - (double)r * cc.inv_klen, 0.0);
    return;
  }
  // ---- lower part (y = r, x = q in [0, r]): g(q) = (n - r) - a q, w = 1 - |g| / D
  {
    const double nr = (double)(n - r);
    const double x0 = nr * cc.inv_a;                      // g >= 0  <=>  q <= x0
    const double t = nr * cc.inv_D, b = cc.a * cc.inv_D;
    const int xs = ix_clamp_int(floor(x0), -1, r);        // last column of the rising piece (<= r)
    const int lo_a = ix_clamp_int(floor(x0 - cc.h) + 1.0, 0, r);
    const int hi_a = xs < r ? xs : r;
    const int lo_b = xs + 1 > 0 ? xs + 1 : 0;
    int hi_b = ix_clamp_int(ceil(x0 + cc.h) - 1.0, -1, r);
    if (hi_b > r) hi_b = r;
    const double al_a = 1.0 - t, be_a = b;                // rising:  w = 1 - (nr - a q) / D
    const double al_b = 1.0 + t, be_b = -b;               // falling: w = 1 + (nr - a q) / D
    num += ix_piece(P, lo_a, hi_a, al_a, be_a) + ix_piece(P, lo_b, hi_b, al_b, be_b);
    den += 2.0 * (ix_linsum(lo_a, hi_a, al_a, be_a) + ix_linsum(lo_b, hi_b, al_b, be_b));
    if (lo_a <= r && r <= hi_a) den -= al_a + be_a * (double)r;        // the diagonal cell counts once
    else if (lo_b <= r && r <= hi_b) den -= al_b + be_b * (double)r;
  }
  // ---- upper part (x = r, y = q in (r, n)): g(q) = (n - a r) - q, w = 1 - |g| / D
  {
    const double q0 = (double)n - cc.a * (double)r;       // g >= 0  <=>  q <= q0
    const double t = q0 * cc.inv_D;
    const int qs = ix_clamp_int(floor(q0), r, n - 1);     // in [r, n]
    const int lo_a = ix_clamp_int(floor(q0 - cc.D) + 1.0, r + 1, n - 1);
    const int hi_a = qs < n - 1 ? qs : n - 1;
    const int lo_b = qs + 1;
    int hi_b = ix_clamp_int(ceil(q0 + cc.D) - 1.0, r, n - 1);
    if (hi_b > n - 1) hi_b = n - 1;
    num += ix_piece(P, lo_a, hi_a, 1.0 - t, cc.inv_D) + ix_piece(P, lo_b, hi_b, 1.0 + t, -cc.inv_D);
  }
}

}  // namespace hia
