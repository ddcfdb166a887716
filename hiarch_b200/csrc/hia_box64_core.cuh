// One line of scipy.ndimage.uniform_filter1d (NI_UniformFilter1D, mode 'reflect', fp64) in scipy's
// exact operation order, shared by the device kernel (hia_filter64.cu) and a host build of the very
// same function (tests/test_box64_core.py compiles it with g++ and compares bit for bit with scipy).
#pragma once

#if defined(__CUDACC__)
#define HIA_BOX_HD __host__ __device__ __forceinline__
#else
#define HIA_BOX_HD inline
#endif

namespace hia {

// scipy.ndimage 'reflect' (d c b a | a b c d | d c b a), any distance
HIA_BOX_HD int reflect_idx64(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}

// steps of loads in flight per thread in the interior (one thread per line: a 25 k-bin block runs
// only ~170 threads per SM, so the bytes in flight have to come from each thread)
constexpr int kBoxUnroll = 16;

// Line c of a [len][lines] array (element (l, c) at in[l * lines + c]), sequentially:
//   tmp = sum of the first window in order; out[0] = tmp / size;
//   tmp += new - old; out[l] = tmp / size.
// Every partial sum is rounded, so the chain is sequential by construction; the interior steps
// (no index reflects) issue kBoxUnroll steps of loads -- independent of the chain -- before consuming them.
HIA_BOX_HD void box64_line(const double* __restrict__ in, double* __restrict__ out, int len, int lines,
                           int c, int size) {
  const int s1 = size / 2;
  const long long ls = lines;
  double tmp = 0.0;
  for (int ll = 0; ll < size; ++ll) tmp += in[(long long)reflect_idx64(ll - s1, len) * ls + c];
  const double fs = (double)size;
  out[c] = tmp / fs;
  const int int_lo = s1 + 1;                 // first step whose outgoing index ll - 1 - s1 is >= 0
  const int int_hi = len - size + s1;        // last step whose incoming index ll + size - 1 - s1 is <= len - 1
  int ll = 1;
  for (; ll < len && ll < int_lo; ++ll) {
    const double nw = in[(long long)reflect_idx64(ll + size - 1 - s1, len) * ls + c];
    const double od = in[(long long)reflect_idx64(ll - 1 - s1, len) * ls + c];
    tmp += nw - od;
    out[(long long)ll * ls + c] = tmp / fs;
  }
  for (; ll + (kBoxUnroll - 1) <= int_hi; ll += kBoxUnroll) {
    const double* pn = in + (long long)(ll + size - 1 - s1) * ls + c;
    const double* po = in + (long long)(ll - 1 - s1) * ls + c;
    double* pw = out + (long long)ll * ls + c;
    double nw[kBoxUnroll], od[kBoxUnroll];
#pragma unroll
    for (int k = 0; k < kBoxUnroll; ++k) { nw[k] = pn[(long long)k * ls]; od[k] = po[(long long)k * ls]; }
#pragma unroll
    for (int k = 0; k < kBoxUnroll; ++k) { tmp += nw[k] - od[k]; pw[(long long)k * ls] = tmp / fs; }
  }
  for (; ll < len; ++ll) {
    const double nw = in[(long long)reflect_idx64(ll + size - 1 - s1, len) * ls + c];
    const double od = in[(long long)reflect_idx64(ll - 1 - s1, len) * ls + c];
    tmp += nw - od;
    out[(long long)ll * ls + c] = tmp / fs;
  }
}

}  // namespace hia
