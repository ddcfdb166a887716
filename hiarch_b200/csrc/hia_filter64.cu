// GF_S1 used_filter (publish/confor/src/S1_preprocess/S1_1_filtering.py:13-16) as ONE op on a
// block: box filter (size x size, reflect) followed by norm_to_zero_mean
// (publish/oe_map/S4_norm_value.py:10-27), computed internally in fp64 with scipy's exact
// operation order.
//
// Why a bit-replica: Hi-C values are discrete (count levels, or %.2f-quantised files), so the
// box-filtered values are massively TIED around the p34 / p66 percentiles that select the
// "core" of norm_to_zero_mean; which tied cells pass the strict `vmin < x < vmax` test is decided
// by the last-bit rounding of scipy's running sums (NI_UniformFilter1D: tmp = sum of the first
// window in order; tmp += new - old; out = tmp / size; axis 0 then axis 1, fp64 intermediate).
// Any other summation order flips those cells and moves the block's mean/std by ~1e-2. So this
// path reproduces that order exactly (verified bit-for-bit against scipy in
// tests/test_host_tables.py::test_uniform_filter_replica), selects the percentiles exactly on
// 64-bit order-preserving keys, and only rounds to fp32 at the very end.
//
// The fp32 input is promoted to the reference's fp64 values: when `decimals >= 0` every value is
// re-quantised as nearbyint(x * 10^decimals) / 10^decimals in fp64, i.e. the double a "%.2f"
// text field parses to (the pipeline hands maps over as %.2f text, new_hic_class.py:2298).
#include "hia_common.cuh"
#include "hia_box64_core.cuh"

namespace hia {
namespace {

__device__ __forceinline__ unsigned long long double_to_ordered(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_to_double(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// grid (column groups of 256, row stripes): no integer division per element
__global__ void __launch_bounds__(256)
load_f64_kernel(const float* __restrict__ x, long long ld, int rows, int cols, double scale,
                double* __restrict__ a) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    double v = (double)x[(long long)r * ld + c];
    if (scale > 0.0) v = nearbyint(v * scale) / scale;   // the double the decimal text parses to
    a[(long long)r * cols + c] = v;
  }
}

// uniform_filter1d along the SLOW axis of a [len][lines] array (element (l, c) at in[l * lines + c]):
// thread per line c, the whole line sequentially in scipy's order (tmp = sum of the first window in
// order; tmp += new - old; out = tmp / size), coalesced across the threads of a warp. The
// arithmetic chain is sequential by construction (every partial sum is rounded), so the only
// parallelism is across lines; memory-level parallelism comes from the interior loop, where no
// index reflects and 8 steps of loads (independent of the chain) are issued before they are consumed.
__global__ void __launch_bounds__(64)
box64_lines_kernel(const double* __restrict__ in, double* __restrict__ out, int len, int lines, int size) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < lines) box64_line(in, out, len, lines, c, size);
}

// out[c][r] = in[r][c] (32 x 32 tiles through shared memory, both sides coalesced): the fast-axis
// pass of the box filter runs as a slow-axis pass on the transposed array.
__global__ void __launch_bounds__(256)
transpose64_kernel(const double* __restrict__ in, double* __restrict__ out, int rows, int cols) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = in[(long long)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) out[(long long)c * rows + r] = tile[tx][i];
  }
}

// ---- exact order statistics on 64-bit keys: 8 passes of 8 bits (3 over the map, 5 over the
// candidates the third one keeps), up to 4 ranks. Ranks whose key
// prefixes coincide share one histogram ("slot": the two ranks of a percentile nearly always do,
// and in pass 0 all four), and the shared-memory adds are aggregated per warp by (slot, digit)
// with match.any: box-filtered maps put almost every key into a handful of leading digits, and
// un-aggregated same-address atomics serialise.
struct Sel64 {
  unsigned long long hist[4][256];     // per slot
  unsigned long long prefix[4];        // per rank: key bits decided so far (right aligned)
  unsigned long long rank[4];          // per rank: residual rank inside its prefix
  unsigned long long slot_prefix[4];
  int slot_of_rank[4];
  int n_slots;
  int pad;
  unsigned long long cand_n;           // candidates kept by the compacting pass
  double gamma[2];
  double pct[2];       // p_hi (66), p_lo (34)
  double sums[6];
  double mean_std[2];
};

__global__ void sel64_init_kernel(Sel64* st, long long n, double q_hi, double q_lo) {
  const int t = threadIdx.x;
  for (int i = t; i < 4 * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0ull;
  if (t < 6) st->sums[t] = 0.0;
  if (t == 0) {
    const double qs[2] = {q_hi, q_lo};
    for (int q = 0; q < 2; ++q) {
      const double h = (double)(n - 1) * qs[q];            // numpy 'linear': (n - 1) * q/100
      double lo = floor(h);
      long long lo_i = (long long)lo;
      if (lo_i < 0) lo_i = 0;
      if (lo_i > n - 1) lo_i = n - 1;
      const long long hi_i = (lo_i + 1 <= n - 1) ? lo_i + 1 : n - 1;
      st->gamma[q] = h - lo;
      st->rank[2 * q] = (unsigned long long)lo_i;
      st->rank[2 * q + 1] = (unsigned long long)hi_i;
    }
    for (int r = 0; r < 4; ++r) { st->prefix[r] = 0ull; st->slot_of_rank[r] = 0; st->slot_prefix[r] = 0ull; }
    st->n_slots = 1;
    st->cand_n = 0ull;
  }
}

// One pass: per-slot histogram of digit `pass` over the keys whose higher bits equal a slot prefix.
// The source is the whole array (n_host elements) or, when n_dev is given, the candidate list a
// COMPACT pass left behind. COMPACT (the pass with 16 decided bits): every element that matches a
// slot prefix is also appended to `cand` -- a few per cent of a map unless it is massively tied --
// so that the remaining five passes stream the candidates instead of the map. The append is
// aggregated per CTA iteration (1024 elements): warp ballots -> shared counter -> ONE global add.
template <bool COMPACT>
__global__ void __launch_bounds__(256)
sel64_pass_kernel(const double* __restrict__ a, long long n_host, const unsigned long long* __restrict__ n_dev,
                  int pass, Sel64* st, int aggregate, double* __restrict__ cand, unsigned long long* cand_count) {
  __shared__ unsigned int s_hist[4 * 256];
  __shared__ unsigned int s_cnt;
  __shared__ unsigned long long s_base;
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) s_hist[i] = 0u;
  const int n_slots = st->n_slots;
  unsigned long long pf[4];                              // ~0 never equals a (<= 56-bit) prefix
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) pf[sl] = sl < n_slots ? st->slot_prefix[sl] : ~0ull;
  const long long n = n_dev ? (long long)*n_dev : n_host;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr long long kChunk = 4 * 256;
  // block-uniform trip count: every thread reaches the votes and the barriers
  for (long long base = (long long)blockIdx.x * kChunk; base < n; base += (long long)gridDim.x * kChunk) {
    double v[4];
    unsigned int agg[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = base + k * 256 + tid;
      v[k] = i < n ? a[i] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long i = base + k * 256 + tid;
      agg[k] = 0xffffffffu;
      if (i < n) {
        const unsigned long long key = double_to_ordered(v[k]);
        const unsigned long long hi = (pass == 0) ? 0ull : (key >> (shift + 8));
        const unsigned int dig = (unsigned int)((key >> shift) & 0xffull);
#pragma unroll
        for (int sl = 0; sl < 4; ++sl)
          if (hi == pf[sl]) agg[k] = (unsigned int)sl * 256u + dig;       // slot prefixes are distinct
      }
      if (!aggregate) {
        if (agg[k] != 0xffffffffu) atomicAdd(&s_hist[agg[k]], 1u);
      } else if (__any_sync(0xffffffffu, agg[k] != 0xffffffffu)) {
        const unsigned int peers = __match_any_sync(0xffffffffu, agg[k]);
        if (agg[k] != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&s_hist[agg[k]], (unsigned int)__popc(peers));
      }
    }
    if (COMPACT) {
      if (tid == 0) s_cnt = 0u;
      __syncthreads();
      unsigned int bal[4], tot = 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) { bal[k] = __ballot_sync(0xffffffffu, agg[k] != 0xffffffffu); tot += __popc(bal[k]); }
      unsigned int woff = 0u;
      if (lane == 0 && tot) woff = atomicAdd(&s_cnt, tot);
      woff = __shfl_sync(0xffffffffu, woff, 0);
      __syncthreads();
      if (tid == 0 && s_cnt) s_base = atomicAdd(cand_count, (unsigned long long)s_cnt);
      __syncthreads();
      if (tot) {
        unsigned long long dst = s_base + woff;
        const unsigned int below = (1u << lane) - 1u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (agg[k] != 0xffffffffu) cand[dst + __popc(bal[k] & below)] = v[k];
          dst += __popc(bal[k]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_slots * 256; i += blockDim.x) {
    const unsigned int c = s_hist[i];
    if (c) atomicAdd(&(&st->hist[0][0])[i], (unsigned long long)c);
  }
}

// after each pass: every rank finds its digit in its slot's histogram, the prefixes grow by 8 bits,
// the slot list (distinct prefixes) is rebuilt, the histograms are cleared
__global__ void sel64_scan_kernel(Sel64* st) {
  __shared__ unsigned long long s_newprefix[4];
  const int r = threadIdx.x;
  if (r < 4) {
    const unsigned long long* h = st->hist[st->slot_of_rank[r]];
    unsigned long long cum = 0, target = st->rank[r];
    int bin = 255;
    for (int i = 0; i < 256; ++i) {
      const unsigned long long c = h[i];
      if (target < cum + c) { bin = i; break; }
      cum += c;
    }
    st->rank[r] = target - cum;
    s_newprefix[r] = (st->prefix[r] << 8) | (unsigned long long)bin;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ns = 0;
    for (int q = 0; q < 4; ++q) {
      st->prefix[q] = s_newprefix[q];
      int found = -1;
      for (int sl = 0; sl < ns; ++sl) if (st->slot_prefix[sl] == s_newprefix[q]) { found = sl; break; }
      if (found < 0) { st->slot_prefix[ns] = s_newprefix[q]; found = ns++; }
      st->slot_of_rank[q] = found;
    }
    st->n_slots = ns;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) (&st->hist[0][0])[i] = 0ull;
}

__global__ void sel64_finish_kernel(Sel64* st) {
  const int q = threadIdx.x;
  if (q >= 2) return;
  const double a = ordered_to_double(st->prefix[2 * q]);
  const double b = ordered_to_double(st->prefix[2 * q + 1]);
  const double t = st->gamma[q], d = b - a;
  st->pct[q] = (t >= 0.5) ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(a, __dmul_rn(d, t));  // numpy _lerp, no FMA
}

__global__ void __launch_bounds__(256)
moments64_kernel(const double* __restrict__ a, long long n, Sel64* st) {
  __shared__ double s_red[8][6];
  const double hi = st->pct[0], lo = st->pct[1];
  const double c = 0.5 * (lo + hi);
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double v = a[i], d = v - c;
    acc[3] += 1.0; acc[4] += d; acc[5] += d * d;
    if (v < hi && v > lo) { acc[0] += 1.0; acc[1] += d; acc[2] += d * d; }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 6; ++k) { const double s = warp_sum(acc[k]); if (lane == 0) s_red[w][k] = s; }
  __syncthreads();
  if (threadIdx.x < 6) {
    double s = 0; for (int i = 0; i < 8; ++i) s += s_red[i][threadIdx.x];
    atomicAdd(&st->sums[threadIdx.x], s);
  }
}

__global__ void moments64_finish_kernel(Sel64* st) {
  if (threadIdx.x != 0) return;
  const double hi = st->pct[0], lo = st->pct[1], c = 0.5 * (lo + hi);
  const bool use_core = (lo < hi) && (st->sums[0] > 0.0);            // S4_norm_value.py:16-22
  const double cnt = use_core ? st->sums[0] : st->sums[3];
  const double s1 = use_core ? st->sums[1] : st->sums[4];
  const double s2 = use_core ? st->sums[2] : st->sums[5];
  const double m = s1 / cnt;
  double var = s2 / cnt - m * m;
  if (var < 0) var = 0;
  st->mean_std[0] = m + c;
  st->mean_std[1] = sqrt(var);
}

__global__ void zscore64_kernel(const double* __restrict__ a, int rows, int cols, const Sel64* st,
                                float* __restrict__ out, long long ld_out, int normalise) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  double v = a[i];
  if (normalise) v = (v - st->mean_std[0]) / st->mean_std[1];
  out[(long long)r * ld_out + c] = (float)v;
}

// the same from the TRANSPOSED fp64 array at[c][r] (what the transposed fast-axis pass leaves)
__global__ void __launch_bounds__(256)
zscore64_t_kernel(const double* __restrict__ at, int rows, int cols, const Sel64* st,
                  float* __restrict__ out, long long ld_out, int normalise) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) tile[i][tx] = at[(long long)c * rows + r];
  }
  __syncthreads();
  const double mean = normalise ? st->mean_std[0] : 0.0, sd = normalise ? st->mean_std[1] : 1.0;
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) {
      double v = tile[tx][i];
      if (normalise) v = (v - mean) / sd;
      out[(long long)r * ld_out + c] = (float)v;
    }
  }
}

}  // namespace
}  // namespace hia

HIA_API int64_t hia_used_filter_f64_workspace_bytes(int32_t rows, int32_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  return 2 * (int64_t)rows * cols * 8 + (int64_t)sizeof(hia::Sel64) + 512;
}

HIA_API int hia_used_filter_f64(const float* x, int32_t rows, int32_t cols, int64_t ld_in, int32_t size,
                                int32_t decimals, int normalise, double mid_vmax_per, float* out,
                                int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !out || !workspace || rows <= 0 || cols <= 0 || ld_in < cols || ld_out < cols || size < 0)
    return HIA_ERR_BAD_ARG;
  if (workspace_bytes < hia_used_filter_f64_workspace_bytes(rows, cols)) return HIA_ERR_WORKSPACE;
  unsigned char* ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  const long long n = (long long)rows * cols;
  double* A = reinterpret_cast<double*>(ws);
  double* B = A + n;
  Sel64* st = reinterpret_cast<Sel64*>(B + n);
  const int blocks = (int)((n + 255) / 256);
  double scale = 0.0;
  if (decimals >= 0) { scale = 1.0; for (int i = 0; i < decimals; ++i) scale *= 10.0; }
  load_f64_kernel<<<dim3((cols + 255) / 256, std::min(rows, 16384)), 256, 0, s>>>(x, ld_in, rows, cols, scale, A);
  HIA_LAUNCH_CHECK();
  const bool filtered = size > 1;                        // scipy: size 1 (and 0) leave the block as is
  if (filtered) {
    // axis 0 on A -> B; axis 1 = the same slow-axis pass on B^T: B -> A (transpose) -> B. The
    // filtered block ends up TRANSPOSED in B; the order statistics / moments below do not care.
    box64_lines_kernel<<<(cols + 63) / 64, 64, 0, s>>>(A, B, rows, cols, size);
    HIA_LAUNCH_CHECK();
    transpose64_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, s>>>(B, A, rows, cols);
    HIA_LAUNCH_CHECK();
    box64_lines_kernel<<<(rows + 63) / 64, 64, 0, s>>>(A, B, cols, rows, size);
    HIA_LAUNCH_CHECK();
  }
  const double* F = filtered ? B : A;
  if (normalise) {
    const double q_hi = mid_vmax_per / 100.0, q_lo = (100.0 - mid_vmax_per) / 100.0;
    sel64_init_kernel<<<1, 256, 0, s>>>(st, n, q_hi, q_lo);
    HIA_LAUNCH_CHECK();
    const int grid = (int)std::min<long long>(blocks, (long long)num_sms() * 8);
    const int pgrid = (int)std::min<long long>((n + 1023) / 1024, (long long)num_sms() * 8);
    const int aggregate = sel_aggregate();
    double* cand = filtered ? A : B;                     // the fp64 array that is free at this point
    constexpr int kCompactPass = 2;                      // 16 bits decided: sign, exponent, 4 mantissa bits
    for (int pass = 0; pass < 8; ++pass) {
      if (pass < kCompactPass)
        sel64_pass_kernel<false><<<pgrid, 256, 0, s>>>(F, n, nullptr, pass, st, aggregate, nullptr, nullptr);
      else if (pass == kCompactPass)
        sel64_pass_kernel<true><<<pgrid, 256, 0, s>>>(F, n, nullptr, pass, st, aggregate, cand, &st->cand_n);
      else
        sel64_pass_kernel<false><<<pgrid, 256, 0, s>>>(cand, 0, &st->cand_n, pass, st, aggregate, nullptr, nullptr);
      HIA_LAUNCH_CHECK();
      sel64_scan_kernel<<<1, 256, 0, s>>>(st);
      HIA_LAUNCH_CHECK();
    }
    sel64_finish_kernel<<<1, 32, 0, s>>>(st);
    HIA_LAUNCH_CHECK();
    moments64_kernel<<<grid, 256, 0, s>>>(F, n, st);
    HIA_LAUNCH_CHECK();
    moments64_finish_kernel<<<1, 32, 0, s>>>(st);
    HIA_LAUNCH_CHECK();
  }
  if (filtered)
    zscore64_t_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, s>>>(F, rows, cols, st, out, ld_out, normalise);
  else
    zscore64_kernel<<<blocks, 256, 0, s>>>(F, rows, cols, st, out, ld_out, normalise);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}
