// a6: whole-matrix percentiles with numpy's linear-interpolation semantics, clip, the
// "core" mean/std of norm_to_zero_mean, z-score.
//
// Replaces tid_off_outliers (publish/oe_map/S3_denoise.py:85-92), norm_to_zero_mean /
// remove_outliers / norm_de_mtx (publish/oe_map/S4_norm_value.py:10-46) and the percentile of
// keep_high_cons (publish/confor/src/S1_preprocess/S1_2_find_anchors/U2_keep_top_cons.py:20-41).
//
// np.percentile(a, q) (method 'linear') needs two order statistics: with h = (m-1) * q/100,
// lo = floor(h), g = h - lo: result = lerp(a[lo], a[min(lo+1, m-1)], g). The order statistics
// are found EXACTLY by a most-significant-digit radix select over order-preserving uint32 keys
// (3 streaming passes of 11/11/10 bits, 4 B/cell each), for up to 4 percentiles (8 ranks) at
// once; ranks that share a key prefix share a histogram. Subset percentiles (`a[a > 0]`,
// `a[a < 0]`) are absolute ranks offset by the subset counts, which pass 1 also produces, so
// everything stays on the device (no host round trip between passes).
#include "hia_common.cuh"

namespace hia {
namespace {

constexpr int kMaxQueries = 4;
constexpr int kMaxRanks = 2 * kMaxQueries;
constexpr int kBins = 2048;
constexpr int kSelThreads = 256;

struct SelState {                      // lives in the workspace (device)
  unsigned long long hist[kMaxRanks][kBins];   // per-slot digit histograms of the current pass
  unsigned long long cnt_pos, cnt_neg, n;
  unsigned long long rank[kMaxRanks];          // residual rank inside the current prefix
  unsigned int prefix[kMaxRanks];              // key bits decided so far (right aligned)
  int slot_of_rank[kMaxRanks];
  unsigned int slot_prefix[kMaxRanks];
  int n_slots;
  int n_ranks;
  int valid[kMaxQueries];
  double gamma[kMaxQueries];
};

__device__ __forceinline__ int pass_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ int pass_bits(int pass) { return pass == 2 ? 10 : 11; }

__global__ void sel_init_kernel(SelState* st, unsigned long long n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long* h = &st->hist[0][0];
  if (i < kMaxRanks * kBins) h[i] = 0ull;
  if (i == 0) { st->cnt_pos = 0; st->cnt_neg = 0; st->n = n; st->n_slots = 1; st->slot_prefix[0] = 0; }
}

// One streaming pass: per-slot histogram of the current digit of all keys whose higher bits
// equal the slot's prefix. Pass 0 has a single slot (empty prefix) and also counts x>0 / x<0.
// Grid (column groups of 1024, row stripes): a thread owns one float4 column group and walks
// down its stripe of rows -- no integer division in the loop, slot prefixes in registers.
template <int PASS>
__global__ void __launch_bounds__(kSelThreads)
sel_pass_kernel(const float* __restrict__ x, int rows, int cols, long long ld, SelState* st, int aggregate) {
  extern __shared__ unsigned int s_hist[];             // [n_slots][kBins]
  __shared__ unsigned long long s_pos, s_neg;
  const int n_slots = (PASS == 0) ? 1 : st->n_slots;
  for (int i = threadIdx.x; i < n_slots * kBins; i += blockDim.x) s_hist[i] = 0u;
  if (threadIdx.x == 0) { s_pos = 0ull; s_neg = 0ull; }
  unsigned int pf[kMaxRanks];                          // 0xffffffff never equals a (<= 22-bit) prefix
#pragma unroll
  for (int s = 0; s < kMaxRanks; ++s) pf[s] = (PASS != 0 && s < n_slots) ? st->slot_prefix[s] : 0xffffffffu;
  __syncthreads();
  constexpr int shift = (PASS == 0) ? 21 : (PASS == 1 ? 10 : 0);
  constexpr unsigned int mask = (PASS == 2) ? 0x3ffu : 0x7ffu;
  constexpr int hshift = (PASS == 0) ? 32 : (PASS == 1 ? 21 : 10);   // bits above the digit
  unsigned int pos = 0, neg = 0;
  const int lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * kSelThreads + threadIdx.x) * 4;
  const int nv = min(4, cols - c0);                    // <= 0: this thread owns no column (it still votes)
  const float* col = x + c0;
  // the trip count depends on blockIdx.y only: every lane of a warp reaches the match instruction.
  // kRowsInFlight rows are loaded before any is consumed (the 64 KB of histograms of passes 1 / 2
  // leave 3 CTAs per SM: the loads in flight per thread make up for the occupancy).
  constexpr int kRowsInFlight = 4;
  for (int r0 = blockIdx.y; r0 < rows; r0 += kRowsInFlight * gridDim.y) {
    float v[kRowsInFlight][4];
    int nvr[kRowsInFlight];
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
      const int r = r0 + u * gridDim.y;
      nvr[u] = r < rows ? nv : 0;
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (nvr[u] == 4) {
        const float4 t = ld_stream4(col + (long long)r * ld);
        v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
      } else if (nvr[u] > 0) {
        const float* p = col + (long long)r * ld;
        for (int j = 0; j < nvr[u]; ++j) v[u][j] = p[j];
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsInFlight; ++u) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned int key = float_to_ordered(v[u][j]);
        const bool on = j < nvr[u];
        if (PASS == 0) {
          pos += (on && v[u][j] > 0.f);
          neg += (on && v[u][j] < 0.f);
          const unsigned int agg = on ? (key >> shift) : 0xffffffffu;
          if (aggregate) {
            // the leading 11 bits of nearly all keys of a map fall into a handful of bins
            const unsigned int peers = __match_any_sync(0xffffffffu, agg);
            if (on && lane == __ffs(peers) - 1) atomicAdd(&s_hist[agg], (unsigned int)__popc(peers));
          } else if (on) {
            atomicAdd(&s_hist[agg], 1u);
          }
        } else {
          const unsigned int hi = (hshift >= 32) ? 0u : (key >> hshift);
          const unsigned int dig = (key >> shift) & mask;
          int slot = -1;                               // slot prefixes are distinct: at most one matches
#pragma unroll
          for (int s = 0; s < kMaxRanks; ++s) slot = (hi == pf[s]) ? s : slot;
          if (on && slot >= 0) atomicAdd(&s_hist[slot * kBins + dig], 1u);
        }
      }
    }
  }
  if (PASS == 0) {
    const unsigned long long wp = warp_sum((unsigned long long)pos), wn = warp_sum((unsigned long long)neg);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_pos, wp); atomicAdd(&s_neg, wn); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_slots * kBins; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&st->hist[i / kBins][i % kBins], (unsigned long long)s_hist[i]);
  if (PASS == 0 && threadIdx.x == 0) { atomicAdd(&st->cnt_pos, s_pos); atomicAdd(&st->cnt_neg, s_neg); }
}

// After pass 0: turn (mode, quantile) queries into absolute ranks (numpy 'linear' indexes).
__global__ void sel_ranks_kernel(SelState* st, const int* __restrict__ modes,
                                 const double* __restrict__ quantiles, int nq) {
  if (threadIdx.x != 0) return;
  st->n_ranks = 2 * nq;
  for (int q = 0; q < nq; ++q) {
    unsigned long long m = st->n, base = 0;
    if (modes[q] == 1) { m = st->cnt_pos; base = st->n - st->cnt_pos; }      // a[a > 0]
    else if (modes[q] == 2) { m = st->cnt_neg; base = 0; }                    // a[a < 0]
    st->valid[q] = m > 0;
    if (m == 0) { st->rank[2 * q] = 0; st->rank[2 * q + 1] = 0; st->gamma[q] = 0.0; continue; }
    const double h = (double)(m - 1) * quantiles[q];                          // (n - 1) * q/100
    double lo = floor(h);
    if (lo < 0) lo = 0;
    unsigned long long lo_i = (unsigned long long)lo;
    if (lo_i > m - 1) lo_i = m - 1;
    unsigned long long hi_i = (lo_i + 1 <= m - 1) ? lo_i + 1 : m - 1;
    st->gamma[q] = h - lo;
    st->rank[2 * q] = base + lo_i;
    st->rank[2 * q + 1] = base + hi_i;
  }
  for (int r = 0; r < 2 * nq; ++r) { st->prefix[r] = 0; st->slot_of_rank[r] = 0; }
}

// After each pass: locate every rank's digit in its slot histogram (warp per rank), extend the
// prefixes, rebuild the slot list (distinct prefixes), clear the histograms for the next pass.
__global__ void __launch_bounds__(kMaxRanks * 32)
sel_scan_kernel(SelState* st, int pass) {
  __shared__ unsigned int s_newprefix[kMaxRanks];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ranks = st->n_ranks;
  const int bits = pass_bits(pass);
  const int nb = 1 << bits;
  if (warp < n_ranks) {
    const unsigned long long* h = st->hist[st->slot_of_rank[warp]];
    const unsigned long long target = st->rank[warp];
    const int per = nb / 32;                            // bins per lane (64 or 32)
    unsigned long long mine = 0;
    for (int i = 0; i < per; ++i) mine += h[lane * per + i];
    unsigned long long incl = mine;                     // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const unsigned long long excl = incl - mine;
    const bool here = (target >= excl) && (target < incl);
    const unsigned int ballot = __ballot_sync(0xffffffffu, here);
    int bin = nb - 1;
    unsigned long long before = 0;
    if (here) {
      unsigned long long cum = excl;
      for (int i = 0; i < per; ++i) {
        const unsigned long long c = h[lane * per + i];
        if (target < cum + c) { bin = lane * per + i; before = cum; break; }
        cum += c;
      }
    }
    const int src = ballot ? (__ffs(ballot) - 1) : 0;
    bin = __shfl_sync(0xffffffffu, bin, src);
    before = __shfl_sync(0xffffffffu, before, src);
    if (lane == 0) {
      st->rank[warp] = ballot ? (target - before) : 0;
      s_newprefix[warp] = (st->prefix[warp] << bits) | (unsigned int)bin;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int ns = 0;
    for (int r = 0; r < n_ranks; ++r) {
      st->prefix[r] = s_newprefix[r];
      int found = -1;
      for (int s = 0; s < ns; ++s) if (st->slot_prefix[s] == s_newprefix[r]) { found = s; break; }
      if (found < 0) { st->slot_prefix[ns] = s_newprefix[r]; found = ns++; }
      st->slot_of_rank[r] = found;
    }
    st->n_slots = ns;
  }
  __syncthreads();
  unsigned long long* h = &st->hist[0][0];
  for (int i = threadIdx.x; i < kMaxRanks * kBins; i += blockDim.x) h[i] = 0ull;
}

// numpy _lerp: a + (b - a) t, or b - (b - a)(1 - t) for t >= 0.5
__global__ void sel_finish_kernel(const SelState* st, int nq, double* __restrict__ out) {
  const int q = threadIdx.x;
  if (q >= nq) return;
  if (!st->valid[q]) { out[q] = nan(""); return; }
  const double a = (double)ordered_to_float(st->prefix[2 * q]);
  const double b = (double)ordered_to_float(st->prefix[2 * q + 1]);
  const double t = st->gamma[q];
  const double d = b - a;
  // numpy rounds the product and the sum separately: no FMA contraction here
  out[q] = (t >= 0.5) ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(a, __dmul_rn(d, t));
}

// ---------------------------------------------------------------------------------------
// clip / moments / z-score (elementwise + reductions; parameters are device scalars)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
clip_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols, long long ld,
            const double* __restrict__ lo, const double* __restrict__ hi) {
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= cols) return;
  const bool has_lo = lo != nullptr, has_hi = hi != nullptr;
  const double dlo = has_lo ? *lo : 0.0, dhi = has_hi ? *hi : 0.0;
  const float flo = (float)dlo, fhi = (float)dhi;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const float* p = x + (long long)r * ld + c0;
    float* o = out + (long long)r * ld + c0;
    float v[4];
    const int nv = min(4, cols - c0);
    if (nv == 4) { const float4 t = ld_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else for (int j = 0; j < 4; ++j) v[j] = (j < nv) ? p[j] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // M[M > vmax] = vmax; M[M < vmin] = vmin   (comparisons in fp64 like the reference)
      if (has_hi && (double)v[j] > dhi) v[j] = fhi;
      if (has_lo && (double)v[j] < dlo) v[j] = flo;
    }
    if (nv == 4) st_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
    else for (int j = 0; j < nv; ++j) o[j] = v[j];
  }
}

// sums[0..2] = count, sum(x - c), sum((x - c)^2) over vmin < x < vmax; sums[3..5] the same over all.
__global__ void __launch_bounds__(256)
moments_kernel(const float* __restrict__ x, int rows, int cols, long long ld,
               const double* __restrict__ vmin, const double* __restrict__ vmax,
               double* __restrict__ sums) {
  __shared__ double s_red[8][6];
  const double lo = *vmin, hi = *vmax;
  const double c = 0.5 * (lo + hi);                    // shift: kills the cancellation in the variance
  double a[6] = {0, 0, 0, 0, 0, 0};
  const int c4 = (cols + 3) / 4;
  const long long total = (long long)rows * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4), c0 = (int)(i % c4) * 4;
    const float* p = x + (long long)r * ld + c0;
    float v[4];
    const int nv = min(4, cols - c0);
    if (nv == 4) { const float4 t = ld_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else for (int j = 0; j < 4; ++j) v[j] = (j < nv) ? p[j] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= nv) break;
      const double d = (double)v[j] - c;
      a[3] += 1.0; a[4] += d; a[5] += d * d;
      if ((double)v[j] < hi && (double)v[j] > lo) { a[0] += 1.0; a[1] += d; a[2] += d * d; }
    }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 6; ++k) { const double s = warp_sum(a[k]); if (lane == 0) s_red[w][k] = s; }
  __syncthreads();
  if (threadIdx.x < 6) {
    double s = 0; for (int i = 0; i < 8; ++i) s += s_red[i][threadIdx.x];
    atomicAdd(&sums[threadIdx.x], s);
  }
}

// norm_to_zero_mean's choice of the "core" (S4_norm_value.py:16-22) -> mean, population std
__global__ void moments_finish_kernel(const double* sums, const double* vmin, const double* vmax,
                                      double* mean_std) {
  if (threadIdx.x != 0) return;
  const double lo = *vmin, hi = *vmax, c = 0.5 * (lo + hi);
  const bool use_core = (lo < hi) && (sums[0] > 0.0);
  const double cnt = use_core ? sums[0] : sums[3];
  const double s1 = use_core ? sums[1] : sums[4];
  const double s2 = use_core ? sums[2] : sums[5];
  const double m = s1 / cnt;
  double var = s2 / cnt - m * m;
  if (var < 0) var = 0;
  mean_std[0] = m + c;
  mean_std[1] = sqrt(var);
}

__global__ void __launch_bounds__(256)
zscore_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols, long long ld,
              const double* __restrict__ mean_std) {
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= cols) return;
  const double mean = mean_std[0], sd = mean_std[1];
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    const float* p = x + (long long)r * ld + c0;
    float* o = out + (long long)r * ld + c0;
    float v[4];
    const int nv = min(4, cols - c0);
    if (nv == 4) { const float4 t = ld_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else for (int j = 0; j < 4; ++j) v[j] = (j < nv) ? p[j] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (float)(((double)v[j] - mean) / sd);   // M -= mean; M /= std
    if (nv == 4) st_stream4(o, make_float4(v[0], v[1], v[2], v[3]));
    else for (int j = 0; j < nv; ++j) o[j] = v[j];
  }
}

// remove_outliers: v = min(p98 of positives, -p2 of negatives); bounds = (-v, v)
__global__ void sym_bounds_kernel(const double* pcts, double* bounds) {
  if (threadIdx.x != 0) return;
  const double v = fmin(pcts[0], -pcts[1]);            // np.min([vmax, -vmin]) propagates NaN
  const double vv = (pcts[0] != pcts[0] || pcts[1] != pcts[1]) ? nan("") : v;
  bounds[0] = -vv;
  bounds[1] = vv;
}

}  // namespace
}  // namespace hia

HIA_API int64_t hia_percentiles_workspace_bytes(void) { return (int64_t)sizeof(hia::SelState) + 256; }

HIA_API int hia_percentiles(const float* x, int32_t rows, int32_t cols, int64_t ld, const int32_t* modes,
                            const double* quantiles, int32_t nq, double* out, void* workspace,
                            int64_t workspace_bytes, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !modes || !quantiles || !out || !workspace || rows <= 0 || cols <= 0 || ld < cols)
    return HIA_ERR_BAD_ARG;
  if (nq < 1 || nq > kMaxQueries) return HIA_ERR_UNSUPPORTED;
  if (!aligned16(x) || (ld & 3)) return HIA_ERR_ALIGN;
  if (workspace_bytes < hia_percentiles_workspace_bytes()) return HIA_ERR_WORKSPACE;
  SelState* st = reinterpret_cast<SelState*>((reinterpret_cast<uintptr_t>(workspace) + 15) & ~(uintptr_t)15);
  const unsigned long long n = (unsigned long long)rows * (unsigned long long)cols;
  sel_init_kernel<<<(kMaxRanks * kBins + 255) / 256, 256, 0, s>>>(st, n);
  HIA_LAUNCH_CHECK();
  const int gx = (cols + 4 * kSelThreads - 1) / (4 * kSelThreads);
  const int gy = std::max(1, std::min(rows, (num_sms() * 8 + gx - 1) / gx));
  const dim3 grid(gx, gy);
  const size_t sh_full = (size_t)kMaxRanks * kBins * 4;        // 64 KB
  // per call: the attribute belongs to the CURRENT device's copy of the function (a process may
  // drive several devices), and setting it is a few hundred nanoseconds
  HIA_CUDA(cudaFuncSetAttribute(sel_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_full));
  HIA_CUDA(cudaFuncSetAttribute(sel_pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_full));
  sel_pass_kernel<0><<<grid, kSelThreads, kBins * 4, s>>>(x, rows, cols, ld, st, sel_aggregate());
  HIA_LAUNCH_CHECK();
  sel_ranks_kernel<<<1, 32, 0, s>>>(st, modes, quantiles, nq);
  HIA_LAUNCH_CHECK();
  sel_scan_kernel<<<1, kMaxRanks * 32, 0, s>>>(st, 0);
  HIA_LAUNCH_CHECK();
  sel_pass_kernel<1><<<grid, kSelThreads, sh_full, s>>>(x, rows, cols, ld, st, 0);
  HIA_LAUNCH_CHECK();
  sel_scan_kernel<<<1, kMaxRanks * 32, 0, s>>>(st, 1);
  HIA_LAUNCH_CHECK();
  sel_pass_kernel<2><<<grid, kSelThreads, sh_full, s>>>(x, rows, cols, ld, st, 0);
  HIA_LAUNCH_CHECK();
  sel_scan_kernel<<<1, kMaxRanks * 32, 0, s>>>(st, 2);
  HIA_LAUNCH_CHECK();
  sel_finish_kernel<<<1, 32, 0, s>>>(st, nq, out);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int hia_clip(const float* x, float* out, int32_t rows, int32_t cols, int64_t ld, const double* lo,
                     const double* hi, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !out || rows <= 0 || cols <= 0 || ld < cols) return HIA_ERR_BAD_ARG;
  if (!aligned16(x) || !aligned16(out) || (ld & 3)) return HIA_ERR_ALIGN;
  dim3 grid((cols + 1023) / 1024, std::min(rows, 2048));
  clip_kernel<<<grid, 256, 0, s>>>(x, out, rows, cols, ld, lo, hi);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int hia_core_moments(const float* x, int32_t rows, int32_t cols, int64_t ld, const double* vmin,
                             const double* vmax, double* mean_std, double* workspace6, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !vmin || !vmax || !mean_std || !workspace6 || rows <= 0 || cols <= 0 || ld < cols)
    return HIA_ERR_BAD_ARG;
  if (!aligned16(x) || (ld & 3)) return HIA_ERR_ALIGN;
  HIA_CUDA(cudaMemsetAsync(workspace6, 0, 6 * sizeof(double), s));
  const long long vec = (long long)rows * ((cols + 3) / 4);
  const int grid = (int)std::min<long long>((vec + 255) / 256, (long long)num_sms() * 8);
  moments_kernel<<<grid, 256, 0, s>>>(x, rows, cols, ld, vmin, vmax, workspace6);
  HIA_LAUNCH_CHECK();
  moments_finish_kernel<<<1, 32, 0, s>>>(workspace6, vmin, vmax, mean_std);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int hia_zscore(const float* x, float* out, int32_t rows, int32_t cols, int64_t ld,
                       const double* mean_std, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !out || !mean_std || rows <= 0 || cols <= 0 || ld < cols) return HIA_ERR_BAD_ARG;
  if (!aligned16(x) || !aligned16(out) || (ld & 3)) return HIA_ERR_ALIGN;
  dim3 grid((cols + 1023) / 1024, std::min(rows, 2048));
  zscore_kernel<<<grid, 256, 0, s>>>(x, out, rows, cols, ld, mean_std);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int hia_sym_clip_bounds(const double* pcts2, double* bounds2, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!pcts2 || !bounds2) return HIA_ERR_BAD_ARG;
  sym_bounds_kernel<<<1, 32, 0, s>>>(pcts2, bounds2);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}
