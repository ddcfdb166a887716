// a10: checkerboard strength -- the diagonal / entropy loop of get_local.
//
// Replaces get_short_sims + get_accordance (publish/complex/src/S2_get_entro.py:40-65) and the
// loop of get_local (publish/complex/src/main_local.py:30-55):
//   for d in [d_lo, d_hi): vals = both +-d diagonals of the (symmetric) distance matrix;
//       keep p2 < v < p98 (np.percentile 'linear', strict); append;
//       whenever more than `min_count` values are pending: Shannon entropy (nats) of their
//       29-bin histogram over np.linspace(0, 1.6, 30) (out-of-range values dropped), reset;
//   score = mean of the entropies (None if there is none).
//
// band_stats_kernel: one CTA per diagonal. The diagonal is staged in shared memory; the two
// percentiles need 4 exact order statistics of the *duplicated* list (each value appears twice:
// lower then upper diagonal), found by a block-wide radix select on order-preserving keys;
// then one pass counts the kept values and bins them. band_reduce_kernel replays the
// sequential "more than 200 pending" chunking over the per-diagonal counts.
#include "hia_common.cuh"

namespace hia {
namespace {

constexpr int kBandThreads = 512;
constexpr int kHistBins = 29;

__device__ unsigned int block_select(const float* s_vals, int cnt, unsigned int rank,
                                     unsigned int* s_hist, unsigned int* s_bcast) {
  unsigned int prefix = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const unsigned int key = float_to_ordered(s_vals[i]);
      if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      unsigned int mine = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) mine += s_hist[lane * 8 + i];
      unsigned int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned int excl = incl - mine;
      if (rank >= excl && rank < incl) {
        unsigned int cum = excl;
        for (int i = 0; i < 8; ++i) {
          const unsigned int c = s_hist[lane * 8 + i];
          if (rank < cum + c) { s_bcast[0] = (unsigned int)(lane * 8 + i); s_bcast[1] = rank - cum; break; }
          cum += c;
        }
      }
    }
    __syncthreads();
    prefix = (prefix << 8) | s_bcast[0];
    rank = s_bcast[1];
    __syncthreads();
  }
  return prefix;
}

__device__ __forceinline__ double np_lerp(double a, double b, double t) {
  const double d = b - a;
  return (t >= 0.5) ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, t))) : __dadd_rn(a, __dmul_rn(d, t));   // no FMA
}

__global__ void __launch_bounds__(kBandThreads)
band_stats_kernel(const float* __restrict__ adj, int n, long long ld, int d_lo,
                  const double* __restrict__ edges, double q_lo, double q_hi,
                  long long* __restrict__ kept, long long* __restrict__ hist) {
  extern __shared__ float s_vals[];
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned int s_bcast[2];
  __shared__ unsigned int s_bins[kHistBins];
  __shared__ unsigned int s_kept;
  __shared__ double s_edges[kHistBins + 1];
  const int d = d_lo + blockIdx.x;
  const int cnt = n - d;                                 // values on the upper diagonal
  long long* my_hist = hist + (long long)blockIdx.x * kHistBins;
  if (cnt <= 0) {                                        // d >= n: get_short_sims returns None
    if (threadIdx.x == 0) kept[blockIdx.x] = 0;
    if (threadIdx.x < kHistBins) my_hist[threadIdx.x] = 0;
    return;
  }
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_vals[i] = adj[(long long)i * ld + i + d];
  if (threadIdx.x <= kHistBins) s_edges[threadIdx.x] = edges[threadIdx.x];
  if (threadIdx.x < kHistBins) s_bins[threadIdx.x] = 0u;
  if (threadIdx.x == 0) s_kept = 0u;
  __syncthreads();
  // the reference's list holds every value twice (d > 0): m = 2 cnt, sorted'[k] = sorted[k / 2]
  const long long m = (d == 0) ? cnt : 2LL * cnt;
  const int dup = (d == 0) ? 1 : 2;
  double thr[2];
  const double qs[2] = {q_lo, q_hi};
  for (int t = 0; t < 2; ++t) {
    const double h = (double)(m - 1) * qs[t];
    long long lo = (long long)floor(h);
    if (lo < 0) lo = 0;
    if (lo > m - 1) lo = m - 1;
    const long long hi = (lo + 1 <= m - 1) ? lo + 1 : m - 1;
    const double g = h - (double)lo;
    const unsigned int ka = block_select(s_vals, cnt, (unsigned int)(lo / dup), s_hist, s_bcast);
    const unsigned int kb = block_select(s_vals, cnt, (unsigned int)(hi / dup), s_hist, s_bcast);
    thr[t] = np_lerp((double)ordered_to_float(ka), (double)ordered_to_float(kb), g);
  }
  unsigned int my_kept = 0;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const double v = (double)s_vals[i];
    if (v < thr[1] && v > thr[0]) {
      ++my_kept;
      // np.histogram: edges[b] <= v < edges[b+1], last bin closed on the right
      if (v >= s_edges[0] && v <= s_edges[kHistBins]) {
        int lo = 0, hi = kHistBins;                       // find b with edges[b] <= v < edges[b+1]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (v >= s_edges[mid]) lo = mid; else hi = mid; }
        atomicAdd(&s_bins[lo], 1u);
      }
    }
  }
  my_kept = (unsigned int)warp_sum((unsigned long long)my_kept);
  if ((threadIdx.x & 31) == 0 && my_kept) atomicAdd(&s_kept, my_kept);
  __syncthreads();
  if (threadIdx.x == 0) kept[blockIdx.x] = (long long)s_kept * dup;
  if (threadIdx.x < kHistBins) my_hist[threadIdx.x] = (long long)s_bins[threadIdx.x] * dup;
}

// Sequential chunking over the diagonals (main_local.py:41-55). One warp; the fallback for very
// wide ranges (chunk table beyond 48 KB of shared memory).
__global__ void band_reduce_seq_kernel(const long long* __restrict__ kept, const long long* __restrict__ hist,
                                       int nd, int min_count, double* __restrict__ out) {
  const int lane = threadIdx.x;
  long long pending = 0;
  double bins = 0.0;                                    // lane b holds bin b of the pending histogram
  double ent_sum = 0.0;
  int n_ent = 0;
  for (int i = 0; i < nd; ++i) {
    pending += kept[i];
    if (lane < kHistBins) bins += (double)hist[(long long)i * kHistBins + lane];
    if (pending > min_count) {
      const double tot = warp_sum(lane < kHistBins ? bins : 0.0);
      const double p = bins / tot;                      // tot == 0 -> NaN like the reference
      double term = (lane < kHistBins && bins > 0.0) ? -p * log(p) : 0.0;
      if (tot == 0.0 && lane == 0) term = nan("");
      ent_sum += warp_sum(term);
      ++n_ent;
      pending = 0;
      bins = 0.0;
    }
  }
  if (lane == 0) { out[0] = n_ent ? ent_sum / n_ent : nan(""); out[1] = (double)n_ent; }
}

// The same in two phases (one CTA): only the chunk BOUNDARIES are sequential ("more than
// min_count pending -> emit, reset": an integer walk over the per-diagonal counts staged in shared
// memory); the entropy of each chunk -- at chromosome scale every diagonal is a chunk of its own --
// is a warp's work, 32 chunks at a time. The entropies are summed in a fixed order.
__global__ void __launch_bounds__(1024)
band_reduce_kernel(const long long* __restrict__ kept, const long long* __restrict__ hist,
                   int nd, int min_count, double* __restrict__ out) {
  extern __shared__ int s_tab[];                        // [nd] counts, then [nd] last diagonal of chunk c
  __shared__ int s_nchunks;
  __shared__ double s_part[32];
  int* s_kept = s_tab;
  int* s_end = s_tab + nd;
  for (int i = threadIdx.x; i < nd; i += blockDim.x) s_kept[i] = (int)kept[i];    // <= 2 n
  __syncthreads();
  if (threadIdx.x == 0) {
    long long pending = 0;
    int nc = 0;
    for (int i = 0; i < nd; ++i) {
      pending += s_kept[i];
      if (pending > min_count) { s_end[nc++] = i; pending = 0; }
    }
    s_nchunks = nc;
  }
  __syncthreads();
  const int nc = s_nchunks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double part = 0.0;
  for (int c = warp; c < nc; c += 32) {
    const int i0 = c ? s_end[c - 1] + 1 : 0, i1 = s_end[c];
    double bins = 0.0;
    if (lane < kHistBins)
      for (int i = i0; i <= i1; ++i) bins += (double)hist[(long long)i * kHistBins + lane];
    const double tot = warp_sum(lane < kHistBins ? bins : 0.0);
    const double p = bins / tot;                        // tot == 0 -> NaN like the reference
    double term = (lane < kHistBins && bins > 0.0) ? -p * log(p) : 0.0;
    if (tot == 0.0 && lane == 0) term = nan("");
    part += warp_sum(term);
  }
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0.0;
    for (int w = 0; w < 32; ++w) e += s_part[w];
    out[0] = nc ? e / nc : nan("");
    out[1] = (double)nc;
  }
}

}  // namespace
}  // namespace hia

HIA_API int64_t hia_band_entropy_workspace_bytes(int32_t n_diag) {
  if (n_diag < 0) n_diag = 0;
  return (int64_t)n_diag * (1 + hia::kHistBins) * 8 + 256;
}

HIA_API int hia_band_entropy(const float* adj, int32_t n, int64_t ld, int32_t d_lo, int32_t d_hi,
                             const double* edges30, int32_t min_count, double per, double* out2,
                             void* workspace, int64_t workspace_bytes, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!adj || !edges30 || !out2 || !workspace || n <= 0 || ld < n || d_lo < 0) return HIA_ERR_BAD_ARG;
  const int nd = d_hi - d_lo;
  if (nd <= 0) {                                         // empty range: no entropy chunk -> None
    band_reduce_seq_kernel<<<1, 32, 0, s>>>(nullptr, nullptr, 0, min_count, out2);
    HIA_LAUNCH_CHECK();
    return HIA_OK;
  }
  if (workspace_bytes < hia_band_entropy_workspace_bytes(nd)) return HIA_ERR_WORKSPACE;
  const size_t sh = (size_t)std::max(1, n - d_lo) * sizeof(float);
  if (sh > 200 * 1024) return HIA_ERR_UNSUPPORTED;       // diagonal must fit in shared memory
  long long* kept = static_cast<long long*>(workspace);
  long long* hist = kept + nd;
  HIA_CUDA(cudaFuncSetAttribute(band_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  // np.percentile(vals, per) / (100 - per): quantiles formed like numpy (q / 100 in fp64)
  band_stats_kernel<<<nd, kBandThreads, sh, s>>>(adj, n, ld, d_lo, edges30, per / 100.0,
                                                 (100.0 - per) / 100.0, kept, hist);
  HIA_LAUNCH_CHECK();
  const size_t sh_tab = (size_t)2 * nd * sizeof(int);
  if (sh_tab <= 40 * 1024)
    band_reduce_kernel<<<1, 1024, sh_tab, s>>>(kept, hist, nd, min_count, out2);
  else
    band_reduce_seq_kernel<<<1, 32, 0, s>>>(kept, hist, nd, min_count, out2);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}
