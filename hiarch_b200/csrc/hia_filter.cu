// a7: box (uniform) and median filters with scipy.ndimage 'reflect' boundaries, per block.
//
// Replaces denoise_one_method(mode='mean' | 'median') -- publish/oe_map/S3_denoise.py:48-57 --
// as used by used_denoise (:76-79, median, size = max(2, int(mean_chr_len * .01))) and by
// used_filter (publish/confor/src/S1_preprocess/S1_1_filtering.py:13-16, box, size =
// int(mean_chr_len * .1)).
//
// scipy semantics reproduced: window of output i along an axis = [i - size/2, i - size/2 + size - 1]
// (origin 0, also for even sizes); 'reflect' = (d c b a | a b c d | d c b a), repeated as often
// as needed; uniform_filter = two separable 1-D passes (axis 0 then axis 1) with running sums
// in fp64; median_filter = element of rank (size*size)/2 of the sorted window.
#include "hia_common.cuh"

namespace hia {
namespace {

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}

// ---- axis 0 (down the rows): thread per column, running window sum over a row segment
constexpr int kBoxRowsPerCta = 512;

__global__ void __launch_bounds__(128)
box_axis0_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols,
                 long long ld_in, long long ld_out, int size) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * kBoxRowsPerCta;
  const int r1 = min(rows, r0 + kBoxRowsPerCta);
  const int s1 = size / 2, s2 = size - s1 - 1;
  double sum = 0.0;
  for (int k = -s1; k <= s2; ++k) sum += (double)x[(long long)reflect_idx(r0 + k, rows) * ld_in + c];
  const double inv = (double)size;
  for (int r = r0; r < r1; ++r) {
    out[(long long)r * ld_out + c] = (float)(sum / inv);
    sum += (double)x[(long long)reflect_idx(r + s2 + 1, rows) * ld_in + c] -
           (double)x[(long long)reflect_idx(r - s1, rows) * ld_in + c];
  }
}

// ---- axis 1 (along a row): CTA per (row, 2048-column chunk); fp64 prefix sums in shared memory
constexpr int kBoxChunk = 2048;
constexpr int kBoxThreads = 256;

__global__ void __launch_bounds__(kBoxThreads)
box_axis1_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols,
                 long long ld_in, long long ld_out, int size) {
  extern __shared__ double s_pre[];                      // [chunk + size] inclusive prefix sums
  __shared__ double s_warp[kBoxThreads / 32];
  const int r = blockIdx.y;
  const int c0 = blockIdx.x * kBoxChunk;
  const int nout = min(kBoxChunk, cols - c0);
  if (nout <= 0) return;
  const int s1 = size / 2, s2 = size - s1 - 1;
  const int next = nout + size - 1;                      // extended line: [c0 - s1, c0 + nout - 1 + s2]
  const float* row = x + (long long)r * ld_in;
  // each thread owns a contiguous slice of the extended line
  const int per = (next + kBoxThreads - 1) / kBoxThreads;
  const int b = threadIdx.x * per, e = min(next, b + per);
  double acc = 0.0;
  for (int i = b; i < e; ++i) {
    acc += (double)row[reflect_idx(c0 - s1 + i, cols)];
    s_pre[i] = acc;                                      // local inclusive prefix
  }
  // block exclusive scan of the per-thread totals
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double incl = acc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  double woff = 0.0;
  for (int i = 0; i < w; ++i) woff += s_warp[i];
  const double off = woff + incl - acc;
  for (int i = b; i < e; ++i) s_pre[i] += off;
  __syncthreads();
  const double inv = (double)size;
  for (int j = threadIdx.x; j < nout; j += kBoxThreads) {
    // window = extended indices [j, j + size - 1]
    const double hi = s_pre[j + size - 1];
    const double lo = j > 0 ? s_pre[j - 1] : 0.0;
    out[(long long)r * ld_out + c0 + j] = (float)((hi - lo) / inv);
  }
}

// ---- median: 16x16 output tile per CTA, window in shared memory, bitwise radix select per pixel
constexpr int kMedTile = 16;

__global__ void __launch_bounds__(kMedTile * kMedTile)
median_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols,
              long long ld_in, long long ld_out, int size) {
  extern __shared__ unsigned int s_keys[];               // (tile + size - 1)^2 ordered keys
  const int tw = kMedTile + size - 1;
  const int s1 = size / 2;
  const int r0 = blockIdx.y * kMedTile, c0 = blockIdx.x * kMedTile;
  for (int idx = threadIdx.x; idx < tw * tw; idx += blockDim.x) {
    const int rr = idx / tw, cc = idx % tw;
    const int gr = reflect_idx(r0 - s1 + rr, rows), gc = reflect_idx(c0 - s1 + cc, cols);
    s_keys[idx] = float_to_ordered(x[(long long)gr * ld_in + gc]);
  }
  __syncthreads();
  const int ty = threadIdx.x / kMedTile, tx = threadIdx.x % kMedTile;
  const int r = r0 + ty, c = c0 + tx;
  if (r >= rows || c >= cols) return;
  unsigned int k = (unsigned int)(size * size) / 2;      // rank of the element wanted (ascending)
  unsigned int prefix = 0;
  for (int bit = 31; bit >= 0; --bit) {
    // among the elements that match `prefix` on the bits above `bit`, count those with bit == 0
    const unsigned int himask = (bit == 31) ? 0u : (0xffffffffu << (bit + 1));
    unsigned int cnt0 = 0;
    for (int rr = 0; rr < size; ++rr) {
      const unsigned int* line = s_keys + (ty + rr) * tw + tx;
      for (int cc = 0; cc < size; ++cc) {
        const unsigned int key = line[cc];
        cnt0 += (((key ^ prefix) & himask) == 0u) & (((key >> bit) & 1u) == 0u);
      }
    }
    if (k >= cnt0) { k -= cnt0; prefix |= (1u << bit); }
  }
  out[(long long)r * ld_out + c] = ordered_to_float(prefix);
}

}  // namespace
}  // namespace hia

HIA_API int hia_box_filter_reflect(const float* x, float* out, float* scratch, int32_t rows, int32_t cols,
                                   int64_t ld, int32_t size, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !out || !scratch || rows <= 0 || cols <= 0 || ld < cols || size < 0) return HIA_ERR_BAD_ARG;
  if (scratch == x || scratch == out) return HIA_ERR_BAD_ARG;
  if (size <= 1) {                                       // scipy: size 1 is the identity
    if (out != x)
      HIA_CUDA(cudaMemcpy2DAsync(out, (size_t)ld * 4, x, (size_t)ld * 4, (size_t)cols * 4, rows,
                                 cudaMemcpyDeviceToDevice, s));
    return HIA_OK;
  }
  dim3 g0((cols + 127) / 128, (rows + kBoxRowsPerCta - 1) / kBoxRowsPerCta);
  box_axis0_kernel<<<g0, 128, 0, s>>>(x, scratch, rows, cols, ld, ld, size);
  HIA_LAUNCH_CHECK();
  const size_t sh = (size_t)(kBoxChunk + size) * sizeof(double);
  if (sh > 200 * 1024) return HIA_ERR_UNSUPPORTED;
  HIA_CUDA(cudaFuncSetAttribute(box_axis1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  dim3 g1((cols + kBoxChunk - 1) / kBoxChunk, rows);
  box_axis1_kernel<<<g1, kBoxThreads, sh, s>>>(scratch, out, rows, cols, ld, ld, size);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int hia_median_filter_reflect(const float* x, float* out, int32_t rows, int32_t cols, int64_t ld,
                                      int32_t size, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !out || x == out || rows <= 0 || cols <= 0 || ld < cols || size < 1) return HIA_ERR_BAD_ARG;
  const int tw = kMedTile + size - 1;
  const size_t sh = (size_t)tw * tw * sizeof(unsigned int);
  if (sh > 200 * 1024) return HIA_ERR_UNSUPPORTED;       // size <= ~208
  HIA_CUDA(cudaFuncSetAttribute(median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  dim3 grid((cols + kMedTile - 1) / kMedTile, (rows + kMedTile - 1) / kMedTile);
  median_kernel<<<grid, kMedTile * kMedTile, sh, s>>>(x, out, rows, cols, ld, ld, size);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}
