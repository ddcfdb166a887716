// a8, precision HIA_PREC_FP64: the same row x row similarity with FP64 tensor-core MMA
// (mma.sync.aligned.m8n8k4.f64) and fp64 accumulation. This is the accuracy mode / on-device
// cross-check of the 3xTF32 tcgen05 path; B200's FP64 tensor peak is ~40 TFLOP/s, so it is
// ~8x slower than 3xTF32 and meant for chromosome-sized blocks.
//
// CTA tile 64 x 64, 8 warps (2 x 4), warp tile 32 x 16, k-step 16; operands are normalised
// to fp64 unit rows on the fly from the fp32 map ((x - mean) * inv_norm, stats in fp64).
#include "hia_common.cuh"

namespace hia {
namespace {

constexpr int TM = 64, TN = 64, TK = 16, LDS = 20;   // LDS: padded smem row stride (doubles)

__global__ void __launch_bounds__(256)
row_stats_kernel(const float* __restrict__ x, int m, int k, long long ld_x, int center,
                 double* __restrict__ mean_out, double* __restrict__ inv_out) {
  __shared__ double s_red[8];
  __shared__ double s_val;
  const int r = blockIdx.x;
  if (r >= m) return;
  const float* row = x + (long long)r * ld_x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  double mean = 0.0;
  if (center) {
    double sum = 0.0;
    for (int j = tid; j < k; j += blockDim.x) sum += (double)row[j];
    sum = warp_sum(sum);
    if (lane == 0) s_red[w] = sum;
    __syncthreads();
    if (tid == 0) { double t = 0; for (int i = 0; i < 8; ++i) t += s_red[i]; s_val = t / (double)k; }
    __syncthreads();
    mean = s_val;
    __syncthreads();
  }
  double ss = 0.0;
  for (int j = tid; j < k; j += blockDim.x) { const double d = (double)row[j] - mean; ss += d * d; }
  ss = warp_sum(ss);
  if (lane == 0) s_red[w] = ss;
  __syncthreads();
  if (tid == 0) {
    double t = 0; for (int i = 0; i < 8; ++i) t += s_red[i];
    mean_out[r] = mean;
    inv_out[r] = 1.0 / sqrt(t);
  }
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
syrk_fp64_kernel(const float* __restrict__ x, int m, int k, long long ld_x,
                 const double* __restrict__ mean, const double* __restrict__ inv,
                 float* __restrict__ out, long long ld_out, int flavour) {
  const int bm = blockIdx.y, bn = blockIdx.x;
  if (bn < bm) return;                                   // upper tiles only; lower = mirror
  __shared__ double sA[TM * LDS];
  __shared__ double sB[TN * LDS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;               // 2 x 4 warps
  const int m0 = bm * TM, n0 = bn * TN;

  double acc[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  // loader mapping: 256 threads x 4 elements = 64 rows x 16 k
  const int lr = tid >> 2;            // 0..63
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  const int ga = m0 + lr, gb = n0 + lr;
  const double mean_a = (ga < m) ? mean[ga] : 0.0, inv_a = (ga < m) ? inv[ga] : 0.0;
  const double mean_b = (gb < m) ? mean[gb] : 0.0, inv_b = (gb < m) ? inv[gb] : 0.0;

  for (int k0 = 0; k0 < k; k0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k0 + lk + j;
      double va = 0.0, vb = 0.0;
      if (kk < k) {
        if (ga < m) va = ((double)x[(long long)ga * ld_x + kk] - mean_a) * inv_a;
        if (gb < m) vb = ((double)x[(long long)gb * ld_x + kk] - mean_b) * inv_b;
      }
      sA[lr * LDS + lk + j] = va;
      sB[lr * LDS + lk + j] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < TK; ks += 4) {
      double a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[(wm * 32 + i * 8 + (lane >> 2)) * LDS + ks + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = sB[(wn * 16 + j * 8 + (lane >> 2)) * LDS + ks + (lane & 3)];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma884(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
  // C fragment: row = lane >> 2, cols = 2 (lane & 3) + {0, 1}
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = m0 + wm * 32 + i * 8 + (lane >> 2);
        const int c = n0 + wn * 16 + j * 8 + 2 * (lane & 3) + e;
        if (r >= m || c >= m || c < r) continue;
        double v = acc[i][j][e];
        if (fabs(v) > 1.0) v = copysign(1.0, v);          // scipy pdist clips |cos| to 1
        float o;
        if (flavour != HIA_SIM_PEARSON) o = (c == r) ? 0.f : (float)(1.0 - v);
        else o = (c == r) ? 1.f : (float)v;
        out[(long long)r * ld_out + c] = o;
        if (c != r) out[(long long)c * ld_out + r] = o;
      }
}

}  // namespace

long long similarity_fp64_workspace(int m, int k) { (void)k; return (long long)m * 16 + 256; }

int similarity_fp64(const float* x, int m, int k, long long ld_x, float* out, long long ld_out,
                    int flavour, void* workspace, long long workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < similarity_fp64_workspace(m, k)) return HIA_ERR_WORKSPACE;
  double* mean = static_cast<double*>(workspace);
  double* inv = mean + m;
  row_stats_kernel<<<m, 256, 0, stream>>>(x, m, k, ld_x, flavour != HIA_SIM_COSINE_DISTANCE ? 1 : 0, mean, inv);
  HIA_LAUNCH_CHECK();
  dim3 grid((m + TN - 1) / TN, (m + TM - 1) / TM);
  syrk_fp64_kernel<<<grid, 256, 0, stream>>>(x, m, k, ld_x, mean, inv, out, ld_out, flavour);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

}  // namespace hia
