// a11 / a14 / a15: global-folding reductions.
//
//  row_profile: keep_high_cons (publish/confor/src/S1_preprocess/S1_2_find_anchors/
//      U2_keep_top_cons.py:20-41) fused with get_inter_rowsum (S2_1_find_by_inter_rowsum.py:26-37)
//      and get_intra_rowsum (S2_3_find_by_intra_low.py:18-27): one 4 B/cell pass produces, per row,
//      sum_{inter cols} f(x) and sum_{intra cols} f(x), f(x) = max(x - t, 0) (reverse: max(t - x, 0)),
//      plus the global sum / count of the positive f values (the `new_data /= mean(positives)`
//      normalisation) -- the thresholded map is never materialised.
//  smooth1d: smooth_rowsum (S2_1:42-55): per-chromosome odd box filter, scipy 'reflect'.
//  pool + finish: pool_mtx / norm_mtx / symmetry (publish/confor/src/S2_to_train_mtx/
//      preprocess.py:8-48) and the template scores of to_confor
//      (publish/confor/src/S5_sim_to_pat/to_confor.py:35-47).
#include "hia_common.cuh"
#include "hia_intra_x_core.cuh"

namespace hia {
namespace {

// ------------------------------------------------------------------------------ row profiles
__global__ void __launch_bounds__(256)
row_profile_kernel(const float* __restrict__ x, int n, long long ld, const int* __restrict__ chr_start,
                   const short* __restrict__ bin_chr, const double* __restrict__ thr, int reverse,
                   double* __restrict__ inter_sum, double* __restrict__ intra_sum,
                   double* __restrict__ pos_stats /* [sum, count] */) {
  __shared__ double s_red[8][4];
  const int r = blockIdx.x;
  const int a = bin_chr[r];
  const int s = chr_start[a], e = chr_start[a + 1];      // [s, e)
  const double t = *thr;
  const float* row = x + (long long)r * ld;
  double acc_inter = 0.0, acc_intra = 0.0, acc_pos = 0.0, cnt_pos = 0.0;
  for (int c0 = threadIdx.x * 4; c0 < n; c0 += blockDim.x * 4) {
    float v[4];
    const int nv = min(4, n - c0);
    if (nv == 4) { const float4 q = ld_stream4(row + c0); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    else for (int j = 0; j < 4; ++j) v[j] = (j < nv) ? row[c0 + j] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= nv) break;
      const double d = reverse ? (t - (double)v[j]) : ((double)v[j] - t);
      const double f = d > 0.0 ? d : 0.0;
      const int c = c0 + j;
      if (c >= s && c < e) acc_intra += f; else acc_inter += f;
      if (f > 0.0) { acc_pos += f; cnt_pos += 1.0; }
    }
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double vals[4] = {acc_inter, acc_intra, acc_pos, cnt_pos};
#pragma unroll
  for (int k = 0; k < 4; ++k) { const double q = warp_sum(vals[k]); if (lane == 0) s_red[w][k] = q; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double q = 0; for (int i = 0; i < 8; ++i) q += s_red[i][threadIdx.x];
    if (threadIdx.x == 0) inter_sum[r] = q;
    else if (threadIdx.x == 1) intra_sum[r] = q;
    else atomicAdd(&pos_stats[threadIdx.x - 2], q);
  }
}

// profile = rowsum / n_cols / mean(positives)
__global__ void profile_finish_kernel(const double* __restrict__ inter_sum, const double* __restrict__ intra_sum,
                                      const double* __restrict__ pos_stats, int n,
                                      const int* __restrict__ chr_start, const short* __restrict__ bin_chr,
                                      int norm, double* __restrict__ inter_prof, double* __restrict__ intra_prof) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int a = bin_chr[r];
  const int len = chr_start[a + 1] - chr_start[a];
  const double mean_pos = norm ? pos_stats[0] / pos_stats[1] : 1.0;
  const int other = n - len;
  inter_prof[r] = other > 0 ? inter_sum[r] / mean_pos / (double)other : 0.0;
  intra_prof[r] = intra_sum[r] / mean_pos / (double)len;
}

// scipy.ndimage 'reflect' (d c b a | a b c d | d c b a), any distance
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}

__global__ void smooth1d_kernel(const double* __restrict__ in, double* __restrict__ out, int n,
                                const int* __restrict__ chr_start, const short* __restrict__ bin_chr,
                                double kernal_per, int min_kernal_len) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int a = bin_chr[r];
  const int s = chr_start[a], len = chr_start[a + 1] - s;
  int kl = (int)((double)len * kernal_per);              // get_kernal_len (S2_1:15-20)
  if (kl < min_kernal_len) kl = min_kernal_len;
  kl = (kl / 2) * 2 + 1;
  const double w = 1.0 / (double)kl;
  const int h = kl / 2;
  double acc = 0.0;
  for (int o = -h; o <= h; ++o) acc += w * in[s + reflect_idx(r - s + o, len)];
  out[r] = acc;
}

// ------------------------------------------------------------------------------ GF_S2 pooling
struct PairDesc { int r0, n1, c0, n2, cen1, cen2, intra, pad; };

// adaptive-average-pool cell range: [floor(i*n/k), ceil((i+1)*n/k))
__device__ __forceinline__ void pool_range(int i, int n, int k, int& lo, int& hi) {
  lo = (int)(((long long)i * n) / k);
  hi = (int)(((long long)(i + 1) * n + k - 1) / k);
}

__global__ void __launch_bounds__(256)
pool_kernel(const float* __restrict__ x, long long ld, const PairDesc* __restrict__ pairs,
            double* __restrict__ pooled /* [P][256] */) {
  __shared__ double s_red[8];
  const PairDesc p = pairs[blockIdx.y];
  const int cell = blockIdx.x, i = cell >> 4, j = cell & 15;
  // pool_mtx: whole-block 16x16 unless a 'used' anchor lies strictly inside on both axes
  const bool split = p.cen1 > 0 && p.cen1 < p.n1 - 1 && p.cen2 > 0 && p.cen2 < p.n2 - 1;
  int r_lo, r_hi, c_lo, c_hi;
  if (!split) {
    pool_range(i, p.n1, 16, r_lo, r_hi);
    pool_range(j, p.n2, 16, c_lo, c_hi);
  } else {
    const int qi = i >> 3, ii = i & 7, qj = j >> 3, jj = j & 7;
    const int rb = qi ? p.cen1 : 0, rl = qi ? p.n1 - p.cen1 : p.cen1;
    const int cb = qj ? p.cen2 : 0, cl = qj ? p.n2 - p.cen2 : p.cen2;
    pool_range(ii, rl, 8, r_lo, r_hi); r_lo += rb; r_hi += rb;
    pool_range(jj, cl, 8, c_lo, c_hi); c_lo += cb; c_hi += cb;
  }
  const int w = c_hi - c_lo, hgt = r_hi - r_lo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double acc = 0.0;
  // warp per row of the cell: scalar head up to the first 16-byte boundary, 128-bit body, scalar
  // tail -- no integer division, 16 B per lane per load in flight
  for (int rr = warp; rr < hgt; rr += 8) {
    const float* rp = x + (long long)(p.r0 + r_lo + rr) * ld + p.c0 + c_lo;
    const int mis = (int)((reinterpret_cast<uintptr_t>(rp) >> 2) & 3u);
    int head = (4 - mis) & 3;
    if (head > w) head = w;
    const int nvec = (w - head) >> 2;
    const int tail0 = head + 4 * nvec;
    if (lane < head) acc += (double)rp[lane];
    const float4* vp = reinterpret_cast<const float4*>(rp + head);
    for (int q = lane; q < nvec; q += 32) {
      const float4 t = __ldg(vp + q);
      acc += ((double)t.x + (double)t.y) + ((double)t.z + (double)t.w);
    }
    if (tail0 + lane < w) acc += (double)rp[tail0 + lane];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double q = 0; for (int k = 0; k < 8; ++k) q += s_red[k];
    pooled[(long long)blockIdx.y * 256 + cell] = q / ((double)w * (double)hgt);
  }
}

// norm_mtx (population std) -> symmetry -> template scores; one CTA (256 threads) per pair.
__global__ void __launch_bounds__(256)
gf_finish_kernel(const double* __restrict__ pooled, const PairDesc* __restrict__ pairs,
                 const double* __restrict__ tmpl_intra, const double* __restrict__ tmpl_inter,
                 double* __restrict__ isym /* [P][256] */, double* __restrict__ xs /* [P][256] */,
                 double* __restrict__ scores /* [P][5] */) {
  __shared__ double s_img[256];
  __shared__ double s_sym[256];
  __shared__ double s_red[8];
  __shared__ double s_stat[2];
  const int pidx = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const PairDesc p = pairs[pidx];
  const double v = pooled[(long long)pidx * 256 + t];
  double q = warp_sum(v);
  if (lane == 0) s_red[w] = q;
  __syncthreads();
  if (t == 0) { double a = 0; for (int k = 0; k < 8; ++k) a += s_red[k]; s_stat[0] = a / 256.0; }
  __syncthreads();
  const double mean = s_stat[0];
  const double d = v - mean;
  q = warp_sum(d * d);
  __syncthreads();
  if (lane == 0) s_red[w] = q;
  __syncthreads();
  if (t == 0) { double a = 0; for (int k = 0; k < 8; ++k) a += s_red[k]; s_stat[1] = sqrt(a / 256.0); }
  __syncthreads();
  const double sd = s_stat[1];
  const double z = sd > 0.0 ? d / sd : d;
  s_img[t] = z;
  isym[(long long)pidx * 256 + t] = z;
  __syncthreads();
  const int i = t >> 4, j = t & 15;
  double sym;
  if (p.intra) {
    sym = (z + s_img[(15 - i) * 16 + (15 - j)]) / 2.0;
  } else {
    sym = z;                                             // x + flipud + fliplr + flipud(fliplr), / 4
    sym += s_img[(15 - i) * 16 + j];
    sym += s_img[i * 16 + (15 - j)];
    sym += s_img[(15 - i) * 16 + (15 - j)];
    sym /= 4.0;
  }
  s_sym[t] = sym;
  xs[(long long)pidx * 256 + t] = sym;
  __syncthreads();
  const double* tm = p.intra ? tmpl_intra : tmpl_inter;
  for (int k = 0; k < 5; ++k) {
    double a = warp_sum(sym * tm[k * 256 + t]);
    double b = p.intra ? 0.0 : warp_sum(s_sym[j * 16 + i] * tm[k * 256 + t]);   // transposed image
    __syncthreads();
    if (lane == 0) { s_red[w] = a; }
    __syncthreads();
    double sa = 0;
    if (t == 0) for (int kk = 0; kk < 8; ++kk) sa += s_red[kk];
    __syncthreads();
    if (lane == 0) { s_red[w] = b; }
    __syncthreads();
    if (t == 0) {
      double sb = 0; for (int kk = 0; kk < 8; ++kk) sb += s_red[kk];
      scores[(long long)pidx * 5 + k] = (!p.intra && sa < sb) ? sb : sa;       // max with transposed
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ a13 intra-x
// get_intra_x (publish/confor/src/S1_preprocess/S1_2_find_anchors/S2_2_find_by_intra_x.py:25-50,
// :116-126): for every candidate centre c the response sum(M * K_c) / sum(K_c), with
//   K_c[y][x] = max(0, (k_len - dist((x, y), line through (0, n) and (c, c))) / k_len)
// evaluated on the lower triangle (y >= x) and mirrored (np.tril(K).T + np.tril(K, -1)).
// The reference builds the dense n x n kernel per centre (O(n^3)); here each centre only
// visits the band of width 2 k_len around its line: CTA per centre, warp per row, lanes over
// the (conservative) column interval, exact per-cell weight in fp64.
__global__ void __launch_bounds__(256)
intra_x_kernel(const float* __restrict__ m, int n, long long ld, int k_len, double* __restrict__ out) {
  __shared__ double s_red[8][2];
  const int c = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vertical = (c == 0);                        // cen_point[0] == telo_point[0]
  double k = 0.0, nrm = 1.0;
  if (!vertical) { k = ((double)c - (double)n) / ((double)c - 0.0); nrm = sqrt(1.0 + k * k); }
  const double b0 = (double)n;
  const double inv_klen = 1.0 / (double)k_len;
  double num = 0.0, den = 0.0;
  for (int y = warp; y < n; y += 8) {
    int x_lo, x_hi;                                      // conservative inclusive column interval
    if (vertical) { x_lo = 0; x_hi = min(y, k_len); }
    else {
      // |k x - y + n| < k_len * nrm  <=>  x between (y - n -+ W) / k ; k < 0 for c < n
      const double W = (double)k_len * nrm;
      if (k == 0.0) { x_lo = 0; x_hi = y; }
      else {
        double xa = ((double)y - b0 - W) / k, xb = ((double)y - b0 + W) / k;
        if (xa > xb) { const double t = xa; xa = xb; xb = t; }
        x_lo = max(0, (int)floor(xa) - 1);
        x_hi = min(y, (int)ceil(xb) + 1);
      }
    }
    for (int x = x_lo + lane; x <= x_hi; x += 32) {
      const double dis = vertical ? fabs((double)x - (double)c)
                                  : fabs(k * (double)x - (double)y + b0) / nrm;
      double w = inv_klen * ((double)k_len - dis);
      if (w <= 0.0) continue;
      const double lower = (double)m[(long long)y * ld + x];
      if (x == y) { num += lower * w; den += w; }
      else {
        const double upper = (double)m[(long long)x * ld + y];   // mirrored kernel cell
        num += (lower + upper) * w;
        den += 2.0 * w;
      }
    }
  }
  num = warp_sum(num);
  den = warp_sum(den);
  if (lane == 0) { s_red[warp][0] = num; s_red[warp][1] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, d = 0;
    for (int i = 0; i < 8; ++i) { a += s_red[i][0]; d += s_red[i][1]; }
    out[c] = a / d;
  }
}


// ------------------------------------------------------------------ a13 intra-x, prefix-sum path
// The same response in O(n^2): along a row of M the kernel of every centre is a tent (two linear
// pieces) in the column index, both for the lower cells (q <= r) and for the mirrored upper cells
// (q > r) -- hia_intra_x_core.cuh. Rows are processed in panels of R rows:
//   ix_prefix_kernel : one CTA per row, block scan of (v, q v) in fp64 -> P[row][0..n] (16 B / cell)
//   ix_eval_kernel   : thread per centre x sub-chunk of 32 rows, <= 8 prefix look-ups per (row, centre),
//                      register accumulators -> partial[sub][c]
//   ix_accum_kernel  : fixed-order sum of the partials into acc[c]; the last panel divides.
// Deterministic (no atomics).
__global__ void __launch_bounds__(256)
ix_prefix_kernel(const float* __restrict__ m, int n, long long ld, int r0, IxPrefix* __restrict__ P) {
  __shared__ float s_in[2][1024];
  __shared__ double s_w[2][8][2];
  const float* row = m + (long long)(r0 + blockIdx.x) * ld;
  IxPrefix* out = P + (size_t)blockIdx.x * (size_t)(n + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { IxPrefix z; z.s0 = 0.0; z.s1 = 0.0; out[0] = z; }
  double c0 = 0.0, c1 = 0.0;                             // carry of the previous tiles (same in every thread)
  int buf = 0;
  for (int tile = 0; tile < n; tile += 1024, buf ^= 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = tile + j * 256 + tid;                // coalesced scalar loads: block views need no alignment
      s_in[buf][j * 256 + tid] = q < n ? row[q] : 0.f;
    }
    __syncthreads();
    const int q0 = tile + 4 * tid;
    double a0[4], a1[4], t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double v = (double)s_in[buf][4 * tid + j];
      t0 += v; t1 += (double)(q0 + j) * v;
      a0[j] = t0; a1[j] = t1;
    }
    double w0 = t0, w1 = t1;                             // warp inclusive scan of the thread totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double y0 = __shfl_up_sync(0xffffffffu, w0, o), y1 = __shfl_up_sync(0xffffffffu, w1, o);
      if (lane >= o) { w0 += y0; w1 += y1; }
    }
    if (lane == 31) { s_w[buf][warp][0] = w0; s_w[buf][warp][1] = w1; }
    double e0 = __shfl_up_sync(0xffffffffu, w0, 1), e1 = __shfl_up_sync(0xffffffffu, w1, 1);
    if (lane == 0) { e0 = 0.0; e1 = 0.0; }               // exclusive within the warp
    __syncthreads();
    double b0 = c0, b1 = c1, tot0 = 0.0, tot1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double x0 = s_w[buf][k][0], x1 = s_w[buf][k][1];
      if (k < warp) { b0 += x0; b1 += x1; }
      tot0 += x0; tot1 += x1;
    }
    b0 += e0; b1 += e1;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q0 + j < n) { IxPrefix p; p.s0 = b0 + a0[j]; p.s1 = b1 + a1[j]; out[q0 + j + 1] = p; }
    c0 += tot0; c1 += tot1;
  }
}

__global__ void __launch_bounds__(128)
ix_eval_kernel(const IxPrefix* __restrict__ P, int n, int k_len, int r0, int rows, int rows_per_sub,
               IxPrefix* __restrict__ partial) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= n) return;
  const IxCentre cc = ix_centre(c, n, k_len);
  const int lo = blockIdx.y * rows_per_sub;
  const int hi = min(rows, lo + rows_per_sub);
  double num = 0.0, den = 0.0;
#pragma unroll 2
  for (int i = lo; i < hi; ++i)
    ix_row_contrib(cc, n, k_len, r0 + i, P + (size_t)i * (size_t)(n + 1), num, den);
  IxPrefix p; p.s0 = num; p.s1 = den;
  partial[(size_t)blockIdx.y * n + c] = p;
}

__global__ void ix_accum_kernel(const IxPrefix* __restrict__ partial, int n, int subs, IxPrefix* __restrict__ acc,
                                int first, int last, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double num = 0.0, den = 0.0;
  if (!first) { const IxPrefix a = acc[c]; num = a.s0; den = a.s1; }
  for (int k = 0; k < subs; ++k) { const IxPrefix p = partial[(size_t)k * n + c]; num += p.s0; den += p.s1; }
  IxPrefix a; a.s0 = num; a.s1 = den;
  acc[c] = a;
  if (last) out[c] = num / den;
}

constexpr int kIxRowsPerSub = 32;
// rows per panel: as many as 512 MB of prefix sums hold, 32 <= R <= 1024, multiple of 32. (A panel
// small enough to stay in L2 -- 224 rows at 25 k bins -- left the scan with 1.5 CTAs per SM and
// both kernels latency-bound: 114 x (54 + 76) us measured; 1024 rows keep every SM busy and the
// 2 x 16 B/cell of DRAM traffic cost less than the idle time did.)
inline int ix_panel_rows(int n) {
  long long r = (512ll << 20) / (16ll * ((long long)n + 1));
  r = (r / 32) * 32;
  if (r < 32) r = 32;
  if (r > 1024) r = 1024;
  const long long n32 = (((long long)n + 31) / 32) * 32;
  return (int)(r < n32 ? r : n32);
}

}  // namespace
}  // namespace hia

HIA_API int hia_intra_x_response(const float* block, int32_t n, int64_t ld, int32_t k_len, double* out,
                                 void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!block || !out || n <= 0 || ld < n || k_len <= 0) return HIA_ERR_BAD_ARG;
  intra_x_kernel<<<n, 256, 0, s>>>(block, n, ld, k_len, out);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int64_t hia_intra_x_workspace_bytes(int32_t n) {
  using namespace hia;
  if (n <= 0) return 0;
  const int64_t R = ix_panel_rows(n), subs = (R + kIxRowsPerSub - 1) / kIxRowsPerSub;
  return 16 * (R * ((int64_t)n + 1) + subs * (int64_t)n + (int64_t)n) + 256;
}

HIA_API int hia_intra_x_response_prefix(const float* block, int32_t n, int64_t ld, int32_t k_len, double* out,
                                        void* workspace, int64_t workspace_bytes, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!block || !out || !workspace || n <= 0 || ld < n || k_len <= 0) return HIA_ERR_BAD_ARG;
  if (workspace_bytes < hia_intra_x_workspace_bytes(n)) return HIA_ERR_WORKSPACE;
  if (!aligned16(workspace)) return HIA_ERR_ALIGN;
  const int R = ix_panel_rows(n);
  const int max_subs = (R + kIxRowsPerSub - 1) / kIxRowsPerSub;
  IxPrefix* P = static_cast<IxPrefix*>(workspace);
  IxPrefix* partial = P + (size_t)R * (size_t)(n + 1);
  IxPrefix* acc = partial + (size_t)max_subs * (size_t)n;
  for (int r0 = 0; r0 < n; r0 += R) {
    const int rows = n - r0 < R ? n - r0 : R;
    const int subs = (rows + kIxRowsPerSub - 1) / kIxRowsPerSub;
    ix_prefix_kernel<<<rows, 256, 0, s>>>(block, n, ld, r0, P);
    HIA_LAUNCH_CHECK();
    ix_eval_kernel<<<dim3((n + 127) / 128, subs), 128, 0, s>>>(P, n, k_len, r0, rows, kIxRowsPerSub, partial);
    HIA_LAUNCH_CHECK();
    ix_accum_kernel<<<(n + 255) / 256, 256, 0, s>>>(partial, n, subs, acc, r0 == 0, r0 + R >= n, out);
    HIA_LAUNCH_CHECK();
  }
  return HIA_OK;
}

HIA_API int hia_row_profiles(const float* x, int32_t n, int64_t ld, const int32_t* chr_start,
                             const int16_t* bin_chr, const double* threshold, int reverse, int norm,
                             double* inter_profile, double* intra_profile, double* workspace,
                             void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !chr_start || !bin_chr || !threshold || !inter_profile || !intra_profile || !workspace ||
      n <= 0 || ld < n)
    return HIA_ERR_BAD_ARG;
  if (!aligned16(x) || (ld & 3)) return HIA_ERR_ALIGN;
  double* inter_sum = workspace;             // [n]
  double* intra_sum = workspace + n;         // [n]
  double* pos_stats = workspace + 2 * (int64_t)n;   // [2]
  HIA_CUDA(cudaMemsetAsync(pos_stats, 0, 2 * sizeof(double), s));
  row_profile_kernel<<<n, 256, 0, s>>>(x, n, ld, chr_start, bin_chr, threshold, reverse, inter_sum,
                                       intra_sum, pos_stats);
  HIA_LAUNCH_CHECK();
  profile_finish_kernel<<<(n + 255) / 256, 256, 0, s>>>(inter_sum, intra_sum, pos_stats, n, chr_start,
                                                        bin_chr, norm, inter_profile, intra_profile);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int hia_smooth1d_reflect(const double* in, double* out, int32_t n, const int32_t* chr_start,
                                 const int16_t* bin_chr, double kernal_per, int32_t min_kernal_len,
                                 void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!in || !out || in == out || !chr_start || !bin_chr || n <= 0) return HIA_ERR_BAD_ARG;
  smooth1d_kernel<<<(n + 255) / 256, 256, 0, s>>>(in, out, n, chr_start, bin_chr, kernal_per, min_kernal_len);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int hia_gf_pool_scores(const float* x, int64_t ld, const int32_t* pairs8, int32_t n_pairs,
                               const double* tmpl_intra, const double* tmpl_inter, double* pooled,
                               double* isym, double* xs, double* scores, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!x || !pairs8 || !tmpl_intra || !tmpl_inter || !pooled || !isym || !xs || !scores || n_pairs <= 0)
    return HIA_ERR_BAD_ARG;
  const PairDesc* pairs = reinterpret_cast<const PairDesc*>(pairs8);
  pool_kernel<<<dim3(256, n_pairs), 256, 0, s>>>(x, ld, pairs, pooled);
  HIA_LAUNCH_CHECK();
  gf_finish_kernel<<<n_pairs, 256, 0, s>>>(pooled, pairs, tmpl_intra, tmpl_inter, isym, xs, scores);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

// Scores of P pooled images against CALLER-SUPPLIED templates (to_confor with a user's .mat,
// publish/confor/src/S5_sim_to_pat/to_confor.py:35-47): scores[p][t] = xs[p] . tmpl[t]; do_trans:
// max with the score of the transposed 16 x 16 image (inter-chromosome pairs, :42-46). One warp per
// (image, template); fp64 like the reference. (Round 1 took this branch through torch / cuBLAS.)
namespace hia {
namespace {
__global__ void __launch_bounds__(256)
template_scores_kernel(const double* __restrict__ xs, const double* __restrict__ tmpl, int n_img, int n_tmpl,
                       int do_trans, double* __restrict__ scores) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_img * n_tmpl) return;
  const int p = w / n_tmpl, t = w % n_tmpl;
  const double* x = xs + (long long)p * 256;
  const double* tm = tmpl + (long long)t * 256;
  double a = 0.0, b = 0.0;
  for (int i = lane; i < 256; i += 32) {
    const double tv = tm[i];
    a += x[i] * tv;
    b += x[(i & 15) * 16 + (i >> 4)] * tv;                 // element i of the transposed image
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) scores[(long long)p * n_tmpl + t] = do_trans ? fmax(a, b) : a;
}
}  // namespace
}  // namespace hia

HIA_API int hia_template_scores(const double* xs, int32_t n_img, const double* tmpl, int32_t n_tmpl, int do_trans,
                                double* scores, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!xs || !tmpl || !scores || n_img <= 0 || n_tmpl <= 0) return HIA_ERR_BAD_ARG;
  const long long warps = (long long)n_img * n_tmpl;
  template_scores_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, s>>>(xs, tmpl, n_img, n_tmpl, do_trans ? 1 : 0, scores);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}
