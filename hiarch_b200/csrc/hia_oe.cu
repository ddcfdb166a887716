// a4/a5: expected contacts per diagonal, O/E distance normalisation, log.
//
// Replaces HiCMtx.obs_d_exp(mode='sgd_mean') -- publish/new_hic_class.py:1121-1290
// (_sgd_mean :1182-1216, monotone clamp :1251-1261, per-diagonal divide :1263-1268, inter
// blocks :1272-1282) -- and ArrayHiCMtx.log (:2076-2094).
//
// All three kernels are HBM-bound streaming passes over an fp32 map:
//   intra_diag_kernel  reads the upper triangle of every intra block once   (4 B/cell read)
//   inter_block_kernel reads every upper inter block once                   (4 B/cell read)
//   oe_apply_kernel    reads + writes the full map once                     (8 B/cell)
// i.e. 2 B/cell + 8 B/cell = 10 B/cell of the full symmetric map.
// Reductions accumulate in fp64 (the reference is fp64 throughout).
#include "hia_common.cuh"

namespace hia {
namespace {

constexpr int kDiagThreads = 256;             // one float4 column group per thread
constexpr int kDiagWindow = kDiagThreads * 4; // diagonals covered by one CTA
constexpr int kDiagRowGroups = 32;            // 4-row groups per CTA  (128 rows)

// ---------------------------------------------------------------------------------------
// Per-diagonal sums of one intra block.
//
// Work item = (diagonal window w, row chunk). Thread L owns the 7 diagonals
// D + [-3, 3], D = 4 (L + 256 w).  For the 4-row group rho (global rows 4 rho .. 4 rho + 3) it
// loads the 16-byte aligned float4 at global column 4 (L + 256 w + rho) of each row: the load
// address advances by one float4 per row group, so the diagonals a thread sees stay fixed --
// coalesced 128-bit loads (a warp reads 512 contiguous bytes per row), no shared memory, no
// shuffles in the loop; the 3-diagonal overlap between neighbouring lanes is merged by the
// fp64 atomics at the end.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDiagThreads)
intra_diag_kernel(const float* __restrict__ map, long long ld, const int* __restrict__ chr_start,
                  double* __restrict__ diag_sum, float* __restrict__ diag_minpos, int row_chunks,
                  int row0, int row_end) {
  // `map` points at global row `row0`; only rows [row0, row_end) are visited (row sharding)
  const int c = blockIdx.y;
  const int s = chr_start[c], e = chr_start[c + 1] - 1;   // inclusive bin range
  const int n = e - s + 1;
  if (n <= 0) return;
  const int w = blockIdx.x / row_chunks;
  const int chunk = blockIdx.x % row_chunks;
  // global 4-row groups covering [s, e]
  const int rho_first = s >> 2;
  const int rho_lo = rho_first + chunk * kDiagRowGroups;
  const int rho_hi = min(rho_lo + kDiagRowGroups, (e >> 2) + 1);
  if (rho_lo >= rho_hi) return;
  const int D = 4 * (threadIdx.x + kDiagThreads * w);     // centre diagonal of this thread
  if (D - 3 > n - 1) {
    // whole warp beyond the last diagonal? (D grows with threadIdx) -> nothing to do
    if (4 * ((threadIdx.x & ~31) + kDiagThreads * w) - 3 > n - 1) return;
  }
  double acc[7];
  float mn[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) { acc[k] = 0.0; mn[k] = INFINITY; }

  for (int rho = rho_lo; rho < rho_hi; ++rho) {
    const int gc0 = 4 * rho + D;                           // aligned global column of the float4
    if (gc0 > e) break;                                    // later row groups are further right
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int gr = 4 * rho + rr;
      if (gr < s || gr > e || gr < row0 || gr >= row_end) continue;
      if (gc0 + 3 < gr) continue;                          // float4 entirely below the diagonal
      const float4 v = *reinterpret_cast<const float4*>(map + (long long)(gr - row0) * ld + gc0);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int gc = gc0 + cc;
        if (gc >= gr && gc <= e) {
          const int k = cc - rr + 3;                       // diagonal D + (cc - rr)
          acc[k] += (double)vv[cc];
          if (vv[cc] > 0.f) mn[k] = fminf(mn[k], vv[cc]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int d = D + k - 3;
    if (d >= 0 && d < n) {
      if (acc[k] != 0.0) atomicAdd(&diag_sum[s + d], acc[k]);
      if (mn[k] != INFINITY) atomic_min_pos_float(&diag_minpos[s + d], mn[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Upper inter blocks (a < b): sum, non-zero count, min positive per block.
// CTA = 64 rows x 1024 columns; a thread owns one float4 column group, so the chromosome of its
// 4 columns is fixed; per-tile partials are fp32 (<= 64 addends), folded into fp64 shared
// tables with warp-aggregated atomics, then one global atomic per (tile, chromosome).
// ---------------------------------------------------------------------------------------
constexpr int kInterThreads = 256;
constexpr int kInterCols = kInterThreads * 4;
constexpr int kInterRows = 64;

__global__ void __launch_bounds__(kInterThreads)
inter_block_kernel(const float* __restrict__ map, int n, long long ld,
                   const int* __restrict__ chr_start, int n_chr,
                   const short* __restrict__ bin_chr, double* __restrict__ inter_sum,
                   unsigned long long* __restrict__ inter_nnz, float* __restrict__ inter_minpos,
                   int row0, int row_end, const int* __restrict__ skip_if) {
  extern __shared__ unsigned char smem_raw[];
  if (skip_if && *skip_if) return;                        // the scatter produced these statistics already
  double* s_sum = reinterpret_cast<double*>(smem_raw);
  unsigned long long* s_nnz = reinterpret_cast<unsigned long long*>(s_sum + n_chr);
  float* s_min = reinterpret_cast<float*>(s_nnz + n_chr);

  const int r0 = row0 + blockIdx.y * kInterRows;           // global rows; `map` points at row0
  if (r0 >= row_end) return;
  const int r1 = min(r0 + kInterRows, row_end);
  const int cta_c1 = min(n, (int)(blockIdx.x + 1) * kInterCols);      // exclusive
  int a_prev = bin_chr[r0];
  // the inter part of a row starts at chr_start[a + 1], which only grows with the row index
  if (cta_c1 <= chr_start[a_prev + 1]) return;
  const int c0 = (blockIdx.x * kInterThreads + threadIdx.x) * 4;
  const bool active = c0 < n;
  int b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = active ? (int)bin_chr[min(c0 + j, n - 1)] : 0;
  for (int i = threadIdx.x; i < n_chr; i += blockDim.x) { s_sum[i] = 0.0; s_nnz[i] = 0ull; s_min[i] = INFINITY; }
  __syncthreads();

  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned int cnt[4] = {0u, 0u, 0u, 0u};
  float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};

  auto flush = [&](int a) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b0 = __shfl_sync(0xffffffffu, b[j], 0);
      if (__all_sync(0xffffffffu, b[j] == b0)) {
        const double ws = warp_sum((double)sum[j]);
        const unsigned long long wc = warp_sum((unsigned long long)cnt[j]);
        const float wm = warp_min(mn[j]);
        if ((threadIdx.x & 31) == 0) {
          if (wc | (ws != 0.0)) { atomicAdd(&s_sum[b0], ws); atomicAdd(&s_nnz[b0], wc); }
          if (wm != INFINITY) atomic_min_pos_float(&s_min[b0], wm);
        }
      } else {
        if (cnt[j] | (sum[j] != 0.f)) { atomicAdd(&s_sum[b[j]], (double)sum[j]); atomicAdd(&s_nnz[b[j]], (unsigned long long)cnt[j]); }
        if (mn[j] != INFINITY) atomic_min_pos_float(&s_min[b[j]], mn[j]);
      }
      sum[j] = 0.f; cnt[j] = 0u; mn[j] = INFINITY;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_chr; i += blockDim.x) {
      if (s_nnz[i] | (s_sum[i] != 0.0)) {
        atomicAdd(&inter_sum[a * n_chr + i], s_sum[i]);
        atomicAdd(&inter_nnz[a * n_chr + i], s_nnz[i]);
      }
      if (s_min[i] != INFINITY) atomic_min_pos_float(&inter_minpos[a * n_chr + i], s_min[i]);
      s_sum[i] = 0.0; s_nnz[i] = 0ull; s_min[i] = INFINITY;
    }
    __syncthreads();
  };

  auto accumulate = [&](const float4& v, int col_lo) {
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = c0 + j;
      if (gc >= col_lo && gc < n) {
        sum[j] += vv[j];
        cnt[j] += (vv[j] != 0.f);
        if (vv[j] > 0.f) mn[j] = fminf(mn[j], vv[j]);
      }
    }
  };
  int r = r0;
  while (r < r1) {
    const int a = bin_chr[r];                              // uniform
    if (a != a_prev) { flush(a_prev); a_prev = a; }
    const int col_lo = chr_start[a + 1];
    // rows of chromosome a inside this tile: 4 independent 128-bit loads in flight per thread
    const int r_stop = min(r1, chr_start[a + 1]);
    if (active && c0 + 3 >= col_lo) {
      const float* p = map + (long long)(r - row0) * ld + c0;   // ld % 4 == 0, ld >= n
      if (c0 >= col_lo && c0 + 3 < n) {
        // all four columns are inter-chromosome cells of this row range: no per-element tests, 8 rows in flight
        auto acc4 = [&](const float4& v) {
          sum[0] += v.x; sum[1] += v.y; sum[2] += v.z; sum[3] += v.w;
          cnt[0] += (v.x != 0.f); cnt[1] += (v.y != 0.f); cnt[2] += (v.z != 0.f); cnt[3] += (v.w != 0.f);
          mn[0] = fminf(mn[0], v.x > 0.f ? v.x : INFINITY); mn[1] = fminf(mn[1], v.y > 0.f ? v.y : INFINITY);
          mn[2] = fminf(mn[2], v.z > 0.f ? v.z : INFINITY); mn[3] = fminf(mn[3], v.w > 0.f ? v.w : INFINITY);
        };
        for (; r + 8 <= r_stop; r += 8, p += 8 * ld) {
          float4 v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = ld_stream4(p + q * ld);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc4(v[q]);
        }
        for (; r < r_stop; ++r, p += ld) acc4(ld_stream4(p));
      } else {
        for (; r + 4 <= r_stop; r += 4, p += 4 * ld) {
          const float4 v0 = ld_stream4(p), v1 = ld_stream4(p + ld), v2 = ld_stream4(p + 2 * ld), v3 = ld_stream4(p + 3 * ld);
          accumulate(v0, col_lo); accumulate(v1, col_lo); accumulate(v2, col_lo); accumulate(v3, col_lo);
        }
        for (; r < r_stop; ++r, p += ld) accumulate(ld_stream4(p), col_lo);
      }
    } else {
      r = r_stop;
    }
  }
  flush(a_prev);
}

__global__ void init_expected_kernel(double* diag_sum, float* diag_minpos, int n, double* inter_sum,
                                     unsigned long long* inter_nnz, float* inter_minpos, int cc,
                                     const int* __restrict__ inter_valid) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { diag_sum[i] = 0.0; diag_minpos[i] = INFINITY; }
  // inter_valid (optional): the scatter already produced the inter-chromosome statistics
  if (i < cc && !(inter_valid && *inter_valid)) { inter_sum[i] = 0.0; inter_nnz[i] = 0ull; inter_minpos[i] = INFINITY; }
}

// ---------------------------------------------------------------------------------------
// Finalize: one CTA per chromosome. SGD pooling, un-flushed last group, monotone clamp.
// ---------------------------------------------------------------------------------------
constexpr int kMaxGroups = 512;
constexpr int kFinThreads = 1024;

__global__ void __launch_bounds__(kFinThreads)
finalize_intra_kernel(const double* __restrict__ diag_sum, const float* __restrict__ diag_minpos,
                      const int* __restrict__ chr_start, const int* __restrict__ pool_gid,
                      const int* __restrict__ use_code, int clamp, double* __restrict__ expected,
                      float* __restrict__ inv_expected, float* __restrict__ min_pos_oe) {
  __shared__ double g_sum[kMaxGroups];
  __shared__ double g_cnt[kMaxGroups];
  __shared__ double s_scan[kFinThreads];
  __shared__ double s_carry;
  __shared__ int s_has_carry;
  __shared__ float s_min[32];

  const int c = blockIdx.x;
  const int s = chr_start[c];
  const int n = chr_start[c + 1] - s;
  if (n <= 0) return;
  const int tid = threadIdx.x;
  for (int g = tid; g < kMaxGroups; g += blockDim.x) { g_sum[g] = 0.0; g_cnt[g] = 0.0; }
  __syncthreads();
  // pooled sums: an 'all'-type diagonal d>0 holds both +-d: sum = 2 * upper sum, 2 (n-d) cells
  for (int d = tid; d < n; d += blockDim.x) {
    const int g = pool_gid[s + d];
    if (g >= 0) {
      const double mult = (d == 0) ? 1.0 : 2.0;
      atomicAdd(&g_sum[g], mult * diag_sum[s + d]);
      atomicAdd(&g_cnt[g], mult * (double)(n - d));
    }
  }
  __syncthreads();
  const double diag0_mean = diag_sum[s] / (double)n;
  // n <= 1: M / mean(M); all-zero block: unchanged  (new_hic_class.py:1229-1236)
  // -> both fall out of the generic path: exp[0] = mean(diag) and exp <= 0 is skipped.

  if (tid == 0) { s_carry = 0.0; s_has_carry = 0; }
  __syncthreads();
  float my_min = INFINITY;
  for (int base = 0; base < n; base += blockDim.x) {
    const int d = base + tid;
    double x = 0.0;
    bool valid = d < n;
    if (valid) {
      const int code = use_code[s + d];
      if (code < 0) {
        x = diag0_mean;
      } else {
        const int g = code >> 1;
        x = g_sum[g] / g_cnt[g];
        if ((code & 1) && x == 0.0) x = 1.0;
      }
    }
    double ex = x;
    if (clamp) {
      // exp'[d] = x > pmin ? pmin : x, pmin = min over the non-zero x before d (carry included)
      // exclusive prefix-min over non-zero values: Hillis-Steele in shared memory
      const double big = INFINITY;
      s_scan[tid] = (valid && x != 0.0) ? x : big;
      __syncthreads();
      for (int off = 1; off < blockDim.x; off <<= 1) {
        double other = (tid >= off) ? s_scan[tid - off] : big;
        __syncthreads();
        s_scan[tid] = fmin(s_scan[tid], other);
        __syncthreads();
      }
      double pmin = (tid > 0) ? s_scan[tid - 1] : big;     // exclusive
      if (s_has_carry) pmin = fmin(pmin, s_carry);
      if (valid && pmin != big && x > pmin) ex = pmin;
      __syncthreads();
      if (tid == blockDim.x - 1) {
        double incl = s_scan[tid];
        if (incl != big) {
          s_carry = s_has_carry ? fmin(s_carry, incl) : incl;
          s_has_carry = 1;
        }
      }
      __syncthreads();
    }
    if (valid) {
      expected[s + d] = ex;
      const float inv = (ex > 0.0) ? (float)(1.0 / ex) : 1.0f;   // exp <= 0 -> skipped (:1264)
      inv_expected[s + d] = inv;
      const float mp = diag_minpos[s + d];
      if (mp != INFINITY) my_min = fminf(my_min, mp * inv);
    }
  }
  my_min = warp_min(my_min);
  if ((tid & 31) == 0) s_min[tid >> 5] = my_min;
  __syncthreads();
  if (tid < 32) {
    float v = (tid < (blockDim.x >> 5)) ? s_min[tid] : INFINITY;
    v = warp_min(v);
    if (tid == 0 && v != INFINITY && v > 0.f) atomic_min_pos_float(min_pos_oe, v);
  }
}

__global__ void finalize_inter_kernel(const double* __restrict__ inter_sum,
                                      const unsigned long long* __restrict__ inter_nnz,
                                      const float* __restrict__ inter_minpos, int n_chr,
                                      float* __restrict__ inter_inv, float* __restrict__ min_pos_oe) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_chr * n_chr) return;
  const int a = i / n_chr, b = i % n_chr;
  if (a == b) { inter_inv[i] = 1.f; return; }
  const int u = (a < b) ? i : (b * n_chr + a);                  // stats live in the upper block
  const unsigned long long cnt = inter_nnz[u];
  float inv = 1.f;
  if (cnt > 0) {
    const double mean = inter_sum[u] / (double)cnt;             // mean of non-zero cells (:1278-1280)
    inv = (float)(1.0 / mean);
    const float mp = inter_minpos[u];
    if (a < b && mp != INFINITY) {
      const float x = mp * inv;
      if (x > 0.f) atomic_min_pos_float(min_pos_oe, x);
    }
  }
  inter_inv[i] = inv;
}

__global__ void init_min_kernel(float* p) { *p = INFINITY; }

// ---------------------------------------------------------------------------------------
// Apply: out = map * inv; optional log. One thread per float4, rows over blockIdx.y.
// ---------------------------------------------------------------------------------------
// ln(1 + x) for x > 0 on the SFU: 1 + x is rounded once (<= 6e-8 absolute in the result), lg2.approx
// has <= 2^-21.4 absolute error near 1 and 2 ulp elsewhere -- three orders below the 1e-5 the
// path has to hold, and ~20 instructions cheaper than log1pf (the pass was ALU-bound with it).
// The fill value of the non-positive cells MUST come from the same function: in the reference it is
// bit-identical to the value of the (many) cells that attain the minimum, and the percentile
// stages downstream (strict inequalities on a large tie group) depend on that tie.
__device__ __forceinline__ float fast_log1p_pos(float x) { return __logf(1.f + x); }

__global__ void __launch_bounds__(256)
oe_apply_kernel(const float* __restrict__ map, float* __restrict__ out, int n, long long ld,
                const int* __restrict__ chr_start, int n_chr, const short* __restrict__ bin_chr,
                const float* __restrict__ inv_expected, const float* __restrict__ inter_inv,
                const float* __restrict__ min_pos_oe, int flags, int row0, int row_end) {
  // map / out point at global row `row0`
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= n) return;
  const bool do_log = flags & HIA_OE_LOG;
  const bool only_intra = flags & HIA_OE_ONLY_INTRA;
  float fill = 0.f;
  if (do_log) {
    const float m = *min_pos_oe;
    fill = (m == INFINITY) ? 0.f : fast_log1p_pos(m);
  }
  int cb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cb[j] = bin_chr[min(c0 + j, n - 1)];
  const bool full4 = c0 + 3 < n;
  const bool one_chr = full4 && cb[0] == cb[3];            // the common case: 4 columns, one chromosome

  auto scale = [&](float (&v)[4], int r, int a) {
    if (one_chr && cb[0] != a) {                           // inter block, one factor for the float4
      const float f = only_intra ? 0.f : __ldg(&inter_inv[a * n_chr + cb[0]]);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= f;
    } else {
      const int s = chr_start[a];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (cb[j] == a) v[j] *= __ldg(&inv_expected[s + abs(c0 + j - r)]);
        else v[j] = only_intra ? 0.f : v[j] * __ldg(&inter_inv[a * n_chr + cb[j]]);
      }
    }
    if (do_log) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (v[j] > 0.f) ? fast_log1p_pos(v[j]) : fill;
    }
  };

  int r = row0 + blockIdx.y;
  if (full4) {
    // two rows per iteration: two independent 128-bit loads in flight per thread
    for (; r + (int)gridDim.y < row_end; r += 2 * gridDim.y) {
      const int r2 = r + gridDim.y;
      const float4 t0 = ld_stream4(map + (long long)(r - row0) * ld + c0);
      const float4 t1 = ld_stream4(map + (long long)(r2 - row0) * ld + c0);
      float v0[4] = {t0.x, t0.y, t0.z, t0.w}, v1[4] = {t1.x, t1.y, t1.z, t1.w};
      scale(v0, r, bin_chr[r]);
      scale(v1, r2, bin_chr[r2]);
      st_stream4(out + (long long)(r - row0) * ld + c0, make_float4(v0[0], v0[1], v0[2], v0[3]));
      st_stream4(out + (long long)(r2 - row0) * ld + c0, make_float4(v1[0], v1[1], v1[2], v1[3]));
    }
  }
  for (; r < row_end; r += gridDim.y) {
    const float* src = map + (long long)(r - row0) * ld + c0;
    float v[4];
    if (full4) {
      const float4 t = ld_stream4(src);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (c0 + j < n) ? src[j] : 0.f;
    }
    scale(v, r, bin_chr[r]);
    float* dst = out + (long long)(r - row0) * ld + c0;
    if (full4) {
      st_stream4(dst, make_float4(v[0], v[1], v[2], v[3]));
    } else {
      for (int j = 0; j < 4 && c0 + j < n; ++j) dst[j] = v[j];
    }
  }
}

// The same from the UPPER triangle only (HIA_OE_FROM_UPPER): O/E and log are symmetric functions of
// a symmetric map (the expected vector depends on |c - r|, inter_inv[a][b] == inter_inv[b][a]), so
// a 64 x 64 tile (bi <= bj) of the upper triangle is read once (128-bit loads), its result written
// to out[r][c] and, transposed through shared memory, to out[c][r] (128-bit stores on both sides):
// 2 + 4 instead of 4 + 4 bytes per cell, and the lower triangle of `map` is never read -- the
// scatter can skip its mirror pass (HIA_MTX_TRIU_UPPER). Bit-identical to oe_apply_kernel on a
// mirrored map. (The first version used 32 x 32 tiles with scalar accesses: 2.3 TB/s.)
constexpr int kUT = 64;                       // tile side
constexpr int kUTPad = kUT + 1;               // shared row stride (floats): <= 2-way bank conflicts both ways

__global__ void __launch_bounds__(256)
oe_apply_upper_kernel(const float* __restrict__ map, float* __restrict__ out, int n, long long ld,
                      const int* __restrict__ chr_start, int n_chr, const short* __restrict__ bin_chr,
                      const float* __restrict__ inv_expected, const float* __restrict__ inter_inv,
                      const float* __restrict__ min_pos_oe, int flags) {
  __shared__ float tile[kUT][kUTPad];
  // 1-D grid over the T (T + 1) / 2 tiles of the upper triangle, row-major: tile (bi, bj), bi <= bj
  const int T = (n + kUT - 1) / kUT;
  const long long k = blockIdx.x;
  const float tt = 2.f * (float)T + 1.f;                 // first guess in fp32, corrected by the two loops below
  int bi = (int)((tt - sqrtf(fmaxf(tt * tt - 8.f * (float)k, 0.f))) * 0.5f);
  bi = max(0, min(bi, T - 1));
  while (bi > 0 && (long long)bi * T - (long long)bi * (bi - 1) / 2 > k) --bi;
  while (bi + 1 < T && (long long)(bi + 1) * T - (long long)(bi + 1) * bi / 2 <= k) ++bi;
  const int bj = bi + (int)(k - ((long long)bi * T - (long long)bi * (bi - 1) / 2));
  const bool do_log = flags & HIA_OE_LOG;
  const bool only_intra = flags & HIA_OE_ONLY_INTRA;
  float fill = 0.f;
  if (do_log) {
    const float m = *min_pos_oe;
    fill = (m == INFINITY) ? 0.f : fast_log1p_pos(m);
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 float4 column groups x 16 rows
  const int c0 = bj * kUT + 4 * tx;
  // ---- fast path (CTA-uniform): a full tile strictly above the diagonal whose rows lie in ONE chromosome and
  // whose columns lie in ONE other chromosome -- ~90 % of the tiles of a genome-wide map. One factor for the
  // whole tile, no per-cell chromosome compare / expected-vector look-up / edge predicate: the generic path
  // below spends ~50 instructions per cell on those and is issue-bound at the clock the step runs at. Same
  // arithmetic, bit-identical results.
  if (bi < bj && bj * kUT + kUT <= n) {
    const int ra = bin_chr[bi * kUT], rb = bin_chr[bi * kUT + kUT - 1];
    const int ca = bin_chr[bj * kUT], cbl = bin_chr[bj * kUT + kUT - 1];
    if (ra == rb && ca == cbl && ra != ca) {
      const float f = only_intra ? 0.f : __ldg(&inter_inv[ra * n_chr + ca]);
      const float* src = map + (long long)(bi * kUT + ty) * ld + c0;
      float* dst = out + (long long)(bi * kUT + ty) * ld + c0;
      float4 raw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) raw[k] = ld_stream4(src + (long long)(16 * k) * ld);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v[4] = {raw[k].x * f, raw[k].y * f, raw[k].z * f, raw[k].w * f};
        if (do_log) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = (v[j] > 0.f) ? fast_log1p_pos(v[j]) : fill;
        }
        st_stream4(dst + (long long)(16 * k) * ld, make_float4(v[0], v[1], v[2], v[3]));
#pragma unroll
        for (int j = 0; j < 4; ++j) tile[ty + 16 * k][4 * tx + j] = v[j];
      }
      __syncthreads();
      float* dt = out + (long long)(bj * kUT + ty) * ld + bi * kUT + 4 * tx;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        st_stream4(dt + (long long)(16 * k) * ld, make_float4(tile[4 * tx][ty + 16 * k], tile[4 * tx + 1][ty + 16 * k],
                                                               tile[4 * tx + 2][ty + 16 * k], tile[4 * tx + 3][ty + 16 * k]));
      return;
    }
  }
  int cb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cb[j] = bin_chr[min(c0 + j, n - 1)];
  const bool diag_tile = bi == bj;
  // ---- phase 1: load, scale, log, direct store, stage in shared memory
  float4 raw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {                           // 4 independent 128-bit loads in flight
    const int r = bi * kUT + ty + 16 * k;
    raw[k] = (r < n && c0 < n) ? ld_stream4(map + (long long)r * ld + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = bi * kUT + ty + 16 * k;
    float v[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
    if (r < n && c0 < n) {
      const int a = bin_chr[r];
      const int s = chr_start[a];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        if (c < n && c >= r) {
          if (cb[j] == a) v[j] *= __ldg(&inv_expected[s + (c - r)]);
          else v[j] = only_intra ? 0.f : v[j] * __ldg(&inter_inv[a * n_chr + cb[j]]);
          if (do_log) v[j] = (v[j] > 0.f) ? fast_log1p_pos(v[j]) : fill;
        } else {
          v[j] = 0.f;                                     // below the diagonal / beyond the edge: never used
        }
      }
      float* dst = out + (long long)r * ld + c0;
      if (c0 >= r && c0 + 3 < n) {
        st_stream4(dst, make_float4(v[0], v[1], v[2], v[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + j < n && c0 + j >= r) dst[j] = v[j];
      }
    } else {
      v[0] = v[1] = v[2] = v[3] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[ty + 16 * k][4 * tx + j] = v[j];
  }
  __syncthreads();
  // ---- phase 2: transposed store out[c][r] for the strictly lower cells
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int R = bj * kUT + ty + 16 * k;                 // output row = column of the tile
    const int C0 = bi * kUT + 4 * tx;                     // output columns = rows of the tile
    if (R >= n || C0 >= n) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = tile[4 * tx + j][ty + 16 * k];
    float* dst = out + (long long)R * ld + C0;
    if (!diag_tile) {                                     // R >= (bi + 1) * 64 > every column of the group
      st_stream4(dst, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (C0 + j < R) dst[j] = v[j];
    }
  }
}

// Raw O/E in fp64 (obs_d_exp without log at class level, new_hic_class.py:1263-1282): the fp32 map
// holds 1e-5 absolute only up to ~50; this variant divides the (exact) fp32 counts by the fp64
// expected vector / the fp64 mean of the non-zero cells like the reference and stores doubles.
__global__ void __launch_bounds__(256)
oe_apply_f64_kernel(const float* __restrict__ map, long long ld, double* __restrict__ out, long long ld_out, int n,
                    const int* __restrict__ chr_start, int n_chr, const short* __restrict__ bin_chr,
                    const double* __restrict__ expected, const double* __restrict__ inter_sum,
                    const unsigned long long* __restrict__ inter_nnz, int only_intra) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int b = bin_chr[c];
  for (int r = blockIdx.y; r < n; r += gridDim.y) {
    const int a = bin_chr[r];
    double v = (double)map[(long long)r * ld + c];
    if (a == b) {
      const double ex = expected[chr_start[a] + abs(c - r)];
      if (ex > 0.0) v = v / ex;                                          // exp <= 0: skipped (:1264)
    } else if (only_intra) {
      v = 0.0;
    } else {
      const int u = (a < b) ? a * n_chr + b : b * n_chr + a;            // statistics live in the upper block
      const unsigned long long cnt = inter_nnz[u];
      if (cnt > 0) v = v / (inter_sum[u] / (double)cnt);                // / mean(non-zero) (:1278-1280)
    }
    out[(long long)r * ld_out + c] = v;
  }
}

// ---------------------------------------------------------------------------------------
// Stand-alone ArrayHiCMtx.log (new_hic_class.py:2076-2094) for maps that did not come out of
// hia_oe_apply: pass 1 = min positive value, pass 2 = x>0 ? ln(x+1) : ln(min_pos + 1).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
min_pos_kernel(const float* __restrict__ x, int rows, int cols, long long ld, float* __restrict__ min_pos) {
  float mn = INFINITY;
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = x[(i / cols) * ld + (i % cols)];
    if (v > 0.f) mn = fminf(mn, v);
  }
  mn = warp_min(mn);
  if ((threadIdx.x & 31) == 0 && mn != INFINITY) atomic_min_pos_float(min_pos, mn);
}

__global__ void __launch_bounds__(256)
log_apply_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int cols, long long ld,
                 const float* __restrict__ min_pos) {
  const float m = *min_pos;
  const float fill = (m == INFINITY) ? 0.f : fast_log1p_pos(m);
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long off = (i / cols) * ld + (i % cols);
    const float v = x[off];
    out[off] = (v > 0.f) ? fast_log1p_pos(v) : fill;
  }
}

}  // namespace
}  // namespace hia

HIA_API int hia_log_map(const float* x, float* out, int32_t rows, int32_t cols, int64_t ld,
                        float* min_pos_ws, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !out || !min_pos_ws || rows <= 0 || cols <= 0 || ld < cols) return HIA_ERR_BAD_ARG;
  init_min_kernel<<<1, 1, 0, stream>>>(min_pos_ws);
  HIA_LAUNCH_CHECK();
  const long long total = (long long)rows * cols;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
  min_pos_kernel<<<grid, 256, 0, stream>>>(x, rows, cols, ld, min_pos_ws);
  HIA_LAUNCH_CHECK();
  log_apply_kernel<<<grid, 256, 0, stream>>>(x, out, rows, cols, ld, min_pos_ws);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

static int expected_pass_impl(const int32_t* inter_valid, const float* map, int32_t n, int64_t ld, const int32_t* chr_start,
                              const int32_t* chr_start_host, int32_t n_chr, double* diag_sum,
                              float* diag_minpos, double* inter_sum, unsigned long long* inter_nnz,
                              float* inter_minpos, const int16_t* bin_chr, int row0, int rows,
                              void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!map || !chr_start || !chr_start_host || !diag_sum || !diag_minpos || !inter_sum ||
      !inter_nnz || !inter_minpos || !bin_chr || n <= 0 || n_chr <= 0 || ld < n || row0 < 0 ||
      rows < 0 || row0 + rows > n)
    return HIA_ERR_BAD_ARG;
  if (!aligned16(map) || (ld & 3)) return HIA_ERR_ALIGN;
  if (n_chr > 1024) return HIA_ERR_UNSUPPORTED;
  const int cc = n_chr * n_chr;
  const int cnt = max(n, cc);
  init_expected_kernel<<<(cnt + 255) / 256, 256, 0, stream>>>(diag_sum, diag_minpos, n, inter_sum,
                                                             inter_nnz, inter_minpos, cc, inter_valid);
  HIA_LAUNCH_CHECK();
  if (rows == 0) return HIA_OK;
  int max_len = 0;
  for (int c = 0; c < n_chr; ++c) max_len = max(max_len, chr_start_host[c + 1] - chr_start_host[c]);
  if (max_len <= 0) return HIA_ERR_BAD_ARG;
  // +1 chunk: chromosome starts are not 4-row aligned
  const int row_chunks = (max_len + 3) / 4 / kDiagRowGroups + 2;
  const int windows = (max_len + 3 + kDiagWindow - 1) / kDiagWindow;
  dim3 grid_d(windows * row_chunks, n_chr);
  intra_diag_kernel<<<grid_d, kDiagThreads, 0, stream>>>(map, ld, chr_start, diag_sum, diag_minpos,
                                                         row_chunks, row0, row0 + rows);
  HIA_LAUNCH_CHECK();
  if (n_chr > 1) {
    dim3 grid_i((n + kInterCols - 1) / kInterCols, (rows + kInterRows - 1) / kInterRows);
    const size_t sh = (size_t)n_chr * (8 + 8 + 4);
    inter_block_kernel<<<grid_i, kInterThreads, sh, stream>>>(map, n, ld, chr_start, n_chr, bin_chr,
                                                              inter_sum, inter_nnz, inter_minpos,
                                                              row0, row0 + rows, inter_valid);
    HIA_LAUNCH_CHECK();
  }
  return HIA_OK;
}

HIA_API int hia_oe_expected_pass(const float* map, int32_t n, int64_t ld, const int32_t* chr_start,
                                 const int32_t* chr_start_host, int32_t n_chr, double* diag_sum,
                                 float* diag_minpos, double* inter_sum,
                                 unsigned long long* inter_nnz, float* inter_minpos,
                                 const int16_t* bin_chr, void* stream) {
  return expected_pass_impl(nullptr, map, n, ld, chr_start, chr_start_host, n_chr, diag_sum, diag_minpos,
                            inter_sum, inter_nnz, inter_minpos, bin_chr, 0, n, stream);
}

// The same when the scatter may already have produced the inter-chromosome statistics
// (hia_scatter_*_sym_stats): *inter_valid (device int) != 0 -> they are kept and the pass over the
// inter-chromosome part of the map is skipped (the diagonal statistics are always computed).
HIA_API int hia_oe_expected_pass_ex(const float* map, int32_t n, int64_t ld, const int32_t* chr_start,
                                    const int32_t* chr_start_host, int32_t n_chr, double* diag_sum,
                                    float* diag_minpos, double* inter_sum, unsigned long long* inter_nnz,
                                    float* inter_minpos, const int16_t* bin_chr, const int32_t* inter_valid,
                                    void* stream) {
  return expected_pass_impl(inter_valid, map, n, ld, chr_start, chr_start_host, n_chr, diag_sum, diag_minpos,
                            inter_sum, inter_nnz, inter_minpos, bin_chr, 0, n, stream);
}

// Row-sharded variant: `map_rows` holds global rows [row0, row0 + rows) of the n x n map. The
// outputs are this rank's PARTIAL statistics: the caller sums diag_sum / inter_sum / inter_nnz and
// takes the min of diag_minpos / inter_minpos over the ranks (two small all-reduces), then runs
// hia_oe_finalize (replicated) and hia_oe_apply_rows.
HIA_API int hia_oe_expected_pass_rows(const float* map_rows, int32_t n, int64_t ld, int32_t row0,
                                      int32_t rows, const int32_t* chr_start,
                                      const int32_t* chr_start_host, int32_t n_chr, double* diag_sum,
                                      float* diag_minpos, double* inter_sum,
                                      unsigned long long* inter_nnz, float* inter_minpos,
                                      const int16_t* bin_chr, void* stream) {
  return expected_pass_impl(nullptr, map_rows, n, ld, chr_start, chr_start_host, n_chr, diag_sum, diag_minpos,
                            inter_sum, inter_nnz, inter_minpos, bin_chr, row0, rows, stream);
}

HIA_API int hia_oe_finalize(const double* diag_sum, const float* diag_minpos, const double* inter_sum,
                            const unsigned long long* inter_nnz, const float* inter_minpos,
                            const int32_t* chr_start, int32_t n_chr, int32_t n,
                            const int32_t* pool_gid, const int32_t* use_code, int counts_must_decrese,
                            double* expected, float* inv_expected, float* inter_inv,
                            float* min_pos_oe, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!diag_sum || !diag_minpos || !inter_sum || !inter_nnz || !inter_minpos || !chr_start ||
      !pool_gid || !use_code || !expected || !inv_expected || !inter_inv || !min_pos_oe ||
      n <= 0 || n_chr <= 0)
    return HIA_ERR_BAD_ARG;
  init_min_kernel<<<1, 1, 0, stream>>>(min_pos_oe);
  HIA_LAUNCH_CHECK();
  finalize_intra_kernel<<<n_chr, kFinThreads, 0, stream>>>(diag_sum, diag_minpos, chr_start, pool_gid,
                                                           use_code, counts_must_decrese ? 1 : 0,
                                                           expected, inv_expected, min_pos_oe);
  HIA_LAUNCH_CHECK();
  const int cc = n_chr * n_chr;
  finalize_inter_kernel<<<(cc + 255) / 256, 256, 0, stream>>>(inter_sum, inter_nnz, inter_minpos, n_chr,
                                                              inter_inv, min_pos_oe);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

static int apply_impl(const float* map, float* out, int32_t n, int64_t ld, const int32_t* chr_start,
                      int32_t n_chr, const int16_t* bin_chr, const float* inv_expected,
                      const float* inter_inv, const float* min_pos_oe, int flags, int row0, int rows,
                      void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!map || !out || !chr_start || !bin_chr || !inv_expected || !inter_inv || !min_pos_oe ||
      n <= 0 || n_chr <= 0 || ld < n || row0 < 0 || rows < 0 || row0 + rows > n)
    return HIA_ERR_BAD_ARG;
  if (!aligned16(map) || !aligned16(out) || (ld & 3)) return HIA_ERR_ALIGN;
  if (rows == 0) return HIA_OK;
  if (flags & HIA_OE_FROM_UPPER) {                       // whole maps only; in place is fine (cell by cell)
    if (row0 != 0 || rows != n) return HIA_ERR_BAD_ARG;
    const long long nt = (n + kUT - 1) / kUT;
    const unsigned int gu = (unsigned int)(nt * (nt + 1) / 2);
    oe_apply_upper_kernel<<<gu, 256, 0, stream>>>(map, out, n, ld, chr_start, n_chr, bin_chr, inv_expected,
                                                  inter_inv, min_pos_oe, flags);
    HIA_LAUNCH_CHECK();
    return HIA_OK;
  }
  const int threads = 256;
  dim3 grid((n + threads * 4 - 1) / (threads * 4), min(rows, 2048));
  oe_apply_kernel<<<grid, threads, 0, stream>>>(map, out, n, ld, chr_start, n_chr, bin_chr,
                                                inv_expected, inter_inv, min_pos_oe, flags, row0,
                                                row0 + rows);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

HIA_API int hia_oe_apply(const float* map, float* out, int32_t n, int64_t ld, const int32_t* chr_start,
                         int32_t n_chr, const int16_t* bin_chr, const float* inv_expected,
                         const float* inter_inv, const float* min_pos_oe, int flags, void* stream) {
  return apply_impl(map, out, n, ld, chr_start, n_chr, bin_chr, inv_expected, inter_inv, min_pos_oe,
                    flags, 0, n, stream);
}

HIA_API int hia_oe_apply_rows(const float* map_rows, float* out_rows, int32_t n, int64_t ld, int32_t row0,
                              int32_t rows, const int32_t* chr_start, int32_t n_chr,
                              const int16_t* bin_chr, const float* inv_expected, const float* inter_inv,
                              const float* min_pos_oe, int flags, void* stream) {
  return apply_impl(map_rows, out_rows, n, ld, chr_start, n_chr, bin_chr, inv_expected, inter_inv,
                    min_pos_oe, flags, row0, rows, stream);
}

HIA_API int hia_oe_apply_f64(const float* map, int32_t n, int64_t ld, double* out, int64_t ld_out,
                             const int32_t* chr_start, int32_t n_chr, const int16_t* bin_chr,
                             const double* expected, const double* inter_sum, const int64_t* inter_nnz,
                             int flags, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!map || !out || !chr_start || !bin_chr || !expected || !inter_sum || !inter_nnz || n <= 0 || n_chr <= 0 ||
      ld < n || ld_out < n || (flags & (HIA_OE_LOG | HIA_OE_FROM_UPPER)))
    return HIA_ERR_BAD_ARG;
  dim3 grid((n + 255) / 256, min(n, 1024));
  oe_apply_f64_kernel<<<grid, 256, 0, stream>>>(map, ld, out, ld_out, n, chr_start, n_chr, bin_chr, expected,
                                                inter_sum, reinterpret_cast<const unsigned long long*>(inter_nnz),
                                                (flags & HIA_OE_ONLY_INTRA) ? 1 : 0);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}
