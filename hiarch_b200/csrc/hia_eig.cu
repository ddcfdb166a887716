// a9: top-k eigenvectors of a dense symmetric fp32 matrix (the Pearson map) -- the
// compartment / checkerboard-sign extraction of the north-star. There is NO eigen-decomposition
// in the reference: parity is against numpy.linalg.eigh on the
// identical matrix ("parity unpinned by the reference").
//
// Method: block Lanczos (block Krylov subspace with full re-orthogonalisation, CGS2 + CholQR2),
// Rayleigh-Ritz on the projected block-tridiagonal matrix with a one-sided Jacobi solver that
// runs in a single CTA -- everything stays on the device, no host round trip.
//   Y = C * Q_j          symm_panel_kernel: the only pass over C (4 B/cell, HBM-bound);
//                        fp32 C x fp32 panel, fp32 per-lane partials folded in fp64
//   H[:,j] = Q^T Y ; Y -= Q H[:,j]   (twice)      gram_kernel / panel_sub_kernel   (fp64)
//   Q_{j+1} R = Y        CholQR twice                                               (fp64)
//   T = sym(H) -> Jacobi -> Ritz vectors U = Q Z[:, top k]; residuals ||C u - lambda u||
// Panels are stored column-major (each vector contiguous, ld = n).
#include "hia_common.cuh"
#include <cooperative_groups.h>
#include <cuda.h>
#include <stdlib.h>

namespace hia {
namespace {
namespace cg = cooperative_groups;

constexpr int kMaxBlock = 16;
constexpr int kSymmCols = 1024;       // columns of C per shared-memory V chunk (2048 measured slower: 844 vs 815 us)
constexpr int kSymmRowsPerWarp = 4;
constexpr int kSymmWarps = 8;
constexpr int kSymmRows = kSymmRowsPerWarp * kSymmWarps;   // 32 rows per CTA

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void random_panel_kernel(double* __restrict__ q, int n, int b, unsigned long long seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * b) return;
  const uint32_t h = hash32((uint32_t)i * 2654435761u + (uint32_t)seed) ^ hash32((uint32_t)(i >> 32) + (uint32_t)(seed >> 32) + 0x9e3779b9u);
  q[i] = (double)hash32(h) * (2.0 / 4294967296.0) - 1.0;
}

__global__ void float_to_double_panel_kernel(const float* __restrict__ src, double* __restrict__ dst, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = (double)src[i];
}

__global__ void double_to_float_panel_kernel(const double* __restrict__ src, float* __restrict__ dst, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = (float)src[i];
}

// packed fp32 FMA (Blackwell FFMA2): two lanes of a 64-bit register pair per instruction
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// Sum NACC (32 or 16) per-lane partials across the 32 lanes with a TRANSPOSING butterfly: at every
// step a lane hands half of its values to its partner and keeps the other half, so the whole
// reduction costs NACC - 1 shuffles (a butterfly per value costs 5 NACC) and lane l ends up with
// the complete sum of value (l % NACC).
template <int NACC>
__device__ __forceinline__ float lane_transpose_sum(float (&v)[NACC], int lane) {
  static_assert(NACC == 32 || NACC == 16, "4 rows x block of 8 or 4");
  if (NACC == 16) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
  }
#pragma unroll
  for (int off = NACC / 2; off >= 1; off >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Y[n x B] (fp64, col-major, += into a zeroed panel) = C[n x n] (fp32 row-major) * V[n x B] (fp32).
// CTA = 1024-column chunk x 256 rows: the V chunk (B x 1024 fp32) is staged in shared memory ONCE
// per CTA and reused by 256 rows; every warp streams its 32 rows in groups of 4 with 128-bit
// loads, fp32 partials over the chunk, butterfly reduce, one fp64 atomicAdd per (row, vector).
constexpr int kSymmRowsPerCta = 256;
template <int B>
__global__ void __launch_bounds__(kSymmWarps * 32)
symm_panel_kernel(const float* __restrict__ c, int rows, int n, long long ld, const float* __restrict__ v,
                  double* __restrict__ y, int row_major) {
  // c: rows x n block of the matrix (rows <= n); v: n x B panel (col-major, ld n);
  // y: rows x B, row-major (row_major != 0) or col-major with ld = rows
  extern __shared__ __align__(16) unsigned char s_symm_raw[];
  float (*s_v)[kSymmCols] = reinterpret_cast<float (*)[kSymmCols]>(s_symm_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kSymmCols;
  for (int idx = threadIdx.x; idx < B * kSymmCols; idx += blockDim.x) {
    const int j = idx / kSymmCols, col = idx % kSymmCols;
    s_v[j][col] = (k0 + col < n) ? v[(long long)j * n + k0 + col] : 0.f;
  }
  __syncthreads();
  const int rbase = blockIdx.y * kSymmRowsPerCta + warp * (kSymmRowsPerCta / kSymmWarps);
#pragma unroll 1
  for (int g = 0; g < kSymmRowsPerCta / kSymmWarps / kSymmRowsPerWarp; ++g) {
    const int r0 = rbase + g * kSymmRowsPerWarp;
    if (r0 >= rows) break;
    // (even-column, odd-column) partial sums: the pairs (a.x, a.y) / (v.x, v.y) are adjacent registers
    // of the 128-bit loads, so one FFMA2 does two of the four products (the pass was issue-bound)
    float2 acc[kSymmRowsPerWarp][B];
#pragma unroll
    for (int r = 0; r < kSymmRowsPerWarp; ++r)
#pragma unroll
      for (int j = 0; j < B; ++j) acc[r][j] = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int cc = lane * 4; cc < kSymmCols; cc += 128) {
      const int gc = k0 + cc;
      if (gc >= n) break;
      float4 a[kSymmRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kSymmRowsPerWarp; ++r) {
        const int gr = r0 + r;
        if (gr < rows) {
          if (gc + 3 < n) a[r] = ld_stream4(c + (long long)gr * ld + gc);
          else {
            const float* p = c + (long long)gr * ld + gc;
            a[r].x = p[0];
            a[r].y = (gc + 1 < n) ? p[1] : 0.f;
            a[r].z = (gc + 2 < n) ? p[2] : 0.f;
            a[r].w = 0.f;
          }
        } else a[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const float4 vv = *reinterpret_cast<const float4*>(&s_v[j][cc]);
#pragma unroll
        for (int r = 0; r < kSymmRowsPerWarp; ++r) {
          acc[r][j] = ffma2(make_float2(a[r].x, a[r].y), make_float2(vv.x, vv.y), acc[r][j]);
          acc[r][j] = ffma2(make_float2(a[r].z, a[r].w), make_float2(vv.z, vv.w), acc[r][j]);
        }
      }
    }
    float part[kSymmRowsPerWarp * B];
#pragma unroll
    for (int r = 0; r < kSymmRowsPerWarp; ++r)
#pragma unroll
      for (int j = 0; j < B; ++j) part[r * B + j] = acc[r][j].x + acc[r][j].y;
    const float total = lane_transpose_sum<kSymmRowsPerWarp * B>(part, lane);
    {
      const int idx = lane % (kSymmRowsPerWarp * B), r = idx / B, j = idx % B;
      if (lane < kSymmRowsPerWarp * B && r0 + r < rows) {
        double* dst = row_major ? &y[(long long)(r0 + r) * B + j] : &y[(long long)j * rows + r0 + r];
        atomicAdd(dst, (double)total);
      }
    }
  }
}

// The same pass with the matrix stream decoupled from the registers: every lane copies ITS float4 of
// each of the 4 rows of an iteration straight into a per-warp shared-memory ring (cp.async, 16 bytes,
// zero fill beyond the edge) `kAsyncStages` iterations ahead, and reads back exactly what it copied
// -- no cross-lane hand-off, so cp.async.wait_group is the only synchronisation. The register-staged
// kernel above keeps 8 float4 per lane in flight and is limited to 2 CTAs per SM by its 124
// registers: ncu (30 894 bins) showed 57 % of DRAM peak at 24 % warp occupancy and 33 % issue
// utilisation -- latency-bound, not issue-bound. Here 2 CTAs x 8 warps x 4 stages x 2 KB = 128 KB of
// loads are in flight per SM.
constexpr int kAsyncStages = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int B>
__global__ void __launch_bounds__(kSymmWarps * 32, 2)
symm_panel_async_kernel(const float* __restrict__ c, int rows, int n, long long ld, const float* __restrict__ v,
                        double* __restrict__ y, int row_major) {
  extern __shared__ __align__(16) unsigned char s_symm_raw[];
  float (*s_v)[kSymmCols] = reinterpret_cast<float (*)[kSymmCols]>(s_symm_raw);
  // ring[warp][stage][row 0..3][lane] float4
  float4* s_ring = reinterpret_cast<float4*>(s_symm_raw + (size_t)B * kSymmCols * sizeof(float));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kSymmCols;
  for (int idx = threadIdx.x; idx < B * kSymmCols; idx += blockDim.x) {
    const int j = idx / kSymmCols, col = idx % kSymmCols;
    s_v[j][col] = (k0 + col < n) ? v[(long long)j * n + k0 + col] : 0.f;
  }
  __syncthreads();
  float4* ring = s_ring + (size_t)warp * kAsyncStages * kSymmRowsPerWarp * 32;
  const int rbase = blockIdx.y * kSymmRowsPerCta + warp * (kSymmRowsPerCta / kSymmWarps);
  constexpr int kGroups = kSymmRowsPerCta / kSymmWarps / kSymmRowsPerWarp;      // 8 row groups per warp
  constexpr int kSteps = kSymmCols / 128;                                      // 8 column steps per group
  constexpr int kIters = kGroups * kSteps;
  // iteration it -> (row group, column step); copies 4 rows x 16 bytes per lane
  auto issue = [&](int it) {
    if (it < kIters) {
      const int g = it / kSteps, cc = (it % kSteps) * 128 + lane * 4;
      const int gc = k0 + cc;
      float4* dst = ring + (size_t)(it % kAsyncStages) * kSymmRowsPerWarp * 32 + lane;
#pragma unroll
      for (int r = 0; r < kSymmRowsPerWarp; ++r) {
        const int gr = rbase + g * kSymmRowsPerWarp + r;
        const bool ok = gr < rows && gc < n;
        const int bytes = ok ? min(16, (n - gc) * 4) : 0;        // zero fill beyond the matrix
        cp_async16(dst + r * 32, ok ? (const void*)(c + (long long)gr * ld + gc) : (const void*)c, bytes);
      }
    }
    cp_async_commit();                                           // (empty groups keep the count uniform)
  };
#pragma unroll
  for (int it = 0; it < kAsyncStages - 1; ++it) issue(it);
  float2 acc[kSymmRowsPerWarp][B];
  for (int it = 0; it < kIters; ++it) {
    const int g = it / kSteps, st = it % kSteps;
    if (st == 0) {
#pragma unroll
      for (int r = 0; r < kSymmRowsPerWarp; ++r)
#pragma unroll
        for (int j = 0; j < B; ++j) acc[r][j] = make_float2(0.f, 0.f);
    }
    issue(it + kAsyncStages - 1);
    cp_async_wait<kAsyncStages - 1>();                           // iteration `it` has landed
    const float4* src = ring + (size_t)(it % kAsyncStages) * kSymmRowsPerWarp * 32 + lane;
    float4 a[kSymmRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kSymmRowsPerWarp; ++r) a[r] = src[r * 32];
    const int cc = st * 128 + lane * 4;
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const float4 vv = *reinterpret_cast<const float4*>(&s_v[j][cc]);
#pragma unroll
      for (int r = 0; r < kSymmRowsPerWarp; ++r) {
        acc[r][j] = ffma2(make_float2(a[r].x, a[r].y), make_float2(vv.x, vv.y), acc[r][j]);
        acc[r][j] = ffma2(make_float2(a[r].z, a[r].w), make_float2(vv.z, vv.w), acc[r][j]);
      }
    }
    if (st == kSteps - 1) {
      const int r0 = rbase + g * kSymmRowsPerWarp;
      float part[kSymmRowsPerWarp * B];
#pragma unroll
      for (int r = 0; r < kSymmRowsPerWarp; ++r)
#pragma unroll
        for (int j = 0; j < B; ++j) part[r * B + j] = acc[r][j].x + acc[r][j].y;
      const float total = lane_transpose_sum<kSymmRowsPerWarp * B>(part, lane);
      const int idx = lane % (kSymmRowsPerWarp * B), r = idx / B, j = idx % B;
      if (lane < kSymmRowsPerWarp * B && r0 + r < rows) {
        double* dstp = row_major ? &y[(long long)(r0 + r) * B + j] : &y[(long long)j * rows + r0 + r];
        atomicAdd(dstp, (double)total);
      }
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Y = C V from the UPPER TRIANGLE of a symmetric C only: half the DRAM bytes of symm_panel_kernel
// (the pass is HBM-bound). A CTA owns a block of 128 columns J and a segment of 1024 rows that lies
// on or above it; lane l of every warp owns the 4 columns c0 + 4 l .. + 3 for the whole CTA, warps take
// groups of 4 rows. Every loaded cell C[r][c] (c >= r) feeds BOTH products
//     row part     y[r] += C[r][c] v[c]            (v of the lane's 4 columns lives in registers)
//     column part  y[c] += C[r][c] v[r]  (c > r)   (accumulators of the lane's 4 columns in registers)
// Row parts are reduced over the 32 lanes (transposing butterfly) and written -- not added -- to a
// partial buffer P[J][r][B]; column parts are reduced over the 8 warps through shared memory and
// written to Pc[segment][c][B]. symm_upper_reduce_kernel then sums the partials of a row in a fixed
// order in fp64 (deterministic; the extra traffic is 2 x 128 B per 2 KB of C, ~12 %).
// ---------------------------------------------------------------------------------------------
constexpr int kUpCols = 128;               // columns per CTA
constexpr int kUpSeg = 1024;               // rows per CTA
constexpr int kUpWarps = 8;

template <int B>
__global__ void __launch_bounds__(kUpWarps * 32)
symm_upper_kernel(const float* __restrict__ c, int n, long long ld, const float* __restrict__ v_cm,
                  float* __restrict__ prow, float* __restrict__ pcol) {
  // v_cm: n x B panel, column-major (ld n). prow: [nJ][n][B], pcol: [nSeg][n][B] (fp32)
  extern __shared__ __align__(16) unsigned char s_up_raw[];
  float (*s_vr)[B] = reinterpret_cast<float (*)[B]>(s_up_raw);                 // v of the segment's rows [kUpSeg][B]
  float (*s_col)[kUpCols * B] = reinterpret_cast<float (*)[kUpCols * B]>(s_up_raw);   // aliased after the loop
  const int J = blockIdx.x, seg = blockIdx.y;
  const int c0 = J * kUpCols, r_seg = seg * kUpSeg;
  if (r_seg > c0 + kUpCols - 1) return;                     // segment entirely below the block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r_end = min(min(n, r_seg + kUpSeg), c0 + kUpCols);   // rows beyond the block's last column: nothing above the diagonal
  for (int idx = threadIdx.x; idx < kUpSeg * B; idx += blockDim.x) {
    const int r = r_seg + idx / B, j = idx % B;
    s_vr[idx / B][j] = (r < n) ? v_cm[(long long)j * n + r] : 0.f;
  }
  const int gc = c0 + 4 * lane;                             // first of the lane's 4 columns
  float vc[4][B];                                           // v at the lane's columns
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < B; ++j) vc[q][j] = (gc + q < n) ? v_cm[(long long)j * n + gc + q] : 0.f;
  float2 cacc[4][B / 2];                                    // column parts: [column][vector pair]
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < B / 2; ++j) cacc[q][j] = make_float2(0.f, 0.f);
  __syncthreads();

  float* prow_j = prow + (long long)J * n * B;
  // 8 rows per warp and iteration: all eight 128-bit loads are issued before the first use
  for (int r0 = r_seg + warp * 8; r0 < r_end; r0 += kUpWarps * 8) {
    float4 a[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int r = r0 + rr;
      if (r < r_end && gc < n && gc + 3 >= r) {
        if (gc + 3 < n) a[rr] = ld_stream4(c + (long long)r * ld + gc);
        else {
          const float* p = c + (long long)r * ld + gc;
          a[rr].x = p[0];
          a[rr].y = (gc + 1 < n) ? p[1] : 0.f;
          a[rr].z = (gc + 2 < n) ? p[2] : 0.f;
          a[rr].w = 0.f;
        }
      } else a[rr] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float2 racc[4][B / 2];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = r0 + 4 * h + rr;
        float e[4] = {a[4 * h + rr].x, a[4 * h + rr].y, a[4 * h + rr].z, a[4 * h + rr].w};
        if (gc < r) {                                       // the float4 straddles the diagonal: mask c < r
#pragma unroll
          for (int q = 0; q < 4; ++q) if (gc + q < r) e[q] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < B / 2; ++j) racc[rr][j] = make_float2(0.f, 0.f);
        // row part: sum_q e[q] * v[c_q]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 ee = make_float2(e[q], e[q]);
#pragma unroll
          for (int j = 0; j < B / 2; ++j) racc[rr][j] = ffma2(ee, make_float2(vc[q][2 * j], vc[q][2 * j + 1]), racc[rr][j]);
        }
        // column part: e[q] * v[r] for c_q > r (the diagonal cell belongs to the row part only)
#pragma unroll
        for (int q = 0; q < 4; ++q) if (gc + q == r) e[q] = 0.f;
        const float4* vr4 = reinterpret_cast<const float4*>(&s_vr[min(r, r_seg + kUpSeg - 1) - r_seg][0]);
        float vr[B];
#pragma unroll
        for (int j4 = 0; j4 < B / 4; ++j4) {
          const float4 t = vr4[j4];
          vr[4 * j4] = t.x; vr[4 * j4 + 1] = t.y; vr[4 * j4 + 2] = t.z; vr[4 * j4 + 3] = t.w;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 ee = make_float2(e[q], e[q]);
#pragma unroll
          for (int j = 0; j < B / 2; ++j) cacc[q][j] = ffma2(ee, make_float2(vr[2 * j], vr[2 * j + 1]), cacc[q][j]);
        }
      }
      float part[4 * B];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int j = 0; j < B / 2; ++j) { part[rr * B + 2 * j] = racc[rr][j].x; part[rr * B + 2 * j + 1] = racc[rr][j].y; }
      const float total = lane_transpose_sum<4 * B>(part, lane);
      const int idx = lane % (4 * B), rr = idx / B, j = idx % B;
      if (lane < 4 * B && r0 + 4 * h + rr < r_end) prow_j[(long long)(r0 + 4 * h + rr) * B + j] = total;
    }
  }
  // column parts: reduce over the warps (fixed order) and write Pc[seg][c][B]
  __syncthreads();                                          // s_vr is dead: alias it
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < B / 2; ++j) {
      s_col[warp][(4 * lane + q) * B + 2 * j] = cacc[q][j].x;
      s_col[warp][(4 * lane + q) * B + 2 * j + 1] = cacc[q][j].y;
    }
  __syncthreads();
  float* pcol_s = pcol + ((long long)seg * n + c0) * B;
  for (int idx = threadIdx.x; idx < kUpCols * B; idx += blockDim.x) {
    if (c0 + idx / B >= n) break;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kUpWarps; ++w) t += s_col[w][idx];
    pcol_s[idx] = t;
  }
}

// y[r][j] (fp64, row-major) = sum over the column blocks J >= block(r) of P[J][r][j]
//                           + sum over the row segments seg <= segment(r) of Pc[seg][r][j]
template <int B>
__global__ void __launch_bounds__(256)
symm_upper_reduce_kernel(const float* __restrict__ prow, const float* __restrict__ pcol, int n,
                         double* __restrict__ y) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * B) return;
  const int r = (int)(idx / B);
  const int nJ = (n + kUpCols - 1) / kUpCols;
  double s = 0.0;
  for (int J = r / kUpCols; J < nJ; ++J) s += (double)prow[(long long)J * n * B + idx];
  // column parts of cell (r', r), r' < r, were produced by the segments that hold rows r' < r of block(r)
  for (int seg = 0; seg <= min(r, (r / kUpCols) * kUpCols + kUpCols - 1) / kUpSeg; ++seg)
    s += (double)pcol[(long long)seg * n * B + idx];
  y[idx] = s;
}

long long symm_upper_ws_bytes(int n, int b) {
  const long long nJ = (n + kUpCols - 1) / kUpCols, nSeg = (n + kUpSeg - 1) / kUpSeg;
  return (nJ + nSeg) * (long long)n * b * 4 + 256;
}

template <int B>
int launch_symm_upper(const float* c, int n, long long ld, const float* v, double* y, void* ws, cudaStream_t s) {
  const int nJ = (n + kUpCols - 1) / kUpCols, nSeg = (n + kUpSeg - 1) / kUpSeg;
  float* prow = static_cast<float*>(ws);
  float* pcol = prow + (long long)nJ * n * B;
  const int smem = max(kUpSeg * B, kUpWarps * kUpCols * B) * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    HIA_CUDA(cudaFuncSetAttribute(symm_upper_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  symm_upper_kernel<B><<<dim3(nJ, nSeg), kUpWarps * 32, smem, s>>>(c, n, ld, v, prow, pcol);
  HIA_LAUNCH_CHECK();
  symm_upper_reduce_kernel<B><<<(int)(((long long)n * B + 255) / 256), 256, 0, s>>>(prow, pcol, n, y);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

// Gpart[chunk][p x b] (row-major) = A[chunk rows, 0:p]^T B[chunk rows, 0:b]; A, B col-major fp64, ld = n.
__global__ void __launch_bounds__(256)
gram_kernel(const double* __restrict__ a, int p, const double* __restrict__ bm, int b, int n,
            int rows_per_cta, double* __restrict__ g) {
  __shared__ double s_red[8][kMaxBlock];
  const int i = blockIdx.x;                  // column of A
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(n, r0 + rows_per_cta);
  double acc[kMaxBlock];
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j) acc[j] = 0.0;
  const double* ai = a + (long long)i * n;
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    const double x = ai[r];
#pragma unroll
    for (int j = 0; j < kMaxBlock; ++j)
      if (j < b) acc[j] += x * bm[(long long)j * n + r];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j) {
    if (j < b) {
      const double s = warp_sum(acc[j]);
      if (lane == 0) s_red[warp][j] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < b) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
    // per-row-chunk partial, summed in a FIXED order by gram_reduce_kernel: the replicated solvers of
    // a row-sharded run must stay bit-identical (atomics would make the rounding order-dependent)
    g[((long long)blockIdx.y * p + i) * b + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256)
gram_reduce_kernel(const double* __restrict__ gpart, int chunks, int count, double* __restrict__ g) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += gpart[(long long)c * count + idx];
  g[idx] = s;
}

// Y[:, j] -= sum_i A[:, i] G[i][j]   (thread per row)
__global__ void __launch_bounds__(256)
panel_sub_kernel(const double* __restrict__ a, int p, const double* __restrict__ g, int b, int n,
                 double* __restrict__ y) {
  extern __shared__ double s_g[];
  for (int idx = threadIdx.x; idx < p * b; idx += blockDim.x) s_g[idx] = g[idx];
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double acc[kMaxBlock];
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j) acc[j] = 0.0;
  for (int i = 0; i < p; ++i) {
    const double x = a[(long long)i * n + r];
#pragma unroll
    for (int j = 0; j < kMaxBlock; ++j)
      if (j < b) acc[j] += x * s_g[i * b + j];
  }
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j)
    if (j < b) y[(long long)j * n + r] -= acc[j];
}

// Cholesky of G (b x b, row-major) = R^T R; writes R (upper) and Rinv (upper). One thread.
// A pivot below eps * max diagonal deflates the column (Rinv column = 0).
__global__ void chol_kernel(const double* __restrict__ g, int b, double* __restrict__ r_out,
                            double* __restrict__ rinv_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double R[kMaxBlock][kMaxBlock], Ri[kMaxBlock][kMaxBlock];
  double dmax = 0.0;
  for (int i = 0; i < b; ++i) dmax = fmax(dmax, g[i * b + i]);
  bool dead[kMaxBlock];
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < b; ++j) { R[i][j] = 0.0; Ri[i][j] = 0.0; }
  for (int j = 0; j < b; ++j) {
    double d = g[j * b + j];
    for (int t = 0; t < j; ++t) d -= R[t][j] * R[t][j];
    dead[j] = !(d > 1e-24 * dmax) || !(dmax > 0.0);
    if (dead[j]) { R[j][j] = 0.0; continue; }
    R[j][j] = sqrt(d);
    for (int c = j + 1; c < b; ++c) {
      double s = 0.5 * (g[j * b + c] + g[c * b + j]);
      for (int t = 0; t < j; ++t) s -= R[t][j] * R[t][c];
      R[j][c] = s / R[j][j];
    }
  }
  // Rinv by back substitution on the live columns
  for (int j = 0; j < b; ++j) {
    if (dead[j]) continue;
    Ri[j][j] = 1.0 / R[j][j];
    for (int i = j - 1; i >= 0; --i) {
      if (dead[i]) continue;
      double s = 0.0;
      for (int t = i + 1; t <= j; ++t) s += R[i][t] * Ri[t][j];
      Ri[i][j] = -s / R[i][i];
    }
  }
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < b; ++j) { r_out[i * b + j] = R[i][j]; rinv_out[i * b + j] = Ri[i][j]; }
}

// Y <- Y * M (M b x b row-major), thread per row; optionally also writes an fp32 copy.
__global__ void __launch_bounds__(256)
panel_mul_kernel(double* __restrict__ y, const double* __restrict__ mtx, int b, int n,
                 float* __restrict__ yf) {
  __shared__ double s_m[kMaxBlock * kMaxBlock];
  for (int idx = threadIdx.x; idx < b * b; idx += blockDim.x) s_m[idx] = mtx[idx];
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double in[kMaxBlock], out[kMaxBlock];
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j) in[j] = (j < b) ? y[(long long)j * n + r] : 0.0;
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j) {
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < kMaxBlock; ++t)
      if (t < b && j < b) s += in[t] * s_m[t * b + j];
    out[j] = s;
  }
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j)
    if (j < b) {
      y[(long long)j * n + r] = out[j];
      if (yf) yf[(long long)j * n + r] = (float)out[j];
    }
}

// dst (b x b) = A * B  (tiny, one thread): R_total = R2 * R1
__global__ void small_matmul_kernel(const double* a, const double* bb, int b, double* dst) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double tmp[kMaxBlock * kMaxBlock];
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < b; ++j) {
      double s = 0.0;
      for (int t = 0; t < b; ++t) s += a[i * b + t] * bb[t * b + j];
      tmp[i * b + j] = s;
    }
  for (int i = 0; i < b * b; ++i) dst[i] = tmp[i];
}

// H[(row0 + i) * D + col0 + j] (+)= src[i * b + j]
__global__ void place_block_kernel(const double* __restrict__ src, int rows, int b, double* __restrict__ h,
                                   int D, int row0, int col0, int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * b) return;
  const int i = idx / b, j = idx % b;
  double* dst = &h[(long long)(row0 + i) * D + col0 + j];
  *dst = accumulate ? (*dst + src[idx]) : src[idx];
}

// One-sided (Hestenes) Jacobi on the symmetric dim x dim matrix built from the lower triangle of H.
// Single CTA, warp per column pair, round-robin pairing. A and V (col-major) live in SHARED memory
// when 2 dim^2 doubles fit (dim <= 118: every default configuration), else in the global scratch
// A_g / V_g; V (and lam / order) are written back to global memory at the end. The global-memory
// version spent 6 ms per solve on L2 round trips between the dependent rotation steps.
__global__ void __launch_bounds__(1024)
jacobi_kernel(const double* __restrict__ h, int D, int dim, double* __restrict__ A_g, double* __restrict__ V_g,
              double* __restrict__ lam, int* __restrict__ order, int max_sweeps, int use_smem) {
  extern __shared__ double s_jac[];
  __shared__ int s_rot;
  double* A = use_smem ? s_jac : A_g;
  double* V = use_smem ? s_jac + (size_t)dim * dim : V_g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int m = (dim + 1) & ~1;                      // padded to even (phantom index dim)
  for (int idx = tid; idx < dim * dim; idx += blockDim.x) {
    const int col = idx / dim, row = idx % dim;
    const double v = (row >= col) ? h[(long long)row * D + col] : h[(long long)col * D + row];
    A[col * dim + row] = v;
    V[col * dim + row] = (row == col) ? 1.0 : 0.0;
  }
  // Round-robin pairing, tabulated once: pair `pi` of step `step` (the first profile of this kernel
  // showed more than half of its instructions in the two integer modulos, 64-bit index arithmetic
  // and loop control around ~20 useful fp64 operations per lane: the kernel is ISSUE-bound on one SM).
  __shared__ unsigned short s_p[136], s_q[136];         // pairs of ONE step (<= 136), rebuilt per step
  __shared__ double s_norm[272];                        // squared column norms, maintained incrementally
  __shared__ double s_amax;
  if (tid == 0) s_amax = 0.0;
  __syncthreads();
  // Noise floor of the column inner products. Every rotation perturbs the columns by ~eps ||T||, so
  // gamma = a_p . a_q cannot be driven below ~eps ||T|| max(||a_p||, ||a_q||): pairs of columns that
  // belong to SMALL Ritz values can never meet the purely relative test. amax = max_p ||a_p||^2 ~ ||T||^2.
  for (int p = warp; p < dim; p += nwarps) {
    double a2 = 0.0;
    for (int r = lane; r < dim; r += 32) { const double x = A[p * dim + r]; a2 += x * x; }
    a2 = warp_sum(a2);
    if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(&s_amax), (unsigned long long)__double_as_longlong(a2));
  }
  __syncthreads();
  const double floor2 = 1.2e-28 * s_amax;                // (50 eps)^2 ||T||^2
  const int npairs = m / 2;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    if (tid == 0) s_rot = 0;
    // exact squared norms at the start of every sweep; inside the sweep a rotation updates them as
    // alpha' = alpha - t gamma, beta' = beta + t gamma (the rotation that annihilates gamma), so a
    // step needs ONE inner product and one warp reduction instead of three
    for (int p = warp; p < dim; p += nwarps) {
      double a2 = 0.0;
      for (int r = lane; r < dim; r += 32) { const double x = A[p * dim + r]; a2 += x * x; }
      a2 = warp_sum(a2);
      if (lane == 0) s_norm[p] = a2;
    }
    __syncthreads();
    for (int step = 0; step < m - 1; ++step) {
      if (tid < npairs) {
        int p, q;
        if (tid == 0) { p = m - 1; q = step; }
        else { p = (step + tid) % (m - 1); q = (step - tid + (m - 1)) % (m - 1); }
        s_p[tid] = (unsigned short)p;
        s_q[tid] = (unsigned short)q;
      }
      __syncthreads();
      for (int pi = warp; pi < npairs; pi += nwarps) {
        const int p = s_p[pi], q = s_q[pi];
        if (p >= dim || q >= dim) continue;
        double* ap = A + p * dim;
        double* aq = A + q * dim;
        double gamma = 0.0;
#pragma unroll 4
        for (int r = lane; r < dim; r += 32) gamma += ap[r] * aq[r];
        gamma = warp_sum(gamma);
        const double alpha = s_norm[p], beta = s_norm[q];
        // |gamma| <= 1e-12 sqrt(alpha beta), compared squared (no sqrt on the dependent chain), or
        // below the noise floor
        const double g2 = gamma * gamma;
        if (g2 <= 1e-24 * alpha * beta || g2 <= floor2 * fmax(alpha, beta) || gamma == 0.0) continue;
        // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (beta - alpha) / (2 gamma), written with
        // ONE division and one square root:  t = sign(d g) |2 gamma| / (|d| + sqrt(d^2 + 4 gamma^2))
        const double d = beta - alpha, tg = 2.0 * gamma;
        const double den = fabs(d) + sqrt(d * d + tg * tg);
        const double t = copysign(fabs(tg) / den, d * tg >= 0.0 ? 1.0 : -1.0);
        const double cs = rsqrt(1.0 + t * t), sn = cs * t;
        if (lane == 0) { s_rot = 1; s_norm[p] = alpha - t * gamma; s_norm[q] = beta + t * gamma; }
        double* vp = V + p * dim;
        double* vq = V + q * dim;
#pragma unroll 4
        for (int r = lane; r < dim; r += 32) {
          const double x = ap[r], y = aq[r];
          ap[r] = cs * x - sn * y; aq[r] = sn * x + cs * y;
          const double vx = vp[r], vy = vq[r];
          vp[r] = cs * vx - sn * vy; vq[r] = sn * vx + cs * vy;
        }
      }
      __syncthreads();
    }
    const int rotated = s_rot;
    __syncthreads();
    if (!rotated) break;
  }
  // eigenvalue_p = v_p . (T v_p) = v_p . a_p   (a_p = T v_p is maintained by the rotations)
  for (int p = warp; p < dim; p += nwarps) {
    double s = 0.0;
    for (int r = lane; r < dim; r += 32) s += V[(long long)p * dim + r] * A[(long long)p * dim + r];
    s = warp_sum(s);
    if (lane == 0) lam[p] = s;
  }
  if (use_smem)
    for (int idx = tid; idx < dim * dim; idx += blockDim.x) V_g[idx] = V[idx];
  __syncthreads();
  if (tid == 0) {                                    // descending order (selection sort, dim <= 272)
    for (int i = 0; i < dim; ++i) order[i] = i;
    for (int i = 0; i < dim; ++i) {
      int best = i;
      for (int j = i + 1; j < dim; ++j) if (lam[order[j]] > lam[order[best]]) best = j;
      const int t = order[i]; order[i] = order[best]; order[best] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Top eigenpairs of the projected matrix WITHOUT a full decomposition (default; the one-sided Jacobi
// above stays as the fallback for dim > 160 and behind HIA_EIG_JACOBI=1). Rayleigh-Ritz needs only
// the `kk` largest pairs (kk = the block size: the wanted pairs plus the rest of the restart block) and
// one more eigenvalue for the gaps; the Jacobi solve needs 13 sweeps on the graded Krylov matrix and is
// issue-bound on one SM (1.35 ms at dim 64, 4.5 ms at dim 96, O(dim^3)). One CTA, three phases:
//   1. Householder tridiagonalisation of T in shared memory (dim - 2 reflectors; matrix-vector
//      product and rank-2 update spread over 32 warps; the reflectors stay in the rows of A),
//   2. the kk + 1 largest eigenvalues of the tridiagonal matrix by multisection on Sturm counts (all
//      of them at once: 1024 threads = kk + 1 groups of shifts, ~8 rounds to 1e-15 relative),
//   3. inverse iteration per eigenvalue (one thread each: pivoted tridiagonal LU, 3 solves; vectors
//      of eigenvalues closer than 1e-3 ||T|| are re-orthogonalised like LAPACK's dstein), then the
//      reflectors applied backwards, one warp per vector, vector in registers.
// Output convention of the consumers: Z column j / lam[j] / order[j] = j for the j-th LARGEST pair,
// j < kk; lam[kk] = the next eigenvalue (gap of the last pair); lam[j > kk] = -1e300.
// A numpy prototype of exactly these steps reproduces numpy.linalg.eigh to 1e-15 on clustered
// (1410.5 / 1404.5 / 1392.2), near-degenerate (1e-9 apart), indefinite and banded test matrices.
// ---------------------------------------------------------------------------------------------
constexpr int kTriMaxDim = 272;
constexpr int kTriSmemDim = 160;            // A (dim^2 doubles) must fit in shared memory
constexpr int kTriMaxVec = 16;

__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e2, int n,
                                           double x, double pivmin) {
  int cnt = 0;
  double q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  cnt += q < 0.0;
  for (int i = 1; i < n; ++i) {
    q = d[i] - x - e2[i - 1] / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
  }
  return cnt;                                // eigenvalues < x
}

__global__ void __launch_bounds__(1024)
tri_eig_kernel(const double* __restrict__ h, int D, int dim, int kk, double* __restrict__ Z,
               double* __restrict__ lam, int* __restrict__ order) {
  extern __shared__ double s_tri[];
  double* A = s_tri;                                    // dim x dim, row-major, symmetric
  __shared__ double s_d[kTriMaxDim], s_e[kTriMaxDim], s_e2[kTriMaxDim], s_beta[kTriMaxDim];
  __shared__ double s_v[kTriMaxDim], s_p[kTriMaxDim];
  __shared__ double s_scal[4];
  __shared__ double s_lo[kTriMaxVec + 1], s_hi[kTriMaxVec + 1], s_lam[kTriMaxVec + 1];
  __shared__ int s_cnt[1024];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = dim;
  for (int idx = tid; idx < n * n; idx += blockDim.x) {
    const int r = idx / n, c = idx % n;
    A[idx] = (r >= c) ? h[(long long)r * D + c] : h[(long long)c * D + r];
  }
  __syncthreads();

  // ---- 1. Householder tridiagonalisation (reflector k acts on rows / columns k + 1 .. n - 1)
  for (int k = 0; k + 2 < n; ++k) {
    const int m = n - k - 1;
    const double* x = A + k * n + k + 1;                // row k right of the diagonal == column k below it
    if (warp == 0) {
      double sig = 0.0;
      for (int i = 1 + lane; i < m; i += 32) sig += x[i] * x[i];
      sig = warp_sum(sig);
      if (lane == 0) {
        const double x0 = x[0];
        double alpha = x0, b = 0.0, v0 = 0.0;
        if (sig != 0.0) {
          alpha = -copysign(sqrt(x0 * x0 + sig), x0);
          v0 = x0 - alpha;
          b = 2.0 / (v0 * v0 + sig);
        }
        s_scal[0] = alpha; s_scal[1] = b; s_scal[2] = v0;
      }
    }
    __syncthreads();
    const double b = s_scal[1];
    if (b != 0.0) {
      for (int i = tid; i < m; i += blockDim.x) s_v[i] = (i == 0) ? s_scal[2] : x[i];
      __syncthreads();
      // p = b * A22 v  (warp per row)
      for (int i = warp; i < m; i += 32) {
        const double* row = A + (k + 1 + i) * n + k + 1;
        double acc = 0.0;
        for (int j = lane; j < m; j += 32) acc += row[j] * s_v[j];
        acc = warp_sum(acc);
        if (lane == 0) s_p[i] = b * acc;
      }
      __syncthreads();
      if (warp == 0) {
        double kq = 0.0;
        for (int i = lane; i < m; i += 32) kq += s_p[i] * s_v[i];
        kq = warp_sum(kq);
        if (lane == 0) s_scal[3] = 0.5 * b * kq;
      }
      __syncthreads();
      for (int i = tid; i < m; i += blockDim.x) s_p[i] -= s_scal[3] * s_v[i];      // w
      __syncthreads();
      for (int i = warp; i < m; i += 32) {                                        // A22 -= v w^T + w v^T
        double* row = A + (k + 1 + i) * n + k + 1;
        const double vi = s_v[i], wi = s_p[i];
        for (int j = lane; j < m; j += 32) row[j] -= vi * s_p[j] + wi * s_v[j];
      }
      // keep the reflector in row k (its old content is no longer needed)
      for (int i = tid; i < m; i += blockDim.x) A[k * n + k + 1 + i] = s_v[i];
    }
    if (tid == 0) { s_e[k] = s_scal[0]; s_beta[k] = b; }
    __syncthreads();
  }
  for (int i = tid; i < n; i += blockDim.x) s_d[i] = A[i * n + i];
  if (tid == 0) {
    if (n >= 2) { s_e[n - 2] = A[(n - 1) * n + n - 2]; s_beta[n - 2] = 0.0; }
    s_e[n - 1] = 0.0;
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) s_e2[i] = s_e[i] * s_e[i];
  __syncthreads();

  // ---- 2. the ne = min(kk + 1, n) largest eigenvalues: multisection on Sturm counts
  const int ne = min(kk + 1, n);
  if (warp == 0) {
    double lo = INFINITY, hi = -INFINITY, emax = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double r = fabs(s_e[i]) + (i > 0 ? fabs(s_e[i - 1]) : 0.0);
      lo = fmin(lo, s_d[i] - r);
      hi = fmax(hi, s_d[i] + r);
      emax = fmax(emax, s_e2[i]);
    }
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    }
    if (lane == 0) {
      const double span = hi - lo;
      s_scal[0] = lo - 1e-3 * span - 1e-300;
      s_scal[1] = hi + 1e-3 * span + 1e-300;
      s_scal[2] = 2.2250738585072014e-308 * fmax(1.0, emax);      // pivmin
      s_scal[3] = fmax(fabs(lo), fabs(hi));                       // ~ ||T||
    }
  }
  __syncthreads();
  const double pivmin = s_scal[2], tnorm = s_scal[3];
  if (tid < ne) { s_lo[tid] = s_scal[0]; s_hi[tid] = s_scal[1]; }
  __syncthreads();
  const int S = (int)blockDim.x / ne;                   // shifts per eigenvalue and round
  const int g = tid / S, sidx = tid % S;
  const int rounds = (int)ceil(56.0 / log2((double)(S + 1)));
  for (int r = 0; r < rounds; ++r) {
    int cnt = 0;
    double xs = 0.0;
    if (g < ne) {
      const double a = s_lo[g], bb = s_hi[g];
      xs = a + (bb - a) * ((double)(sidx + 1) / (double)(S + 1));
      cnt = sturm_count(s_d, s_e2, n, xs, pivmin);
    }
    s_cnt[tid] = cnt;
    __syncthreads();
    if (g < ne) {
      const int idx = n - 1 - g;                        // ascending index of the g-th largest eigenvalue
      // bracket: the last shift with count <= idx and the first with count > idx
      const bool above = cnt > idx;
      const bool prev_above = (sidx > 0) ? (s_cnt[tid - 1] > idx) : false;
      if (above && !prev_above) s_hi[g] = xs;
      const bool next_above = (sidx + 1 < S) ? (s_cnt[tid + 1] > idx) : true;
      if (!above && next_above) s_lo[g] = xs;
    }
    __syncthreads();
  }
  if (tid < ne) s_lam[tid] = 0.5 * (s_lo[tid] + s_hi[tid]);
  __syncthreads();

  // ---- 3a. inverse iteration on the tridiagonal matrix, one thread per eigenvalue
  const int nv = min(kk, n);
  double dl[kTriMaxDim], dd[kTriMaxDim], du[kTriMaxDim], du2[kTriMaxDim];
  unsigned char piv[kTriMaxDim];
  if (tid < nv) {
    const double lamj = s_lam[tid];
    for (int i = 0; i < n; ++i) { dl[i] = s_e[i]; dd[i] = s_d[i] - lamj; du[i] = s_e[i]; du2[i] = 0.0; piv[i] = 0; }
    const double tiny = 2.220446049250313e-16 * fmax(tnorm, 1e-300);
    for (int i = 0; i + 1 < n; ++i) {
      if (fabs(dd[i]) >= fabs(dl[i])) {
        if (dd[i] == 0.0) dd[i] = tiny;
        const double f = dl[i] / dd[i];
        dl[i] = f;
        dd[i + 1] -= f * du[i];
      } else {
        piv[i] = 1;
        const double f = dd[i] / dl[i];
        dd[i] = dl[i];
        dl[i] = f;
        const double t = du[i];
        du[i] = dd[i + 1];
        dd[i + 1] = t - f * dd[i + 1];
        if (i + 2 < n) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
      }
    }
    if (dd[n - 1] == 0.0) dd[n - 1] = tiny;
  }
  double* X = Z;                                        // columns 0 .. nv - 1 of the output hold the iterates
  if (tid < nv) {                                       // start vectors: deterministic, not orthogonal to anything special
    uint32_t st = 0x9e3779b9u * (uint32_t)(tid + 1);
    for (int i = 0; i < n; ++i) {
      st = hash32(st + (uint32_t)i);
      X[(long long)tid * n + i] = (double)st * (2.0 / 4294967296.0) - 1.0;
    }
  }
  __syncthreads();
  for (int it = 0; it < 3; ++it) {
    if (tid < nv) {
      double* x = X + (long long)tid * n;
      for (int i = 0; i + 1 < n; ++i) {                 // forward substitution with the row interchanges
        if (!piv[i]) x[i + 1] -= dl[i] * x[i];
        else { const double t = x[i]; x[i] = x[i + 1]; x[i + 1] = t - dl[i] * x[i]; }
      }
      x[n - 1] /= dd[n - 1];
      if (n > 1) x[n - 2] = (x[n - 2] - du[n - 2] * x[n - 1]) / dd[n - 2];
      for (int i = n - 3; i >= 0; --i) x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / dd[i];
    }
    __syncthreads();
    if (warp == 0) {                                    // modified Gram-Schmidt inside clusters + normalisation
      for (int j = 0; j < nv; ++j) {
        double* xj = X + (long long)j * n;
        for (int p = 0; p < j; ++p) {
          if (fabs(s_lam[p] - s_lam[j]) > 1e-3 * tnorm) continue;
          const double* xp = X + (long long)p * n;
          double dot = 0.0;
          for (int i = lane; i < n; i += 32) dot += xp[i] * xj[i];
          dot = warp_sum(dot);
          for (int i = lane; i < n; i += 32) xj[i] -= dot * xp[i];
          __syncwarp();
        }
        double nrm = 0.0;
        for (int i = lane; i < n; i += 32) nrm += xj[i] * xj[i];
        nrm = warp_sum(nrm);
        const double inv = rsqrt(nrm);
        for (int i = lane; i < n; i += 32) xj[i] *= inv;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  // ---- 3b. back-transformation: z = H_0 H_1 ... H_{n-3} x, one warp per vector, vector in registers
  for (int j = warp; j < nv; j += 32) {
    double* xj = X + (long long)j * n;
    double xr[(kTriMaxDim + 31) / 32];
#pragma unroll
    for (int q = 0; q < (kTriMaxDim + 31) / 32; ++q) xr[q] = (lane + 32 * q < n) ? xj[lane + 32 * q] : 0.0;
    for (int k = n - 3; k >= 0; --k) {
      const double b = s_beta[k];
      if (b == 0.0) continue;
      const double* v = A + k * n;                      // v_i at A[k][k + 1 + i] <-> element index k + 1 + i
      double dot = 0.0;
#pragma unroll
      for (int q = 0; q < (kTriMaxDim + 31) / 32; ++q) {
        const int i = lane + 32 * q;
        if (i > k && i < n) dot += v[i] * xr[q];
      }
      dot = b * warp_sum(dot);
#pragma unroll
      for (int q = 0; q < (kTriMaxDim + 31) / 32; ++q) {
        const int i = lane + 32 * q;
        if (i > k && i < n) xr[q] -= dot * v[i];
      }
    }
#pragma unroll
    for (int q = 0; q < (kTriMaxDim + 31) / 32; ++q) if (lane + 32 * q < n) xj[lane + 32 * q] = xr[q];
  }
  for (int i = tid; i < n; i += blockDim.x) {
    lam[i] = (i < ne) ? s_lam[i] : -1e300;
    order[i] = i;
  }
}

// After `done` Lanczos steps (dim = done * b Ritz pairs in Z / lam / order): the Lanczos relation
// C Q = Q T + Q_next B E^T gives the residual of the Ritz pair (theta, z) WITHOUT a pass over C:
// ||C u - theta u|| = ||B z_last||, B = H[dim : dim + b, dim - b : dim] (the R of the newest block),
// z_last = the last b entries of z. gap_j = distance of theta_j to the nearest other Ritz value
// (the sin-theta theorem bounds the eigenvector angle by residual / gap).
// out[0..k) = estimates, out[k..2k) = gaps, flag = 1 when every pair j < k satisfies
// est_j <= rel_tol |theta_1|  or  est_j <= angle_tol gap_j.
__global__ void ritz_check_kernel(const double* __restrict__ h, int D, int dim, int b,
                                  const double* __restrict__ Z, const double* __restrict__ lam,
                                  const int* __restrict__ order, int k, double rel_tol, double angle_tol,
                                  double* __restrict__ out, int* __restrict__ flag) {
  __shared__ int s_ok[kMaxBlock];
  const int j = threadIdx.x;
  if (j < k) {
    const double* z = Z + (long long)order[j] * dim;
    double r2 = 0.0;
    for (int i = 0; i < b; ++i) {
      double acc = 0.0;
      for (int c = 0; c < b; ++c) acc += h[(long long)(dim + i) * D + (dim - b + c)] * z[dim - b + c];
      r2 += acc * acc;
    }
    const double est = sqrt(r2), theta = lam[order[j]];
    double gap = INFINITY;
    for (int i = 0; i < dim; ++i)
      if (i != order[j]) gap = fmin(gap, fabs(theta - lam[i]));
    out[j] = est;
    out[k + j] = gap;
    s_ok[j] = (est <= rel_tol * fabs(lam[order[0]])) || (est <= angle_tol * gap);
  }
  __syncthreads();
  if (j == 0) {
    int ok = 1;
    for (int i = 0; i < k; ++i) ok &= s_ok[i];
    *flag = ok;
  }
}

// gaps[j] = distance of the j-th largest Ritz value to the nearest other Ritz value
__global__ void ritz_gap_kernel(const double* __restrict__ lam, const int* __restrict__ order, int dim, int k,
                                double* __restrict__ gaps) {
  const int j = threadIdx.x;
  if (j >= k) return;
  const double theta = lam[order[j]];
  double gap = INFINITY;
  for (int i = 0; i < dim; ++i)
    if (i != order[j]) gap = fmin(gap, fabs(theta - lam[i]));
  gaps[j] = gap;
}

// U[:, j] = Q[:, 0:dim] * Z[:, order[j]], j < k  (thread per row); fp64 + fp32 copies, eigenvalues out
__global__ void __launch_bounds__(256)
ritz_kernel(const double* __restrict__ q, int n, int dim, const double* __restrict__ Z,
            const int* __restrict__ order, const double* __restrict__ lam, int k,
            double* __restrict__ u, float* __restrict__ uf, double* __restrict__ eigvals) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < k) eigvals[threadIdx.x] = lam[order[threadIdx.x]];
  if (r >= n) return;
  double acc[kMaxBlock];
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j) acc[j] = 0.0;
  for (int i = 0; i < dim; ++i) {
    const double x = q[(long long)i * n + r];
#pragma unroll
    for (int j = 0; j < kMaxBlock; ++j)
      if (j < k) acc[j] += x * Z[(long long)order[j] * dim + i];
  }
#pragma unroll
  for (int j = 0; j < kMaxBlock; ++j)
    if (j < k) { u[(long long)j * n + r] = acc[j]; uf[(long long)j * n + r] = (float)acc[j]; }
}

// resid[j] = || W[:, j] - lambda_j U[:, j] ||   (sum of squares via atomics, sqrt later)
__global__ void __launch_bounds__(256)
resid_kernel(const double* __restrict__ w, const double* __restrict__ u, const double* __restrict__ eigvals,
             int n, int k, double* __restrict__ resid2) {
  const int j = blockIdx.y;
  double s = 0.0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const double d = w[(long long)j * n + r] - eigvals[j] * u[(long long)j * n + r];
    s += d * d;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(&resid2[j], s);
}
__global__ void sqrt_kernel(double* x, int k) { if (threadIdx.x < k) x[threadIdx.x] = sqrt(x[threadIdx.x]); }

// ---------------------------------------------------------------------------------------------
// One Lanczos step after the pass over the matrix, as ONE cooperative kernel (grid-wide barriers
// between the phases instead of ~20 dependent launches: gram + reduce + subtract + place, twice;
// gram + reduce + Cholesky + multiply, twice; R2 R1; place -- the launch list of round 1 showed
// ~2.2 ms of such kernels plus ~3 ms of launch gaps per 8-step cycle at 30 894 bins):
//   Qn = Y (row-major staging -> column-major block j + 1 of Q)
//   twice:  G = Q[:, 0:p]^T Qn;  Qn -= Q G;  H[0:p, j-block] (+)= G                       (CGS2)
//   twice:  S = Qn^T Qn = R^T R;  Qn = Qn R^-1                                            (CholQR2)
//   H[(j+1)-block, j-block] = R2 R1;  fp32 copy of Qn = the next panel
// CTA c owns the rows [c * rows_per_cta, ...). Every Gram matrix is reduced in a FIXED order (per-CTA
// partials, then entry e summed over the CTAs by CTA e mod G): bit-identical on every rank of a
// row-sharded run, like the kernels it replaces. The b x b Cholesky is computed redundantly by every
// CTA (same inputs, same code: same bits) -- cheaper than another grid-wide barrier.
// ---------------------------------------------------------------------------------------------
constexpr int kStepThreads = 256;
constexpr int kMaxStepGrid = 256;          // upper bound of the grid (workspace sizing)

struct StepArgs {
  const double* y_full;                    // n x b row-major (all rows)
  double* Q;                               // n x D column-major
  double* H;                               // D x D row-major
  double* Gpart;                           // [grid][p * b]
  double* G;                               // [p * b] reduced (also b * b for the Cholesky steps)
  double* R1;                              // b x b
  double* R2;                              // b x b
  float* Vf;                               // n x b column-major fp32 (next panel)
  int n, b, D, j, rows_per_cta;
};

// Cholesky of the symmetric b x b matrix g (row-major, shared) -> R (upper), Rinv (upper), by one warp's
// lane 0 ... small and on the critical path of every CTA, so plain registers / shared arrays only.
__device__ void chol_small(const double* g, int b, double* R, double* Ri) {
  double dmax = 0.0;
  for (int i = 0; i < b; ++i) dmax = fmax(dmax, g[i * b + i]);
  bool dead[kMaxBlock];
  for (int i = 0; i < b * b; ++i) { R[i] = 0.0; Ri[i] = 0.0; }
  for (int j = 0; j < b; ++j) {
    double d = g[j * b + j];
    for (int t = 0; t < j; ++t) d -= R[t * b + j] * R[t * b + j];
    dead[j] = !(d > 1e-24 * dmax) || !(dmax > 0.0);
    if (dead[j]) { R[j * b + j] = 0.0; continue; }
    const double rjj = sqrt(d);
    R[j * b + j] = rjj;
    const double inv = 1.0 / rjj;
    for (int c = j + 1; c < b; ++c) {
      double s = 0.5 * (g[j * b + c] + g[c * b + j]);
      for (int t = 0; t < j; ++t) s -= R[t * b + j] * R[t * b + c];
      R[j * b + c] = s * inv;
    }
  }
  for (int j = 0; j < b; ++j) {
    if (dead[j]) continue;
    Ri[j * b + j] = 1.0 / R[j * b + j];
    for (int i = j - 1; i >= 0; --i) {
      if (dead[i]) continue;
      double s = 0.0;
      for (int t = i + 1; t <= j; ++t) s += R[i * b + t] * Ri[t * b + j];
      Ri[i * b + j] = -s / R[i * b + i];
    }
  }
}

template <int B>
__global__ void __launch_bounds__(kStepThreads)
lanczos_step_kernel(const StepArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_g[272 * B];                         // reduced Gram matrix (p <= 272 rows)
  __shared__ double s_r[B * B], s_ri[B * B], s_r1[B * B];
  const int n = a.n, D = a.D, j = a.j;
  const int p = (j + 1) * B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int r_lo = min(n, cta * a.rows_per_cta), r_hi = min(n, r_lo + a.rows_per_cta);
  double* Qn = a.Q + (long long)p * n;                    // block j + 1

  // ---- Qn <- Y (own rows)
  for (int r = r_lo + tid; r < r_hi; r += kStepThreads) {
#pragma unroll
    for (int c = 0; c < B; ++c) Qn[(long long)c * n + r] = a.y_full[(long long)r * B + c];
  }
  __syncthreads();

  // Gram partial of this CTA: out[i][c] = sum_{r in slice} A[r][i] Qn[r][c], i < pa; A column-major (ld n).
  // Warp w owns the columns i = w, w + 8, ...; four of them per sweep over the slice (the B values of
  // a Qn row are loaded once per sweep and row), lanes stride over the rows, butterfly at the end.
  auto gram_partial = [&](const double* A, int pa) {
    double* out = a.Gpart + (long long)cta * pa * B;
    const int per_warp = (pa + 7) / 8;
    for (int ii0 = 0; ii0 < per_warp; ii0 += 4) {
      double acc[4][B];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < B; ++c) acc[u][c] = 0.0;
      for (int r = r_lo + lane; r < r_hi; r += 32) {
        double q[B];
#pragma unroll
        for (int c = 0; c < B; ++c) q[c] = Qn[(long long)c * n + r];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = warp + 8 * (ii0 + u);
          if (i < pa) {
            const double x = A[(long long)i * n + r];
#pragma unroll
            for (int c = 0; c < B; ++c) acc[u][c] += x * q[c];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = warp + 8 * (ii0 + u);
#pragma unroll
        for (int c = 0; c < B; ++c) acc[u][c] = warp_sum(acc[u][c]);
        if (i < pa && lane == 0) {
#pragma unroll
          for (int c = 0; c < B; ++c) out[i * B + c] = acc[u][c];
        }
      }
    }
  };
  // fixed-order reduction of the per-CTA partials into a.G, then into shared memory of every CTA
  auto gram_reduce = [&](int pa) {
    grid.sync();
    const int cnt = pa * B;
    for (int e = cta * kStepThreads + tid; e < cnt; e += G * kStepThreads) {
      double s = 0.0;
      for (int c = 0; c < G; ++c) s += a.Gpart[(long long)c * cnt + e];
      a.G[e] = s;
    }
    grid.sync();
    for (int e = tid; e < cnt; e += kStepThreads) s_g[e] = a.G[e];
    __syncthreads();
  };

  // ---- CGS2 against all previous blocks
  for (int pass = 0; pass < 2; ++pass) {
    gram_partial(a.Q, p);
    gram_reduce(p);
    for (int r = r_lo + tid; r < r_hi; r += kStepThreads) {
      double acc[B];
#pragma unroll
      for (int c = 0; c < B; ++c) acc[c] = 0.0;
      for (int i = 0; i < p; ++i) {
        const double x = a.Q[(long long)i * n + r];
#pragma unroll
        for (int c = 0; c < B; ++c) acc[c] += x * s_g[i * B + c];
      }
#pragma unroll
      for (int c = 0; c < B; ++c) Qn[(long long)c * n + r] -= acc[c];
    }
    if (cta == 0) {
      for (int e = tid; e < p * B; e += kStepThreads) {
        double* dst = &a.H[(long long)(e / B) * D + j * B + (e % B)];
        *dst = pass ? (*dst + s_g[e]) : s_g[e];
      }
    }
    __syncthreads();                                        // s_g is rewritten by the next reduction
  }
  // ---- CholQR2: Qn = Q_{j+1} (R2 R1)
  for (int pass = 0; pass < 2; ++pass) {
    gram_partial(Qn, B);
    gram_reduce(B);
    if (tid == 0) chol_small(s_g, B, s_r, s_ri);
    __syncthreads();
    if (pass == 0 && tid < B * B) s_r1[tid] = s_r[tid];
    const bool last = pass == 1;
    for (int r = r_lo + tid; r < r_hi; r += kStepThreads) {
      double in[B], out[B];
#pragma unroll
      for (int c = 0; c < B; ++c) in[c] = Qn[(long long)c * n + r];
#pragma unroll
      for (int c = 0; c < B; ++c) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < B; ++t) s += in[t] * s_ri[t * B + c];
        out[c] = s;
      }
#pragma unroll
      for (int c = 0; c < B; ++c) {
        Qn[(long long)c * n + r] = out[c];
        if (last) a.Vf[(long long)c * n + r] = (float)out[c];
      }
    }
    __syncthreads();
  }
  if (cta == 0 && tid < B * B) {                            // H[(j+1)-block, j-block] = R2 R1
    const int i = tid / B, c = tid % B;
    double s = 0.0;
    for (int t = 0; t < B; ++t) s += s_r[i * B + t] * s_r1[t * B + c];
    a.H[(long long)(p + i) * D + j * B + c] = s;
    a.R1[tid] = s_r1[tid];
    a.R2[tid] = s_r[tid];
  }
}

// grid of the cooperative step kernel: one CTA per SM at most (always co-resident), fewer for small n
int step_grid(int n) { return max(1, min(min(num_sms(), kMaxStepGrid), (n + 127) / 128)); }

template <int B>
int launch_step(const StepArgs& a, cudaStream_t s) {
  StepArgs args = a;
  const int G = step_grid(a.n);
  args.rows_per_cta = (a.n + G - 1) / G;
  void* params[] = {&args};
  HIA_CUDA(cudaLaunchCooperativeKernel((const void*)lanczos_step_kernel<B>, dim3(G), dim3(kStepThreads), params, 0, s));
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

struct EigWs {
  double *Q, *Y, *U, *H, *G, *G2, *R1, *Rinv, *R2, *Rt, *JA, *JV, *lam, *resid2, *Gpart;
  float *Vf;
  int* order;
  double* check;      // [2 * kMaxBlock] estimates + gaps of the last convergence check
  int* flag;
};

constexpr int kGramRowsPerCta = 1024;   // 8192 left <= 32 CTAs for the b x b Gram: 86 us, latency-bound
inline int gram_chunks(int n) { return (n + kGramRowsPerCta - 1) / kGramRowsPerCta; }

// G (p x b) = A[:, 0:p]^T B, deterministic: partials per row chunk + ordered reduction
int gram(const double* a, int p, const double* bm, int b, int n, double* gpart, double* g, cudaStream_t s) {
  dim3 grid(p, gram_chunks(n));
  gram_kernel<<<grid, 256, 0, s>>>(a, p, bm, b, n, kGramRowsPerCta, gpart);
  HIA_LAUNCH_CHECK();
  gram_reduce_kernel<<<(p * b + 255) / 256, 256, 0, s>>>(gpart, gram_chunks(n), p * b, g);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

long long carve(unsigned char* base, int n, int b, int steps, EigWs* w);

long long carve(unsigned char* base, int n, int b, int steps, EigWs* w) {
  const long long D = (long long)b * (steps + 1);
  long long off = 0;
  auto take = [&](long long bytes) { long long o = off; off += (bytes + 255) / 256 * 256; return base ? base + o : nullptr; };
  unsigned char* p;
  p = take((long long)n * D * 8); if (w) w->Q = (double*)p;
  p = take((long long)n * b * 8); if (w) w->Y = (double*)p;
  p = take((long long)n * b * 8); if (w) w->U = (double*)p;
  p = take((long long)n * b * 4); if (w) w->Vf = (float*)p;
  p = take(D * D * 8); if (w) w->H = (double*)p;
  p = take(D * b * 8); if (w) w->G = (double*)p;
  p = take(D * b * 8); if (w) w->G2 = (double*)p;
  p = take(b * b * 8); if (w) w->R1 = (double*)p;
  p = take(b * b * 8); if (w) w->Rinv = (double*)p;
  p = take(b * b * 8); if (w) w->R2 = (double*)p;
  p = take(b * b * 8); if (w) w->Rt = (double*)p;
  p = take(D * D * 8); if (w) w->JA = (double*)p;
  p = take(D * D * 8); if (w) w->JV = (double*)p;
  p = take(D * 8); if (w) w->lam = (double*)p;
  p = take(b * 8); if (w) w->resid2 = (double*)p;
  p = take(D * 4); if (w) w->order = (int*)p;
  p = take(2 * kMaxBlock * 8); if (w) w->check = (double*)p;
  p = take((long long)max(gram_chunks(n), kMaxStepGrid) * D * b * 8); if (w) w->Gpart = (double*)p;
  p = take(256); if (w) w->flag = (int*)p;
  return off + 256;
}

// The same pass with the matrix streamed by TMA: ONE lane per warp issues a 2-D bulk tensor copy of a
// 4-row x BOXC-column box of C into the warp's shared-memory ring (cp.async.bulk.tensor, completion on an
// mbarrier); the edge of the matrix is zero-filled by the copy engine, so the per-lane address / bounds
// arithmetic of the cp.async version (4 copies x ~7 instructions per lane and iteration, ~25 % of that
// kernel's instructions) disappears. Why it matters: inside the step the pass runs at the ~1.35 GHz the
// power-capped SYRK leaves behind, where the cp.async kernel's 63 % issue utilisation (at 1.9 GHz, ncu)
// becomes the limiter. The warp is its own producer and consumer: the box of iteration it + S - 1 is
// requested right after the lanes are done with iteration it - 1 (its slot), no empty barrier needed.
__device__ __forceinline__ uint32_t eig_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool eig_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(eig_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps after ~4 s instead of hanging the GPU box.
__device__ __forceinline__ void eig_mbar_wait(uint64_t* bar, uint32_t parity) {
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  while (!eig_mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fffu) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}

template <int B, int BOXC, int S>
__global__ void __launch_bounds__(kSymmWarps * 32, 2)
symm_panel_tma_kernel(const __grid_constant__ CUtensorMap tmap, int rows, int n, const float* __restrict__ v,
                      double* __restrict__ y, int row_major) {
  extern __shared__ __align__(128) unsigned char s_tma_raw[];
  constexpr int kBoxFloats = kSymmRowsPerWarp * BOXC;
  constexpr int kRingBytes = kSymmWarps * S * kBoxFloats * (int)sizeof(float);
  float* s_ring = reinterpret_cast<float*>(s_tma_raw);                          // [warp][stage][4][BOXC]
  float (*s_v)[kSymmCols] = reinterpret_cast<float (*)[kSymmCols]>(s_tma_raw + kRingBytes);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tma_raw + kRingBytes + (size_t)B * kSymmCols * sizeof(float));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * kSymmCols;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSymmWarps * S; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(eig_smem_u32(&s_bar[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = threadIdx.x; idx < B * kSymmCols; idx += blockDim.x) {
    const int j = idx / kSymmCols, col = idx % kSymmCols;
    s_v[j][col] = (k0 + col < n) ? v[(long long)j * n + k0 + col] : 0.f;
  }
  __syncthreads();
  float* ring = s_ring + (size_t)warp * S * kBoxFloats;
  uint64_t* bar = s_bar + warp * S;
  const int rbase = blockIdx.y * kSymmRowsPerCta + warp * (kSymmRowsPerCta / kSymmWarps);
  constexpr int kGroups = kSymmRowsPerCta / kSymmWarps / kSymmRowsPerWarp;      // 8 row groups per warp
  constexpr int kBoxes = kSymmCols / BOXC;                                     // boxes per row group
  constexpr int kIters = kGroups * kBoxes;
  auto issue = [&](int it) {
    if (lane == 0 && it < kIters) {
      const int g = it / kBoxes, bx = it % kBoxes;
      const int slot = it % S;
      const uint32_t b = eig_smem_u32(&bar[slot]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                   :: "r"(b), "r"(kBoxFloats * (int)sizeof(float)) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
          :: "r"(eig_smem_u32(ring + (size_t)slot * kBoxFloats)), "l"(&tmap), "r"(b),
             "r"(k0 + bx * BOXC), "r"(rbase + g * kSymmRowsPerWarp) : "memory");
    }
  };
#pragma unroll
  for (int it = 0; it < S - 1; ++it) issue(it);
  float2 acc[kSymmRowsPerWarp][B];
  for (int it = 0; it < kIters; ++it) {
    const int g = it / kBoxes, bx = it % kBoxes;
    if (bx == 0) {
#pragma unroll
      for (int r = 0; r < kSymmRowsPerWarp; ++r)
#pragma unroll
        for (int j = 0; j < B; ++j) acc[r][j] = make_float2(0.f, 0.f);
    }
    __syncwarp();                                        // every lane is done with the slot of iteration it - 1
    issue(it + S - 1);
    eig_mbar_wait(&bar[it % S], (uint32_t)((it / S) & 1));
    const float* src = ring + (size_t)(it % S) * kBoxFloats + lane * 4;
#pragma unroll 1
    for (int sub = 0; sub < BOXC / 128; ++sub) {
      float4 a[kSymmRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kSymmRowsPerWarp; ++r) a[r] = *reinterpret_cast<const float4*>(src + r * BOXC + sub * 128);
      const int cc = bx * BOXC + sub * 128 + lane * 4;
#pragma unroll
      for (int j = 0; j < B; ++j) {
        const float4 vv = *reinterpret_cast<const float4*>(&s_v[j][cc]);
#pragma unroll
        for (int r = 0; r < kSymmRowsPerWarp; ++r) {
          acc[r][j] = ffma2(make_float2(a[r].x, a[r].y), make_float2(vv.x, vv.y), acc[r][j]);
          acc[r][j] = ffma2(make_float2(a[r].z, a[r].w), make_float2(vv.z, vv.w), acc[r][j]);
        }
      }
    }
    if (bx == kBoxes - 1) {
      const int r0 = rbase + g * kSymmRowsPerWarp;
      float part[kSymmRowsPerWarp * B];
#pragma unroll
      for (int r = 0; r < kSymmRowsPerWarp; ++r)
#pragma unroll
        for (int j = 0; j < B; ++j) part[r * B + j] = acc[r][j].x + acc[r][j].y;
      const float total = lane_transpose_sum<kSymmRowsPerWarp * B>(part, lane);
      const int idx = lane % (kSymmRowsPerWarp * B), r = idx / B, j = idx % B;
      if (lane < kSymmRowsPerWarp * B && r0 + r < rows) {
        double* dstp = row_major ? &y[(long long)(r0 + r) * B + j] : &y[(long long)j * rows + r0 + r];
        atomicAdd(dstp, (double)total);
      }
    }
  }
}

typedef CUresult (*PFN_eigEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// rows x n fp32 matrix with row stride ld, boxes of kSymmRowsPerWarp rows x boxc columns, no swizzle, zero fill
int make_matrix_map(CUtensorMap* map, const float* c, int rows, int n, long long ld, int boxc) {
  static PFN_eigEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !p) { set_last_cuda_error(e); return HIA_ERR_CUDA; }
    fn = reinterpret_cast<PFN_eigEncodeTiled>(p);
  }
  cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)boxc, (cuuint32_t)kSymmRowsPerWarp};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(c), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_cuda_error(cudaErrorInvalidValue); return HIA_ERR_CUDA; }
  return HIA_OK;
}

template <int B, int BOXC, int S>
int launch_symm_tma(const float* c, int rows, int n, long long ld, const float* v, double* y, int row_major,
                    dim3 grid, cudaStream_t s) {
  constexpr int smem = kSymmWarps * S * kSymmRowsPerWarp * BOXC * (int)sizeof(float) +
                       B * kSymmCols * (int)sizeof(float) + kSymmWarps * S * (int)sizeof(uint64_t);
  static bool attr_set = false;                            // once per process and instantiation
  if (!attr_set) {
    HIA_CUDA(cudaFuncSetAttribute(symm_panel_tma_kernel<B, BOXC, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  CUtensorMap map;
  int st = make_matrix_map(&map, c, rows, n, ld, BOXC);
  if (st != HIA_OK) return st;
  symm_panel_tma_kernel<B, BOXC, S><<<grid, kSymmWarps * 32, smem, s>>>(map, rows, n, v, y, row_major);
  count_launch();
  return cuda_ok(cudaGetLastError());
}

template <int B>
int launch_symm(const float* c, int rows, int n, long long ld, const float* v, double* y, int row_major,
                cudaStream_t s) {
  HIA_CUDA(cudaMemsetAsync(y, 0, (size_t)rows * B * sizeof(double), s));
  dim3 grid((n + kSymmCols - 1) / kSymmCols, (rows + kSymmRowsPerCta - 1) / kSymmRowsPerCta);
  const int smem = B * kSymmCols * (int)sizeof(float);
  const int smem_async = smem + kSymmWarps * kAsyncStages * kSymmRowsPerWarp * 32 * (int)sizeof(float4);
  static bool attr_set = false;                            // once per process and instantiation
  if (!attr_set) {
    HIA_CUDA(cudaFuncSetAttribute(symm_panel_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    HIA_CUDA(cudaFuncSetAttribute(symm_panel_async_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_async));
    attr_set = true;
  }
  const char* e = getenv("HIA_EIG_ASYNC");                 // read per call: tools time both in one process
  const bool use_async = !(e && atoi(e) == 0) && ((reinterpret_cast<uintptr_t>(c) & 15u) == 0) && (ld & 3) == 0;
  // TMA boxes of 4 rows x 256 columns, 2 slots per warp (default). Measured inside the step at 30 894 bins
  // (eigen stage, 8 passes): cp.async ring 8.65 ms, this 7.65 ms; 4 x 128 boxes with 4 slots 8.63 ms; 3 slots
  // of 4 x 256 (one CTA per SM) 9.97 ms. HIA_EIG_TMA=0 keeps the cp.async ring.
  const char* et = getenv("HIA_EIG_TMA");
  if (use_async && !(et && atoi(et) == 0))
    return launch_symm_tma<B, 256, 2>(c, rows, n, ld, v, y, row_major, grid, s);
  if (use_async)
    symm_panel_async_kernel<B><<<grid, kSymmWarps * 32, smem_async, s>>>(c, rows, n, ld, v, y, row_major);
  else
    symm_panel_kernel<B><<<grid, kSymmWarps * 32, smem, s>>>(c, rows, n, ld, v, y, row_major);
  count_launch();
  return cuda_ok(cudaGetLastError());
}
int symm(const float* c, int rows, int n, long long ld, const float* v, double* y, int b, int row_major,
         cudaStream_t s) {
  switch (b) {
    case 4: return launch_symm<4>(c, rows, n, ld, v, y, row_major, s);
    case 8: return launch_symm<8>(c, rows, n, ld, v, y, row_major, s);
    default: return HIA_ERR_UNSUPPORTED;
  }
}

// dst (col-major n x b) <- src (row-major n x b)
__global__ void rowmajor_to_colmajor_kernel(const double* __restrict__ src, double* __restrict__ dst, int n, int b) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * b) return;
  const int j = (int)(i / n), r = (int)(i % n);
  dst[i] = src[(long long)r * b + j];
}

// Orthonormalise Y (n x b) in place by CholQR; returns R in r_out (b x b), fp32 copy in yf.
int cholqr(double* Y, int n, int b, EigWs& w, double* r_out, float* yf, cudaStream_t s) {
  int gst = gram(Y, b, Y, b, n, w.Gpart, w.G2, s);
  if (gst != HIA_OK) return gst;
  chol_kernel<<<1, 32, 0, s>>>(w.G2, b, r_out, w.Rinv);
  HIA_LAUNCH_CHECK();
  panel_mul_kernel<<<(n + 255) / 256, 256, 0, s>>>(Y, w.Rinv, b, n, yf);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

int launch_jacobi(const EigWs& w, int D, int dim, cudaStream_t s, int kk) {
  const char* e = getenv("HIA_EIG_JACOBI");               // read per call: tools time both in one process
  const bool want_jacobi = e && atoi(e) == 1;
  if (!want_jacobi && dim <= kTriSmemDim && kk <= kTriMaxVec) {
    static bool tri_attr_set = false;
    if (!tri_attr_set) {
      HIA_CUDA(cudaFuncSetAttribute(tri_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kTriSmemDim * kTriSmemDim * (int)sizeof(double)));
      tri_attr_set = true;
    }
    tri_eig_kernel<<<1, 1024, (size_t)dim * dim * sizeof(double), s>>>(w.H, D, dim, kk, w.JV, w.lam, w.order);
    HIA_LAUNCH_CHECK();
    return HIA_OK;
  }
  const size_t need = (size_t)2 * dim * dim * sizeof(double);
  const int use_smem = need <= 222 * 1024;
  static bool attr_set = false;                            // once: the largest size that is ever requested
  if (use_smem && !attr_set) {
    HIA_CUDA(cudaFuncSetAttribute(jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
    attr_set = true;
  }
  jacobi_kernel<<<1, 1024, use_smem ? need : 0, s>>>(w.H, D, dim, w.JA, w.JV, w.lam, w.order, 15, use_smem);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

int check_shape(int n, int block, int steps) {
  if (n <= 0 || steps <= 0) return HIA_ERR_BAD_ARG;
  if (block != 4 && block != 8) return HIA_ERR_UNSUPPORTED;
  if ((long long)block * (steps + 1) > n) return HIA_ERR_BAD_ARG;       // Krylov space larger than n
  if ((long long)block * (steps + 1) > 272) return HIA_ERR_UNSUPPORTED;
  return HIA_OK;
}

}  // namespace
}  // namespace hia

HIA_API int64_t hia_topk_eig_workspace_bytes(int32_t n, int32_t block, int32_t steps) {
  if (n <= 0 || block <= 0 || steps <= 0) return 0;
  return hia::carve(nullptr, n, block, steps, nullptr);
}

// ---- phase 0: start block (caller-provided vectors for a restart + random columns), orthonormal
HIA_API int hia_eig_begin(int32_t n, int32_t block, int32_t steps, uint64_t seed, const float* init_vecs,
                          int32_t n_init, void* workspace, int64_t workspace_bytes, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int st = check_shape(n, block, steps);
  if (st != HIA_OK) return st;
  if (!workspace || n_init < 0 || n_init > block || (n_init > 0 && !init_vecs)) return HIA_ERR_BAD_ARG;
  if (workspace_bytes < hia_topk_eig_workspace_bytes(n, block, steps)) return HIA_ERR_WORKSPACE;
  const int b = block, D = b * (steps + 1);
  EigWs w;
  carve(static_cast<unsigned char*>(workspace), n, b, steps, &w);
  random_panel_kernel<<<(int)(((long long)n * b + 255) / 256), 256, 0, s>>>(w.Q, n, b, seed);
  HIA_LAUNCH_CHECK();
  if (n_init > 0) {
    float_to_double_panel_kernel<<<(int)(((long long)n * n_init + 255) / 256), 256, 0, s>>>(init_vecs, w.Q, (long long)n * n_init);
    HIA_LAUNCH_CHECK();
  }
  st = cholqr(w.Q, n, b, w, w.R1, nullptr, s);
  if (st != HIA_OK) return st;
  st = cholqr(w.Q, n, b, w, w.R1, w.Vf, s);
  if (st != HIA_OK) return st;
  HIA_CUDA(cudaMemsetAsync(w.H, 0, (size_t)D * D * 8, s));
  return HIA_OK;
}

// ---- the only pass over the matrix: y_rows (rows x block, row-major fp64) = C[rows, :] * panel,
// where the panel is the current fp32 block held in the workspace (Q_j, or the Ritz vectors).
HIA_API int hia_eig_symm_rows(const float* c_rows, int32_t rows, int32_t n, int64_t ld, int32_t block,
                              int32_t steps, void* workspace, double* y_rows, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (!c_rows || !workspace || !y_rows || rows <= 0 || rows > n || ld < n) return HIA_ERR_BAD_ARG;
  if (!aligned16(c_rows) || (ld & 3)) return HIA_ERR_ALIGN;
  int st = check_shape(n, block, steps);
  if (st != HIA_OK) return st;
  EigWs w;
  carve(static_cast<unsigned char*>(workspace), n, block, steps, &w);
  return symm(c_rows, rows, n, ld, w.Vf, y_rows, block, 1, s);
}

// ---- step j: Y (n x block, row-major: all rows gathered) -> H[:, j], Q_{j+1}, next fp32 panel
HIA_API int hia_eig_absorb(int32_t n, int32_t block, int32_t steps, int32_t j, const double* y_full,
                           void* workspace, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int st = check_shape(n, block, steps);
  if (st != HIA_OK) return st;
  if (!y_full || !workspace || j < 0 || j >= steps) return HIA_ERR_BAD_ARG;
  const int b = block, D = b * (steps + 1);
  EigWs w;
  carve(static_cast<unsigned char*>(workspace), n, b, steps, &w);
  const char* env_coop = getenv("HIA_EIG_COOP");         // read per call: tools time both in one process
  const bool coop = !(env_coop && atoi(env_coop) == 0);
  if (coop) {                                           // one cooperative kernel for the whole step
    StepArgs a;
    a.y_full = y_full; a.Q = w.Q; a.H = w.H; a.Gpart = w.Gpart; a.G = w.G; a.R1 = w.R1; a.R2 = w.R2; a.Vf = w.Vf;
    a.n = n; a.b = b; a.D = D; a.j = j; a.rows_per_cta = 0;
    return b == 8 ? launch_step<8>(a, s) : launch_step<4>(a, s);
  }
  double* Qn = w.Q + (long long)(j + 1) * b * n;
  const int p = (j + 1) * b;                         // columns of Q so far
  rowmajor_to_colmajor_kernel<<<(int)(((long long)n * b + 255) / 256), 256, 0, s>>>(y_full, Qn, n, b);
  HIA_LAUNCH_CHECK();
  for (int pass = 0; pass < 2; ++pass) {             // CGS2 against all previous blocks
    st = gram(w.Q, p, Qn, b, n, w.Gpart, w.G, s);
    if (st != HIA_OK) return st;
    panel_sub_kernel<<<(n + 255) / 256, 256, (size_t)p * b * 8, s>>>(w.Q, p, w.G, b, n, Qn);
    HIA_LAUNCH_CHECK();
    place_block_kernel<<<(p * b + 255) / 256, 256, 0, s>>>(w.G, p, b, w.H, D, 0, j * b, pass);
    HIA_LAUNCH_CHECK();
  }
  st = cholqr(Qn, n, b, w, w.R1, nullptr, s);        // CholQR2: Y = Q_{j+1} (R2 R1)
  if (st != HIA_OK) return st;
  st = cholqr(Qn, n, b, w, w.R2, w.Vf, s);
  if (st != HIA_OK) return st;
  small_matmul_kernel<<<1, 32, 0, s>>>(w.R2, w.R1, b, w.Rt);
  HIA_LAUNCH_CHECK();
  place_block_kernel<<<(b * b + 255) / 256, 256, 0, s>>>(w.Rt, b, b, w.H, D, (j + 1) * b, j * b, 0);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

// ---- convergence check after `steps_done` absorbed steps, without a pass over the matrix: Jacobi on
// the (steps_done * block)-dimensional projected matrix + the Lanczos residual estimates. Writes
// *converged_host (synchronises the stream); est_gap_host (optional) receives k estimates + k gaps.
HIA_API int hia_eig_check(int32_t n, int32_t block, int32_t steps, int32_t steps_done, int32_t k, double rel_tol,
                          double angle_tol, int32_t* converged_host, double* est_gap_host, void* workspace,
                          void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int st = check_shape(n, block, steps);
  if (st != HIA_OK) return st;
  if (!converged_host || !workspace || k <= 0 || k > block || steps_done < 1 || steps_done > steps)
    return HIA_ERR_BAD_ARG;
  const int b = block, D = b * (steps + 1), dim = steps_done * b;
  EigWs w;
  carve(static_cast<unsigned char*>(workspace), n, b, steps, &w);
  st = launch_jacobi(w, D, dim, s, b);
  if (st != HIA_OK) return st;
  ritz_check_kernel<<<1, 32, 0, s>>>(w.H, D, dim, b, w.JV, w.lam, w.order, k, rel_tol, angle_tol, w.check, w.flag);
  HIA_LAUNCH_CHECK();
  int flag = 0;
  HIA_CUDA(cudaMemcpyAsync(&flag, w.flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  if (est_gap_host) HIA_CUDA(cudaMemcpyAsync(est_gap_host, w.check, (size_t)2 * k * 8, cudaMemcpyDeviceToHost, s));
  HIA_CUDA(cudaStreamSynchronize(s));
  *converged_host = flag;
  return HIA_OK;
}

// ---- Rayleigh-Ritz on the first `steps_used` blocks: eigvals[k], eigvecs[k][n], gaps[k] (optional:
// distance to the nearest other Ritz value); the fp32 panel becomes [U, 0] for the residual pass
HIA_API int hia_eig_ritz(int32_t n, int32_t block, int32_t steps, int32_t steps_used, int32_t reuse_check,
                         int32_t k, double* eigvals, float* eigvecs, double* gaps, void* workspace,
                         void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int st = check_shape(n, block, steps);
  if (st != HIA_OK) return st;
  if (!eigvals || !eigvecs || !workspace || k <= 0 || k > block || steps_used < 1 || steps_used > steps)
    return HIA_ERR_BAD_ARG;
  const int b = block, D = b * (steps + 1), dim = steps_used * b;
  EigWs w;
  carve(static_cast<unsigned char*>(workspace), n, b, steps, &w);
  if (!reuse_check) {          // reuse_check: the last hia_eig_check already decomposed exactly this matrix
    st = launch_jacobi(w, D, dim, s, b);
    if (st != HIA_OK) return st;
  }
  if (gaps) {
    ritz_gap_kernel<<<1, 32, 0, s>>>(w.lam, w.order, dim, k, gaps);
    HIA_LAUNCH_CHECK();
  }
  ritz_kernel<<<(n + 255) / 256, 256, 0, s>>>(w.Q, n, dim, w.JV, w.order, w.lam, k, w.U, w.Vf, eigvals);
  HIA_LAUNCH_CHECK();
  HIA_CUDA(cudaMemcpyAsync(eigvecs, w.Vf, (size_t)n * k * 4, cudaMemcpyDeviceToDevice, s));
  if (k < b) HIA_CUDA(cudaMemsetAsync(w.Vf + (long long)k * n, 0, (size_t)(b - k) * n * 4, s));
  return HIA_OK;
}

// ---- row-sharded callers: install the Ritz vectors chosen by ONE rank (broadcast by the caller) as
// the current panel [U, 0] and as U for the residual. Every rank decomposes its own replicated copy
// of the projected matrix; with clustered Ritz values the order (or the rotation inside a cluster)
// can differ between ranks by rounding alone, and panels that differ across ranks turn the
// all-gathered product into garbage.
__global__ void set_ritz_kernel(const float* __restrict__ src, long long count, double* __restrict__ u,
                                float* __restrict__ vf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) { const float v = src[i]; u[i] = (double)v; vf[i] = v; }
}
HIA_API int hia_eig_set_ritz(int32_t n, int32_t block, int32_t steps, int32_t k, const float* eigvecs,
                             void* workspace, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int st = check_shape(n, block, steps);
  if (st != HIA_OK) return st;
  if (!eigvecs || !workspace || k <= 0 || k > block) return HIA_ERR_BAD_ARG;
  EigWs w;
  carve(static_cast<unsigned char*>(workspace), n, block, steps, &w);
  const long long count = (long long)n * k;
  set_ritz_kernel<<<(int)((count + 255) / 256), 256, 0, s>>>(eigvecs, count, w.U, w.Vf);
  HIA_LAUNCH_CHECK();
  if (k < block) HIA_CUDA(cudaMemsetAsync(w.Vf + (long long)k * n, 0, (size_t)(block - k) * n * 4, s));
  return HIA_OK;
}

// ---- residuals ||C u - lambda u|| from y_full = C * [U, 0] (n x block, row-major)
HIA_API int hia_eig_resid(int32_t n, int32_t block, int32_t steps, int32_t k, const double* y_full,
                          const double* eigvals, double* resid, void* workspace, void* stream_) {
  using namespace hia;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  int st = check_shape(n, block, steps);
  if (st != HIA_OK) return st;
  if (!y_full || !eigvals || !resid || !workspace || k <= 0 || k > block) return HIA_ERR_BAD_ARG;
  EigWs w;
  carve(static_cast<unsigned char*>(workspace), n, block, steps, &w);
  rowmajor_to_colmajor_kernel<<<(int)(((long long)n * block + 255) / 256), 256, 0, s>>>(y_full, w.Y, n, block);
  HIA_LAUNCH_CHECK();
  HIA_CUDA(cudaMemsetAsync(w.resid2, 0, (size_t)block * 8, s));
  resid_kernel<<<dim3(64, k), 256, 0, s>>>(w.Y, w.U, eigvals, n, k, w.resid2);
  HIA_LAUNCH_CHECK();
  sqrt_kernel<<<1, 32, 0, s>>>(w.resid2, k);
  HIA_LAUNCH_CHECK();
  HIA_CUDA(cudaMemcpyAsync(resid, w.resid2, (size_t)k * 8, cudaMemcpyDeviceToDevice, s));
  return HIA_OK;
}

// Cadence of the pass-free convergence checks. A check solves the projected eigenproblem and reads one
// flag back (~0.4 ms with the partial solver); a Lanczos step costs a pass over the matrix (0.7 ms at
// 30 894 bins, ~20 us at 2 000). Large matrices: first check after min(7, 2/3 of the steps) steps, then
// after EVERY step (measured at 30 894 bins: first check at 5 / 6 / 7 -> 9.5 / 9.1 / 8.5 ms per solve, against
// 10.0 ms with "step 8, then every second step"). Small matrices (chromosome-sized chains, where the
// checks are the cost): step min(8, 2/3 of the steps), then every second step.
HIA_API int32_t hia_eig_first_check(int32_t n, int32_t steps) {
  const char* e = getenv("HIA_EIG_FIRST_CHECK");        // tuning knob (tools)
  if (e && atoi(e) > 0) return atoi(e);
  const int cap = n >= 16384 ? 7 : 8;
  const int f = steps * 2 / 3 < cap ? steps * 2 / 3 : cap;
  return f > 4 ? f : 4;
}
HIA_API int32_t hia_eig_check_every(int32_t n) {
  const char* e = getenv("HIA_EIG_CHECK_EVERY");
  if (e && atoi(e) > 0) return atoi(e);
  return n >= 16384 ? 1 : 2;
}

// ---- single-device driver: all phases on one stream. early_rel_tol / early_angle_tol > 0 enable the
// pass-free convergence check (hia_eig_check, one tiny D2H flag) after every second step from
// hia_eig_first_check(steps) on: the cycle stops as soon as the Lanczos residual estimates meet the tolerance instead of
// always spending `steps` passes over the matrix. *steps_used_host (optional) = steps actually taken.
// The projected (Rayleigh-Ritz) eigenproblem on its own, for tests and tools: h = D x D row-major, the
// lower triangle of its leading dim x dim block is the symmetric matrix. Outputs (device): z (dim x dim
// doubles, column-major: column j = eigenvector of the j-th largest eigenvalue for j < kk; the partial
// solver fills only those), lam[dim], order[dim] (consumers address pair j as lam[order[j]] / column
// order[j]). scratch: dim x dim doubles. solver 0 = default (partial solver up to dim 160), 1 = Jacobi.
HIA_API int hia_projected_eig(const double* h, int32_t D, int32_t dim, int32_t kk, int solver, double* z,
                              double* lam, int32_t* order, double* scratch, void* stream_) {
  using namespace hia;
  if (!h || !z || !lam || !order || !scratch || dim <= 0 || dim > D || dim > kTriMaxDim || kk <= 0 || kk > kMaxBlock)
    return HIA_ERR_BAD_ARG;
  EigWs w = {};
  w.H = const_cast<double*>(h);
  w.JA = scratch;
  w.JV = z;
  w.lam = lam;
  w.order = order;
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  if (solver == 1) {
    const size_t need = (size_t)2 * dim * dim * sizeof(double);
    const int use_smem = need <= 222 * 1024;
    if (use_smem) HIA_CUDA(cudaFuncSetAttribute(jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024));
    jacobi_kernel<<<1, 1024, use_smem ? need : 0, s>>>(w.H, D, dim, w.JA, w.JV, w.lam, w.order, 15, use_smem);
    HIA_LAUNCH_CHECK();
    return HIA_OK;
  }
  return launch_jacobi(w, D, dim, s, kk);
}

HIA_API int64_t hia_eig_symm_upper_workspace_bytes(int32_t n, int32_t block) {
  if (n <= 0 || (block != 4 && block != 8)) return 0;
  return hia::symm_upper_ws_bytes(n, block);
}

// upper_ws (optional, hia_eig_symm_upper_workspace_bytes): the matrix is SYMMETRIC and only its upper
// triangle is read by the passes (half the DRAM bytes; the lower triangle is never touched).
// k_keep (k <= k_keep <= block; 0 = k): Ritz pairs RETURNED (eigvals / eigvecs / resid / gaps hold k_keep
// entries). Convergence is judged on the first k only; the extra pairs make better restart blocks --
// on a clustered spectrum (25 kb count map: 1410.5 / 1404.5 / 1392.2) a restart from k = 3 Ritz vectors
// plus random columns throws away what the cycle learnt about the next eigenvalues.
HIA_API int hia_topk_eig(const float* c, int32_t n, int64_t ld, int32_t k, int32_t k_keep, int32_t block,
                         int32_t steps, uint64_t seed, const float* init_vecs, int32_t n_init, double early_rel_tol,
                         double early_angle_tol, double* eigvals, float* eigvecs, double* resid, double* gaps,
                         int32_t* steps_used_host, void* workspace, int64_t workspace_bytes, void* upper_ws,
                         int64_t upper_ws_bytes, void* stream) {
  using namespace hia;
  if (!c || !eigvals || !eigvecs || !resid || !workspace || ld < n || k <= 0) return HIA_ERR_BAD_ARG;
  if (k_keep <= 0) k_keep = k;
  if (k > block || k_keep < k || k_keep > block) return HIA_ERR_BAD_ARG;
  if (upper_ws && upper_ws_bytes < hia_eig_symm_upper_workspace_bytes(n, block)) return HIA_ERR_WORKSPACE;
  if (upper_ws && (!aligned16(upper_ws) || !aligned16(c) || (ld & 3))) return HIA_ERR_ALIGN;
  int st = hia_eig_begin(n, block, steps, seed, init_vecs, n_init, workspace, workspace_bytes, stream);
  if (st != HIA_OK) return st;
  EigWs w;
  carve(static_cast<unsigned char*>(workspace), n, block, steps, &w);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto pass = [&](double* y) -> int {                 // y (n x block, row-major fp64) = C * current fp32 panel
    if (upper_ws)
      return block == 8 ? launch_symm_upper<8>(c, n, ld, w.Vf, y, upper_ws, s)
                        : launch_symm_upper<4>(c, n, ld, w.Vf, y, upper_ws, s);
    return hia_eig_symm_rows(c, n, n, ld, block, steps, workspace, y, stream);
  };
  double* ystage = w.Y;                               // n x block staging (row-major)
  const bool early = early_rel_tol > 0.0 || early_angle_tol > 0.0;
  int used = steps, reuse = 0;
  for (int j = 0; j < steps; ++j) {
    st = pass(ystage);
    if (st != HIA_OK) return st;
    st = hia_eig_absorb(n, block, steps, j, ystage, workspace, stream);
    if (st != HIA_OK) return st;
    const int done = j + 1;
    if (early && done >= hia_eig_first_check(n, steps) &&
        ((done - hia_eig_first_check(n, steps)) % hia_eig_check_every(n)) == 0 && done < steps) {
      int conv = 0;
      st = hia_eig_check(n, block, steps, done, k, early_rel_tol, early_angle_tol, &conv, nullptr, workspace, stream);
      if (st != HIA_OK) return st;
      if (conv) { used = done; reuse = 1; break; }
    }
  }
  if (steps_used_host) *steps_used_host = used;
  st = hia_eig_ritz(n, block, steps, used, reuse, k_keep, eigvals, eigvecs, gaps, workspace, stream);
  if (st != HIA_OK) return st;
  // the residual pass needs a second staging buffer: w.Y is the transposed target of hia_eig_resid
  double* ystage2 = w.Q + (long long)used * block * n;    // Q_{used} is no longer needed
  st = pass(ystage2);
  if (st != HIA_OK) return st;
  return hia_eig_resid(n, block, steps, k_keep, ystage2, eigvals, resid, workspace, stream);
}
