// a1/a2: sparse (i, j, v) -> dense symmetric fp32 map.
//
// Replaces read_sps_file -> coo_matrix -> tocsr (duplicates summed), SpsHiCMtx.to_all
// (triu(M,0).T + triu(M,1)) and SpsHiCMtx.to_dense (absent cells 0, or min(data) when
// has_neg; explicit zero entries stay 0) -- publish/new_hic_class.py:2576-2602, :2229-2246,
// :2209-2224. Per-block min fill follows yield_by_chro(triu=True)+to_dense as used by
// complex/src/main_local.py:112-115.
//
// HBM-bound: 12 B/nnz read + 4 B/cell written (fill) + 4..8 B/nnz scattered.
// Zero fill, entries sorted by row (every COO that comes out of a CSR matrix or a .mtx file the
//             pipeline writes; detected on the device, no host round trip):
//             row_ptr_kernel (reads the row indices once: 4 B/nnz) -> fill_rows_sorted_kernel: one
//             CTA per (row, 16 K-column segment) zeroes a shared-memory row buffer, adds the row's
//             entries into it (shared-memory atomics: duplicates are summed) and writes the segment
//             with coalesced stores (8 B/nnz read + 4 B per upper cell written; no memset, no
//             read-modify-write of DRAM sectors) -> tiled mirror (triu input).
// Zero fill, any order:  zero fill -> RED.ADD scatter of the direct entries -> tiled mirror.
// Min fill:   NaN sentinel fill -> store 0 at present cells -> RED.ADD scatter -> min over the
//             *scattered* cells (so duplicates are summed before the min, like coo->csr then
//             np.min(data)) -> replace the remaining sentinels by the (per-block) min -> mirror.
// Duplicates are summed with fp32 RED atomics; with no duplicates the result is bit-exact
// (0 + v == v), which is the case for every file the pipeline writes.
#include "hia_common.cuh"
#include <stdlib.h>

namespace hia {
namespace {

constexpr int kMaxChr = 1024;

__device__ __forceinline__ int find_chr(const int* s_start, int n_chr, int bin) {
  // largest c with s_start[c] <= bin
  int lo = 0, hi = n_chr - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (s_start[mid] <= bin) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void init_keys_kernel(unsigned int* keys, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) keys[i] = 0xffffffffu;
}

// min over valid entries, per chromosome-pair block (BLOCK_MIN) and overall (slot C*C).
__global__ void __launch_bounds__(256)
block_min_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
                 const float* __restrict__ dense, long long ld, long long nnz, int n, int mtx_type,
                 const int* __restrict__ chr_start, int n_chr, int per_block,
                 unsigned int* __restrict__ keys) {
  extern __shared__ int s_start[];
  for (int i = threadIdx.x; i <= n_chr; i += blockDim.x) s_start[i] = chr_start[i];
  __syncthreads();
  const int slot_all = n_chr * n_chr;
  int cur_slot = -1;
  unsigned int cur_key = 0xffffffffu;
  unsigned int all_key = 0xffffffffu;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    int r = rows[e], c = cols[e];
    if (r < 0 || c < 0 || r >= n || c >= n) continue;
    if (mtx_type == HIA_MTX_TRIU && c < r) continue;
    unsigned int k = float_to_ordered(dense[(long long)r * ld + c]);   // summed value
    all_key = min(all_key, k);
    if (per_block) {
      int slot = find_chr(s_start, n_chr, r) * n_chr + find_chr(s_start, n_chr, c);
      if (slot != cur_slot) {
        if (cur_slot >= 0) atomicMin(&keys[cur_slot], cur_key);
        cur_slot = slot;
        cur_key = k;
      } else {
        cur_key = min(cur_key, k);
      }
    }
  }
  if (per_block && cur_slot >= 0) atomicMin(&keys[cur_slot], cur_key);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) all_key = min(all_key, __shfl_xor_sync(0xffffffffu, all_key, o));
  if ((threadIdx.x & 31) == 0 && all_key != 0xffffffffu) atomicMin(&keys[slot_all], all_key);
}

__global__ void __launch_bounds__(256)
fill_const_kernel(float* __restrict__ out, int n, long long ld, float value) {
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= n) return;
  for (int r = blockIdx.y; r < n; r += gridDim.y) {
    float* p = out + (long long)r * ld + c0;
    if (c0 + 3 < n) {
      *reinterpret_cast<float4*>(p) = make_float4(value, value, value, value);
    } else {
      for (int j = 0; j < 4 && c0 + j < n; ++j) p[j] = value;
    }
  }
}

// Replace the NaN sentinels (absent cells) by fill(r, c). One thread per 4 columns.
__global__ void __launch_bounds__(256)
fixup_kernel(float* __restrict__ out, int n, long long ld, int fill_mode, int mtx_type,
             const int* __restrict__ chr_start, int n_chr, const unsigned int* __restrict__ keys) {
  extern __shared__ int s_start[];
  for (int i = threadIdx.x; i <= n_chr; i += blockDim.x) s_start[i] = chr_start[i];
  __syncthreads();
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= n) return;
  float gfill = 0.f;
  if (fill_mode == HIA_FILL_MIN) {
    unsigned int k = keys[n_chr * n_chr];
    gfill = (k == 0xffffffffu) ? 0.f : ordered_to_float(k);
  }
  int cb[4];
  if (fill_mode == HIA_FILL_BLOCK_MIN) {
#pragma unroll
    for (int j = 0; j < 4; ++j) cb[j] = find_chr(s_start, n_chr, min(c0 + j, n - 1));
  }
  for (int r = blockIdx.y; r < n; r += gridDim.y) {
    float* p = out + (long long)r * ld + c0;
    float v[4];
    const bool full = c0 + 3 < n;
    if (full) {
      const float4 t = *reinterpret_cast<const float4*>(p);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      for (int j = 0; j < 4; ++j) v[j] = (c0 + j < n) ? p[j] : 0.f;
    }
    if (!(v[0] != v[0] || v[1] != v[1] || v[2] != v[2] || v[3] != v[3])) continue;
    const int a = (fill_mode == HIA_FILL_BLOCK_MIN) ? find_chr(s_start, n_chr, r) : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (v[j] == v[j]) continue;
      float f = gfill;
      if (fill_mode == HIA_FILL_BLOCK_MIN) {
        const int b = cb[j];
        // triu input: only upper blocks hold entries; the lower block mirrors the upper one
        const int slot = (mtx_type == HIA_MTX_TRIU && b < a) ? (b * n_chr + a) : (a * n_chr + b);
        const unsigned int k = keys[slot];
        f = (k == 0xffffffffu) ? 0.f : ordered_to_float(k);
      }
      v[j] = f;
    }
    if (full) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int j = 0; j < 4 && c0 + j < n; ++j) p[j] = v[j];
    }
  }
}

// Direct entries only: out[r][c] (+)= v. For a 'triu' input the lower triangle is produced
// afterwards by mirror_kernel (coalesced), not by a second, uncoalesced atomic per entry.
template <bool kAdd>
__global__ void __launch_bounds__(256)
scatter_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
               const float* __restrict__ vals, long long nnz, int n, int mtx_type,
               float* __restrict__ out, long long ld, const int* __restrict__ only_if_flag) {
  if (only_if_flag && !*only_if_flag) return;              // the sorted-row path did the work
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    int r = rows[e], c = cols[e];
    if (r < 0 || c < 0 || r >= n || c >= n) continue;
    if (mtx_type == HIA_MTX_TRIU && c < r) continue;
    if (kAdd) atomicAdd(out + (long long)r * ld + c, vals[e]);
    else out[(long long)r * ld + c] = 0.f;
  }
}

// ---- sorted-row fast path ------------------------------------------------------------------
// row_ptr[q] = first entry of row q (q in [0, n]); *flag |= 1 when the row indices are not
// non-decreasing or out of range (the caller-independent fallback kernels then do the work).
__global__ void __launch_bounds__(256)
row_ptr_kernel(const int* __restrict__ rows, long long nnz, int n, long long* __restrict__ row_ptr,
               int* __restrict__ flag) {
  // 4 consecutive entries per thread (one 128-bit load + the entry before them)
  const long long groups = (nnz + 3) >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  bool bad = false;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const long long e0 = g << 2;
    int rv[4];
    if (e0 + 3 < nnz) {
      const int4 t = *reinterpret_cast<const int4*>(rows + e0);     // cudaMalloc'd arrays: 16-byte aligned
      rv[0] = t.x; rv[1] = t.y; rv[2] = t.z; rv[3] = t.w;
    } else {
      for (int i = 0; i < 4; ++i) rv[i] = (e0 + i < nnz) ? rows[e0 + i] : 0;
    }
    // (taking the entry before the group from the neighbouring lane's registers -- a shuffle instead of this
    //  scalar load -- measured 0.1 ms SLOWER in-step: the load overlaps the 128-bit one, the shuffle waits for it)
    int prev = (e0 > 0) ? rows[e0 - 1] : -1;
    // the common group: all four entries continue the previous entry's row and none of them is the last entry
    // -- nothing to record. (A run of one invalid row index starts in a group that takes the slow path: its
    // first entry differs from its predecessor, or it is entry 0 whose predecessor is -1.)
    if (e0 + 4 < nnz && prev >= 0 && ((rv[0] ^ prev) | (rv[1] ^ prev) | (rv[2] ^ prev) | (rv[3] ^ prev)) == 0) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long e = e0 + i;
      if (e >= nnz) break;
      const int r = rv[i];
      if (r < 0 || r >= n || r < prev) { bad = true; prev = r; continue; }
      if (r != prev)
        for (int q = max(prev, -1) + 1; q <= r; ++q) row_ptr[q] = e;
      if (e == nnz - 1)
        for (int q = r + 1; q <= n; ++q) row_ptr[q] = nnz;
      prev = r;
    }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// HIA_FILL_ATOMIC=1: the all-atomic fill for every row (A/B timing; read per call)
inline bool fill_atomic_only() {
  const char* e = getenv("HIA_FILL_ATOMIC");
  return e && atoi(e) == 1;
}

constexpr int kSegCols = 16384;              // 64 KB of shared memory per CTA, 3 CTAs per SM
constexpr int kSegThreads = 512;

// 4 consecutive column indices (entry index % 4 == 0): one 128-bit (int32) or 64-bit (uint16) load
__device__ __forceinline__ int4 load_cols4(const int* p) { return *reinterpret_cast<const int4*>(p); }
__device__ __forceinline__ int4 load_cols4(const unsigned short* p) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  return make_int4((int)(t.x & 0xffffu), (int)(t.x >> 16), (int)(t.y & 0xffffu), (int)(t.y >> 16));
}

// The default fill. One CTA per (row, 16 384-column segment): the segment is built in shared memory and
// written with 128-bit stores. Entries whose columns increase STRICTLY inside the row (what a sorted COO /
// canonical CSR holds) hit distinct cells, so they are placed with plain shared-memory stores; every entry
// is compared with its predecessor on the way, and a row that breaks the order (duplicates, unsorted
// columns) is redone by the same CTA with shared-memory atomics -- a CTA-local decision, both segment
// CTAs of a row see the same entries. Why: a shared-memory float atomicAdd is a compare-and-swap loop
// (~8 instructions + replays); ncu showed the all-atomic kernel (fill_rows_sorted_kernel below, still used
// for the fused statistics) at 36 % issue utilisation and 44 ... 57 % of DRAM peak, issue-bound at the
// clock the step runs at. The buffer is shifted by the 16-byte phase of the output segment so that the
// write-out is aligned 128-bit shared loads + 128-bit global stores.
template <typename CT>
__global__ void __launch_bounds__(kSegThreads)
fill_rows_fast_kernel(const CT* __restrict__ cols, const float* __restrict__ vals,
                      const long long* __restrict__ row_ptr, const int* __restrict__ flag, int n,
                      long long ld, int mtx_type, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_buf[];            // kSegCols + 4 floats
  if (flag && *flag) return;                               // rows not sorted: the any-order kernels do the work
  // (all loops below run on 32-bit indices with the compile-time stride kSegThreads: with blockDim.x and 64-bit
  //  entry indices the compiler's unrolling computed every trip count by an emulated division -- a third of the
  //  kernel's instructions for loops of ~8 iterations)
  constexpr int T = kSegThreads;
  const int tid = threadIdx.x;
  const int segs = (n + kSegCols - 1) / kSegCols;
  const int r = blockIdx.x / segs;
  const int c_lo = (mtx_type == HIA_MTX_TRIU) ? r : 0;     // a 'triu' row starts on the diagonal
  const int seg_lo = c_lo + (blockIdx.x - r * segs) * kSegCols;
  if (seg_lo >= n) return;
  const int seg_n = min(kSegCols, n - seg_lo);
  const long long e0 = row_ptr[r], e1 = row_ptr[r + 1];    // in flight while the buffer is zeroed
  const int off = (int)(((long long)r * ld + seg_lo) & 3);  // cell i of the segment lives in s_buf[i + off]
  float* s_row = s_buf + off;
  const int nq = (off + seg_n + 3) >> 2;                   // float4 groups of the buffer in use
  float4* s_q = reinterpret_cast<float4*>(s_buf);
  for (int i = tid; i < nq; i += T) s_q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  // the row's entries (a row holds < 2^31 of them): scalar head up to a 16-byte boundary, 128-bit body (two
  // groups in flight), scalar tail; indices relative to the row's first entry
  const CT* rc = cols + e0;
  const float* rv = vals + e0;
  const int len = (int)(e1 - e0);
  const int h = min(len, (int)((4 - (e0 & 3)) & 3));       // entries before the first 16-byte aligned one
  const int ng = (len - h) >> 2;                           // aligned groups of 4
  const int t0 = h + 4 * ng;                               // first tail entry
  bool bad = false;
  auto put = [&](int prev, int c, float v) {
    bad |= (c <= prev);
    c -= seg_lo;
    if (c >= 0 && c < seg_n) s_row[c] = v;
  };
  auto before = [&](int e) { return e > 0 ? (int)rc[e - 1] : -1; };
  if (tid < h) put(before(tid), (int)rc[tid], rv[tid]);
  int g = tid;
  for (; g + T < ng; g += 2 * T) {
    const int ea = h + 4 * g, eb = ea + 4 * T;
    const int4 c0 = load_cols4(rc + ea);
    const float4 v0 = *reinterpret_cast<const float4*>(rv + ea);
    const int4 c1 = load_cols4(rc + eb);
    const float4 v1 = *reinterpret_cast<const float4*>(rv + eb);
    const int p0 = before(ea), p1 = before(eb);
    put(p0, c0.x, v0.x); put(c0.x, c0.y, v0.y); put(c0.y, c0.z, v0.z); put(c0.z, c0.w, v0.w);
    put(p1, c1.x, v1.x); put(c1.x, c1.y, v1.y); put(c1.y, c1.z, v1.z); put(c1.z, c1.w, v1.w);
  }
  if (g < ng) {
    const int ea = h + 4 * g;
    const int4 c0 = load_cols4(rc + ea);
    const float4 v0 = *reinterpret_cast<const float4*>(rv + ea);
    const int p0 = before(ea);
    put(p0, c0.x, v0.x); put(c0.x, c0.y, v0.y); put(c0.y, c0.z, v0.z); put(c0.z, c0.w, v0.w);
  }
  if (t0 + tid < len) put(before(t0 + tid), (int)rc[t0 + tid], rv[t0 + tid]);
  if (__syncthreads_or(bad)) {
    // duplicates or unsorted columns in this row: redo the segment with atomics
    for (int i = tid; i < nq; i += T) s_q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int e = tid; e < len; e += T) {
      const int c = (int)rc[e] - seg_lo;
      if (c >= 0 && c < seg_n) atomicAdd(&s_row[c], rv[e]);
    }
    __syncthreads();
  }
  // write-out: buffer group q <-> cells 4 q - off ... 4 q - off + 3 of the segment (16-byte aligned in `out`)
  float* dst = out + (long long)r * ld + seg_lo - off;
  for (int q = tid; q < nq; q += T) {
    const int i0 = 4 * q - off;
    const float4 v = s_q[q];
    if (i0 >= 0 && i0 + 3 < seg_n) {
      *reinterpret_cast<float4*>(dst + 4 * q) = v;
    } else {
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + j >= 0 && i0 + j < seg_n) dst[4 * q + j] = e[j];
    }
  }
}

// CT = int (COO / CSR with 32-bit columns) or unsigned short (CSR with 16-bit columns, n <= 65 536:
// 6 instead of 8 bytes per entry over PCIe and out of HBM)
// Optional by-product of the fill (STATS): the inter-chromosome statistics of the O/E expected pass
// (sum, non-zero count and smallest positive value of every upper block a < b; what
// inter_block_kernel of hia_oe.cu computes from the dense map) taken from the finished row segment
// while it is still in shared memory -- saves the 1.8 GB re-read of the upper triangle (0.5 ms at
// 30 894 bins). Every thread keeps (chromosome, sum, count, min) for the run of columns it walks and
// flushes into per-CTA tables when the chromosome changes; one global atomic per (row, chromosome).
// MEASURED AND REJECTED as a default: correct (tests), but a thread's columns are 2 048 apart, so it
// flushes on almost every iteration, and 512 threads hammering a few fp64 shared-memory words (CAS
// loops) turned the 1.0 ms fill into 12 ms. Kept behind the *_stats entry points / HIA_FUSE_STATS=1;
// a per-chromosome block reduction over the finished segment would be the way to make it pay.
constexpr int kStatsMaxChr = 64;
struct InterStats {
  const int* chr_start;
  const short* bin_chr;
  int n_chr;
  double* sum;
  unsigned long long* nnz;
  float* minpos;
  int* valid;               // set to 1 by the kernel (the sorted-row path ran: the statistics are complete)
};

template <typename CT, bool STATS, int SEGC = kSegCols, int THR = kSegThreads>
__global__ void __launch_bounds__(THR)
fill_rows_sorted_kernel(const CT* __restrict__ cols, const float* __restrict__ vals,
                        const long long* __restrict__ row_ptr, const int* __restrict__ flag, int n,
                        long long ld, int mtx_type, float* __restrict__ out, const InterStats st) {
  extern __shared__ float s_row[];
  __shared__ double s_sum[STATS ? kStatsMaxChr : 1];
  __shared__ unsigned long long s_nnz[STATS ? kStatsMaxChr : 1];
  __shared__ float s_min[STATS ? kStatsMaxChr : 1];
  if (flag && *flag) return;
  if (STATS) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *st.valid = 1;
    for (int i = threadIdx.x; i < st.n_chr; i += blockDim.x) { s_sum[i] = 0.0; s_nnz[i] = 0ull; s_min[i] = INFINITY; }
  }
  // 1-D grid, the segments of a row in ADJACENT CTAs: every segment's CTA walks all entries of the row,
  // and the first layout (segments on blockIdx.y, i.e. ~n CTAs apart) re-read them from DRAM (ncu:
  // 3.07 GB read for 1.8 GB of entries, L2 hit rate 7 %); neighbours find them in L2.
  const int segs = (n + SEGC - 1) / SEGC;
  const int r = blockIdx.x / segs;
  const int c_lo = (mtx_type == HIA_MTX_TRIU) ? r : 0;     // a 'triu' row starts on the diagonal
  const int seg_lo = c_lo + (blockIdx.x % segs) * SEGC;
  if (seg_lo >= n) return;
  const int seg_n = min(SEGC, n - seg_lo);
  const long long e0 = row_ptr[r], e1 = row_ptr[r + 1];    // in flight while the buffer is zeroed
  for (int i = threadIdx.x; i < seg_n; i += blockDim.x) s_row[i] = 0.f;
  __syncthreads();
  // the row's entries: scalar head up to a 16-byte boundary, 128-bit body (two groups in flight), scalar tail
  const long long a0 = min(e1, (e0 + 3) & ~3ll), a1 = max(a0, e1 & ~3ll);
  auto add = [&](int c, float v) {
    c -= seg_lo;
    if (c >= 0 && c < seg_n) atomicAdd(&s_row[c], v);
  };
  for (long long e = e0 + threadIdx.x; e < a0; e += blockDim.x) add((int)cols[e], vals[e]);
  const long long ng = (a1 - a0) >> 2;
  long long g = threadIdx.x;
  for (; g + blockDim.x < ng; g += 2 * blockDim.x) {
    const int4 c0 = load_cols4(cols + a0 + 4 * g);
    const float4 v0 = *reinterpret_cast<const float4*>(vals + a0 + 4 * g);
    const int4 c1 = load_cols4(cols + a0 + 4 * (g + blockDim.x));
    const float4 v1 = *reinterpret_cast<const float4*>(vals + a0 + 4 * (g + blockDim.x));
    add(c0.x, v0.x); add(c0.y, v0.y); add(c0.z, v0.z); add(c0.w, v0.w);
    add(c1.x, v1.x); add(c1.y, v1.y); add(c1.z, v1.z); add(c1.w, v1.w);
  }
  for (; g < ng; g += blockDim.x) {
    const int4 c0 = load_cols4(cols + a0 + 4 * g);
    const float4 v0 = *reinterpret_cast<const float4*>(vals + a0 + 4 * g);
    add(c0.x, v0.x); add(c0.y, v0.y); add(c0.z, v0.z); add(c0.w, v0.w);
  }
  for (long long e = a1 + threadIdx.x; e < e1; e += blockDim.x) add((int)cols[e], vals[e]);
  __syncthreads();
  float* dst = out + (long long)r * ld + seg_lo;
  // per-thread running statistics of the inter-chromosome cells it writes (STATS)
  int cur_b = -1, col_lo = 0x7fffffff, a_chr = 0;
  float fsum = 0.f, fmin = INFINITY;
  unsigned int fcnt = 0;
  if (STATS) { a_chr = st.bin_chr[r]; col_lo = st.chr_start[a_chr + 1]; }
  auto flush = [&]() {
    if (cur_b >= 0 && (fcnt | (fsum != 0.f))) {
      atomicAdd(&s_sum[cur_b], (double)fsum);
      atomicAdd(&s_nnz[cur_b], (unsigned long long)fcnt);
      if (fmin != INFINITY) atomic_min_pos_float(&s_min[cur_b], fmin);
    }
    fsum = 0.f; fcnt = 0; fmin = INFINITY;
  };
  auto stat = [&](int c, float v) {                       // c: global column of a written cell
    if (c < col_lo) return;
    const int b = st.bin_chr[c];
    if (b != cur_b) { flush(); cur_b = b; }
    fsum += v;
    fcnt += (v != 0.f);
    if (v > 0.f) fmin = fminf(fmin, v);
  };
  // 128-bit stores from the first 16-byte aligned cell of the segment on
  const int head = min(seg_n, (int)((4 - ((((long long)r * ld + seg_lo)) & 3)) & 3));
  for (int i = threadIdx.x; i < head; i += blockDim.x) {
    dst[i] = s_row[i];
    if (STATS) stat(seg_lo + i, s_row[i]);
  }
  const int nv = (seg_n - head) >> 2;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const int o = head + 4 * i;
    const float4 q = make_float4(s_row[o], s_row[o + 1], s_row[o + 2], s_row[o + 3]);
    *reinterpret_cast<float4*>(dst + o) = q;
    if (STATS && seg_lo + o + 3 >= col_lo) {
      stat(seg_lo + o, q.x); stat(seg_lo + o + 1, q.y); stat(seg_lo + o + 2, q.z); stat(seg_lo + o + 3, q.w);
    }
  }
  for (int i = head + 4 * nv + threadIdx.x; i < seg_n; i += blockDim.x) {
    dst[i] = s_row[i];
    if (STATS) stat(seg_lo + i, s_row[i]);
  }
  if (STATS) {
    flush();
    __syncthreads();
    for (int i = threadIdx.x; i < st.n_chr; i += blockDim.x) {
      if (s_nnz[i] | (s_sum[i] != 0.0)) {
        atomicAdd(&st.sum[a_chr * st.n_chr + i], s_sum[i]);
        atomicAdd(&st.nnz[a_chr * st.n_chr + i], s_nnz[i]);
      }
      if (s_min[i] != INFINITY) atomic_min_pos_float(&st.minpos[a_chr * st.n_chr + i], s_min[i]);
    }
  }
}

// fallback zero fill (any entry order): does nothing when the sorted-row path ran
__global__ void __launch_bounds__(256)
zero_fill_if_kernel(float* __restrict__ out, int n, long long ld, const int* __restrict__ only_if_flag) {
  if (!*only_if_flag) return;
  const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c0 >= n) return;
  for (int r = blockIdx.y; r < n; r += gridDim.y) {
    float* p = out + (long long)r * ld + c0;
    if (c0 + 3 < n) *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
    else for (int j = 0; j < 4 && c0 + j < n; ++j) p[j] = 0.f;
  }
}

// out[c][r] = out[r][c] for r < c: 32 x 32 tiles through shared memory, both sides coalesced.
__global__ void __launch_bounds__(256)
mirror_kernel(float* __restrict__ out, int n, long long ld) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;            // tile (bi, bj) of the upper triangle
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = bi * 32 + ty + 8 * k, c = bj * 32 + tx;
    tile[ty + 8 * k][tx] = (r < n && c < n) ? out[(long long)r * ld + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = bj * 32 + ty + 8 * k, c = bi * 32 + tx;   // transposed position
    if (r < n && c < n && r > c) out[(long long)r * ld + c] = tile[tx][ty + 8 * k];
  }
}

// ------------------------------------------------------------------ f1: condense / trim as a remap
// out[map[r]][map[c]] += v (and the mirrored cell for a 'triu' input, at the FINE level: a fine
// off-diagonal entry whose two bins fall into one coarse bin is counted twice on the coarse
// diagonal, exactly like condensing the symmetrised matrix, new_hic_class.py:2426-2465).
__global__ void __launch_bounds__(256)
scatter_remap_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
                     const float* __restrict__ vals, long long nnz, int n_fine,
                     const int* __restrict__ index_map, int mtx_type, float* __restrict__ out,
                     long long ld) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    const int r = rows[e], c = cols[e];
    if (r < 0 || c < 0 || r >= n_fine || c >= n_fine) continue;
    if (mtx_type == HIA_MTX_TRIU && c < r) continue;
    const int rr = index_map[r], cc = index_map[c];
    if (rr < 0 || cc < 0) continue;
    const float v = vals[e];
    if (mtx_type == HIA_MTX_TRIU) {
      // canonical (upper) coarse cell only; mirror_kernel copies it below the diagonal afterwards,
      // so the coarse map is symmetric BIT FOR BIT (the O/E passes read the upper triangle only).
      // An off-diagonal fine entry folded onto the coarse diagonal counts twice (2v is exact).
      const int lo = min(rr, cc), hi = max(rr, cc);
      atomicAdd(out + (long long)lo * ld + hi, (lo == hi && r != c) ? 2.f * v : v);
    } else {
      atomicAdd(out + (long long)rr * ld + cc, v);
    }
  }
}

// non-zero stored entries per chromosome pair of the symmetrised ('all') matrix
__global__ void __launch_bounds__(256)
coo_block_nnz_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
                     const float* __restrict__ vals, long long nnz, int n, int mtx_type,
                     const short* __restrict__ bin_chr, int n_chr,
                     unsigned long long* __restrict__ counts) {
  int cur = -1;
  unsigned long long acc = 0;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    const int r = rows[e], c = cols[e];
    if (r < 0 || c < 0 || r >= n || c >= n) continue;
    if (mtx_type == HIA_MTX_TRIU && c < r) continue;
    if (vals[e] == 0.f) continue;
    const int a = bin_chr[r], b = bin_chr[c];
    const bool twice = (mtx_type == HIA_MTX_TRIU && r != c);
    if (a == b) {
      const int slot = a * n_chr + a;
      if (slot != cur) { if (cur >= 0) atomicAdd(&counts[cur], acc); cur = slot; acc = 0; }
      acc += twice ? 2 : 1;
    } else {
      atomicAdd(&counts[a * n_chr + b], 1ull);
      if (twice) atomicAdd(&counts[b * n_chr + a], 1ull);
    }
  }
  if (cur >= 0) atomicAdd(&counts[cur], acc);
}

// fp64 row sums of the symmetrised, remapped matrix straight from the fine COO (no dense map):
// decides which bins tid_off_low_bins keeps, so it must not depend on fp32 rounding of the map.
__global__ void __launch_bounds__(256)
coo_bin_sums_kernel(const int* __restrict__ rows, const int* __restrict__ cols,
                    const float* __restrict__ vals, long long nnz, int n_fine,
                    const int* __restrict__ index_map, int mtx_type, double* __restrict__ bin_sum) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    const int r = rows[e], c = cols[e];
    if (r < 0 || c < 0 || r >= n_fine || c >= n_fine) continue;
    if (mtx_type == HIA_MTX_TRIU && c < r) continue;
    const int rr = index_map[r], cc = index_map[c];
    if (rr < 0 || cc < 0) continue;
    const double v = (double)vals[e];
    atomicAdd(bin_sum + rr, v);
    if (mtx_type == HIA_MTX_TRIU && r != c) atomicAdd(bin_sum + cc, v);
  }
}

// per-row sum (fp64) and non-zero count of a dense map; CTA per row
__global__ void __launch_bounds__(256)
bin_stats_kernel(const float* __restrict__ x, int n, long long ld, double* __restrict__ row_sum,
                 long long* __restrict__ row_nnz) {
  __shared__ double s_sum[8];
  __shared__ unsigned long long s_cnt[8];
  const int r = blockIdx.x;
  const float* row = x + (long long)r * ld;
  double acc = 0.0;
  unsigned long long cnt = 0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) { const float v = row[c]; acc += (double)v; cnt += (v != 0.f); }
  acc = warp_sum(acc);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = acc; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0; unsigned long long k = 0;
    for (int i = 0; i < 8; ++i) { a += s_sum[i]; k += s_cnt[i]; }
    row_sum[r] = a;
    row_nnz[r] = (long long)k;
  }
}

}  // namespace
}  // namespace hia

HIA_API int hia_scatter_coo_remap(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                                  int32_t n_fine, const int32_t* index_map, int32_t n_out, float* out,
                                  int64_t ld, int mtx_type, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!out || !index_map || n_fine <= 0 || n_out <= 0 || ld < n_out || nnz < 0) return HIA_ERR_BAD_ARG;
  if (nnz > 0 && (!rows || !cols || !vals)) return HIA_ERR_BAD_ARG;
  if (mtx_type != HIA_MTX_TRIU && mtx_type != HIA_MTX_ALL) return HIA_ERR_BAD_ARG;
  HIA_CUDA(cudaMemset2DAsync(out, (size_t)ld * 4, 0, (size_t)n_out * 4, n_out, stream));
  if (nnz > 0) {
    const int grid = (int)max((int64_t)1, min((int64_t)num_sms() * 16, (nnz + 255) / 256));
    scatter_remap_kernel<<<grid, 256, 0, stream>>>(rows, cols, vals, nnz, n_fine, index_map, mtx_type, out, ld);
    HIA_LAUNCH_CHECK();
    if (mtx_type == HIA_MTX_TRIU) {
      dim3 gm((n_out + 31) / 32, (n_out + 31) / 32);
      mirror_kernel<<<gm, 256, 0, stream>>>(out, n_out, ld);
      HIA_LAUNCH_CHECK();
    }
  }
  return HIA_OK;
}

HIA_API int hia_coo_block_nnz(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                              int32_t n, int mtx_type, const int16_t* bin_chr, int32_t n_chr,
                              unsigned long long* counts, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!bin_chr || !counts || n <= 0 || n_chr <= 0 || nnz < 0) return HIA_ERR_BAD_ARG;
  if (nnz > 0 && (!rows || !cols || !vals)) return HIA_ERR_BAD_ARG;
  HIA_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_chr * n_chr * 8, stream));
  if (nnz > 0) {
    const int grid = (int)max((int64_t)1, min((int64_t)num_sms() * 16, (nnz + 255) / 256));
    coo_block_nnz_kernel<<<grid, 256, 0, stream>>>(rows, cols, vals, nnz, n, mtx_type, bin_chr, n_chr, counts);
    HIA_LAUNCH_CHECK();
  }
  return HIA_OK;
}

HIA_API int hia_coo_bin_sums(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                             int32_t n_fine, const int32_t* index_map, int32_t n_out, int mtx_type,
                             double* bin_sum, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!bin_sum || !index_map || n_fine <= 0 || n_out <= 0 || nnz < 0) return HIA_ERR_BAD_ARG;
  if (nnz > 0 && (!rows || !cols || !vals)) return HIA_ERR_BAD_ARG;
  if (mtx_type != HIA_MTX_TRIU && mtx_type != HIA_MTX_ALL) return HIA_ERR_BAD_ARG;
  HIA_CUDA(cudaMemsetAsync(bin_sum, 0, (size_t)n_out * 8, stream));
  if (nnz > 0) {
    const int grid = (int)max((int64_t)1, min((int64_t)num_sms() * 16, (nnz + 255) / 256));
    coo_bin_sums_kernel<<<grid, 256, 0, stream>>>(rows, cols, vals, nnz, n_fine, index_map, mtx_type, bin_sum);
    HIA_LAUNCH_CHECK();
  }
  return HIA_OK;
}

HIA_API int hia_bin_stats(const float* x, int32_t n, int64_t ld, double* row_sum, long long* row_nnz,
                          void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !row_sum || !row_nnz || n <= 0 || ld < n) return HIA_ERR_BAD_ARG;
  bin_stats_kernel<<<n, 256, 0, stream>>>(x, n, ld, row_sum, row_nnz);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

// workspace: [keys: n_chr^2 + 1 u32 (min fill)] [flag: 1 int, 256-byte slot] [row_ptr: n + 1 i64]
static int64_t scatter_keys_bytes(int32_t n_chr) {
  if (n_chr < 1) n_chr = 1;
  return ((int64_t)((int64_t)n_chr * n_chr + 1) * 4 + 255) / 256 * 256;
}
HIA_API int64_t hia_scatter_workspace_bytes(int32_t n, int32_t n_chr) {
  if (n < 0) n = 0;
  return scatter_keys_bytes(n_chr) + 256 + ((int64_t)n + 1) * 8;
}

namespace hia {
namespace {
__global__ void init_inter_stats_kernel(double* sum, unsigned long long* nnz, float* minpos, int cc, int* valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cc) { sum[i] = 0.0; nnz[i] = 0ull; minpos[i] = INFINITY; }
  if (i == 0) *valid = 0;
}
int init_inter_stats(const InterStats& st, cudaStream_t stream) {
  const int cc = st.n_chr * st.n_chr;
  init_inter_stats_kernel<<<(cc + 255) / 256, 256, 0, stream>>>(st.sum, st.nnz, st.minpos, cc, st.valid);
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}
}  // namespace
}  // namespace hia

static int scatter_coo_impl(const int32_t* rows, const int32_t* cols, const float* vals,
                            int64_t nnz, int32_t n, float* out, int64_t ld, int mtx_type,
                            int fill_mode, const int32_t* chr_start, int32_t n_chr,
                            void* workspace, void* stream_, const hia::InterStats* stats) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (stats) {
    const int st = init_inter_stats(*stats, stream);
    if (st != HIA_OK) return st;
  }
  if (!out || n <= 0 || ld < n || nnz < 0) return HIA_ERR_BAD_ARG;
  if (nnz > 0 && (!rows || !cols || !vals)) return HIA_ERR_BAD_ARG;
  if (mtx_type != HIA_MTX_TRIU && mtx_type != HIA_MTX_ALL && mtx_type != HIA_MTX_TRIU_UPPER) return HIA_ERR_BAD_ARG;
  if (fill_mode < HIA_FILL_ZERO || fill_mode > HIA_FILL_BLOCK_MIN) return HIA_ERR_BAD_ARG;
  const bool mirror = (mtx_type == HIA_MTX_TRIU);
  if (mtx_type == HIA_MTX_TRIU_UPPER) {                  // upper triangle only: no mirror pass
    if (fill_mode != HIA_FILL_ZERO) return HIA_ERR_BAD_ARG;
    mtx_type = HIA_MTX_TRIU;
  }
  if (!aligned16(out) || (ld & 3)) return HIA_ERR_ALIGN;
  if (fill_mode != HIA_FILL_ZERO && (!workspace || !chr_start || n_chr < 1)) return HIA_ERR_BAD_ARG;
  if (n_chr > kMaxChr) return HIA_ERR_UNSUPPORTED;

  const int threads = 256;
  const int grid_e = (int)max((int64_t)1, min((int64_t)num_sms() * 16, (nnz + threads - 1) / threads));
  if (fill_mode == HIA_FILL_ZERO && workspace && nnz > 0 && aligned16(workspace) && aligned16(rows) &&
      aligned16(cols) && aligned16(vals)) {
    // sorted-row path, with the any-order kernels as a device-side fallback (flag)
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    int* flag = reinterpret_cast<int*>(ws + scatter_keys_bytes(n_chr));
    long long* row_ptr = reinterpret_cast<long long*>(ws + scatter_keys_bytes(n_chr) + 256);
    HIA_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), stream));
    row_ptr_kernel<<<grid_e, threads, 0, stream>>>(rows, nnz, n, row_ptr, flag);
    HIA_LAUNCH_CHECK();
    static bool attr_set = false;                          // once per process (idempotent)
    if (!attr_set) {
      HIA_CUDA((cudaFuncSetAttribute(fill_rows_sorted_kernel<int, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSegCols * (int)sizeof(float))));
      HIA_CUDA((cudaFuncSetAttribute(fill_rows_sorted_kernel<int, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kSegCols * (int)sizeof(float))));
      HIA_CUDA((cudaFuncSetAttribute(fill_rows_fast_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (kSegCols + 4) * (int)sizeof(float))));
      attr_set = true;
    }
    dim3 grid_f((unsigned)n * ((n + kSegCols - 1) / kSegCols), 1);
    // (smaller segments -- more CTAs per SM in different phases -- measured slower in-step at 30 894 bins:
    //  8192 columns x 256 threads 2.10 ms, 8192 x 512 2.37 ms, 4096 x 256 2.39 ms against 1.89 ms for this)
    if (stats)
      fill_rows_sorted_kernel<int, true><<<grid_f, kSegThreads, kSegCols * sizeof(float), stream>>>(
          cols, vals, row_ptr, flag, n, ld, mtx_type, out, *stats);
    else if (fill_atomic_only())
      fill_rows_sorted_kernel<int, false><<<grid_f, kSegThreads, kSegCols * sizeof(float), stream>>>(
          cols, vals, row_ptr, flag, n, ld, mtx_type, out, InterStats{});
    else
      fill_rows_fast_kernel<int><<<grid_f, kSegThreads, (kSegCols + 4) * sizeof(float), stream>>>(
          cols, vals, row_ptr, flag, n, ld, mtx_type, out);
    HIA_LAUNCH_CHECK();
    // (a 4096-row grid of CTAs that only read the flag and exit cost 69 us per call: launch list r1)
    dim3 grid_z((n + threads * 4 - 1) / (threads * 4), min(n, 64));
    zero_fill_if_kernel<<<grid_z, threads, 0, stream>>>(out, n, ld, flag);
    HIA_LAUNCH_CHECK();
    scatter_kernel<true><<<grid_e, threads, 0, stream>>>(rows, cols, vals, nnz, n, mtx_type, out, ld, flag);
    HIA_LAUNCH_CHECK();
  } else if (fill_mode == HIA_FILL_ZERO) {
    if (ld == n) {
      HIA_CUDA(cudaMemsetAsync(out, 0, (size_t)n * n * sizeof(float), stream));
    } else {
      HIA_CUDA(cudaMemset2DAsync(out, (size_t)ld * 4, 0, (size_t)n * 4, n, stream));
    }
    if (nnz > 0) {
      scatter_kernel<true><<<grid_e, threads, 0, stream>>>(rows, cols, vals, nnz, n, mtx_type, out, ld, nullptr);
      HIA_LAUNCH_CHECK();
    }
  } else {
    unsigned int* keys = static_cast<unsigned int*>(workspace);
    const int nk = n_chr * n_chr + 1;
    init_keys_kernel<<<(nk + 255) / 256, 256, 0, stream>>>(keys, nk);
    HIA_LAUNCH_CHECK();
    const size_t sh = (size_t)(n_chr + 1) * sizeof(int);
    dim3 grid((n + threads * 4 - 1) / (threads * 4), min(n, 4096));
    fill_const_kernel<<<grid, threads, 0, stream>>>(out, n, ld, nanf(""));
    HIA_LAUNCH_CHECK();
    if (nnz > 0) {
      scatter_kernel<false><<<grid_e, threads, 0, stream>>>(rows, cols, vals, nnz, n, mtx_type, out, ld, nullptr);
      HIA_LAUNCH_CHECK();
      scatter_kernel<true><<<grid_e, threads, 0, stream>>>(rows, cols, vals, nnz, n, mtx_type, out, ld, nullptr);
      HIA_LAUNCH_CHECK();
      block_min_kernel<<<grid_e, threads, sh, stream>>>(rows, cols, out, ld, nnz, n, mtx_type, chr_start,
                                                          n_chr, fill_mode == HIA_FILL_BLOCK_MIN, keys);
      HIA_LAUNCH_CHECK();
    }
    fixup_kernel<<<grid, threads, sh, stream>>>(out, n, ld, fill_mode, mtx_type, chr_start, n_chr, keys);
    HIA_LAUNCH_CHECK();
  }
  if (mirror) {                                            // to_all: triu(M,0).T + triu(M,1)
    dim3 gm((n + 31) / 32, (n + 31) / 32);
    mirror_kernel<<<gm, 256, 0, stream>>>(out, n, ld);
    HIA_LAUNCH_CHECK();
  }
  return HIA_OK;
}

HIA_API int hia_scatter_coo_sym(const int32_t* rows, const int32_t* cols, const float* vals,
                                int64_t nnz, int32_t n, float* out, int64_t ld, int mtx_type,
                                int fill_mode, const int32_t* chr_start, int32_t n_chr,
                                void* workspace, void* stream) {
  return scatter_coo_impl(rows, cols, vals, nnz, n, out, ld, mtx_type, fill_mode, chr_start, n_chr, workspace,
                          stream, nullptr);
}

// The same (zero fill) with the inter-chromosome statistics of the O/E expected pass as a by-product
// of the row-sorted fill: inter_sum / inter_nnz / inter_minpos [n_chr^2] as hia_oe_expected_pass
// defines them, *stats_valid (device int) = 1 when they are complete (row-sorted input; the any-order
// fallback leaves 0 and hia_oe_expected_pass_ex then computes them from the dense map). n_chr <= 64.
HIA_API int hia_scatter_coo_sym_stats(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                                      int32_t n, float* out, int64_t ld, int mtx_type, const int32_t* chr_start,
                                      int32_t n_chr, const int16_t* bin_chr, double* inter_sum,
                                      unsigned long long* inter_nnz, float* inter_minpos, int32_t* stats_valid,
                                      void* workspace, void* stream) {
  using namespace hia;
  if (!chr_start || !bin_chr || !inter_sum || !inter_nnz || !inter_minpos || !stats_valid || n_chr < 1 ||
      n_chr > kStatsMaxChr || !workspace)
    return HIA_ERR_BAD_ARG;
  InterStats st = {chr_start, bin_chr, n_chr, inter_sum, inter_nnz, inter_minpos, stats_valid};
  return scatter_coo_impl(rows, cols, vals, nnz, n, out, ld, mtx_type, HIA_FILL_ZERO, nullptr, 0, workspace, stream, &st);
}

// CSR input (what SpsHiCMtx.matrix is: scipy CSR, publish/new_hic_class.py:2158-2161): row_ptr[n + 1]
// (int64, device), column indices as int32 (col_bytes = 4) or uint16 (col_bytes = 2, n <= 65 536), values
// fp32. Zero fill only. Rows are sorted by construction, so there is no row-pointer pass and no
// fallback: 8 (or 6) bytes per entry instead of the 12 of the COO triple -- the bytes that cross PCIe
// in the end-to-end path. Entries with a column outside [0, n) (or below the diagonal for the
// 'triu' types) are ignored; duplicates are summed.
static int scatter_csr_impl(const int64_t* row_ptr, const void* cols, int col_bytes, const float* vals,
                            int32_t n, float* out, int64_t ld, int mtx_type, void* stream_,
                            const hia::InterStats* stats) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (stats) {
    const int st = init_inter_stats(*stats, stream);
    if (st != HIA_OK) return st;
  }
  if (!row_ptr || !cols || !vals || !out || n <= 0 || ld < n) return HIA_ERR_BAD_ARG;
  if (col_bytes != 4 && col_bytes != 2) return HIA_ERR_BAD_ARG;
  if (col_bytes == 2 && n > 65536) return HIA_ERR_UNSUPPORTED;
  if (mtx_type != HIA_MTX_TRIU && mtx_type != HIA_MTX_ALL && mtx_type != HIA_MTX_TRIU_UPPER) return HIA_ERR_BAD_ARG;
  if (!aligned16(out) || (ld & 3) || !aligned16(cols) || !aligned16(vals)) return HIA_ERR_ALIGN;
  const bool mirror = (mtx_type == HIA_MTX_TRIU);
  const int mt = (mtx_type == HIA_MTX_ALL) ? HIA_MTX_ALL : HIA_MTX_TRIU;
  static bool attr_set = false;
  if (!attr_set) {
    const int bytes = kSegCols * (int)sizeof(float);
    HIA_CUDA((cudaFuncSetAttribute(fill_rows_sorted_kernel<int, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)));
    HIA_CUDA((cudaFuncSetAttribute(fill_rows_sorted_kernel<int, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)));
    HIA_CUDA((cudaFuncSetAttribute(fill_rows_sorted_kernel<unsigned short, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)));
    HIA_CUDA((cudaFuncSetAttribute(fill_rows_sorted_kernel<unsigned short, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)));
    HIA_CUDA((cudaFuncSetAttribute(fill_rows_fast_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 16)));
    HIA_CUDA((cudaFuncSetAttribute(fill_rows_fast_kernel<unsigned short>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   bytes + 16)));
    attr_set = true;
  }
  dim3 grid_f((unsigned)n * ((n + kSegCols - 1) / kSegCols), 1);
  const long long* rp = reinterpret_cast<const long long*>(row_ptr);
  const size_t sh = kSegCols * sizeof(float);
  const InterStats none = {};
  if (col_bytes == 4 && stats)
    fill_rows_sorted_kernel<int, true><<<grid_f, kSegThreads, sh, stream>>>(static_cast<const int*>(cols), vals, rp,
                                                                            nullptr, n, ld, mt, out, *stats);
  else if (col_bytes == 4 && !fill_atomic_only())
    fill_rows_fast_kernel<int><<<grid_f, kSegThreads, sh + 16, stream>>>(static_cast<const int*>(cols), vals, rp, nullptr,
                                                                         n, ld, mt, out);
  else if (col_bytes == 4)
    fill_rows_sorted_kernel<int, false><<<grid_f, kSegThreads, sh, stream>>>(static_cast<const int*>(cols), vals, rp,
                                                                             nullptr, n, ld, mt, out, none);
  else if (stats)
    fill_rows_sorted_kernel<unsigned short, true><<<grid_f, kSegThreads, sh, stream>>>(
        static_cast<const unsigned short*>(cols), vals, rp, nullptr, n, ld, mt, out, *stats);
  else if (!fill_atomic_only())
    fill_rows_fast_kernel<unsigned short><<<grid_f, kSegThreads, sh + 16, stream>>>(
        static_cast<const unsigned short*>(cols), vals, rp, nullptr, n, ld, mt, out);
  else
    fill_rows_sorted_kernel<unsigned short, false><<<grid_f, kSegThreads, sh, stream>>>(
        static_cast<const unsigned short*>(cols), vals, rp, nullptr, n, ld, mt, out, none);
  HIA_LAUNCH_CHECK();
  if (mirror) {
    dim3 gm((n + 31) / 32, (n + 31) / 32);
    mirror_kernel<<<gm, 256, 0, stream>>>(out, n, ld);
    HIA_LAUNCH_CHECK();
  }
  return HIA_OK;
}

HIA_API int hia_scatter_csr_sym(const int64_t* row_ptr, const void* cols, int col_bytes, const float* vals,
                                int32_t n, float* out, int64_t ld, int mtx_type, void* stream) {
  return scatter_csr_impl(row_ptr, cols, col_bytes, vals, n, out, ld, mtx_type, stream, nullptr);
}

// CSR input + the inter-chromosome statistics of the O/E expected pass (see hia_scatter_coo_sym_stats);
// CSR rows are sorted by construction, so *stats_valid is always 1 afterwards.
HIA_API int hia_scatter_csr_sym_stats(const int64_t* row_ptr, const void* cols, int col_bytes, const float* vals,
                                      int32_t n, float* out, int64_t ld, int mtx_type, const int32_t* chr_start,
                                      int32_t n_chr, const int16_t* bin_chr, double* inter_sum,
                                      unsigned long long* inter_nnz, float* inter_minpos, int32_t* stats_valid,
                                      void* stream) {
  using namespace hia;
  if (!chr_start || !bin_chr || !inter_sum || !inter_nnz || !inter_minpos || !stats_valid || n_chr < 1 ||
      n_chr > kStatsMaxChr)
    return HIA_ERR_BAD_ARG;
  InterStats st = {chr_start, bin_chr, n_chr, inter_sum, inter_nnz, inter_minpos, stats_valid};
  return scatter_csr_impl(row_ptr, cols, col_bytes, vals, n, out, ld, mtx_type, stream, &st);
}
