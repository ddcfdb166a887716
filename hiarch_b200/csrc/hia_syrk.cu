// a8: row x row similarity of an m x k map block -- cosine distance (get_adj_mtx ==
// squareform(pdist(M,'cosine')), publish/complex/src/S1_get_adj_mtx.py:8-14) and Pearson
// correlation (the same contraction on row-centred rows; north-star flavour).
//
// out[i][j] = f( sum_k xn[i][k] * xn[j][k] ),  xn = unit-norm (optionally centred) rows.
//
// Design (sm_100a):
//   row_prep_kernel   fp64 row mean / norm, writes xn split into two TF32-exact fp32 planes
//                     hi = rna_tf32(xn), lo = rna_tf32(xn - hi)           (12 B/cell, HBM-bound)
//   syrk_tf32x3_kernel  persistent, warp-specialised tcgen05 kernel:
//       warp 0   TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, 128 B = 32 fp32 per row)
//       warp 1   MMA issuer     (tcgen05.mma.cta_group::1.kind::tf32, M=128, N=256, K=8 per
//                                instruction; 3 MMAs per k-step: hi.hi + hi.lo + lo.hi)
//       warps 2-9 epilogue      (tcgen05.ld 32x32b.x32 -> fp32 running sums in registers ->
//                                1-c / c -> coalesced direct + mirrored stores)
//     Accumulators live in tensor memory (2 x 256 columns, double buffered). To bound the
//     rounding of long fp32 accumulation chains the K loop is cut into chunks (k_chunk
//     elements): each chunk accumulates in TMEM, the 8 epilogue warps fold it into fp32
//     running sums held in registers (128 per thread) while the next chunk's MMAs run.
//   Only tiles touching the upper triangle (and the requested |i-j| band) are computed; the
//   lower triangle is written as the exact mirror, the diagonal as the exact constant.
//
// HIA_PREC_3XF16: the same split on kind::f16 -- fp16 has the SAME 11-bit significand as TF32, so
// hi = fp16(S xn), lo = fp16(S xn - hi) (S = 2^15, exact; |xn| <= 1 keeps hi finite) carry the same
// 22 bits, the products are exact in fp32, and one MMA instruction covers K = 16 instead of 8 in
// the same 128 cycles and the same 12 KB of shared-memory operands: half the MMA time, half the
// plane bytes. The products carry 2^30 -- the unit of the epilogue's fixed-point fold -- and the epilogue
// rescales by 2^-30 (exact).
//
// 3xTF32 error budget: |hi.hi + hi.lo + lo.hi - x.y| <= ~2^-21 |x||y| per product; fp32
// accumulation over <= k_chunk/8 MMA steps in TMEM (truncating: the bias grows with the chunk), then
// an EXACT fixed-point fold of the <= k/k_chunk partials in the epilogue registers (see there).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "hia_common.cuh"

namespace hia {
namespace {

constexpr int BM = 128;                        // rows of A per CTA
constexpr int BN = 256;                        // columns of the tile (rows of B)
constexpr int ROW_BYTES = 128;                 // one SW128 atom of K per row and k-block
constexpr int UMMA_K_BYTES = 32;               // K extent of one MMA instruction, all kinds
constexpr int A_BYTES = BM * ROW_BYTES;        // 16 KB
// elements per k-block: 32 (tf32 planes, fp32 storage) or 64 (fp16 planes)
__host__ __device__ constexpr int bk_of(int prec) { return prec == HIA_PREC_3XF16 ? 64 : 32; }
__host__ __device__ constexpr int elem_bytes(int prec) { return prec == HIA_PREC_3XF16 ? 2 : 4; }
constexpr float kF16Scale = 32768.f;                         // 2^15: products of two planes carry 2^30
// fixed-point fold of the chunk partials: |sum| <= 1 (unit rows) is held as an int32 in units of 2^-30
template <int PREC> struct FixScale { static constexpr float value = 1073741824.f; };          // tf32 planes: x 2^30
template <> struct FixScale<HIA_PREC_3XF16> { static constexpr float value = 1.f; };          // fp16 planes carry 2^30
constexpr float kFixUnscale = 1.f / 1073741824.f;            // 2^-30
constexpr int kFixNaN = (int)0x80000000;                     // fixed-point marker of a NaN sum
constexpr int EPI_WARPS = 8;                   // 2 per TMEM lane quarter x 128 columns each
constexpr int NUM_THREADS = 32 * (2 + EPI_WARPS);
constexpr int TRANS_PAD = 33;
constexpr int SMEM_TRANS = EPI_WARPS * 32 * TRANS_PAD * 4;            // 16896
constexpr int TMEM_COLS = 512;

// CTAS = 1: one CTA per 128 x 256 tile (cta_group::1).
// CTAS = 2: a CTA pair (cluster of 2 on one TPC, cta_group::2) per 256 x 256 tile: each CTA stages
//           its 128 rows of A and HALF of B (128 of the 256 B rows); the pair's tensor cores read
//           both halves. Per CTA and k-block that is 64 KB of TMA writes + 8 KB of operand reads
//           per MMA instead of 96 KB + 12 KB: the 1-CTA kernel was shared-memory-bandwidth bound
//           (156 B/clk demanded of 128 B/clk -> 78 % tensor-pipe activity measured).
template <int CTAS> struct Geo {
  static constexpr int B_ROWS = BN / CTAS;
  static constexpr int B_BYTES = B_ROWS * ROW_BYTES;              // 32 KB | 16 KB
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);     // hi + lo planes: 96 KB | 64 KB
  static constexpr int STAGES = (CTAS == 2) ? 3 : 2;
  static constexpr int SMEM_TILES = STAGES * STAGE_BYTES;         // 192 KB
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + SMEM_TILES + SMEM_TRANS + 256 /*barriers*/;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      :: "r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug traps (kernel error) after ~4 s instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fffu) == 0) {
      const unsigned long long now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
template <int CTAS>
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c_inner, int c_outer) {
  if constexpr (CTAS == 2) {
    // executed by both CTAs of the pair; the peer bit of the barrier address is cleared so that
    // the transaction bytes land on the LEADER's barrier (cute SM100_TMA_2SM_LOAD)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c_inner), "r"(c_outer)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer) : "memory");
  }
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int PREC, int CTAS>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                     uint32_t idesc, uint32_t accumulate) {
#define HIA_UMMA(GROUP, KIND)                                                              \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                           \
               "tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%0], %1, %2, %3, p;\n\t}" \
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory")
  if constexpr (PREC == HIA_PREC_3XF16 && CTAS == 2) HIA_UMMA("2", "f16");
  else if constexpr (PREC == HIA_PREC_3XF16) HIA_UMMA("1", "f16");
  else if constexpr (CTAS == 2) HIA_UMMA("2", "tf32");
  else HIA_UMMA("1", "tf32");
#undef HIA_UMMA
}
// MMA completion -> mbarrier (of this CTA, or of both CTAs of the pair at the same offset)
template <int CTAS>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if constexpr (CTAS == 2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, set 1) |
//   [32,46) SBO >> 4 = 1024 B (8 rows x 128 B) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a/b format [7,10),[10,13)
// (F16 = 0, TF32 = 2), K-major both, N>>3 at [17,23), M>>4 at [24,29) (M = 256 for the CTA pair).
__host__ __device__ constexpr uint32_t idesc_of(int prec, int ctas) {
  const uint32_t fmt = prec == HIA_PREC_3XF16 ? 0u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) |
         ((uint32_t)((BM * ctas) >> 4) << 24);
}

struct TileCoord { int m0; int n0; };            // first A row / first B row of a TILE_M x BN tile

// Everything the kernel needs besides the four tensor maps. A (rows of the output) and B (columns
// of the output) may be DIFFERENT plane sets (row-sharded ring: A = own rows, B = a peer block
// staged in local HBM); the tile coordinates are plane-row indices of A and of B.
//   out  [(r - row0) * ld_out + (c - col0)]   direct store, r in [row0, row_end), c in [col0, col_end)
//   out_t[(c - col0) * ld_t   + (r - row0)]   transposed store of the same cell (optional). In the
//         symmetric single-map case out_t == out; in the ring it is the PEER GPU's row block
//         (a P2P pointer: the mirrored half of the symmetric result travels over NVLink straight
//         out of the epilogue, there is no separate exchange step).
//   symmetric: A == B, only cells c >= r are produced directly and c > r transposed.
//   has_diag : A and B index the same rows, cells with c == r get the exact constant.
struct SyrkArgs {
  const TileCoord* tiles;
  int num_tiles, num_k_blocks, kb_per_chunk, flavour;
  float* out;
  long long ld_out;
  float* out_t;
  long long ld_t;
  unsigned int* sync_ctr;
  int sync_every, row0, row_end, col0, col_end, symmetric, has_diag, same_planes;
};

// ------------------------------------------------------------------ the kernel
template <int PREC, int CTAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
syrk_split3_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                   const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                   const SyrkArgs a) {
  const TileCoord* __restrict__ tiles = a.tiles;
  const int num_tiles = a.num_tiles, num_k_blocks = a.num_k_blocks, kb_per_chunk = a.kb_per_chunk;
  const int flavour = a.flavour, sync_every = a.sync_every, symmetric = a.symmetric;
  unsigned int* __restrict__ sync_ctr = a.sync_ctr;
  using G = Geo<CTAS>;
  constexpr int STAGES = G::STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* s_trans = reinterpret_cast<float*>(smem + G::SMEM_TILES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::SMEM_TILES + SMEM_TRANS);
  uint64_t* full_bar = bars;                 // [STAGES]   (pair: the leader's is the one used)
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]   (pair: one per CTA, multicast commit)
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]        (pair: one per CTA, multicast commit)
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]    (pair: the leader's, both CTAs' epilogues arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if constexpr (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const int unit = blockIdx.x / CTAS;               // tile-processing unit: a CTA or a CTA pair
  const int num_units = gridDim.x / CTAS;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_b_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_b_lo) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], EPI_WARPS * CTAS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   :: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();   // barriers visible to the peer too
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int chunks_per_tile = (num_k_blocks + kb_per_chunk - 1) / kb_per_chunk;

  if (warp == 0) {
    // ===================================================== TMA producer (one per CTA)
    if (lane == 0) {
      // one plane set (A == B): both operands go through the SAME two descriptors
      const CUtensorMap* mb_hi = a.same_planes ? &map_hi : &map_b_hi;
      const CUtensorMap* mb_lo = a.same_planes ? &map_lo : &map_b_lo;
      uint32_t it = 0;
      for (int t = unit; t < num_tiles; t += num_units) {
        const int m0 = tiles[t].m0 + (int)cta_rank * BM;             // this CTA's rows of A
        const int nb0 = tiles[t].n0 + (int)cta_rank * G::B_ROWS;     // this CTA's rows of B
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          if (sync_every > 0 && (it % sync_every) == 0) {
            // Soft rendezvous of all TMA producers: keeps the CTAs that share operand tiles
            // within the L2 residency window of each other (measured: without it they drift
            // apart over a 1.4 ms tile and every operand tile is re-read from HBM).
            // Timing only -- bounded wait, then proceed: cannot deadlock, carries no data.
            const uint32_t step = it / sync_every;
            const int seq = (t - unit) / num_units;
            const uint32_t expected = (uint32_t)(CTAS * min(num_units, num_tiles - seq * num_units));
            atomicAdd(&sync_ctr[step], 1u);
            for (int spin = 0; spin < 256; ++spin) {
              if (*reinterpret_cast<volatile unsigned int*>(&sync_ctr[step]) >= expected) break;
              __nanosleep(64);
            }
          }
          mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* st = smem + s * G::STAGE_BYTES;
          // the leader's barrier counts the bytes of BOTH CTAs of a pair
          if (cta_rank == 0) mbar_expect_tx(&full_bar[s], G::STAGE_BYTES * CTAS);
          const int k0 = kb * bk_of(PREC);
          tma_load_2d<CTAS>(st, &map_hi, &full_bar[s], k0, m0);                               // A hi
          tma_load_2d<CTAS>(st + A_BYTES, &map_lo, &full_bar[s], k0, m0);                     // A lo
          unsigned char* sb = st + 2 * A_BYTES;
          tma_load_2d<CTAS>(sb, mb_hi, &full_bar[s], k0, nb0);                                // B hi
          tma_load_2d<CTAS>(sb + G::B_BYTES, mb_lo, &full_bar[s], k0, nb0);                   // B lo
          if constexpr (CTAS == 1) {                                                         // rows [128,256) of B
            tma_load_2d<CTAS>(sb + A_BYTES, mb_hi, &full_bar[s], k0, nb0 + BM);
            tma_load_2d<CTAS>(sb + G::B_BYTES + A_BYTES, mb_lo, &full_bar[s], k0, nb0 + BM);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread; pair: the leader's)
    if (lane == 0 && cta_rank == 0) {
      uint32_t it = 0, chunk_it = 0;
      for (int t = unit; t < num_tiles; t += num_units) {
        for (int ch = 0; ch < chunks_per_tile; ++ch, ++chunk_it) {
          const uint32_t buf = chunk_it & 1, tph = (chunk_it >> 1) & 1;
          mbar_wait(&tempty_bar[buf], tph ^ 1);           // epilogue(s) drained this accumulator
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + buf * BN;
          const int kb_end = min(num_k_blocks, (ch + 1) * kb_per_chunk);
          for (int kb = ch * kb_per_chunk; kb < kb_end; ++kb, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tcgen05_fence_after();
            const uint32_t st = smem_u32(smem + s * G::STAGE_BYTES);
            const uint64_t a_hi = make_kmajor_sw128_desc(st);
            const uint64_t a_lo = make_kmajor_sw128_desc(st + A_BYTES);
            const uint64_t b_hi = make_kmajor_sw128_desc(st + 2 * A_BYTES);
            const uint64_t b_lo = make_kmajor_sw128_desc(st + 2 * A_BYTES + G::B_BYTES);
            constexpr uint32_t kIdesc = idesc_of(PREC, CTAS);
#pragma unroll
            for (int j = 0; j < ROW_BYTES / UMMA_K_BYTES; ++j) {
              const uint64_t adv = (uint64_t)((j * UMMA_K_BYTES) >> 4);   // +32 B per k-step
              const uint32_t first = (kb == ch * kb_per_chunk && j == 0) ? 0u : 1u;
              // small cross terms first, the dominant hi.hi term last
              umma<PREC, CTAS>(d_tmem, a_lo + adv, b_hi + adv, kIdesc, first);
              umma<PREC, CTAS>(d_tmem, a_hi + adv, b_lo + adv, kIdesc, 1u);
              umma<PREC, CTAS>(d_tmem, a_hi + adv, b_hi + adv, kIdesc, 1u);
            }
            umma_commit<CTAS>(&empty_bar[s]);             // frees the smem stage (of both CTAs) when MMAs retire
          }
          umma_commit<CTAS>(&tfull_bar[buf]);             // accumulator chunk complete (in both CTAs)
        }
      }
    }
  } else {
    // ===================================================== epilogue warps
    // TMEM lane quarter = warp % 4 (hardware rule); the two warps of a quarter split the 256
    // accumulator columns. The fp32 running sums of the K chunks live in REGISTERS (128 per
    // thread): a chunk costs 4 x (tcgen05.ld + 32 FADD) per warp, no memory traffic.
    // (pair: each CTA's TMEM holds ITS 128 rows of the 256 x 256 accumulator.)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* trans = s_trans + (warp - 2) * 32 * TRANS_PAD;
    uint32_t chunk_it = 0;
    for (int t = unit; t < num_tiles; t += num_units) {
      const int m0 = tiles[t].m0 + (int)cta_rank * BM, n0 = tiles[t].n0 + half * (BN / 2);
      // The chunk partials are folded in FIXED POINT: unit rows give |sum| <= 1, so an int32 in units of 2^-30
      // (9.3e-10) holds the running sum and the adds are exact; each partial is rounded to that grid once
      // (<= 4.7e-10). An fp32 running sum loses up to half an ulp of the SUM (3e-8) per add, and between
      // neighbouring rows of a sparse high-resolution map that does not average out: the few chunks near the
      // diagonal carry almost all of the sum, and the hundreds of small, nearly equal background partials
      // that follow fall below half an ulp of it and are rounded away one after the other
      // (tools/accuracy_split.py at 25 kb: fold 3.8e-6 mean / 6.8e-6 max of a 1.1e-5 total at k_chunk 512,
      // 1.6e-5 mean at k_chunk 128 -- a shorter chunk made it worse).
      // A NaN partial (all-zero row: every chunk is NaN) is caught on the last chunk and kept as INT_MIN.
      int accr[BN / 2];
#pragma unroll
      for (int j = 0; j < BN / 2; ++j) accr[j] = 0;
      for (int ch = 0; ch < chunks_per_tile; ++ch, ++chunk_it) {
        const uint32_t buf = chunk_it & 1, tph = (chunk_it >> 1) & 1;
        mbar_wait(&tfull_bar[buf], tph);
        tcgen05_fence_after();
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * (BN / 2);
        const bool last = ch == chunks_per_tile - 1;
#pragma unroll
        for (int cg = 0; cg < BN / 64; ++cg) {
          uint32_t v[32];
          tmem_ld32(taddr0 + cg * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            accr[cg * 32 + j] += (FixScale<PREC>::value == 1.f)
                                     ? __float2int_rn(__uint_as_float(v[j]))
                                     : __float2int_rn(__uint_as_float(v[j]) * FixScale<PREC>::value);
          }
          if (last) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float pv = __uint_as_float(v[j]);
              if (pv != pv) accr[cg * 32 + j] = kFixNaN;
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {                                    // accumulator buffer is free again
          if constexpr (CTAS == 2) mbar_arrive_cluster(&tempty_bar[buf], 0);
          else mbar_arrive(&tempty_bar[buf]);
        }
      }
      // ---- finalize + store (overlaps the next tile's MMAs)
      const int gr_base = m0 + q * 32;                      // first A row of this warp
      const float diag_val = (flavour == HIA_SIM_PEARSON) ? 1.f : 0.f;       // distances: exact 0 diagonal
      float* __restrict__ out = a.out;
      float* __restrict__ out_t = a.out_t;
      const long long ld_out = a.ld_out, ld_t = a.ld_t;
      const int row0 = a.row0, row_end = a.row_end, col0 = a.col0, col_end = a.col_end;
      const int has_diag = a.has_diag;
#pragma unroll
      for (int cg = 0; cg < BN / 64; ++cg) {
        const int gc_base = n0 + cg * 32;
        // symmetric: skip column groups entirely below the diagonal for this warp's rows
        const bool needed = symmetric ? (gc_base + 31 >= gr_base) : true;
        if (needed && gr_base < row_end && gc_base < col_end) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int fx = accr[cg * 32 + j];
            const float raw = (fx == kFixNaN) ? __int_as_float(0x7fc00000) : (float)fx * kFixUnscale;
            float c = fminf(1.f, fmaxf(-1.f, raw));         // |cos| clipped to 1 like scipy
            if (raw != raw) c = raw;                        // NaN stays NaN
            f[j] = (flavour == HIA_SIM_PEARSON) ? c : (1.f - c);
          }
          const int gr = gr_base + lane;
          if (out_t != nullptr && gr < row_end) {
            // transposed store out_t[c][r] (lanes = consecutive r -> one 128-byte line per column)
            float* tp = out_t + (long long)(gc_base - col0) * ld_t + (gr - row0);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int gc = gc_base + j;
              if (gc < col_end && (!symmetric || gc > gr)) tp[(long long)j * ld_t] = f[j];
            }
          }
          // direct store out[r][c] through a 32x32 shared transpose (coalesced rows)
#pragma unroll
          for (int j = 0; j < 32; ++j) trans[lane * TRANS_PAD + j] = f[j];
          __syncwarp();
          const int gc = gc_base + lane;
          if (gc < col_end) {
#pragma unroll 4
            for (int i = 0; i < 32; ++i) {
              const int r = gr_base + i;
              if (r < row_end && (gc >= r || !symmetric))
                out[(long long)(r - row0) * ld_out + (gc - col0)] =
                    (has_diag && gc == r) ? diag_val : trans[i * TRANS_PAD + lane];
            }
          }
          __syncwarp();
        }
      }
    }
  }

  __syncwarp();
  tcgen05_fence_before();
  // pair: neither CTA may exit (or free TMEM) while the other can still touch its barriers / smem
  if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ row prologue
// One CTA per row: fp64 sum and sum of squares in ONE pass (128-bit loads; the row then sits in
// L1/L2), mean / norm, then the hi/lo planes of the unit row in a second pass
// (TF32-exact fp32 planes, or fp16 planes of the row scaled by 2^15).
template <int PREC> struct PlaneT { typedef float type; };
template <> struct PlaneT<HIA_PREC_3XF16> { typedef __half type; };

// hi/lo planes of a value that already carries the plane's scale (2^15 for fp16, 1 for tf32)
template <int PREC>
__device__ __forceinline__ void split_planes(float xs, typename PlaneT<PREC>::type& h, typename PlaneT<PREC>::type& l) {
  if constexpr (PREC == HIA_PREC_3XF16) {
    h = __float2half_rn(xs);
    l = __float2half_rn(xs - __half2float(h));            // the difference is exact in fp32
  } else {
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(xs));
    h = __uint_as_float(hb);
    const float rem = xs - h;                             // exact in fp32
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
    l = __uint_as_float(lb);
  }
}

template <int PREC>
__device__ __forceinline__ void split_hi_lo(float xn, typename PlaneT<PREC>::type& h, typename PlaneT<PREC>::type& l) {
  if constexpr (PREC == HIA_PREC_3XF16) {
    const float xs = xn * kF16Scale;                      // exact (power of two)
    h = __float2half_rn(xs);
    l = __float2half_rn(xs - __half2float(h));            // the difference is exact in fp32
  } else {
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(xn));
    h = __uint_as_float(hb);
    const float rem = xn - h;                             // exact in fp32
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(rem));
    l = __uint_as_float(lb);
  }
}

// ncu (N = 30 894): the first version read the map TWICE from DRAM (7.4 GB): 148 x 8 CTAs x 123 KB rows
// in flight = 146 MB > the 126 MB L2, so the second pass missed. Now 512-thread CTAs (<= 4 per SM:
// 73 MB of rows in flight) and the second pass walks the row BACKWARDS, i.e. most-recently-read
// first, which is what an LRU-like L2 still holds.
constexpr int kPrepThreads = 512;

// Instruction budget (ncu, 30 894 bins: 49 % issue utilisation at ~1.9 GHz, i.e. issue-bound at the ~1.35 GHz
// the step runs at): the first version converted every cell to fp64 twice (statistics pass, then
// ((double)x - mean) * inv) -- 64-bit conversions and fp64 adds issue at a fraction of the fp32 rate. Now
// the statistics are fp32 inside a float4 (x - shift, 4-term sum and sum of squares) and fp64 across
// float4s (one conversion pair per 4 cells; the per-cell roundings are unbiased and average out, ~1e-9 of
// the row's spread on the mean), and the unit row is ((x - mean_hi) - mean_lo) * inv in fp32 with the mean
// split into two floats: 2.4e-7 relative per cell, the size of the hi/lo split's own quantisation, random
// in sign -- the Pearson sum averages it to ~1e-9 (measured stripe error unchanged). fp16 pairs are
// converted with the packed 2-wide instructions.
template <int PREC, int THREADS = kPrepThreads>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS)
row_prep_kernel(const float* __restrict__ x, int m, int k, long long ld_x, int kpad, int center,
                typename PlaneT<PREC>::type* __restrict__ hi, typename PlaneT<PREC>::type* __restrict__ lo) {
  typedef typename PlaneT<PREC>::type P;
  __shared__ double s_red[THREADS / 32][2];
  __shared__ float s_val[3];
  const int r = blockIdx.x;
  if (r >= m) return;
  const float* row = x + (long long)r * ld_x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & 15u) == 0);
  // Shift by an ESTIMATE OF THE MEAN (32 cells spread over the row, warp 0): sum((x - s)^2) - n (mean - s)^2
  // cancels by a factor ((mean - s) / sigma)^2, and with fp32 squares the rounding of the (repeated)
  // background value of a sparse row is the same in every cell -- with s = row[0] (a count cell, several sigma
  // from the mean) that put 1e-6 into the norm of 25 kb rows (tools/accuracy_split.py: representation error
  // 8e-7 mean / 3.6e-6 max against 6e-9 with fp64 squares).
  __shared__ float s_shift;
  if (w == 0) {
    float v = row[(int)(((long long)lane * k) >> 5)];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_shift = (v == v && fabsf(v) < INFINITY) ? v * (1.f / 32.f) : row[0];
  }
  __syncthreads();
  const float shift = s_shift;
  double s1 = 0.0, s2 = 0.0;
  const unsigned long long keep = l2_policy_evict_last();
  auto one = [&](float v) { const float a = v - shift; s1 += (double)a; s2 += (double)(a * a); };
  if (vec_ok) {
    const int k4 = k >> 2;
    for (int j = tid; j < k4; j += THREADS) {
      const float4 v = ld_keep4(row + 4 * j, keep);        // stays in L2 for the second pass
      const float a = v.x - shift, b = v.y - shift, c = v.z - shift, d = v.w - shift;
      s1 += (double)((a + b) + (c + d));
      s2 += (double)((a * a + b * b) + (c * c + d * d));
    }
    for (int j = (k4 << 2) + tid; j < k; j += THREADS) one(row[j]);
  } else {
    for (int j = tid; j < k; j += THREADS) one(row[j]);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) { s_red[w][0] = s1; s_red[w][1] = s2; }
  __syncthreads();
  if (tid == 0) {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < THREADS / 32; ++i) { t1 += s_red[i][0]; t2 += s_red[i][1]; }
    // Pearson: centre on the mean; cosine: centre on 0 (= shift back)
    const double mean_s = center ? t1 / (double)k : -(double)shift;      // mean of (x - shift)
    const double ss = t2 - 2.0 * mean_s * t1 + (double)k * mean_s * mean_s;   // sum (x - shift - mean_s)^2
    const double mean = mean_s + (double)shift;
    const double inv = 1.0 / sqrt(ss > 0.0 ? ss : 0.0);  // inf for an all-zero row -> NaN, like pdist
    const float mean_hi = (float)mean;
    s_val[0] = mean_hi;
    s_val[1] = (float)(mean - (double)mean_hi);
    s_val[2] = (float)(PREC == HIA_PREC_3XF16 ? inv * (double)kF16Scale : inv);   // the fp16 planes hold 2^15 x
  }
  __syncthreads();
  const float mean_hi = s_val[0], mean_lo = s_val[1], inv = s_val[2];
  auto unit = [&](float v) { return ((v - mean_hi) - mean_lo) * inv; };
  P* ph = hi + (long long)r * kpad;
  P* pl = lo + (long long)r * kpad;
  // second pass, back to front: 4 cells per thread (128-bit loads; 128-/64-bit plane stores; kpad % 4
  // == 0 and the plane rows are 16-byte aligned by construction)
  if (vec_ok) {
    const int n4 = kpad >> 2;
    for (int j4 = n4 - 1 - tid; j4 >= 0; j4 -= THREADS) {
      const int j = j4 << 2;
      float xs[4] = {0.f, 0.f, 0.f, 0.f};
      if (j + 3 < k) {
        const float4 v = ld_last4(row + j);                  // last use: evict first
        xs[0] = unit(v.x); xs[1] = unit(v.y); xs[2] = unit(v.z); xs[3] = unit(v.w);
      } else {
        for (int e = 0; e < 4; ++e) if (j + e < k) xs[e] = unit(row[j + e]);   // beyond k: zero padding
      }
      // streaming stores: the planes must not evict the rows that still wait for their second pass
      if constexpr (PREC == HIA_PREC_3XF16) {
        const __half2 h01 = __floats2half2_rn(xs[0], xs[1]), h23 = __floats2half2_rn(xs[2], xs[3]);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(xs[0] - f01.x, xs[1] - f01.y);   // the differences are exact in fp32
        const __half2 l23 = __floats2half2_rn(xs[2] - f23.x, xs[3] - f23.y);
        uint2 qh, ql;
        qh.x = *reinterpret_cast<const unsigned int*>(&h01); qh.y = *reinterpret_cast<const unsigned int*>(&h23);
        ql.x = *reinterpret_cast<const unsigned int*>(&l01); ql.y = *reinterpret_cast<const unsigned int*>(&l23);
        st_cs_u2(ph + j, qh);
        st_cs_u2(pl + j, ql);
      } else {
        P h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_planes<PREC>(xs[e], h[e], l[e]);
        st_cs_f4(ph + j, *reinterpret_cast<const float4*>(h));
        st_cs_f4(pl + j, *reinterpret_cast<const float4*>(l));
      }
    }
  } else {
    for (int j = kpad - 1 - tid; j >= 0; j -= THREADS) {
      P h = P(0.f), l = P(0.f);
      if (j < k) split_planes<PREC>(unit(row[j]), h, l);
      ph[j] = h;
      pl[j] = l;
    }
  }
}

// Register-staged variant (opt-in, HIA_PREP_REG=1; rows 16-byte aligned and <= 262 144 long):
// the row is read from DRAM exactly ONCE. A cluster of <= 8 CTAs of T = 256 / 512 / 1024 threads owns a
// row (T = the smallest that covers the row with 8 CTAs: small CTAs, several per SM, so that the
// load / reduce / store phases of different rows overlap on one SM -- one 1024-thread CTA per SM was
// 2.35 ms at 30 894 bins against 2.0 ms of the two-pass kernel); every thread loads 8 float4 (all
// 128-bit loads issued before the first use), accumulates the fp64 sum / sum of squares of its 32 cells, the cluster combines the CTA
// partials through distributed shared memory (every CTA stores its pair into every CTA of the
// cluster, one cluster barrier, fixed summation order), and the hi/lo planes are produced straight
// from the registers. 4 + 4 (f16) or 4 + 8 (tf32) bytes per cell, nothing re-read.
constexpr int kRegVec = 8;                                   // float4 per thread

template <int PREC, int kRegThreads>
__global__ void __launch_bounds__(kRegThreads, 1024 / kRegThreads)
row_prep_reg_kernel(const float* __restrict__ x, int m, int k, long long ld_x, int kpad, int center,
                    typename PlaneT<PREC>::type* __restrict__ hi, typename PlaneT<PREC>::type* __restrict__ lo) {
  typedef typename PlaneT<PREC>::type P;
  __shared__ double s_red[kRegThreads / 32][2];
  __shared__ double s_part[8][2];                            // partials of the (<= 8) CTAs of the cluster
  uint32_t cs, cr;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cr));
  const int r = blockIdx.x / (int)cs;                        // grid = m * cs: never beyond the last row
  const float* row = x + (long long)r * ld_x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n4 = kpad >> 2;                                  // float4 groups of the plane row
  const int g0 = (int)cr * kRegThreads * kRegVec;            // first group of this CTA
  const float shift_f = __ldg(row);
  const double shift = (double)shift_f;
  float4 v[kRegVec];
  // cells beyond the row take the value of the shift: they add exactly 0 to both sums (no masks below)
#pragma unroll
  for (int i = 0; i < kRegVec; ++i) {
    const int j = (g0 + tid + i * kRegThreads) << 2;
    if (j + 3 < k) v[i] = ld_stream4(row + j);
    else {
      v[i].x = (j < k) ? row[j] : shift_f;
      v[i].y = (j + 1 < k) ? row[j + 1] : shift_f;
      v[i].z = (j + 2 < k) ? row[j + 2] : shift_f;
      v[i].w = shift_f;
    }
  }
  double s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int i = 0; i < kRegVec; ++i) {
    const double a = (double)v[i].x - shift, b = (double)v[i].y - shift, c = (double)v[i].z - shift,
                 d = (double)v[i].w - shift;
    s1 += (a + b) + (c + d);
    s2 += (a * a + b * b) + (c * c + d * d);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) { s_red[w][0] = s1; s_red[w][1] = s2; }
  __syncthreads();
  if (tid == 0) {
    double t1 = 0, t2 = 0;
    for (int i = 0; i < kRegThreads / 32; ++i) { t1 += s_red[i][0]; t2 += s_red[i][1]; }
    if (cs == 1) { s_part[0][0] = t1; s_part[0][1] = t2; }
    else {
      const uint32_t local = smem_u32(&s_part[cr][0]);
      for (uint32_t dst = 0; dst < cs; ++dst) {              // my pair into slot `cr` of every CTA
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(dst));
        asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(remote), "d"(t1) : "memory");
        asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(remote + 8), "d"(t2) : "memory");
      }
    }
  }
  if (cs == 1) __syncthreads(); else cluster_sync_all();
  double t1 = 0.0, t2 = 0.0;
  for (uint32_t i = 0; i < cs; ++i) { t1 += s_part[i][0]; t2 += s_part[i][1]; }
  // Pearson: centre on the mean; cosine: centre on 0 (= shift back)
  const double mean_s = center ? t1 / (double)k : -shift;
  const double ss = t2 - 2.0 * mean_s * t1 + (double)k * mean_s * mean_s;
  const double mean = mean_s + shift;
  const double inv = 1.0 / sqrt(ss > 0.0 ? ss : 0.0);        // inf for an all-zero row -> NaN, like pdist
  P* ph = hi + (long long)r * kpad;
  P* pl = lo + (long long)r * kpad;
#pragma unroll
  for (int i = 0; i < kRegVec; ++i) {
    const int g = g0 + tid + i * kRegThreads;
    if (g >= n4) continue;
    const int j = g << 2;
    const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    P h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (j + q < k) split_hi_lo<PREC>((float)(((double)e[q] - mean) * inv), h[q], l[q]);
      else { h[q] = P(0.f); l[q] = P(0.f); }
    }
    if constexpr (PREC == HIA_PREC_3XF16) {
      *reinterpret_cast<uint2*>(ph + j) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(pl + j) = *reinterpret_cast<const uint2*>(l);
    } else {
      *reinterpret_cast<float4*>(ph + j) = *reinterpret_cast<const float4*>(h);
      *reinterpret_cast<float4*>(pl + j) = *reinterpret_cast<const float4*>(l);
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_plane_map(CUtensorMap* map, void* base, int rows, int kpad, int prec) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !p) { set_last_cuda_error(e); return HIA_ERR_CUDA; }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kpad * elem_bytes(prec)};
  cuuint32_t box[2] = {(cuuint32_t)bk_of(prec), (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, prec == HIA_PREC_3XF16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  2, base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_cuda_error(cudaErrorInvalidValue); return HIA_ERR_CUDA; }
  return HIA_OK;
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline bool split_prec(int p) { return p == HIA_PREC_3XTF32 || p == HIA_PREC_3XF16; }

// The tile list: tile_m x BN tiles of A rows [row0, row_end) x B rows [col0, col_end) in 2048 x 2048
// super-blocks (concurrently running CTAs then share operand row blocks through L2). symmetric: only
// tiles that touch the upper triangle and the requested |i - j| band are kept. The list is built ON
// THE DEVICE (one CTA, ordered compaction): no pageable host -> device copy that would synchronise the
// host with the stream on every call (and would make the call impossible to capture in a CUDA graph);
// the host only COUNTS the tiles (same predicate) to size the grid.
struct TileSpace {
  int row0, row_end, col0, col_end, tile_m, symmetric, band_lo, band_hi;
  __host__ __device__ int mt() const { return (row_end - row0 + tile_m - 1) / tile_m; }
  __host__ __device__ int nt() const { return (col_end - col0 + BN - 1) / BN; }
  __host__ __device__ int sbm() const { return 2048 / tile_m; }
  __host__ __device__ int nsbn() const { return (nt() + 2048 / BN - 1) / (2048 / BN); }
  __host__ __device__ int nsbm() const { return (mt() + sbm() - 1) / sbm(); }
  __host__ __device__ long long candidates() const { return (long long)nsbm() * nsbn() * sbm() * (2048 / BN); }
  // candidate c (super-block major, then row tile, then column tile) -> kept?, coordinates
  __host__ __device__ bool tile(long long c, TileCoord* t) const {
    const int SBN = 2048 / BN, per_sb = sbm() * SBN;
    const int sb = (int)(c / per_sb), w = (int)(c % per_sb);
    const int mi = (sb / nsbn()) * sbm() + w / SBN, nj = (sb % nsbn()) * SBN + w % SBN;
    if (mi >= mt() || nj >= nt()) return false;
    const int r0 = row0 + mi * tile_m, c0 = col0 + nj * BN;
    if (symmetric) {
      const int r1 = (r0 + tile_m < row_end ? r0 + tile_m : row_end) - 1;
      const int c1 = (c0 + BN < col_end ? c0 + BN : col_end) - 1;
      if (c1 < r0) return false;                         // entirely below the diagonal
      // distance range of the upper cells of the tile: [max(0, c0 - r1), c1 - r0]
      const int dmin = c0 - r1 > 0 ? c0 - r1 : 0, dmax = c1 - r0;
      if (band_hi >= 0 && (dmin > band_hi || dmax < band_lo)) return false;
    }
    t->m0 = r0;
    t->n0 = c0;
    return true;
  }
  int count() const {
    if (!symmetric) return mt() * nt();
    int cnt = 0;
    TileCoord t;
    for (long long c = 0, e = candidates(); c < e; ++c) cnt += tile(c, &t) ? 1 : 0;
    return cnt;
  }
};

__global__ void __launch_bounds__(1024) tile_list_kernel(const TileSpace ts, TileCoord* __restrict__ dst) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  const long long total = ts.candidates();
  for (long long c0 = 0; c0 < total; c0 += 1024) {
    TileCoord t;
    const bool keep = (c0 + tid < total) && ts.tile(c0 + tid, &t);
    const unsigned int bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int before = 0, all = 0;
    for (int i = 0; i < 32; ++i) { const int v = s_warp[i]; before += (i < w) ? v : 0; all += v; }
    if (keep) dst[s_base + before + __popc(bal & ((1u << lane) - 1u))] = t;
    __syncthreads();
    if (tid == 0) s_base += all;
    __syncthreads();
  }
}

template <int PREC, int T>
int launch_prep_reg(const float* x, int rows, int k, long long ld_x, int kpad, int center, void* hi, void* lo,
                    int cs, cudaStream_t stream) {
  typedef typename PlaneT<PREC>::type P;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)rows * cs, 1, 1);
  cfg.blockDim = dim3(T, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  HIA_CUDA(cudaLaunchKernelEx(&cfg, row_prep_reg_kernel<PREC, T>, x, rows, k, ld_x, kpad, center,
                              static_cast<P*>(hi), static_cast<P*>(lo)));
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}
template <int PREC>
int launch_prep_reg_t(const float* x, int rows, int k, long long ld_x, int kpad, int center, void* hi, void* lo,
                      cudaStream_t stream) {
  const int cells256 = 256 * kRegVec * 4;                  // 8 192 cells per 256-thread CTA
  auto cs_of = [&](int t) { return (kpad + t * kRegVec * 4 - 1) / (t * kRegVec * 4); };
  if (kpad <= 8 * cells256) return launch_prep_reg<PREC, 256>(x, rows, k, ld_x, kpad, center, hi, lo, cs_of(256), stream);
  if (kpad <= 16 * cells256) return launch_prep_reg<PREC, 512>(x, rows, k, ld_x, kpad, center, hi, lo, cs_of(512), stream);
  return launch_prep_reg<PREC, 1024>(x, rows, k, ld_x, kpad, center, hi, lo, cs_of(1024), stream);
}

int launch_prep(const float* x, int rows, int k, long long ld_x, int center, void* hi, void* lo, int prec,
                cudaStream_t stream) {
  const int kpad = round_up(k, bk_of(prec));
  // HIA_PREP_REG=1: one DRAM read per cell, rows staged in the registers of a CTA cluster. Measured at
  // 30 894 bins: 2.28 ms against 1.95 ms of the two-pass kernel below (whose second pass mostly hits L2),
  // so the two-pass kernel stays the default.
  const char* env_reg = getenv("HIA_PREP_REG");           // read per call: tools time both in one process
  const bool use_reg = env_reg && atoi(env_reg) == 1;
  if (use_reg && kpad <= 8 * 1024 * kRegVec * 4 && aligned16(x) && (ld_x & 3) == 0) {
    if (prec == HIA_PREC_3XF16) return launch_prep_reg_t<HIA_PREC_3XF16>(x, rows, k, ld_x, kpad, center, hi, lo, stream);
    return launch_prep_reg_t<HIA_PREC_3XTF32>(x, rows, k, ld_x, kpad, center, hi, lo, stream);
  }
  // Rows in flight = CTAs per SM x SMs x row bytes must stay inside L2 for the second pass to hit it.
  // HIA_PREP_CTAS (1 .. 4, default 4 = the register limit) caps the CTAs per SM by padding the launch with
  // unused dynamic shared memory (tuning knob; measured in profiles/).
  const char* e_ctas = getenv("HIA_PREP_CTAS");
  const int ctas = e_ctas ? max(1, min(4, atoi(e_ctas))) : 4;
  const size_t pad = ctas >= 4 ? 0 : (size_t)(200 * 1024 / ctas);
  if (pad) {
    HIA_CUDA(cudaFuncSetAttribute(row_prep_kernel<HIA_PREC_3XF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HIA_CUDA(cudaFuncSetAttribute(row_prep_kernel<HIA_PREC_3XTF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  // long rows: 1024-thread CTAs, 2 per SM -- the same threads per SM with half the rows in flight, so more of
  // the second pass hits L2 (30 894 bins: 1.61 against 1.63 ms in-step; HIA_PREP_THREADS=512 keeps the small CTAs)
  const char* e_thr = getenv("HIA_PREP_THREADS");
  const bool big_ctas = e_thr ? atoi(e_thr) == 1024 : k >= 16384;
  if (prec == HIA_PREC_3XF16 && big_ctas && !pad)
    row_prep_kernel<HIA_PREC_3XF16, 1024><<<rows, 1024, 0, stream>>>(x, rows, k, ld_x, kpad, center,
                                                                     static_cast<__half*>(hi), static_cast<__half*>(lo));
  else if (prec == HIA_PREC_3XF16)
    row_prep_kernel<HIA_PREC_3XF16><<<rows, kPrepThreads, pad, stream>>>(x, rows, k, ld_x, kpad, center,
                                                               static_cast<__half*>(hi), static_cast<__half*>(lo));
  else
    row_prep_kernel<HIA_PREC_3XTF32><<<rows, kPrepThreads, pad, stream>>>(x, rows, k, ld_x, kpad, center,
                                                                static_cast<float*>(hi), static_cast<float*>(lo));
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

}  // namespace

int similarity_fp64(const float* x, int m, int k, long long ld_x, float* out, long long ld_out,
                    int flavour, void* workspace, long long workspace_bytes, cudaStream_t stream);
long long similarity_fp64_workspace(int m, int k);

}  // namespace hia

// workspace layout (split precisions): [hi plane][lo plane][tile list][rendezvous counters]
constexpr int kSyncEvery = 32;          // k-blocks between producer rendezvous
static int64_t sync_counters(int m, int k) {
  using namespace hia;
  const int64_t ntiles = (int64_t)((m + BM - 1) / BM) * ((m + BN - 1) / BN);
  const int64_t per_cta = ntiles / num_sms() + 2;
  return per_cta * (round_up(k, 32) / 32) / kSyncEvery + 64;      // sized for the smaller k-block
}
static int64_t split_ws_layout(int m, int k, int prec, int64_t* off_lo, int64_t* off_tiles,
                               int64_t* off_sync = nullptr) {
  using namespace hia;
  const int64_t kpad = round_up(k, bk_of(prec));
  const int64_t plane = (int64_t)m * kpad * elem_bytes(prec);
  const int64_t plane_al = (plane + 1023) / 1024 * 1024;
  const int64_t ntiles = (int64_t)((m + BM - 1) / BM) * ((m + BN - 1) / BN);
  if (off_lo) *off_lo = plane_al;
  if (off_tiles) *off_tiles = 2 * plane_al;
  const int64_t tiles_al = (ntiles * (int64_t)sizeof(TileCoord) + 1023) / 1024 * 1024;
  if (off_sync) *off_sync = 2 * plane_al + tiles_al;
  return 2 * plane_al + tiles_al + sync_counters(m, k) * 4 + 1024;
}

HIA_API int64_t hia_similarity_workspace_bytes(int32_t m, int32_t k, int precision) {
  if (m <= 0 || k <= 0) return 0;
  if (precision == HIA_PREC_FP64) return hia::similarity_fp64_workspace(m, k);
  if (!hia::split_prec(precision)) return 0;
  return split_ws_layout(m, k, precision, nullptr, nullptr);
}

static int similarity_impl(const float* x, int32_t m, int32_t k, int64_t ld_x, float* out,
                           int64_t ld_out, int flavour, int precision, int32_t band_lo,
                           int32_t band_hi, int32_t k_chunk, void* workspace,
                           int64_t workspace_bytes, void* stream_, int phases);
// One contraction job: out = f(A planes . B planes^T) on a range of A rows x B rows (see SyrkArgs).
struct GemmJob {
  void *a_hi, *a_lo;  int a_plane_rows;       // planes of the output's rows
  void *b_hi, *b_lo;  int b_plane_rows;       // planes of the output's columns (may be the same buffers)
  int prec, kpad;
  int row0, row_end, col0, col_end;           // plane-row ranges (row0 % tile_m == 0, col0 % 256 == 0)
  int symmetric, has_diag, band_lo, band_hi;
  float* out;  int64_t ld_out;                // out[(r - row0)][(c - col0)]
  float* out_t;  int64_t ld_t;                // optional transposed copy out_t[(c - col0)][(r - row0)]
  int flavour, k_chunk;
  hia::TileCoord* d_tiles;  unsigned int* sync_ctr;  int64_t n_sync;
};

static int launch_gemm(const GemmJob& j, cudaStream_t stream);
static int check_sm100();

HIA_API int hia_similarity(const float* x, int32_t m, int32_t k, int64_t ld_x, float* out,
                           int64_t ld_out, int flavour, int precision, int32_t band_lo,
                           int32_t band_hi, int32_t k_chunk, void* workspace,
                           int64_t workspace_bytes, void* stream) {
  return similarity_impl(x, m, k, ld_x, out, ld_out, flavour, precision, band_lo, band_hi, k_chunk,
                         workspace, workspace_bytes, stream, 3);
}
HIA_API int hia_similarity_prep(const float* x, int32_t m, int32_t k, int64_t ld_x, int flavour,
                                int precision, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!hia::split_prec(precision)) return HIA_ERR_BAD_ARG;
  return similarity_impl(x, m, k, ld_x, reinterpret_cast<float*>(workspace), m, flavour, precision,
                         0, -1, 0, workspace, workspace_bytes, stream, 1);
}
HIA_API int hia_similarity_gemm(int32_t m, int32_t k, float* out, int64_t ld_out, int flavour,
                                int precision, int32_t band_lo, int32_t band_hi, int32_t k_chunk,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  if (!hia::split_prec(precision)) return HIA_ERR_BAD_ARG;
  return similarity_impl(reinterpret_cast<const float*>(workspace), m, k, k, out, ld_out, flavour,
                         precision, band_lo, band_hi, k_chunk, workspace, workspace_bytes, stream, 2);
}

static int similarity_impl(const float* x, int32_t m, int32_t k, int64_t ld_x, float* out,
                           int64_t ld_out, int flavour, int precision, int32_t band_lo,
                           int32_t band_hi, int32_t k_chunk, void* workspace,
                           int64_t workspace_bytes, void* stream_, int phases) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !out || !workspace || m <= 0 || k <= 0 || ld_x < k || ld_out < m) return HIA_ERR_BAD_ARG;
  if (flavour != HIA_SIM_COSINE_DISTANCE && flavour != HIA_SIM_PEARSON && flavour != HIA_SIM_CORRELATION_DISTANCE)
    return HIA_ERR_BAD_ARG;
  if (precision != HIA_PREC_FP64 && !split_prec(precision)) return HIA_ERR_BAD_ARG;
  if (workspace_bytes < hia_similarity_workspace_bytes(m, k, precision)) return HIA_ERR_WORKSPACE;
  if (precision == HIA_PREC_FP64)
    return similarity_fp64(x, m, k, ld_x, out, ld_out, flavour, workspace, workspace_bytes, stream);
  if (!aligned16(workspace)) return HIA_ERR_ALIGN;
  if (band_hi >= 0 && band_lo > band_hi) return HIA_ERR_BAD_ARG;
  if (band_lo < 0) band_lo = 0;

  { const int st = check_sm100(); if (st != HIA_OK) return st; }

  int64_t off_lo, off_tiles, off_sync;
  split_ws_layout(m, k, precision, &off_lo, &off_tiles, &off_sync);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  void* hi = ws;
  void* lo = ws + off_lo;
  TileCoord* d_tiles = reinterpret_cast<TileCoord*>(ws + off_tiles);
  const int kpad = round_up(k, bk_of(precision));

  if (phases & 1) {
    const int st = launch_prep(x, m, k, ld_x, flavour != HIA_SIM_COSINE_DISTANCE ? 1 : 0, hi, lo, precision, stream);
    if (st != HIA_OK) return st;
  }
  if (!(phases & 2)) return HIA_OK;

  GemmJob j = {};
  j.a_hi = j.b_hi = hi;
  j.a_lo = j.b_lo = lo;
  j.a_plane_rows = j.b_plane_rows = m;
  j.prec = precision;
  j.kpad = kpad;
  j.row0 = j.col0 = 0;
  j.row_end = j.col_end = m;
  j.symmetric = j.has_diag = 1;
  j.band_lo = band_lo;
  j.band_hi = band_hi;
  j.out = j.out_t = out;                       // the lower triangle is the exact mirror
  j.ld_out = j.ld_t = ld_out;
  j.flavour = flavour;
  j.k_chunk = k_chunk;
  j.d_tiles = d_tiles;
  j.sync_ctr = reinterpret_cast<unsigned int*>(ws + off_sync);
  j.n_sync = sync_counters(m, k);
  return launch_gemm(j, stream);
}

// CTA pairs (cta_group::2) unless HIA_SYRK_CTAS=1 selects the single-CTA kernel.
static int syrk_ctas() {
  static const int v = [] {
    const char* e = getenv("HIA_SYRK_CTAS");
    return (e && atoi(e) == 1) ? 1 : 2;
  }();
  return v;
}

// Per (kernel instantiation, device): opt-in shared memory once, resident cluster count once.
template <int PREC, int CTAS>
static int syrk_units(int* units_out) {
  using namespace hia;
  using G = Geo<CTAS>;
  static int cached_units[16] = {0};
  int dev = 0;
  HIA_CUDA(cudaGetDevice(&dev));
  const int slot = dev & 15;
  int units = __atomic_load_n(&cached_units[slot], __ATOMIC_ACQUIRE);
  if (units == 0) {
    auto kern = syrk_split3_kernel<PREC, CTAS>;
    HIA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    const int sms = num_sms();
    units = sms / CTAS;
    if (CTAS == 2) {
      cudaLaunchConfig_t cfg = {};
      cfg.blockDim = dim3(NUM_THREADS, 1, 1);
      cfg.dynamicSmemBytes = G::SMEM_BYTES;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cfg.gridDim = dim3(sms / 2 * 2, 1, 1);
      int max_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0)
        units = min(units, max_clusters);                 // persistent: never more pairs than can be resident
      else
        (void)cudaGetLastError();
    }
    __atomic_store_n(&cached_units[slot], units, __ATOMIC_RELEASE);
  }
  *units_out = units;
  return HIA_OK;
}

template <int PREC, int CTAS>
static int launch_syrk(const CUtensorMap* maps, const hia::SyrkArgs& args, cudaStream_t stream) {
  using namespace hia;
  using G = Geo<CTAS>;
  auto kern = syrk_split3_kernel<PREC, CTAS>;
  int units = 0;
  const int st = syrk_units<PREC, CTAS>(&units);
  if (st != HIA_OK) return st;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = G::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (CTAS == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  units = min(units, args.num_tiles);
  cfg.gridDim = dim3(units * CTAS, 1, 1);
  HIA_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], args));
  HIA_LAUNCH_CHECK();
  return HIA_OK;
}

// Tile list (device) + tensor maps + launch. Everything is stream-ordered and capturable: no host
// allocation, no pageable copy, no per-call attribute / occupancy query.
static int launch_gemm(const GemmJob& j, cudaStream_t stream) {
  using namespace hia;
  const int ctas = syrk_ctas();
  TileSpace ts = {j.row0, j.row_end, j.col0, j.col_end, BM * ctas, j.symmetric, j.band_lo, j.band_hi};
  const int num_tiles = ts.count();
  if (num_tiles == 0) return HIA_OK;
  tile_list_kernel<<<1, 1024, 0, stream>>>(ts, j.d_tiles);
  HIA_LAUNCH_CHECK();

  CUtensorMap maps[4];
  int st = make_plane_map(&maps[0], j.a_hi, j.a_plane_rows, j.kpad, j.prec);
  if (st == HIA_OK) st = make_plane_map(&maps[1], j.a_lo, j.a_plane_rows, j.kpad, j.prec);
  if (st == HIA_OK) st = make_plane_map(&maps[2], j.b_hi, j.b_plane_rows, j.kpad, j.prec);
  if (st == HIA_OK) st = make_plane_map(&maps[3], j.b_lo, j.b_plane_rows, j.kpad, j.prec);
  if (st != HIA_OK) return st;

  const int bk = bk_of(j.prec);
  // The TMEM accumulate truncates, so its bias grows with the length of one chain: K = 20 000 in one chain gave
  // 1e-4. With the exact fixed-point fold of the chunk partials in the epilogue the error is proportional to
  // k_chunk and to nothing else -- 3xF16, neighbouring rows of a 25 kb count map (K = 123 541, |rho| ~ 0.67),
  // max / mean abs error: 1024 -> 6.1e-6 / 1.9e-6, 512 -> 3.0e-6 / 9.2e-7, 256 -> 1.6e-6 / 4.3e-7 (+3.5 %
  // kernel time), 128 -> 1.0e-6 / 2.7e-7 (epilogue-bound, +44 %); 100 kb (K = 30 894): 2.8e-7 at 512.
  // HIA_K_CHUNK overrides the default for sweeps (tools/accuracy_probe.py).
  static const int env_chunk = getenv("HIA_K_CHUNK") ? atoi(getenv("HIA_K_CHUNK")) : 0;
  const int k_chunk = j.k_chunk > 0 ? j.k_chunk
                                    : (env_chunk > 0 ? env_chunk : (j.prec == HIA_PREC_3XF16 ? 512 : 256));
  static const int sync_every = getenv("HIA_SYRK_SYNC") ? atoi(getenv("HIA_SYRK_SYNC")) : kSyncEvery;
  HIA_CUDA(cudaMemsetAsync(j.sync_ctr, 0, (size_t)j.n_sync * 4, stream));
  SyrkArgs a;
  a.tiles = j.d_tiles;
  a.num_tiles = num_tiles;
  a.num_k_blocks = j.kpad / bk;
  a.kb_per_chunk = max(1, k_chunk / bk);
  a.flavour = j.flavour;
  a.out = j.out;
  a.ld_out = j.ld_out;
  a.out_t = j.out_t;
  a.ld_t = j.ld_t;
  a.sync_ctr = j.sync_ctr;
  a.sync_every = (sync_every >= kSyncEvery) ? sync_every : 0;
  a.row0 = j.row0;
  a.row_end = j.row_end;
  a.col0 = j.col0;
  a.col_end = j.col_end;
  a.symmetric = j.symmetric;
  a.has_diag = j.has_diag;
  a.same_planes = (j.a_hi == j.b_hi && j.a_lo == j.b_lo && j.a_plane_rows == j.b_plane_rows) ? 1 : 0;
  if (j.prec == HIA_PREC_3XF16)
    return ctas == 2 ? launch_syrk<HIA_PREC_3XF16, 2>(maps, a, stream) : launch_syrk<HIA_PREC_3XF16, 1>(maps, a, stream);
  return ctas == 2 ? launch_syrk<HIA_PREC_3XTF32, 2>(maps, a, stream) : launch_syrk<HIA_PREC_3XTF32, 1>(maps, a, stream);
}

static int check_sm100() {
  static int major = 0;
  if (major == 0) {
    int dev = 0, mj = 0;
    HIA_CUDA(cudaGetDevice(&dev));
    HIA_CUDA(cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev));
    major = mj;
  }
  return major == 10 ? HIA_OK : HIA_ERR_ARCH;
}

// ------------------------------------------------------------------ row-sharded entry points
// (the genome-wide high-resolution similarity is the one step with an
//  exchange: every rank prepares the hi/lo planes of ITS rows, the planes are all-gathered by the
//  caller (NCCL), then every rank contracts its row block against all rows.)
HIA_API int32_t hia_sim_kpad(int32_t k, int precision) { return hia::round_up(k, hia::bk_of(precision)); }
HIA_API int32_t hia_sim_plane_elem_bytes(int precision) { return hia::elem_bytes(precision); }

HIA_API int hia_sim_prep_rows(const float* x_rows, int32_t rows, int32_t k, int64_t ld_x, int flavour,
                              int precision, void* hi_rows, void* lo_rows, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x_rows || !hi_rows || !lo_rows || rows <= 0 || k <= 0 || ld_x < k || !split_prec(precision))
    return HIA_ERR_BAD_ARG;
  if (!aligned16(hi_rows) || !aligned16(lo_rows)) return HIA_ERR_ALIGN;
  return launch_prep(x_rows, rows, k, ld_x, flavour != HIA_SIM_COSINE_DISTANCE ? 1 : 0, hi_rows, lo_rows, precision, stream);
}

HIA_API int64_t hia_sim_gemm_rows_workspace_bytes(int32_t n, int32_t k) {
  if (n <= 0 || k <= 0) return 0;
  const int64_t ntiles = (int64_t)((n + hia::BM - 1) / hia::BM) * ((n + hia::BN - 1) / hia::BN);
  return (ntiles * (int64_t)sizeof(hia::TileCoord) + 1023) / 1024 * 1024 + sync_counters(n, k) * 4 + 1024;
}

static int rows_ws(void* workspace, int64_t workspace_bytes, int32_t n, int32_t k, GemmJob* j) {
  using namespace hia;
  if (!aligned16(workspace)) return HIA_ERR_ALIGN;
  if (workspace_bytes < hia_sim_gemm_rows_workspace_bytes(n, k)) return HIA_ERR_WORKSPACE;
  const int64_t ntiles = (int64_t)((n + BM - 1) / BM) * ((n + BN - 1) / BN);
  const int64_t tiles_al = (ntiles * (int64_t)sizeof(TileCoord) + 1023) / 1024 * 1024;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  j->d_tiles = reinterpret_cast<TileCoord*>(ws);
  j->sync_ctr = reinterpret_cast<unsigned int*>(ws + tiles_al);
  j->n_sync = sync_counters(n, k);
  return HIA_OK;
}

HIA_API int hia_sim_gemm_rows(void* hi, void* lo, int precision, int32_t plane_rows, int32_t n, int32_t k,
                              int32_t row0, int32_t rows, float* out_rows, int64_t ld_out, int flavour,
                              int32_t k_chunk, void* workspace, int64_t workspace_bytes, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!hi || !lo || !out_rows || !workspace || n <= 0 || k <= 0 || plane_rows < n || rows <= 0 ||
      row0 < 0 || ld_out < n || !split_prec(precision))
    return HIA_ERR_BAD_ARG;
  if (row0 % BM) return HIA_ERR_ALIGN;                    // row blocks start on a 128-row boundary
  if (!aligned16(hi) || !aligned16(lo)) return HIA_ERR_ALIGN;
  int st = check_sm100();
  if (st != HIA_OK) return st;
  GemmJob j = {};
  st = rows_ws(workspace, workspace_bytes, n, k, &j);
  if (st != HIA_OK) return st;
  j.a_hi = j.b_hi = hi;
  j.a_lo = j.b_lo = lo;
  j.a_plane_rows = j.b_plane_rows = plane_rows;
  j.prec = precision;
  j.kpad = round_up(k, bk_of(precision));
  j.row0 = row0;
  j.row_end = min(row0 + rows, n);
  j.col0 = 0;
  j.col_end = n;
  j.symmetric = 0;
  j.has_diag = 1;
  j.band_lo = 0;
  j.band_hi = -1;
  j.out = out_rows;
  j.ld_out = ld_out;
  j.out_t = nullptr;
  j.ld_t = 0;
  j.flavour = flavour;
  j.k_chunk = k_chunk;
  return launch_gemm(j, stream);
}

// One block of the row-sharded contraction (the ring schedule of hiarch_b200/sharded.py):
//   out  [i - a_row0][j - b_row0] = f(A[i] . B[j])   i in [a_row0, a_row0 + a_rows), j in [b_row0, b_row0 + b_rows)
//   out_t[j - b_row0][i - a_row0] = the same value   (optional; may be a peer GPU's memory)
// A and B are separate plane sets (row indices are plane rows). symmetric != 0: B must be A and the
// ranges equal: only the upper triangle is contracted, out_t receives the strict upper cells
// transposed (pass out_t = out, ld_t = ld_out for a full symmetric block) and the diagonal is the
// exact constant.
HIA_API int hia_sim_gemm_block(void* a_hi, void* a_lo, int32_t a_plane_rows, void* b_hi, void* b_lo,
                               int32_t b_plane_rows, int precision, int32_t k, int32_t a_row0, int32_t a_rows,
                               int32_t b_row0, int32_t b_rows, int symmetric, float* out, int64_t ld_out,
                               float* out_t, int64_t ld_t, int flavour, int32_t k_chunk, void* workspace,
                               int64_t workspace_bytes, void* stream_) {
  using namespace hia;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a_hi || !a_lo || !b_hi || !b_lo || !out || !workspace || k <= 0 || a_rows <= 0 || b_rows <= 0 ||
      a_row0 < 0 || b_row0 < 0 || a_row0 + a_rows > a_plane_rows || b_row0 + b_rows > b_plane_rows ||
      ld_out < b_rows || (out_t && ld_t < a_rows) || !split_prec(precision))
    return HIA_ERR_BAD_ARG;
  if (symmetric && (a_hi != b_hi || a_lo != b_lo || a_row0 != b_row0 || a_rows != b_rows)) return HIA_ERR_BAD_ARG;
  if ((a_row0 % BM) || (b_row0 % BM)) return HIA_ERR_ALIGN;
  if (!aligned16(a_hi) || !aligned16(a_lo) || !aligned16(b_hi) || !aligned16(b_lo)) return HIA_ERR_ALIGN;
  int st = check_sm100();
  if (st != HIA_OK) return st;
  GemmJob j = {};
  st = rows_ws(workspace, workspace_bytes, max(a_rows, b_rows), k, &j);
  if (st != HIA_OK) return st;
  j.a_hi = a_hi;
  j.a_lo = a_lo;
  j.a_plane_rows = a_plane_rows;
  j.b_hi = b_hi;
  j.b_lo = b_lo;
  j.b_plane_rows = b_plane_rows;
  j.prec = precision;
  j.kpad = round_up(k, bk_of(precision));
  j.row0 = a_row0;
  j.row_end = a_row0 + a_rows;
  j.col0 = b_row0;
  j.col_end = b_row0 + b_rows;
  j.symmetric = symmetric ? 1 : 0;
  j.has_diag = symmetric ? 1 : 0;
  j.band_lo = 0;
  j.band_hi = -1;
  j.out = out;
  j.ld_out = ld_out;
  j.out_t = out_t;
  j.ld_t = ld_t;
  j.flavour = flavour;
  j.k_chunk = k_chunk;
  return launch_gemm(j, stream);
}
