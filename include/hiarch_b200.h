/*
 * hiarch_b200.h -- C ABI of the B200-native (sm_100a) HiArch matrix hot path.
 *
 * The reference (xjtu-omics/HiArch) has no FFI: its operator boundary is the set of Python
 * methods/functions listed in SURVEY.md section 8(a/b). Each entry point below names the
 * reference interface it replaces (file:line relative to /root/reference/publish/).
 * `INTEGRATION.md` shows the ctypes stub a reference maintainer would add per call site.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns all memory (inputs, outputs, workspaces) and keeps it alive until the
 *     stream is synchronised; the library allocates nothing persistent;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, re-entrant,
 *     no global state;
 *   - return value: 0 (HIA_OK) or a negative hia_status; nothing is thrown across the ABI;
 *   - dense maps are row-major fp32 with leading dimension `ld` (elements); `ld % 4 == 0`
 *     and 16-byte aligned bases are required where stated (128-bit loads / TMA);
 *   - chromosome layout is `chr_start[C+1]` (int32, device): chromosome c owns bins
 *     [chr_start[c], chr_start[c+1]) -- GenomeIndex.chr_seps (new_hic_class.py:92-102).
 */
#ifndef HIARCH_B200_H
#define HIARCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  HIA_OK = 0,
  HIA_ERR_BAD_ARG = -1,       /* null pointer, negative size, bad flag            */
  HIA_ERR_ALIGN = -2,         /* pointer / ld alignment requirement not met       */
  HIA_ERR_CUDA = -3,          /* a CUDA runtime call or launch failed             */
  HIA_ERR_UNSUPPORTED = -4,   /* size beyond what this build supports             */
  HIA_ERR_WORKSPACE = -5,     /* workspace too small                              */
  HIA_ERR_ARCH = -6           /* device is not sm_100                             */
} hia_status;

int hia_version(void);
const char* hia_status_string(int status);
/* last CUDA error text seen by this thread inside the library (for HIA_ERR_CUDA) */
const char* hia_last_cuda_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t hia_launch_count(void);

/* ------------------------------------------------------------------------------------ */
/* a1/a2  sparse (i, j, v) -> dense symmetric                                            */
/* replaces read_sps_file -> SpsHiCMtx (new_hic_class.py:2576-2602, 2152-2198),          */
/*          SpsHiCMtx.to_all (:2229-2246), SpsHiCMtx.to_dense (:2209-2224),              */
/*          ArrayHiCMtx.to_all (:1991-2008)                                              */
/* ------------------------------------------------------------------------------------ */
#define HIA_MTX_TRIU 0   /* entries with col < row are ignored; result mirrored */
#define HIA_MTX_ALL  1   /* entries stored as given                             */
#define HIA_MTX_TRIU_UPPER 2 /* like TRIU, HIA_FILL_ZERO only: the upper triangle is scattered, the lower
                              * triangle of `out` is left UNWRITTEN (for consumers that read the upper
                              * triangle only: hia_oe_expected_pass, hia_oe_apply(HIA_OE_FROM_UPPER)) */
#define HIA_FILL_ZERO      0  /* absent cells = 0                (has_neg == False)          */
#define HIA_FILL_MIN       1  /* absent cells = min(all entries) (has_neg, whole matrix)     */
#define HIA_FILL_BLOCK_MIN 2  /* absent cells = min(entries of that chromosome-pair block)   */
                              /* (yield_by_chro(triu=True) + to_dense, main_local.py:112-115) */
/* workspace: hia_scatter_workspace_bytes(n, n_chr) bytes, 16-byte aligned (min-fill keys + the row
 * pointers of the sorted-row path). Optional for HIA_FILL_ZERO: with a workspace, entries that are
 * sorted by row (any COO taken from a CSR matrix / a .mtx file of the pipeline; checked on the
 * device) are written row by row through shared memory -- no memset, no read-modify-write --
 * and any other order falls back, on the device, to zero fill + atomic adds; without a workspace
 * only the latter runs. Duplicate (i,j) are summed (coo->csr). Explicit zero entries are stored
 * as 0 (they are not "absent"). out is n x n, ld >= n. */
int64_t hia_scatter_workspace_bytes(int32_t n, int32_t n_chr);
int hia_scatter_coo_sym(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                        int32_t n, float* out, int64_t ld, int mtx_type, int fill_mode,
                        const int32_t* chr_start, int32_t n_chr, void* workspace, void* stream);
/* The same from CSR (SpsHiCMtx.matrix is a scipy CSR matrix, publish/new_hic_class.py:2158-2161):
 * row_ptr[n + 1] int64, column indices int32 (col_bytes = 4) or uint16 (col_bytes = 2, n <= 65 536),
 * values fp32; zero fill. 8 (6) bytes per entry instead of 12: the bytes that cross PCIe in the
 * end-to-end path. No row-pointer pass, no fallback (CSR rows are sorted by construction).
 * Row r of an 'ALL' map starts at column 0; a 'triu' row at the diagonal (earlier columns ignored). */
int hia_scatter_csr_sym(const int64_t* row_ptr, const void* cols, int col_bytes, const float* vals,
                        int32_t n, float* out, int64_t ld, int mtx_type, void* stream);
/* Zero-fill scatter + the inter-chromosome statistics of the O/E expected pass as a by-product of the
 * row-sorted fill (taken from the finished row segment while it is in shared memory: saves the re-read
 * of the upper triangle by hia_oe_expected_pass). inter_sum / inter_nnz / inter_minpos [n_chr^2] as the
 * expected pass defines them; *stats_valid (device int) = 1 when complete (row-sorted COO, any CSR),
 * 0 after the any-order fallback -- hand it to hia_oe_expected_pass_ex. n_chr <= 64. */
int hia_scatter_coo_sym_stats(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz, int32_t n,
                              float* out, int64_t ld, int mtx_type, const int32_t* chr_start, int32_t n_chr,
                              const int16_t* bin_chr, double* inter_sum, unsigned long long* inter_nnz,
                              float* inter_minpos, int32_t* stats_valid, void* workspace, void* stream);
int hia_scatter_csr_sym_stats(const int64_t* row_ptr, const void* cols, int col_bytes, const float* vals, int32_t n,
                              float* out, int64_t ld, int mtx_type, const int32_t* chr_start, int32_t n_chr,
                              const int16_t* bin_chr, double* inter_sum, unsigned long long* inter_nnz,
                              float* inter_minpos, int32_t* stats_valid, void* stream);



/* ------------------------------------------------------------------------------------ */
/* f1  coarse-graining / trimming as an index remap fused into the scatter                 */
/* replaces SpsHiCMtx.condense (new_hic_class.py:2426-2465), tid_off_low_bins (:1671-1702), */
/*          tid_off_zero_chros / keep_long_chros / get_multi_chros_mtx (row+column           */
/*          selection), as composed by preprocess_mtx (oe_map/S1_preprocess.py:48-83)        */
/* ------------------------------------------------------------------------------------ */
/* out[map[r]][map[c]] += v for every entry (index_map[n_fine] int32, -1 = bin dropped); a
 * 'triu' input is mirrored at the fine level. out (n_out x n_out) is zero-filled first. */
int hia_scatter_coo_remap(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                          int32_t n_fine, const int32_t* index_map, int32_t n_out, float* out,
                          int64_t ld, int mtx_type, void* stream);
/* counts[C*C] (u64): non-zero stored entries per chromosome pair of the symmetrised matrix
 * (HiCMtx.get_nonzero_ratio on the blocks of yield_by_chro, :1040-1065). */
int hia_coo_block_nnz(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                      int32_t n, int mtx_type, const int16_t* bin_chr, int32_t n_chr,
                      unsigned long long* counts, void* stream);
/* bin_sum[n_out] (fp64): row sums of the symmetrised, remapped matrix straight from the fine COO
 * (the statistic get_high_coverage_bins :1658-1669 thresholds; fp64 so that the kept-bin set
 * does not depend on the fp32 storage of the map). */
int hia_coo_bin_sums(const int32_t* rows, const int32_t* cols, const float* vals, int64_t nnz,
                     int32_t n_fine, const int32_t* index_map, int32_t n_out, int mtx_type,
                     double* bin_sum, void* stream);
/* row_sum[n] (fp64) and row_nnz[n] (i64) of a dense map (get_high_coverage_bins :1658-1669,
 * get_nonzero_ratio). */
int hia_bin_stats(const float* x, int32_t n, int64_t ld, double* row_sum, long long* row_nnz,
                  void* stream);

/* ------------------------------------------------------------------------------------ */
/* a4/a5  expected contacts per diagonal, O/E, log                                       */
/* replaces HiCMtx.obs_d_exp(mode='sgd_mean') (new_hic_class.py:1121-1290) and           */
/*          ArrayHiCMtx.log (:2076-2094); oe_map_core (oe_map/S2_oe.py:10-15)            */
/* ------------------------------------------------------------------------------------ */
/* Pass 1 (reads the upper triangle once): per intra block, per diagonal d: sum of the
 * upper-diagonal cells (fp64) and min positive cell; per inter block (a<b): sum, non-zero
 * count and min positive. The map must be symmetric ('all' type from the scatter).
 * diag_sum[N] fp64, diag_minpos[N] fp32 (indexed chr_start[c] + d), inter_sum[C*C] fp64,
 * inter_nnz[C*C] u64, inter_minpos[C*C] fp32 -- all zero/inf-initialised by the call.
 * bin_chr[N] (int16) = chromosome index of every bin; chr_start_host = host copy (grid sizing). */
int hia_oe_expected_pass(const float* map, int32_t n, int64_t ld, const int32_t* chr_start,
                         const int32_t* chr_start_host, int32_t n_chr, double* diag_sum,
                         float* diag_minpos, double* inter_sum, unsigned long long* inter_nnz,
                         float* inter_minpos, const int16_t* bin_chr, void* stream);

/* Pass 1b (tiny): SGD pooling + un-flushed-last-group rule + monotone clamp.
 * pool_gid[N], use_code[N] are the data-independent tables the host derives from
 * ScaledGenomicDistance (new_hic_class.py:788-1003, scalar lookup quirk :865-875):
 *   pool_gid[s+d]  = stored-group index diagonal d is pooled into, or -1;
 *   use_code[s+d]  = -1 -> mean of the main diagonal (d == 0);
 *                    2*g   -> stored mean of group g as is (fallback to the largest key);
 *                    2*g+1 -> stored mean of group g, 0 replaced by 1 (:1213).
 * Outputs: expected[N] fp64 (the reference's exp_counts), inv_expected[N] fp32 (1/exp, or 1
 * where exp <= 0 -> "skip", :1264-1267), inter_inv[C*C] fp32 (1/mean of non-zero cells, both
 * (a,b) and (b,a)), min_pos_oe[1] fp32 = smallest positive O/E value (for the log fill). */
int hia_oe_finalize(const double* diag_sum, const float* diag_minpos, const double* inter_sum,
                    const unsigned long long* inter_nnz, const float* inter_minpos,
                    const int32_t* chr_start, int32_t n_chr, int32_t n, const int32_t* pool_gid,
                    const int32_t* use_code, int counts_must_decrese, double* expected,
                    float* inv_expected, float* inter_inv, float* min_pos_oe, void* stream);

/* Pass 2 (streaming, 8 B/cell): out = map * inv (intra: per diagonal; inter: per block);
 * flags: HIA_OE_LOG -> x>0 ? ln(x+1) : min over positives of ln(x+1)   (ArrayHiCMtx.log)
 *        HIA_OE_ONLY_INTRA -> inter blocks are written as 0            (only_intra=True) */
#define HIA_OE_LOG 1
#define HIA_OE_ONLY_INTRA 2
/* hia_oe_apply only (whole maps): read the upper triangle of `map` only and write both halves of
 * `out` (the transposed half through shared memory). Same result bit for bit on a symmetric map;
 * pairs with HIA_MTX_TRIU_UPPER of the scatter (no mirror pass). Opt-in, see DESIGN.md 6. */
#define HIA_OE_FROM_UPPER 4
int hia_oe_apply(const float* map, float* out, int32_t n, int64_t ld, const int32_t* chr_start,
                 int32_t n_chr, const int16_t* bin_chr, const float* inv_expected,
                 const float* inter_inv, const float* min_pos_oe, int flags, void* stream);

/* Row-sharded O/E: a rank that holds global rows [row0, row0 + rows) of the map computes PARTIAL
 * statistics (sum the diag_sum / inter_sum / inter_nnz arrays and take the element-wise min of
 * diag_minpos / inter_minpos over the ranks -- two small all-reduces), runs hia_oe_finalize on the
 * reduced arrays (replicated) and applies to its own rows. The upper-triangle rule (column >=
 * row) makes the partial sums disjoint. */
/* hia_oe_expected_pass when the scatter may already hold the inter-chromosome statistics
 * (hia_scatter_*_sym_stats): *inter_valid (device int) != 0 -> kept, the inter part of the map is not read. */
int hia_oe_expected_pass_ex(const float* map, int32_t n, int64_t ld, const int32_t* chr_start,
                            const int32_t* chr_start_host, int32_t n_chr, double* diag_sum, float* diag_minpos,
                            double* inter_sum, unsigned long long* inter_nnz, float* inter_minpos,
                            const int16_t* bin_chr, const int32_t* inter_valid, void* stream);
int hia_oe_expected_pass_rows(const float* map_rows, int32_t n, int64_t ld, int32_t row0, int32_t rows,
                              const int32_t* chr_start, const int32_t* chr_start_host, int32_t n_chr,
                              double* diag_sum, float* diag_minpos, double* inter_sum,
                              unsigned long long* inter_nnz, float* inter_minpos,
                              const int16_t* bin_chr, void* stream);
int hia_oe_apply_rows(const float* map_rows, float* out_rows, int32_t n, int64_t ld, int32_t row0,
                      int32_t rows, const int32_t* chr_start, int32_t n_chr, const int16_t* bin_chr,
                      const float* inv_expected, const float* inter_inv, const float* min_pos_oe,
                      int flags, void* stream);
/* Raw O/E in fp64 (class-level obs_d_exp without log): out[r][c] = map[r][c] / expected[|c - r|]
 * inside a chromosome, / mean of the non-zero cells of the block between chromosomes, in double like
 * the reference (new_hic_class.py:1263-1282). `expected`, `inter_sum`, `inter_nnz` are the outputs of
 * hia_oe_expected_pass + hia_oe_finalize. flags: HIA_OE_ONLY_INTRA only. The fp32 map of
 * hia_oe_apply holds 1e-5 absolute up to values of ~50; this one everywhere. */
int hia_oe_apply_f64(const float* map, int32_t n, int64_t ld, double* out, int64_t ld_out,
                     const int32_t* chr_start, int32_t n_chr, const int16_t* bin_chr,
                     const double* expected, const double* inter_sum, const int64_t* inter_nnz,
                     int flags, void* stream);


/* Stand-alone ArrayHiCMtx.log (new_hic_class.py:2076-2094) for a map that did not come out of
 * hia_oe_apply: x>0 -> ln(x+1); x<=0 -> min over positives of ln(x+1). min_pos_ws: 1 float. */
int hia_log_map(const float* x, float* out, int32_t rows, int32_t cols, int64_t ld, float* min_pos_ws,
                void* stream);

/* ------------------------------------------------------------------------------------ */
/* a8  row x row similarity: cosine distance (get_adj_mtx, complex/src/S1_get_adj_mtx.py  */
/*     :8-14 == squareform(pdist(M,'cosine'))) and Pearson correlation (north-star)       */
/* ------------------------------------------------------------------------------------ */
#define HIA_SIM_COSINE_DISTANCE 0   /* out = 1 - cos, exact 0 diagonal (pdist/squareform) */
#define HIA_SIM_PEARSON 1           /* out = corr,     exact 1 diagonal                    */
#define HIA_SIM_CORRELATION_DISTANCE 2   /* out = 1 - corr, exact 0 diagonal (pdist 'correlation') */
#define HIA_PREC_3XTF32 0           /* tcgen05 kind::tf32, hi/lo split, fp32 accumulate    */
#define HIA_PREC_FP64 1             /* FP64 DMMA (mma.sync m8n8k4), fp64 accumulate        */
#define HIA_PREC_3XF16 2            /* tcgen05 kind::f16, hi/lo split of the row scaled by 2^15 (fp16 has
                                       TF32's 11-bit significand: same 22 bits, K = 16 per MMA instead of
                                       8 -> half the tensor time and half the plane bytes), fp32 accumulate */
/* X is m x k (ld_x), out is m x m (ld_out). band_lo/band_hi (in bins, band_hi < 0 = all):
 * only tiles intersecting band_lo <= |i-j| <= band_hi are computed (checkerboard needs
 * 5%..15% of the chromosome length only); other cells of `out` are left untouched.
 * k_chunk: K elements accumulated inside tensor memory before the partial sum is folded
 * into an fp32 running tile (0 = default). */
int64_t hia_similarity_workspace_bytes(int32_t m, int32_t k, int precision);
int hia_similarity(const float* x, int32_t m, int32_t k, int64_t ld_x, float* out, int64_t ld_out,
                   int flavour, int precision, int32_t band_lo, int32_t band_hi, int32_t k_chunk,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* The two phases of hia_similarity(precision = 3XTF32 | 3XF16) as separate calls, so that a
 * caller can time (or overlap) them: prep = fp64 row statistics + hi/lo planes into the workspace
 * (4 B/cell read + 8 (tf32) or 4 (f16) B/cell written, HBM-bound); gemm = the tcgen05 contraction
 * + epilogue (tensor-bound). The workspace must be sized for the same precision. */
int hia_similarity_prep(const float* x, int32_t m, int32_t k, int64_t ld_x, int flavour, int precision,
                        void* workspace, int64_t workspace_bytes, void* stream);
int hia_similarity_gemm(int32_t m, int32_t k, float* out, int64_t ld_out, int flavour, int precision,
                        int32_t band_lo, int32_t band_hi, int32_t k_chunk, void* workspace,
                        int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------ */
/* a6  whole-matrix percentiles (numpy 'linear'), clip, core mean/std, z-score           */
/* replaces tid_off_outliers (oe_map/S3_denoise.py:85-92), norm_to_zero_mean /            */
/*          remove_outliers / norm_de_mtx (oe_map/S4_norm_value.py:10-46), the percentile */
/*          of keep_high_cons (confor/.../U2_keep_top_cons.py:20-41)                      */
/* ------------------------------------------------------------------------------------ */
#define HIA_PCT_ALL 0   /* np.percentile(a, q)          */
#define HIA_PCT_POS 1   /* np.percentile(a[a > 0], q)   */
#define HIA_PCT_NEG 2   /* np.percentile(a[a < 0], q)   */
/* Up to 4 percentiles of a rows x cols fp32 matrix in one call (3 streaming passes, exact
 * order statistics by radix select, numpy _lerp interpolation). modes[nq] (int32) and
 * quantiles[nq] (fp64, = q/100 computed by the caller) are DEVICE arrays; out[nq] fp64 on the
 * device (NaN for an empty subset). */
int64_t hia_percentiles_workspace_bytes(void);
int hia_percentiles(const float* x, int32_t rows, int32_t cols, int64_t ld, const int32_t* modes,
                    const double* quantiles, int32_t nq, double* out, void* workspace,
                    int64_t workspace_bytes, void* stream);
/* out = x with x > *hi -> *hi, x < *lo -> *lo (either bound may be NULL). lo/hi device fp64. */
int hia_clip(const float* x, float* out, int32_t rows, int32_t cols, int64_t ld, const double* lo,
             const double* hi, void* stream);
/* mean_std[2] (device fp64) = mean and population std of the values strictly inside
 * (*vmin, *vmax) -- or of all values when vmin >= vmax or the core is empty
 * (S4_norm_value.py:14-22). workspace6: 6 doubles. */
int hia_core_moments(const float* x, int32_t rows, int32_t cols, int64_t ld, const double* vmin,
                     const double* vmax, double* mean_std, double* workspace6, void* stream);
/* out = (x - mean_std[0]) / mean_std[1] */
int hia_zscore(const float* x, float* out, int32_t rows, int32_t cols, int64_t ld,
               const double* mean_std, void* stream);
/* bounds2 = (-v, v), v = min(pcts2[0], -pcts2[1])   (remove_outliers, S4_norm_value.py:34-37) */
int hia_sym_clip_bounds(const double* pcts2, double* bounds2, void* stream);

/* ------------------------------------------------------------------------------------ */
/* a10  checkerboard strength: the diagonal / entropy loop of get_local                   */
/* replaces get_short_sims / get_accordance (complex/src/S2_get_entro.py:40-65) and the   */
/*          loop of get_local (complex/src/main_local.py:30-55)                           */
/* ------------------------------------------------------------------------------------ */
/* adj: symmetric n x n distance matrix (only the diagonals d_lo <= d < d_hi are read).
 * edges30: the 30 histogram edges (device fp64; np.linspace(0, 1.6, 30)). per = 2 (percent).
 * out2 (device fp64): [mean entropy or NaN, number of entropy chunks]. */
int64_t hia_band_entropy_workspace_bytes(int32_t n_diag);
int hia_band_entropy(const float* adj, int32_t n, int64_t ld, int32_t d_lo, int32_t d_hi,
                     const double* edges30, int32_t min_count, double per, double* out2,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------ */
/* a11  GF_S1 row profiles: keep_high_cons fused with get_inter_rowsum / get_intra_rowsum */
/* (confor/.../U2_keep_top_cons.py:20-41, S2_1_find_by_inter_rowsum.py:26-37,             */
/*  S2_3_find_by_intra_low.py:18-27) and smooth_rowsum (S2_1:42-55)                       */
/* ------------------------------------------------------------------------------------ */
/* f(x) = max(x - *threshold, 0)  (reverse: max(*threshold - x, 0)); profiles[N] fp64:
 *   inter[i] = sum_{j not in chr(i)} f / (#such j) / mean(f > 0)   (0 for a single chromosome)
 *   intra[i] = sum_{j in chr(i)} f / len(chr(i)) / mean(f > 0)     (norm = 0 skips the mean)
 * workspace: (2 N + 2) doubles. */
int hia_row_profiles(const float* x, int32_t n, int64_t ld, const int32_t* chr_start,
                     const int16_t* bin_chr, const double* threshold, int reverse, int norm,
                     double* inter_profile, double* intra_profile, double* workspace, void* stream);
/* per-chromosome box smoothing, odd length max(min_kernal_len, int(len * kernal_per)) | 1,
 * scipy 'reflect' boundary. in != out. */
int hia_smooth1d_reflect(const double* in, double* out, int32_t n, const int32_t* chr_start,
                         const int16_t* bin_chr, double kernal_per, int32_t min_kernal_len,
                         void* stream);

/* a13  intra-x response (confor/.../S2_2_find_by_intra_x.py:25-50, :116-126): for every     */
/* candidate centre c of an n x n intra block: sum(M * K_c) / sum(K_c), K_c = the X-shaped    */
/* line-distance kernel of half-width k_len = max(3, int(.05 n)). out[n] fp64 (device).       */
int hia_intra_x_response(const float* block, int32_t n, int64_t ld, int32_t k_len, double* out,
                         void* stream);
/* The same response in O(n^2) instead of O(n^2 k_len): along a row of M every centre's kernel  */
/* is a tent in the column index, so row prefix sums of (v, column * v) give each (row, centre)  */
/* term from <= 8 look-ups. Deterministic; fp64 throughout; `block` needs no alignment. The      */
/* direct kernel above stays as the on-device cross-check.                                       */
int64_t hia_intra_x_workspace_bytes(int32_t n);
int hia_intra_x_response_prefix(const float* block, int32_t n, int64_t ld, int32_t k_len, double* out,
                                void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------ */
/* a7  box (uniform) and median filters, scipy.ndimage 'reflect' boundary                 */
/* replaces denoise_one_method(mode='mean'|'median') (oe_map/S3_denoise.py:48-57), used by  */
/* used_denoise (:76-79) and used_filter (confor/src/S1_preprocess/S1_1_filtering.py:13-16) */
/* ------------------------------------------------------------------------------------ */
/* x / out / scratch: rows x cols views sharing the leading dimension ld; out may alias x for
 * the box filter (scratch must not); size <= 1 is the identity (like scipy). */
int hia_box_filter_reflect(const float* x, float* out, float* scratch, int32_t rows, int32_t cols,
                           int64_t ld, int32_t size, void* stream);
/* out != x. element of rank size*size/2 of each size x size window. */
int hia_median_filter_reflect(const float* x, float* out, int32_t rows, int32_t cols, int64_t ld,
                              int32_t size, void* stream);

/* GF_S1 used_filter (confor/src/S1_preprocess/S1_1_filtering.py:13-16) on one block as a single
 * op: box filter (size) + norm_to_zero_mean (mid_vmax_per = 66), fp64 inside with scipy's exact
 * running-sum order, because the core selection of norm_to_zero_mean is decided by ties (see
 * hia_filter64.cu). decimals >= 0: the fp32 input is first re-quantised to that many decimals in
 * fp64 (= the double a "%.2f" field parses to); -1: plain promotion. normalise = 0: box filter
 * only. x / out: rows x cols views (any alignment). */
int64_t hia_used_filter_f64_workspace_bytes(int32_t rows, int32_t cols);
int hia_used_filter_f64(const float* x, int32_t rows, int32_t cols, int64_t ld_in, int32_t size,
                        int32_t decimals, int normalise, double mid_vmax_per, float* out,
                        int64_t ld_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------ */
/* a14/a15  GF_S2: adaptive pool 16x16 (4 quadrants of 8x8 around a 'used' anchor),        */
/* z-score, flip-symmetrise, 5 template scores (inter: max with the transposed image)      */
/* replaces pool_mtx / norm_mtx / symmetry (confor/src/S2_to_train_mtx/preprocess.py:8-48) */
/*          and to_confor (confor/src/S5_sim_to_pat/to_confor.py:35-47)                    */
/* ------------------------------------------------------------------------------------ */
/* pairs8[P][8] int32 (device): {row0, n_rows, col0, n_cols, cen1 (-1 = none), cen2, intra, 0}.
 * tmpl_*: 5 x 256 fp64, already L1-normalised. Outputs (device fp64): pooled/isym/xs [P][256],
 * scores [P][5]. */
int hia_gf_pool_scores(const float* x, int64_t ld, const int32_t* pairs8, int32_t n_pairs,
                       const double* tmpl_intra, const double* tmpl_inter, double* pooled,
                       double* isym, double* xs, double* scores, void* stream);
/* Scores of n_img pooled 16 x 16 images (xs [n_img][256] fp64, device) against CALLER-SUPPLIED,
 * already normalised templates (tmpl [n_tmpl][256] fp64): scores[p][t] = xs[p] . tmpl[t]; do_trans:
 * max with the transposed image's score (inter pairs). to_confor with a user's template file,
 * confor/src/S5_sim_to_pat/to_confor.py:35-47. */
int hia_template_scores(const double* xs, int32_t n_img, const double* tmpl, int32_t n_tmpl, int do_trans,
                        double* scores, void* stream);


/* ------------------------------------------------------------------------------------ */
/* f4: wavelet denoise of one dense block. Replaces                                       */
/*   skimage.restoration.denoise_wavelet(arr, wavelet='db3', method="VisuShrink")         */
/*   at publish/oe_map/S3_denoise.py:43-47 (called per chromosome block, :76-79, :95-107) */
/* ------------------------------------------------------------------------------------ */
/* Multi-level separable 2-D db3 transform (half-sample symmetric extension, PyWavelets' dwt/idwt
 * conventions), sigma = median(|dd_finest| != 0) / 0.6745, threshold = sigma sqrt(2 ln(rows cols)),
 * soft threshold of every detail band, inverse transform cut to rows x cols. fp64 arithmetic.
 * The arithmetic is third-party and unversioned in the reference: oracle/wavelet_oracle.py writes the
 * definition down (parity unpinned).
 * levels <= 0: scikit-image's default, max(min over axes of floor(log2(n / 5)) - 3, 1).
 * in / out: device, rows x cols, elem_bytes = 4 (fp32) or 8 (fp64), must not overlap.
 * stats (device, 3 doubles): sigma, threshold, number of nonzero finest-diagonal coefficients. */
int32_t hia_wavelet_default_levels(int32_t rows, int32_t cols);
int64_t hia_wavelet_workspace_bytes(int32_t rows, int32_t cols, int32_t levels);
int hia_wavelet_denoise_db3(const void* in, int64_t ld_in, int32_t rows, int32_t cols, int32_t elem_bytes,
                            int32_t levels, void* out, int64_t ld_out, double* stats, void* ws,
                            int64_t ws_bytes, void* stream);


/* Row-sharded similarity (the one exchange step of the path): every rank
 * prepares the hi/lo planes of ITS rows (hia_sim_prep_rows; planes are row-major with
 * hia_sim_kpad(k, precision) columns of hia_sim_plane_elem_bytes(precision) bytes: fp32 for
 * 3XTF32, fp16 for 3XF16), the caller all-gathers the planes (NCCL), then every rank contracts
 * its row block [row0, row0 + rows) against all rows (hia_sim_gemm_rows; row0 % 128 == 0).
 * plane_rows >= n rows are addressable in hi/lo (rows >= n must be zero). out_rows: rows x n. */
int32_t hia_sim_kpad(int32_t k, int precision);
int32_t hia_sim_plane_elem_bytes(int precision);
int hia_sim_prep_rows(const float* x_rows, int32_t rows, int32_t k, int64_t ld_x, int flavour,
                      int precision, void* hi_rows, void* lo_rows, void* stream);
int64_t hia_sim_gemm_rows_workspace_bytes(int32_t n, int32_t k);
int hia_sim_gemm_rows(void* hi, void* lo, int precision, int32_t plane_rows, int32_t n, int32_t k,
                      int32_t row0, int32_t rows, float* out_rows, int64_t ld_out, int flavour,
                      int32_t k_chunk, void* workspace, int64_t workspace_bytes, void* stream);
/* One block of the row-sharded contraction (ring schedule of hiarch_b200/sharded.py; SURVEY.md
 * section 8e "by symmetry only s >= r tiles are needed"):
 *   out  [i - a_row0][j - b_row0] = f(A[i] . B[j])   i in [a_row0, a_row0 + a_rows), j in [b_row0, b_row0 + b_rows)
 *   out_t[j - b_row0][i - a_row0] = the same value   (optional, NULL = no transposed copy)
 * A (a_hi/a_lo, a_plane_rows addressable rows) and B are separate plane sets; a_row0 % 128 == 0,
 * b_row0 % 128 == 0. out_t may point into a PEER GPU's memory (hia_ipc_open): the mirrored half of the
 * symmetric result then leaves the epilogue over NVLink, there is no separate exchange step.
 * symmetric != 0 (B must be A, equal ranges): only the upper triangle is contracted, the diagonal
 * is the exact constant and out_t receives the strict upper cells transposed (out_t = out,
 * ld_t = ld_out gives the full symmetric block). Workspace: hia_sim_gemm_rows_workspace_bytes(
 * max(a_rows, b_rows), k). */
int hia_sim_gemm_block(void* a_hi, void* a_lo, int32_t a_plane_rows, void* b_hi, void* b_lo,
                       int32_t b_plane_rows, int precision, int32_t k, int32_t a_row0, int32_t a_rows,
                       int32_t b_row0, int32_t b_rows, int symmetric, float* out, int64_t ld_out,
                       float* out_t, int64_t ld_t, int flavour, int32_t k_chunk, void* workspace,
                       int64_t workspace_bytes, void* stream);

/* Peer memory of the row-sharded path (one process per GPU on one NVSwitch box): a rank exports a
 * device buffer (hia_ipc_export: 64-byte CUDA IPC handle of the allocation + the pointer's offset
 * in it), the other ranks map it (hia_ipc_open -> base pointer valid in the calling process until
 * hia_ipc_close). hia_peer_copy is a stream-ordered copy between any two such pointers; it runs on
 * a copy engine over NVLink, so it overlaps the tensor-core kernel without taking SMs from it. */
int hia_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset);
int hia_ipc_open(const void* handle64, void** base_out);
int hia_ipc_close(void* base);
int hia_peer_copy(void* dst, const void* src, int64_t bytes, void* stream);

/* ------------------------------------------------------------------------------------ */
/* a9  top-k eigenvectors of a dense symmetric fp32 matrix (largest algebraic).          */
/* No reference implementation exists (the reference never decomposes anything); the     */
/* north-star asks for compartment eigenvectors: oracle = numpy.linalg.eigh.             */
/* Block Lanczos (block in {4, 8}, `steps` Krylov steps, full re-orthogonalisation),      */
/* Rayleigh-Ritz by an on-device Jacobi solver. block * (steps + 1) <= min(n, 272).       */
/* init_vecs: optional n_init <= block start vectors (fp32, each contiguous, length n)    */
/* for restarts. Outputs (device): eigvals[k] fp64 descending, eigvecs[k][n] fp32 (each   */
/* vector contiguous, unit norm, sign arbitrary), resid[k] = ||C u - lambda u||_2 fp64.   */
/* Exact for positive semi-definite input (Pearson maps); for indefinite input pairs      */
/* +lambda/-lambda of equal magnitude may mix.                                            */
/* ------------------------------------------------------------------------------------ */
int64_t hia_topk_eig_workspace_bytes(int32_t n, int32_t block, int32_t steps);
/* early_rel_tol / early_angle_tol > 0: after every hia_eig_check_every(n)-th Lanczos step (from step
 * hia_eig_first_check(n, steps) on; n >= 16 384: step 7, then every step; smaller: step 8, every second) the
 * Ritz residuals are ESTIMATED from the Lanczos relation (no pass over the matrix; one tiny D2H
 * flag) and the cycle ends as soon as every wanted pair has est <= early_rel_tol |lambda_1| or
 * est <= early_angle_tol * gap (gap = distance to the nearest other Ritz value: the sin-theta
 * bound on the eigenvector angle). resid[k] are always the TRUE residuals ||C u - lambda u|| of
 * one final pass; gaps[k] (optional, device) the Ritz gaps; *steps_used_host (optional) the steps
 * taken. upper_ws (optional; hia_eig_symm_upper_workspace_bytes(n, block) bytes, 16-byte aligned):
 * every pass reads the UPPER TRIANGLE of the (symmetric) matrix only -- each loaded cell feeds
 * both y[r] += c v[col] and y[col] += c v[r] -- i.e. half the DRAM bytes of the HBM-bound pass;
 * the lower triangle is never touched. NULL = the full-matrix pass. */
/* The projected (Rayleigh-Ritz) eigenproblem of the Lanczos solver on its own (tests / tools): h is D x D
 * row-major, the lower triangle of its leading dim x dim block is the symmetric matrix (dim <= 272).
 * z: dim x dim doubles column-major, lam[dim], order[dim] (device): pair j (j-th largest, j < kk <= 16) is
 * lam[order[j]] with eigenvector column order[j]. solver 0: Householder tridiagonalisation + Sturm
 * multisection + inverse iteration for the kk largest pairs only (dim <= 160; lam[order[kk]] is the next
 * eigenvalue, the rest -1e300); solver 1: one-sided Jacobi (all pairs). scratch: dim x dim doubles. */
int hia_projected_eig(const double* h, int32_t D, int32_t dim, int32_t kk, int solver, double* z, double* lam,
                      int32_t* order, double* scratch, void* stream);
int64_t hia_eig_symm_upper_workspace_bytes(int32_t n, int32_t block);
/* k_keep (k <= k_keep <= block, 0 = k): Ritz pairs returned (eigvals / eigvecs / resid / gaps hold k_keep
 * entries); convergence is judged on the first k, the rest make better restart blocks. */
int hia_topk_eig(const float* c, int32_t n, int64_t ld, int32_t k, int32_t k_keep, int32_t block, int32_t steps,
                 uint64_t seed, const float* init_vecs, int32_t n_init, double early_rel_tol,
                 double early_angle_tol, double* eigvals, float* eigvecs, double* resid, double* gaps,
                 int32_t* steps_used_host, void* workspace, int64_t workspace_bytes, void* upper_ws,
                 int64_t upper_ws_bytes, void* stream);

/* The phases of hia_topk_eig as separate calls, so that a row-sharded caller can all-gather the
 * panel between the pass over the matrix and the (replicated, tiny) orthogonalisation:
 *   hia_eig_begin      start block (restart vectors + random), orthonormalised
 *   hia_eig_symm_rows  y_rows[rows x block] (row-major fp64) = C[rows, :] * current fp32 panel
 *   hia_eig_absorb     step j: y_full[n x block] (row-major, all rows) -> H[:, j], Q_{j+1}
 *   hia_eig_check      after steps_done steps: Jacobi + Lanczos residual estimates -> *converged_host
 *                      (synchronises the stream; est_gap_host optional: k estimates + k gaps)
 *   hia_eig_ritz       Jacobi on the first steps_used blocks (skipped when reuse_check != 0: the last
 *                      hia_eig_check ran with steps_done == steps_used) -> eigvals[k], eigvecs[k][n],
 *                      gaps[k] (optional); panel = [U, 0]
 *   hia_eig_resid      y_full = C [U, 0] -> resid[k] = ||C u - lambda u||
 * All take the same (n, block, steps) and the workspace of hia_topk_eig_workspace_bytes. */
int hia_eig_begin(int32_t n, int32_t block, int32_t steps, uint64_t seed, const float* init_vecs,
                  int32_t n_init, void* workspace, int64_t workspace_bytes, void* stream);
int hia_eig_symm_rows(const float* c_rows, int32_t rows, int32_t n, int64_t ld, int32_t block,
                      int32_t steps, void* workspace, double* y_rows, void* stream);
int hia_eig_absorb(int32_t n, int32_t block, int32_t steps, int32_t j, const double* y_full,
                   void* workspace, void* stream);
int32_t hia_eig_first_check(int32_t n, int32_t steps);   /* first pass-free convergence check after this many steps */
int32_t hia_eig_check_every(int32_t n);                  /* steps between two checks after the first one */
int hia_eig_check(int32_t n, int32_t block, int32_t steps, int32_t steps_done, int32_t k, double rel_tol,
                  double angle_tol, int32_t* converged_host, double* est_gap_host, void* workspace,
                  void* stream);
int hia_eig_ritz(int32_t n, int32_t block, int32_t steps, int32_t steps_used, int32_t reuse_check,
                 int32_t k, double* eigvals, float* eigvecs, double* gaps, void* workspace, void* stream);
/* row-sharded callers: after hia_eig_ritz, broadcast eigvals / eigvecs / gaps from one rank and
 * install them with hia_eig_set_ritz, so that the panel of the residual pass (and the restart
 * vectors) are identical on every rank even when clustered Ritz values order differently. */
int hia_eig_set_ritz(int32_t n, int32_t block, int32_t steps, int32_t k, const float* eigvecs,
                     void* workspace, void* stream);
int hia_eig_resid(int32_t n, int32_t block, int32_t steps, int32_t k, const double* y_full,
                  const double* eigvals, double* resid, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------ */
/* f2  the text wire format between the pipeline's processes (HOST functions, HOST memory) */
/* replaces pd.read_table(sps_file, names=['row','col','value']) + dropna (read_sps_file,  */
/*          new_hic_class.py:2580-2581) and np.savetxt(fmt="%d\t%d\t%.2f")                  */
/*          (SpsHiCMtx.to_output, new_hic_class.py:2290-2297)                              */
/* ------------------------------------------------------------------------------------ */
/* Exact statistics: entries the file holds after blank-line skipping and dropna (>= 0), or a negative
 * hia_status (cannot open, or a field that is neither a number nor a missing-value marker).
 * max_row / max_col / has_neg are optional outputs. threads <= 0: all hardware threads. */
int64_t hia_mtx_scan(const char* path, int32_t threads, int64_t* max_row, int64_t* max_col,
                     int32_t* has_neg);
/* Upper bound on the entries: the number of lines (no parsing). */
int64_t hia_mtx_count_lines(const char* path, int32_t threads);
/* Parse (one pass) into caller-owned host arrays of `capacity` entries (>= the entries of the file:
 * hia_mtx_count_lines or hia_mtx_scan), file order kept; *n_out = entries written. Any output
 * pointer may be NULL; rows32 / cols32 / vals32 are the int32 / fp32 forms the device kernels take
 * (point them at pinned memory to upload without another pass). */
int hia_mtx_parse(const char* path, int32_t threads, int64_t capacity, int64_t* rows, int64_t* cols,
                  double* vals, int32_t* rows32, int32_t* cols32, float* vals32, int64_t* n_out);
/* value_format 0: "%d\t%d\t%.2f" (correctly rounded like CPython / glibc); 1: "%d\t%d\t%d". */
int hia_mtx_write(const char* path, int64_t n, const int64_t* rows, const int64_t* cols, const double* vals,
                  int value_format, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* HIARCH_B200_H */
