"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy, fp64) of HiArch's matrix hot path.

This is the *oracle* the CUDA path is checked against. It is never imported by the product
package `hiarch_b200/`; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it (as the checker / the CPU baseline).

Parity pin: every function here is checked against the reference's own code, imported
read-only through `oracle/ref_shim.py`, by `tests/test_oracle_vs_reference.py` (runs where
`/root/reference` exists) and against the committed outputs of that same reference code in
`tests/golden/*.npz` (made by `tests/golden/gen_golden.py`; runs everywhere).
Exceptions, stated: `topk_eig` has NO reference implementation (the reference contains no
eigen-decomposition) -> "parity unpinned" by the reference, oracle = numpy.linalg.eigh; the
wavelet denoise step is not restated (scikit-image absent and unpinned).

All `file:line` citations are relative to /root/reference/publish/.
All arithmetic is fp64 like the reference (new_hic_class.py:1227 `astype(float)`).
"""
import numpy as np

# ----------------------------------------------------------------------------------------
# a1/a2  sparse (i, j, v)  ->  dense symmetric
# ----------------------------------------------------------------------------------------

def scatter_coo_sym(rows, cols, vals, n, mtx_type="triu", has_neg=False, seps=None):
    """COO -> CSR (duplicates summed) -> dense -> 'all' (symmetric).

    new_hic_class.py:2596 (coo_matrix), :2158-2161 (tocsr sums duplicates),
    :2209-2224 (to_dense: absent cells are 0, or min(data) when has_neg; *explicit* zero
    entries such as a "0.00" line of a %.2f file stay 0), :1991-2008 / :2229-2246 (to_all:
    triu(M,0).T + triu(M,1)).

    seps=None : one fill value = min over all stored entries (main_filter path,
                confor/src/S1_preprocess/S1_1_filtering.py:21-24).
    seps given: every chromosome-pair block is densified on its own, so its fill value is the
                min over the entries stored in *that* block (checkerboard path,
                complex/src/main_local.py:112-115 via yield_by_chro(triu=True))."""
    import scipy.sparse as sps
    m = sps.coo_matrix((np.asarray(vals, dtype=np.float64),
                        (np.asarray(rows), np.asarray(cols))), shape=(n, n)).tocsr().tocoo()
    r, c, v = m.row, m.col, m.data
    if mtx_type == "tril":
        r, c = c, r
    elif mtx_type == "triu":
        keep = c >= r
        r, c, v = r[keep], c[keep], v[keep]
    elif mtx_type != "all":
        raise ValueError(mtx_type)
    dense = np.zeros((n, n))
    if has_neg and len(v):
        if seps is None:
            dense[:] = np.min(v)
        else:
            for (s1, e1) in seps:
                for (s2, e2) in seps:
                    if s2 < s1 and mtx_type != "all":
                        continue
                    inb = (r >= s1) & (r <= e1) & (c >= s2) & (c <= e2)
                    if inb.any():
                        dense[s1:e1 + 1, s2:e2 + 1] = np.min(v[inb])
    dense[r, c] = v
    if mtx_type != "all":
        dense = np.triu(dense, k=0).T + np.triu(dense, k=1)
    return dense


# ----------------------------------------------------------------------------------------
# a3  ScaledGenomicDistance
# ----------------------------------------------------------------------------------------

def sgd_size_dict(bin_size, n_bins, k=0.2):
    """Matrix-distance -> re-indexed SGD bin. new_hic_class.py:899-1000 as called from
    obs_d_exp (:1241-1244): s=bin_size, e=bin_size*n, auto_change_end=True."""
    s = bin_size
    e_in = bin_size * n_bins
    cur_max_size = (np.log2(e_in) + k) / k                      # :929
    e = np.power(2, k * np.floor(cur_max_size) - k)             # :930-931
    size_max = int(np.floor((np.log2(e) + k) / k)) + 1          # :940
    size_dict = {}
    for diff in range(int(s), int(e) + 1, int(bin_size)):       # :943
        size = np.floor((np.log2(diff) + k) / k)
        if size >= size_max:
            break
        size_dict[int(diff / bin_size)] = size
    raw_sizes = sorted(set(size_dict.values()))                 # :956-962
    reordered = {raw: i for i, raw in enumerate(raw_sizes)}
    return {d: reordered[v] for d, v in size_dict.items()}


def sgd_lookup(size_dict, dis, bin_size):
    """Scalar __getitem__, including the `dis >= bin_size` quirk. new_hic_class.py:865-875."""
    if not dis < bin_size:
        dis = dis / bin_size
    return size_dict.get(dis)


# ----------------------------------------------------------------------------------------
# a4  obs_d_exp(mode='sgd_mean')
# ----------------------------------------------------------------------------------------

def _diag_values(M, d, mtx_type="all"):
    """Values of 'diagonal d' in the order of yield_diag_idxes. new_hic_class.py:1538-1573."""
    n = M.shape[0]
    if d == 0:
        return np.diagonal(M).copy()
    idx = np.arange(n - d)
    if mtx_type == "all":
        return np.concatenate([M[idx + d, idx], M[idx, idx + d]])   # lower then upper
    if mtx_type == "triu":
        return M[idx, idx + d].copy()
    if mtx_type == "tril":
        return M[idx + d, idx].copy()
    raise ValueError(mtx_type)


def sgd_mean_expected(M, bin_size, k=0.2, mtx_type="all"):
    """Expected counts per diagonal. new_hic_class.py:1182-1216 (_sgd_mean).
    The last SGD group is never stored (no flush after the loop) - reproduced."""
    n = M.shape[0]
    sd = sgd_size_dict(bin_size, n, k)
    means = {}
    cur = None
    lines, num = [], 0
    for d in range(n):
        g = sgd_lookup(sd, d, bin_size)
        if g is None:
            continue
        vals = _diag_values(M, d, mtx_type)
        if cur is None or cur == g:
            lines.append(vals)
            num += len(vals)
            cur = g
        else:
            means[cur] = np.sum(np.concatenate(lines)) / num
            lines, num = [vals], len(vals)
            cur = g
    out = []
    for d in range(n):
        g = sgd_lookup(sd, d, bin_size)
        if g is None or g not in means:
            if d == 0:
                v = np.mean(np.diagonal(M))
            else:
                v = means[max(means.keys())]          # ValueError on empty, like np.max([])
        else:
            v = 1 if means[g] == 0 else means[g]
        out.append(v)
    return out


def monotone_clamp(exp_counts):
    """counts_must_decrese. new_hic_class.py:1251-1261."""
    exp_counts = list(exp_counts)
    cur = None
    for i in range(len(exp_counts)):
        if cur is None:
            if exp_counts[i] != 0:
                cur = exp_counts[i]
        elif exp_counts[i] > cur:
            exp_counts[i] = cur
        else:
            if exp_counts[i] != 0:
                cur = exp_counts[i]
    return exp_counts


def oe_intra_block(M, bin_size, k=0.2, counts_must_decrese=True, mtx_type="all"):
    """One intra-chromosomal block. new_hic_class.py:1224-1270. Returns (O/E block, expected)."""
    M = np.array(M, dtype=np.float64)
    n = M.shape[0]
    if n <= 1:
        return M / np.mean(M), None                              # :1229-1232
    if np.sum(M) == 0:
        return M, None                                           # :1234-1236
    exp = sgd_mean_expected(M, bin_size, k, mtx_type)
    if counts_must_decrese:
        exp = monotone_clamp(exp)
    idx0 = np.arange(n)
    for d in range(n):                                           # :1263-1268
        if exp[d] <= 0:
            continue
        if d == 0:
            M[idx0, idx0] = M[idx0, idx0] / exp[0]
            continue
        i = np.arange(n - d)
        if mtx_type in ("all", "tril"):
            M[i + d, i] = M[i + d, i] / exp[d]
        if mtx_type in ("all", "triu"):
            M[i, i + d] = M[i, i + d] / exp[d]
    return M, np.array(exp, dtype=np.float64)


def oe_inter_block(B):
    """Inter-chromosomal block: divide by the mean of the non-zero cells. :1272-1282."""
    B = np.array(B, dtype=np.float64)
    nz = np.nonzero(B)
    if len(nz[0]) > 0:
        B = B / np.mean(B[nz])
    return B


def obs_d_exp(dense, seps, bin_size, k=0.2, counts_must_decrese=True, only_intra=False,
              return_expected=False):
    """Whole-genome O/E. new_hic_class.py:1121-1290 with mode='sgd_mean' on an 'all'-type map.
    `seps` = [(start, end_inclusive)] per chromosome in index order (GenomeIndex.chr_seps)."""
    dense = np.asarray(dense, dtype=np.float64)
    out = np.zeros_like(dense)            # concat_mtxes -> coo -> dense: untouched cells are 0
    expected = []
    for a, (s1, e1) in enumerate(seps):
        for b, (s2, e2) in enumerate(seps):
            blk = dense[s1:e1 + 1, s2:e2 + 1]
            if a == b:
                o, ex = oe_intra_block(blk, bin_size, k, counts_must_decrese)
                expected.append(ex)
                out[s1:e1 + 1, s2:e2 + 1] = o
            elif not only_intra:
                out[s1:e1 + 1, s2:e2 + 1] = oe_inter_block(blk)
    if return_expected:
        return out, expected
    return out


# ----------------------------------------------------------------------------------------
# a5  log
# ----------------------------------------------------------------------------------------

def log_map(M, has_neg=False):
    """ArrayHiCMtx.log. new_hic_class.py:2076-2094."""
    M = np.asarray(M, dtype=np.float64)
    if has_neg:
        return M.copy()
    pos = M > 0
    out = M.copy()
    out[pos] = np.log(out[pos] + 1)
    out[M <= 0] = np.min(out[pos])
    return out


def oe_map_core(dense, seps, bin_size, counts_must_decrese=True):
    """oe_map/S2_oe.py:10-15: obs_d_exp(sgd_mean, k=.2) -> log."""
    return log_map(obs_d_exp(dense, seps, bin_size, 0.2, counts_must_decrese))


# ----------------------------------------------------------------------------------------
# a6  percentile clip / z-score / keep_high_cons
# ----------------------------------------------------------------------------------------

def tid_off_outliers(M, vmax_per=98):
    """oe_map/S3_denoise.py:85-92."""
    M = np.array(M, dtype=np.float64)
    vmax = np.percentile(M, vmax_per)
    vmin = np.percentile(M, 100 - vmax_per)
    M[M > vmax] = vmax
    M[M < vmin] = vmin
    return M


def norm_to_zero_mean(M, mid_vmax_per=66):
    """oe_map/S4_norm_value.py:10-27."""
    M = np.array(M, dtype=np.float64)
    core = M.copy()
    vmax = np.percentile(core, mid_vmax_per)
    vmin = np.percentile(core, 100 - mid_vmax_per)
    if vmin < vmax:
        sub = core[core < vmax]
        sub = sub[sub > vmin]
        if len(sub) > 0:
            core = sub
    M -= np.mean(core)
    M /= np.std(core)
    return M


def remove_outliers(M, default_vmax_per=98):
    """oe_map/S4_norm_value.py:30-40."""
    M = np.array(M, dtype=np.float64)
    vmax = np.percentile(M[M > 0], default_vmax_per)
    vmin = np.percentile(M[M < 0], 100 - default_vmax_per)
    v = np.min([vmax, -vmin])
    M[M > v] = v
    M[M < -v] = -v
    return M


def norm_de_mtx(M):
    """oe_map/S4_norm_value.py:43-46."""
    return remove_outliers(norm_to_zero_mean(M))


def keep_high_cons(M, per=100 / 3, reverse=False, norm=True):
    """confor/src/S1_preprocess/S1_2_find_anchors/U2_keep_top_cons.py:20-41."""
    M = np.asarray(M, dtype=np.float64)
    new = M.copy()
    if not reverse:
        t = np.percentile(new, 100 - per)
        new[M < t] = t
        new -= t
    else:
        t = np.percentile(new, per)
        new[M > t] = t
        new *= -1
        new += t
    if norm:
        new /= np.mean(new[new > 0])
    return new


# ----------------------------------------------------------------------------------------
# a7  box / median filters (scipy.ndimage *is* the reference arithmetic here)
# ----------------------------------------------------------------------------------------

def box_filter_reflect(M, size):
    """denoise_one_method(mode='mean'): uniform_filter(size, mode='reflect'). S3_denoise.py:53-57."""
    from scipy.ndimage import uniform_filter
    return uniform_filter(np.asarray(M, dtype=np.float64), size=size, mode="reflect")


def _reflect(i, n):
    p = 2 * n
    i %= p
    return i if i < n else p - 1 - i


def box_filter_replica(M, size):
    """Bit-for-bit restatement of scipy.ndimage.uniform_filter(mode='reflect') = two passes of
    NI_UniformFilter1D (axis 0 then axis 1): tmp = sum of the first window in order;
    tmp += new - old; out = tmp / size. The GPU fp64 path (csrc/hia_filter64.cu) follows this
    order; `tests/test_host_tables.py::test_uniform_filter_replica` pins it against scipy."""
    def one_axis(a, axis):
        a = np.moveaxis(np.asarray(a, dtype=np.float64), axis, 0).copy()
        n = a.shape[0]
        out = np.empty_like(a)
        s1 = size // 2
        s2 = size - s1 - 1
        ext = a[[_reflect(i, n) for i in range(-s1, n + s2)]]
        tmp = np.zeros(a.shape[1:])
        for ll in range(size):
            tmp = tmp + ext[ll]
        out[0] = tmp / float(size)
        for ll in range(1, n):
            tmp = tmp + (ext[ll + size - 1] - ext[ll - 1])
            out[ll] = tmp / float(size)
        return np.moveaxis(out, 0, axis)
    if size <= 1:
        return np.array(M, dtype=np.float64)
    return one_axis(one_axis(M, 0), 1)


def median_filter_reflect(M, size):
    """denoise_one_method(mode='median'): median_filter(size, mode='reflect'). S3_denoise.py:48-52."""
    from scipy.ndimage import median_filter
    return median_filter(np.asarray(M, dtype=np.float64), size=size, mode="reflect")


def main_denoise(M, seps, wavelet=None):
    """main_denoise, oe_map/S3_denoise.py:95-107 = tid_off_outliers, then per chromosome(-pair)
    block used_denoise (:76-79): [wavelet ->] median filter of size max(2, int(len(row chromosome)
    * .01)) ('reflect'). The wavelet (skimage db3 / VisuShrink, :45-47) is third-party, absent and
    unpinned: `wavelet` is a callable applied per block, None = identity ("wavelet=identity")."""
    M = tid_off_outliers(M)
    out = np.empty_like(M)
    for (s1, e1) in seps:
        size = int(np.mean(np.array([e1 - s1 + 1]) * .01))
        size = 2 if size < 2 else size
        for (s2, e2) in seps:
            blk = M[s1:e1 + 1, s2:e2 + 1]
            if wavelet is not None:
                blk = wavelet(blk)
            out[s1:e1 + 1, s2:e2 + 1] = median_filter_reflect(blk, size)
    return out


def filter_sizes(mean_chr_len):
    """(median size, box size): S3_denoise.py:50-51 with ratio .01 (:76-79); :55 with ratio .1."""
    med = int(np.mean(np.array([mean_chr_len]) * .01))
    med = 2 if med < 2 else med
    box = int(np.mean(np.array([mean_chr_len]) * .1))
    return med, box


def main_filter(dense, seps, has_neg=True):
    """confor/src/S1_preprocess/S1_1_filtering.py:13-32: per block box filter (size from the
    block's *row* chromosome length) + norm_to_zero_mean; then global remove_outliers.
    concat_mtxes drops exact zeros (they come back as 0 in the dense result)."""
    dense = np.asarray(dense, dtype=np.float64)
    out = np.zeros_like(dense)
    for (s1, e1) in seps:
        size = int(np.mean(np.array([e1 - s1 + 1]) * .1))
        for (s2, e2) in seps:
            blk = dense[s1:e1 + 1, s2:e2 + 1]
            out[s1:e1 + 1, s2:e2 + 1] = norm_to_zero_mean(box_filter_reflect(blk, size))
    return remove_outliers(out)


# ----------------------------------------------------------------------------------------
# a8  row x row cosine distance (and the Pearson flavour)
# ----------------------------------------------------------------------------------------

def cosine_distance(M):
    """squareform(pdist(M, 'cosine')). complex/src/S1_get_adj_mtx.py:8-14, restated as a GEMM:
    1 - x_i.x_j / (|x_i||x_j|), |cos| clipped to 1 like scipy's C kernel, exact-zero diagonal."""
    M = np.asarray(M, dtype=np.float64)
    nrm = np.sqrt(np.einsum("ij,ij->i", M, M))
    with np.errstate(invalid="ignore", divide="ignore"):
        c = (M @ M.T) / np.outer(nrm, nrm)
    c = np.where(np.abs(c) > 1, np.copysign(1.0, c), c)
    d = 1.0 - c
    d = 0.5 * (d + d.T)
    np.fill_diagonal(d, 0.0)
    return d


def pearson(M):
    """Pearson correlation of rows (north-star flavour; not in the reference): the same
    contraction on row-centred rows; unit diagonal."""
    M = np.asarray(M, dtype=np.float64)
    Mc = M - M.mean(axis=1, keepdims=True)
    nrm = np.sqrt(np.einsum("ij,ij->i", Mc, Mc))
    with np.errstate(invalid="ignore", divide="ignore"):
        c = (Mc @ Mc.T) / np.outer(nrm, nrm)
    c = np.clip(c, -1.0, 1.0)
    c = 0.5 * (c + c.T)
    np.fill_diagonal(c, 1.0)
    return c


# ----------------------------------------------------------------------------------------
# a9  top-k eigenvectors -- NO reference implementation: parity unpinned by the reference
# ----------------------------------------------------------------------------------------

def topk_eig(C, k=3):
    """Largest-algebraic k eigenpairs of a symmetric matrix (numpy.linalg.eigh).
    Returns (eigenvalues descending [k], eigenvectors [n, k])."""
    w, v = np.linalg.eigh(np.asarray(C, dtype=np.float64))
    order = np.argsort(w)[::-1][:k]
    return w[order], v[:, order]


# ----------------------------------------------------------------------------------------
# a10  checkerboard strength
# ----------------------------------------------------------------------------------------

HIST_BINS = np.linspace(0, 1.6, 30)          # complex/src/S2_get_entro.py:61
MIN_CHRO_LEN = 50                            # complex/src/main_local.py:16
MIN_N_SHORT_SIMS = 200                       # complex/src/main_local.py:17


def short_sims(adj, d, per=2):
    """get_short_sims(adj, d, d) on an 'all'-type matrix. S2_get_entro.py:40-58."""
    n = adj.shape[0]
    if d >= n:
        return None
    vals = _diag_values(adj, d, "all")
    if len(vals) < 1:
        return None
    lo = np.percentile(vals, per)
    hi = np.percentile(vals, 100 - per)
    return vals[(vals < hi) & (vals > lo)]


def accordance(vals):
    """get_accordance. S2_get_entro.py:61-65 (scipy.stats.entropy, natural log)."""
    from scipy.stats import entropy
    h, _ = np.histogram(vals, bins=HIST_BINS)
    return entropy(h / np.sum(h))


def checkerboard_from_adj(adj, L, short_dis_max_per=.15):
    """The diagonal/entropy loop of get_local. complex/src/main_local.py:30-55.
    L = mean chromosome length of the block's row window (S2_get_entro.py:6-10)."""
    smax = int(L * short_dis_max_per)
    smin = int(L * .05)
    ent, cur = [], []
    for d in range(smin, smax):
        sub = short_sims(adj, d)
        if sub is not None:
            cur += list(sub)
        if len(cur) > MIN_N_SHORT_SIMS:
            ent.append(accordance(cur))
            cur = []
    if len(ent) == 0:
        return None
    return float(np.mean(ent))


def get_local(M, L=None, short_dis_max_per=.15):
    """get_local. complex/src/main_local.py:19-55."""
    M = np.asarray(M, dtype=np.float64)
    if min(M.shape) < MIN_CHRO_LEN:
        return None
    L = M.shape[0] if L is None else L
    return checkerboard_from_adj(cosine_distance(M), L, short_dis_max_per)


def main_local(dense, seps, chros, short_dis_max_per=.15):
    """main_local on a dense min-filled 'all' map: intra + upper inter blocks, then the
    transposed inter block (skipped when the first orientation returned None).
    complex/src/main_local.py:108-141. Returns [(accord_chro, other_chro, accord)]."""
    out = []
    for a, (s1, e1) in enumerate(seps):
        for b, (s2, e2) in enumerate(seps):
            if s1 > s2:
                continue
            blk = dense[s1:e1 + 1, s2:e2 + 1]
            L = e1 - s1 + 1
            acc = get_local(blk, L, short_dis_max_per)
            if acc is None:
                continue
            out.append((chros[a], chros[b], acc))
            if a != b:
                acc = get_local(blk.T, L, short_dis_max_per)   # row_window stays chro1's
                if acc is None:
                    continue
                out.append((chros[b], chros[a], acc))
    return out


# ----------------------------------------------------------------------------------------
# a11  GF_S1 row profiles
# ----------------------------------------------------------------------------------------

def inter_rowsum(M, seps):
    """get_inter_rowsum. .../S2_1_find_by_inter_rowsum.py:26-37."""
    n = M.shape[0]
    out = np.zeros(n)
    for (s, e) in seps:
        other = np.hstack((M[s:e + 1, :s], M[s:e + 1, e + 1:]))
        if other.shape[1] == 0:
            continue
        out[s:e + 1] = other.sum(axis=1) / other.shape[1]
    return out


def intra_rowsum(M, seps):
    """get_intra_rowsum. .../S2_3_find_by_intra_low.py:18-27."""
    out = np.zeros(M.shape[1])
    for (s, e) in seps:
        blk = M[s:e + 1, s:e + 1]
        out[s:e + 1] = blk.sum(axis=1) / blk.shape[1]
    return out


def kernal_len(n, kp, mkl, only_odd=False):
    """get_kernal_len. .../S2_1_find_by_inter_rowsum.py:15-20."""
    kl = int(n * kp)
    kl = mkl if kl < mkl else kl
    if only_odd:
        kl = (kl // 2) * 2 + 1
    return kl


def smooth_rowsum(profile, seps):
    """smooth_rowsum: per chromosome box convolve, reflect. S2_1:42-55."""
    from scipy.ndimage import convolve
    out = np.zeros_like(profile, dtype=np.float64)
    for (s, e) in seps:
        seg = np.asarray(profile[s:e + 1], dtype=np.float64).reshape(-1, 1)
        kl = kernal_len(seg.shape[0], .1, 5, only_odd=True)
        out[s:e + 1] = convolve(seg, weights=np.ones((kl, 1)) / kl, mode="reflect")[:, 0]
    return out


# ----------------------------------------------------------------------------------------
# a13  intra-x response
# ----------------------------------------------------------------------------------------

def intra_x_kernal(cen_pos, n, k_len):
    """get_intra_x_kernal. .../S2_2_find_by_intra_x.py:25-50."""
    telo = [0, n]
    cen = [cen_pos, cen_pos]
    x, y = np.meshgrid(range(n), range(n))
    if cen[0] != telo[0]:
        k = (cen[1] - telo[1]) / (cen[0] - telo[0])
        b = telo[1]
        dis = np.abs(k * x - y + b) / np.sqrt(1 + k ** 2)
    else:
        dis = np.abs(x - cen[0])
    ker = 1 / k_len * (k_len - dis)
    ker[ker < 0] = 0
    return np.tril(ker).T + np.tril(ker, k=-1)


def intra_x(M):
    """get_intra_x. .../S2_2_find_by_intra_x.py:116-126."""
    M = np.asarray(M, dtype=np.float64)
    n = M.shape[0]
    k_len = kernal_len(n, .05, 3)
    out = np.zeros(n)
    for c in range(n):
        ker = intra_x_kernal(c, n, k_len)
        out[c] = np.sum(np.multiply(M, ker)) / np.sum(ker)
    return out


# ----------------------------------------------------------------------------------------
# a14/a15  GF_S2: pool -> z-score -> symmetrise -> template scores
# ----------------------------------------------------------------------------------------

X_SIZE = 16                                   # confor/src/S2_to_train_mtx/x_size.py


def _adaptive_pool(B, out):
    """torch.nn.AdaptiveAvgPool2d(out) semantics: cell i covers
    [floor(i*n/out), ceil((i+1)*n/out))."""
    n1, n2 = B.shape
    res = np.zeros((out, out))
    for i in range(out):
        r0, r1 = (i * n1) // out, -((-(i + 1) * n1) // out)
        for j in range(out):
            c0, c1 = (j * n2) // out, -((-(j + 1) * n2) // out)
            res[i, j] = B[r0:r1, c0:c1].mean()
    return res


def pool_mtx(B, c1=None, c2=None):
    """pool_mtx. confor/src/S2_to_train_mtx/preprocess.py:8-29."""
    B = np.asarray(B, dtype=np.float64)
    if (c1 is None or c2 is None or c1 == 0 or c1 >= B.shape[0] - 1
            or c2 == 0 or c2 >= B.shape[1] - 1):
        return _adaptive_pool(B, X_SIZE)
    h = X_SIZE // 2
    top = np.hstack([_adaptive_pool(B[:c1, :c2], h), _adaptive_pool(B[:c1, c2:], h)])
    bot = np.hstack([_adaptive_pool(B[c1:, :c2], h), _adaptive_pool(B[c1:, c2:], h)])
    return np.vstack([top, bot])


def norm_mtx(x):
    """preprocess.py:44-48 (population std)."""
    if np.std(x) > 0:
        return (x - np.mean(x)) / np.std(x)
    return x - np.mean(x)


def symmetry(x, intra=True):
    """preprocess.py:32-41."""
    if intra:
        return (x + np.flipud(np.fliplr(x))) / 2
    s = x.copy()
    s += np.flipud(x)
    s += np.fliplr(x)
    s += np.flipud(np.fliplr(x))
    s /= 4
    return s


def gf_image(B, intra, c1=None, c2=None):
    """pool -> norm -> symmetry; to_train_mtx.py:85-92. Returns the 256-vector."""
    return symmetry(norm_mtx(pool_mtx(B, c1, c2)), intra=intra).reshape(-1)


def normalise_templates(templates):
    """get_n_norm_ave_maps: t / sum|t|. confor/src/S5_sim_to_pat/to_confor.py:12-25.
    `templates` [5, 256] in sorted-key order (change_map_name, :28-32)."""
    t = np.asarray(templates)
    return np.vstack([row / np.sum(np.abs(row)) for row in t])


def template_scores(xs, templates_norm, do_trans=False):
    """to_confor scores. to_confor.py:35-47."""
    xs = np.asarray(xs, dtype=np.float64)
    scores = np.matmul(xs, templates_norm.T)
    if do_trans:
        tidx = np.arange(X_SIZE ** 2).reshape((X_SIZE, X_SIZE)).T.reshape((-1,))
        ts = np.matmul(xs[:, tidx], templates_norm.T)
        change = scores < ts
        scores[change] = ts[change]
    return scores
