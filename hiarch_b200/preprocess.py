"""f1 / f3: the steps around the hot path that decide its shapes.

preprocess_mtx   publish/oe_map/S1_preprocess.py:48-83 (tid_off_zero_chros, filt_chr_by_name,
                 rough_concat -> SpsHiCMtx.condense, tid_off_low_bins, the non-zero-ratio gate and
                 its condense search, keep_long_chros)
correct_map      publish/HiArch/CorrectMap.py:30-47 (drop listed chromosomes, reorder by length)

Design: every one of these steps is a row/column *index map* (drop / fold / permute). The map
is composed on the host -- integer logic, bit-exact -- and applied by ONE kernel that scatters the
original fine COO straight into the coarse dense map (`hia_scatter_coo_remap`); the
data-dependent decisions read tiny statistics back from the device (`hia_coo_block_nnz`,
`hia_bin_stats`). The fine matrix is never densified and no intermediate sparse matrix is built.
"""
import re

import numpy as np
import torch

from . import _lib, ops
from .new_hic_class import ArrayHiCMtx, GenomeIndex, SpsHiCMtx, read_sps_file

min_nonzero_ratio = .95
max_searching_n_con = 10
min_len_for_kept_chro = 20
tid_off_low_bin_ratio = 5
max_bin_num_for_chros = 600
log_base = 2


class _Remapped:
    """A fine COO on the device + the composed index map + the current GenomeIndex."""

    def __init__(self, sps_mtx, device="cuda"):
        self.lib = _lib.load()
        self.device = device
        if sps_mtx.mtx_type is None:
            sps_mtx.get_mtx_type()
        self.mtx_type = ops.MTX_TRIU if sps_mtx.mtx_type == 'triu' else ops.MTX_ALL
        if sps_mtx.mtx_type == 'tril':
            raise NotImplementedError("'tril' inputs")
        self.rows, self.cols, self.vals = sps_mtx.device_coo(device)
        self.n_fine = sps_mtx.shape[0]
        self.index = sps_mtx.row_window
        self.map = np.arange(self.n_fine, dtype=np.int64)          # fine bin -> current bin (-1 dropped)
        self.src = sps_mtx

    # ---- index-map composition ------------------------------------------------------------
    def _apply(self, trans):
        """trans: current index -> new index (or -1)."""
        m = self.map
        out = np.full_like(m, -1)
        ok = m >= 0
        out[ok] = trans[m[ok]]
        self.map = out

    def keep_chros(self, chros):
        new_index, old_index = self.index.keep_chros(chros, return_old_index=True)
        if new_index is None:
            return False
        trans = np.full(self.index.max_idx + 1, -1, dtype=np.int64)
        trans[old_index] = np.arange(len(old_index))
        self._apply(trans)
        self.index = new_index
        return True

    def keep_long_chros(self, min_chro_len):
        """new_hic_class.py:1599-1630 (row-window call is the verbose one)."""
        short = [c for c in self.index.chros if self.index.chr_lens[c] <= min_chro_len]
        if short:
            print(f'{" ".join(short)} is too short.')
        return self.keep_chros([c for c in self.index.chros if self.index.chr_lens[c] > min_chro_len])

    def condense(self, fold):
        new_index, trans = self.index.condense(fold)
        self._apply(trans)
        self.index = new_index

    def keep_bins(self, idxes):
        idxes = np.asarray(idxes, dtype=np.int64)
        trans = np.full(self.index.max_idx + 1, -1, dtype=np.int64)
        trans[idxes] = np.arange(len(idxes))
        new_index = self.index.get_sub_gen_index(idxes=idxes)
        self._apply(trans)
        self.index = new_index

    # ---- device work ------------------------------------------------------------------------
    def dense(self):
        n = int(self.index.max_idx) + 1
        out = ops.alloc_map(n, self.device)
        d_map = torch.from_numpy(self.map.astype(np.int32)).to(self.device)
        st = self.lib.hia_scatter_coo_remap(self.rows.data_ptr(), self.cols.data_ptr(), self.vals.data_ptr(),
                                            self.rows.numel(), self.n_fine, d_map.data_ptr(), n,
                                            out.data_ptr(), out.stride(0), self.mtx_type, ops._stream())
        _lib.check(st, "hia_scatter_coo_remap")
        return out

    def bin_sums(self):
        """fp64 row sums of the current (remapped, symmetrised) matrix, straight from the COO."""
        n = int(self.index.max_idx) + 1
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        d_map = torch.from_numpy(self.map.astype(np.int32)).to(self.device)
        st = self.lib.hia_coo_bin_sums(self.rows.data_ptr(), self.cols.data_ptr(), self.vals.data_ptr(),
                                       self.rows.numel(), self.n_fine, d_map.data_ptr(), n, self.mtx_type,
                                       out.data_ptr(), ops._stream())
        _lib.check(st, "hia_coo_bin_sums")
        return out.cpu().numpy()

    def bin_stats(self, dense):
        n = dense.shape[0]
        rs = torch.empty(n, dtype=torch.float64, device=self.device)
        rn = torch.empty(n, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.hia_bin_stats(dense.data_ptr(), n, dense.stride(0), rs.data_ptr(), rn.data_ptr(),
                                          ops._stream()), "hia_bin_stats")
        return rs.cpu().numpy(), rn.cpu().numpy()

    def intra_nnz_fine(self):
        """non-zero stored entries of every intra block of the symmetrised fine matrix."""
        layout = self.index.layout(self.device)
        cnt = torch.empty(layout.n_chr * layout.n_chr, dtype=torch.int64, device=self.device)
        st = self.lib.hia_coo_block_nnz(self.rows.data_ptr(), self.cols.data_ptr(), self.vals.data_ptr(),
                                        self.rows.numel(), self.n_fine, self.mtx_type,
                                        layout.bin_chr.data_ptr(), layout.n_chr, cnt.data_ptr(), ops._stream())
        _lib.check(st, "hia_coo_block_nnz")
        c = cnt.cpu().numpy().reshape(layout.n_chr, layout.n_chr)
        return {chro: int(c[i, i]) for i, chro in enumerate(self.index.chros)}

    # ---- the reference's steps ---------------------------------------------------------------
    def tid_off_low_bins(self, ratio):
        """new_hic_class.py:1658-1702: keep bins whose row+column sum exceeds the `ratio`-th
        percentile of all bins. Returns the dense map AFTER the trim and its (row_sum, row_nnz)."""
        row_sum = self.bin_sums()
        bin_sum = row_sum + row_sum                               # sum(axis=0) + sum(axis=1), symmetric
        remain = np.nonzero(bin_sum > np.percentile(bin_sum, ratio))[0]
        print(f'### Tid off bins. Remain: {len(remain)}/{len(bin_sum)}')
        self.keep_bins(self.index.window_df['index'].values[remain])
        d = self.dense()
        return d, self.bin_stats(d)


def preprocess_mtx(sps_mtx, do_filt_chr_by_name=False):
    """publish/oe_map/S1_preprocess.py:48-83. Returns a dense 'all' ArrayHiCMtx on the device (the
    reference returns the equivalent SpsHiCMtx) or None when no coarse-graining reaches 95 %
    non-zero cells."""
    if not isinstance(sps_mtx, SpsHiCMtx):
        raise TypeError("preprocess_mtx expects the SpsHiCMtx read from a .mtx file")
    w = _Remapped(sps_mtx)
    # tid_off_zero_chros (new_hic_class.py:1631-1656): drop chromosomes whose intra block is > 99 % zero
    nnz = w.intra_nnz_fine()
    kept, dropped = [], []
    for chro in w.index.chros:
        ratio = nnz[chro] / float(w.index.chr_lens[chro] ** 2)
        (kept if 1 - ratio <= .99 else dropped).append(chro)
    if dropped:
        print(f'{" ".join(dropped)} contains too much zero.')
    if not w.keep_chros(kept):
        return None
    if do_filt_chr_by_name:                                       # S1_preprocess.py:5-14
        lens = w.index.chr_lens
        longest = max(lens, key=lens.get)
        prefix = re.search(r'([A-Z]*)\d', longest).group(1)
        w.keep_chros([c for c in w.index.chros if c.startswith(prefix)])
    # rough_concat (S1_preprocess.py:17-26)
    max_len = np.max(list(w.index.chr_lens.values()))
    times = np.log(max_len / max_bin_num_for_chros) / np.log(log_base)
    times = 0 if times < 0 else times
    times = np.power(log_base, int(times))
    print(f'# Will concat raw max len ({int(max_len)}) {times} times')
    if times > 1:
        w.condense(int(times))
    dense, (row_sum, row_nnz) = w.tid_off_low_bins(tid_off_low_bin_ratio)
    n = dense.shape[0]
    nonzero_ratio = float(row_nnz.sum()) / float(n * n) if n > 0 else 0
    print(f'Non-zero ratio is {nonzero_ratio}.')
    if nonzero_ratio > min_nonzero_ratio:
        if not w.keep_long_chros(min_len_for_kept_chro):
            return None
        return ArrayHiCMtx(w.dense(), mtx_file=sps_mtx.mtx_file, row_window=w.index, mtx_type='all', has_neg=False)
    print(f'### Non-zero ratio is smaller than {min_nonzero_ratio}. Will try to concat.')
    t = np.sqrt(min_nonzero_ratio / nonzero_ratio)
    print(f'Estimated concat time is {int(t) + 1} ({t}).')
    est = int(t) + 1
    base_map, base_index = w.map.copy(), w.index
    new_ratio, n_con = None, None
    for n_con in range(est, est + max_searching_n_con):
        w.map, w.index = base_map.copy(), base_index
        w.condense(n_con)
        dense, (row_sum, row_nnz) = w.tid_off_low_bins(tid_off_low_bin_ratio)
        if not w.keep_long_chros(min_len_for_kept_chro):
            return None
        dense = w.dense()
        _, row_nnz = w.bin_stats(dense)
        n = dense.shape[0]
        new_ratio = float(row_nnz.sum()) / float(n * n)
        if new_ratio > min_nonzero_ratio:
            print(f'Found mtx with nonzero {new_ratio} > {min_nonzero_ratio} with condense fold {n_con}')
            return ArrayHiCMtx(dense, mtx_file=sps_mtx.mtx_file, row_window=w.index, mtx_type='all', has_neg=False)
    print(f'### Warning: Can not find mtx with nonzero > {min_nonzero_ratio}. '
          f'Max nonzero_ratio is {new_ratio} with {n_con}')
    return None


def correct_map(sps_file, index_file, output, ab_chrs=None, fig_output=None, do_reorder_chrs=True):
    """publish/HiArch/CorrectMap.py:30-47: remove the listed chromosomes (1-based, negative counts
    from the end), reorder the rest by length (descending, stable), rewrite `.clean_de_ode.mtx` and
    `.clean_window.bed`. Pure index permutation of the sparse entries + text I/O: host side."""
    m = read_sps_file(sps_file, index_file)
    m.mtx_name = sps_file
    m.mtx_type = 'triu'
    gi = m.row_window
    kept = list(gi.chros)
    if ab_chrs is not None:
        raw = list(gi.chros)
        for idx in ab_chrs:
            if idx > 0:
                idx -= 1
            kept.remove(raw[idx])
    new_index, old_index = gi.keep_chros(kept, return_old_index=True)
    if do_reorder_chrs:
        order = [c for c, _ in sorted(new_index.chr_lens.items(), key=lambda x: x[1], reverse=True)]
        new_index2, old2 = new_index.keep_chros(order, return_old_index=True)
        old_index = old_index[old2]
        new_index = new_index2
    trans = np.full(gi.max_idx + 1, -1, dtype=np.int64)
    trans[old_index] = np.arange(len(old_index))
    coo = m.matrix.tocoo()                                        # stored upper triangle
    keep = coo.col >= coo.row
    r, c, v = trans[coo.row[keep]], trans[coo.col[keep]], coo.data[keep]
    ok = (r >= 0) & (c >= 0) & (v != 0)                            # scipy's `+` in to_all drops stored zeros
    r, c, v = r[ok], c[ok], v[ok]
    lo, hi = np.minimum(r, c), np.maximum(r, c)                   # to_all -> permute -> to_triu
    order = np.lexsort((hi, lo))
    arr = np.vstack([lo[order], hi[order], v[order]]).T
    from .mtx_io import write_mtx
    write_mtx(f'{output}.clean_de_ode.mtx', arr[:, 0], arr[:, 1], arr[:, 2])
    new_index.to_output(f'{output}.clean_window.bed')
    import scipy.sparse as sps
    n = len(old_index)
    return SpsHiCMtx(sps.coo_matrix((v, (lo, hi)), shape=(n, n)), mtx_file=sps_file, row_window=new_index,
                     mtx_type='triu', has_neg=m.has_neg)
