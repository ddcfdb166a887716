"""Drop-in mirror of the reference's data model (publish/new_hic_class.py) for the hot path.

Same class / function names, argument order, defaults and misspellings
(`counts_must_decrese`) as the reference, so L3/L4 code written against
`publish/new_hic_class.py` runs unchanged for the paths listed in SURVEY.md section 8(b):

    GenomeIndex            (:19-458)    bin index: chros, chr_seps, chr_lens, bin_size ...
    ScaledGenomicDistance  (:788-1003)  re-exported from hiarch_b200.sgd
    SpsHiCMtx              (:2152-2465) sparse container (host COO / scipy CSR)
    ArrayHiCMtx            (:1940-2149) dense container -- the map lives in HBM (fp32)
    read_sps_file          (:2576-2602), concat_mtxes (:2516-2551), change_suffix (:13-14)

Differences, all deliberate: dense maps are device-resident fp32 torch views (`.dmatrix`);
`.matrix` of an ArrayHiCMtx is a *host fp64 copy made on demand* for callers that index it
with numpy. Every dense operation launches the CUDA kernels behind `hiarch_b200.ops`; there is
no CPU path.
"""
import copy
import statistics
import warnings

import numpy as np
import pandas as pd
import scipy.sparse as sps
import torch

from . import ops
from .sgd import ScaledGenomicDistance  # noqa: F401  (same public name as the reference)


def change_suffix(file_name, changed_suffix):
    """publish/new_hic_class.py:13-14 (names containing '.' are cut at the first dot)."""
    return f"{'/'.join(file_name.split('/')[:-1])}/{file_name.split('/')[-1].split('.')[0]}{changed_suffix}"


# ------------------------------------------------------------------------------------------
class GenomeIndex:
    """publish/new_hic_class.py:19-458 (the parts the hot path uses)."""

    def __init__(self, input_window, corr_index=None, read_corr=True):
        self.is_diplo, self.corr_index = None, None
        if isinstance(input_window, str):
            self.window_df = pd.read_table(input_window, names=['chr', 'start', 'end', 'index'],
                                           dtype={'chr': str, 'start': int, 'end': int, 'index': int},
                                           comment="#")
        elif isinstance(input_window, pd.DataFrame):
            self.window_df = input_window
        elif isinstance(input_window, GenomeIndex):
            self.window_df = input_window.window_df
        else:
            raise ValueError(f'Unrecognized input window type: {type(input_window)}')
        self.window_df.index = self.window_df['index']
        self.window_df = self.window_df.sort_index()
        self.ori_start_idx = None
        self._layouts = {}
        self.get_index_features()

    def __getitem__(self, item):
        return self.window_df[item]

    def copy(self):
        return copy.copy(self)

    def get_index_features(self):
        """:83-105."""
        df = self.window_df
        self.min_idx = df['index'].min()
        self.max_idx = df['index'].max()
        self.all_len = self.max_idx - self.min_idx + 1
        self.chr_seps, self.chr_lens, self.chr_abs_lens = {}, {}, {}
        chrs = df['chr'].values
        idx = df['index'].values
        ends = df['end'].values
        for chro in pd.unique(chrs):
            m = chrs == chro
            self.chr_seps[chro] = [idx[m].min(), idx[m].max()]
            self.chr_lens[chro] = int(m.sum())
            self.chr_abs_lens[chro] = ends[m].max()
        self.chros = [c for c, _ in sorted(self.chr_seps.items(), key=lambda x: x[1][0])]
        bin_sizes = np.array(df['end'].iloc[:10] - df['start'].iloc[:10])
        self.bin_size = statistics.mode(bin_sizes)

    def get_region_startend_idx(self, chros, starts=None, ends=None, is_mtx_idx=True):
        """Whole-chromosome form only (what yield_by_chro needs)."""
        if starts is not None or ends is not None:
            raise NotImplementedError("sub-chromosome regions are not on the hot path")
        s, e = self.chr_seps[chros]
        return s, e

    def seps(self):
        """[(start, end_inclusive)] in chromosome order."""
        return [(int(self.chr_seps[c][0]), int(self.chr_seps[c][1])) for c in self.chros]

    def layout(self, device="cuda"):
        key = str(device)
        if key not in self._layouts:
            self._layouts[key] = ops.GenomeLayout(self.seps(), device)
        return self._layouts[key]

    def sub_index(self, chro):
        """Index of one chromosome re-based to 0 (get_region_mtx, :1413-1417)."""
        s, e = self.chr_seps[chro]
        sub = self.window_df.loc[s:e, :].copy()
        sub.loc[:, 'ori_index'] = sub['index'].copy()
        sub.loc[:, 'index'] = sub['index'] - s
        g = GenomeIndex(sub)
        g.ori_start_idx = s
        return g

    def to_output(self, out_file):
        """4-column .window.bed (:419-421)."""
        self.window_df[['chr', 'start', 'end', 'index']].to_csv(out_file, sep="\t", header=False, index=False)

    # ---- index surgery used by preprocess_mtx / CorrectMap (host, integer, bit-exact) ----------
    def condense(self, condense_fold):
        """:423-458. Returns (new GenomeIndex, index_trans: old index -> new index, int64 array)."""
        df = self.window_df
        idx = df['index'].values
        chrs = df['chr'].values
        new_index = np.zeros(len(df), dtype=np.int64)
        current = 0
        for chro in self.chros:
            lo = self.chr_seps[chro][0]
            m = chrs == chro
            ni = (idx[m] - lo) // condense_fold + current
            new_index[m] = ni
            current = int(ni.max()) + 1
        # groupby('new_index'): chr min / start min / end max per coarse bin. A coarse bin never spans two
        # chromosomes (`current` restarts per chromosome), so the min over its equal names is any of them;
        # done with reduceat on the bins sorted by coarse index (pandas falls back to a per-group Python loop
        # for the string column: 0.13 s of a 0.6 s NormDis step at 6 895 bins).
        order = np.argsort(new_index, kind='stable')
        ni_sorted = new_index[order]
        first = np.flatnonzero(np.r_[True, ni_sorted[1:] != ni_sorted[:-1]])
        new_df = pd.DataFrame({'chr': chrs[order][first],
                               'start': np.minimum.reduceat(df['start'].values[order], first),
                               'end': np.maximum.reduceat(df['end'].values[order], first)},
                              index=pd.Index(ni_sorted[first], name='new_index'))
        new_df['index'] = new_df.index
        trans = np.full(int(idx.max()) + 1, -1, dtype=np.int64)
        trans[idx] = new_index
        return GenomeIndex(new_df.reset_index(drop=True)), trans

    def keep_chros(self, chros, return_old_index=False):
        """:256-278: chromosomes re-packed in the given order; returns the old index per new bin."""
        if len(chros) == 0:
            return (None, None) if return_old_index else None
        parts = [self.window_df[self.window_df['chr'] == c] for c in chros]
        new_df = pd.concat(parts, axis=0).reset_index(drop=True)
        old_index = np.array(new_df['index'])
        new_df = new_df.drop(['index'], axis=1)
        new_df['index'] = np.arange(len(new_df))
        g = GenomeIndex(new_df[['chr', 'start', 'end', 'index']])
        return (g, old_index) if return_old_index else g

    def keep_long_chros(self, min_chro_len=5, return_old_index=False, verbose=True):
        """:306-327: chromosomes with MORE than min_chro_len bins."""
        keep = [c for c in self.chros if self.chr_lens[c] > min_chro_len]
        drop = [c for c in self.chros if self.chr_lens[c] <= min_chro_len]
        if verbose and drop:
            print(f'{" ".join(drop)} is too short.')
        return self.keep_chros(keep, return_old_index)

    def get_sub_gen_index(self, start=None, end=None, idxes=None, reset_index=True):
        """:229-249 for a list of kept indexes."""
        if idxes is None:
            raise NotImplementedError
        sub = self.window_df.loc[idxes, :].copy()
        sub.loc[:, 'ori_index'] = sub['index'].copy()
        sub.reset_index(inplace=True, drop=True)
        sub.loc[:, 'index'] = sub.index
        g = GenomeIndex(sub)
        g.ori_start_idx = int(np.min(idxes)) if len(idxes) else None
        return g


# ------------------------------------------------------------------------------------------
class HiCMtx:
    def __init__(self):
        self.mtx_file = None
        self.mtx_name = None
        self.row_window = None
        self.col_window = None
        self.mtx_type = None
        self.bin_size = None
        self.intra = None
        self.has_neg = False

    def copy(self):
        return copy.copy(self)

    def _init_meta(self, mtx_file, row_window, col_window, mtx_type, intra, has_neg, copy_mtx_paras):
        self.has_neg = None
        if copy_mtx_paras is not None:
            for a in ("mtx_file", "row_window", "col_window", "mtx_type", "intra", "has_neg"):
                setattr(self, a, getattr(copy_mtx_paras, a))
        if row_window is None and col_window is not None:
            self.row_window, self.col_window = col_window, col_window
        elif col_window is None and row_window is not None:
            self.row_window, self.col_window = row_window, row_window
        elif col_window is not None and row_window is not None:
            self.row_window, self.col_window = row_window, col_window
        if self.row_window is not None:
            self.bin_size = self.row_window.bin_size
        if mtx_file is not None:
            self.mtx_file = mtx_file
        if mtx_type is not None:
            self.mtx_type = mtx_type
        if intra is not None:
            self.intra = intra
        if has_neg is not None:
            self.has_neg = has_neg
        elif self.has_neg is None:
            self.get_has_neg()

    # ---- O/E: HiCMtx.obs_d_exp (:1121-1290) -------------------------------------------------
    def obs_d_exp(self, mode='sgd_mean', k=.2, verbose=False, only_intra=False,
                  return_exp_counts=False, counts_must_decrese=True, min_dis_mean=1):
        if mode != 'sgd_mean':
            raise NotImplementedError("only mode='sgd_mean' (the pipeline's mode, S2_oe.py:5) is on the hot path")
        dense = self if isinstance(self, ArrayHiCMtx) else self.to_dense()
        if dense.mtx_type != 'all':
            dense = dense.to_all(inplace=False)
        layout = dense.row_window.layout(dense.dmatrix.device)
        out, state = ops.obs_d_exp(dense.dmatrix, layout, dense.row_window.bin_size, k=k,
                                   counts_must_decrese=counts_must_decrese, only_intra=only_intra)
        new_mtx = ArrayHiCMtx(out, copy_mtx_paras=dense, mtx_type='all', has_neg=False)
        new_mtx._oe_state = state
        new_mtx._oe_src = (dense.dmatrix, layout, k, counts_must_decrese, only_intra)
        if return_exp_counts:
            s, e = layout.seps[-1]                        # the reference returns the last block's vector
            return new_mtx, state.expected[s:e + 1].cpu().numpy()
        return new_mtx

    def oe_float64(self):
        """Not in the reference: the raw O/E map of the `obs_d_exp` call that produced this object
        as a float64 CUDA tensor, divided in double like the reference (1e-5 absolute for entries
        of any size; the fp32 `dmatrix` carries 6e-8 relative, i.e. 1e-5 only up to values of ~50)."""
        src = getattr(self, "_oe_src", None)
        if src is None:
            raise ValueError("oe_float64() is available on the result of obs_d_exp() only")
        dmatrix, layout, k, clamp, only_intra = src
        out, _ = ops.obs_d_exp(dmatrix, layout, self.row_window.bin_size, k=k, counts_must_decrese=clamp,
                               only_intra=only_intra, fp64_out=True)
        return out

    # ---- chromosome selection / reordering (:1492-1506, :1663-1667; what CorrectMap.py calls) ----
    def get_multi_chros_mtx(self, chros):
        new_row_window, kept_r = self.row_window.keep_chros(chros, return_old_index=True)
        new_col_window, kept_c = self.col_window.keep_chros(chros, return_old_index=True)
        if isinstance(self, ArrayHiCMtx):
            d = self.dmatrix
            ir = torch.as_tensor(kept_r, device=d.device, dtype=torch.long)
            ic = torch.as_tensor(kept_c, device=d.device, dtype=torch.long)
            out = ops.alloc_map(len(kept_r), d.device, cols=len(kept_c))
            out.copy_(d[ir][:, ic])
            return ArrayHiCMtx(out, copy_mtx_paras=self, row_window=new_row_window, col_window=new_col_window)
        new_matrix = self.matrix[kept_r, :][:, kept_c]
        return SpsHiCMtx(new_matrix, copy_mtx_paras=self, row_window=new_row_window, col_window=new_col_window)

    def reorder_chros_by_len(self):
        chr_lens = self.row_window.chr_lens
        sorted_chros = [c for c, _ in sorted(chr_lens.items(), key=lambda x: x[1], reverse=True)]
        return self.get_multi_chros_mtx(sorted_chros)

    # ---- block iteration (:1372-1396) -------------------------------------------------------
    def yield_by_chro(self, only_intra=False, triu=False):
        for chro1 in self.row_window.chros:
            for chro2 in self.col_window.chros:
                if triu and self.row_window.chr_seps[chro1][0] > self.col_window.chr_seps[chro2][0]:
                    continue
                if chro1 != chro2 and only_intra:
                    continue
                chr_mtx = self.get_region_mtx(chro1, chro2=chro2)
                chr_mtx.intra = chro1 == chro2
                chr_mtx.mtx_type = self.mtx_type if chro1 == chro2 else 'all'
                chr_mtx.has_neg = self.has_neg
                if triu and chro1 == chro2:
                    chr_mtx.to_triu(inplace=True)
                if only_intra:
                    yield chro1, chr_mtx
                else:
                    yield chro1, chro2, chr_mtx


# ------------------------------------------------------------------------------------------
class ArrayHiCMtx(HiCMtx):
    """Dense map in HBM. `matrix` may be a CUDA fp32 tensor (kept as is, must be a view with a
    leading dimension that is a multiple of 4 for the vectorised kernels) or a numpy array
    (uploaded)."""

    def __init__(self, matrix, mtx_file=None, row_window=None, col_window=None,
                 mtx_type=None, intra=None, has_neg=None, copy_mtx_paras=None):
        super().__init__()
        if isinstance(matrix, torch.Tensor):
            self.dmatrix = matrix
        else:
            arr = np.asarray(matrix)
            d = ops.alloc_map(arr.shape[0], cols=arr.shape[1])
            d.copy_(torch.from_numpy(np.ascontiguousarray(arr)).to(device="cuda", dtype=torch.float32))
            self.dmatrix = d
        self.shape = tuple(self.dmatrix.shape)
        self._init_meta(mtx_file, row_window, col_window, mtx_type, intra, has_neg, copy_mtx_paras)

    @property
    def matrix(self):
        """Host fp64 copy (a D2H transfer) for numpy-style consumers."""
        return self.dmatrix.cpu().numpy().astype(np.float64)

    @matrix.setter
    def matrix(self, value):
        if isinstance(value, torch.Tensor):
            self.dmatrix = value
        else:
            self.dmatrix = ArrayHiCMtx(value).dmatrix
        self.shape = tuple(self.dmatrix.shape)

    def get_has_neg(self):
        self.has_neg = bool((self.dmatrix.min() < 0).item())

    def to_dense(self):
        new = copy.copy(self)
        d = ops.alloc_map(self.shape[0], self.dmatrix.device, cols=self.shape[1])
        d.copy_(self.dmatrix)
        new.dmatrix = d
        return new

    def to_sps(self):
        """csr_matrix(dense): exact zeros vanish (:1985-1986)."""
        idx = self.dmatrix.nonzero(as_tuple=False)
        vals = self.dmatrix[idx[:, 0], idx[:, 1]]
        return SpsHiCMtx((vals.cpu().numpy().astype(np.float64),
                          (idx[:, 0].cpu().numpy(), idx[:, 1].cpu().numpy())),
                         copy_mtx_paras=self, shape=self.shape)

    def get_mtx_type(self):
        d = self.dmatrix
        if d.shape[0] != d.shape[1] or float(d.sum()) <= 0:
            self.mtx_type = 'all'
            return
        tl, tu = float(torch.tril(d, -1).sum()), float(torch.triu(d, 1).sum())
        diag = float(torch.diagonal(d).sum())
        if (diag != 0 and tl == 0 and tu == 0) or (tl == 0 and tu != 0):
            self.mtx_type = 'triu'
        elif tl != 0 and tu == 0:
            self.mtx_type = 'tril'
        else:
            self.mtx_type = 'all'

    def to_all(self, inplace=False):
        """:1991-2008."""
        if self.mtx_type is None:
            self.get_mtx_type()
        d = self.dmatrix
        if self.mtx_type == 'all':
            new = d
        elif self.mtx_type == 'triu':
            new = ops.alloc_map(d.shape[0], d.device, cols=d.shape[1])
            new.copy_(torch.triu(d, 0).t() + torch.triu(d, 1))
        elif self.mtx_type == 'tril':
            new = ops.alloc_map(d.shape[0], d.device, cols=d.shape[1])
            new.copy_(torch.tril(d, 0).t() + torch.tril(d, 1))
        else:
            raise ValueError(f'Unrecognized mtx type: {self.mtx_type}')
        if inplace:
            self.mtx_type, self.dmatrix = 'all', new
            return None
        return ArrayHiCMtx(new, copy_mtx_paras=self, mtx_type='all')

    def to_triu(self, k=0, inplace=False):
        if self.mtx_type is None:
            self.get_mtx_type()
        src = self.dmatrix.t() if self.mtx_type == 'tril' else self.dmatrix
        new = ops.alloc_map(src.shape[0], src.device, cols=src.shape[1])
        new.copy_(torch.triu(src, k))
        if inplace:
            self.mtx_type, self.dmatrix = 'triu', new
            return None
        return ArrayHiCMtx(new, copy_mtx_paras=self, mtx_type='triu')

    def log(self, inplace=False, verbose=True):
        """:2076-2094: x>0 -> ln(x+1); x<=0 -> min over positives of ln(x+1)."""
        if self.has_neg:
            warnings.warn('Contain negative values, cannot do log.')
            new = self.to_dense().dmatrix
        else:
            new = ops.log_map(self.dmatrix)
        if inplace:
            self.dmatrix = new
            return None
        return ArrayHiCMtx(new, copy_mtx_paras=self)

    def get_region_mtx(self, chro, start=None, end=None, chro2=None, start2=None, end2=None,
                       col_is_inter=False, col_is_all=False, inplace=False):
        """Whole-chromosome blocks as VIEWS of the genome-wide map (:1398-1490)."""
        if col_is_inter or col_is_all or start is not None or end is not None or inplace:
            raise NotImplementedError("only whole chromosome-pair blocks are on the hot path")
        chro2 = chro if chro2 is None else chro2
        s1, e1 = self.row_window.chr_seps[chro]
        s2, e2 = self.col_window.chr_seps[chro2]
        blk = self.dmatrix[s1:e1 + 1, s2:e2 + 1]
        rw, cw = self.row_window.sub_index(chro), self.col_window.sub_index(chro2)
        return ArrayHiCMtx(blk, mtx_file=self.mtx_file, row_window=rw, col_window=cw,
                           mtx_type='all' if chro != chro2 else self.mtx_type,
                           intra=chro == chro2, has_neg=self.has_neg)

    def to_output(self, out_file, out_window=False, out_int=False, float_format="%.4f",
                  compression='infer', only_row_window=True):
        """:2029-2047 (dense text)."""
        if out_window:
            self.row_window.to_output(change_suffix(out_file, '.window.bed'))
        m = pd.DataFrame(self.matrix)
        if out_int:
            m, float_format = m.astype(int), '%d'
        m.to_csv(out_file, header=False, index=False, sep="\t", float_format=float_format,
                 compression=compression)


# ------------------------------------------------------------------------------------------
class SpsHiCMtx(HiCMtx):
    """Sparse container on the host (scipy CSR, duplicates summed like the reference's
    coo->csr); densification is the GPU scatter."""

    def __init__(self, sps_mtx, mtx_file=None, row_window=None, col_window=None,
                 mtx_type=None, intra=None, has_neg=None, copy_mtx_paras=None, shape=None):
        super().__init__()
        if isinstance(sps_mtx, tuple):                  # (vals, (rows, cols)) like scipy's coo_matrix
            vals, (rows, cols) = sps_mtx
            fast = None
            if shape is not None:                       # row-major entries (torch.nonzero order): no sort
                fast = _csr_from_sorted_coo(np.asarray(rows), np.asarray(cols), np.asarray(vals), tuple(shape))
            sps_mtx = fast if fast is not None else sps.coo_matrix(sps_mtx, shape=shape)
        self.matrix = sps_mtx if isinstance(sps_mtx, sps.csr_matrix) else sps_mtx.tocsr()
        self.shape = self.matrix.shape
        self._init_meta(mtx_file, row_window, col_window, mtx_type, intra, has_neg, copy_mtx_paras)

    def get_has_neg(self):
        self.has_neg = bool(self.matrix.data.size and np.min(self.matrix.data) < 0)

    def get_mtx_type(self):
        """:1294-1322."""
        m = self.matrix
        if m.sum() <= 0:
            warnings.warn('Sum of input matrix less than 0.')
        if m.shape[0] != m.shape[1]:
            self.mtx_type = 'all'
            return
        diag = m.diagonal().sum()
        tl, tu = sps.tril(m, k=-1).sum(), sps.triu(m, k=1).sum()
        if (diag != 0 and tl == 0 and tu == 0) or (tl == 0 and tu != 0):
            self.mtx_type = 'triu'
        elif tl != 0 and tu == 0:
            self.mtx_type = 'tril'
        else:
            self.mtx_type = 'all'

    def to_sps(self):
        return copy.deepcopy(self)

    def to_all(self, inplace=False):
        """:2229-2246."""
        if self.mtx_type is None:
            self.get_mtx_type()
        if self.mtx_type == 'all':
            new = self.matrix
        elif self.mtx_type == 'triu':
            new = sps.triu(self.matrix, k=0).T + sps.triu(self.matrix, k=1)
        elif self.mtx_type == 'tril':
            new = sps.tril(self.matrix, k=0).T + sps.tril(self.matrix, k=1)
        else:
            raise ValueError(f'Unrecognized mtx type: {self.mtx_type}')
        if inplace:
            self.mtx_type, self.matrix = 'all', new.tocsr()
            return None
        return SpsHiCMtx(new, copy_mtx_paras=self, mtx_type='all')

    def to_triu(self, k=0, inplace=False):
        if self.mtx_type is None:
            self.get_mtx_type()
        src = self.matrix.T if self.mtx_type == 'tril' else self.matrix
        new = sps.triu(src, k=k)
        if inplace:
            self.mtx_type, self.matrix = 'triu', new.tocsr()
            return None
        return SpsHiCMtx(new, copy_mtx_paras=self, mtx_type='triu')

    def device_coo(self, device="cuda"):
        coo = self.matrix.tocoo()
        return (torch.from_numpy(coo.row.astype(np.int32)).to(device),
                torch.from_numpy(coo.col.astype(np.int32)).to(device),
                torch.from_numpy(coo.data.astype(np.float32)).to(device))

    def to_dense(self, zero_fill='min', block_min=False, device="cuda"):
        """:2209-2224 via the scatter kernel. Absent cells: 0, or min(data) when has_neg
        (`block_min=True`: the min of each chromosome-pair block, which is what densifying the
        blocks of yield_by_chro(triu=True) one by one gives, main_local.py:112-115)."""
        if zero_fill != 'min':
            raise NotImplementedError("only zero_fill='min' is used by the pipeline")
        if self.mtx_type is None:
            self.get_mtx_type()
        n = self.shape[0]
        r, c, v = self.device_coo(device)
        layout = self.row_window.layout(device) if self.row_window is not None else None
        fill = ops.FILL_ZERO
        if self.has_neg:
            fill = ops.FILL_BLOCK_MIN if block_min else ops.FILL_MIN
        if self.mtx_type == 'triu':
            d = ops.scatter_coo_sym(r, c, v, n, ops.MTX_TRIU, fill, layout)
            # the scatter mirrors; a 'triu' dense map keeps the upper triangle only
            out = ArrayHiCMtx(d, copy_mtx_paras=self, mtx_type='all')
            out._from_triu = True
            return out
        d = ops.scatter_coo_sym(r, c, v, n, ops.MTX_ALL, fill, layout)
        return ArrayHiCMtx(d, copy_mtx_paras=self)

    def to_output(self, out_file, out_window=True, out_near_num=None, out_triu=True, out_int=False,
                  only_out_intra=False, only_row_window=True, fmt=None):
        """:2267-2308: upper-triangular COO text, default '%d\\t%d\\t%.2f'."""
        if out_near_num is not None or only_out_intra:
            raise NotImplementedError
        if out_triu:
            self.to_triu(inplace=True)
        m = self.matrix.tocoo()
        if out_int or fmt is None:
            from .mtx_io import write_mtx
            write_mtx(out_file, m.row, m.col, m.data, out_int=out_int)      # native writer, byte-identical
        else:
            np.savetxt(out_file, np.vstack([m.row, m.col, m.data]).T, fmt=fmt)
        if out_window:
            self.row_window.to_output(change_suffix(out_file, '.window.bed'))
        print("Window file and sparse matrix file have been outputted.")


# ------------------------------------------------------------------------------------------
def concat_mtxes(mtxes, row_window=None, col_window=None, **other_mtx_kwargs):
    """:2516-2551 for dense device blocks that are views of one parent map: nothing to copy.
    Kept for API compatibility; returns the parent as an ArrayHiCMtx."""
    parent = other_mtx_kwargs.get('copy_mtx_paras')
    if parent is None or not isinstance(parent, ArrayHiCMtx):
        raise NotImplementedError("concat_mtxes needs copy_mtx_paras= the dense parent map")
    return parent


def _csr_from_sorted_coo(rows, cols, vals, shape):
    """CSR straight from entries that are already in canonical order (row-major, columns strictly
    increasing inside a row, no duplicates) -- every file the pipeline writes (`to_output` walks a
    CSR matrix). The same arrays scipy's coo -> csr produces (:2596-2600), without its sort
    (0.7 s per 12 M entries). Returns None when the order is anything else."""
    n = len(rows)
    if n == 0 or shape[0] <= 0 or shape[1] <= 0:
        return None
    if rows.min() < 0 or cols.min() < 0 or rows.max() >= shape[0] or cols.max() >= shape[1]:
        return None                                 # let scipy raise its own error
    key = rows.astype(np.int64) * np.int64(shape[1]) + cols
    if n > 1 and not np.all(key[1:] > key[:-1]):
        return None
    idx_dtype = np.int32 if max(shape[0], shape[1], n) < 2 ** 31 - 1 else np.int64
    indptr = np.zeros(shape[0] + 1, dtype=idx_dtype)
    np.cumsum(np.bincount(rows, minlength=shape[0]), out=indptr[1:])
    m = sps.csr_matrix((np.asarray(vals, dtype=np.float64), cols.astype(idx_dtype), indptr), shape=shape)
    m.has_sorted_indices = True
    m.has_canonical_format = True
    return m


def read_sps_file(sps_file, row_window_file=None, col_window_file=None,
                  mtx_type=None, intra=None, has_neg=None):
    """:2576-2602: tab separated (row, col, value) -> SpsHiCMtx."""
    from .mtx_io import read_mtx
    r_, c_, v_ = read_mtx(sps_file)               # native parser: pd.read_table(...) + dropna
    df = {'row': r_, 'col': c_, 'value': v_}
    if row_window_file is None and col_window_file is None:
        row_window, col_window = None, None
        shape = (int(np.max(df['row'])) + 1, int(np.max(df['col'])) + 1)
    elif row_window_file is not None and col_window_file is None:
        row_window, col_window = GenomeIndex(row_window_file), None
        shape = (row_window.max_idx + 1, row_window.max_idx + 1)
    elif row_window_file is None:
        row_window, col_window = None, GenomeIndex(col_window_file)
        shape = (col_window.max_idx + 1, col_window.max_idx + 1)
    else:
        row_window, col_window = GenomeIndex(row_window_file), GenomeIndex(col_window_file)
        shape = (row_window.max_idx + 1, col_window.max_idx + 1)
    m = _csr_from_sorted_coo(df['row'], df['col'], df['value'], shape)
    if m is None:                                   # any other order / duplicates: scipy's coo -> csr
        m = sps.coo_matrix((df['value'], (df['row'], df['col'])), shape=shape)
    out = SpsHiCMtx(m, mtx_file=sps_file, row_window=row_window, col_window=col_window,
                    mtx_type=mtx_type, intra=intra, has_neg=has_neg)
    out.mtx_name = sps_file.split('/')[-1].split('.sps.mtx')[0]
    return out


def read_matrix(mtx_file, index_file=None, thread=1):
    """publish/iutils/read_matrix.py:21-30."""
    if not mtx_file.endswith('.mtx'):
        raise ValueError('Unrecognized input matrix file.')
    m = read_sps_file(mtx_file, index_file)
    m.mtx_name = mtx_file
    m.mtx_type = 'triu'
    if 'ode' in mtx_file.split('/')[-1].split('.')[-2]:
        m.has_neg = True
    return m
