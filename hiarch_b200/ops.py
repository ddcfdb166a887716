"""Thin Python wrappers over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Every function here launches hand-written sm_100a kernels through libhiarch_b200.so on the
current torch stream; torch only owns the memory. No op has a CPU or torch fallback.
"""
import collections
import ctypes
import os
import threading

import numpy as np
import torch

from . import _lib
from .sgd import oe_tables

MTX_TRIU, MTX_ALL, MTX_TRIU_UPPER = 0, 1, 2      # TRIU_UPPER: no mirror pass, lower triangle unwritten
FILL_ZERO, FILL_MIN, FILL_BLOCK_MIN = 0, 1, 2
OE_LOG, OE_ONLY_INTRA, OE_FROM_UPPER = 1, 2, 4
SIM_COSINE_DISTANCE, SIM_PEARSON, SIM_CORRELATION_DISTANCE = 0, 1, 2
PREC_3XTF32, PREC_FP64, PREC_3XF16 = 0, 1, 2
# Default similarity precision: the fp16 hi/lo split (same 22-bit operands as 3xTF32, half the
# tensor time; header of csrc/hia_syrk.cu). 3XTF32 and FP64 stay selectable per call.
PREC_DEFAULT = PREC_3XF16
SPLIT_PRECS = (PREC_3XTF32, PREC_3XF16)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.HiaError("hiarch_b200 ops need CUDA tensors (there is no CPU fallback)")


def padded_ld(n):
    """Leading dimension used for dense maps: multiple of 4 (128-bit loads, TMA strides)."""
    return (n + 3) // 4 * 4


def alloc_map(n, device="cuda", cols=None):
    """Uninitialised n x cols fp32 map as a view into an [n, ld] buffer (ld % 4 == 0)."""
    cols = n if cols is None else cols
    buf = torch.empty((n, padded_ld(cols)), dtype=torch.float32, device=device)
    return buf[:, :cols]


class GenomeLayout:
    """Chromosome layout of a map on the device: chr_start[C+1] (int32), bin_chr[N] (int16).

    Built from GenomeIndex.chr_seps-style inclusive (start, end) pairs
    (publish/new_hic_class.py:92-102). Also caches the O/E pooling tables per bin size."""

    def __init__(self, seps, device="cuda"):
        seps = [(int(s), int(e)) for s, e in seps]
        starts = [s for s, _ in seps]
        assert starts == sorted(starts), "chromosomes must be in index order"
        for (s0, e0), (s1, _) in zip(seps[:-1], seps[1:]):
            assert e0 + 1 == s1, "chromosomes must tile the index range contiguously"
        assert seps[0][0] == 0
        self.seps = seps
        self.n = seps[-1][1] + 1
        self.n_chr = len(seps)
        self.device = device
        self.chr_start_host = np.array(starts + [self.n], dtype=np.int32)
        self.chr_start = torch.from_numpy(self.chr_start_host).to(device)
        bin_chr = np.concatenate([np.full(e - s + 1, c, dtype=np.int16)
                                  for c, (s, e) in enumerate(seps)])
        self.bin_chr = torch.from_numpy(bin_chr).to(device)
        self._tables = {}

    def oe_tables(self, bin_size, k=.2):
        key = (int(bin_size), float(k))
        if key not in self._tables:
            pools, uses = [], []
            for (s, e) in self.seps:
                p, u = oe_tables(bin_size, e - s + 1, k)
                pools.append(p)
                uses.append(u)
            self._tables[key] = (torch.from_numpy(np.concatenate(pools)).to(self.device),
                                 torch.from_numpy(np.concatenate(uses)).to(self.device))
        return self._tables[key]


# ------------------------------------------------------------------------------- a1/a2 scatter
_scatter_ws = collections.OrderedDict()     # (device, n, n_chr, stream) -> workspace, bounded LRU
_scatter_ws_lock = threading.Lock()
_SCATTER_WS_MAX = 64


def scatter_coo_sym(rows, cols, vals, n, mtx_type=MTX_TRIU, fill_mode=FILL_ZERO, layout=None,
                    out=None, oe_state=None):
    """COO (int32 rows, int32 cols, fp32 vals; CUDA) -> dense symmetric fp32 [n, n] view.
    Replaces read_sps_file/to_all/to_dense (publish/new_hic_class.py:2576-2602, :2229-2246,
    :2209-2224). oe_state (an OEState of `layout`, zero fill only): the inter-chromosome statistics
    of the O/E expected pass are taken from the row segments while they are in shared memory
    (`oe_state.inter_valid` then tells obs_d_exp to skip its pass over the inter part of the map)."""
    lib = _lib.load()
    _need_cuda(rows, cols, vals)
    assert rows.dtype == torch.int32 and cols.dtype == torch.int32 and vals.dtype == torch.float32
    rows, cols, vals = rows.contiguous(), cols.contiguous(), vals.contiguous()
    if out is None:
        out = alloc_map(n, rows.device)
    n_chr = 0
    chr_start = None
    if fill_mode != FILL_ZERO:
        if layout is None:
            layout = GenomeLayout([(0, n - 1)], rows.device)
        n_chr, chr_start = layout.n_chr, layout.chr_start
    key = (str(rows.device), n, n_chr, _stream())               # one scratch per stream: calls may overlap
    with _scatter_ws_lock:                      # HotPathPool calls this from several host threads
        ws = _scatter_ws.get(key)
        if ws is None:
            ws = torch.empty(lib.hia_scatter_workspace_bytes(n, n_chr), dtype=torch.uint8, device=rows.device)
            _scatter_ws[key] = ws
            while len(_scatter_ws) > _SCATTER_WS_MAX:
                _scatter_ws.popitem(last=False)
        else:
            _scatter_ws.move_to_end(key)
    if oe_state is not None and fill_mode == FILL_ZERO and layout is not None and layout.n_chr <= 64:
        st = lib.hia_scatter_coo_sym_stats(_ptr(rows), _ptr(cols), _ptr(vals), rows.numel(), n, _ptr(out),
                                           out.stride(0), mtx_type, _ptr(layout.chr_start), layout.n_chr,
                                           _ptr(layout.bin_chr), _ptr(oe_state.inter_sum), _ptr(oe_state.inter_nnz),
                                           _ptr(oe_state.inter_minpos), _ptr(oe_state.inter_valid), _ptr(ws), _stream())
        _lib.check(st, "hia_scatter_coo_sym_stats")
        return out
    st = lib.hia_scatter_coo_sym(_ptr(rows), _ptr(cols), _ptr(vals), rows.numel(), n, _ptr(out),
                                 out.stride(0), mtx_type, fill_mode, _ptr(chr_start), n_chr,
                                 _ptr(ws), _stream())
    _lib.check(st, "hia_scatter_coo_sym")
    return out


def scatter_csr_sym(row_ptr, cols, vals, n, mtx_type=MTX_TRIU, out=None, layout=None, oe_state=None):
    """CSR (int64 row_ptr [n + 1], int32 or uint16 cols, fp32 vals; CUDA) -> dense fp32 [n, n] view, zero
    fill: the scatter for matrices that arrive as CSR (scipy's layout of SpsHiCMtx.matrix), 8 or 6
    bytes per entry instead of the 12 of a COO triple."""
    lib = _lib.load()
    _need_cuda(row_ptr, cols, vals)
    assert row_ptr.dtype == torch.int64 and row_ptr.numel() == n + 1 and vals.dtype == torch.float32
    assert cols.dtype in (torch.int32, torch.uint16, torch.int16)
    if out is None:
        out = alloc_map(n, vals.device)
    if oe_state is not None and layout is not None and layout.n_chr <= 64:      # + the inter statistics of the O/E pass
        st = lib.hia_scatter_csr_sym_stats(_ptr(row_ptr), _ptr(cols), cols.element_size(), _ptr(vals), n, _ptr(out),
                                           out.stride(0), mtx_type, _ptr(layout.chr_start), layout.n_chr,
                                           _ptr(layout.bin_chr), _ptr(oe_state.inter_sum), _ptr(oe_state.inter_nnz),
                                           _ptr(oe_state.inter_minpos), _ptr(oe_state.inter_valid), _stream())
        _lib.check(st, "hia_scatter_csr_sym_stats")
        return out
    st = lib.hia_scatter_csr_sym(_ptr(row_ptr), _ptr(cols), cols.element_size(), _ptr(vals), n, _ptr(out),
                                 out.stride(0), mtx_type, _stream())
    _lib.check(st, "hia_scatter_csr_sym")
    return out


def coo_to_csr_host(rows, cols, vals, n, cols16=None):
    """Row-sorted host COO (numpy / CPU tensors) -> (row_ptr int64 [n + 1], cols int32 | uint16, vals fp32)
    pinned CPU tensors. cols16 (default: n <= 65 536): 16-bit column indices."""
    rows = np.asarray(rows)
    cols = np.asarray(cols)
    vals = np.asarray(vals, dtype=np.float32)
    if rows.size and np.any(rows[1:] < rows[:-1]):
        order = np.argsort(rows, kind="stable")
        rows, cols, vals = rows[order], cols[order], vals[order]
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=row_ptr[1:])
    cols16 = (n <= 65536) if cols16 is None else cols16
    ct = np.uint16 if cols16 else np.int32
    out = []
    for a in (row_ptr, cols.astype(ct), vals):
        t = torch.from_numpy(np.ascontiguousarray(a))
        out.append(t.pin_memory() if torch.cuda.is_available() else t)
    return tuple(out)


# ------------------------------------------------------------------------------- a4/a5 O/E
class OEState:
    """Device-side by-products of the O/E passes (kept for inspection / return_exp_counts)."""

    def __init__(self, layout):
        dev, n, cc = layout.device, layout.n, layout.n_chr * layout.n_chr
        self.diag_sum = torch.empty(n, dtype=torch.float64, device=dev)
        self.diag_minpos = torch.empty(n, dtype=torch.float32, device=dev)
        self.inter_sum = torch.empty(cc, dtype=torch.float64, device=dev)
        self.inter_nnz = torch.empty(cc, dtype=torch.int64, device=dev)
        self.inter_minpos = torch.empty(cc, dtype=torch.float32, device=dev)
        self.expected = torch.empty(n, dtype=torch.float64, device=dev)
        self.inv_expected = torch.empty(n, dtype=torch.float32, device=dev)
        self.inter_inv = torch.empty(cc, dtype=torch.float32, device=dev)
        self.min_pos_oe = torch.empty(1, dtype=torch.float32, device=dev)
        self.inter_valid = torch.zeros(1, dtype=torch.int32, device=dev)   # set by scatter_*_sym(oe_state=...)


def obs_d_exp(dense, layout, bin_size, k=.2, counts_must_decrese=True, only_intra=False,
              log=False, out=None, state=None, from_upper=False, fp64_out=False, inter_from_scatter=False):
    """O/E (mode='sgd_mean') of a symmetric dense map, optionally fused with ArrayHiCMtx.log.
    Replaces HiCMtx.obs_d_exp (publish/new_hic_class.py:1121-1290) + log (:2076-2094).
    `out` may alias `dense` (in place). Returns (out, OEState). from_upper=True: only the upper
    triangle of `dense` is read (it may come from scatter_coo_sym(mtx_type=MTX_TRIU_UPPER)); the
    result is the same full symmetric map. fp64_out=True (raw O/E only, log=False): the result is
    a float64 [n, n] tensor divided in double like the reference -- 1e-5 absolute for entries of any
    size (the fp32 map carries its 6e-8 relative rounding, i.e. 1e-5 only up to values of ~50).
    inter_from_scatter=True: `state` was handed to the scatter that produced `dense`
    (scatter_*_sym(oe_state=state)); where that scatter already took the inter-chromosome statistics
    (`state.inter_valid` on the device) the pass over the inter part of the map is skipped."""
    lib = _lib.load()
    _need_cuda(dense)
    n = layout.n
    assert dense.shape == (n, n) and dense.dtype == torch.float32 and dense.stride(1) == 1
    pool_gid, use_code = layout.oe_tables(bin_size, k)
    state = OEState(layout) if state is None else state
    if out is None and not fp64_out:
        out = alloc_map(n, dense.device)
    assert fp64_out or out.stride(0) == dense.stride(0), "in/out maps must share the leading dimension"
    stream = _stream()
    st = lib.hia_oe_expected_pass_ex(_ptr(dense), n, dense.stride(0), _ptr(layout.chr_start),
                                     layout.chr_start_host.ctypes.data, layout.n_chr,
                                     _ptr(state.diag_sum), _ptr(state.diag_minpos),
                                     _ptr(state.inter_sum), _ptr(state.inter_nnz),
                                     _ptr(state.inter_minpos), _ptr(layout.bin_chr),
                                     _ptr(state.inter_valid) if inter_from_scatter else 0, stream)
    _lib.check(st, "hia_oe_expected_pass_ex")
    st = lib.hia_oe_finalize(_ptr(state.diag_sum), _ptr(state.diag_minpos), _ptr(state.inter_sum),
                             _ptr(state.inter_nnz), _ptr(state.inter_minpos),
                             _ptr(layout.chr_start), layout.n_chr, n, _ptr(pool_gid),
                             _ptr(use_code), 1 if counts_must_decrese else 0,
                             _ptr(state.expected), _ptr(state.inv_expected),
                             _ptr(state.inter_inv), _ptr(state.min_pos_oe), stream)
    _lib.check(st, "hia_oe_finalize")
    if fp64_out:
        if log or from_upper:
            raise _lib.HiaError("fp64_out is the raw O/E of a full symmetric map (log=False, from_upper=False)")
        out64 = torch.empty((n, n), dtype=torch.float64, device=dense.device)
        st = lib.hia_oe_apply_f64(_ptr(dense), n, dense.stride(0), _ptr(out64), n, _ptr(layout.chr_start),
                                  layout.n_chr, _ptr(layout.bin_chr), _ptr(state.expected), _ptr(state.inter_sum),
                                  _ptr(state.inter_nnz), OE_ONLY_INTRA if only_intra else 0, stream)
        _lib.check(st, "hia_oe_apply_f64")
        return out64, state
    flags = (OE_LOG if log else 0) | (OE_ONLY_INTRA if only_intra else 0) | (OE_FROM_UPPER if from_upper else 0)
    st = lib.hia_oe_apply(_ptr(dense), _ptr(out), n, dense.stride(0), _ptr(layout.chr_start),
                          layout.n_chr, _ptr(layout.bin_chr), _ptr(state.inv_expected),
                          _ptr(state.inter_inv), _ptr(state.min_pos_oe), flags, stream)
    _lib.check(st, "hia_oe_apply")
    return out, state


def log_map(x, out=None):
    """ArrayHiCMtx.log (publish/new_hic_class.py:2076-2094) on an arbitrary rows x cols view."""
    lib = _lib.load()
    _need_cuda(x)
    if out is None:
        out = alloc_map(x.shape[0], x.device, cols=x.shape[1])
    assert out.stride(0) == x.stride(0)
    ws = torch.empty(1, dtype=torch.float32, device=x.device)
    st = lib.hia_log_map(_ptr(x), _ptr(out), x.shape[0], x.shape[1], x.stride(0), _ptr(ws), _stream())
    _lib.check(st, "hia_log_map")
    return out


# ------------------------------------------------------------------------------- a8 similarity

def similarity(x, flavour=SIM_COSINE_DISTANCE, precision=None, band=None, k_chunk=0,
               out=None, workspace=None):
    """Row x row cosine distance / Pearson of an [m, k] fp32 block -> [m, m] fp32.
    Replaces get_adj_mtx (publish/complex/src/S1_get_adj_mtx.py:8-14)."""
    lib = _lib.load()
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.stride(1) == 1
    precision = PREC_DEFAULT if precision is None else precision
    m, k = x.shape
    if out is None:
        out = alloc_map(m, x.device)
        if band is not None:
            out.zero_()
    need = lib.hia_similarity_workspace_bytes(m, k, precision)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=x.device)
    band_lo, band_hi = (0, -1) if band is None else (int(band[0]), int(band[1]))
    st = lib.hia_similarity(_ptr(x), m, k, x.stride(0), _ptr(out), out.stride(0), flavour, precision,
                            band_lo, band_hi, k_chunk, _ptr(workspace), workspace.numel(), _stream())
    _lib.check(st, "hia_similarity")
    return out


# ------------------------------------------------------------------------------- a9 eigen
def eig_converged(resid, lam, gaps, tol, angle_tol):
    """Every wanted pair has ||C u - lambda u|| <= tol |lambda_1|, or <= angle_tol * gap (sin-theta
    theorem: the angle to the true eigenvector is at most residual / gap)."""
    if np.all(resid <= tol * max(abs(lam[0]), 1e-300)):
        return True
    return bool(angle_tol > 0 and np.all(resid <= angle_tol * gaps))


def _eig_shape(n, k, block, steps):
    if n < 2 * block:
        block = 4
    if steps is None:
        steps = 19              # Krylov dimension <= 160 (the partial projected eigensolver's limit); cycles
    steps = max(1, min(steps, n // block - 1, 272 // block - 1))     # end early at the first passed check
    if block * (steps + 1) > n or k > block:
        raise _lib.HiaError(f"matrix of {n} bins is too small for k={k}")
    return block, steps


class EigBuffers:
    """Workspaces + outputs of one eigen solve of an n x n matrix, allocated once (HotPath keeps one;
    a CUDA-graph capture needs every buffer to exist before the capture starts)."""

    def __init__(self, n, k, block, steps, device, symmetric=True):
        lib = _lib.load()
        self.n, self.k, self.block, self.steps = n, k, block, steps
        self.kk = block                                  # Ritz pairs kept: the whole block (restart quality)
        self.ws = torch.empty(lib.hia_topk_eig_workspace_bytes(n, block, steps), dtype=torch.uint8, device=device)
        self.upper = None
        # measured at 30 894 bins: the upper-triangle pass halves the DRAM bytes but not the FMAs or the
        # reductions, and the pass is issue-bound: 1.14 ms against 0.82 ms of the full pass -> opt-in
        if symmetric and os.environ.get("HIA_EIG_UPPER", "0") == "1":
            self.upper = torch.empty(lib.hia_eig_symm_upper_workspace_bytes(n, block), dtype=torch.uint8, device=device)
        self.eigvals = torch.empty(self.kk, dtype=torch.float64, device=device)
        self.eigvecs = torch.empty((self.kk, n), dtype=torch.float32, device=device)
        self.stats = torch.empty(2 * self.kk, dtype=torch.float64, device=device)  # residuals | gaps: one D2H read
        self.used = ctypes.c_int32(0)


def topk_eig_cycle(c, buf, seed=0, init=None, n_init=0, early_tol=0.0, early_angle_tol=0.0):
    """ONE Lanczos cycle enqueued on the current stream; results stay on the device in `buf`
    (eigvals, eigvecs, stats = true residuals | Ritz gaps). With early_tol = early_angle_tol = 0 the
    cycle runs all `buf.steps` steps and contains no host synchronisation at all (capturable in a
    CUDA graph); otherwise the in-cycle convergence checks read one flag each."""
    lib = _lib.load()
    n, k, kk = buf.n, buf.k, buf.kk
    upper = buf.upper if (buf.upper is not None and c.data_ptr() % 16 == 0 and c.stride(0) % 4 == 0) else None
    st = lib.hia_topk_eig(_ptr(c), n, c.stride(0), k, kk, buf.block, buf.steps, seed, _ptr(init), n_init,
                          early_tol, early_angle_tol, _ptr(buf.eigvals), _ptr(buf.eigvecs), buf.stats.data_ptr(),
                          buf.stats.data_ptr() + 8 * kk, ctypes.addressof(buf.used), _ptr(buf.ws), buf.ws.numel(),
                          _ptr(upper), 0 if upper is None else upper.numel(), _stream())
    _lib.check(st, "hia_topk_eig")


def topk_eig(c, k=3, block=8, steps=None, tol=1e-6, angle_tol=1e-3, max_restarts=8, seed=0,
             return_info=False, symmetric=True, workspace=None):
    """Largest-algebraic k eigenpairs of a dense symmetric fp32 matrix (e.g. the Pearson map).
    No reference implementation exists (oracle: numpy.linalg.eigh) -- north-star item.
    Block Lanczos on the device. A cycle of at most `steps` passes over the matrix ends as soon as
    the Lanczos residual ESTIMATES of the wanted pairs meet the tolerance (no extra pass); the
    true residuals ||C u - lambda u|| of one final pass decide convergence: each <= tol |lambda_1|,
    or each <= angle_tol * (distance to the nearest other Ritz value), i.e. an eigenvector angle of
    at most angle_tol (|cos| >= 1 - angle_tol^2 / 2; the north star asks for 0.9999). Otherwise the
    cycle restarts from the Ritz vectors. Returns (eigvals[k] fp64, eigvecs[n, k] fp32).
    symmetric=True with HIA_EIG_UPPER=1: the passes read the upper triangle of `c` only (half the DRAM
    bytes, but measured slower than the full pass, which is issue-bound: opt-in).
    workspace: optional dict reused across calls (buffers keyed by shape)."""
    _need_cuda(c)
    n = c.shape[0]
    assert c.shape == (n, n) and c.dtype == torch.float32 and c.stride(1) == 1
    block, steps = _eig_shape(n, k, block, steps)
    cache = workspace if workspace is not None else {}
    key = ("eig", n, k, block, steps, str(c.device), bool(symmetric))
    if key not in cache:
        cache[key] = EigBuffers(n, k, block, steps, c.device, symmetric)
    buf = cache[key]
    init, n_init, info = None, 0, []
    for it in range(max_restarts):
        # the in-cycle estimates use half the tolerances: the final, true residual must pass too
        topk_eig_cycle(c, buf, seed + it, init, n_init, 0.5 * tol, 0.5 * angle_tol)
        sg = buf.stats.cpu().numpy()
        lam = buf.eigvals.cpu().numpy()[:k]
        r, gaps = sg[:k], sg[buf.kk:buf.kk + k]
        info.append((float(r.max()), float(abs(lam[0])), int(buf.used.value)))
        if eig_converged(r, lam, gaps, tol, angle_tol):
            break
        init, n_init = buf.eigvecs.clone(), buf.kk          # restart from the whole block of Ritz vectors
    # fresh tensors: the buffers are reused by the next call
    out = (buf.eigvals[:k].clone(), buf.eigvecs[:k].t().clone() if workspace is not None else buf.eigvecs[:k].t())
    return out + (info,) if return_info else out


# ------------------------------------------------------------------------------- a6 selection
PCT_ALL, PCT_POS, PCT_NEG = 0, 1, 2


def percentiles(x, qs, modes=None):
    """np.percentile(x[, subset], q) for up to 4 (q, mode) pairs -> fp64 device tensor [nq].
    q in percent like numpy; the quantile q/100 is formed on the host in fp64 exactly like
    numpy does (true_divide(q, 100))."""
    lib = _lib.load()
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.stride(1) == 1
    qs = list(qs)
    modes = [PCT_ALL] * len(qs) if modes is None else list(modes)
    quant = np.true_divide(np.asarray(qs, dtype=np.float64), 100)
    d_quant = torch.from_numpy(quant).to(x.device)
    d_modes = torch.tensor(modes, dtype=torch.int32, device=x.device)
    out = torch.empty(len(qs), dtype=torch.float64, device=x.device)
    ws = torch.empty(lib.hia_percentiles_workspace_bytes(), dtype=torch.uint8, device=x.device)
    st = lib.hia_percentiles(_ptr(x), x.shape[0], x.shape[1], x.stride(0), _ptr(d_modes), _ptr(d_quant),
                             len(qs), _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(st, "hia_percentiles")
    return out


def _like(x, out):
    if out is None:
        out = alloc_map(x.shape[0], x.device, cols=x.shape[1])
    assert out.stride(0) == x.stride(0) and out.shape == x.shape
    return out


def clip(x, lo=None, hi=None, out=None):
    """x clipped to [lo, hi]; lo / hi are 1-element fp64 device tensors (or None)."""
    lib = _lib.load()
    _need_cuda(x)
    out = _like(x, out)
    st = lib.hia_clip(_ptr(x), _ptr(out), x.shape[0], x.shape[1], x.stride(0), _ptr(lo), _ptr(hi), _stream())
    _lib.check(st, "hia_clip")
    return out


def tid_off_outliers(x, vmax_per=98, out=None):
    """publish/oe_map/S3_denoise.py:85-92."""
    p = percentiles(x, [vmax_per, 100 - vmax_per])
    return clip(x, lo=p[1:2], hi=p[0:1], out=out)


def norm_to_zero_mean(x, mid_vmax_per=66, out=None):
    """publish/oe_map/S4_norm_value.py:10-27."""
    lib = _lib.load()
    _need_cuda(x)
    out = _like(x, out)
    p = percentiles(x, [mid_vmax_per, 100 - mid_vmax_per])         # vmax, vmin
    ms = torch.empty(2, dtype=torch.float64, device=x.device)
    ws6 = torch.empty(6, dtype=torch.float64, device=x.device)
    st = lib.hia_core_moments(_ptr(x), x.shape[0], x.shape[1], x.stride(0), _ptr(p[1:2]), _ptr(p[0:1]),
                              _ptr(ms), _ptr(ws6), _stream())
    _lib.check(st, "hia_core_moments")
    st = lib.hia_zscore(_ptr(x), _ptr(out), x.shape[0], x.shape[1], x.stride(0), _ptr(ms), _stream())
    _lib.check(st, "hia_zscore")
    return out


def remove_outliers(x, default_vmax_per=98, out=None):
    """publish/oe_map/S4_norm_value.py:30-40."""
    lib = _lib.load()
    p = percentiles(x, [default_vmax_per, 100 - default_vmax_per], [PCT_POS, PCT_NEG])
    b = torch.empty(2, dtype=torch.float64, device=x.device)
    _lib.check(lib.hia_sym_clip_bounds(_ptr(p), _ptr(b), _stream()), "hia_sym_clip_bounds")
    return clip(x, lo=b[0:1], hi=b[1:2], out=out)


def norm_de_mtx(x, out=None):
    """publish/oe_map/S4_norm_value.py:43-46."""
    z = norm_to_zero_mean(x, out=out)
    return remove_outliers(z, out=z)


# ------------------------------------------------------------------------------- a10 checkerboard
MIN_CHRO_LEN = 50            # publish/complex/src/main_local.py:16
MIN_N_SHORT_SIMS = 200       # publish/complex/src/main_local.py:17
_edges_cache = {}


def _hist_edges(device):
    key = str(device)
    if key not in _edges_cache:
        _edges_cache[key] = torch.from_numpy(np.linspace(0, 1.6, 30)).to(device)   # S2_get_entro.py:61
    return _edges_cache[key]


def band_entropy(adj, d_lo, d_hi, per=2):
    """The diagonal/entropy loop of get_local on a symmetric distance matrix.
    Returns a 2-element fp64 device tensor [score or NaN, number of entropy chunks]."""
    lib = _lib.load()
    _need_cuda(adj)
    n = adj.shape[0]
    out = torch.empty(2, dtype=torch.float64, device=adj.device)
    nd = max(0, d_hi - d_lo)
    ws = torch.empty(lib.hia_band_entropy_workspace_bytes(nd), dtype=torch.uint8, device=adj.device)
    st = lib.hia_band_entropy(_ptr(adj), n, adj.stride(0), d_lo, d_hi, _ptr(_hist_edges(adj.device)),
                              MIN_N_SHORT_SIMS, float(per), _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(st, "hia_band_entropy")
    return out


def get_local(block, L=None, short_dis_max_per=.15, precision=None):
    """Checkerboard score of one dense block (rows x cols view of a min-filled symmetric map).
    publish/complex/src/main_local.py:19-55. L = mean chromosome length of the block's row
    window (S2_get_entro.py:6-10); defaults to the number of rows. Returns float or None."""
    if min(block.shape) < MIN_CHRO_LEN:
        return None
    L = block.shape[0] if L is None else L
    smax = int(L * short_dis_max_per)
    smin = int(L * .05)
    if smax <= smin:
        return None
    if block.stride(1) != 1 or block.stride(0) % 4 or block.data_ptr() % 16:
        tmp = alloc_map(block.shape[0], block.device, cols=block.shape[1])
        tmp.copy_(block)
        block = tmp
    n = block.shape[0]
    precision = PREC_DEFAULT if precision is None else precision
    # only the tiles of the distance matrix that touch the band are computed
    adj = similarity(block, flavour=SIM_COSINE_DISTANCE, precision=precision,
                     band=(smin, min(smax - 1, n - 1)) if precision in SPLIT_PRECS else None,
                     out=alloc_map(n, block.device))
    res = band_entropy(adj, smin, smax).cpu().numpy()
    return None if res[1] == 0 else float(res[0])


def main_local(dense, layout, chros, short_dis_max_per=.15):
    """publish/complex/src/main_local.py:108-141 on a dense (per-block min-filled) 'all' map:
    intra + upper inter blocks, then the transposed inter block (skipped when the first
    orientation gave None; its L stays the first chromosome's length).
    Returns [(accord_chro, other_chro, accord)]."""
    out = []
    for a, (s1, e1) in enumerate(layout.seps):
        for b, (s2, e2) in enumerate(layout.seps):
            if s1 > s2:
                continue
            blk = dense[s1:e1 + 1, s2:e2 + 1]
            L = e1 - s1 + 1
            acc = get_local(blk, L, short_dis_max_per)
            if acc is None:
                continue
            out.append((chros[a], chros[b], acc))
            if a != b:
                acc = get_local(blk.t(), L, short_dis_max_per)
                if acc is None:
                    continue
                out.append((chros[b], chros[a], acc))
    return out


# ------------------------------------------------------------------------------- a11 GF_S1 profiles
def row_profiles(x, layout, threshold, reverse=False, norm=True):
    """(inter_profile, intra_profile) fp64 [N]: keep_high_cons fused with get_inter_rowsum /
    get_intra_rowsum. `threshold`: 1-element fp64 device tensor (the 66.7 / 33.3 percentile)."""
    lib = _lib.load()
    _need_cuda(x)
    n = layout.n
    inter = torch.empty(n, dtype=torch.float64, device=x.device)
    intra = torch.empty(n, dtype=torch.float64, device=x.device)
    ws = torch.empty(2 * n + 2, dtype=torch.float64, device=x.device)
    st = lib.hia_row_profiles(_ptr(x), n, x.stride(0), _ptr(layout.chr_start), _ptr(layout.bin_chr),
                              _ptr(threshold), 1 if reverse else 0, 1 if norm else 0, _ptr(inter),
                              _ptr(intra), _ptr(ws), _stream())
    _lib.check(st, "hia_row_profiles")
    return inter, intra


def smooth_rowsum(profile, layout, kernal_per=.1, min_kernal_len=5):
    """smooth_rowsum (S2_1_find_by_inter_rowsum.py:42-55)."""
    lib = _lib.load()
    out = torch.empty_like(profile)
    st = lib.hia_smooth1d_reflect(_ptr(profile), _ptr(out), layout.n, _ptr(layout.chr_start),
                                  _ptr(layout.bin_chr), float(kernal_per), int(min_kernal_len), _stream())
    _lib.check(st, "hia_smooth1d_reflect")
    return out


def inter_rowsum_profile(x, layout, per=100 / 3):
    """keep_high_cons -> get_inter_rowsum -> smooth_rowsum (main_find_anchors_by_inter_rowsum,
    S2_1:374-396). Returns (raw, smoothed) fp64 device tensors."""
    t = percentiles(x, [100 - per])
    inter, _ = row_profiles(x, layout, t[0:1], reverse=False)
    return inter, smooth_rowsum(inter, layout)


def intra_low_profile(x, layout, per=100 / 3):
    """keep_high_cons(reverse) -> get_intra_rowsum -> smooth_rowsum (S2_3:32-56)."""
    t = percentiles(x, [per])
    _, intra = row_profiles(x, layout, t[0:1], reverse=True)
    return intra, smooth_rowsum(intra, layout)


# ------------------------------------------------------------------------------- a14/a15 GF_S2
_tmpl_cache = {}


def _templates(device):
    """The shipped template maps (publish/confor/{intra,inter}_ave_maps.mat, keys sorted
    numerically), L1-normalised in float32 exactly like get_n_norm_ave_maps
    (to_confor.py:12-25), promoted to fp64."""
    import os
    key = str(device)
    if key not in _tmpl_cache:
        data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ave_maps.npz"))
        out = []
        for tag in ("intra", "inter"):
            t = data[tag]                                            # float32 [5, 256]
            norm = np.vstack([row / np.sum(np.abs(row)) for row in t])
            out.append(torch.from_numpy(norm.astype(np.float64)).to(device))
        _tmpl_cache[key] = out
    return _tmpl_cache[key]


INTRA_AVE_NAMES = ['End-whole', 'Large-center', 'Center', 'Center-end-axis', 'Center-whole']
INTER_AVE_NAMES = ['Center-center', 'Center-end-axis', 'Center-whole', 'End-whole', 'End-end']


def gf_scores(dense, layout, centers=None):
    """GF_S2 on a dense zero-filled 'all' map: every ordered chromosome pair -> pooled 16x16
    image -> z-score -> symmetrise -> 5 template scores. `centers[c]` = chromosome-relative
    index of the 'used' anchor or None. Returns dict of host arrays:
    pairs [(a, b)], xs / isym [P, 256], scores [P, 5] (fp64)."""
    lib = _lib.load()
    _need_cuda(dense)
    C = layout.n_chr
    centers = [None] * C if centers is None else centers
    desc, pairs = [], []
    for a, (s1, e1) in enumerate(layout.seps):
        for b, (s2, e2) in enumerate(layout.seps):
            c1 = -1 if centers[a] is None else int(centers[a])
            c2 = -1 if centers[b] is None else int(centers[b])
            desc.append([s1, e1 - s1 + 1, s2, e2 - s2 + 1, c1, c2, 1 if a == b else 0, 0])
            pairs.append((a, b))
    P = len(desc)
    d_desc = torch.tensor(desc, dtype=torch.int32, device=dense.device)
    t_intra, t_inter = _templates(dense.device)
    pooled = torch.empty((P, 256), dtype=torch.float64, device=dense.device)
    isym = torch.empty_like(pooled)
    xs = torch.empty_like(pooled)
    scores = torch.empty((P, 5), dtype=torch.float64, device=dense.device)
    st = lib.hia_gf_pool_scores(_ptr(dense), dense.stride(0), _ptr(d_desc), P, _ptr(t_intra),
                                _ptr(t_inter), _ptr(pooled), _ptr(isym), _ptr(xs), _ptr(scores),
                                _stream())
    _lib.check(st, "hia_gf_pool_scores")
    return {"pairs": pairs, "pooled": pooled.cpu().numpy(), "isym": isym.cpu().numpy(),
            "xs": xs.cpu().numpy(), "scores": scores.cpu().numpy()}


# ------------------------------------------------------------------------------- a7 filters
def box_filter_reflect(block, size, out=None, scratch=None):
    """scipy.ndimage.uniform_filter(block, size, mode='reflect') on a rows x cols view
    (denoise_one_method 'mean', publish/oe_map/S3_denoise.py:53-57)."""
    lib = _lib.load()
    _need_cuda(block)
    assert block.dtype == torch.float32 and block.stride(1) == 1
    rows, cols = block.shape
    out = block if out is None else out
    if scratch is None:
        scratch = torch.empty((rows, block.stride(0)), dtype=torch.float32, device=block.device)[:, :cols]
    assert out.stride(0) == block.stride(0) == scratch.stride(0)
    st = lib.hia_box_filter_reflect(_ptr(block), _ptr(out), _ptr(scratch), rows, cols, block.stride(0),
                                    int(size), _stream())
    _lib.check(st, "hia_box_filter_reflect")
    return out


def median_filter_reflect(block, size, out=None):
    """scipy.ndimage.median_filter(block, size, mode='reflect') (denoise_one_method 'median',
    publish/oe_map/S3_denoise.py:48-52)."""
    lib = _lib.load()
    _need_cuda(block)
    assert block.dtype == torch.float32 and block.stride(1) == 1
    rows, cols = block.shape
    if out is None:
        out = torch.empty((rows, block.stride(0)), dtype=torch.float32, device=block.device)[:, :cols]
    assert out.stride(0) == block.stride(0)
    st = lib.hia_median_filter_reflect(_ptr(block), _ptr(out), rows, cols, block.stride(0), int(size),
                                       _stream())
    _lib.check(st, "hia_median_filter_reflect")
    return out


def wavelet_default_levels(rows, cols):
    """scikit-image's default number of db3 levels for a rows x cols block."""
    return int(_lib.load().hia_wavelet_default_levels(int(rows), int(cols)))


def wavelet_denoise_db3(block, levels=None, out=None, workspace=None, return_stats=False):
    """skimage.restoration.denoise_wavelet(block, wavelet='db3', method="VisuShrink") of one dense
    block (denoise_one_method 'wavelet', publish/oe_map/S3_denoise.py:43-47): multi-level db3
    transform, soft threshold at sigma sqrt(2 ln size), inverse. fp32 or fp64 block, fp64 inside.
    `return_stats`: also the device tensor [sigma, threshold, nonzero finest-diagonal count]."""
    lib = _lib.load()
    _need_cuda(block)
    assert block.dim() == 2 and block.dtype in (torch.float32, torch.float64) and (block.numel() == 0 or block.stride(1) == 1)
    rows, cols = block.shape
    if out is None:
        if block.dtype == torch.float32:
            out = alloc_map(rows, block.device, cols=cols)
        else:
            out = torch.empty((rows, cols), dtype=block.dtype, device=block.device)
    assert out.dtype == block.dtype and out.shape == block.shape
    stats = torch.zeros(3, dtype=torch.float64, device=block.device)
    if rows and cols:
        lv = 0 if levels is None else int(levels)
        need = lib.hia_wavelet_workspace_bytes(rows, cols, lv)
        if need < 0:
            raise _lib.HiaError(f"wavelet_denoise_db3: levels={levels} not supported")
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=block.device)
        st = lib.hia_wavelet_denoise_db3(_ptr(block), block.stride(0), rows, cols, block.element_size(), lv,
                                         _ptr(out), out.stride(0), _ptr(stats), _ptr(workspace), workspace.numel(),
                                         _stream())
        _lib.check(st, "hia_wavelet_denoise_db3")
    return (out, stats) if return_stats else out


def used_filter(block, size, out, decimals=2, normalise=True, workspace=None):
    """used_filter of GF_S1 on one block: box filter + norm_to_zero_mean, fp64 inside with
    scipy's exact summation order (see csrc/hia_filter64.cu for why)."""
    lib = _lib.load()
    _need_cuda(block)
    rows, cols = block.shape
    need = lib.hia_used_filter_f64_workspace_bytes(rows, cols)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=block.device)
    st = lib.hia_used_filter_f64(_ptr(block), rows, cols, block.stride(0), int(size), int(decimals),
                                 1 if normalise else 0, 66.0, _ptr(out), out.stride(0), _ptr(workspace),
                                 workspace.numel(), _stream())
    _lib.check(st, "hia_used_filter_f64")
    return out


def main_filter(dense, layout, out=None, decimals=2):
    """GF_S1 main_filter (publish/confor/src/S1_preprocess/S1_1_filtering.py:21-32): per block
    box filter (size = int(len(row chromosome) * .1)) + norm_to_zero_mean, then one global
    remove_outliers. `dense` is the globally min-filled 'all' map; `decimals` = the decimal
    quantisation of its values (2 for a map read from a %.2f .mtx; -1 = not quantised)."""
    lib = _lib.load()
    n = layout.n
    out = alloc_map(n, dense.device) if out is None else out
    assert out.data_ptr() != dense.data_ptr()
    max_len = max(e - s + 1 for s, e in layout.seps)
    ws = torch.empty(lib.hia_used_filter_f64_workspace_bytes(max_len, max_len), dtype=torch.uint8,
                     device=dense.device)
    for (s1, e1) in layout.seps:
        size = int(np.mean(np.array([e1 - s1 + 1]) * .1))
        for (s2, e2) in layout.seps:
            used_filter(dense[s1:e1 + 1, s2:e2 + 1], size, out[s1:e1 + 1, s2:e2 + 1], decimals,
                        workspace=ws)
    return remove_outliers(out, out=out)


def norm_to_zero_mean_block(block, mid_vmax_per=66):
    """norm_to_zero_mean in place on an arbitrary (possibly unaligned) block view."""
    if block.data_ptr() % 16 == 0 and block.stride(0) % 4 == 0:
        return norm_to_zero_mean(block, mid_vmax_per, out=block)
    tmp = alloc_map(block.shape[0], block.device, cols=block.shape[1])
    tmp.copy_(block)
    norm_to_zero_mean(tmp, mid_vmax_per, out=tmp)
    block.copy_(tmp)
    return block


# ------------------------------------------------------------------------------- a13 intra-x
def intra_x_response(block, method=None, workspace=None):
    """get_intra_x (S2_2_find_by_intra_x.py:116-126) of one intra block -> fp64 device [n].

    method 'prefix' (default; env HIA_INTRA_X overrides): row prefix sums, O(n^2)
    (`hia_intra_x_response_prefix`); 'direct': every band cell of every centre, O(n^2 k_len)
    (`hia_intra_x_response`, the on-device cross-check)."""
    lib = _lib.load()
    _need_cuda(block)
    n = block.shape[0]
    assert block.shape == (n, n) and block.dtype == torch.float32 and block.stride(1) == 1
    k_len = max(3, int(n * .05))
    out = torch.empty(n, dtype=torch.float64, device=block.device)
    method = method or os.environ.get("HIA_INTRA_X", "prefix")
    if method not in ("prefix", "direct"):
        raise ValueError(f"intra_x_response: unknown method {method!r}")
    if method == "prefix":
        need = lib.hia_intra_x_workspace_bytes(n)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=block.device)
        st = lib.hia_intra_x_response_prefix(_ptr(block), n, block.stride(0), k_len, _ptr(out),
                                             _ptr(workspace), workspace.numel(), _stream())
        _lib.check(st, "hia_intra_x_response_prefix")
        return out
    st = lib.hia_intra_x_response(_ptr(block), n, block.stride(0), k_len, _ptr(out), _stream())
    _lib.check(st, "hia_intra_x_response")
    return out
