"""Thin Python wrappers over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Every function here launches hand-written sm_100a kernels through libhiarch_b200.so on the
current torch stream; torch only owns the memory. No op has a CPU or torch fallback.
"""
import numpy as np
import torch

from . import _lib
from .sgd import oe_tables

MTX_TRIU, MTX_ALL = 0, 1
FILL_ZERO, FILL_MIN, FILL_BLOCK_MIN = 0, 1, 2
OE_LOG, OE_ONLY_INTRA = 1, 2
SIM_COSINE_DISTANCE, SIM_PEARSON = 0, 1
PREC_3XTF32, PREC_FP64 = 0, 1


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
def scatter_coo_sym(rows, cols, vals, n, mtx_type=MTX_TRIU, fill_mode=FILL_ZERO, layout=None,
                    out=None):
    """COO (int32 rows, int32 cols, fp32 vals; CUDA) -> dense symmetric fp32 [n, n] view.
    Replaces read_sps_file/to_all/to_dense (publish/new_hic_class.py:2576-2602, :2229-2246,
    :2209-2224)."""
    lib = _lib.load()
    _need_cuda(rows, cols, vals)
    assert rows.dtype == torch.int32 and cols.dtype == torch.int32 and vals.dtype == torch.float32
    rows, cols, vals = rows.contiguous(), cols.contiguous(), vals.contiguous()
    if out is None:
        out = alloc_map(n, rows.device)
    ws = None
    n_chr = 0
    chr_start = None
    if fill_mode != FILL_ZERO:
        if layout is None:
            layout = GenomeLayout([(0, n - 1)], rows.device)
        n_chr, chr_start = layout.n_chr, layout.chr_start
        ws = torch.empty(lib.hia_scatter_workspace_bytes(n_chr) // 4, dtype=torch.float32,
                         device=rows.device)
    st = lib.hia_scatter_coo_sym(_ptr(rows), _ptr(cols), _ptr(vals), rows.numel(), n, _ptr(out),
                                 out.stride(0), mtx_type, fill_mode, _ptr(chr_start), n_chr,
                                 _ptr(ws), _stream())
    _lib.check(st, "hia_scatter_coo_sym")
    return out


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


def obs_d_exp(dense, layout, bin_size, k=.2, counts_must_decrese=True, only_intra=False,
              log=False, out=None, state=None):
    """O/E (mode='sgd_mean') of a symmetric dense map, optionally fused with ArrayHiCMtx.log.
    Replaces HiCMtx.obs_d_exp (publish/new_hic_class.py:1121-1290) + log (:2076-2094).
    `out` may alias `dense` (in place). Returns (out, OEState)."""
    lib = _lib.load()
    _need_cuda(dense)
    n = layout.n
    assert dense.shape == (n, n) and dense.dtype == torch.float32 and dense.stride(1) == 1
    pool_gid, use_code = layout.oe_tables(bin_size, k)
    state = OEState(layout) if state is None else state
    if out is None:
        out = alloc_map(n, dense.device)
    assert out.stride(0) == dense.stride(0), "in/out maps must share the leading dimension"
    stream = _stream()
    st = lib.hia_oe_expected_pass(_ptr(dense), n, dense.stride(0), _ptr(layout.chr_start),
                                  layout.chr_start_host.ctypes.data, layout.n_chr,
                                  _ptr(state.diag_sum), _ptr(state.diag_minpos),
                                  _ptr(state.inter_sum), _ptr(state.inter_nnz),
                                  _ptr(state.inter_minpos), _ptr(layout.bin_chr), stream)
    _lib.check(st, "hia_oe_expected_pass")
    st = lib.hia_oe_finalize(_ptr(state.diag_sum), _ptr(state.diag_minpos), _ptr(state.inter_sum),
                             _ptr(state.inter_nnz), _ptr(state.inter_minpos),
                             _ptr(layout.chr_start), layout.n_chr, n, _ptr(pool_gid),
                             _ptr(use_code), 1 if counts_must_decrese else 0,
                             _ptr(state.expected), _ptr(state.inv_expected),
                             _ptr(state.inter_inv), _ptr(state.min_pos_oe), stream)
    _lib.check(st, "hia_oe_finalize")
    flags = (OE_LOG if log else 0) | (OE_ONLY_INTRA if only_intra else 0)
    st = lib.hia_oe_apply(_ptr(dense), _ptr(out), n, dense.stride(0), _ptr(layout.chr_start),
                          layout.n_chr, _ptr(layout.bin_chr), _ptr(state.inv_expected),
                          _ptr(state.inter_inv), _ptr(state.min_pos_oe), flags, stream)
    _lib.check(st, "hia_oe_apply")
    return out, state


# ------------------------------------------------------------------------------- a8 similarity
_ws_cache = {}


def similarity(x, flavour=SIM_COSINE_DISTANCE, precision=PREC_3XTF32, band=None, k_chunk=0,
               out=None, workspace=None):
    """Row x row cosine distance / Pearson of an [m, k] fp32 block -> [m, m] fp32.
    Replaces get_adj_mtx (publish/complex/src/S1_get_adj_mtx.py:8-14)."""
    lib = _lib.load()
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.stride(1) == 1
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
def topk_eig(c, k=3, block=8, steps=None, tol=1e-6, max_restarts=8, seed=0, return_info=False):
    """Largest-algebraic k eigenpairs of a dense symmetric fp32 matrix (e.g. the Pearson map).
    No reference implementation exists (oracle: numpy.linalg.eigh) -- north-star item.
    Block Lanczos on the device; restarted from the current Ritz vectors until every residual
    ||C u - lambda u|| <= tol * |lambda_1|. Returns (eigvals[k] fp64, eigvecs[n, k] fp32)."""
    lib = _lib.load()
    _need_cuda(c)
    n = c.shape[0]
    assert c.shape == (n, n) and c.dtype == torch.float32 and c.stride(1) == 1
    if n < 2 * block:
        block = 4
    if steps is None:
        steps = 12
    steps = max(1, min(steps, n // block - 1, 272 // block - 1))
    if block * (steps + 1) > n or k > block:
        raise _lib.HiaError(f"matrix of {n} bins is too small for k={k}")
    dev = c.device
    ws = torch.empty(lib.hia_topk_eig_workspace_bytes(n, block, steps), dtype=torch.uint8, device=dev)
    eigvals = torch.empty(k, dtype=torch.float64, device=dev)
    eigvecs = torch.empty((k, n), dtype=torch.float32, device=dev)
    resid = torch.empty(k, dtype=torch.float64, device=dev)
    init, n_init, info = None, 0, []
    for it in range(max_restarts):
        st = lib.hia_topk_eig(_ptr(c), n, c.stride(0), k, block, steps, seed + it, _ptr(init), n_init,
                              _ptr(eigvals), _ptr(eigvecs), _ptr(resid), _ptr(ws), ws.numel(),
                              _stream())
        _lib.check(st, "hia_topk_eig")
        r = resid.cpu().numpy()
        lam = eigvals.cpu().numpy()
        info.append((float(r.max()), float(abs(lam[0]))))
        if np.all(r <= tol * max(abs(lam[0]), 1e-300)):
            break
        init, n_init = eigvecs.clone(), k
    out = (eigvals, eigvecs.t())
    return out + (info,) if return_info else out
