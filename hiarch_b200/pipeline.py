"""The genome-wide hot path as one call: sparse (i, j, v) -> dense symmetric map -> O/E + log
-> row x row Pearson (or cosine distance) -> top-k eigenvectors.

`HotPath` owns every device buffer (maps, workspaces) so that repeated calls do not allocate;
this is the object `bench.py` times and the call a user of the class-level API makes for the
BASELINE.json configs that the shipped CLI pipeline never reaches.
"""
import numpy as np
import torch

from . import _lib, ops


class HotPath:
    def __init__(self, seps, bin_size, k_eig=3, flavour=ops.SIM_PEARSON, device="cuda",
                 max_nnz=None, eig_block=8, eig_steps=12, precision=None):
        self.lib = _lib.load()
        self.layout = ops.GenomeLayout(seps, device)
        self.n = self.layout.n
        self.bin_size = int(bin_size)
        self.k_eig = k_eig
        self.flavour = flavour
        self.device = device
        n = self.n
        self.raw = ops.alloc_map(n, device)          # dense symmetric counts
        self.oe = ops.alloc_map(n, device)           # O/E + log
        self.sim = ops.alloc_map(n, device)          # Pearson / cosine distance
        self.oe_state = ops.OEState(self.layout)
        self.precision = ops.PREC_DEFAULT if precision is None else precision
        assert self.precision in ops.SPLIT_PRECS
        self.sim_ws = torch.empty(self.lib.hia_similarity_workspace_bytes(n, n, self.precision),
                                  dtype=torch.uint8, device=device)
        self.eig_block, self.eig_steps = eig_block, eig_steps
        self.layout.oe_tables(self.bin_size)         # build + upload the SGD pooling tables once
        self.max_nnz = max_nnz
        if max_nnz:
            self.d_rows = torch.empty(max_nnz, dtype=torch.int32, device=device)
            self.d_cols = torch.empty(max_nnz, dtype=torch.int32, device=device)
            self.d_vals = torch.empty(max_nnz, dtype=torch.float32, device=device)

    # ---- stages -------------------------------------------------------------------------
    def scatter(self, rows, cols, vals):
        return ops.scatter_coo_sym(rows, cols, vals, self.n, out=self.raw)

    def normalise(self):
        ops.obs_d_exp(self.raw, self.layout, self.bin_size, log=True, out=self.oe, state=self.oe_state)
        return self.oe

    def similarity_prep(self):
        st = self.lib.hia_similarity_prep(self.oe.data_ptr(), self.n, self.n, self.oe.stride(0),
                                          self.flavour, self.precision, self.sim_ws.data_ptr(),
                                          self.sim_ws.numel(), ops._stream())
        _lib.check(st, "hia_similarity_prep")

    def similarity_gemm(self):
        st = self.lib.hia_similarity_gemm(self.n, self.n, self.sim.data_ptr(), self.sim.stride(0),
                                          self.flavour, self.precision, 0, -1, 0, self.sim_ws.data_ptr(),
                                          self.sim_ws.numel(), ops._stream())
        _lib.check(st, "hia_similarity_gemm")

    def similarity(self):
        self.similarity_prep()
        self.similarity_gemm()
        return self.sim

    def eigen(self, tol=1e-6):
        return ops.topk_eig(self.sim, k=self.k_eig, block=self.eig_block, steps=self.eig_steps, tol=tol)

    # ---- whole path ---------------------------------------------------------------------
    def run_device(self, rows, cols, vals):
        """Inputs already in HBM. Returns (eigvals [k] fp64, eigvecs [n, k] fp32) on the device."""
        self.scatter(rows, cols, vals)
        self.normalise()
        self.similarity()
        return self.eigen()

    def run_host(self, rows_h, cols_h, vals_h):
        """End to end from pinned HOST buffers: H2D of the COO, the whole path, D2H of the
        eigenpairs. Returns (eigvals ndarray [k], eigvecs ndarray [n, k])."""
        nnz = rows_h.numel()
        assert self.max_nnz and nnz <= self.max_nnz
        r, c, v = self.d_rows[:nnz], self.d_cols[:nnz], self.d_vals[:nnz]
        r.copy_(rows_h, non_blocking=True)
        c.copy_(cols_h, non_blocking=True)
        v.copy_(vals_h, non_blocking=True)
        w, vecs = self.run_device(r, c, v)
        return w.cpu().numpy(), vecs.cpu().numpy()

    @property
    def h2d_bytes(self):
        return 12

    def cells(self):
        return self.n * self.n


def dense_to_coo(dense):
    """Upper-triangle non-zeros of a symmetric device map as (rows i32, cols i32, vals f32)."""
    up = torch.triu(dense)
    idx = up.nonzero(as_tuple=False)
    rows = idx[:, 0].to(torch.int32).contiguous()
    cols = idx[:, 1].to(torch.int32).contiguous()
    vals = up[idx[:, 0], idx[:, 1]].contiguous()
    return rows, cols, vals
