"""Row-sharded genome-wide similarity + top-k eigenvectors across the GPUs of one box.

The single exchange step of the hot path (BASELINE.json config C5:
123 541 bins = 61 GB per fp32 map, too large for one pass on one GPU). Rank r owns the rows
[r * per, (r + 1) * per) of the normalised map, `per` a multiple of the 128-row MMA tile:

  1. every rank builds the unit-norm hi/lo planes of ITS rows (fp16 split by   (hia_sim_prep_rows)
     default: 2 x 2 B/cell to exchange instead of 2 x 4)
  2. NCCL all-gather of the planes over NVLink (torch.distributed)            -> full planes
  3. every rank contracts its row block against all rows on tcgen05           (hia_sim_gemm_rows)
  4. eigenvectors: block Lanczos where the only pass over the matrix,
     Y[rows] = C[rows, :] V, is local and the n x 8 fp64 panel is all-gathered per step;
     the tiny orthogonalisation / Rayleigh-Ritz is replicated (bit-identical on all ranks).

`ShardedSimilarity` is the plain version (NCCL all-gather of the planes, every rank contracts its
full row block); `RingSimilarity` is the fast one: peer memory instead of collectives, each
unordered pair of row blocks contracted once, plane transfers on the copy engines behind the MMAs.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops


def rows_per_rank(n, world):
    per = (n + world - 1) // world
    return (per + 127) // 128 * 128


class ShardedSimilarity:
    def __init__(self, n, k, flavour=ops.SIM_PEARSON, group=None, device=None, precision=None):
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n, self.k, self.flavour = n, k, flavour
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.per = rows_per_rank(n, self.world)
        self.n_pad = self.per * self.world
        self.row0 = self.rank * self.per
        self.rows = max(0, min(n, self.row0 + self.per) - self.row0)
        self.precision = ops.PREC_DEFAULT if precision is None else precision
        assert self.precision in ops.SPLIT_PRECS
        self.kpad = self.lib.hia_sim_kpad(k, self.precision)
        pdt = torch.float16 if self.lib.hia_sim_plane_elem_bytes(self.precision) == 2 else torch.float32
        # full planes (all-gather targets); rows >= n stay zero
        self.hi = torch.zeros((self.n_pad, self.kpad), dtype=pdt, device=self.device)
        self.lo = torch.zeros((self.n_pad, self.kpad), dtype=pdt, device=self.device)
        self.own_hi = torch.zeros((self.per, self.kpad), dtype=pdt, device=self.device)
        self.own_lo = torch.zeros((self.per, self.kpad), dtype=pdt, device=self.device)
        self.ws = torch.empty(self.lib.hia_sim_gemm_rows_workspace_bytes(n, k), dtype=torch.uint8,
                              device=self.device)

    def run(self, x_rows, out_rows=None, k_chunk=0):
        """x_rows: this rank's rows of the map [rows, k] (fp32, stride(1) == 1).
        Returns out_rows [rows, n] = similarity of the own rows against all rows."""
        lib = self.lib
        assert x_rows.shape == (self.rows, self.k) and x_rows.dtype == torch.float32
        if self.rows > 0:
            st = lib.hia_sim_prep_rows(x_rows.data_ptr(), self.rows, self.k, x_rows.stride(0), self.flavour,
                                       self.precision, self.own_hi.data_ptr(), self.own_lo.data_ptr(),
                                       ops._stream())
            _lib.check(st, "hia_sim_prep_rows")
        if self.world > 1:
            dist.all_gather_into_tensor(self.hi.view(-1), self.own_hi.view(-1), group=self.group)
            dist.all_gather_into_tensor(self.lo.view(-1), self.own_lo.view(-1), group=self.group)
        else:
            self.hi.copy_(self.own_hi)
            self.lo.copy_(self.own_lo)
        if out_rows is None:
            out_rows = torch.empty((max(self.rows, 1), ops.padded_ld(self.n)), dtype=torch.float32,
                                   device=self.device)[:self.rows, :self.n]
        if self.rows > 0:
            st = lib.hia_sim_gemm_rows(self.hi.data_ptr(), self.lo.data_ptr(), self.precision, self.n_pad,
                                       self.n, self.k,
                                       self.row0, self.rows, out_rows.data_ptr(), out_rows.stride(0),
                                       self.flavour, k_chunk, self.ws.data_ptr(), self.ws.numel(),
                                       ops._stream())
            _lib.check(st, "hia_sim_gemm_rows")
        return out_rows


def ring_plan(world, rank, per, n):
    """The block products of `rank` in the symmetric ring schedule ("by symmetry only
    s >= r tiles are needed"). Every unordered pair of row blocks is contracted exactly once; the
    mirrored half C[R_s, R_r] = C[R_r, R_s]^T is written by the same kernel into the peer's rows.

      step 0          the diagonal block (upper triangle only: half a block product)
      step t          rank r against block s = r + t (mod world), 0 < 2 t < world: a full block
      step world / 2  (even world) the antipodal pair splits ITS block in two: the lower rank
                      contracts its first h rows against all of the peer's, the higher rank all
                      of its rows against the peer's rows [h, ...): half a block each, so every
                      rank does world / 2 block products in total (8 GPUs: 4 instead of 8).

    Returns a list of jobs (dict): t, peer, symmetric, a0, a_rows (own rows, block-local),
    b0, b_rows (the peer's rows, block-local). Jobs with no rows are omitted."""
    def valid(r):
        return max(0, min(n, (r + 1) * per) - r * per)

    jobs = []
    mine = valid(rank)
    if mine > 0:
        jobs.append(dict(t=0, peer=rank, symmetric=True, a0=0, a_rows=mine, b0=0, b_rows=mine))
    for t in range(1, world // 2 + 1):
        s = (rank + t) % world
        if 2 * t < world:
            a0, a_rows, b0, b_rows = 0, mine, 0, valid(s)
        elif 2 * t == world:
            h = (per // 2) // 128 * 128
            if rank < s:
                a0, a_rows, b0, b_rows = 0, min(h, mine), 0, valid(s)
            else:
                a0, a_rows, b0, b_rows = 0, mine, min(h, valid(s)), valid(s) - min(h, valid(s))
        else:
            continue
        if a_rows > 0 and b_rows > 0:
            jobs.append(dict(t=t, peer=s, symmetric=False, a0=a0, a_rows=a_rows, b0=b0, b_rows=b_rows))
    return jobs


class RingSimilarity:
    """Row-sharded similarity over peer memory: the symmetry is used (world / 2 block products per
    rank instead of world) and nothing but the tensor-core kernel runs on the SMs.

    Every rank exports the hi/lo planes of ITS rows and its output rows (CUDA IPC, peer.py). Per job of
    `ring_plan`: the peer block's planes are PULLED into one of two local staging slots by a copy
    engine on a side stream (7.7 GB per block at 25 kb over NVLink against a ~120 ms block product:
    off the critical path, and -- unlike an NCCL send/recv -- it takes no SM from the persistent
    148-CTA kernel, so the overlap is real); `hia_sim_gemm_block` contracts own rows x staged rows,
    stores C[R_r, R_s] into the own output rows and its transpose straight into the PEER's output
    rows over NVLink from the epilogue (bit-identical mirror by construction). Two tiny all-reduces
    (entry: every rank's planes exist and its previous result is no longer read; exit: every
    mirrored store has landed) are the only collectives."""

    def __init__(self, n, k, flavour=ops.SIM_PEARSON, group=None, device=None, precision=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n, self.k, self.flavour = n, k, flavour
        self.per = rows_per_rank(n, self.world)
        self.row0 = self.rank * self.per
        self.rows = self.valid(self.rank)
        self.jobs = ring_plan(self.world, self.rank, self.per, n)
        self.precision = ops.PREC_DEFAULT if precision is None else precision
        self.device = device
        self.stats = {}
        self._alloc()

    def valid(self, r):
        """Rows of block r that exist (the last block may be ragged or empty)."""
        return max(0, min(self.n, (r + 1) * self.per) - r * self.per)

    # ---- backend: CUDA library + peer memory (overridden by the CPU stand-in of the tests) ----------
    def _alloc(self):
        from .peer import PeerBuffer
        self.lib = _lib.load()
        assert self.precision in ops.SPLIT_PRECS
        self.device = self.device or torch.device("cuda", torch.cuda.current_device())
        self.kpad = self.lib.hia_sim_kpad(self.k, self.precision)
        self.elem = self.lib.hia_sim_plane_elem_bytes(self.precision)
        pdt = torch.float16 if self.elem == 2 else torch.float32
        # own planes [2 (hi, lo), per, kpad] (exported; rows >= self.rows stay zero) and two staging slots
        self.own = torch.zeros((2, self.per, self.kpad), dtype=pdt, device=self.device)
        self.slots = [torch.zeros((2, self.per, self.kpad), dtype=pdt, device=self.device) for _ in range(2)] \
            if self.world > 1 else []
        self.ld = ops.padded_ld(self.n)
        self.out_buf = torch.empty((self.per, self.ld), dtype=torch.float32, device=self.device)   # exported
        self.ws = torch.empty(self.lib.hia_sim_gemm_rows_workspace_bytes(self.per, self.k), dtype=torch.uint8,
                              device=self.device)
        self.peer_planes = PeerBuffer(self.own, self.group)
        self.peer_out = PeerBuffer(self.out_buf, self.group)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _prep(self, x_rows):
        if self.rows > 0:
            st = self.lib.hia_sim_prep_rows(x_rows.data_ptr(), self.rows, self.k, x_rows.stride(0), self.flavour,
                                            self.precision, self.own[0].data_ptr(), self.own[1].data_ptr(),
                                            ops._stream())
            _lib.check(st, "hia_sim_prep_rows")

    def _rendezvous(self):
        """Stream-ordered barrier: when it completes on this rank's stream, every rank has finished
        everything it enqueued before its own call."""
        if self.world > 1:
            dist.all_reduce(self._flag, group=self.group)

    def _pull(self, slot, job):
        """Planes rows [b0, b0 + b_rows) of the peer -> the same rows of staging slot `slot`, on the
        copy stream (the caller orders it after the slot's previous reader)."""
        row_bytes = self.kpad * self.elem
        for plane in range(2):
            off = (plane * self.per + job["b0"]) * row_bytes
            dst = self.slots[slot][plane, job["b0"]:job["b0"] + job["b_rows"]]
            self.peer_planes.pull(dst, job["peer"], off, job["b_rows"] * row_bytes, self.copy_stream.cuda_stream)

    def block_offsets(self, job):
        """Element offsets (backend independent) of a job's two stores: cell (i, j) of the block
        (i = own row a0 + i', j = peer row b0 + j') goes to
            own  output buffer[out_off + i' * ld + j']        (columns of block `peer`)
            peer output buffer[t_off   + j' * ld + i']        (columns of block `rank`, transposed)."""
        r, s, per, ld = self.rank, job["peer"], self.per, self.ld
        return job["a0"] * ld + s * per + job["b0"], job["b0"] * ld + r * per + job["a0"]

    def _block(self, slot, job):
        per, ld = self.per, self.ld
        out_off, t_off = self.block_offsets(job)
        out = self.out_buf.data_ptr() + 4 * out_off
        if job["symmetric"]:
            b_planes, out_t = self.own, out
        else:
            b_planes, out_t = self.slots[slot], self.peer_out.ptr(job["peer"]) + 4 * t_off
        st = self.lib.hia_sim_gemm_block(self.own[0].data_ptr(), self.own[1].data_ptr(), per,
                                         b_planes[0].data_ptr(), b_planes[1].data_ptr(), per,
                                         self.precision, self.k, job["a0"], job["a_rows"], job["b0"], job["b_rows"],
                                         1 if job["symmetric"] else 0, out, ld, out_t, ld, self.flavour, 0,
                                         self.ws.data_ptr(), self.ws.numel(), ops._stream())
        _lib.check(st, "hia_sim_gemm_block")

    def _streams(self):
        return torch.cuda.current_stream(), self.copy_stream

    def _event(self, timing=False):
        return torch.cuda.Event(enable_timing=timing)

    # ---- schedule (backend independent) ---------------------------------------------------------
    def run(self, x_rows, timing=False):
        """x_rows: this rank's rows of the map [rows, k]. Returns out_rows [rows, n] (a view of the
        exported output buffer: valid until the next run). timing=True records per-job CUDA events;
        `self.stats` then holds kernel / exposed-wait milliseconds after `collect_stats()`."""
        assert x_rows.shape == (self.rows, self.k)
        main, side = self._streams()
        self._prep(x_rows)
        self._rendezvous()                      # all planes exist; nobody still reads the previous result
        ready = self._event()
        ready.record(main)
        remote = [j for j in self.jobs if not j["symmetric"]]
        pulled, freed, marks = {}, {}, []

        def post_pull(i):
            slot = i & 1
            side.wait_event(freed.get(slot, ready))       # the slot's previous reader is done
            self._pull(slot, remote[i])
            ev = self._event()
            ev.record(side)
            pulled[i] = ev

        if remote:
            post_pull(0)
        ri = 0
        for job in self.jobs:
            slot = 0
            if not job["symmetric"]:
                slot = ri & 1
                if timing:
                    w0 = self._event(True)
                    w0.record(main)
                main.wait_event(pulled[ri])
                if ri + 1 < len(remote):
                    post_pull(ri + 1)                     # overlaps this job's contraction
            if timing:
                k0 = self._event(True)
                k0.record(main)
            self._block(slot, job)
            done = self._event(timing)
            done.record(main)
            if not job["symmetric"]:
                freed[slot] = done
                ri += 1
            if timing:
                marks.append((job, w0 if not job["symmetric"] else None, k0, done))
        self._rendezvous()                      # every rank's mirrored stores have landed
        self._marks = marks
        return self.out_buf[:self.rows, :self.n]

    def collect_stats(self):
        """After a synchronised run(timing=True): per-job kernel ms and the ms the main stream waited
        for planes (the exposed, non-overlapped part of the exchange)."""
        kern = [k0.elapsed_time(done) for _, _, k0, done in self._marks]
        wait = [w0.elapsed_time(k0) for _, w0, k0, _ in self._marks if w0 is not None]
        row_bytes = self.kpad * self.elem
        self.stats = {"block_ms": kern, "exposed_wait_ms": float(sum(wait)),
                      "pulled_bytes": int(sum(2 * j["b_rows"] * row_bytes for j in self.jobs if not j["symmetric"])),
                      "p2p_store_bytes": int(sum(4 * j["a_rows"] * j["b_rows"] for j in self.jobs
                                                 if not j["symmetric"])),
                      "block_products": float(sum((0.5 if j["symmetric"] else 1.0) * j["a_rows"] * j["b_rows"]
                                                  for j in self.jobs) / max(1, self.per) ** 2)}
        return self.stats


def topk_eig_rows(c_rows, n, k=3, block=8, steps=19, tol=1e-6, angle_tol=1e-3, max_restarts=8, seed=0,
                  group=None, phases=None):
    """Top-k eigenpairs of the symmetric n x n matrix whose rows [row0, row0 + rows) this rank
    holds in `c_rows` (row layout of ShardedSimilarity). Returns (eigvals [k] fp64 device,
    eigvecs [n, k] fp32 device, info) -- identical on every rank. `phases`: a dict that receives
    the device time per phase in ms (pass / gather / absorb / check / ritz; bench.py)."""
    lib = _lib.load()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    per = rows_per_rank(n, world)
    row0 = rank * per
    rows = max(0, min(n, row0 + per) - row0)
    assert c_rows.shape[0] == rows and c_rows.shape[1] == n
    dev = c_rows.device
    steps = max(1, min(steps, n // block - 1, 272 // block - 1))
    ws = torch.empty(lib.hia_topk_eig_workspace_bytes(n, block, steps), dtype=torch.uint8, device=dev)
    y_rows = torch.zeros((per, block), dtype=torch.float64, device=dev)
    y_full = torch.zeros((per * world, block), dtype=torch.float64, device=dev)
    kk = block                   # Ritz pairs kept: the whole block (restart quality on clustered spectra)
    eigvals = torch.empty(kk, dtype=torch.float64, device=dev)
    eigvecs = torch.empty((kk, n), dtype=torch.float32, device=dev)
    resid = torch.empty(kk, dtype=torch.float64, device=dev)
    gaps = torch.empty(kk, dtype=torch.float64, device=dev)
    conv = ctypes.c_int32(0)
    s = ops._stream

    marks = [] if phases is not None else None

    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    def matvec():
        mark("other")
        if rows > 0:
            _lib.check(lib.hia_eig_symm_rows(c_rows.data_ptr(), rows, n, c_rows.stride(0), block, steps,
                                             ws.data_ptr(), y_rows.data_ptr(), s()), "hia_eig_symm_rows")
        mark("pass")
        if world > 1:
            dist.all_gather_into_tensor(y_full.view(-1), y_rows.view(-1), group=group)
        else:
            y_full.copy_(y_rows)
        mark("gather")

    vote = torch.zeros(1, dtype=torch.int32, device=dev)

    def agree(flag):
        """Every rank must take the SAME branch (the collectives of the next step have to pair up).
        The replicated Lanczos state is equal only up to the rounding of the fp64 atomics in the Gram
        kernels, so a borderline estimate could flip on one rank: decide by a MIN all-reduce."""
        if world == 1:
            return bool(flag)
        vote.fill_(1 if flag else 0)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN, group=group)
        return bool(vote.item())

    init, n_init, info = None, 0, []
    for it in range(max_restarts):
        _lib.check(lib.hia_eig_begin(n, block, steps, seed + it, ops._ptr(init), n_init, ws.data_ptr(),
                                     ws.numel(), s()), "hia_eig_begin")
        used, reuse = steps, 0
        first_check, every = lib.hia_eig_first_check(n, steps), lib.hia_eig_check_every(n)
        for j in range(steps):
            matvec()
            _lib.check(lib.hia_eig_absorb(n, block, steps, j, y_full.data_ptr(), ws.data_ptr(), s()),
                       "hia_eig_absorb")
            mark("absorb")
            done = j + 1
            if done >= first_check and (done - first_check) % every == 0 and done < steps:
                # replicated on every rank; the branch is taken by agreement (see agree())
                _lib.check(lib.hia_eig_check(n, block, steps, done, k, 0.5 * tol, 0.5 * angle_tol,
                                             ctypes.addressof(conv), None, ws.data_ptr(), s()), "hia_eig_check")
                ok = agree(conv.value)
                mark("check")
                if ok:
                    used, reuse = done, 1
                    break
        _lib.check(lib.hia_eig_ritz(n, block, steps, used, reuse, kk, eigvals.data_ptr(), eigvecs.data_ptr(),
                                    gaps.data_ptr(), ws.data_ptr(), s()), "hia_eig_ritz")
        if world > 1:
            # one rank's Ritz pairs for everybody: with clustered Ritz values the replicated Jacobi
            # solves can order (or rotate) them differently by rounding alone, and a panel that
            # differs across ranks turns the all-gathered product C [U, 0] into garbage
            src = dist.get_global_rank(group, 0) if group is not None else 0
            for t in (eigvals, eigvecs, gaps):
                dist.broadcast(t, src=src, group=group)
            _lib.check(lib.hia_eig_set_ritz(n, block, steps, kk, eigvecs.data_ptr(), ws.data_ptr(), s()),
                       "hia_eig_set_ritz")
        mark("ritz")
        matvec()
        _lib.check(lib.hia_eig_resid(n, block, steps, kk, y_full.data_ptr(), eigvals.data_ptr(),
                                     resid.data_ptr(), ws.data_ptr(), s()), "hia_eig_resid")
        r, lam = resid.cpu().numpy()[:k], eigvals.cpu().numpy()[:k]
        mark("ritz")
        info.append((float(r.max()), float(abs(lam[0])), used))
        if agree(ops.eig_converged(r, lam, gaps.cpu().numpy()[:k], tol, angle_tol)):
            break
        init, n_init = eigvecs.clone(), kk                  # restart from the whole block of Ritz vectors
    if marks:
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            phases[name] = phases.get(name, 0.0) + e0.elapsed_time(e1)
            phases[f"n_{name}"] = phases.get(f"n_{name}", 0) + 1
    return eigvals[:k], eigvecs[:k].t(), info


def obs_d_exp_rows(map_rows, layout, bin_size, row0, k=.2, counts_must_decrese=True, only_intra=False,
                   log=True, out_rows=None, group=None, marks=None):
    """Row-sharded O/E (+log): `map_rows` = global rows [row0, row0 + rows) of the symmetric map.
    Partial diagonal / inter-block statistics are all-reduced (N + 3 C^2 numbers), the SGD pooling
    + clamp is replicated, the apply pass touches the own rows only. Returns (out_rows, state)."""
    lib = _lib.load()
    n = layout.n
    rows = map_rows.shape[0]
    assert map_rows.shape[1] == n and map_rows.dtype == torch.float32 and map_rows.stride(1) == 1
    pool_gid, use_code = layout.oe_tables(bin_size, k)
    st_ = ops.OEState(layout)
    s = ops._stream()

    def mark(name):                      # optional per-phase CUDA events (bench.py)
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    _lib.check(lib.hia_oe_expected_pass_rows(map_rows.data_ptr(), n, map_rows.stride(0), row0, rows,
                                             layout.chr_start.data_ptr(), layout.chr_start_host.ctypes.data,
                                             layout.n_chr, st_.diag_sum.data_ptr(), st_.diag_minpos.data_ptr(),
                                             st_.inter_sum.data_ptr(), st_.inter_nnz.data_ptr(),
                                             st_.inter_minpos.data_ptr(), layout.bin_chr.data_ptr(), s),
               "hia_oe_expected_pass_rows")
    mark("expected_pass")
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in (st_.diag_sum, st_.inter_sum, st_.inter_nnz):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        for t in (st_.diag_minpos, st_.inter_minpos):
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    mark("all_reduce")
    _lib.check(lib.hia_oe_finalize(st_.diag_sum.data_ptr(), st_.diag_minpos.data_ptr(), st_.inter_sum.data_ptr(),
                                   st_.inter_nnz.data_ptr(), st_.inter_minpos.data_ptr(),
                                   layout.chr_start.data_ptr(), layout.n_chr, n, pool_gid.data_ptr(),
                                   use_code.data_ptr(), 1 if counts_must_decrese else 0,
                                   st_.expected.data_ptr(), st_.inv_expected.data_ptr(),
                                   st_.inter_inv.data_ptr(), st_.min_pos_oe.data_ptr(), s), "hia_oe_finalize")
    mark("finalize")
    if out_rows is None:
        out_rows = torch.empty((max(rows, 1), map_rows.stride(0)), dtype=torch.float32,
                               device=map_rows.device)[:rows, :n]
    assert out_rows.stride(0) == map_rows.stride(0)
    flags = (ops.OE_LOG if log else 0) | (ops.OE_ONLY_INTRA if only_intra else 0)
    _lib.check(lib.hia_oe_apply_rows(map_rows.data_ptr(), out_rows.data_ptr(), n, map_rows.stride(0), row0, rows,
                                     layout.chr_start.data_ptr(), layout.n_chr, layout.bin_chr.data_ptr(),
                                     st_.inv_expected.data_ptr(), st_.inter_inv.data_ptr(),
                                     st_.min_pos_oe.data_ptr(), flags, s), "hia_oe_apply_rows")
    mark("apply")
    return out_rows, st_


class RowShardedPath:
    """The genome-wide high-resolution chain on the GPUs of one box (BASELINE.json config C5):
    count rows -> row-sharded O/E + log -> RingSimilarity (Pearson) -> top-k eigenvectors.
    Every rank holds the rows [rank * per, (rank + 1) * per) of every map. `ring=False` keeps the
    plain all-gather version (ShardedSimilarity), e.g. where peer mapping is not available."""

    def __init__(self, seps, bin_size, k_eig=3, group=None, device=None, ring=True, precision=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.layout = ops.GenomeLayout(seps, self.device)
        self.n = self.layout.n
        self.bin_size = int(bin_size)
        self.k_eig = k_eig
        self.layout.oe_tables(self.bin_size)
        self.per = rows_per_rank(self.n, self.world)
        self.row0 = self.rank * self.per
        self.rows = max(0, min(self.n, self.row0 + self.per) - self.row0)
        self.ring = ring
        if ring:
            self.sim = RingSimilarity(self.n, self.n, flavour=ops.SIM_PEARSON, group=group, device=self.device,
                                      precision=precision)
        else:
            self.sim = ShardedSimilarity(self.n, self.n, flavour=ops.SIM_PEARSON, group=group, device=self.device,
                                         precision=precision)
            self._out = torch.empty((max(self.rows, 1), ops.padded_ld(self.n)), dtype=torch.float32,
                                    device=self.device)[:self.rows, :self.n]
        self.oe_rows = None

    def normalise(self, count_rows, out_rows=None, marks=None):
        self.oe_rows, self.oe_state = obs_d_exp_rows(count_rows, self.layout, self.bin_size, self.row0, log=True,
                                                     out_rows=out_rows, group=self.group, marks=marks)
        return self.oe_rows

    def similarity(self, timing=False):
        if self.ring:
            self.sim_rows = self.sim.run(self.oe_rows, timing=timing)
        else:
            self.sim_rows = self.sim.run(self.oe_rows, out_rows=self._out)
        return self.sim_rows

    def eigen(self, tol=1e-6, phases=None):
        return topk_eig_rows(self.sim_rows, self.n, k=self.k_eig, tol=tol, group=self.group, phases=phases)

    def run(self, count_rows):
        self.normalise(count_rows)
        self.similarity()
        return self.eigen()

    def spot_check_fp64(self, n_own=16, n_other=64, seed=0):
        """Max |sim_rows[i, j] - Pearson_fp64(oe[i], oe[j])| over n_own own rows x n_other rows drawn
        from the whole genome (the same on every rank; their O/E rows are summed in from their
        owners with one all-reduce), in torch fp64 -- independent of the tcgen05 path. Returns
        (max abs error over all ranks, entries checked over all ranks)."""
        g = np.random.default_rng(seed)
        cols = np.sort(g.choice(self.n, size=min(n_other, self.n), replace=False))
        buf = torch.zeros((len(cols), self.n), dtype=torch.float32, device=self.device)
        for q, j in enumerate(cols):
            if self.row0 <= j < self.row0 + self.rows:
                buf[q] = self.oe_rows[j - self.row0]
        if self.world > 1:
            dist.all_reduce(buf, group=self.group)
        err, cnt = 0.0, 0
        if self.rows > 0:
            gi = np.random.default_rng(seed + 1 + self.rank)
            own = np.sort(gi.choice(self.rows, size=min(n_own, self.rows), replace=False))
            a = self.oe_rows[torch.as_tensor(own, device=self.device)].double()
            b = buf.double()
            a = a - a.mean(dim=1, keepdim=True)
            b = b - b.mean(dim=1, keepdim=True)
            ref = (a / a.norm(dim=1, keepdim=True)) @ (b / b.norm(dim=1, keepdim=True)).t()
            ref = ref.clamp(-1.0, 1.0)
            got = self.sim_rows[torch.as_tensor(own, device=self.device)][:, torch.as_tensor(cols, device=self.device)]
            same = torch.as_tensor(own + self.row0, device=self.device)[:, None] == \
                torch.as_tensor(cols, device=self.device)[None, :]
            ref = torch.where(same, torch.ones_like(ref), ref)
            err = float((got.double() - ref).abs().max())
            cnt = ref.numel()
        if self.world > 1:
            t = torch.tensor([err, float(cnt)], dtype=torch.float64, device=self.device)
            dist.all_reduce(t[0:1], op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM, group=self.group)
            err, cnt = float(t[0]), int(t[1])
        return err, cnt
