#!/usr/bin/env python
"""bench.py -- the HiArch matrix hot path on B200: scatter -> O/E+log -> Pearson -> top-3 eigvec.

    python bench.py --gpus N --steps K --warmup W            # our arm (N=1: plain python;
    python -m torch.distributed.run ... bench.py --gpus N    #  N>1: one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the CPU reference arm (oracle port)

Metric (BASELINE.json): "O/E+Pearson+eigen cells/s". A cell = one entry of the dense N x N map.
Workload at N GPUs (weak scaling, no data-path collective): every rank processes its own
synthetic genome-wide human map at 100 kb (configs[1]: 30 894 bins, 9.54e8 cells, 24 hg38
chromosomes), i.e. samples are sharded by rank like the reference's ParallelFunc
(publish/run_hic_dataset.py:573-603).

One step = sparse (i,j,v) upper triangle resident in HBM -> dense symmetric scatter -> expected /
O-E / log (sgd_mean, k=.2) -> row x row Pearson on tcgen05 (3-term fp16 hi/lo split by default,
--precision tf32x3 for the TF32 split) -> top-3 eigenvectors (block
Lanczos). `value` = cells of all ranks / max-over-ranks device time. `e2e` = the same through
HotPath.run_host: COO in pinned host memory, H2D inside the timed region, eigenpairs read back.
Inputs (0.9 GB COO / 3.8 GB maps) are far larger than the 126 MB L2, so no explicit L2 flush.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "O/E+Pearson+eigen cells/s"
UNIT = "cells/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=100000, help="bin size (bp)")
    ap.add_argument("--chroms", type=int, default=24, help="first C hg38 chromosomes")
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "tf32x3"],
                    help="operand split of the Pearson contraction (both carry 22 significant bits)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-checks", action="store_true", help="skip the in-run fp64 parity checks")
    ap.add_argument("--no-row-sharded", action="store_true", help="N > 1: skip the row-sharded genome-wide chain")
    ap.add_argument("--no-c3", action="store_true", help="skip the per-chromosome batch (config C3)")
    ap.add_argument("--only-row-sharded", action="store_true",
                    help="N > 1: run only the row-sharded chain and print its record (development sweeps)")
    ap.add_argument("--c3-samples", type=int, default=64)
    ap.add_argument("--row-sharded-res", type=int, default=0,
                    help="bin size of the row-sharded chain (default: 50 kb at 2 GPUs, 25 kb at 4 / 8)")
    ap.add_argument("--cpu-sample-chroms", default="chr16,chr17,chr18,chr19,chr20,chr21,chr22")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy/BLAS restatement of the reference, oracle/hia_oracle.py)
# --------------------------------------------------------------------------------------------
def cpu_sample(args):
    from hiarch_b200 import synth
    chroms = args.cpu_sample_chroms.split(",")
    bins = synth.hg38_bins(args.res, chroms)
    lens = [n for _, n in bins]
    lam, seps = synth.rate_matrix(lens, seed=0)
    up = synth.sample_counts(lam, seed=0, jitter=False)
    dense = up + np.triu(up, 1).T
    return dense, seps, chroms


def cpu_step(dense, seps, res):
    """One pass of the hot path on the host: O/E+log -> Pearson -> top-3 eigenpairs. The
    eigenpairs come from ARPACK (scipy eigsh, k=3, 'LA': a Lanczos solver like ours) -- the oracle's
    full `eigh` is the exact checker of the tests, but as a TIMED baseline it would charge the CPU
    an O(n^3) decomposition nobody needs (4.8 s of 7.8 s at 4 750 bins vs 0.1 s)."""
    from oracle import hia_oracle as O
    from scipy.sparse.linalg import eigsh
    oe = O.oe_map_core(dense, seps, res)
    c = O.pearson(oe)
    w, v = eigsh(c, k=3, which="LA")
    order = np.argsort(w)[::-1]
    return w[order], v[:, order]


def workload_name(args, n):
    return (f"genome-wide human {args.res // 1000} kb ({n} bins, {float(n) * n:.3g} cells) "
            "scatter -> O/E+log -> Pearson -> top-3 eigvec, one sample per GPU")


def config_of(args, n, n_chr):
    """The SAME dict in both arms (the driver compares them)."""
    return {"workload": workload_name(args, n), "bins": n, "chromosomes": n_chr,
            "l2": "inputs (>=0.9 GB) exceed the 126 MB L2; no explicit flush",
            "sharding": "samples by rank, no collective (the row_sharded sub-record is the path with an exchange)"}


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        t = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(t) if t else os.cpu_count()
    except Exception:
        return os.cpu_count()


def all_host_cores():
    """Context manager: BLAS on every host core, whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1 for N > 1, which round 1's reference arm silently inherited)."""
    from threadpoolctl import threadpool_limits
    return threadpool_limits(limits=os.cpu_count(), user_api="blas")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dense, seps, chroms = cpu_sample(args)
    n = dense.shape[0]
    from hiarch_b200 import synth
    n_full = sum(nb for _, nb in synth.hg38_bins(args.res)[: args.chroms])
    steps = max(1, args.steps)
    with all_host_cores():
        cores = blas_threads()
        for _ in range(max(0, args.warmup)):
            cpu_step(dense, seps, args.res)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_step(dense, seps, args.res)
        dt = (time.perf_counter() - t0) / steps
    value = n * n / dt
    sample = (f"{'+'.join(chroms)} at {args.res // 1000} kb = {n} bins ({n * n:.3g} cells) per step; "
              "O/E+log numpy per-diagonal loops 1 thread, Pearson GEMM (BLAS, all threads) + ARPACK eigsh(k=3)")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference",
        "config": config_of(args, n_full, min(args.chroms, 24)),
        "sample": f"each step = a bounded sample of that workload: {n} bins ({'+'.join(chroms)})",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}



# --------------------------------------------------------------------------------------------
# in-run checks, the like-for-like sample, the row-sharded chain
# --------------------------------------------------------------------------------------------
def parity_checks(hp, rows, cols, vals, w_h, v_h, stripe=384):
    """Asserted on every rank after the timed steps (state of the last step):
      scatter  bit-exact against a torch index_put of the same COO (sampled rows);
      O/E+log  one inter-chromosome block against  ln(1 + v / mean(non-zero))  in fp64 (<= 1e-5);
      Pearson  a `stripe`-row stripe against fp64 centred / normalised dot products (<= 1e-5);
      eigen    ||C u - lambda u|| / lambda_1 and orthonormality of the returned vectors in fp64.
    The oracle comparisons proper (incl. the intra-chromosome SGD pooling) are tests/."""
    import torch
    n = hp.n
    dev = hp.raw.device
    out = {}
    g = torch.Generator(device="cpu")
    g.manual_seed(11)
    # scatter: rows r0..r0+255 of the dense map rebuilt from the COO
    r0 = int(torch.randint(0, max(1, n - 256), (1,), generator=g))
    sel = (rows >= r0) & (rows < r0 + 256)
    ref = torch.zeros((256, n), dtype=torch.float32, device=dev)
    ref.index_put_(((rows[sel] - r0).long(), cols[sel].long()), vals[sel], accumulate=True)
    got = hp.raw[r0:r0 + 256]
    cmask = torch.arange(n, device=dev)[None, :] >= (torch.arange(256, device=dev) + r0)[:, None]
    out["scatter_upper_mismatches"] = int(((got != ref) & cmask).sum())
    out["raw_map"] = "upper triangle only (no mirror pass)" if hp.upper_only else "symmetric"
    if not hp.upper_only:
        d = hp.raw[r0:r0 + 256, r0:r0 + 256]
        out["scatter_symmetry_mismatches"] = int((d != d.t()).sum())
        assert out["scatter_symmetry_mismatches"] == 0, out
    assert out["scatter_upper_mismatches"] == 0, out
    blk_oe = hp.oe[r0:r0 + 256, r0 + 300:r0 + 556] if r0 + 556 <= n else None
    if blk_oe is not None:                                   # the O/E map itself is symmetric bit for bit
        out["oe_symmetry_mismatches"] = int((blk_oe != hp.oe[r0 + 300:r0 + 556, r0:r0 + 256].t()).sum())
        assert out["oe_symmetry_mismatches"] == 0, out
    # O/E + log on the inter block (chr 1 rows, last chromosome's columns)
    (s1, e1), (s2, e2) = hp.layout.seps[0], hp.layout.seps[-1]
    if len(hp.layout.seps) > 1:
        blk = hp.raw[s1:e1 + 1, s2:e2 + 1].double()
        nz = blk != 0
        ref_oe = torch.log1p(blk / (blk.sum() / nz.sum()))
        got_oe = hp.oe[s1:e1 + 1, s2:e2 + 1].double()
        out["oe_inter_block_max_abs_err"] = float((got_oe - ref_oe)[blk > 0].abs().max())
        assert out["oe_inter_block_max_abs_err"] <= 1e-5, out
    # Pearson stripe in fp64
    q0 = int(torch.randint(0, max(1, n - stripe), (1,), generator=g))
    def unit(a):
        a = a - a.mean(dim=1, keepdim=True)
        return a / a.norm(dim=1, keepdim=True)
    a = unit(hp.oe[q0:q0 + stripe].double())
    err = 0.0
    for c0 in range(0, n, 4096):
        b = unit(hp.oe[c0:c0 + 4096].double())
        ref_c = (a @ b.t()).clamp(-1.0, 1.0)
        got_c = hp.sim[q0:q0 + stripe, c0:c0 + 4096].double()
        ii = torch.arange(q0, q0 + a.shape[0], device=dev)[:, None]
        jj = torch.arange(c0, c0 + b.shape[0], device=dev)[None, :]
        ref_c = torch.where(ii == jj, torch.ones_like(ref_c), ref_c)
        err = max(err, float((got_c - ref_c).abs().max()))
    out["pearson_stripe_rows"] = int(a.shape[0])
    out["pearson_stripe_max_abs_err"] = err
    assert err <= 1e-5, out
    # eigenpairs: residual against the GPU's own C in fp64, row block by row block
    v = torch.from_numpy(np.ascontiguousarray(v_h)).to(dev).double()
    lam = torch.from_numpy(np.ascontiguousarray(w_h)).to(dev).double()
    cu = torch.empty_like(v)
    for c0 in range(0, n, 4096):
        cu[c0:c0 + 4096] = hp.sim[c0:c0 + 4096].double() @ v
    res = (cu - v * lam[None, :]).norm(dim=0) / lam[0].abs()
    gram = v.t() @ v
    out["eig_rel_residual_max"] = float(res.max())
    out["eig_orthonormality_err"] = float((gram - torch.eye(v.shape[1], device=dev, dtype=torch.float64)).abs().max())
    gaps = (lam[:-1] - lam[1:]).abs().min() if lam.numel() > 1 else lam[0].abs()
    # sin(angle to the true eigenvector) <= residual / gap; |cos| >= 0.9999 <=> sin <= 1.4e-2
    out["eig_angle_bound"] = float((res * lam[0].abs()).max() / gaps)
    assert out["eig_angle_bound"] <= 1.4e-2 and out["eig_orthonormality_err"] <= 1e-5, out
    return out


def gpu_on_sample(args, dense_host, seps, prec, steps=5, warmup=3):
    """The GPU arm on the CPU baseline's own bounded sample (same map, same chain)."""
    import torch
    from hiarch_b200.pipeline import HotPath
    up = np.triu(dense_host)
    r, c = np.nonzero(up)
    rows = torch.from_numpy(r.astype(np.int32)).cuda()
    cols = torch.from_numpy(c.astype(np.int32)).cuda()
    vals = torch.from_numpy(up[r, c].astype(np.float32)).cuda()
    nnz, n = rows.numel(), dense_host.shape[0]
    hp = HotPath(seps, args.res, k_eig=3, max_nnz=nnz, precision=prec)
    rows_h, cols_h, vals_h = (t.cpu().pin_memory() for t in (rows, cols, vals))
    for _ in range(warmup):
        hp.run_device(rows, cols, vals)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        hp.run_device(rows, cols, vals)
    b.record()
    torch.cuda.synchronize()
    dev_ms = a.elapsed_time(b) / steps
    for _ in hp.run_host_many([(rows_h, cols_h, vals_h)] * 2):
        pass
    torch.cuda.synchronize()
    a.record()
    for _ in hp.run_host_many([(rows_h, cols_h, vals_h)] * steps):
        pass
    b.record()
    torch.cuda.synchronize()
    e2e_ms = a.elapsed_time(b) / steps
    return {"gpu_ms_per_step": dev_ms, "gpu_cells_per_s": n * n / (dev_ms * 1e-3),
            "gpu_e2e_ms_per_step": e2e_ms, "gpu_e2e_cells_per_s": n * n / (e2e_ms * 1e-3)}


def run_c3(args, world, rank, max_over_ranks, barrier, res=50000, workers=8):
    """BASELINE.json config C3: `--c3-samples` (64) samples x 24 per-chromosome maps at 50 kb (935 ... 4 980
    bins, 1.9e8 cells per sample), sharded by sample over the ranks like the reference's ParallelFunc
    (publish/run_hic_dataset.py:259-269, :573-603), no communication. Each chromosome goes through
    scatter -> O/E + log -> Pearson -> top-3 eigenvectors; the 24 chains of a sample run on 8 host
    threads / streams (pipeline.HotPathPool: chromosome-sized chains are latency-bound). Every rank
    re-runs ITS synthetic sample (device-resident COO) samples / world times."""
    import torch
    from hiarch_b200 import synth
    from hiarch_b200.pipeline import HotPath, HotPathPool, dense_to_coo
    bins = synth.hg38_bins(res)
    paths, coos, cells = [], [], 0
    for i, (_, n) in enumerate(bins):
        dense, seps = synth.device_dense_map([n], seed=1000 * rank + 100 + i)
        coos.append(dense_to_coo(dense))
        paths.append(HotPath(seps, res, k_eig=3))
        cells += n * n
        del dense
    pool = HotPathPool(paths, workers=workers)
    n_local = [args.c3_samples // world + (1 if r < args.c3_samples % world else 0) for r in range(world)][rank]
    for _ in range(2):
        pool.run_device(coos)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n_local):
        out = pool.run_device(coos)
        w_last = out[0][0].cpu()                                  # the sample's results are read back
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b))
    pool.close()
    del pool, paths, coos
    torch.cuda.empty_cache()
    return {"workload": f"{args.c3_samples} samples x {len(bins)} per-chromosome maps at {res // 1000} kb "
                        f"({cells:.3g} cells per sample), sharded by sample over {world} GPU(s)",
            "samples": args.c3_samples, "ms_total": ms, "samples_per_s": args.c3_samples / (ms * 1e-3),
            "ms_per_sample_per_gpu": ms / max(1, -(-args.c3_samples // world)),
            "cells_per_s": args.c3_samples * cells / (ms * 1e-3), "host_threads_per_gpu": workers}


def run_row_sharded(args, world, rank, local, pk, steps=2):
    """N > 1: ONE fixed genome-wide map row-sharded over the N ranks (BASELINE.json config C5 at 4 / 8
    GPUs: 25 kb, 123 541 bins; 50 kb, 61 775 bins at 2 GPUs): synthetic count rows -> row-sharded
    O/E + log (all-reduce of the diagonal statistics) -> Pearson by the symmetric ring over peer
    memory (copy-engine plane pulls + P2P stores of the mirrored blocks) -> top-3 eigenvectors
    (all-gather of the n x 8 panel per Lanczos step). Strong scaling: total work is fixed."""
    import torch
    import torch.distributed as dist
    from hiarch_b200 import synth
    from hiarch_b200.sharded import RowShardedPath

    res = args.row_sharded_res or (50000 if world == 2 else 25000)
    lens = [nb for _, nb in synth.hg38_bins(res)]
    n = sum(lens)
    dev = torch.device("cuda", local)

    def mx(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    path_kind, ipc_error = "ring over peer memory (CUDA IPC + copy-engine pulls + P2P mirrored stores)", None
    try:
        path = RowShardedPath(synth.chr_seps_from_lens(lens), res, k_eig=3, ring=True)
        ok = 1
    except Exception as e:                                   # peer mapping unavailable on this box
        ipc_error, ok = repr(e)[:300], 0
    t = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if int(t.item()) == 0:
        path_kind = "NCCL all-gather of the planes (peer mapping unavailable)"
        path = RowShardedPath(synth.chr_seps_from_lens(lens), res, k_eig=3, ring=False)
    counts, _ = synth.device_count_rows(lens, path.row0, path.rows, 5, dev)
    oe_rows = torch.empty((max(path.rows, 1), counts.stride(0)), dtype=torch.float32, device=dev)[:path.rows, :n]
    ms = {"oe_log": [], "similarity": [], "eig": [], "total": []}
    info = None
    for it in range(steps + 1):                              # first iteration = warm-up
        barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        oe_marks = []
        path.normalise(counts, out_rows=oe_rows, marks=oe_marks)
        e[1].record()
        path.similarity(timing=path.ring)
        e[2].record()
        w, v, info = path.eigen()
        e[3].record()
        barrier()
        if it > 0:
            for key, (i, j) in (("oe_log", (0, 1)), ("similarity", (1, 2)), ("eig", (2, 3)), ("total", (0, 3))):
                ms[key].append(mx(e[i].elapsed_time(e[j])))
    eig_phases = {}
    path.eigen(phases=eig_phases)                            # one more, untimed, with per-phase events
    barrier()
    ring_stats = path.sim.collect_stats() if path.ring else {}
    oe_phases = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(oe_marks[:-1], oe_marks[1:])}
    err, cnt = path.spot_check_fp64(n_own=16, n_other=64)
    # eigen residual of the sharded result: || C[rows, :] u - lambda u[rows] ||, all-reduced
    vv = v.double()
    lam = w.double()
    cu = torch.cat([path.sim_rows[i:i + 2048].double() @ vv for i in range(0, max(path.rows, 1), 2048)])
    r2 = ((cu - vv[path.row0:path.row0 + path.rows] * lam[None, :]) ** 2).sum(dim=0)
    dist.all_reduce(r2)
    eig_res = float((r2.sqrt() / lam[0].abs()).max())
    exposed = mx(ring_stats.get("exposed_wait_ms", 0.0))
    sim_ms = float(np.mean(ms["similarity"]))
    total_ms = float(np.mean(ms["total"]))
    blocks = ring_stats.get("block_products", float(world))
    flop_exec = 3 * 2.0 * blocks * path.per * path.per * n          # executed MMA flops of this rank
    rec = {"workload": f"genome-wide human {res // 1000} kb ({n} bins, {float(n) * n:.3g} cells) row-sharded over "
                       f"{world} GPUs: count rows -> O/E+log -> Pearson -> top-3 eigvec",
           "scaling": "strong", "bins": n, "rows_per_rank": path.per, "path": path_kind, "peer_mapping_error": ipc_error,
           "ms": {k: float(np.mean(v_)) for k, v_ in ms.items()},
           "oe_log_phases_ms_rank0": oe_phases,
           "cells_per_s": float(n) * n / (total_ms * 1e-3),
           "similarity": {"block_products_per_rank": blocks, "block_ms": ring_stats.get("block_ms"),
                          "executed_mma_tflops_per_gpu": flop_exec / (sim_ms * 1e-3) / 1e12,
                          "plane_bytes_pulled_per_rank": ring_stats.get("pulled_bytes"),
                          "p2p_store_bytes_per_rank": ring_stats.get("p2p_store_bytes"),
                          "exposed_comm_ms": exposed,
                          "note": "exposed_comm_ms = time the compute stream waited for a plane pull (max over "
                                  "ranks, last timed step); the pulls run on copy engines behind the previous block"},
           "fp64_spot_check": {"entries": cnt, "max_abs_err": err, "tolerance": 1e-5, "within_tolerance": bool(err <= 1e-5)},
           "eig": {"eigvals": [float(x) for x in w.cpu().numpy()], "cycles": info, "rel_residual_max": eig_res,
                   "phases_ms_rank0": {k: (round(v_, 3) if isinstance(v_, float) else v_) for k, v_ in eig_phases.items()}}}
    # reported, not asserted: a sampled entry over the budget must show up in the record instead of taking the
    # whole line down (measured 2.4e-6 at 25 kb; tests/test_gpu_scale.py asserts the hardest block on one GPU)
    if err > 1e-5 and rank == 0:
        print(f"PARITY WARNING: row-sharded fp64 spot check {err:.3e} > 1e-5 ({cnt} entries)", file=sys.stderr)
    return rec

# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hiarch_b200 import _lib, ops, synth
    from hiarch_b200.pipeline import HotPath, dense_to_coo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    pk, pk_src = peaks()
    if args.only_row_sharded and world > 1:
        rec = run_row_sharded(args, world, rank, local, pk)
        if rank == 0:
            print(json.dumps({"row_sharded": rec}), flush=True)
        dist.destroy_process_group()
        return

    bins = synth.hg38_bins(args.res)[: args.chroms]
    lens = [n for _, n in bins]
    dense, seps = synth.device_dense_map(lens, seed=rank)        # one sample per rank
    n = dense.shape[0]
    rows, cols, vals = dense_to_coo(dense)
    del dense
    torch.cuda.empty_cache()
    nnz = rows.numel()
    prec = ops.PREC_3XF16 if args.precision == "f16x3" else ops.PREC_3XTF32
    hp = HotPath(seps, args.res, k_eig=3, max_nnz=nnz, precision=prec)
    rows_h = torch.empty(nnz, dtype=torch.int32).pin_memory().copy_(rows)
    cols_h = torch.empty(nnz, dtype=torch.int32).pin_memory().copy_(cols)
    vals_h = torch.empty(nnz, dtype=torch.float32).pin_memory().copy_(vals)
    cells = n * n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm ---------------------------------------------------------------
    for _ in range(args.warmup):
        hp.run_device(rows, cols, vals)
    barrier()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    launches0 = lib.hia_launch_count()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    stage_ev = []
    profiling = os.environ.get("HIA_BENCH_PROFILE") == "1"    # `ncu --profile-from-start off`: the timed steps only
    if profiling:
        torch.cuda.profiler.start()
    for _ in range(args.steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        e[0].record()
        hp.scatter(rows, cols, vals)
        e[1].record()
        hp.normalise()
        e[2].record()
        hp.similarity_prep()
        e[3].record()
        hp.similarity_gemm()                       # the dominant kernel, timed on its own stream
        e[4].record()
        w, vecs = hp.eigen()
        e[5].record()
        stage_ev.append(e)
    t_end.record()
    if profiling:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    launches = lib.hia_launch_count() - launches0
    ms_total = max_over_ranks(t_start.elapsed_time(t_end))
    ms_step = ms_total / args.steps
    stage = [float(np.mean([e[i].elapsed_time(e[i + 1]) for e in stage_ev])) for i in range(5)]
    scat, oe_ms, prep, gemm, eig_ms = stage
    clk = clocks.stop() if rank == 0 else None

    # ---- end-to-end arm (pinned host CSR -> eigenpairs on host) --------------------------------
    # HotPath.run_host_many: the public streaming call; every step copies ITS inputs H2D from pinned
    # memory and reads its eigenpairs back inside the timed region; the copy of step i + 1 runs on a
    # second stream while step i computes (double buffering). The sample travels as CSR -- what the
    # reference's SpsHiCMtx holds (scipy CSR) -- with 16-bit column indices when n <= 65 536: 6 bytes per
    # entry instead of the 12 of the COO triple (round 1: 2.69 GB per step; all ranks of a box pull from
    # the same host memory, which is what bent the 1 -> 8 GPU e2e curve).
    from hiarch_b200.pipeline import CsrSample
    csr_h = CsrSample(*ops.coo_to_csr_host(rows_h.numpy(), cols_h.numpy(), vals_h.numpy(), n))
    h2d_bytes = hp.sample_bytes(csr_h)
    for _ in hp.run_host_many([csr_h] * min(args.warmup, 2)):
        pass
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_stats = {}
    for w_h, v_h in hp.run_host_many([csr_h] * args.steps, stats=e2e_stats):
        pass
    t1.record()
    barrier()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1)) / args.steps
    e2e_parts = {k: max_over_ranks(float(np.mean(v[1:] if len(v) > 1 else v))) for k, v in e2e_stats.items()}
    # the same without overlap (one blocking call per sample), for the record
    t0s = torch.cuda.Event(enable_timing=True)
    t1s = torch.cuda.Event(enable_timing=True)
    t0s.record()
    hp.run_host(csr_h)
    t1s.record()
    barrier()
    e2e_serial_ms = max_over_ranks(t0s.elapsed_time(t1s))
    # H2D alone, all ranks at once (they share the host memory system): the slowest rank's GB/s
    t0c = torch.cuda.Event(enable_timing=True)
    t1c = torch.cuda.Event(enable_timing=True)
    barrier()
    t0c.record()
    hp._upload(csr_h, (hp.d_rows, hp.d_cols, hp.d_vals, hp.d_rowptr))
    t1c.record()
    barrier()
    h2d_alone_ms = max_over_ranks(t0c.elapsed_time(t1c))
    # ... and the COO triple through the same call (round 1's wire format), for comparison
    t0o = torch.cuda.Event(enable_timing=True)
    t1o = torch.cuda.Event(enable_timing=True)
    for _ in hp.run_host_many([(rows_h, cols_h, vals_h)] * 2):
        pass
    barrier()
    t0o.record()
    for _ in hp.run_host_many([(rows_h, cols_h, vals_h)] * args.steps):
        pass
    t1o.record()
    barrier()
    e2e_coo_ms = max_over_ranks(t0o.elapsed_time(t1o)) / args.steps

    # ---- in-run parity checks on the state of the last step (torch fp64 on the device; the oracle
    # comparisons proper live in tests/): a bench line of a wrong result is worthless
    checks = parity_checks(hp, rows, cols, vals, w_h, v_h) if not args.no_checks else None

    # passes over the matrix of one eigen solve (untimed repeat of the last step's solve, same state): every
    # Lanczos step of a cycle plus its true-residual pass reads the 4 B/cell correlation map once
    from hiarch_b200 import ops as _ops
    _, _, eig_info = _ops.topk_eig(hp.sim, k=hp.k_eig, block=hp.eig_block, steps=hp.eig_steps, return_info=True,
                                   workspace=hp._eig_ws)
    eig_passes = int(sum(used + 1 for _, _, used in eig_info))

    # ---- config C3: 64 samples x 24 per-chromosome maps at 50 kb, sharded by sample over the ranks -----
    c3 = run_c3(args, world, rank, max_over_ranks, barrier) if not args.no_c3 else None

    # ---- the row-sharded genome-wide chain (the one path with an exchange step), all ranks ------
    row_sharded = None
    if world > 1 and not args.no_row_sharded:
        del hp, rows, cols, vals, rows_h, cols_h, vals_h, csr_h
        torch.cuda.empty_cache()
        row_sharded = run_row_sharded(args, world, rank, local, pk)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    plane_elem_bytes = 2 if args.precision == "f16x3" else 4

    # ---- roofline of the dominant kernel (syrk_split3_kernel) ----------------------------------
    # algorithmic flops per launch: N^2 * K multiply-adds counted as the symmetric half,
    # i.e. 2 * N^2 * K / 2 = N^3 logical flops; each logical flop costs 3 MMAs
    # (hi.hi + hi.lo + lo.hi) at the fp16 (= measured bf16) or TF32 (= bf16 / 2) dense rate.
    logical = float(n) ** 3
    achieved = logical / (gemm * 1e-3) / 1e12
    rate_div = 1.0 if args.precision == "f16x3" else 2.0
    peak_logical = pk["bf16_tflops_sustained"] / rate_div / 3.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", f"syrk_traffic_{args.precision}.json")
    if os.path.exists(tpath) and n == 30894:            # the capture was taken on this workload
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
    kind = "fp16" if args.precision == "f16x3" else "TF32"
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_logical, "unit": "TFLOP/s",
                "frac": achieved / peak_logical, "traffic": traffic,
                "traffic_source": (f"profiles/{os.path.basename(tpath)}: dram__bytes_read.sum + dram__bytes_write.sum of "
                                   "one ncu --set full capture of this kernel on this workload (not an in-run counter)")
                if traffic is not None else None,
                "frac_of_burst_peak": achieved / (pk["bf16_tflops"] / rate_div / 3.0),
                "kernel": f"syrk_split3_kernel<{args.precision}>", "ms_per_launch": gemm,
                "note": ("achieved = N^3 logical flops (upper-triangle SYRK) / CUDA-event time of the "
                         f"kernel; peak = bf16_tflops_sustained{'' if rate_div == 1 else '/2'} ({kind} dense) "
                         f"/3 (3-term split), {pk_src}; executed {kind} rate = {3 * achieved:.1f} TFLOP/s")}

    cpu_baseline, same_sample = None, None
    if not args.no_cpu_baseline:
        d_cpu, seps_cpu, chroms_cpu = cpu_sample(args)
        with all_host_cores():
            cores = blas_threads()
            t0c = time.perf_counter()
            cpu_step(d_cpu, seps_cpu, args.res)
            dtc = time.perf_counter() - t0c
        nc = d_cpu.shape[0]
        # the GPU arm on the SAME bounded sample (device-resident and end to end): one like-for-like pair
        same_sample = gpu_on_sample(args, d_cpu, seps_cpu, prec)
        same_sample.update({"bins": nc, "chroms": "+".join(chroms_cpu), "cpu_cells_per_s": nc * nc / dtc,
                            "cpu_cores": cores})
        cpu_baseline = {"value": nc * nc / dtc, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": (f"{'+'.join(chroms_cpu)} at {args.res // 1000} kb = {nc} bins, "
                                   f"{dtc:.1f} s: oracle port (numpy restatement of the reference; "
                                   "O/E numpy per-diagonal loops 1 thread, Pearson GEMM on all BLAS threads, ARPACK eigsh(k=3))")}

    line = {
        "metric": METRIC, "value": world * cells / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": f"{args.precision} (fp32 storage, fp64 reductions)",
        "data": "synthetic",
        "config": config_of(args, n, len(lens)), "nnz_upper": nnz,
        "clocks": clk, "gpu_launches": int(launches),
        "e2e": {"value": world * cells / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 3 * 8 + 3 * n * 4,
                "api": "HotPath.run_host_many on CsrSample (H2D of step i+1 overlaps the compute of step i)",
                "wire_format": f"CSR: int64 row_ptr + {'uint16' if n <= 65536 else 'int32'} cols + fp32 vals "
                               f"({h2d_bytes / nnz:.2f} B/entry)",
                "ms_single_blocking_call": e2e_serial_ms,
                "in_loop_ms_slowest_rank": e2e_parts,       # H2D of the next sample / wait for it / compute, per step
                "h2d_alone_ms_slowest_rank": h2d_alone_ms,
                "h2d_gbs_per_rank_all_ranks_copying": h2d_bytes / h2d_alone_ms / 1e6,
                "coo_triple": {"ms_per_step": e2e_coo_ms, "h2d_bytes_per_step": 12 * nnz,
                               "value": world * cells / (e2e_coo_ms * 1e-3)}},
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "breakdown_ms": {"scatter": scat, "oe_log": oe_ms, "similarity_prep": prep, "similarity_gemm": gemm,
                         "topk_eig": eig_ms},
        "hbm_kernels": {                     # achieved algorithmic GB/s of the HBM-bound stages (SURVEY 8d bytes)
            "scatter": {"bytes": 12 * nnz + 4 * cells, "gbs": (12 * nnz + 4 * cells) / scat / 1e6,
                        "frac": (12 * nnz + 4 * cells) / scat / 1e6 / pk["hbm_gbs"]},
            "oe_log": {"bytes": 10 * cells, "gbs": 10 * cells / oe_ms / 1e6, "frac": 10 * cells / oe_ms / 1e6 / pk["hbm_gbs"]},
            "similarity_prep": {"bytes": (4 + 2 * plane_elem_bytes) * cells,
                                "gbs": (4 + 2 * plane_elem_bytes) * cells / prep / 1e6,
                                "frac": (4 + 2 * plane_elem_bytes) * cells / prep / 1e6 / pk["hbm_gbs"]},
            "topk_eig": {"passes": eig_passes, "cycles": [list(c) for c in eig_info], "bytes": 4 * cells * eig_passes,
                         "gbs": 4 * cells * eig_passes / eig_ms / 1e6,
                         "frac": 4 * cells * eig_passes / eig_ms / 1e6 / pk["hbm_gbs"],
                         "note": "passes x 4 B/cell over the WHOLE stage time (step kernels, projected solve, "
                                 "read-backs included); the pass kernel alone: profiles/launches_r2_bench_summary.txt"},
            "all_three": {"bytes": 12 * nnz + (4 + 10 + 4 + 2 * plane_elem_bytes) * cells,
                          "gbs": (12 * nnz + (18 + 2 * plane_elem_bytes) * cells) / (scat + oe_ms + prep) / 1e6,
                          "frac": (12 * nnz + (18 + 2 * plane_elem_bytes) * cells) / (scat + oe_ms + prep) / 1e6
                          / pk["hbm_gbs"]}},
        "eigvals": [float(x) for x in w_h],
        "checks": checks,
        "same_sample": same_sample,
        "c3": c3,
    }
    if row_sharded is not None:
        line["row_sharded"] = row_sharded
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
