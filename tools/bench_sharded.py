"""Row-sharded genome-wide Pearson + top-3 eigenvectors (BASELINE.json config C5) across the GPUs
of one box. Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29700 tools/bench_sharded.py --res 25000

Every rank synthesises ITS rows of a normalised (O/E-like) map: low-rank compartment structure
+ noise, deterministic in (seed, global row), so no rank ever holds the 61 GB map. Timed with
CUDA events, max over ranks. Spot-checks entries against fp64 dot products of the gathered planes.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import torch.distributed as dist

from hiarch_b200 import dist as D
from hiarch_b200 import ops, synth
from hiarch_b200.sharded import RingSimilarity, ShardedSimilarity, topk_eig_rows


def make_rows(n, row0, rows, seed, device, chunk=2048):
    """rows x n block of a synthetic normalised map: sum_m beta_m c_m[i] c_m[j] + noise."""
    rng = np.random.default_rng(seed)
    comps = np.stack([np.sqrt(b) * synth.compartment_track(n, rng, max(3, n // (400 * (m + 1))))
                      for m, b in enumerate((1.0, 0.6, 0.4))], axis=1)          # [n, 3]
    comp = torch.from_numpy(comps).to(device=device, dtype=torch.float32)
    x = torch.empty((rows, ops.padded_ld(n)), dtype=torch.float32, device=device)[:, :n]
    g = torch.Generator(device=device)
    for r in range(0, rows, chunk):
        r1 = min(rows, r + chunk)
        g.manual_seed(seed * 1000003 + row0 + r)
        blk = comp[row0 + r:row0 + r1] @ comp.T
        blk += 0.8 * torch.randn((r1 - r, n), device=device, generator=g)
        x[r:r1] = blk
    return x


make_count_rows = synth.device_count_rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from-counts", action="store_true",
                    help="start from synthetic COUNT rows and run the row-sharded O/E+log first")
    ap.add_argument("--res", type=int, default=25000)
    ap.add_argument("--bins", type=int, default=0, help="override the bin count")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--no-eig", action="store_true")
    ap.add_argument("--ring", action="store_true",
                    help="also time RingSimilarity (symmetric ring schedule, plane exchange hidden behind the "
                         "contraction) and compare its rows with the all-gather path")
    args = ap.parse_args()
    # a rank that dies must not leave the others in a 10-minute NCCL watchdog wait
    os.environ.setdefault("TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC", "120")
    import datetime
    rank, local, world = D.env_world()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=120))
    dev = torch.device("cuda", local)
    lens = [nb for _, nb in synth.hg38_bins(args.res)]
    n = args.bins or sum(lens)
    oe_ms = []
    from hiarch_b200.sharded import rows_per_rank
    per = rows_per_rank(n, world)
    my_row0 = rank * per
    my_rows = max(0, min(n, my_row0 + per) - my_row0)
    if args.from_counts:
        from hiarch_b200.sharded import obs_d_exp_rows
        assert not args.bins
        # generate + normalise BEFORE the 2 x (n_pad x kpad) plane buffers exist (memory headroom)
        counts, seps = make_count_rows(lens, my_row0, my_rows, 5, dev)
        torch.cuda.empty_cache()
        layout = ops.GenomeLayout(seps, dev)
        layout.oe_tables(args.res)
        x = torch.empty((max(my_rows, 1), counts.stride(0)), dtype=torch.float32, device=dev)[:my_rows, :n]
        for it in range(2):
            D.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            obs_d_exp_rows(counts, layout, args.res, my_row0, log=True, out_rows=x)
            b.record()
            D.barrier()
            oe_ms.append(D.max_over_ranks(a.elapsed_time(b)))
        # symmetry spot check of the generator + O/E across ranks: O/E[i, j] on this rank vs [j, i]
        del counts
        torch.cuda.empty_cache()
    else:
        x = make_rows(n, my_row0, my_rows, 5, dev)
    sh = ShardedSimilarity(n, n, flavour=ops.SIM_PEARSON)
    out = torch.empty((max(sh.rows, 1), ops.padded_ld(n)), dtype=torch.float32, device=dev)[:sh.rows, :n]
    D.barrier()
    times = {"similarity": [], "eig": []}
    for it in range(args.steps + 1):                    # first iteration = warm-up
        D.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        sh.run(x, out_rows=out)
        e1.record()
        if not args.no_eig:
            w, v, info = topk_eig_rows(out, n, k=3)
        e2.record()
        D.barrier()
        if it > 0:
            times["similarity"].append(D.max_over_ranks(e0.elapsed_time(e1)))
            times["eig"].append(D.max_over_ranks(e1.elapsed_time(e2)))
    # spot check: entries of the own row block against fp64 dot products of the gathered planes
    errs = []
    rng = np.random.default_rng(rank)
    for _ in range(24):
        i = int(rng.integers(0, sh.rows)) if sh.rows else 0
        j = int(rng.integers(0, n))
        unscale = 2.0 ** -14 if sh.hi.dtype == torch.float16 else 1.0     # 3xF16 planes are scaled by 2^14
        a = (sh.hi[sh.row0 + i].double() + sh.lo[sh.row0 + i].double()) * unscale
        b = (sh.hi[j].double() + sh.lo[j].double()) * unscale
        ref = float(a @ b)
        ref = 1.0 if (sh.row0 + i) == j else max(-1.0, min(1.0, ref))
        errs.append(abs(float(out[i, j]) - ref))
    err = D.max_over_ranks(max(errs) if errs else 0.0)
    meta = {"per": sh.per, "kpad": sh.kpad, "elem": sh.hi.element_size(), "prec": sh.precision}
    ring_ms, ring_diff = None, None
    if args.ring:
        del sh.hi, sh.lo, sh.own_hi, sh.own_lo              # make room: the ring keeps its own (smaller) buffers
        torch.cuda.empty_cache()
        ring = RingSimilarity(n, n, flavour=ops.SIM_PEARSON)
        out2 = torch.empty((max(ring.rows, 1), ops.padded_ld(n)), dtype=torch.float32, device=dev)[:ring.rows, :n]
        rt = []
        for it in range(args.steps + 1):
            D.barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            ring.run(x, out_rows=out2)
            r1.record()
            D.barrier()
            if it > 0:
                rt.append(D.max_over_ranks(r0.elapsed_time(r1)))
        ring_ms = float(np.mean(rt))
        ring_diff = D.max_over_ranks(float((out2 - out).abs().max()) if ring.rows else 0.0)
    if rank == 0:
        sim = float(np.mean(times["similarity"]))
        eig = float(np.mean(times["eig"]))
        cells = float(n) * n
        rec = {"workload": f"row-sharded Pearson + top-3 eigvec, {n} bins ({cells:.3g} cells), {world} GPUs",
               "n_gpus": world, "bins": n, "rows_per_rank": meta["per"],
               "oe_log_ms": (oe_ms[-1] if oe_ms else None),
               "similarity_ms": sim, "eig_ms": eig,
               "cells_per_s": cells / ((sim + eig + (oe_ms[-1] if oe_ms else 0.0)) * 1e-3),
               "precision": {0: "tf32x3", 2: "f16x3"}[meta["prec"]],
               "executed_mma_tflops_per_gpu": 3 * 2 * float(meta["per"]) * n * n / (sim * 1e-3) / 1e12,
               "allgather_bytes_per_rank": 2 * (world - 1) * meta["per"] * meta["kpad"] * meta["elem"],
               "max_abs_err_vs_fp64_spotcheck": err,
               "ring_similarity_ms": ring_ms, "ring_vs_allgather_max_abs_diff": ring_diff}
        if not args.no_eig:
            rec["eigvals"] = [float(t) for t in w.cpu().numpy()]
            rec["eig_info"] = info
        print(json.dumps(rec), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        tag = "counts_" if args.from_counts else ""
        with open(f"gpurun_out/sharded_{tag}{world}gpu_{n}.json", "w") as fh:
            json.dump(rec, fh, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
