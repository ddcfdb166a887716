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
from hiarch_b200.sharded import ShardedSimilarity, topk_eig_rows


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=25000)
    ap.add_argument("--bins", type=int, default=0, help="override the bin count")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--no-eig", action="store_true")
    args = ap.parse_args()
    rank, local, world = D.init("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = args.bins or sum(nb for _, nb in synth.hg38_bins(args.res))
    sh = ShardedSimilarity(n, n, flavour=ops.SIM_PEARSON)
    x = make_rows(n, sh.row0, sh.rows, 5, dev)
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
        a = (sh.hi[sh.row0 + i].double() + sh.lo[sh.row0 + i].double())
        b = (sh.hi[j].double() + sh.lo[j].double())
        ref = float(a @ b)
        ref = 1.0 if (sh.row0 + i) == j else max(-1.0, min(1.0, ref))
        errs.append(abs(float(out[i, j]) - ref))
    err = D.max_over_ranks(max(errs) if errs else 0.0)
    if rank == 0:
        sim = float(np.mean(times["similarity"]))
        eig = float(np.mean(times["eig"]))
        cells = float(n) * n
        rec = {"workload": f"row-sharded Pearson + top-3 eigvec, {n} bins ({cells:.3g} cells), {world} GPUs",
               "n_gpus": world, "bins": n, "rows_per_rank": sh.per,
               "similarity_ms": sim, "eig_ms": eig,
               "cells_per_s": cells / ((sim + eig) * 1e-3),
               "executed_tf32_tflops_per_gpu": 3 * 2 * float(sh.per) * n * n / (sim * 1e-3) / 1e12,
               "allgather_bytes_per_rank": 2 * (world - 1) * sh.per * sh.kpad * 4,
               "max_abs_err_vs_fp64_spotcheck": err}
        if not args.no_eig:
            rec["eigvals"] = [float(t) for t in w.cpu().numpy()]
            rec["eig_info"] = info
        print(json.dumps(rec), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/sharded_{world}gpu_{n}.json", "w") as fh:
            json.dump(rec, fh, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
