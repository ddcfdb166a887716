"""Worker for tests/test_gpu_multi.py: run under torchrun with one rank per GPU."""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from hiarch_b200 import dist as D
from hiarch_b200 import ops
from hiarch_b200.sharded import (RingSimilarity, RowShardedPath, ShardedSimilarity, topk_eig_rows, obs_d_exp_rows,
                                 rows_per_rank)
from hiarch_b200 import synth


def main():
    rank, local, world = D.init("nccl")
    torch.cuda.set_device(local)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    g = torch.Generator(device="cuda")
    g.manual_seed(7)                                    # same map on every rank
    base = torch.randn(6, n, device="cuda", generator=g)
    load = torch.randn(n, 6, device="cuda", generator=g) * torch.tensor([3., 2., 1.5, 1., .5, .5], device="cuda")
    x = ops.alloc_map(n)
    x.copy_(load @ base + 0.5 * torch.randn(n, n, device="cuda", generator=g) + 0.3)
    full = ops.similarity(x, flavour=ops.SIM_PEARSON)   # single-GPU result, computed on every rank
    sh = ShardedSimilarity(n, n, flavour=ops.SIM_PEARSON)
    mine = sh.run(x[sh.row0:sh.row0 + sh.rows])
    ref_rows = full[sh.row0:sh.row0 + sh.rows]
    sim_err = float((mine - ref_rows).abs().max()) if sh.rows else 0.0
    # the symmetric ring over peer memory (copy-engine pulls + P2P stores of the mirrored blocks), twice:
    # the second run reuses the exported buffers and must not race with readers of the first result
    ring = RingSimilarity(n, n, flavour=ops.SIM_PEARSON)
    ring_err = 0.0
    for _ in range(2):
        ring_rows = ring.run(x[ring.row0:ring.row0 + ring.rows], timing=True)
        torch.cuda.synchronize()
        ring_err = max(ring_err, float((ring_rows - ref_rows).abs().max()) if ring.rows else 0.0)
    ring_stats = ring.collect_stats()
    ring_sym = float((ring_rows[:, ring.row0:ring.row0 + ring.rows] -
                      ring_rows[:, ring.row0:ring.row0 + ring.rows].t()).abs().max()) if ring.rows else 0.0
    w1, v1, _ = ops.topk_eig(full, k=3, return_info=True)
    w2, v2, info = topk_eig_rows(mine, n, k=3)
    cos = [float(abs((v1[:, j].double() @ v2[:, j].double())) / (v1[:, j].double().norm() * v2[:, j].double().norm()))
           for j in range(3)]
    # forced restarts + in-cycle convergence checks (short cycles, unreachable tolerance): the ranks
    # must keep taking the same branches, and the result must still match the single-GPU solver
    w3, v3, info3 = ops.topk_eig(full, k=3, steps=6, tol=1e-14, angle_tol=0.0, max_restarts=3, return_info=True)
    w4, v4, info4 = topk_eig_rows(mine, n, k=3, steps=6, tol=1e-14, angle_tol=0.0, max_restarts=3)
    restart_rel = float(((w3 - w4).abs() / w3.abs()).max())
    restart_cycles = [len(info3), len(info4)]
    # row-sharded O/E + log against the single-GPU passes, on a multi-chromosome count map
    lens = [n // 2, n // 3, n - n // 2 - n // 3]
    lam, seps = synth.rate_matrix(lens, seed=3)
    up = synth.sample_counts(lam, seed=3)
    counts = ops.alloc_map(n)
    counts.copy_(torch.from_numpy(up + np.triu(up, 1).T).to("cuda", torch.float32))
    layout = ops.GenomeLayout(seps)
    ref_oe, _ = ops.obs_d_exp(counts, layout, 100000, log=True)
    per = rows_per_rank(n, world)
    r0 = rank * per
    nr = max(0, min(n, r0 + per) - r0)
    got_oe, _ = obs_d_exp_rows(counts[r0:r0 + nr], layout, 100000, r0)
    oe_err = float((got_oe - ref_oe[r0:r0 + nr]).abs().max()) if nr else 0.0
    # the whole chain from count rows (O/E -> ring -> eigen) + its in-run fp64 spot check
    path = RowShardedPath(seps, 100000, k_eig=3)
    w5, v5, info5 = path.run(counts[r0:r0 + nr])
    chain_err, chain_cnt = path.spot_check_fp64()
    full_oe = ops.similarity(ref_oe, flavour=ops.SIM_PEARSON)
    w6, v6, _ = ops.topk_eig(full_oe, k=3, return_info=True)
    chain_eig_rel = float(((w5 - w6).abs() / w6.abs()).max())
    chain_cos = [float(abs((v5[:, j].double() @ v6[:, j].double())) / (v5[:, j].double().norm() * v6[:, j].double().norm()))
                 for j in range(3)]
    rec = {"rank": rank, "ring_sym": ring_sym, "ring_blocks": ring_stats["block_products"],
           "chain_err": chain_err, "chain_cnt": chain_cnt, "chain_eig_rel": chain_eig_rel, "chain_cos": chain_cos, "world": world, "rows": [sh.row0, sh.rows], "sim_err": sim_err, "oe_err": oe_err,
           "eig_rel": float(((w1 - w2).abs() / w1.abs()).max()), "cos": cos, "info": info,
           "restart_rel": restart_rel, "restart_cycles": restart_cycles, "ring_err": ring_err}
    D.barrier()
    sys.stdout.write("REC " + json.dumps(rec) + "\n")
    sys.stdout.flush()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
