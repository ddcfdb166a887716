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
from hiarch_b200.sharded import ShardedSimilarity, topk_eig_rows


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
    w1, v1, _ = ops.topk_eig(full, k=3, return_info=True)
    w2, v2, info = topk_eig_rows(mine, n, k=3)
    cos = [float(abs((v1[:, j].double() @ v2[:, j].double())) / (v1[:, j].double().norm() * v2[:, j].double().norm()))
           for j in range(3)]
    rec = {"rank": rank, "world": world, "rows": [sh.row0, sh.rows], "sim_err": sim_err,
           "eig_rel": float(((w1 - w2).abs() / w1.abs()).max()), "cos": cos, "info": info}
    D.barrier()
    sys.stdout.write("REC " + json.dumps(rec) + "\n")
    sys.stdout.flush()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
