"""World-size-N gloo worker for tests/test_dist_cpu.py: the exchange logic of
hiarch_b200.sharded.RingSimilarity (who contracts which pair of row blocks, which planes travel
where, how the mirrored blocks come back) with a torch CPU stand-in for the CUDA kernels."""
import contextlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from hiarch_b200 import sharded  # noqa: E402


class CpuRing(sharded.RingSimilarity):
    """Same schedule, planes = the centred unit rows in fp64, block product = a matmul."""

    def _alloc(self):
        self.kpad = self.k
        comb = self.per_pad + self.per
        self.hi = [torch.zeros((comb, self.kpad), dtype=torch.float64) for _ in range(2)]
        self.blocks_done = []

    def _prep(self, x_rows):
        if self.rows > 0:
            x = x_rows.double()
            x = x - x.mean(dim=1, keepdim=True)
            x = x / x.norm(dim=1, keepdim=True)
            self.hi[0][:self.rows] = x
        self.hi[1][:self.per].copy_(self.hi[0][:self.per])

    def _own_planes(self):
        return [self.hi[0][:self.per]]

    def _peer_slots(self, b):
        return [self.hi[b][self.per_pad:]]

    def _block(self, b, peer, out_rows, diagonal):
        cols = self.valid(peer)
        self.blocks_done.append(peer)
        if self.rows == 0 or cols == 0:
            return
        col0 = 0 if diagonal else self.per_pad
        a = self.hi[b][:self.rows]
        bb = self.hi[b][col0:col0 + cols]
        # store exactly like the kernel behind RingSimilarity._block: element (r, gc) of the stacked
        # problem, gc in [col0, col0 + cols), goes to base[(r - row0) * ld + gc] where `base` is the
        # output pointer shifted by (peer * per - col0) columns
        flat, ld = out_rows.view(-1), out_rows.stride(0)
        shift = peer * self.per - col0
        idx = shift + torch.arange(self.rows)[:, None] * ld + torch.arange(col0, col0 + cols)[None, :]
        assert int(idx.min()) >= 0 and int(idx.max()) < flat.numel()
        flat[idx.reshape(-1)] = (a @ bb.t()).reshape(-1)

    def _new_out(self):
        return torch.full((self.rows, self.n), float("nan"), dtype=torch.float64)

    def _side(self):
        return contextlib.nullcontext, None

    def _fence(self, stream):
        pass


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, k = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((n, k)))
    ring = CpuRing(n, k)
    out = ring.run(x[ring.row0:ring.row0 + ring.rows])
    xc = x - x.mean(dim=1, keepdim=True)
    xc = xc / xc.norm(dim=1, keepdim=True)
    ref = (xc @ xc.t())[ring.row0:ring.row0 + ring.rows]
    err = float((out - ref).abs().max()) if ring.rows else 0.0
    nan = int(torch.isnan(out).sum())
    rec = {"rank": rank, "rows": ring.rows, "err": err, "nan": nan, "blocks": ring.blocks_done,
           "expected_blocks": [s for t, pairs in enumerate(ring.steps) for (r, s) in pairs if r == rank]}
    sys.stdout.write("REC " + json.dumps(rec) + "\n")
    sys.stdout.flush()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
