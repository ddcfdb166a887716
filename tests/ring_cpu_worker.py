"""World-size-N gloo worker for tests/test_dist_cpu.py: the schedule of
hiarch_b200.sharded.RingSimilarity (who contracts which pair of row blocks, which plane rows are
pulled from whom into which staging slot, where the direct and the mirrored stores land) with a
torch CPU stand-in for the CUDA side: "peer memory" = file-backed shared tensors under /dev/shm
that every rank maps, a pull = a copy out of the peer's mapping, the block kernel = an fp64
matmul that stores through the same element offsets (`block_offsets`) the CUDA backend turns
into pointers."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from hiarch_b200 import sharded  # noqa: E402


class _Ev:
    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 0.0


class _Stream:
    def wait_event(self, ev):
        pass


def _shared(tag, rank, numel, create):
    path = f"/dev/shm/hia_ring_{os.environ.get('MASTER_PORT', '0')}_{tag}_{rank}"
    if create:
        with open(path, "wb") as fh:
            fh.truncate(numel * 8)
    return torch.from_file(path, shared=True, size=numel, dtype=torch.float64), path


class CpuRing(sharded.RingSimilarity):
    """Same schedule and offsets; planes = the centred unit rows in fp64 (one plane)."""

    def _alloc(self):
        self.kpad, self.elem, self.ld = self.k, 8, self.n + 3
        self.paths = []
        own, p1 = _shared("planes", self.rank, self.per * self.kpad, True)
        out, p2 = _shared("out", self.rank, self.per * self.ld, True)
        self.paths += [p1, p2]
        own.zero_()
        out.fill_(float("nan"))
        self.own = own.view(self.per, self.kpad)
        self.out_buf = out.view(self.per, self.ld)
        self.slots = [torch.full((self.per, self.kpad), float("nan"), dtype=torch.float64) for _ in range(2)]
        if self.world > 1:
            dist.barrier()
        self.peer_planes_t, self.peer_out_t = {}, {}
        for r in range(self.world):
            self.peer_planes_t[r] = _shared("planes", r, self.per * self.kpad, False)[0].view(self.per, self.kpad)
            self.peer_out_t[r] = _shared("out", r, self.per * self.ld, False)[0]
        self.log = []

    def _prep(self, x_rows):
        if self.rows > 0:
            x = x_rows.double()
            x = x - x.mean(dim=1, keepdim=True)
            self.own[:self.rows] = x / x.norm(dim=1, keepdim=True)

    def _rendezvous(self):
        if self.world > 1:
            dist.barrier()

    def _pull(self, slot, job):
        b0, nb = job["b0"], job["b_rows"]
        self.slots[slot][b0:b0 + nb] = self.peer_planes_t[job["peer"]][b0:b0 + nb]
        self.log.append(("pull", job["t"], job["peer"], slot, b0, nb))

    def _block(self, slot, job):
        a0, na, b0, nb = job["a0"], job["a_rows"], job["b0"], job["b_rows"]
        a = self.own[a0:a0 + na]
        b = self.own[b0:b0 + nb] if job["symmetric"] else self.slots[slot][b0:b0 + nb]
        v = a @ b.t()
        assert not torch.isnan(v).any(), "a staging row was read before it was pulled"
        if job["symmetric"]:
            v.fill_diagonal_(1.0)
        out_off, t_off = self.block_offsets(job)
        ii, jj = torch.arange(na)[:, None], torch.arange(nb)[None, :]
        mine, peer = self.out_buf.view(-1), self.peer_out_t[job["peer"]]
        idx = (out_off + ii * self.ld + jj).reshape(-1)
        idx_t = (t_off + jj * self.ld + ii).reshape(-1)
        assert int(idx.min()) >= 0 and int(idx.max()) < mine.numel()
        assert int(idx_t.min()) >= 0 and int(idx_t.max()) < peer.numel()
        mine[idx] = v.reshape(-1)
        peer[idx_t] = v.reshape(-1)
        self.log.append(("block", job["t"], job["peer"], slot, na * nb))

    def _streams(self):
        return _Stream(), _Stream()

    def _event(self, timing=False):
        return _Ev()

    def cleanup(self):
        for p in self.paths:
            try:
                os.unlink(p)
            except OSError:
                pass


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, k = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((n, k)))
    ring = CpuRing(n, k)
    out = ring.run(x[ring.row0:ring.row0 + ring.rows]).clone()
    xc = x - x.mean(dim=1, keepdim=True)
    xc = xc / xc.norm(dim=1, keepdim=True)
    ref = (xc @ xc.t())[ring.row0:ring.row0 + ring.rows]
    err = float((out - ref).abs().max()) if ring.rows else 0.0
    nan = int(torch.isnan(out).sum())
    work = sum(e[4] for e in ring.log if e[0] == "block" and e[2] != rank) + \
        0.5 * sum(e[4] for e in ring.log if e[0] == "block" and e[2] == rank)
    rec = {"rank": rank, "rows": ring.rows, "err": err, "nan": nan, "per": ring.per,
           "blocks": [e[2] for e in ring.log if e[0] == "block"],
           "pulls": [e[2] for e in ring.log if e[0] == "pull"], "work": work}
    sys.stdout.write("REC " + json.dumps(rec) + "\n")
    sys.stdout.flush()
    dist.barrier()
    ring.cleanup()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
