"""CPU: the N>1 plumbing with world_size-2 gloo processes, and the sharding rules."""
import os
import subprocess
import sys

import numpy as np

from hiarch_b200 import dist as D

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
from hiarch_b200 import dist as D
rank, local, world = D.init("gloo")
assert world == 2
weights = [9, 1, 4, 4, 16, 2, 2]
mine = D.shard_for_rank(weights, world, rank)
D.barrier()
slow = D.max_over_ranks(10.0 + rank)          # the slowest rank defines the step time
total = D.sum_over_ranks(sum(weights[i] for i in mine))
sys.stdout.write("REC " + json.dumps({"rank": rank, "mine": mine, "slow": slow, "total": total}) + "\n"); sys.stdout.flush()
'''


def test_two_rank_gloo_sharding_and_reductions(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    import re
    recs = [json.loads(m) for m in re.findall(r"REC (\{.*?\})(?=\s|REC|$)", out.stdout)]
    assert len(recs) == 2
    items = sorted(recs[0]["mine"] + recs[1]["mine"])
    assert items == list(range(7))                              # disjoint + complete
    assert all(r["slow"] == 11.0 for r in recs)                 # max over ranks
    assert all(r["total"] == 38.0 for r in recs)
    loads = [sum([9, 1, 4, 4, 16, 2, 2][i] for i in r["mine"]) for r in recs]
    assert abs(loads[0] - loads[1]) <= 6                        # largest-first balance


def test_yield_mp_idxes_matches_reference_rule():
    # publish/run_hic_dataset.py:259-269: shards of int(size/thread); remainder -> extra shards
    assert [list(x) for x in D.yield_mp_idxes(7, 3)] == [[0, 1], [2, 3], [4, 5], [6]]
    assert [list(x) for x in D.yield_mp_idxes(2, 5)] == [[0], [1]]
    assert [list(x) for x in D.yield_mp_idxes(6, 3)] == [[0, 1], [2, 3], [4, 5]]


def test_shard_for_rank_single():
    assert D.shard_for_rank([3, 1, 2], 1, 0) == [0, 1, 2]


def test_parallel_func_sharding_without_gpu(tmp_path):
    from hiarch_b200.run_hic_dataset import ParallelFunc, CaptureParas
    files = {f"cell{i}": (f"/x/cell{i}.mtx", "/x/w.bed") for i in range(5)}
    pf = ParallelFunc("sp", "ds", lambda m: 0, mtx_files=files, exclude_pats="cell3")
    assert pf.file_names == ["cell0", "cell1", "cell2", "cell4"]
    assert pf.shards(2) == [["cell0", "cell1"], ["cell2", "cell4"]]
    cap = CaptureParas()
    cap.write("#para-near_num: 12\nnoise\n")
    assert cap.paras == {"near_num": "12"}


def test_ring_schedule_covers_every_block_pair_once():
    from hiarch_b200.sharded import ring_schedule
    for world in range(1, 10):
        steps = ring_schedule(world)
        assert len(steps) == world // 2 + 1
        pairs = [tuple(sorted(p)) for st in steps for p in st]
        assert sorted(pairs) == sorted((a, b) for a in range(world) for b in range(a, world))   # once each
        per_rank = [sum(1 for st in steps for (r, _) in st if r == k) for k in range(world)]
        assert max(per_rank) - min(per_rank) <= 1                                             # balanced
        assert max(per_rank) == world // 2 + 1


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _run_ring(world, n, k, port=None):
    import json
    import re
    port = port or _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "ring_cpu_worker.py"), str(n), str(k)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    recs = [json.loads(m) for m in re.findall(r"REC (\{.*?\})(?=\s|REC|$)", out.stdout)]
    assert len(recs) == world
    return sorted(recs, key=lambda r: r["rank"])


def test_ring_similarity_exchange_logic_on_gloo():
    """RingSimilarity's schedule / plane exchange / mirrored-block return with a torch stand-in for
    the kernels: every rank ends up with exactly its rows of the full Pearson matrix (ragged and
    empty last blocks included), having contracted only world/2 + 1 blocks."""
    for world, n in ((2, 200), (3, 300), (4, 400), (4, 300)):
        recs = _run_ring(world, n, 24)
        assert sum(r["rows"] for r in recs) == n
        for r in recs:
            assert r["nan"] == 0, r                              # every column block was filled
            assert r["err"] < 1e-12, r
            assert r["blocks"] == r["expected_blocks"] and len(r["blocks"]) <= world // 2 + 1


def test_ring_similarity_single_process():
    """world = 1 (no process group): the schedule degenerates to the diagonal block."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ring_cpu_worker import CpuRing
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal((150, 40)))
    ring = CpuRing(150, 40)
    out = ring.run(x)
    xc = x - x.mean(dim=1, keepdim=True)
    xc = xc / xc.norm(dim=1, keepdim=True)
    assert ring.blocks_done == [0] and float((out - xc @ xc.t()).abs().max()) < 1e-12
