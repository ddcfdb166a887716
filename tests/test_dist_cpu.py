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


def test_ring_plan_covers_every_cell_once_and_is_balanced():
    """Pure schedule: over all ranks the direct + mirrored stores of `ring_plan` write every cell of
    the n x n matrix exactly once, and every rank contracts ~world / 2 block products."""
    from hiarch_b200.sharded import ring_plan, rows_per_rank
    for world, n in [(1, 300), (2, 1000), (3, 1000), (4, 1100), (5, 900), (8, 4000), (8, 3000), (4, 300), (6, 2000),
                     (8, 4096), (2, 1024), (4, 2048), (7, 7 * 256)]:
        per = rows_per_rank(n, world)
        cover = np.zeros((n, n), dtype=np.int32)
        work = []
        for r in range(world):
            w = 0.0
            for j in ring_plan(world, r, per, n):
                i0, j0 = r * per + j["a0"], j["peer"] * per + j["b0"]
                assert j["a0"] % 128 == 0 and j["b0"] % 128 == 0
                if j["symmetric"]:
                    cover[i0:i0 + j["a_rows"], j0:j0 + j["b_rows"]] += 1
                    w += 0.5 * j["a_rows"] * j["b_rows"]
                else:
                    cover[i0:i0 + j["a_rows"], j0:j0 + j["b_rows"]] += 1
                    cover[j0:j0 + j["b_rows"], i0:i0 + j["a_rows"]] += 1
                    w += j["a_rows"] * j["b_rows"]
            work.append(w / per ** 2)
        assert cover.min() == 1 and cover.max() == 1, (world, n)
        assert max(work) <= world / 2 + 1e-9, (world, n, work)          # never more than world / 2 blocks
        if n == world * per and per % 256 == 0:                         # full blocks: perfectly balanced
            assert max(work) - min(work) <= 1e-9 and abs(work[0] - world / 2) <= 1e-9, (world, n, work)


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
    """RingSimilarity's schedule on shared-memory "peer" buffers with a torch stand-in for the
    kernels: every rank ends up with exactly its rows of the full Pearson matrix (ragged and empty
    last blocks included), every remote block was pulled before it was contracted, and no rank
    contracted more than world / 2 block products."""
    for world, n in ((2, 600), (3, 700), (4, 1100), (4, 300)):
        recs = _run_ring(world, n, 24)
        assert sum(r["rows"] for r in recs) == n
        for r in recs:
            assert r["nan"] == 0, r                              # every column block was filled
            assert r["err"] < 1e-12, r
            assert r["pulls"] == [b for b in r["blocks"] if b != r["rank"]], r
            assert r["work"] <= world / 2 * r["per"] ** 2 + 1e-9, r


def test_ring_similarity_single_process():
    """world = 1 (no process group): the schedule degenerates to the diagonal block."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ring_cpu_worker import CpuRing
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal((150, 40)))
    ring = CpuRing(150, 40)
    try:
        out = ring.run(x)
        xc = x - x.mean(dim=1, keepdim=True)
        xc = xc / xc.norm(dim=1, keepdim=True)
        assert [e[2] for e in ring.log] == [0] and float((out - xc @ xc.t()).abs().max()) < 1e-12
    finally:
        ring.cleanup()
