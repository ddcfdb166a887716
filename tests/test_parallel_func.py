"""CPU: the batch harness ParallelFunc (mirror of publish/run_hic_dataset.py:259-269, :314-356,
:573-603) -- contiguous shards, one spawned worker per shard, '#para-' capture, one result table.
The mapped function here only touches the host-side sparse matrix (no kernel calls), so the
sharding / process plumbing is covered without a GPU."""
import numpy as np
import pandas as pd
import pytest

from hiarch_b200 import synth
from hiarch_b200.dist import yield_mp_idxes
from hiarch_b200.run_hic_dataset import ParallelFunc


def count_entries(mtx, scale=1):
    """Module-level (picklable for spawn): a per-sample statistic of the parsed file."""
    m = mtx.matrix
    print(f"#para-bins: {mtx.shape[0]}")
    print("an ordinary line")
    return {"nnz": int(m.nnz) * scale, "total": float(m.sum())}


def table_result(mtx):
    return pd.DataFrame({"chr": list(mtx.row_window.chros), "len": [3] * len(mtx.row_window.chros)})


def failing(mtx):
    raise RuntimeError("boom")


@pytest.fixture(scope="module")
def samples(tmp_path_factory):
    root = tmp_path_factory.mktemp("cells")
    files = {}
    for i in range(5):
        files[f"cell{i}"] = synth.write_sample(str(root / f"cell{i}"), ["chr1", "chr2"], [30 + i, 20], 100000, seed=i)
    return files


def test_yield_mp_idxes_contiguous_and_complete():
    for n, t in ((7, 2), (5, 5), (3, 8), (10, 3), (1, 1)):
        shards = [list(s) for s in yield_mp_idxes(n, t)]
        flat = [i for s in shards for i in s]
        assert flat == list(range(n))                            # contiguous, ordered, complete
        per = int(n / t) if n > t else 1                         # the reference's rule: the remainder
        assert [len(s) for s in shards] == [min(per, n - i) for i in range(0, n, per)]   # makes EXTRA shards


def test_single_worker_table_and_captured_paras(samples, tmp_path):
    pf = ParallelFunc("synth", "ds", count_entries, mtx_files=samples)
    pf.update_func_kwargs({"scale": 2})
    out = tmp_path / "res.tsv"
    t = pf.vali_parallel(str(out), thread=1)
    assert list(t["cell"]) == sorted(samples)
    assert (t["nnz"] > 0).all() and (t["nnz"] % 2 == 0).all()
    assert list(t["bins"].astype(int)) == [50 + i for i in range(5)]      # harvested from '#para-bins:'
    back = pd.read_table(out)
    assert list(back.columns) == list(t.columns) and len(back) == 5


def test_two_spawned_workers_give_the_same_table(samples):
    pf = ParallelFunc("synth", "ds", count_entries, mtx_files=samples)
    one = pf.vali_parallel(thread=1)
    two = pf.vali_parallel()                                      # reference default: thread=2
    pd.testing.assert_frame_equal(one, two)
    assert [len(s) for s in pf.shards(2)] == [len(list(s)) for s in yield_mp_idxes(5, 2)]


def test_dataframe_results_are_concatenated_per_cell(samples):
    pf = ParallelFunc("synth", "ds", table_result, n_cell=3, mtx_files=samples)
    t = pf.vali_parallel(thread=1, capture_paras=False)
    assert len(t) == 6 and set(t["cell"]) == {"cell0", "cell1", "cell2"}
    assert list(t.columns) == ["cell", "chr", "len"]


def test_include_exclude_patterns(samples):
    pf = ParallelFunc("synth", "ds", count_entries, include_pats=["cell1", "cell3"], exclude_pats="cell3",
                      mtx_files=samples)
    assert pf.file_names == ["cell1"]


def test_worker_failure_is_reported(samples):
    pf = ParallelFunc("synth", "ds", failing, mtx_files=samples)
    with pytest.raises(RuntimeError):
        pf.vali_parallel(thread=2)


def out_path_result(mtx, output=None):
    return {"output": output}


def test_update_para_output_survives_spawned_workers(samples, tmp_path):
    """update_para_output stored a lambda, which cannot be pickled for the spawned workers."""
    pf = ParallelFunc("synth", "ds", out_path_result, mtx_files=samples)
    pf.update_para_output(str(tmp_path / "res{1}"))
    t = pf.vali_parallel(thread=2, capture_paras=False)
    assert list(t["output"]) == [f"{tmp_path}/res{{1}}/{c}" for c in sorted(samples)]
    pf.update_func_kwargs({"output": lambda cell: cell})          # a lambda: refused up front, clearly
    with pytest.raises(TypeError, match="cannot be sent to the worker processes"):
        pf.vali_parallel(thread=2)
    assert list(pf.vali_parallel(thread=1, capture_paras=False)["output"]) == sorted(samples)   # in-process is fine
