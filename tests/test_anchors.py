"""a12 host logic: the anchor peak search must reproduce the reference's tables exactly
(integer columns bit-exact, float columns to 1e-12, the file text byte for byte)."""
import numpy as np
import pytest

from hiarch_b200 import anchors as A
from hiarch_b200 import synth


def _genome(g):
    lens = [int(x) for x in g["in.lens"]]
    names = [str(x) for x in g["in.names"]]
    bs = int(g["in.bin_size"])
    seps = synth.chr_seps_from_lens(lens)
    starts = np.concatenate([np.arange(L) * bs for L in lens])
    ends = starts + bs
    return names, seps, starts, ends


def _check(table, g, method):
    cols = [str(c) for c in g[f"anchors.{method}.columns"]]
    assert table.columns == cols
    n_rows = len(g[f"anchors.{method}.cen"]) if f"anchors.{method}.cen" in g else 0
    assert len(table.rows) == n_rows
    for c in cols:
        ref = g[f"anchors.{method}.{c}"]
        got = table.column(c)
        if c in ("type", "chr"):
            assert [str(x) for x in ref] == got, c
        elif c in ("rs", "re", "cen", "cs", "ce"):
            assert np.array_equal(np.asarray(got, dtype=np.int64), ref.astype(np.int64)), c   # bit-exact
        else:
            assert np.allclose(np.asarray(got, dtype=np.float64), ref, rtol=1e-12, atol=0, equal_nan=True), c


def test_inter_rowsum_intra_low_no_anchor_tables(golden):
    g = golden("g3chr")
    names, seps, starts, ends = _genome(g)
    t = A.find_anchors_by_rowsum(g["inter_rowsum_smooth"], names, seps, starts, ends)
    _check(t, g, "inter_rowsum")
    assert t.to_text() == str(g["anchors.inter_rowsum.text"])        # the file, byte for byte
    assert any(r["type"] == "used" for r in t.rows)
    t2 = A.find_anchors_by_rowsum(g["intra_rowsum_smooth"], names, seps, starts, ends)
    _check(t2, g, "intra_low")
    assert t2.to_text() == str(g["anchors.intra_low.text"])
    t3 = A.remove_all_used(A.find_anchors_by_rowsum(g["inter_rowsum_smooth"], names, seps, starts, ends))
    _check(t3, g, "no_anchor")
    assert not any(r["type"] == "used" for r in t3.rows)


def test_intra_x_and_single_chromosome_tables(golden):
    g = golden("g1chr")
    names, seps, starts, ends = _genome(g)
    t = A.find_anchors_by_intra_x(g["intra_x"], names, seps, starts, ends)
    _check(t, g, "intra_x")
    assert t.to_text() == str(g["anchors.intra_x.text"])
    # a single-chromosome genome has an all-zero inter profile -> empty table
    t = A.find_anchors_by_rowsum(g["inter_rowsum_smooth"], names, seps, starts, ends)
    assert len(t.rows) == 0 and np.all(g["inter_rowsum_smooth"] == 0)


def test_change_anchor_and_kernal_len():
    assert A.get_kernal_len(593, .1, 5, True) == 59 and A.get_kernal_len(30, .1, 5, True) == 5
    assert A.parse_anchor_change("1:2,3:0") == [(1, 2), (3, 0)]
    rows = [dict(chr="a", start=5.0, type="used"), dict(chr="a", start=1.0, type="normal"),
            dict(chr="b", start=0.0, type="used")]
    t = A.change_anchor(A.AnchorTable(rows, ["chr", "start", "type"]), ["a", "b"], [(1, 1), (2, 0)])
    assert [r["type"] for r in t.rows] == ["normal", "used", "normal"]


@pytest.mark.parametrize("seed", range(6))
def test_peak_logic_invariants_on_random_profiles(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 300))
    a = np.convolve(rng.random(n + 20), np.ones(21) / 21, mode="valid")[:n] - 0.3
    regions = A.get_anchor_regions(a)
    for r in regions:
        assert r["rs"] <= r["cen"] <= r["re"] and a[r["cen"]] > 0
    cens = [r["cen"] for r in regions]
    assert cens == sorted(cens)
