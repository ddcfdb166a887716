"""Host-side (CPU) logic of the product: SGD class + pooling tables + the clamp formula the
finalize kernel uses, checked against the oracle and the reference's golden vectors."""
import numpy as np
import pytest

from oracle import hia_oracle as O
from hiarch_b200 import synth
from hiarch_b200.sgd import ScaledGenomicDistance, oe_tables


def expected_from_tables(M, bin_size, clamp=True):
    """numpy model of hia_oe_finalize (finalize_intra_kernel) driven by the host tables."""
    n = M.shape[0]
    pool, use = oe_tables(bin_size, n)
    up = np.array([np.trace(M, d) for d in range(n)])
    ng = pool.max() + 1
    gsum, gcnt = np.zeros(ng), np.zeros(ng)
    for d in range(n):
        if pool[d] >= 0:
            mult = 1.0 if d == 0 else 2.0
            gsum[pool[d]] += mult * up[d]
            gcnt[pool[d]] += mult * (n - d)
    x = np.zeros(n)
    for d in range(n):
        if use[d] < 0:
            x[d] = up[0] / n
        else:
            v = gsum[use[d] >> 1] / gcnt[use[d] >> 1]
            x[d] = 1.0 if ((use[d] & 1) and v == 0) else v
    if clamp:
        out = x.copy()
        pmin = np.inf
        for d in range(n):
            if pmin != np.inf and x[d] > pmin:
                out[d] = pmin
            if x[d] != 0:
                pmin = min(pmin, x[d])
        x = out
    return x


def test_sgd_class_matches_golden(golden):
    g = golden("gsgd")
    for bs, n in g["shapes"]:
        bs, n = int(bs), int(n)
        sgd = ScaledGenomicDistance(s=bs, e=bs * n, bin_size=bs, k=.2)
        look = np.array([-1 if sgd[d] is None else sgd[d] for d in range(n)], dtype=np.int32)
        assert np.array_equal(look, g[f"lookup.{bs}.{n}"]), (bs, n)


@pytest.mark.parametrize("case", ["g3chr", "g1chr", "gquirk"])
def test_tables_reproduce_expected(golden, case):
    g = golden(case)
    bs = int(g["in.bin_size"])
    n = int(g["in.lens"].sum())
    dense = O.scatter_coo_sym(g["in.rows"], g["in.cols"], g["in.vals"], n)
    for (s, e) in synth.chr_seps_from_lens([int(x) for x in g["in.lens"]]):
        blk = dense[s:e + 1, s:e + 1]
        for clamp in (True, False):
            ref = np.array(O.sgd_mean_expected(blk, bs), dtype=np.float64)
            if clamp:
                ref = np.array(O.monotone_clamp(ref))
            mine = expected_from_tables(blk, bs, clamp)
            assert np.allclose(mine, ref, rtol=1e-12, atol=0), (case, s, clamp)
    # last block's expected vector equals what the reference returned
    s, e = synth.chr_seps_from_lens([int(x) for x in g["in.lens"]])[-1]
    mine = expected_from_tables(dense[s:e + 1, s:e + 1], bs, True)
    assert np.allclose(mine, g["expected_last_block"], rtol=1e-12, atol=0)


def test_tables_on_ragged_low_depth_chromosomes(golden):
    """The numpy model of the finalize kernel on the ragged fixture (70 / 12 / 31 / 6 bins, all-zero
    diagonals): expected vectors equal the oracle's (pinned to the real reference by
    test_oracle_golden.py::test_option_and_edge_variants), and dividing by them reproduces the
    reference's O/E block."""
    g = golden("goptions")
    dense, bs = g["r.dense"], 1000000
    s = 0
    for L in [int(x) for x in g["r.lens"]]:
        blk = dense[s:s + L, s:s + L]
        for clamp in (True, False):
            ref = np.array(O.sgd_mean_expected(blk, bs), dtype=np.float64)
            if clamp:
                ref = np.array(O.monotone_clamp(ref))
            assert np.allclose(expected_from_tables(blk, bs, clamp), ref, rtol=1e-12, atol=0), (L, clamp)
        exp = expected_from_tables(blk, bs, True)
        d = np.abs(np.subtract.outer(np.arange(L), np.arange(L)))
        with np.errstate(divide="ignore", invalid="ignore"):
            oe = blk / exp[d]
        ref_oe = g["oe.ragged"][s:s + L, s:s + L]
        ok = np.isfinite(oe)
        assert np.allclose(oe[ok], ref_oe[ok], rtol=1e-12, atol=0)
        s += L


def test_clamp_formula_on_adversarial_vectors():
    rng = np.random.default_rng(0)
    for _ in range(200):
        x = rng.choice([0.0, 1.0, 2.0, 0.5, 3.0, -1.0, 0.25], size=rng.integers(1, 30))
        ref = np.array(O.monotone_clamp(x))
        out = x.copy()
        pmin = np.inf
        for d in range(len(x)):
            if pmin != np.inf and x[d] > pmin:
                out[d] = pmin
            if x[d] != 0:
                pmin = min(pmin, x[d])
        assert np.array_equal(out, ref), x


def test_tiny_chromosome_raises_like_reference():
    with pytest.raises(ValueError):
        oe_tables(100000, 2)
    oe_tables(100000, 1)


def test_uniform_filter_replica():
    """The summation order the GPU fp64 filter path follows is scipy's, bit for bit."""
    rng = np.random.default_rng(0)
    for shape, size in [((120, 100), 12), ((100, 80), 10), ((80, 80), 8), ((37, 53), 5),
                        ((30, 30), 2), ((12, 40), 30), ((50, 64), 7)]:
        a = np.round(rng.standard_normal(shape) * 3, 2)
        assert np.array_equal(O.box_filter_replica(a, size), O.box_filter_reflect(a, size))
