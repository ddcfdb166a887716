"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs and against the reference's golden vectors. Run on the B200 box: pytest -m gpu.

Tolerances (BASELINE.json north_star): scattered non-zeros bit-exact; O/E (logged) and
similarity entries within 1e-5 absolute; raw O/E additionally 2e-7 relative (fp32 map storage).
"""
import numpy as np
import pytest
import torch

from oracle import hia_oracle as O
from hiarch_b200 import synth

pytestmark = pytest.mark.gpu

ATOL = 1e-5


def _dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype)


def _coo(g, prefix):
    return (_dev(g[f"{prefix}.rows"], torch.int32), _dev(g[f"{prefix}.cols"], torch.int32),
            _dev(g[f"{prefix}.vals"], torch.float32))


def _seps(g):
    return synth.chr_seps_from_lens([int(x) for x in g["in.lens"]])


# --------------------------------------------------------------------------------- scatter
@pytest.mark.parametrize("case", ["g3chr", "g1chr", "gquirk"])
def test_scatter_zero_fill_bit_exact(golden, case):
    from hiarch_b200 import ops
    g = golden(case)
    n = int(g["in.lens"].sum())
    r, c, v = _coo(g, "in")
    dense = ops.scatter_coo_sym(r, c, v, n)
    ref = O.scatter_coo_sym(g["in.rows"], g["in.cols"], g["in.vals"], n).astype(np.float32)
    assert np.array_equal(dense.cpu().numpy(), ref)


@pytest.mark.parametrize("case", ["g3chr", "g1chr", "gquirk"])
@pytest.mark.parametrize("cols16", [False, True])
def test_scatter_from_csr_bit_exact(golden, case, cols16):
    """CSR input (int32 and uint16 column indices) gives the oracle's dense map bit for bit, for the
    mirrored, the upper-only and the 'all' matrix types; unsorted columns, duplicates, an empty row
    and out-of-range columns included."""
    from hiarch_b200 import ops
    g = golden(case)
    n = int(g["in.lens"].sum())
    rows, cols, vals = g["in.rows"].copy(), g["in.cols"].copy(), g["in.vals"].copy()
    rng = np.random.default_rng(1)
    # shuffle the entries inside every row (CSR does not promise sorted columns) and add duplicates
    dup = rng.choice(rows.size, size=50, replace=False)
    drows, dcols = rows[dup], cols[dup]
    rows, cols, vals = np.r_[rows, drows], np.r_[cols, dcols], np.r_[vals, vals[dup] * 0.5]
    order = np.lexsort((rng.random(rows.size), rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    ref = O.scatter_coo_sym(rows, cols, vals, n).astype(np.float32)
    rp, cc, vv = ops.coo_to_csr_host(rows, cols, vals, n, cols16=cols16)
    assert cc.dtype == (torch.uint16 if cols16 else torch.int32) and rp.dtype == torch.int64
    rp, cc, vv = rp.cuda(), cc.cuda(), vv.cuda()
    full = ops.scatter_csr_sym(rp, cc, vv, n, mtx_type=ops.MTX_TRIU)
    got = full.cpu().numpy()
    # duplicates are summed in fp32 in an order of their own: last bit there, exact everywhere else
    dmask = np.zeros((n, n), dtype=bool)
    dmask[drows, dcols] = True
    dmask |= dmask.T
    assert np.array_equal(got[~dmask], ref[~dmask])
    assert np.allclose(got[dmask], ref[dmask], rtol=3e-7, atol=0)
    up = ops.alloc_map(n)
    up.fill_(float("nan"))
    ops.scatter_csr_sym(rp, cc, vv, n, mtx_type=ops.MTX_TRIU_UPPER, out=up)
    iu = np.triu_indices(n)
    assert np.array_equal(up.cpu().numpy()[iu], got[iu])             # upper triangle written, lower untouched
    assert torch.isnan(torch.tril(up, -1)[np.tril_indices(n, -1)]).all()
    # 'all' type: the symmetric matrix as CSR of every cell
    fr, fc = np.nonzero(ref)
    rp2, cc2, vv2 = ops.coo_to_csr_host(fr, fc, ref[fr, fc], n, cols16=cols16)
    allm = ops.scatter_csr_sym(rp2.cuda(), cc2.cuda(), vv2.cuda(), n, mtx_type=ops.MTX_ALL)
    assert np.array_equal(allm.cpu().numpy(), ref)


@pytest.mark.parametrize("case", ["g3chr", "g1chr"])
def test_scatter_min_fill_bit_exact(golden, case):
    from hiarch_b200 import ops
    g = golden(case)
    n = int(g["in.lens"].sum())
    seps = _seps(g)
    layout = ops.GenomeLayout(seps)
    r, c, v = _coo(g, "de_file")
    ref_g = O.scatter_coo_sym(g["de_file.rows"], g["de_file.cols"], g["de_file.vals"], n, has_neg=True)
    ref_b = O.scatter_coo_sym(g["de_file.rows"], g["de_file.cols"], g["de_file.vals"], n, has_neg=True,
                              seps=seps)
    d_g = ops.scatter_coo_sym(r, c, v, n, fill_mode=ops.FILL_MIN, layout=layout)
    d_b = ops.scatter_coo_sym(r, c, v, n, fill_mode=ops.FILL_BLOCK_MIN, layout=layout)
    assert np.array_equal(d_g.cpu().numpy(), ref_g.astype(np.float32))
    assert np.array_equal(d_b.cpu().numpy(), ref_b.astype(np.float32))
    # the block the reference densified itself (explicit 0.00 entries stay 0)
    s0, e0 = seps[0]
    assert np.array_equal(d_b[s0:e0 + 1, s0:e0 + 1].cpu().numpy(), g["checker.block0"].astype(np.float32))


def test_scatter_duplicates_lower_entries_and_empty():
    from hiarch_b200 import ops
    rng = np.random.default_rng(5)
    n = 257
    rows = rng.integers(0, n, 5000).astype(np.int32)
    cols = rng.integers(0, n, 5000).astype(np.int32)
    vals = rng.integers(-8, 9, 5000).astype(np.float32)       # small ints: any sum order is exact
    for has_neg, mode in ((False, ops.FILL_ZERO), (True, ops.FILL_MIN)):
        ref = O.scatter_coo_sym(rows, cols, vals, n, has_neg=has_neg)
        got = ops.scatter_coo_sym(_dev(rows, torch.int32), _dev(cols, torch.int32),
                                  _dev(vals, torch.float32), n, fill_mode=mode)
        assert np.array_equal(got.cpu().numpy(), ref.astype(np.float32))
    ref = O.scatter_coo_sym(rows, cols, vals, n, mtx_type="all")
    got = ops.scatter_coo_sym(_dev(rows, torch.int32), _dev(cols, torch.int32), _dev(vals, torch.float32),
                              n, mtx_type=ops.MTX_ALL)
    assert np.array_equal(got.cpu().numpy(), ref.astype(np.float32))
    empty = ops.scatter_coo_sym(torch.empty(0, dtype=torch.int32, device="cuda"),
                                torch.empty(0, dtype=torch.int32, device="cuda"),
                                torch.empty(0, dtype=torch.float32, device="cuda"), 5)
    assert float(empty.abs().sum()) == 0.0


@pytest.mark.parametrize("n", [257, 17000])
def test_scatter_sorted_row_path_and_fallback(n):
    """Row-sorted entries take the shared-memory row path (duplicates, lower-triangle entries of a
    'triu' input, empty rows, rows longer than one 16 K-column segment); the same entries shuffled
    take the atomic fallback; both must equal the oracle bit for bit (small-integer values: any
    summation order is exact)."""
    from hiarch_b200 import ops
    rng = np.random.default_rng(n)
    nnz = 40 * n
    rows = rng.integers(0, n, nnz).astype(np.int32)
    rows[rows % 7 == 3] = 5                                    # empty rows + one very long row
    cols = rng.integers(0, n, nnz).astype(np.int32)
    vals = rng.integers(-8, 9, nnz).astype(np.float32)
    order = np.argsort(rows, kind="stable")                    # row-sorted, columns in random order
    # row- AND column-sorted: rows without duplicates take the plain-store path of the fill kernel, the ~5 % of
    # rows with a duplicate column (and the very long row) are redone with atomics by their own CTA
    canonical = np.lexsort((cols, rows))
    for mtx_type, name in ((ops.MTX_TRIU, "triu"), (ops.MTX_ALL, "all")):
        ref = O.scatter_coo_sym(rows, cols, vals, n, mtx_type=name).astype(np.float32)
        for idx in (order, canonical, rng.permutation(nnz)):
            got = ops.scatter_coo_sym(_dev(rows[idx], torch.int32), _dev(cols[idx], torch.int32),
                                      _dev(vals[idx], torch.float32), n, mtx_type=mtx_type)
            assert np.array_equal(got.cpu().numpy(), ref)
    # out-of-range entries are ignored on both paths (like the any-order kernel always did)
    bad_r = np.concatenate([rows[order], [n - 1]]).astype(np.int32)
    bad_c = np.concatenate([cols[order], [n + 3]]).astype(np.int32)
    bad_v = np.concatenate([vals[order], [5.0]]).astype(np.float32)
    got = ops.scatter_coo_sym(_dev(bad_r, torch.int32), _dev(bad_c, torch.int32), _dev(bad_v, torch.float32), n,
                              mtx_type=ops.MTX_ALL)
    assert np.array_equal(got.cpu().numpy(), O.scatter_coo_sym(rows, cols, vals, n, mtx_type="all").astype(np.float32))


# --------------------------------------------------------------------------------- O/E + log
def _check_oe(dense64, seps, bs):
    from hiarch_b200 import ops
    n = dense64.shape[0]
    layout = ops.GenomeLayout(seps)
    dmap = ops.alloc_map(n)
    dmap.copy_(_dev(dense64, torch.float32))
    ref_oe, ref_exp = O.obs_d_exp(dense64, seps, bs, return_expected=True)
    ref_log = O.log_map(ref_oe)
    got_oe, state = ops.obs_d_exp(dmap, layout, bs)
    got_log, _ = ops.obs_d_exp(dmap, layout, bs, log=True)
    exp = state.expected.cpu().numpy()
    for (s, e), rexp in zip(seps, ref_exp):
        if rexp is not None:
            assert np.allclose(exp[s:e + 1], rexp, rtol=1e-6, atol=0)
    go = got_oe.cpu().numpy().astype(np.float64)
    assert np.all(np.abs(go - ref_oe) <= ATOL + 2e-7 * np.abs(ref_oe))
    # the fp64 output holds the north star's 1e-5 ABSOLUTE with no relative term, whatever the value
    got64, _ = ops.obs_d_exp(dmap, layout, bs, fp64_out=True)
    assert got64.dtype == torch.float64 and np.abs(got64.cpu().numpy() - ref_oe).max() <= ATOL
    got64_oi, _ = ops.obs_d_exp(dmap, layout, bs, only_intra=True, fp64_out=True)
    assert np.abs(got64_oi.cpu().numpy() - O.obs_d_exp(dense64, seps, bs, only_intra=True)).max() <= ATOL
    assert np.abs(got_log.cpu().numpy().astype(np.float64) - ref_log).max() <= ATOL
    # no clamp, only_intra
    got_nc, _ = ops.obs_d_exp(dmap, layout, bs, counts_must_decrese=False, log=True)
    ref_nc = O.log_map(O.obs_d_exp(dense64, seps, bs, counts_must_decrese=False))
    assert np.abs(got_nc.cpu().numpy().astype(np.float64) - ref_nc).max() <= ATOL
    got_oi, _ = ops.obs_d_exp(dmap, layout, bs, only_intra=True)
    ref_oi = O.obs_d_exp(dense64, seps, bs, only_intra=True)
    goi = got_oi.cpu().numpy().astype(np.float64)
    assert np.all(np.abs(goi - ref_oi) <= ATOL + 2e-7 * np.abs(ref_oi))
    # in place
    work = ops.alloc_map(n)
    work.copy_(dmap)
    ops.obs_d_exp(work, layout, bs, log=True, out=work)
    assert torch.equal(work, got_log)


@pytest.mark.parametrize("case", ["g3chr", "g1chr", "gquirk"])
def test_oe_against_reference_golden(golden, case):
    from hiarch_b200 import ops
    g = golden(case)
    n = int(g["in.lens"].sum())
    bs = int(g["in.bin_size"])
    seps = _seps(g)
    r, c, v = _coo(g, "in")
    dense = ops.scatter_coo_sym(r, c, v, n)
    layout = ops.GenomeLayout(seps)
    got, state = ops.obs_d_exp(dense, layout, bs)
    got_log, _ = ops.obs_d_exp(dense, layout, bs, log=True)
    ref = g["oe"]
    assert np.all(np.abs(got.cpu().numpy() - ref) <= ATOL + 2e-7 * np.abs(ref))
    got64, _ = ops.obs_d_exp(dense, layout, bs, fp64_out=True)          # raw O/E: 1e-5 absolute, no relative term
    assert np.abs(got64.cpu().numpy() - ref).max() <= ATOL
    assert np.abs(got_log.cpu().numpy() - g["oe_log"]).max() <= ATOL
    s, e = seps[-1]
    assert np.allclose(state.expected.cpu().numpy()[s:e + 1], g["expected_last_block"], rtol=1e-6)


@pytest.mark.parametrize("lens,bs,seed", [([700, 333, 129, 50], 100000, 11), ([1500], 50000, 12),
                                          ([37, 1030, 64], 25000, 13)])
def test_oe_against_oracle_synthetic(lens, bs, seed):
    lam, seps = synth.rate_matrix(lens, seed)
    up = synth.sample_counts(lam, seed)
    dense = up + np.triu(up, 1).T
    _check_oe(dense, seps, bs)


def test_oe_degenerate_blocks():
    """all-zero chromosome, zero rows, an inter block with no contacts."""
    lens = [60, 40, 30]
    lam, seps = synth.rate_matrix(lens, 3)
    up = synth.sample_counts(lam, 3)
    dense = up + np.triu(up, 1).T
    s, e = seps[1]
    dense[s:e + 1, s:e + 1] = 0.0            # all-zero intra block (np.sum == 0 -> unchanged)
    s2, e2 = seps[2]
    dense[s:e + 1, s2:e2 + 1] = 0.0          # empty inter block
    dense[s2:e2 + 1, s:e + 1] = 0.0
    dense[5, :] = 0.0
    dense[:, 5] = 0.0
    _check_oe(dense, seps, 100000)


# --------------------------------------------------------------------------------- similarity
def _sim_ref(x64, flavour):
    return O.cosine_distance(x64) if flavour == 0 else O.pearson(x64)


@pytest.mark.parametrize("m,k", [(50, 50), (130, 77), (257, 300), (640, 1000), (1000, 33)])
@pytest.mark.parametrize("flavour", [0, 1])
@pytest.mark.parametrize("precision", [0, 1, 2])        # 3xTF32, FP64 DMMA, 3xF16
def test_similarity_small(m, k, flavour, precision):
    from hiarch_b200 import ops
    rng = np.random.default_rng(m * 1000 + k)
    x = rng.standard_normal((m, k)).astype(np.float32)
    x[:, : k // 3] += 0.7                                 # non-zero means, correlated rows
    x[: m // 2] += rng.standard_normal(k).astype(np.float32)
    xd = ops.alloc_map(m, cols=k)
    xd.copy_(_dev(x, torch.float32))
    got = ops.similarity(xd, flavour=flavour, precision=precision).cpu().numpy().astype(np.float64)
    ref = _sim_ref(x.astype(np.float64), flavour)
    assert np.abs(got - ref).max() <= (5e-7 if precision == 1 else ATOL)
    assert np.array_equal(got, got.T)
    assert np.all(np.diag(got) == (0.0 if flavour == 0 else 1.0))


def test_similarity_long_k_chunked_accumulation():
    """K = 20 000: many MMA steps per output; 1e-5 must hold for both chunk settings."""
    from hiarch_b200 import ops
    rng = np.random.default_rng(77)
    m, k = 512, 20000
    base = rng.standard_normal(k)
    x = (0.8 * base[None, :] + 0.6 * rng.standard_normal((m, k)) + 0.3).astype(np.float32)
    xd = ops.alloc_map(m, cols=k)
    xd.copy_(_dev(x, torch.float32))
    ref = O.pearson(x.astype(np.float64))
    refc = O.cosine_distance(x.astype(np.float64))
    for precision in (ops.PREC_3XTF32, ops.PREC_3XF16):
        for k_chunk in (0, 256, 1024, 1 << 20):              # 0 = the precision's default
            got = ops.similarity(xd, flavour=1, k_chunk=k_chunk, precision=precision).cpu().numpy().astype(np.float64)
            err = np.abs(got - ref).max()
            print(f"pearson precision={precision} k_chunk={k_chunk}: max abs err {err:.3e}")
            if k_chunk <= 1024:
                assert err <= ATOL
        got = ops.similarity(xd, flavour=0, precision=precision).cpu().numpy().astype(np.float64)
        assert np.abs(got - refc).max() <= ATOL


def test_similarity_f16_split_extreme_rows():
    """3xF16 planes hold the unit row scaled by 2^14: rows dominated by ONE entry (|xn| -> 1, the
    fp16 range limit), rows with a huge dynamic range (tiny entries fall into fp16 subnormals) and
    an all-zero row (NaN like pdist) must behave like the fp64 reference."""
    from hiarch_b200 import ops
    rng = np.random.default_rng(9)
    m, k = 300, 1500
    x = rng.standard_normal((m, k)).astype(np.float32)
    x[0, :] = 0.0; x[0, 7] = 5.0                              # a single non-zero entry: xn = 1 exactly
    x[1, :] *= 1e-4; x[1, 11] = 1e3                           # one dominant + tiny entries
    x[2, :] = np.exp(rng.uniform(-18, 4, k)).astype(np.float32)     # 9 decades of dynamic range
    x[3, :] = 0.0                                             # zero row
    xd = ops.alloc_map(m, cols=k)
    xd.copy_(_dev(x, torch.float32))
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = O.cosine_distance(x.astype(np.float64))
    got = ops.similarity(xd, flavour=0, precision=ops.PREC_3XF16).cpu().numpy().astype(np.float64)
    ok = np.ones(m, bool); ok[3] = False
    assert np.abs(got[np.ix_(ok, ok)] - ref[np.ix_(ok, ok)]).max() <= ATOL
    off = np.arange(m) != 3
    assert np.all(np.isnan(got[3, off])) and np.all(np.isnan(got[off, 3]))


def test_similarity_golden_adj_and_band(golden):
    from hiarch_b200 import ops
    g = golden("g3chr")
    blk = g["checker.block0"]
    xd = ops.alloc_map(blk.shape[0], cols=blk.shape[1])
    xd.copy_(_dev(blk, torch.float32))
    got = ops.similarity(xd).cpu().numpy().astype(np.float64)
    assert np.abs(got - g["checker.adj0"]).max() <= ATOL          # vs the reference's pdist
    # band-limited: cells inside the band equal the full result
    rng = np.random.default_rng(1)
    m = 900
    x = rng.standard_normal((m, 400)).astype(np.float32)
    xd = ops.alloc_map(m, cols=400)
    xd.copy_(_dev(x, torch.float32))
    full = ops.similarity(xd).cpu().numpy()
    lo, hi = 45, 135
    band = ops.similarity(xd, band=(lo, hi)).cpu().numpy()
    i, j = np.indices((m, m))
    inb = (np.abs(i - j) >= lo) & (np.abs(i - j) <= hi)
    assert np.array_equal(band[inb], full[inb])


# --------------------------------------------------------------------------------- top-k eigen
def _check_eig(C64, k=3, tol=1e-6):
    from hiarch_b200 import ops
    n = C64.shape[0]
    cd = ops.alloc_map(n)
    cd.copy_(_dev(C64, torch.float32))
    c32 = cd.cpu().numpy().astype(np.float64)             # the identical matrix the GPU sees
    w_ref, v_ref = O.topk_eig(c32, k)
    w, v, info = ops.topk_eig(cd, k=k, tol=tol, return_info=True)
    w, v = w.cpu().numpy(), v.cpu().numpy().astype(np.float64)
    print("eig restarts/residuals:", info)
    assert np.allclose(w, w_ref, rtol=1e-5)
    for j in range(k):
        cos = abs(v[:, j] @ v_ref[:, j]) / np.linalg.norm(v[:, j])
        assert cos >= 0.9999, (j, cos)                    # north-star: |cos| >= 0.9999
    assert np.allclose(np.linalg.norm(v, axis=0), 1.0, atol=1e-4)


def test_topk_eig_planted_spectrum():
    rng = np.random.default_rng(3)
    n = 700
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([[50.0, 31.0, 17.0, 9.0], rng.uniform(0.0, 4.0, n - 4)])
    _check_eig((q * lam) @ q.T)


def test_topk_eig_of_pearson_map(golden):
    g = golden("g3chr")
    _check_eig(O.pearson(g["norm_de"]))


def test_topk_eig_small_and_k1():
    rng = np.random.default_rng(4)
    a = rng.standard_normal((64, 20))
    _check_eig(a @ a.T + np.diag(np.linspace(0, 30, 64)), k=1)
    _check_eig(a @ a.T + np.diag(np.linspace(0, 30, 64)), k=3)


# --------------------------------------------------------------------------------- a6 selection
def _mapdev(a):
    from hiarch_b200 import ops
    d = ops.alloc_map(a.shape[0], cols=a.shape[1])
    d.copy_(_dev(a, torch.float32))
    return d


def test_percentiles_exact_numpy_semantics():
    from hiarch_b200 import ops
    rng = np.random.default_rng(9)
    cases = [rng.standard_normal((301, 257)), np.round(rng.standard_normal((64, 1000)), 1),
             rng.integers(-3, 4, (129, 130)).astype(np.float64), np.abs(rng.standard_normal((50, 51))),
             np.full((40, 44), 2.5)]
    for a in cases:
        a32 = a.astype(np.float32)
        d = _mapdev(a32)
        a64 = a32.astype(np.float64)
        got = ops.percentiles(d, [2, 98, 66, 100 / 3]).cpu().numpy()
        ref = np.array([np.percentile(a64, q) for q in (2, 98, 66, 100 / 3)])
        assert np.array_equal(got, ref), (got, ref)
        got = ops.percentiles(d, [98, 2, 0, 100], [ops.PCT_POS, ops.PCT_NEG, ops.PCT_ALL, ops.PCT_ALL]).cpu().numpy()
        pos, neg = a64[a64 > 0], a64[a64 < 0]
        ref = [np.percentile(pos, 98) if pos.size else np.nan, np.percentile(neg, 2) if neg.size else np.nan,
               a64.min(), a64.max()]
        assert np.array_equal(got, np.array(ref), equal_nan=True), (got, ref)


def test_clip_zscore_chain_against_golden(golden):
    from hiarch_b200 import ops
    g = golden("g3chr")
    d = _mapdev(g["oe_log"])
    x64 = d.cpu().numpy().astype(np.float64)                       # what the GPU starts from
    t = ops.tid_off_outliers(d)
    assert np.abs(t.cpu().numpy() - O.tid_off_outliers(x64)).max() <= 1e-6
    assert np.abs(t.cpu().numpy() - g["tid_off"]).max() <= ATOL
    z = ops.norm_to_zero_mean(t)
    assert np.abs(z.cpu().numpy() - g["zscore"]).max() <= ATOL
    nd = ops.norm_de_mtx(t)
    assert np.abs(nd.cpu().numpy() - g["norm_de"]).max() <= ATOL
    # in place
    w = _mapdev(g["oe_log"])
    ops.tid_off_outliers(w, out=w)
    assert torch.equal(w, t)


def test_norm_to_zero_mean_degenerate_core():
    """vmin >= vmax (constant matrix region) -> falls back to all values."""
    from hiarch_b200 import ops
    a = np.zeros((60, 64))
    a[:3] = 5.0                                                    # p34 == p66 == 0
    got = ops.norm_to_zero_mean(_mapdev(a)).cpu().numpy()
    assert np.abs(got - O.norm_to_zero_mean(a)).max() <= ATOL


# --------------------------------------------------------------------------------- a10 checkerboard
@pytest.mark.parametrize("case", ["g3chr", "g1chr"])
def test_checkerboard_table_against_reference(golden, case):
    """.mtx (%.2f) -> per-block min fill -> cosine distance band -> entropy, vs main_local."""
    from hiarch_b200 import ops
    g = golden(case)
    n = int(g["in.lens"].sum())
    seps = _seps(g)
    names = [str(x) for x in g["in.names"]]
    layout = ops.GenomeLayout(seps)
    r, c, v = _coo(g, "de_file")
    dense = ops.scatter_coo_sym(r, c, v, n, fill_mode=ops.FILL_BLOCK_MIN, layout=layout)
    table = ops.main_local(dense, layout, names)
    assert [t[0] for t in table] == list(g["checker.accord_chro"])
    assert [t[1] for t in table] == list(g["checker.other_chro"])
    got = np.array([t[2] for t in table])
    assert np.all(np.abs(got - g["checker.accord"]) <= 1e-4 * np.abs(g["checker.accord"])), \
        (got, g["checker.accord"])                          # north-star: 1e-4 relative


def test_band_entropy_against_oracle_on_exact_adj():
    """Same fp32 distance matrix on both sides: the reduction itself must match to 1e-12."""
    from hiarch_b200 import ops
    rng = np.random.default_rng(21)
    for n, L in ((180, 180), (60, 60), (400, 333)):
        x = rng.standard_normal((n, 90)) + rng.standard_normal(90)
        adj = O.cosine_distance(x).astype(np.float32)
        d = _mapdev(adj)
        smin, smax = int(L * .05), int(L * .15)
        res = ops.band_entropy(d, smin, smax).cpu().numpy()
        ref = O.checkerboard_from_adj(adj.astype(np.float64), L)
        if ref is None:
            assert res[1] == 0
        else:
            assert abs(res[0] - ref) <= 1e-12 * abs(ref), (res, ref)
    assert ops.get_local(_mapdev(rng.standard_normal((40, 200)))) is None      # < 50 rows


# --------------------------------------------------------------------------------- a11 GF_S1 profiles
@pytest.mark.parametrize("case", ["g3chr", "g1chr"])
def test_gf_s1_profiles(golden, case):
    from hiarch_b200 import ops
    g = golden(case)
    seps = _seps(g)
    layout = ops.GenomeLayout(seps)
    d = _mapdev(g["filt"])
    raw, sm = ops.inter_rowsum_profile(d, layout)
    assert np.abs(raw.cpu().numpy() - g["inter_rowsum"]).max() <= 1e-6
    assert np.abs(sm.cpu().numpy() - g["inter_rowsum_smooth"]).max() <= 1e-6
    if "intra_rowsum" in g:
        raw, sm = ops.intra_low_profile(d, layout)
        assert np.abs(raw.cpu().numpy() - g["intra_rowsum"]).max() <= 1e-6
        assert np.abs(sm.cpu().numpy() - g["intra_rowsum_smooth"]).max() <= 1e-6


# --------------------------------------------------------------------------------- a14/a15 GF_S2
def test_gf_s2_images_and_scores(golden):
    from hiarch_b200 import ops
    g = golden("g3chr")
    n = int(g["in.lens"].sum())
    seps = _seps(g)
    names = [str(x) for x in g["in.names"]]
    layout = ops.GenomeLayout(seps)
    r, c, v = _coo(g, "filt_file")
    dense = ops.scatter_coo_sym(r, c, v, n)                 # scipy .todense(): absent = 0
    for method in ("inter_rowsum", "no_anchor"):
        used = {}
        for ch, ty, cen in zip(g[f"anchors.{method}.chr"], g[f"anchors.{method}.type"], g[f"anchors.{method}.cen"]):
            if ty == "used" and str(ch) not in used:
                used[str(ch)] = int(cen) - seps[names.index(str(ch))][0]
        res = ops.gf_scores(dense, layout, [used.get(nm) for nm in names])
        intra = [i for i, (a, b) in enumerate(res["pairs"]) if a == b]
        inter = [i for i, (a, b) in enumerate(res["pairs"]) if a != b]
        assert np.abs(res["xs"][intra] - g[f"gf.{method}.intra_xs"]).max() <= 1e-6
        assert np.abs(res["xs"][inter] - g[f"gf.{method}.inter_xs"]).max() <= 1e-6
        si = res["scores"][intra].T.reshape(-1)             # golden is melted column-major
        so = res["scores"][inter].T.reshape(-1)
        ri, ro = g[f"gf.{method}.intra_score"], g[f"gf.{method}.inter_score"]
        assert np.all(np.abs(si - ri) <= 1e-4 * np.abs(ri) + 1e-9)      # north-star: 1e-4 relative
        assert np.all(np.abs(so - ro) <= 1e-4 * np.abs(ro) + 1e-9)


# --------------------------------------------------------------------------------- a7 filters
@pytest.mark.parametrize("shape,size", [((120, 120), 12), ((100, 80), 10), ((80, 120), 8), ((64, 33), 5),
                                        ((30, 30), 2), ((12, 300), 30), ((50, 50), 1), ((2500, 300), 250)])
def test_box_filter_vs_scipy(shape, size):
    from hiarch_b200 import ops
    rng = np.random.default_rng(size)
    a = rng.standard_normal(shape).astype(np.float32)
    d = _mapdev(a)
    got = ops.box_filter_reflect(d, size, out=_mapdev(np.zeros(shape))).cpu().numpy()
    ref = O.box_filter_reflect(a.astype(np.float64), size)
    assert np.abs(got - ref).max() <= 2e-6


@pytest.mark.parametrize("shape,size", [((120, 120), 2), ((100, 80), 5), ((64, 33), 3), ((40, 40), 8),
                                        ((300, 200), 24)])
def test_median_filter_vs_scipy(shape, size):
    from hiarch_b200 import ops
    rng = np.random.default_rng(size)
    a = np.round(rng.standard_normal(shape), 2).astype(np.float32)
    got = ops.median_filter_reflect(_mapdev(a), size).cpu().numpy()
    ref = O.median_filter_reflect(a.astype(np.float64), size)
    assert np.array_equal(got, ref.astype(np.float32))          # a selection: bit-exact


@pytest.mark.parametrize("case", ["g3chr", "g1chr"])
def test_main_filter_against_reference(golden, case):
    """GF_S1 main_filter through the fp64 scipy-order replica (csrc/hia_filter64.cu): the core
    selection of norm_to_zero_mean sits on massive ties, so only the exact summation order
    reproduces the reference."""
    from hiarch_b200 import ops
    g = golden(case)
    n = int(g["in.lens"].sum())
    seps = _seps(g)
    layout = ops.GenomeLayout(seps)
    # (1) un-quantised input (fp32 values promoted as they are)
    x32 = g["norm_de"].astype(np.float32)
    ref = O.main_filter(x32.astype(np.float64), seps)
    got = ops.main_filter(_mapdev(x32), layout, decimals=-1).cpu().numpy()
    assert np.all(np.abs(got - ref) <= 2e-6 + 2e-7 * np.abs(ref)), np.abs(got - ref).max()
    # (2) the pipeline's %.2f file -> the reference's own output
    r, c, v = _coo(g, "de_file")
    dense = ops.scatter_coo_sym(r, c, v, n, fill_mode=ops.FILL_MIN, layout=layout)
    filt = ops.main_filter(dense, layout, decimals=2).cpu().numpy().astype(np.float64)
    assert np.all(np.abs(filt - g["filt"]) <= 2e-6 + 2e-7 * np.abs(g["filt"])), np.abs(filt - g["filt"]).max()


def test_used_filter_box_only_is_bit_exact_vs_scipy():
    from hiarch_b200 import ops
    rng = np.random.default_rng(8)
    for shape, size in (((120, 100), 12), ((37, 53), 5), ((12, 40), 30), ((64, 64), 1)):
        a = np.round(rng.standard_normal(shape) * 3, 2)
        d = _mapdev(a)
        out = _mapdev(np.zeros(shape))
        ops.used_filter(d, size, out, decimals=2, normalise=False)
        ref = O.box_filter_reflect(a, size) if size > 1 else a
        assert np.array_equal(out.cpu().numpy(), ref.astype(np.float32))


def test_used_filter_normalised_on_massively_tied_blocks():
    """used_filter = box filter + norm_to_zero_mean on count-like blocks: a few distinct levels
    (nearly every cell is a candidate of the exact 64-bit select: the compacting pass keeps almost the
    whole block), size 1 (no filter: the candidate list lives in the other fp64 buffer), a constant
    block interior, rectangular shapes."""
    from hiarch_b200 import ops
    rng = np.random.default_rng(21)
    for shape, size, levels in (((90, 70), 1, 3), ((90, 70), 4, 6), ((257, 129), 7, 2), ((64, 200), 5, 40),
                                ((300, 300), 30, 4)):
        a = rng.integers(0, levels, shape).astype(np.float64)
        a[: shape[0] // 3, : shape[1] // 3] = 1.0                      # a constant patch: exact ties after the filter
        d = _mapdev(a)
        out = _mapdev(np.zeros(shape))
        ops.used_filter(d, size, out, decimals=2, normalise=True)
        ref = O.norm_to_zero_mean(O.box_filter_reflect(a, size) if size > 1 else a)
        got = out.cpu().numpy().astype(np.float64)
        assert np.all(np.isfinite(got))
        assert np.all(np.abs(got - ref) <= 2e-6 + 2e-7 * np.abs(ref)), (shape, size, np.abs(got - ref).max())


def test_oe_and_min_fill_on_the_option_fixtures(golden):
    """goptions.npz (real reference, tests/golden/gen_golden_options.py): ragged chromosomes
    (70 / 12 / 31 / 6 bins, all-zero diagonals, an all-zero inter block), only_intra, and the global
    minimum fill of a has_neg file with explicit 0.00 entries."""
    from hiarch_b200 import ops
    g = golden("goptions")

    def seps_of(lens):
        out, s = [], 0
        for L in lens:
            out.append((s, s + int(L) - 1))
            s += int(L)
        return out

    _check_oe(g["r.dense"], seps_of(g["r.lens"]), 1000000)
    _check_oe(g["a.dense"], seps_of(g["a.lens"]), 400000)
    n = int(g["a.lens"].sum())
    got = ops.scatter_coo_sym(_dev(g["de.rows"], torch.int32), _dev(g["de.cols"], torch.int32),
                              _dev(g["de.vals"], torch.float32), n, fill_mode=ops.FILL_MIN)
    assert np.array_equal(got.cpu().numpy(), g["dense.global_min"].astype(np.float32))


# --------------------------------------------------------------------------------- a13 intra-x
@pytest.mark.parametrize("method", ["prefix", "direct"])
def test_intra_x_response(golden, method):
    from hiarch_b200 import ops
    g = golden("g1chr")
    got = ops.intra_x_response(_mapdev(g["filt"]), method=method).cpu().numpy()
    ref = g["intra_x"]
    assert np.all(np.abs(got - ref) <= 1e-4 * np.abs(ref) + 1e-7)
    rng = np.random.default_rng(2)
    for n in (1, 2, 61, 150):
        a = rng.standard_normal((n, n))
        if n != 150:
            a = (a + a.T) / 2                                    # n = 150: unsymmetric block
        got = ops.intra_x_response(_mapdev(a), method=method).cpu().numpy()
        ref = O.intra_x(a.astype(np.float32).astype(np.float64))
        assert np.all(np.abs(got - ref) <= 1e-9 + 1e-9 * np.abs(ref))


def test_intra_x_prefix_equals_direct_on_a_multi_panel_block_view():
    """O(n^2) prefix-sum path vs the direct band evaluation: several row panels (n > 1024), a
    ragged last tile, and an unaligned view into a larger map (odd offset, ld of the parent)."""
    from hiarch_b200 import ops
    rng = np.random.default_rng(11)
    n_par, off, n = 2400, 131, 2113
    d = np.abs(np.arange(n_par)[:, None] - np.arange(n_par)[None, :])
    parent = (rng.poisson(50.0 * (d + 1.0) ** -1.0 + 1.0) + rng.random((n_par, n_par))).astype(np.float32)
    dev = _mapdev(parent)
    view = dev[off:off + n, off:off + n]
    a = ops.intra_x_response(view, method="prefix").cpu().numpy()
    b = ops.intra_x_response(view, method="direct").cpu().numpy()
    assert np.all(np.isfinite(a))
    assert np.all(np.abs(a - b) <= 1e-10 * np.abs(b)), (np.abs(a - b) / np.abs(b)).max()
    a2 = ops.intra_x_response(view, method="prefix").cpu().numpy()
    assert np.array_equal(a, a2)                                 # deterministic (no atomics)


# --------------------------------------------------------------------------------- f1 preprocess
@pytest.mark.parametrize("case", ["gpre_fold", "gpre_sparse", "gpre_none"])
def test_preprocess_mtx_matches_reference(golden, case, tmp_path):
    """S1_preprocess.py:48-83 through the remap-scatter / COO statistics kernels: the log (every
    fold / trim / gate decision), the bin table (bit-exact) and the coarse map."""
    from hiarch_b200.preprocess import preprocess_mtx
    from test_host_preprocess import check_preprocess_case
    check_preprocess_case(case, golden, tmp_path, preprocess_mtx)


def test_remap_scatter_and_coo_statistics_vs_numpy():
    from hiarch_b200 import _lib, ops
    lib = _lib.load()
    rng = np.random.default_rng(5)
    n_fine, nnz = 700, 60000
    r = rng.integers(0, n_fine, nnz); c = rng.integers(0, n_fine, nnz)
    r, c = np.minimum(r, c), np.maximum(r, c)
    lin = np.unique(r * n_fine + c)                     # duplicate-free, like the host's coo -> csr
    r, c, nnz = lin // n_fine, lin % n_fine, len(lin)
    v = np.round(rng.gamma(2.0, 3.0, nnz), 2)
    v[rng.random(nnz) < .02] = 0.0
    fmap = (np.arange(n_fine) // 3).astype(np.int32)
    fmap[rng.random(n_fine) < .1] = -1
    kept = np.unique(fmap[fmap >= 0]); rel = np.full(fmap.max() + 1, -1); rel[kept] = np.arange(len(kept))
    fmap = np.where(fmap >= 0, rel[np.maximum(fmap, 0)], -1).astype(np.int32)
    n_out = int(fmap.max()) + 1
    off = r != c
    rr = np.concatenate([fmap[r], fmap[c[off]]]); cc = np.concatenate([fmap[c], fmap[r[off]]])
    vv = np.concatenate([v, v[off]]).astype(np.float32).astype(np.float64)
    ok = (rr >= 0) & (cc >= 0)
    ref = np.zeros((n_out, n_out)); np.add.at(ref, (rr[ok], cc[ok]), vv[ok])
    dr, dc, dv, dm = _dev(r, torch.int32), _dev(c, torch.int32), _dev(v, torch.float32), _dev(fmap, torch.int32)
    out = ops.alloc_map(n_out, "cuda")
    _lib.check(lib.hia_scatter_coo_remap(dr.data_ptr(), dc.data_ptr(), dv.data_ptr(), nnz, n_fine, dm.data_ptr(),
                                         n_out, out.data_ptr(), out.stride(0), ops.MTX_TRIU, ops._stream()), "remap")
    got = out.cpu().numpy().astype(np.float64)
    assert np.abs(got - ref).max() <= 2e-6 * ref.max()
    assert np.array_equal(got, got.T)                                   # symmetric cell by cell? (same addends)
    bs = torch.empty(n_out, dtype=torch.float64, device="cuda")
    _lib.check(lib.hia_coo_bin_sums(dr.data_ptr(), dc.data_ptr(), dv.data_ptr(), nnz, n_fine, dm.data_ptr(), n_out,
                                    ops.MTX_TRIU, bs.data_ptr(), ops._stream()), "bin_sums")
    assert np.allclose(bs.cpu().numpy(), ref.sum(axis=1), rtol=1e-12, atol=0)
    rs = torch.empty(n_out, dtype=torch.float64, device="cuda"); rn = torch.empty(n_out, dtype=torch.int64, device="cuda")
    _lib.check(lib.hia_bin_stats(out.data_ptr(), n_out, out.stride(0), rs.data_ptr(), rn.data_ptr(), ops._stream()), "stats")
    assert np.array_equal(rn.cpu().numpy(), (got != 0).sum(axis=1))
    assert np.allclose(rs.cpu().numpy(), got.sum(axis=1), rtol=1e-12)
    # per-chromosome-pair stored non-zeros of the symmetrised fine matrix
    lens = [300, 250, 150]
    layout = ops.GenomeLayout(synth.chr_seps_from_lens(lens))
    cnt = torch.empty(9, dtype=torch.int64, device="cuda")
    _lib.check(lib.hia_coo_block_nnz(dr.data_ptr(), dc.data_ptr(), dv.data_ptr(), nnz, n_fine, ops.MTX_TRIU,
                                     layout.bin_chr.data_ptr(), 3, cnt.data_ptr(), ops._stream()), "block_nnz")
    chr_of = np.repeat(np.arange(3), lens)
    nz = v != 0
    ref_cnt = np.zeros((3, 3), dtype=np.int64)
    np.add.at(ref_cnt, (chr_of[r[nz]], chr_of[c[nz]]), 1)
    np.add.at(ref_cnt, (chr_of[c[nz & off]], chr_of[r[nz & off]]), 1)
    assert np.array_equal(cnt.cpu().numpy().reshape(3, 3), ref_cnt)


def test_get_adj_mtx_metrics_against_scipy_pdist():
    """get_adj_mtx(metric=...) (S1_get_adj_mtx.py:8-14 forwards any metric to pdist): 'cosine' and
    'correlation' run on the tensor cores and match scipy to 1e-5; anything else refuses loudly."""
    from scipy.spatial.distance import pdist, squareform
    from hiarch_b200 import complex as cx
    from hiarch_b200.new_hic_class import ArrayHiCMtx
    rng = np.random.default_rng(8)
    x = rng.standard_normal((300, 417)) + 0.7 * rng.standard_normal((300, 1)) + 0.4
    m = ArrayHiCMtx(x.astype(np.float32), mtx_type='all', has_neg=True)
    x32 = x.astype(np.float32).astype(np.float64)
    for metric in ("cosine", "correlation"):
        got = cx.get_adj_mtx(m, metric=metric).matrix
        ref = squareform(pdist(x32, metric))
        assert np.abs(got - ref).max() <= ATOL, metric
        assert np.all(np.diag(got) == 0.0)
    with pytest.raises(NotImplementedError):
        cx.get_adj_mtx(m, metric="euclidean")


def test_template_scores_for_caller_supplied_templates():
    """to_confor with a caller's template dict (to_confor.py:35-47) on the native kernel: plain and
    transposed-max scores against numpy fp64."""
    from hiarch_b200 import confor
    rng = np.random.default_rng(4)
    xs = rng.standard_normal((37, 256))
    tm = {f"t{i}": rng.standard_normal(256) for i in range(5)}
    t = np.vstack(list(tm.values()))
    plain = confor._template_scores_device(xs, tm, False)
    assert np.abs(plain - xs @ t.T).max() <= 1e-12
    xt = xs.reshape(-1, 16, 16).transpose(0, 2, 1).reshape(-1, 256)
    both = confor._template_scores_device(xs, tm, True)
    assert np.abs(both - np.maximum(xs @ t.T, xt @ t.T)).max() <= 1e-12


def _projected_cases():
    rng = np.random.default_rng(21)
    cases = []
    for n in (1, 2, 3, 8, 24, 64, 96, 104, 160):
        a = rng.standard_normal((n, n))
        cases.append((f"wishart{n}", a @ a.T / max(n, 1)))
    q, _ = np.linalg.qr(rng.standard_normal((72, 72)))
    cases.append(("clustered", (q * np.r_[1410.5, 1404.5, 1392.2, 900, 800, np.linspace(500, 0.01, 67)]) @ q.T))
    q, _ = np.linalg.qr(rng.standard_normal((104, 104)))
    cases.append(("near_degenerate_indefinite", (q * np.r_[100.0, 100.0 + 1e-9, 99.0, np.linspace(50, -3, 101)]) @ q.T))
    t = np.zeros((64, 64))
    for i in range(64):
        for j in range(max(0, i - 8), min(64, i + 9)):
            t[i, j] = rng.standard_normal() / (1 + abs(i - j))
    cases.append(("banded_graded", (t + t.T) / 2 + np.diag(np.r_[202.0, 113.0, 62.0, 41.0, np.geomspace(3.5, 0.02, 60)])))
    cases.append(("diagonal", np.diag(np.linspace(5, 1, 16))))
    return cases


@pytest.mark.parametrize("solver", [0, 1])
def test_projected_eigensolvers_against_eigh(solver):
    """The Rayleigh-Ritz step of the Lanczos solver on its own: the partial solver (Householder +
    Sturm multisection + inverse iteration, top kk pairs) and the one-sided Jacobi against
    numpy.linalg.eigh -- Wishart, clustered (the 25 kb spectrum), near-degenerate + indefinite, banded
    and graded like a Krylov matrix, diagonal, and the 1 x 1 ... 3 x 3 corner cases."""
    from hiarch_b200 import _lib, ops
    lib = _lib.load()
    for name, t in _projected_cases():
        t = (t + t.T) / 2
        n = t.shape[0]
        kk = min(8, n)
        D = n + 8
        h = np.zeros((D, D))
        h[:n, :n] = np.tril(t)
        dh = torch.from_numpy(h).cuda()
        z = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        lam = torch.zeros(n, dtype=torch.float64, device="cuda")
        order = torch.zeros(n, dtype=torch.int32, device="cuda")
        scratch = torch.zeros((n, n), dtype=torch.float64, device="cuda")
        _lib.check(lib.hia_projected_eig(dh.data_ptr(), D, n, kk, solver, z.data_ptr(), lam.data_ptr(),
                                         order.data_ptr(), scratch.data_ptr(), ops._stream()), "hia_projected_eig")
        torch.cuda.synchronize()
        w, u = np.linalg.eigh(t)
        w, u = w[::-1], u[:, ::-1]
        o = order.cpu().numpy()
        lam_h = lam.cpu().numpy()[o]
        zc = z.cpu().numpy()                                   # [column][row] because column-major
        scale = max(abs(w[0]), abs(w[-1]), 1e-300)
        got = lam_h[:kk]
        assert np.abs(got - w[:kk]).max() <= 1e-12 * scale, (name, solver)
        if solver == 0 and n > kk:
            assert abs(lam_h[kk] - w[kk]) <= 1e-12 * scale, (name, "next eigenvalue")
        vecs = np.stack([zc[o[j]] for j in range(kk)], axis=1)       # n x kk
        res = np.linalg.norm(t @ vecs - vecs * got[None, :], axis=0)
        assert res.max() <= 1e-10 * scale, (name, solver, res.max())
        assert np.abs(vecs.T @ vecs - np.eye(kk)).max() <= 1e-9, (name, solver)


@pytest.mark.parametrize("mode", ["coo_sorted", "coo_shuffled", "csr32", "csr16"])
def test_scatter_takes_the_inter_statistics_of_the_oe_pass(mode):
    """hia_scatter_*_sym_stats: the inter-chromosome statistics the fill kernel takes from its row
    segments equal those of the expected pass over the dense map (counts and minima exactly, sums to
    fp64 rounding), for row-sorted COO and CSR; an unsorted COO falls back (valid = 0) and the expected
    pass computes them. The O/E map that follows is the same to 1e-6."""
    from hiarch_b200 import ops
    lens, bs = [260, 33, 140, 75], 100000
    rows, cols, vals, seps = synth.synth_coo(lens, seed=17, lam_inter=1.5)
    n = sum(lens)
    # duplicates (summed before the statistics are taken) and an explicit zero entry
    rows = np.r_[rows, rows[:40], rows[100]]
    cols = np.r_[cols, cols[:40], cols[100]]
    vals = np.r_[vals, 0.5 * vals[:40], -vals[100]]
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    layout = ops.GenomeLayout(seps)
    ref_state = ops.OEState(layout)
    dense = ops.scatter_coo_sym(_dev(rows, torch.int32), _dev(cols, torch.int32), _dev(vals, torch.float32), n,
                                mtx_type=ops.MTX_TRIU_UPPER)
    ref_oe, _ = ops.obs_d_exp(dense, layout, bs, log=True, state=ref_state, from_upper=True)
    st = ops.OEState(layout)
    out = ops.alloc_map(n)
    if mode.startswith("coo"):
        r, c, v = rows, cols, vals
        if mode == "coo_shuffled":
            p = np.random.default_rng(0).permutation(rows.size)
            r, c, v = rows[p], cols[p], vals[p]
        ops.scatter_coo_sym(_dev(r, torch.int32), _dev(c, torch.int32), _dev(v, torch.float32), n,
                            mtx_type=ops.MTX_TRIU_UPPER, layout=layout, out=out, oe_state=st)
    else:
        rp, cc, vv = ops.coo_to_csr_host(rows, cols, vals, n, cols16=(mode == "csr16"))
        ops.scatter_csr_sym(rp.cuda(), cc.cuda(), vv.cuda(), n, mtx_type=ops.MTX_TRIU_UPPER, out=out, layout=layout,
                            oe_state=st)
    torch.cuda.synchronize()
    assert int(st.inter_valid.item()) == (0 if mode == "coo_shuffled" else 1)
    iu = np.triu_indices(n)
    assert np.array_equal(out.cpu().numpy()[iu], dense.cpu().numpy()[iu]) or mode == "coo_shuffled"
    if mode != "coo_shuffled":
        assert torch.equal(st.inter_nnz, ref_state.inter_nnz)
        assert torch.equal(st.inter_minpos, ref_state.inter_minpos)
        assert np.allclose(st.inter_sum.cpu().numpy(), ref_state.inter_sum.cpu().numpy(), rtol=1e-6, atol=0)
    got_oe, _ = ops.obs_d_exp(out, layout, bs, log=True, state=st, from_upper=True, inter_from_scatter=True)
    assert float((got_oe - ref_oe).abs().max()) <= 1e-6
