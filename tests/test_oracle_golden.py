"""Pin the CPU oracle (oracle/hia_oracle.py) against golden vectors produced by the REAL
reference (tests/golden/gen_golden.py). CPU only."""
import numpy as np
import pytest

from oracle import hia_oracle as O
from hiarch_b200 import synth


def _seps(g):
    return synth.chr_seps_from_lens([int(x) for x in g["in.lens"]])


@pytest.mark.parametrize("case", ["g3chr", "g1chr", "gquirk"])
def test_scatter_and_oe_chain(golden, case):
    g = golden(case)
    n = int(g["in.lens"].sum())
    bs = int(g["in.bin_size"])
    dense = O.scatter_coo_sym(g["in.rows"], g["in.cols"], g["in.vals"], n)
    if "dense_all" in g:
        assert np.array_equal(dense, g["dense_all"])
    seps = _seps(g)
    oe, exps = O.obs_d_exp(dense, seps, bs, return_expected=True)
    assert np.abs(oe - g["oe"]).max() == 0.0
    assert np.abs(np.asarray(exps[-1]) - g["expected_last_block"]).max() == 0.0
    assert np.abs(O.log_map(oe) - g["oe_log"]).max() == 0.0
    if "oe_noclamp" in g:
        assert np.abs(O.obs_d_exp(dense, seps, bs, counts_must_decrese=False) - g["oe_noclamp"]).max() == 0.0


def test_sgd_tables(golden):
    g = golden("gsgd")
    for bs, n in g["shapes"]:
        sd = O.sgd_size_dict(int(bs), int(n))
        look = np.array([-1 if O.sgd_lookup(sd, d, int(bs)) is None else O.sgd_lookup(sd, d, int(bs))
                         for d in range(int(n))], dtype=np.int32)
        assert np.array_equal(look, g[f"lookup.{bs}.{n}"]), (bs, n)
    # the quirk really triggers: bin_size=100 < n=350 -> d=100,200,300 map to bins 0,1,2
    look = g["lookup.100.350"]
    assert look[100] == 0 and look[200] == 1 and look[300] == 2 and look[101] == -1


def test_clip_zscore(golden):
    g = golden("g3chr")
    t = O.tid_off_outliers(g["oe_log"])
    assert np.abs(t - g["tid_off"]).max() == 0.0
    z = O.norm_to_zero_mean(t)
    assert np.abs(z - g["zscore"]).max() == 0.0
    assert np.abs(O.norm_de_mtx(t) - g["norm_de"]).max() == 0.0


@pytest.mark.parametrize("case", ["g3chr", "g1chr"])
def test_checkerboard(golden, case):
    g = golden(case)
    n = int(g["in.lens"].sum())
    names = [str(x) for x in g["in.names"]]
    dense = O.scatter_coo_sym(g["de_file.rows"], g["de_file.cols"], g["de_file.vals"], n,
                              has_neg=True, seps=_seps(g))
    s0, e0 = _seps(g)[0]
    assert np.array_equal(dense[s0:e0 + 1, s0:e0 + 1], g["checker.block0"])
    adj = O.cosine_distance(g["checker.block0"])
    assert np.abs(adj - g["checker.adj0"]).max() < 1e-13          # GEMM restatement vs pdist
    table = O.main_local(dense, _seps(g), names)
    assert [t[0] for t in table] == list(g["checker.accord_chro"])
    assert [t[1] for t in table] == list(g["checker.other_chro"])
    assert np.allclose([t[2] for t in table], g["checker.accord"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("case", ["g3chr", "g1chr"])
def test_gf_s1_filter_and_profiles(golden, case):
    g = golden(case)
    n = int(g["in.lens"].sum())
    seps = _seps(g)
    dense = O.scatter_coo_sym(g["de_file.rows"], g["de_file.cols"], g["de_file.vals"], n,
                              has_neg=True)
    filt = O.main_filter(dense, seps)
    assert np.abs(filt - g["filt"]).max() < 1e-12
    high = O.keep_high_cons(g["filt"])
    assert np.abs(high - g["high"]).max() < 1e-13
    ir = O.inter_rowsum(high, seps)
    assert np.abs(ir - g["inter_rowsum"]).max() < 1e-13
    assert np.abs(O.smooth_rowsum(ir, seps) - g["inter_rowsum_smooth"]).max() < 1e-13
    if "low" in g:
        low = O.keep_high_cons(g["filt"], reverse=True)
        assert np.abs(low - g["low"]).max() < 1e-13
        ia = O.intra_rowsum(low, seps)
        assert np.abs(ia - g["intra_rowsum"]).max() < 1e-13
        assert np.abs(O.smooth_rowsum(ia, seps) - g["intra_rowsum_smooth"]).max() < 1e-13


def test_intra_x(golden):
    g = golden("g1chr")
    r = O.intra_x(g["filt"])
    assert np.abs(r - g["intra_x"]).max() < 1e-12


def test_gf_s2_images_and_scores(golden):
    g = golden("g3chr")
    n = int(g["in.lens"].sum())
    seps = _seps(g)
    names = [str(x) for x in g["in.names"]]
    # GF_S2 densifies with scipy .todense(): absent cells are 0, not min-filled
    dense = O.scatter_coo_sym(g["filt_file.rows"], g["filt_file.cols"], g["filt_file.vals"], n,
                              has_neg=False)
    tmpl = np.load("hiarch_b200/data/ave_maps.npz")
    for method in ("inter_rowsum", "no_anchor"):
        used = {}
        chrs = g[f"anchors.{method}.chr"]
        types = g[f"anchors.{method}.type"]
        cens = g[f"anchors.{method}.cen"]
        for c, t, cen in zip(chrs, types, cens):
            if t == "used" and str(c) not in used:
                used[str(c)] = int(cen) - seps[names.index(str(c))][0]
        intra, inter = [], []
        for a, (s1, e1) in enumerate(seps):
            for b, (s2, e2) in enumerate(seps):
                img = O.gf_image(dense[s1:e1 + 1, s2:e2 + 1], a == b,
                                 used.get(names[a]), used.get(names[b]))
                (intra if a == b else inter).append(img)
        assert np.abs(np.vstack(intra) - g[f"gf.{method}.intra_xs"]).max() < 1e-12
        assert np.abs(np.vstack(inter) - g[f"gf.{method}.inter_xs"]).max() < 1e-12
        si = O.template_scores(np.vstack(intra), O.normalise_templates(tmpl["intra"]))
        # golden scores are melted column-major: (type0: all names), (type1: ...), ...
        assert np.abs(si.T.reshape(-1) - g[f"gf.{method}.intra_score"]).max() < 1e-13
        so = O.template_scores(np.vstack(inter), O.normalise_templates(tmpl["inter"]), do_trans=True)
        assert np.abs(so.T.reshape(-1) - g[f"gf.{method}.inter_score"]).max() < 1e-13


def test_topk_eig_is_selfconsistent(golden):
    """No reference implementation exists (parity unpinned): check the oracle's own contract."""
    g = golden("g3chr")
    C = O.pearson(g["norm_de"])
    w, v = O.topk_eig(C, 3)
    assert np.all(np.diff(w) <= 0)
    assert np.abs(C @ v - v * w).max() < 1e-9


@pytest.mark.parametrize("case", ["gnormdis", "gnormdis_db3"])
def test_main_denoise_and_normdis_chain(golden, case):
    """NormDis after the O/E map (oe_map/main_get_oe.py:40-50): main_denoise (clip, per-block
    [wavelet ->] median, see tests/golden/gen_golden_normdis.py) -> norm_de_mtx, and the %.2f text
    the step writes -- the oracle against the REAL reference's output. `gnormdis`: wavelet = identity;
    `gnormdis_db3`: the reference ran with oracle/wavelet_oracle.py standing in for the absent
    skimage call (this pins the composition around the wavelet, not the wavelet's arithmetic)."""
    g = golden(case)
    if case == "gnormdis":
        assert str(g["wavelet"]) == "identity"
        wavelet = None
    else:
        from oracle import wavelet_oracle
        assert str(g["wavelet"]).startswith("db3")
        wavelet = wavelet_oracle.denoise_wavelet_db3_visushrink
    lens = [int(x) for x in g["pro.lens"]]
    seps, s = [], 0
    for L in lens:
        seps.append((s, s + L - 1))
        s += L
    den = O.main_denoise(g["ode.dense"], seps, wavelet=wavelet)
    assert np.abs(den - g["denoised.dense"]).max() == 0.0
    de = O.norm_de_mtx(den)
    assert np.abs(de - g["de_ode.dense"]).max() < 1e-12
    # SpsHiCMtx.to_output (new_hic_class.py:2267-2297): upper triangle, "%d\t%d\t%.2f"
    up = np.triu(de)
    r, c = np.nonzero(up)
    assert np.array_equal(r, g["de_ode.rows"]) and np.array_equal(c, g["de_ode.cols"])
    txt = np.array([float("%.2f" % v) for v in up[r, c]])
    assert np.array_equal(txt, g["de_ode.vals"])


def _seps_of(lens):
    out, s = [], 0
    for L in lens:
        out.append((s, s + int(L) - 1))
        s += int(L)
    return out


def test_option_and_edge_variants(golden):
    """Oracle vs the REAL reference on the variants the main fixtures do not reach
    (tests/golden/gen_golden_options.py): only_intra (with and without the clamp), log of a map with
    zero blocks, ragged chromosomes (70 / 12 / 31 / 6 bins) at low depth, the global minimum fill of
    a has_neg file in the pipeline's order (explicit 0.00 entries stay 0), main_local with
    short_dis_max_per=.3."""
    g = golden("goptions")
    sa, sr = _seps_of(g["a.lens"]), _seps_of(g["r.lens"])
    oi = O.obs_d_exp(g["a.dense"], sa, 400000, only_intra=True)
    assert np.array_equal(oi, g["oe.only_intra"])
    assert np.array_equal(O.obs_d_exp(g["a.dense"], sa, 400000, counts_must_decrese=False, only_intra=True),
                          g["oe.noclamp_only_intra"])
    assert np.array_equal(O.log_map(oi), g["log.only_intra"])
    oer = O.obs_d_exp(g["r.dense"], sr, 1000000)
    assert np.array_equal(oer, g["oe.ragged"])
    assert np.array_equal(O.log_map(oer), g["log.ragged"])
    assert int(g["dense.n_explicit_zero"]) > 0
    n = int(g["a.lens"].sum())
    dense = O.scatter_coo_sym(g["de.rows"], g["de.cols"], g["de.vals"], n, has_neg=True)
    assert np.array_equal(dense, g["dense.global_min"])
    blockwise = O.scatter_coo_sym(g["de.rows"], g["de.cols"], g["de.vals"], n, has_neg=True, seps=sa)
    table = O.main_local(blockwise, sa, ["chr1", "chr2", "chr3"], short_dis_max_per=.3)
    assert [t[0] for t in table] == list(g["checker.sd30.accord_chro"])
    assert [t[1] for t in table] == list(g["checker.sd30.other_chro"])
    assert np.abs(np.array([t[2] for t in table]) - g["checker.sd30.accord"]).max() < 1e-12
