"""CPU: f1/f3 host logic -- index surgery and correct_map against the reference's outputs."""
import numpy as np
import pandas as pd

from hiarch_b200 import synth


def _index(names, lens, bs):
    from hiarch_b200.new_hic_class import GenomeIndex
    rows = synth.window_table(names, lens, bs)
    return GenomeIndex(pd.DataFrame(rows, columns=['chr', 'start', 'end', 'index']))


def test_genome_index_condense_keep_and_sub():
    gi = _index(["a", "b", "c"], [10, 7, 3], 100)
    g2, trans = gi.condense(4)
    assert g2.chr_lens == {"a": 3, "b": 2, "c": 1} and g2.chros == ["a", "b", "c"]
    assert list(trans) == [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 5]
    assert list(g2.window_df['start']) == [0, 400, 800, 0, 400, 0]
    assert list(g2.window_df['end']) == [400, 800, 1000, 400, 700, 300]
    g3, old = gi.keep_chros(["c", "a"], return_old_index=True)
    assert g3.chros == ["c", "a"] and list(old) == [17, 18, 19] + list(range(10))
    g4 = gi.keep_long_chros(5)
    assert g4.chros == ["a", "b"]
    g5 = gi.get_sub_gen_index(idxes=np.array([1, 2, 12, 19]))
    assert g5.chr_lens == {"a": 2, "b": 1, "c": 1} and list(g5.window_df['ori_index']) == [1, 2, 12, 19]


def test_correct_map_matches_reference_files(golden, tmp_path):
    from hiarch_b200.preprocess import correct_map
    g3, gc = golden("g3chr"), golden("gcorrect")
    names = [str(x) for x in g3["in.names"]]
    lens = [int(x) for x in g3["in.lens"]]
    bed = f"{tmp_path}/c.window.bed"
    with open(bed, "w") as fh:
        for (nm, s, e, i) in synth.window_table(names, lens, int(g3["in.bin_size"])):
            fh.write(f"{nm}\t{s}\t{e}\t{i}\n")
    de = f"{tmp_path}/c.de_ode.mtx"
    np.savetxt(de, np.vstack([g3["de_file.rows"], g3["de_file.cols"], g3["de_file.vals"]]).T, fmt="%d\t%d\t%.2f")
    for tag, ab, reorder in (("reorder", None, True), ("drop2", [2], True), ("keep", None, False), ("dropneg", [-1], True)):
        out = f"{tmp_path}/cm_{tag}"
        correct_map(de, bed, out, ab_chrs=ab, do_reorder_chrs=reorder)
        same_bed = open(out + ".clean_window.bed").read() == str(gc[f"{tag}.bed"])
        same_mtx = open(out + ".clean_de_ode.mtx").read() == str(gc[f"{tag}.mtx"])        # byte for byte
        assert same_bed and same_mtx, tag       # booleans: pytest must not diff 45 k-line strings


# ---- preprocess_mtx: the host decisions (index maps, folds, gates) with the three device statistics
# emulated in numpy by a TEST-ONLY subclass; the CUDA kernels themselves are checked against the same
# golden files in tests/test_gpu_parity.py::test_preprocess_mtx_matches_reference.
def _numpy_remapped():
    import torch
    from hiarch_b200 import preprocess as P

    class NumpyRemapped(P._Remapped):
        def __init__(self, sps_mtx, device="cpu"):
            coo = sps_mtx.matrix.tocoo()
            self.device = "cpu"
            self.r, self.c, self.v = coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(np.float64)
            assert sps_mtx.mtx_type == 'triu'
            self.n_fine = sps_mtx.shape[0]
            self.index = sps_mtx.row_window
            self.map = np.arange(self.n_fine, dtype=np.int64)

        def _sym(self, m):
            up = self.c >= self.r
            r, c, v = self.r[up], self.c[up], self.v[up]
            off = r != c
            rr = np.concatenate([m[r], m[c[off]]]); cc = np.concatenate([m[c], m[r[off]]])
            vv = np.concatenate([v, v[off]])
            ok = (rr >= 0) & (cc >= 0)
            return rr[ok], cc[ok], vv[ok]

        def dense(self):
            n = int(self.index.max_idx) + 1
            rr, cc, vv = self._sym(self.map)
            d = np.zeros((n, n))
            np.add.at(d, (rr, cc), vv)
            return torch.from_numpy(d.astype(np.float32))

        def bin_sums(self):
            n = int(self.index.max_idx) + 1
            rr, cc, vv = self._sym(self.map)
            return np.bincount(rr, weights=vv, minlength=n)

        def bin_stats(self, dense):
            d = dense.numpy().astype(np.float64)
            return d.sum(axis=1), (d != 0).sum(axis=1).astype(np.int64)

        def intra_nnz_fine(self):
            rr, cc, vv = self._sym(np.arange(self.n_fine))
            out = {}
            for chro in self.index.chros:
                s, e = self.index.chr_seps[chro]
                out[chro] = int(((rr >= s) & (rr <= e) & (cc >= s) & (cc <= e) & (vv != 0)).sum())
            return out

    return P, NumpyRemapped


PRE_CASES = {
    "gpre_fold": dict(names=["chr1", "chr2"], lens=[1250, 700], bs=100000, seed=21, depth=1.0),
    "gpre_sparse": dict(names=["chrA", "chrB"], lens=[400, 300], bs=50000, seed=22, depth=0.004),
    "gpre_none": dict(names=["chrA", "chrB"], lens=[400, 300], bs=50000, seed=22, depth=0.0015),
}


def check_preprocess_case(name, golden, tmp_path, preprocess_mtx):
    """Shared by the CPU (emulated statistics) and the GPU (CUDA kernels) tests."""
    import contextlib
    import io
    from hiarch_b200.new_hic_class import read_matrix
    c, g = PRE_CASES[name], golden(name)
    mtx, bed = synth.write_sample(f"{tmp_path}/{name}", c["names"], c["lens"], c["bs"], seed=c["seed"],
                                  depth=c["depth"])
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        pro = preprocess_mtx(read_matrix(mtx, bed))
    strip = lambda t: [ln for ln in str(t).split("\n") if ln and not ln.startswith("Warning: Condense in trunc")]
    assert strip(log.getvalue()) == strip(g["log"])           # every decision: folds, kept bins, ratios
    if "none" in g:
        assert pro is None
        return
    df = pro.row_window.window_df
    assert [str(x) for x in df["chr"]] == [str(x) for x in g["chr"]]
    for col in ("start", "end", "index"):
        assert np.array_equal(df[col].values.astype(np.int64), g[col]), col      # bit-exact bin tables
    dense = pro.dmatrix.cpu().numpy()
    ref = g["dense32"]
    assert dense.shape == ref.shape
    assert np.array_equal(dense != 0, ref != 0)
    assert np.abs(dense - ref).max() <= 2e-6 * np.abs(ref).max()     # fp32 sums of <= fold^2 entries, any order
    assert np.allclose(dense.astype(np.float64).sum(axis=1), g["row_sum"], rtol=1e-6)


def test_preprocess_mtx_host_decisions(golden, tmp_path, monkeypatch):
    P, NumpyRemapped = _numpy_remapped()
    monkeypatch.setattr(P, "_Remapped", NumpyRemapped)
    for name in PRE_CASES:
        check_preprocess_case(name, golden, tmp_path, P.preprocess_mtx)
