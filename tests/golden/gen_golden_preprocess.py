"""Golden vectors for f1 (preprocess_mtx) and f3 (correct_map) from the REAL reference.

    python tests/golden/gen_golden_preprocess.py

The inputs are NOT stored: they are regenerated deterministically by
`hiarch_b200.synth.synth_coo(lens, seed, depth=...)` at test time.
  gpre_fold    2 chromosomes (1250 / 700 bins @ 100 kb): rough_concat folds by 2, then the 5 %
               lowest-coverage bins are dropped (dense, non-zero ratio > .95 on the first try)
  gpre_sparse  2 chromosomes (400 / 300 bins) at 0.4 % depth: the non-zero ratio gate fails and
               the condense-search loop (S1_preprocess.py:66-83) picks fold 13 (last candidate)
  gpre_none    the same at 0.15 % depth: no fold reaches 95 % non-zero -> preprocess_mtx returns None
"""
import contextlib
import io
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from hiarch_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore")

CASES = {
    "gpre_fold": dict(names=["chr1", "chr2"], lens=[1250, 700], bs=100000, seed=21, depth=1.0),
    "gpre_sparse": dict(names=["chrA", "chrB"], lens=[400, 300], bs=50000, seed=22, depth=0.004),
    "gpre_none": dict(names=["chrA", "chrB"], lens=[400, 300], bs=50000, seed=22, depth=0.0015),
}


def main():
    ref_shim.install()
    from iutils.read_matrix import read_matrix
    from oe_map.S1_preprocess import preprocess_mtx
    from CorrectMap import correct_map
    with tempfile.TemporaryDirectory() as tmp:
        for name, c in CASES.items():
            mtx, bed = synth.write_sample(f"{tmp}/{name}", c["names"], c["lens"], c["bs"], seed=c["seed"],
                                          depth=c["depth"])
            sps_mtx = read_matrix(mtx, bed)
            log = io.StringIO()
            with contextlib.redirect_stdout(log):
                pro = preprocess_mtx(sps_mtx)
            g = {"log": np.array(log.getvalue())}
            if pro is None:
                g["none"] = np.array(1)
            else:
                df = pro.row_window.window_df
                g["chr"] = np.array([str(x) for x in df["chr"]])
                g["start"] = df["start"].values.astype(np.int64)
                g["end"] = df["end"].values.astype(np.int64)
                g["index"] = df["index"].values.astype(np.int64)
                dense = np.array(pro.matrix.todense())
                g["dense32"] = dense.astype(np.float32)
                g["row_sum"] = dense.sum(axis=1)
                g["nonzero_ratio"] = np.array(pro.get_nonzero_ratio())
            np.savez_compressed(f"{HERE}/{name}.npz", **g)
            print(name, {k: getattr(v, "shape", None) for k, v in g.items()}, "\n", log.getvalue()[-400:])

        # f3: CorrectMap on the g3chr de_ode file: drop chromosome 2, reorder by length
        g3 = dict(np.load(f"{HERE}/g3chr.npz"))
        names = [str(x) for x in g3["in.names"]]
        lens = [int(x) for x in g3["in.lens"]]
        bed = f"{tmp}/c.window.bed"
        with open(bed, "w") as fh:
            for (nm, s, e, i) in synth.window_table(names, lens, int(g3["in.bin_size"])):
                fh.write(f"{nm}\t{s}\t{e}\t{i}\n")
        de = f"{tmp}/c.de_ode.mtx"
        np.savetxt(de, np.vstack([g3["de_file.rows"], g3["de_file.cols"], g3["de_file.vals"]]).T, fmt="%d\t%d\t%.2f")
        out = {}
        import CorrectMap as cm
        cm.heatmap_plot = lambda *a, **k: None
        for tag, ab, reorder in (("reorder", None, True), ("drop2", [2], True), ("keep", None, False), ("dropneg", [-1], True)):
            o = f"{tmp}/cm_{tag}"
            with contextlib.redirect_stdout(io.StringIO()):
                correct_map(de, bed, o, ab_chrs=ab, do_reorder_chrs=reorder)
            out[f"{tag}.mtx"] = np.array(open(f"{o}.clean_de_ode.mtx").read())
            out[f"{tag}.bed"] = np.array(open(f"{o}.clean_window.bed").read())
        np.savez_compressed(f"{HERE}/gcorrect.npz", **out)
        print("gcorrect", {k: len(str(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
