"""Golden vectors from the REAL reference for the option / edge variants the main fixtures do not
reach (small maps so the fixture stays small):

  oe.only_intra / oe.noclamp_only_intra   HiCMtx.obs_d_exp(only_intra=True[, counts_must_decrese=False])
  oe.ragged                               chromosomes of 70 / 12 / 31 / 6 bins at low depth: all-zero
                                          diagonals and an all-zero inter block
  log.*                                   ArrayHiCMtx.log on maps with zeros (has_neg False)
  checker.sd30                            main_local(short_dis_max_per=.3) on a min-filled map
  dense.global_min                        SpsHiCMtx.to_dense() + to_all of a has_neg map (global minimum fill,
                                          explicit 0.00 entries stay 0)

    python tests/golden/gen_golden_options.py        -> tests/golden/goptions.npz
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

A = dict(names=["chr1", "chr2", "chr3"], lens=[90, 70, 60], bin_size=400000, seed=51)
R = dict(names=["c1", "c2", "c3", "c4"], lens=[70, 12, 31, 6], bin_size=1000000, seed=52, depth=0.002, lam_inter=0.02)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    nhc = ref_shim.install()
    from iutils.read_matrix import read_matrix
    from complex.src.main_local import main_local
    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        mtx, bed_a = synth.write_sample(f"{tmp}/a", **A)
        allm = quiet(read_matrix, mtx, bed_a).to_all()
        g["a.dense"] = np.array(allm.matrix.todense())
        g["a.lens"] = np.array(A["lens"])
        g["oe.only_intra"] = quiet(allm.obs_d_exp, mode="sgd_mean", k=.2, only_intra=True).matrix
        g["oe.noclamp_only_intra"] = quiet(allm.obs_d_exp, mode="sgd_mean", k=.2, only_intra=True,
                                           counts_must_decrese=False).matrix
        oi = quiet(allm.obs_d_exp, mode="sgd_mean", k=.2, only_intra=True)
        g["log.only_intra"] = quiet(oi.log, verbose=False).matrix          # inter blocks are 0 -> filled

        mtx, bed = synth.write_sample(f"{tmp}/r", **R)
        allr = quiet(read_matrix, mtx, bed).to_all()
        g["r.dense"] = np.array(allr.matrix.todense())
        g["r.lens"] = np.array(R["lens"])
        oer = quiet(allr.obs_d_exp, mode="sgd_mean", k=.2)
        g["oe.ragged"] = oer.matrix
        g["log.ragged"] = quiet(oer.log, verbose=False).matrix

        # a has_neg map through the %.2f wire format: min fill + checkerboard with a wider band
        from oe_map.S2_oe import oe_map_core
        from oe_map.S4_norm_value import norm_de_mtx
        nde = norm_de_mtx(quiet(oe_map_core, allm).to_dense())
        de_file = f"{tmp}/a.clean_de_ode.mtx"
        quiet(nde.to_sps().to_output, de_file, out_window=False)
        de = np.loadtxt(de_file)
        g["de.rows"], g["de.cols"], g["de.vals"] = de[:, 0].astype(np.int32), de[:, 1].astype(np.int32), de[:, 2]
        sm = quiet(nhc.read_sps_file, de_file, bed_a)
        assert sm.has_neg
        # the pipeline's order (S1_1_filtering.py:22-23, main_local.py:112-115): to_dense FIRST (absent
        # cells get the minimum, explicit "0.00" entries stay 0), then to_all. The other order
        # (to_all on the sparse matrix first) drops the stored zeros in scipy's `+` and min-fills them.
        fm = quiet(read_matrix, de_file, bed_a)                     # mtx_type 'triu', has_neg from the file name
        assert fm.has_neg and fm.mtx_type == 'triu'
        d = quiet(fm.to_dense)
        quiet(d.to_all, inplace=True)
        g["dense.global_min"] = np.asarray(d.matrix, dtype=np.float64)
        g["dense.n_explicit_zero"] = np.array(int((de[:, 2] == 0).sum()))
        np.random.seed(0)
        table = quiet(main_local, sm, short_dis_max_per=.3)
        g["checker.sd30.accord_chro"] = np.array([str(x) for x in table["accord_chro"]])
        g["checker.sd30.other_chro"] = np.array([str(x) for x in table["other_chro"]])
        g["checker.sd30.accord"] = np.asarray(table["accord"], dtype=np.float64)
    np.savez_compressed(f"{HERE}/goptions.npz", **g)
    print({k: getattr(v, "shape", None) for k, v in g.items()})


if __name__ == "__main__":
    main()
