"""Golden files of the WHOLE one_click_pipeline.sh chain (BASELINE.json config C1 shape: the five step
scripts of publish/HiArch chained through their text files) from the REAL reference:

    NormDis.norm_dis -> CorrectMap.correct_map -> complex.src.main_local.local_io (Checkerboard.py)
    -> GF_S1_get_center.anchors_io (anchor_method 'no_anchor', the pipeline's default) -> GF_S2_get_score.get_confor

exactly as `publish/HiArch/one_click_pipeline.sh:40-169` launches them (same arguments, same file names),
with the plot calls no-op'd and wavelet = identity (oracle/ref_shim.py: scikit-image is absent and
unpinned). Every output file is stored: .mtx files as (rows, cols, vals) arrays, the small tables as text.

    python tests/golden/gen_golden_chain.py                 -> tests/golden/gchain.npz
    python tests/golden/gen_golden_chain.py --wavelet db3   -> tests/golden/gchain_db3.npz
        (the reference's own code with oracle/wavelet_oracle.py standing in for the one absent call)

The input is NOT stored: `hiarch_b200.synth.write_sample(**CASE)` regenerates it at test time.
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

CASE = dict(names=["chr1", "chr2", "chr3"], lens=[420, 300, 180], bin_size=100000, seed=77, lam_inter=6.0)
NAME = "sample_normalized"


def mtx_arrays(path):
    a = np.loadtxt(path, ndmin=2)
    return a[:, 0].astype(np.int32), a[:, 1].astype(np.int32), a[:, 2]


def main():
    db3 = sys.argv[1:] == ["--wavelet", "db3"]
    assert db3 or not sys.argv[1:], sys.argv
    ref_shim.install()
    ref_shim.set_wavelet("db3" if db3 else "identity")
    import NormDis
    import CorrectMap
    import GF_S1_get_center
    import GF_S2_get_score
    from complex.src.main_local import local_io
    CorrectMap.heatmap_plot = lambda *a, **k: None
    np.random.seed(0)                                  # main_local draws from np.random for its plot choice
    g = {"wavelet": np.array("db3 (oracle/wavelet_oracle.py)" if db3 else "identity")}
    with tempfile.TemporaryDirectory() as tmp:
        mtx, bed = synth.write_sample(f"{tmp}/sample", **CASE)
        d = {k: f"{tmp}/result/{k}/mtx" for k in ("normDis", "correctMap", "checkerBoard", "globalFolding_s1",
                                                   "globalFolding_s2")}
        for v in d.values():
            os.makedirs(v)
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            # one_click_pipeline.sh:40-46  (-df True -cmd True)
            NormDis.norm_dis(mtx, bed, f"{d['normDis']}/{NAME}", f"{tmp}/fig/{NAME}", "True", "True")
            # :71-77  (-ac <empty> -drc True)
            CorrectMap.correct_map(f"{d['normDis']}/{NAME}.de_ode.mtx", f"{d['normDis']}/{NAME}.window.bed",
                                   f"{d['correctMap']}/{NAME}", [], f"{tmp}/fig/{NAME}", "True")
            # :102-107
            local_io(sps_file=f"{d['correctMap']}/{NAME}.clean_de_ode.mtx",
                     index_file=f"{d['correctMap']}/{NAME}.clean_window.bed",
                     out_file=f"{d['checkerBoard']}/{NAME}.checkerBoard", fig_output=None, short_dis_max_per=.15)
            # :132-139  (-am no_anchor -ac <empty> -ue True)
            GF_S1_get_center.anchors_io(index_file=f"{d['correctMap']}/{NAME}.clean_window.bed",
                                        output=f"{d['globalFolding_s1']}/{NAME}",
                                        sps_file=f"{d['correctMap']}/{NAME}.clean_de_ode.mtx",
                                        fig_output=f"{tmp}/fig/{NAME}", use_exist_filt="True",
                                        anchor_method="no_anchor", anchor_change=None)
            # :164-169
            GF_S2_get_score.get_confor(f"{d['globalFolding_s1']}/{NAME}.filt_ode.mtx",
                                       f"{d['globalFolding_s1']}/{NAME}.window.bed",
                                       f"{d['globalFolding_s1']}/{NAME}.anchors.txt", d["globalFolding_s2"],
                                       f"{tmp}/fig2")
        g["log"] = np.array(log.getvalue())
        for tag, path in (("ode", f"{d['normDis']}/{NAME}.ode.mtx"), ("de_ode", f"{d['normDis']}/{NAME}.de_ode.mtx"),
                          ("clean_de_ode", f"{d['correctMap']}/{NAME}.clean_de_ode.mtx"),
                          ("filt_ode", f"{d['globalFolding_s1']}/{NAME}.filt_ode.mtx")):
            r, c, v = mtx_arrays(path)
            g[f"{tag}.rows"], g[f"{tag}.cols"], g[f"{tag}.vals"] = r, c, v
        for tag, path in (("normDis.window_bed", f"{d['normDis']}/{NAME}.window.bed"),
                          ("clean_window_bed", f"{d['correctMap']}/{NAME}.clean_window.bed"),
                          ("checkerBoard", f"{d['checkerBoard']}/{NAME}.checkerBoard"),
                          ("s1.window_bed", f"{d['globalFolding_s1']}/{NAME}.window.bed"),
                          ("anchors_txt", f"{d['globalFolding_s1']}/{NAME}.anchors.txt")):
            g[tag] = np.array(open(path).read())
        confor = sorted(f for f in os.listdir(d["globalFolding_s2"]) if f.endswith(".txt"))
        g["confor.files"] = np.array(confor)
        for f in confor:
            g[f"confor.{f}"] = np.array(open(os.path.join(d["globalFolding_s2"], f)).read())
    np.savez_compressed(f"{HERE}/gchain_db3.npz" if db3 else f"{HERE}/gchain.npz", **g)
    print({k: getattr(v, "shape", None) for k, v in g.items()})
    print(str(g["log"])[-1500:])
    print(str(g["checkerBoard"]))
    print(str(g["anchors_txt"])[:600])
    for f in confor:
        print(f, str(g[f"confor.{f}"])[:600])


if __name__ == "__main__":
    main()
