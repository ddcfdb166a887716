"""Golden vectors for the NormDis step (publish/HiArch/NormDis.py -> oe_map/main_get_oe.py:10-59)
from the REAL reference, wavelet = identity (the skimage db3/VisuShrink call of
S3_denoise.py:45-47 has no implementation in this image and no pinned version -- oracle/ref_shim.py
stubs it to the identity; everything else of the step is the reference's own code):

    preprocess_mtx -> oe_map_core -> .ode.mtx (+ .window.bed)
                   -> main_denoise (clip p2/p98, per-block [wavelet=identity ->] median filter)
                   -> norm_de_mtx -> .de_ode.mtx

    python tests/golden/gen_golden_normdis.py                 -> tests/golden/gnormdis.npz
    python tests/golden/gen_golden_normdis.py --wavelet db3   -> tests/golden/gnormdis_db3.npz
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

CASE = dict(names=["chr1", "chr2", "chr3"], lens=[150, 120, 90], bin_size=100000, seed=31, lam_inter=6.0)


def main():
    db3 = sys.argv[1:] == ["--wavelet", "db3"]
    assert db3 or not sys.argv[1:], sys.argv
    ref_shim.install()
    ref_shim.set_wavelet("db3" if db3 else "identity")
    from oe_map.main_get_oe import oe_map_io
    from oe_map.S3_denoise import main_denoise
    from oe_map.S2_oe import oe_map_core
    from oe_map.S1_preprocess import preprocess_mtx
    from iutils.read_matrix import read_matrix
    with tempfile.TemporaryDirectory() as tmp:
        mtx, bed = synth.write_sample(f"{tmp}/s", **CASE)
        out = f"{tmp}/res"
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            de = oe_map_io(mtx, out, None, do_filt_chr_by_name=True, counts_must_decrese=True, index_file=bed)
        g = {"log": np.array(log.getvalue()), "wavelet": np.array("db3 (oracle/wavelet_oracle.py)" if db3 else "identity")}
        for tag in ("ode", "de_ode"):
            a = np.loadtxt(f"{out}.{tag}.mtx")
            g[f"{tag}.rows"] = a[:, 0].astype(np.int32)
            g[f"{tag}.cols"] = a[:, 1].astype(np.int32)
            g[f"{tag}.vals"] = a[:, 2]
        g["window_bed"] = np.array(open(f"{out}.window.bed").read())
        assert not os.path.exists(f"{out}.de_window.bed")
        # the dense intermediates, for the oracle pin on the CPU
        with contextlib.redirect_stdout(io.StringIO()):
            pro = preprocess_mtx(read_matrix(mtx, bed), True)
            ode = oe_map_core(pro, counts_must_decrese=True)
            den = main_denoise(ode)
        g["pro.names"] = np.array([str(c) for c in pro.row_window.chros])
        g["pro.lens"] = np.array([int(pro.row_window.chr_lens[c]) for c in pro.row_window.chros])   # map order
        g["ode.dense"] = np.asarray(ode.to_dense().matrix, dtype=np.float64)
        g["denoised.dense"] = np.asarray(den.matrix, dtype=np.float64)
        g["de_ode.dense"] = np.asarray(de.matrix, dtype=np.float64)
        np.savez_compressed(f"{HERE}/gnormdis_db3.npz" if db3 else f"{HERE}/gnormdis.npz", **g)
        print({k: getattr(v, "shape", None) for k, v in g.items()})
        print(log.getvalue())


if __name__ == "__main__":
    main()
