"""Generate golden input/output vectors by running the REAL HiArch reference.

Run in the authoring container only (`/root/reference` mounted read-only):

    python tests/golden/gen_golden.py

It imports the reference through `oracle/ref_shim.py` (plot stubs; wavelet = identity and NOT
on any path recorded here), feeds it seeded synthetic maps written in the reference's own wire
format, and dumps the reference's outputs at every hot-path boundary into
`tests/golden/*.npz`.  The committed fixtures travel to the GPU box; this script and the
reference do not need to.

Cases
  g3chr   3 chromosomes (120/100/80 bins @ 400 kb): O/E chain, clip/z-score, checkerboard
          (through `.mtx` %.2f files), GF_S1 filter + profiles + anchors (3 methods),
          GF_S2 images + scores.
  g1chr   1 chromosome, 400 bins @ 100 kb, Rabl/X pattern: O/E, cosine distance sample,
          get_local, intra_x response + anchors.
  gquirk  1 chromosome, 350 bins with bin_size = 100 bp (< n): the `dis >= bin_size` SGD quirk
          (publish/new_hic_class.py:865-875).
  gsgd    SGD tables for a list of (bin_size, n) shapes incl. hg38 chr1 at 100/50/25/10 kb.
"""
import os
import sys
import io
import contextlib
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from hiarch_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def anchors_to_arrays(loc_df, prefix):
    out = {}
    for col in loc_df.columns:
        v = loc_df[col].values
        if col in ("type", "chr"):
            out[f"{prefix}.{col}"] = np.array([str(x) for x in v])
        else:
            out[f"{prefix}.{col}"] = np.asarray(v, dtype=np.float64)
    out[f"{prefix}.columns"] = np.array(list(loc_df.columns))
    return out


def case_genome(nhc, tmp, name, names, lens, bs, seed, x_shape=0.0, do_gf=True, methods=()):
    from iutils.read_matrix import read_matrix
    from oe_map.S2_oe import oe_map_core
    from oe_map.S3_denoise import tid_off_outliers
    from oe_map.S4_norm_value import norm_de_mtx, norm_to_zero_mean, remove_outliers

    g = {}
    mtx, bed = synth.write_sample(f"{tmp}/{name}", names, lens, bs, seed=seed, x_shape=x_shape)
    coo = np.loadtxt(mtx)
    g["in.rows"] = coo[:, 0].astype(np.int32)
    g["in.cols"] = coo[:, 1].astype(np.int32)
    g["in.vals"] = coo[:, 2]
    g["in.lens"] = np.array(lens)
    g["in.names"] = np.array(names)
    g["in.bin_size"] = np.array(bs)

    sps_mtx = quiet(read_matrix, mtx, bed)
    allm = sps_mtx.to_all()
    g["dense_all"] = np.array(allm.matrix.todense())

    ode = quiet(allm.obs_d_exp, mode="sgd_mean", k=.2, counts_must_decrese=True)
    g["oe"] = ode.matrix
    ode_noclamp = quiet(allm.obs_d_exp, mode="sgd_mean", k=.2, counts_must_decrese=False)
    g["oe_noclamp"] = ode_noclamp.matrix
    _, exp_last = quiet(allm.obs_d_exp, mode="sgd_mean", k=.2, return_exp_counts=True)
    g["expected_last_block"] = np.asarray(exp_last, dtype=np.float64)
    oel = quiet(oe_map_core, allm)
    g["oe_log"] = oel.matrix

    clipped = tid_off_outliers(oel.to_dense())
    g["tid_off"] = clipped.matrix.copy()
    z = norm_to_zero_mean(clipped)
    g["zscore"] = z.matrix.copy()
    nde = norm_de_mtx(clipped)
    g["norm_de"] = nde.matrix.copy()

    # ---- through the %.2f wire format (what Checkerboard / GF_S1 really read) ----
    de_file = f"{tmp}/{name}.clean_de_ode.mtx"
    quiet(nde.to_sps().to_output, de_file, out_window=False)
    de = np.loadtxt(de_file)
    g["de_file.rows"] = de[:, 0].astype(np.int32)
    g["de_file.cols"] = de[:, 1].astype(np.int32)
    g["de_file.vals"] = de[:, 2]

    from complex.src.main_local import main_local
    from complex.src.S1_get_adj_mtx import get_adj_mtx
    sm = quiet(nhc.read_sps_file, de_file, bed)
    np.random.seed(0)
    table = quiet(main_local, sm)
    g["checker.accord_chro"] = np.array([str(x) for x in table["accord_chro"]])
    g["checker.other_chro"] = np.array([str(x) for x in table["other_chro"]])
    g["checker.accord"] = np.asarray(table["accord"], dtype=np.float64)
    # adjacency of the first intra block (min-filled, symmetrised), for entry-wise checks
    blk = None
    for c1, c2, cm in sm.yield_by_chro(triu=True):
        cm = cm.to_dense()
        cm.to_all(inplace=True)
        blk = cm
        break
    g["checker.block0"] = blk.matrix.copy()
    g["checker.adj0"] = get_adj_mtx(blk).matrix

    if do_gf:
        from confor.src.S1_preprocess.S1_1_filtering import filter_io
        from confor.src.S1_preprocess.S1_2_find_anchor_main import main_find_anchor
        from confor.src.S1_preprocess.S1_2_find_anchors.U2_keep_top_cons import keep_high_cons
        from confor.src.S1_preprocess.S1_2_find_anchors.S2_1_find_by_inter_rowsum import (
            get_inter_rowsum, smooth_rowsum)
        from confor.src.S1_preprocess.S1_2_find_anchors.S2_3_find_by_intra_low import get_intra_rowsum
        out = f"{tmp}/{name}_gf"
        filt = quiet(filter_io, de_file, bed, out, None, False)
        g["filt"] = filt.matrix.copy()
        ff = np.loadtxt(f"{out}.filt_ode.mtx")
        g["filt_file.rows"] = ff[:, 0].astype(np.int32)
        g["filt_file.cols"] = ff[:, 1].astype(np.int32)
        g["filt_file.vals"] = ff[:, 2]
        fd = filt.to_dense()
        fd.to_all(inplace=True)
        high = keep_high_cons(fd)
        g["high"] = high.matrix.copy()
        ir = get_inter_rowsum(high)
        g["inter_rowsum"] = np.squeeze(np.array(ir.values))
        g["inter_rowsum_smooth"] = np.squeeze(np.array(smooth_rowsum(ir).values))
        low = keep_high_cons(fd, reverse=True)
        g["low"] = low.matrix.copy()
        ia = get_intra_rowsum(low)
        g["intra_rowsum"] = np.squeeze(np.array(ia.values))
        g["intra_rowsum_smooth"] = np.squeeze(np.array(smooth_rowsum(ia).values))
        for method in methods:
            o2 = f"{tmp}/{name}_{method}"
            try:
                quiet(main_find_anchor, filt, o2, None, anchor_method=method)
            except Exception as exc:                       # e.g. intra_x on a pattern-less map
                print(f"  [{name}] anchor method {method} raised {type(exc).__name__}: {exc}")
                continue
            import pandas as pd
            df = pd.read_table(f"{o2}.anchors.txt", header=0, dtype={"chr": str})
            g.update(anchors_to_arrays(df, f"anchors.{method}"))
            with open(f"{o2}.anchors.txt") as fh:
                g[f"anchors.{method}.text"] = np.array(fh.read())

        # ---- GF_S2 ----
        from confor.src.S2_to_train_mtx.to_train_mtx import to_train_mtx
        from confor.src.S5_sim_to_pat.to_confor import (get_n_norm_ave_maps, change_map_name,
                                                        to_confor)
        from confor.confor_class import (intra_ave_names, inter_ave_names, intra_ave_map_file,
                                         inter_ave_map_file)
        from scipy.io import loadmat
        for method in methods:
            af = f"{tmp}/{name}_{method}.anchors.txt"
            if not os.path.exists(af):
                continue
            intra_d, inter_d = quiet(to_train_mtx, f"{out}.filt_ode.mtx", f"{out}.window.bed"
                                     if os.path.exists(f"{out}.window.bed") else bed, af, name)
            if intra_d.xs is not None:
                g[f"gf.{method}.intra_xs"] = intra_d.xs
                g[f"gf.{method}.intra_isym"] = intra_d.isym_xs
                g[f"gf.{method}.intra_names"] = np.array([s.strip() for s in intra_d.out_names()])
                maps = change_map_name(get_n_norm_ave_maps(loadmat(intra_ave_map_file)), intra_ave_names)
                sc = quiet(to_confor, intra_d, maps)
                g[f"gf.{method}.intra_score"] = np.asarray(sc["score"], dtype=np.float64)
                g[f"gf.{method}.intra_score_name"] = np.array([str(x) for x in sc["name"]])
                g[f"gf.{method}.intra_score_type"] = np.array([str(x) for x in sc["type"]])
            if inter_d.xs is not None:
                g[f"gf.{method}.inter_xs"] = inter_d.xs
                g[f"gf.{method}.inter_names"] = np.array([s.strip() for s in inter_d.out_names()])
                maps = change_map_name(get_n_norm_ave_maps(loadmat(inter_ave_map_file)), inter_ave_names)
                sc = quiet(to_confor, inter_d, maps, do_trans=True)
                g[f"gf.{method}.inter_score"] = np.asarray(sc["score"], dtype=np.float64)
                g[f"gf.{method}.inter_score_name"] = np.array([str(x) for x in sc["name"]])
                g[f"gf.{method}.inter_score_type"] = np.array([str(x) for x in sc["type"]])
    return g


def case_intra_x(nhc, g):
    """intra_x response of the first intra block of the filtered map (reference get_intra_x)."""
    from confor.src.S1_preprocess.S1_2_find_anchors.S2_2_find_by_intra_x import get_intra_x
    filt = g["filt"]

    class _B:
        pass
    b = _B()
    b.matrix = filt
    b.shape = filt.shape
    g["intra_x"] = get_intra_x(b)
    return g


def case_sgd(nhc):
    g = {}
    shapes = [(400000, 593), (100000, 2500), (100000, 2490), (50000, 4980), (25000, 9959),
              (10000, 24896), (1000000, 64), (100, 350), (1000, 9500), (400000, 120),
              (400000, 100), (400000, 80), (100000, 400), (40000, 51)]
    g["shapes"] = np.array(shapes)
    for bs, n in shapes:
        sgd = nhc.ScaledGenomicDistance(s=bs, e=bs * n, bin_size=bs, k=.2)
        # scalar lookups exactly as obs_d_exp does them (with the quirk)
        look = np.array([-1 if sgd[d] is None else int(sgd[d]) for d in range(n)], dtype=np.int32)
        g[f"lookup.{bs}.{n}"] = look
    return g


def main():
    nhc = ref_shim.install()
    with tempfile.TemporaryDirectory() as tmp:
        g = case_genome(nhc, tmp, "g3chr", ["chr1", "chr2", "chr3"], [120, 100, 80], 400000,
                        seed=1, methods=("inter_rowsum", "intra_low", "no_anchor"))
        np.savez_compressed(f"{HERE}/g3chr.npz", **g)
        print("g3chr", {k: getattr(v, "shape", None) for k, v in list(g.items())[:4]}, len(g))

        g = case_genome(nhc, tmp, "g1chr", ["chr1"], [400], 100000, seed=2, x_shape=60.0,
                        methods=("intra_x", "inter_rowsum"))
        g = case_intra_x(nhc, g)
        for k in ("oe_noclamp", "tid_off", "zscore", "low", "intra_rowsum", "intra_rowsum_smooth"):
            g.pop(k, None)
        np.savez_compressed(f"{HERE}/g1chr.npz", **g)
        print("g1chr", len(g), [k for k in g if k.startswith("anchors") and k.endswith("cen")])

        g = case_genome(nhc, tmp, "gquirk", ["chrA"], [350], 100, seed=3, do_gf=False)
        g = {k: v for k, v in g.items() if k.startswith("in.") or k in
             ("oe", "oe_log", "expected_last_block", "norm_de")}
        np.savez_compressed(f"{HERE}/gquirk.npz", **g)
        print("gquirk", len(g))

    g = case_sgd(nhc)
    np.savez_compressed(f"{HERE}/gsgd.npz", **g)
    print("gsgd", len(g))

    # shipped template maps -> product data (inputs to a15; 5 x 256 float32 each, sorted by key)
    from scipy.io import loadmat
    from confor.confor_class import intra_ave_map_file, inter_ave_map_file
    data = {}
    for tag, f in (("intra", intra_ave_map_file), ("inter", inter_ave_map_file)):
        m = loadmat(f)
        keys = sorted([int(float(k)) for k in m if not k.startswith("_")])
        data[tag] = np.vstack([np.asarray(m[str(k)]).reshape(1, -1) for k in keys])
        print(tag, data[tag].shape, data[tag].dtype)
    os.makedirs(f"{ROOT}/hiarch_b200/data", exist_ok=True)
    np.savez(f"{ROOT}/hiarch_b200/data/ave_maps.npz", **data)


if __name__ == "__main__":
    main()
