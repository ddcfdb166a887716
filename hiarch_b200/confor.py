"""Mirror of publish/confor (global-folding score) on the CUDA ops.

used_filter / main_filter / filter_io   publish/confor/src/S1_preprocess/S1_1_filtering.py:13-52
main_find_anchor               publish/confor/src/S1_preprocess/S1_2_find_anchor_main.py:8-45
to_train_mtx                   publish/confor/src/S2_to_train_mtx/to_train_mtx.py:75-129
to_confor / main_to_confor     publish/confor/src/S5_sim_to_pat/to_confor.py:35-93
anchors_io / get_confor        publish/HiArch/GF_S1_get_center.py:12-27, GF_S2_get_score.py:11-25
"""
import os

import numpy as np
import pandas as pd

from . import anchors as A
from . import ops
from .new_hic_class import ArrayHiCMtx, SpsHiCMtx, read_matrix

filt_mtx_suffix = '.filt_ode.mtx'
anchor_suffix = '.anchors.txt'
confor_suffix = '.confor.txt'
intra_ave_names = ops.INTRA_AVE_NAMES
inter_ave_names = ops.INTER_AVE_NAMES
X_SIZE = 16


def used_filter(sps_mtx, decimals=2):
    """S1_1_filtering.py:13-16 on one chromosome(-pair) block: box filter of size
    int(mean(chr_lens) * .1) ('reflect') + norm_to_zero_mean, through the fused fp64 scipy-order
    replica (`hia_used_filter_f64`). `decimals`: decimal quantisation of the block's values (2 for a
    map read from a %.2f .mtx; -1 = not quantised)."""
    d = sps_mtx.to_dense() if isinstance(sps_mtx, SpsHiCMtx) else sps_mtx
    size = int(np.mean(np.array(list(sps_mtx.row_window.chr_lens.values())) * .1))
    out = ops.alloc_map(d.shape[0], d.dmatrix.device, cols=d.shape[1])
    ops.used_filter(d.dmatrix, size, out, decimals=decimals, normalise=True)
    return ArrayHiCMtx(out, copy_mtx_paras=d, has_neg=True)


def main_filter(sps_map, decimals=2):
    """S1_1_filtering.py:21-32. `sps_map`: the min-filled map of a `.mtx` (values quantised to
    `decimals`)."""
    d = sps_map.to_dense() if isinstance(sps_map, SpsHiCMtx) else sps_map
    if d.mtx_type != 'all':
        d = d.to_all(inplace=False)
    layout = d.row_window.layout(d.dmatrix.device)
    out = ops.main_filter(d.dmatrix, layout, decimals=decimals)
    return ArrayHiCMtx(out, copy_mtx_paras=d, mtx_type='all', has_neg=True)


def filter_io(sps_file, index_file, output=None, fig_output=None, use_exist_filt=False):
    """S1_1_filtering.py:35-52 (resume from an existing .filt_ode.mtx when asked)."""
    if output is not None and os.path.exists(f'{output}{filt_mtx_suffix}') and use_exist_filt:
        m = read_matrix(f'{output}{filt_mtx_suffix}', index_file)
        filtered = m.to_dense()
    else:
        if sps_file is None:
            raise ValueError('Plz input sparse matrix.')
        filtered = main_filter(read_matrix(sps_file, index_file))
        if output is not None:
            filtered.to_sps().to_output(f'{output}{filt_mtx_suffix}')
    print('Filtering mtx DONE.')
    return filtered


def _bins(gen_index):
    df = gen_index.window_df
    return df['start'].values, df['end'].values


def main_find_anchor(filtered_mtx, output, fig_output, anchor_method=None, anchor_change=None):
    """S1_2_find_anchor_main.py:8-45. Profiles on the GPU, peak logic on the host."""
    d = filtered_mtx
    gi = d.row_window
    layout = gi.layout(d.dmatrix.device)
    names, seps = gi.chros, gi.seps()
    starts, ends = _bins(gi)
    if anchor_method in ('inter_rowsum', 'no_anchor'):
        _, smooth = ops.inter_rowsum_profile(d.dmatrix, layout)
        table = A.find_anchors_by_rowsum(smooth.cpu().numpy(), names, seps, starts, ends)
    elif anchor_method == 'intra_low':
        _, smooth = ops.intra_low_profile(d.dmatrix, layout)
        table = A.find_anchors_by_rowsum(smooth.cpu().numpy(), names, seps, starts, ends)
    elif anchor_method == 'intra_x':
        prof = np.zeros(layout.n)
        for (s, e) in seps:
            blk = d.dmatrix[s:e + 1, s:e + 1]
            if blk.data_ptr() % 4:
                blk = blk.contiguous()
            prof[s:e + 1] = ops.intra_x_response(blk).cpu().numpy()
        table = A.find_anchors_by_intra_x(prof, names, seps, starts, ends)
    else:
        raise ValueError('Unrecognized anchor method.')
    if anchor_change is not None:
        changes = A.parse_anchor_change(anchor_change) if isinstance(anchor_change, str) else anchor_change
        if changes:
            table = A.change_anchor(table, names, changes)
    if anchor_method == 'no_anchor':
        table = A.remove_all_used(table)
    if output is not None:
        table.write(f'{output}{anchor_suffix}')
    return table


def anchors_io(index_file, output, sps_file=None, fig_output=None, use_exist_filt=True,
               anchor_method='inter_rowsum', anchor_change=None):
    """publish/HiArch/GF_S1_get_center.py:12-27."""
    filtered = filter_io(sps_file, index_file, output, fig_output, use_exist_filt)
    return main_find_anchor(filtered, output, fig_output, anchor_method=anchor_method,
                            anchor_change=anchor_change)


class ConforData:
    """publish/confor/confor_class.py:37-94 (names, xs, isym_xs)."""

    def __init__(self, names, xs, isym_xs):
        self.names = names
        self.xs = xs
        self.isym_xs = isym_xs

    def out_names(self):
        return list(self.names)


def read_anchor_table(center_file, gen_index):
    df = pd.read_table(center_file, header=0, dtype={'chr': str})
    rows = df.to_dict('records')
    return A.AnchorTable(rows, list(df.columns))


def to_train_mtx(sps_file, window_file, center_file, spe_name, do_plot=False, fig_out_dir=None,
                 n_plot_intra=2, n_plot_inter=3, is_circle=False):
    """to_train_mtx.py:75-129: every ordered chromosome pair -> 16x16 image (zero-filled blocks:
    the reference densifies with scipy .todense() here)."""
    m = read_matrix(sps_file, window_file)
    m.has_neg = False                                   # .todense(): absent cells are 0
    dense = m.to_dense()
    gi = dense.row_window
    layout = gi.layout(dense.dmatrix.device)
    table = read_anchor_table(center_file, gi)
    centers = [table.used_center(c, gi.chr_seps[c][0]) for c in gi.chros]
    res = ops.gf_scores(dense.dmatrix, layout, centers)
    names_i, names_o, idx_i, idx_o = [], [], [], []
    for p, (a, b) in enumerate(res["pairs"]):
        if a == b:
            names_i.append(f'{spe_name}:{gi.chros[a]}')
            idx_i.append(p)
        else:
            names_o.append(f'{spe_name}:{gi.chros[a]}:{gi.chros[b]}')
            idx_o.append(p)
    intra = ConforData(names_i, res["xs"][idx_i], res["isym"][idx_i])
    inter = ConforData(names_o, res["xs"][idx_o], res["isym"][idx_o])
    intra.scores, inter.scores = res["scores"][idx_i], res["scores"][idx_o]
    return intra, inter


def get_n_norm_ave_maps(ave_maps):
    """to_confor.py:12-25: every template / sum(|template|), flattened (keys starting with '_' are
    loadmat's metadata)."""
    return {k: (np.asarray(v, dtype=np.float64) / np.sum(np.abs(v))).reshape((-1,))
            for k, v in ave_maps.items() if not k.startswith('_')}


def change_map_name(ave_maps, ave_names):
    """to_confor.py:28-32."""
    u_names = sorted(int(float(i)) for i in ave_maps if not i.startswith('_'))
    return {ave_names[i]: ave_maps[str(u_names[i])] for i in range(len(u_names))}


def _template_scores_device(xs_arr, ave_maps, do_trans):
    """scores = X (P x 256) . T^T (256 x k) for caller-supplied templates, max with the transposed
    images for inter pairs (to_confor.py:39-46): `hia_template_scores`, fp64 like the reference (the
    built-in templates never come here, their scores are produced by `hia_gf_pool_scores` together
    with the images)."""
    import torch
    from . import _lib
    lib = _lib.load()
    x = torch.as_tensor(np.ascontiguousarray(np.asarray(xs_arr, dtype=np.float64)), device="cuda")
    t = torch.as_tensor(np.ascontiguousarray(np.vstack([np.asarray(v, dtype=np.float64).reshape(-1)
                                                        for v in ave_maps.values()])), device="cuda")
    assert x.shape[1] == X_SIZE * X_SIZE == t.shape[1] == 256
    scores = torch.empty((x.shape[0], t.shape[0]), dtype=torch.float64, device="cuda")
    st = lib.hia_template_scores(x.data_ptr(), x.shape[0], t.data_ptr(), t.shape[0], 1 if do_trans else 0,
                                 scores.data_ptr(), ops._stream())
    _lib.check(st, "hia_template_scores")
    return scores.cpu().numpy()


def to_confor(xs, ave_maps, do_trans=False):
    """to_confor.py:35-64: long-format table (name, species, chr1, chr2, type, score).
    `ave_maps`: the reference's dict name -> normalised 256-vector (scores are then taken against
    THESE templates), or just the template names in order -- then the products the GPU already
    took against the shipped templates (`xs.scores`, inter = max with transposed) are used."""
    if xs.xs is None or len(xs.names) == 0:
        return pd.DataFrame([])
    if isinstance(ave_maps, dict):
        values = _template_scores_device(xs.xs, ave_maps, do_trans)
        columns = list(ave_maps.keys())
    else:
        values, columns = xs.scores, list(ave_maps)
    scores = pd.DataFrame(values, xs.out_names(), columns=columns)
    scores.reset_index(inplace=True, names=['name'])
    scores['name'] = scores['name'].str.rstrip()
    parts = scores['name'].str.split(':', expand=True)
    scores['species'] = parts.iloc[:, 0]
    scores['chr1'] = parts.iloc[:, 1]
    scores['chr2'] = parts.iloc[:, 1] if parts.shape[1] == 2 else parts.iloc[:, 2]
    return pd.melt(scores, id_vars=['name', 'species', 'chr1', 'chr2'], var_name='type', value_name='score')


# the shipped templates (publish/confor/{intra,inter}_ave_maps.mat, carried as data/ave_maps.npz)
intra_ave_map_file = None
inter_ave_map_file = None


def main_to_confor(intra_xs, inter_xs, out_dir, intra_ave_file=intra_ave_map_file,
                   inter_ave_file=inter_ave_map_file):
    """to_confor.py:77-93 + out_confor (:67-74). `*_ave_file`: a .mat of templates (keys '0'..'4')
    to score against instead of the shipped ones (None)."""
    def maps(path, names):
        if path is None:
            return names
        from scipy.io import loadmat
        return change_map_name(get_n_norm_ave_maps(loadmat(path)), names)

    allc = pd.concat([to_confor(intra_xs, maps(intra_ave_file, intra_ave_names)),
                      to_confor(inter_xs, maps(inter_ave_file, inter_ave_names), do_trans=True)], axis=0)
    allc = allc.sort_values(['species', 'chr1', 'chr2'])
    for species in allc['species'].unique():
        if os.path.exists(f'{out_dir}/{species}'):
            allc[allc['species'] == species].to_csv(f'{out_dir}/{species}/{species}{confor_suffix}',
                                                    sep="\t", index=False)
    return allc


def get_confor(sps_file, index_file, anchor_file, out_dir, fig_outdir=None):
    """publish/HiArch/GF_S2_get_score.py:11-25 (result lands in <out_dir>/<species>.confor.txt)."""
    spe_name = sps_file.split('/')[-1].split('.')[0]
    intra, inter = to_train_mtx(sps_file, index_file, anchor_file, spe_name)
    os.makedirs(f'{out_dir}/{spe_name}', exist_ok=True)
    table = main_to_confor(intra, inter, out_dir)
    src = f'{out_dir}/{spe_name}/{spe_name}{confor_suffix}'
    if os.path.exists(src):
        os.replace(src, f'{out_dir}/{spe_name}{confor_suffix}')
    try:
        os.rmdir(f'{out_dir}/{spe_name}')
    except OSError:
        pass
    return table
