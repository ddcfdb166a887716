"""Mirror of publish/complex (checkerboard score) on the CUDA ops.

get_adj_mtx   publish/complex/src/S1_get_adj_mtx.py:8-14
get_local / main_local / local_io   publish/complex/src/main_local.py:19-148
"""
import pandas as pd

from . import ops
from .new_hic_class import ArrayHiCMtx, SpsHiCMtx, read_sps_file

min_chro_len = ops.MIN_CHRO_LEN
min_n_short_sims = ops.MIN_N_SHORT_SIMS


def get_adj_mtx(input_mtx, metric='cosine'):
    """squareform(pdist(M, metric)) on tcgen05. metric: 'cosine' (the pipeline's, S1_get_adj_mtx.py:8) or
    'correlation' (1 - Pearson, the same contraction on row-centred rows). scipy's other pdist
    metrics are not inner-product shaped and are not on the hot path: NotImplementedError."""
    flavours = {'cosine': ops.SIM_COSINE_DISTANCE, 'correlation': ops.SIM_CORRELATION_DISTANCE}
    if metric not in flavours:
        raise NotImplementedError(f"metric={metric!r}: only 'cosine' and 'correlation' run on the tensor cores")
    blk = input_mtx.dmatrix
    if blk.data_ptr() % 16 or blk.stride(0) % 4 or blk.stride(1) != 1:
        tmp = ops.alloc_map(blk.shape[0], blk.device, cols=blk.shape[1])
        tmp.copy_(blk)
        blk = tmp
    adj = ops.similarity(blk, flavour=flavours[metric])
    return ArrayHiCMtx(adj, copy_mtx_paras=input_mtx, has_neg=False, col_window=input_mtx.row_window,
                       mtx_type='all')


def get_local(input_mtx, do_plot=False, fig_output=None, chro_str=None, input_axes=None,
              short_dis_max_per=.15):
    """main_local.py:19-105 (plots are out of scope). L = mean chromosome length of the block's
    row window (S2_get_entro.py:6-10)."""
    lens = list(input_mtx.row_window.chr_lens.values())
    L = sum(lens) / len(lens)
    return ops.get_local(input_mtx.dmatrix, L, short_dis_max_per)


def main_local(input_data, out_file=None, do_plot=False, fig_output=None, short_dis_max_per=.15):
    """main_local.py:108-141. A sparse input is densified once with the per-block min fill (what
    densifying the blocks of yield_by_chro(triu=True) one by one gives); blocks are views."""
    if isinstance(input_data, SpsHiCMtx):
        dense = input_data.to_dense(block_min=True)
    else:
        dense = input_data
    layout = dense.row_window.layout(dense.dmatrix.device)
    rows = ops.main_local(dense.dmatrix, layout, dense.row_window.chros, short_dis_max_per)
    complexes = pd.DataFrame({'accord_chro': [r[0] for r in rows], 'other_chro': [r[1] for r in rows],
                              'accord': [r[2] for r in rows]})
    if out_file is not None:
        complexes.to_csv(out_file, sep="\t", index=False)
    return complexes


def local_io(sps_file, index_file, out_file, do_plot=False, fig_output=None, short_dis_max_per=.15):
    """main_local.py:144-148 (Checkerboard.py step)."""
    sps_mtx = read_sps_file(sps_file, index_file)
    return main_local(sps_mtx, out_file=out_file, fig_output=fig_output, do_plot=do_plot,
                      short_dis_max_per=short_dis_max_per)
