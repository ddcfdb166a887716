"""Mirror of publish/oe_map (the O/E normalisation package) on the CUDA ops.

oe_map_core    publish/oe_map/S2_oe.py:10-15            obs_d_exp(sgd_mean, k=.2) -> log, fused
tid_off_outliers / main_denoise   publish/oe_map/S3_denoise.py:85-107
norm_to_zero_mean / remove_outliers / norm_de_mtx       publish/oe_map/S4_norm_value.py:10-46
oe_map_main / oe_map_io / norm_dis                      publish/oe_map/main_get_oe.py:10-76, publish/HiArch/NormDis.py:10-17
"""
import numpy as np

from . import ops
from .new_hic_class import ArrayHiCMtx, SpsHiCMtx

mode = 'sgd_mean'
k = .2
counts_must_decrese = True
mid_vmax_per = 66
default_vmax_per = 98


def _dense_all(m):
    d = m.to_dense() if isinstance(m, SpsHiCMtx) else m
    if d.mtx_type != 'all':
        d = d.to_all(inplace=False)
    return d


def oe_map_core(sps_mtx, counts_must_decrese=True):
    """S2_oe.py:10-15: one fused call (expected pass + finalize + apply+log)."""
    d = _dense_all(sps_mtx)
    layout = d.row_window.layout(d.dmatrix.device)
    out, state = ops.obs_d_exp(d.dmatrix, layout, d.row_window.bin_size, k=k,
                               counts_must_decrese=counts_must_decrese, log=True)
    res = ArrayHiCMtx(out, copy_mtx_paras=d, mtx_type='all', has_neg=False)
    res._oe_state = state
    return res


def tid_off_outliers(sps_map, vamx_per=default_vmax_per):
    d = _dense_all(sps_map)
    return ArrayHiCMtx(ops.tid_off_outliers(d.dmatrix, vamx_per), copy_mtx_paras=d)


def norm_to_zero_mean(input_mtx):
    d = _dense_all(input_mtx)
    return ArrayHiCMtx(ops.norm_to_zero_mean(d.dmatrix, mid_vmax_per), copy_mtx_paras=d, has_neg=True)


def remove_outliers(input_mtx):
    d = _dense_all(input_mtx)
    return ArrayHiCMtx(ops.remove_outliers(d.dmatrix, default_vmax_per), copy_mtx_paras=d)


def norm_de_mtx(input_mtx):
    return remove_outliers(norm_to_zero_mean(input_mtx))


def denoise_one_method(sps_mtx, mode='wavelet', sigma=None, spatial_sigma_ratio=None):
    """S3_denoise.py:38-73, modes 'wavelet', 'median' and 'mean'. 'wavelet' is
    skimage.restoration.denoise_wavelet(arr, wavelet='db3', method="VisuShrink") (:43-47) on the
    device (csrc/hia_wavelet.cu); that arithmetic is third-party and unversioned in the reference, so
    its parity is pinned to oracle/wavelet_oracle.py's written-down definition only. `sigma` is
    accepted and, as in the reference, not passed on. 'low-pass' / 'bilateral' are unused by the
    pipeline and not built."""
    d = sps_mtx.to_dense() if isinstance(sps_mtx, SpsHiCMtx) else sps_mtx
    if mode == 'wavelet':
        return ArrayHiCMtx(ops.wavelet_denoise_db3(d.dmatrix), copy_mtx_paras=d)
    ratio = .1 if spatial_sigma_ratio is None else spatial_sigma_ratio
    size = int(np.mean(np.array(list(sps_mtx.row_window.chr_lens.values())) * ratio))
    if mode == 'median':
        size = 2 if size < 2 else size
        src = d.dmatrix
        if src.data_ptr() % 16 or src.stride(0) % 4:
            tmp = ops.alloc_map(src.shape[0], src.device, cols=src.shape[1])
            tmp.copy_(src)
            src = tmp
        out = ops.median_filter_reflect(src, size)
    elif mode == 'mean':
        out = ops.alloc_map(d.shape[0], d.dmatrix.device, cols=d.shape[1])
        ops.used_filter(d.dmatrix, size, out, decimals=-1, normalise=False)
    else:
        raise NotImplementedError(f"denoise mode {mode!r}: not used by the pipeline, not built")
    return ArrayHiCMtx(out, copy_mtx_paras=d)


def used_denoise(sps_mtx, wavelet=None):
    """S3_denoise.py:76-79: wavelet, then median (size from 1 % of the mean chromosome length)."""
    wavelet = _resolve_wavelet(wavelet)
    if wavelet is not None:
        sps_mtx = wavelet(sps_mtx)
    return denoise_one_method(sps_mtx, 'median', None, .01)


def _native_wavelet(blk):
    return denoise_one_method(blk, 'wavelet', None, None)


def _resolve_wavelet(wavelet):
    """None / 'db3' -> the reference's step on the device; 'identity' -> skip it knowingly (the
    fixtures generated before the step existed say wavelet=identity); callable -> applied per block."""
    if wavelet is None or (isinstance(wavelet, str) and wavelet == 'db3'):
        return _native_wavelet
    if isinstance(wavelet, str):
        if wavelet != 'identity':
            raise ValueError(f"wavelet={wavelet!r}: pass None / 'db3', 'identity' or a callable")
        return None
    return wavelet


def main_denoise(sps_map, wavelet=None):
    """S3_denoise.py:95-107: clip to [p2, p98], then per chromosome block wavelet -> median.
    `wavelet`: None / 'db3' = the reference's db3 VisuShrink step (device kernels), 'identity' =
    leave it out, callable = applied per block instead. The result carries `.wavelet_applied`."""
    wavelet = _resolve_wavelet(wavelet)
    d = _dense_all(sps_map)
    clipped = tid_off_outliers(d)
    out = ops.alloc_map(d.shape[0], d.dmatrix.device)
    res = ArrayHiCMtx(out, copy_mtx_paras=clipped)
    for c1, c2, blk in clipped.yield_by_chro():
        if wavelet is not None:
            blk = wavelet(blk)
        den = denoise_one_method(blk, 'median', None, .01)
        s1, e1 = d.row_window.chr_seps[c1]
        s2, e2 = d.col_window.chr_seps[c2]
        out[s1:e1 + 1, s2:e2 + 1].copy_(den.dmatrix)
    res.wavelet_applied = wavelet is not None
    return res


def oe_map_main(sps_mtx, output=None, fig_output=None, centr_loc=None,
                do_filt_chr_by_name=False, counts_must_decrese=True, wavelet=None):
    """main_get_oe.py:10-59, the NormDis step on an already parsed map: preprocess_mtx ->
    oe_map_core -> `<output>.ode.mtx` (+ `.window.bed`) -> main_denoise -> norm_de_mtx ->
    `<output>.de_ode.mtx`. The plots (`fig_output`, `centr_loc`) are out of scope and ignored;
    `wavelet` goes to main_denoise (None = the db3 VisuShrink step, see there)."""
    from .preprocess import preprocess_mtx
    pro_mtx = preprocess_mtx(sps_mtx, do_filt_chr_by_name)
    if pro_mtx is None:
        print(f'Warning: Not found proper matrix for {sps_mtx.mtx_file}')
        return None
    print(f'Processed matrix size is {tuple(pro_mtx.shape)}')
    chr_lens_list = list(pro_mtx.row_window.chr_lens.values())
    print(f'Chr num is {len(chr_lens_list)}. Max chr size is {np.max(chr_lens_list)}, min is {np.min(chr_lens_list)}.')
    ode_mtx = oe_map_core(pro_mtx, counts_must_decrese=counts_must_decrese)
    print('Matrix normalization DONE.')
    if output is not None:
        ode_mtx.to_sps().to_output(f'{output}.ode.mtx')
    de_ode_mtx = norm_de_mtx(main_denoise(ode_mtx, wavelet=wavelet))
    print('Denoise normalized matrix DONE.')
    if output is not None:
        de_ode_mtx.to_sps().to_output(f'{output}.de_ode.mtx', out_window=False)
    return de_ode_mtx


def oe_map_io(mtx_file, output, fig_output=None, do_filt_chr_by_name=False, counts_must_decrese=True,
              index_file=None, input_centr_loc=None, thread=1, wavelet=None):
    """main_get_oe.py:64-76."""
    from .new_hic_class import read_matrix
    sps_mtx = read_matrix(mtx_file, index_file, thread=thread)
    return oe_map_main(sps_mtx, output, fig_output, None, do_filt_chr_by_name, counts_must_decrese, wavelet=wavelet)


def norm_dis(sps_file, index_file, output, fig_output=None, do_filt_chr_by_name=True,
             counts_must_decrese=True, wavelet=None):
    """publish/HiArch/NormDis.py:10-17 (step 1 of one_click_pipeline.sh)."""
    return oe_map_io(sps_file, output, fig_output, index_file=index_file,
                     do_filt_chr_by_name=do_filt_chr_by_name, counts_must_decrese=counts_must_decrese,
                     wavelet=wavelet)
