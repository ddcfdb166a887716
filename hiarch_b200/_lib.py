"""ctypes binding of libhiarch_b200.so (the C ABI declared in include/hiarch_b200.h).

There is no CPU fallback: if the library is missing or an entry point is absent the import of
an op fails loudly. torch tensors are only the buffer type -- every call passes raw device
pointers, sizes and the current CUDA stream.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhiarch_b200.so")

c_i32, c_i64, c_f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
c_int, c_vp = ctypes.c_int, ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol include/hiarch_b200.h declares
PROTOTYPES = {
    "hia_version": (c_int, []),
    "hia_status_string": (ctypes.c_char_p, [c_int]),
    "hia_last_cuda_error": (ctypes.c_char_p, []),
    "hia_launch_count": (ctypes.c_uint64, []),
    "hia_scatter_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "hia_scatter_coo_sym": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_int, c_int,
                                    c_vp, c_i32, c_vp, c_vp]),
    "hia_scatter_csr_sym": (c_int, [c_vp, c_vp, c_int, c_vp, c_i32, c_vp, c_i64, c_int, c_vp]),
    "hia_scatter_coo_sym_stats": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i64, c_int, c_vp, c_i32, c_vp, c_vp,
                                          c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hia_scatter_csr_sym_stats": (c_int, [c_vp, c_vp, c_int, c_vp, c_i32, c_vp, c_i64, c_int, c_vp, c_i32, c_vp, c_vp,
                                          c_vp, c_vp, c_vp, c_vp]),
    "hia_scatter_coo_remap": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_i64, c_int, c_vp]),
    "hia_coo_block_nnz": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_int, c_vp, c_i32, c_vp, c_vp]),
    "hia_coo_bin_sums": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_i32, c_int, c_vp, c_vp]),
    "hia_bin_stats": (c_int, [c_vp, c_i32, c_i64, c_vp, c_vp, c_vp]),
    "hia_oe_expected_pass": (c_int, [c_vp, c_i32, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_vp, c_vp]),
    "hia_oe_expected_pass_ex": (c_int, [c_vp, c_i32, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, c_vp]),
    "hia_oe_finalize": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_int,
                                c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hia_oe_apply": (c_int, [c_vp, c_vp, c_i32, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_int,
                             c_vp]),
    "hia_oe_expected_pass_rows": (c_int, [c_vp, c_i32, c_i64, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp,
                                          c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hia_oe_apply_rows": (c_int, [c_vp, c_vp, c_i32, c_i64, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp,
                                  c_vp, c_int, c_vp]),
    "hia_oe_apply_f64": (c_int, [c_vp, c_i32, c_i64, c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "hia_log_map": (c_int, [c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "hia_similarity_workspace_bytes": (c_i64, [c_i32, c_i32, c_int]),
    "hia_similarity": (c_int, [c_vp, c_i32, c_i32, c_i64, c_vp, c_i64, c_int, c_int, c_i32, c_i32,
                               c_i32, c_vp, c_i64, c_vp]),
    "hia_similarity_prep": (c_int, [c_vp, c_i32, c_i32, c_i64, c_int, c_int, c_vp, c_i64, c_vp]),
    "hia_similarity_gemm": (c_int, [c_i32, c_i32, c_vp, c_i64, c_int, c_int, c_i32, c_i32, c_i32, c_vp, c_i64,
                                    c_vp]),
    "hia_percentiles_workspace_bytes": (c_i64, []),
    "hia_percentiles": (c_int, [c_vp, c_i32, c_i32, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "hia_clip": (c_int, [c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp, c_vp]),
    "hia_core_moments": (c_int, [c_vp, c_i32, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hia_zscore": (c_int, [c_vp, c_vp, c_i32, c_i32, c_i64, c_vp, c_vp]),
    "hia_sym_clip_bounds": (c_int, [c_vp, c_vp, c_vp]),
    "hia_band_entropy_workspace_bytes": (c_i64, [c_i32]),
    "hia_band_entropy": (c_int, [c_vp, c_i32, c_i64, c_i32, c_i32, c_vp, c_i32, ctypes.c_double, c_vp,
                                 c_vp, c_i64, c_vp]),
    "hia_row_profiles": (c_int, [c_vp, c_i32, c_i64, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp,
                                 c_vp]),
    "hia_smooth1d_reflect": (c_int, [c_vp, c_vp, c_i32, c_vp, c_vp, ctypes.c_double, c_i32, c_vp]),
    "hia_gf_pool_scores": (c_int, [c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hia_template_scores": (c_int, [c_vp, c_i32, c_vp, c_i32, c_int, c_vp, c_vp]),
    "hia_wavelet_default_levels": (c_i32, [c_i32, c_i32]),
    "hia_wavelet_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32]),
    "hia_wavelet_denoise_db3": (c_int, [c_vp, c_i64, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_i64,
                                        c_vp]),
    "hia_intra_x_response": (c_int, [c_vp, c_i32, c_i64, c_i32, c_vp, c_vp]),
    "hia_intra_x_workspace_bytes": (c_i64, [c_i32]),
    "hia_intra_x_response_prefix": (c_int, [c_vp, c_i32, c_i64, c_i32, c_vp, c_vp, c_i64, c_vp]),
    "hia_box_filter_reflect": (c_int, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp]),
    "hia_median_filter_reflect": (c_int, [c_vp, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp]),
    "hia_used_filter_f64_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "hia_used_filter_f64": (c_int, [c_vp, c_i32, c_i32, c_i64, c_i32, c_i32, c_int, ctypes.c_double, c_vp,
                                    c_i64, c_vp, c_i64, c_vp]),
    "hia_topk_eig_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32]),
    "hia_projected_eig": (c_int, [c_vp, c_i32, c_i32, c_i32, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hia_eig_symm_upper_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "hia_topk_eig": (c_int, [c_vp, c_i32, c_i64, c_i32, c_i32, c_i32, c_i32, ctypes.c_uint64, c_vp, c_i32,
                             ctypes.c_double, ctypes.c_double, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp,
                             c_i64, c_vp]),
    "hia_sim_kpad": (c_i32, [c_i32, c_int]),
    "hia_sim_plane_elem_bytes": (c_i32, [c_int]),
    "hia_sim_prep_rows": (c_int, [c_vp, c_i32, c_i32, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    "hia_sim_gemm_rows_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "hia_sim_gemm_rows": (c_int, [c_vp, c_vp, c_int, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i64, c_int,
                                  c_i32, c_vp, c_i64, c_vp]),
    "hia_sim_gemm_block": (c_int, [c_vp, c_vp, c_i32, c_vp, c_vp, c_i32, c_int, c_i32, c_i32, c_i32, c_i32, c_i32,
                                   c_int, c_vp, c_i64, c_vp, c_i64, c_int, c_i32, c_vp, c_i64, c_vp]),
    "hia_ipc_export": (c_int, [c_vp, c_vp, c_vp]),
    "hia_ipc_open": (c_int, [c_vp, c_vp]),
    "hia_ipc_close": (c_int, [c_vp]),
    "hia_peer_copy": (c_int, [c_vp, c_vp, c_i64, c_vp]),
    "hia_mtx_scan": (c_i64, [ctypes.c_char_p, c_i32, c_vp, c_vp, c_vp]),
    "hia_mtx_count_lines": (c_i64, [ctypes.c_char_p, c_i32]),
    "hia_mtx_parse": (c_int, [ctypes.c_char_p, c_i32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hia_mtx_write": (c_int, [ctypes.c_char_p, c_i64, c_vp, c_vp, c_vp, c_int, c_i32]),
    "hia_eig_begin": (c_int, [c_i32, c_i32, c_i32, ctypes.c_uint64, c_vp, c_i32, c_vp, c_i64, c_vp]),
    "hia_eig_symm_rows": (c_int, [c_vp, c_i32, c_i32, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "hia_eig_absorb": (c_int, [c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "hia_eig_first_check": (c_i32, [c_i32, c_i32]),
    "hia_eig_check_every": (c_i32, [c_i32]),
    "hia_eig_check": (c_int, [c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.c_double, ctypes.c_double, c_vp, c_vp,
                              c_vp, c_vp]),
    "hia_eig_ritz": (c_int, [c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hia_eig_set_ritz": (c_int, [c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "hia_eig_resid": (c_int, [c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
}

_lib = None


class HiaError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every prototype. Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HiaError(
            f"{LIB_PATH} not found: build it with `python -m hiarch_b200.build` "
            "(hiarch_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        lib = load()
        msg = lib.hia_status_string(status).decode()
        if status == -3:
            msg += ": " + lib.hia_last_cuda_error().decode()
        raise HiaError(f"{what} failed: {msg} ({status})")
