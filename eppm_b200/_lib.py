"""ctypes binding of eppm_b200/libeppm_b200.so (include/eppm.h).  The library is the product; this module only
declares its prototypes.  Loading fails loudly when the shared object has not been built: there is no fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EPPM_LIB_PATH") or os.path.join(_HERE, "libeppm_b200.so")  # override = A/B builds while tuning


class EppmParams(C.Structure):
    """struct eppm_params (include/eppm.h); defaults = defs.h:31-76 of the reference."""
    _fields_ = [
        ("pyr_levels", C.c_int), ("num_iter", C.c_int), ("patch_r", C.c_int), ("patch_stride", C.c_int),
        ("search_range", C.c_int), ("search_radius_min", C.c_int), ("num_rand_guess", C.c_int), ("prop_seg_length", C.c_int),
        ("lambda_ad", C.c_float), ("lambda_census", C.c_float), ("pm_sig_r", C.c_float),
        ("stat_radius", C.c_int), ("stat_sim_thresh", C.c_int), ("wmf_radius", C.c_int), ("wmf_sig_r", C.c_float),
        ("wmf_iters", C.c_int), ("blf_sig_s", C.c_int), ("blf_sig_r", C.c_float), ("rng_mode", C.c_int),
        ("seed", C.c_ulonglong), ("inplace_filters", C.c_int), ("subpixel_final", C.c_int), ("reserved", C.c_int * 6),
    ]


class EppmFlowError(C.Structure):
    """struct eppm_flow_error (include/eppm.h)."""
    _fields_ = [("epe", C.c_double), ("aae_deg", C.c_double), ("outlier_frac", C.c_double), ("n_valid", C.c_longlong), ("n_known", C.c_longlong)]


# every symbol include/eppm.h declares: name -> (restype, argtypes)
EPPM_SYMBOLS = {
    "eppm_default_params": (None, [C.POINTER(EppmParams)]),
    "eppm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(EppmParams)]),
    "eppm_destroy": (None, [C.c_void_p]),
    "eppm_last_error": (C.c_char_p, []),
    "eppm_version": (C.c_char_p, []),
    "eppm_num_levels": (C.c_int, [C.c_void_p]),
    "eppm_level_dims": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "eppm_compute_batch_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "eppm_compute_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "eppm_compute_stream_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "eppm_compute_stream_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "eppm_synchronize": (C.c_int, [C.c_void_p]),
    "eppm_stream": (C.c_void_p, [C.c_void_p]),
    "eppm_stage_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "eppm_stage_patchmatch": (C.c_int, [C.c_void_p]),
    "eppm_stage_consistency": (C.c_int, [C.c_void_p]),
    "eppm_stage_c2f": (C.c_int, [C.c_void_p, C.c_void_p]),
    "eppm_read_plane": (C.c_long, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eppm_write_plane": (C.c_long, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "eppm_stage_patchmatch_partial": (C.c_int, [C.c_void_p, C.c_int]),
    "eppm_set_band": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "eppm_band_rows": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "eppm_tiled_pm_steps": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "eppm_tiled_c2f_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "eppm_device_plane": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int]),
    "eppm_tiled_unique_id": (C.c_int, [C.c_void_p]),
    "eppm_tiled_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "eppm_compute_tiled_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eppm_compute_tiled_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "eppm_tiled_shutdown": (C.c_int, [C.c_void_p]),
    "eppm_selftest_const_div": (C.c_longlong, [C.c_float, C.c_uint, C.c_uint]),
    "eppm_smooth_uses_fast_div": (C.c_int, [C.c_void_p]),
    "eppm_smooth_uses_tma": (C.c_int, [C.c_void_p]),
    "eppm_selftest_affine_sites": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "eppm_selftest_affine_sites_stride": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "eppm_selftest_volume_tables": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint)]),
    "eppm_refine_uses_site_table": (C.c_int, [C.c_void_p, C.c_int]),
    "eppm_eval_flow": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.POINTER(EppmFlowError)]),
    "eppm_write_flo": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "eppm_read_flo": (C.c_int, [C.c_char_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_size_t]),
    "eppm_launch_count": (C.c_ulonglong, [C.c_int]),
    "eppm_last_stage_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float * 5)]),
    "eppm_last_kernel_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
}

# the reference's stage functions re-exported with the reference's signatures (include/eppm_legacy_abi.h)
_P, _S, _I = C.c_void_p, C.c_size_t, C.c_int
LEGACY_SYMBOLS = {
    "baoCudaPatchMatchMultiscalePrepare": (None, [_P] * 10 + [_I, _P, _P, _I, _I]),
    "baoCudaCensusTransform": (None, [_P, _P, _P, _P, _I, _I, _S, _S]),
    "baoCudaPatchMatch": (None, [_P] * 6 + [_I, _I, _S, _S, _S, _S]),
    "baoCudaLeftRightCheck": (None, [_P, _P, _P, _P, _I, _I, _S, _S]),
    "baoCudaOutlierRemoval": (None, [_P, _P, _I, _I, _S, _S]),
    "baoCudaWeightedMedianFilter": (None, [_P, _P, _P, _I, _I, _S, _S, _S, _I, C.c_bool]),
    "baoCudaFillHole": (None, [_P, _P, _P, _I, _I, _S, _S, _S]),
    "baoCudaNNF2Flow": (None, [_P, _P, _I, _I, _S, _S]),
    "baoCudaBLF_C2F": (None, [_P] * 11 + [_I]),
    "baoCudaBLFCostFilterRefine": (None, [_P] * 5 + [_I, _I, _S, _S]),
    "baoCudaFlowSmoothing": (None, [_P, _P, _I, _I, _S, _S]),
    "baoCudaLeftRightCheck_Buffered": (None, [_P] * 6 + [_I, _I, _S, _S]),
    "baoCudaFlow2NNF": (None, [_P, _P, _I, _I, _S, _S]),
    "baoCudaFlowCutoff": (None, [_P, _I, _I, _S, C.c_float]),
    "baoEliminateStillRegionFlow": (None, [_P, _P, _P, _I, _I, _S]),
    "baoCudaImageSmoothing": (None, [_P, _P, _I, _I, _S]),
    "baoCudaFlowBilteralUpsampling": (None, [_P, _P, _I, _I, _S, _P, _I, _I, C.c_float]),
    "baoCudaPatchMatch_Scaled": (None, [_P] * 7 + [_I, _I, _S, _S, _S, _S, _S]),
    "baoCudaPatchMatch_PlaneFitting": (None, [_P] * 6 + [_I, _I, _S, _S, _S, _S]),
    "baoCudaCensusTransform_Bicubic": (None, [_P, _P, _I, _I, _S, _P, _P, _I, _I, _S]),
    "baoCudaSubpixRefine": (None, [_P] * 6 + [_I, _I, _S, _S, _S, _S]),
}

_lib = None


def bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


def load():
    """dlopen libeppm_b200.so and attach prototypes.  Raises OSError if it was not built (run __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  eppm_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        bind(lib, EPPM_SYMBOLS)
        bind(lib, LEGACY_SYMBOLS)
        _lib = lib
    return _lib
