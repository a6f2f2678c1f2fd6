"""ctypes binding of libmsig.so (the C ABI declared in include/msig.h).

PyTorch is used only for device memory (caching allocator) and streams; every kernel on the hot
path is in libmsig.so. There is no fallback: if the library is missing or the device is not an
sm_100a GPU, calls raise RuntimeError.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_longlong,
                    c_size_t, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSIG_LIB") or os.path.join(_HERE, "libmsig.so")   # MSIG_LIB: probe builds only

# enums (mirror include/msig.h)
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
AUX_NONE, AUX_ADD, AUX_RELU_MASK, AUX_LRELU_MASK = 0, 1, 2, 3
OUT_BF16_NHWC, OUT_F32_NCHW, OUT_F32_NHWC = 0, 1, 2
(WPACK_FWD, WPACK_DGRAD_S1, WPACK_DGRAD_S2, WPACK_CONVT_FWD, WPACK_CONVT_DGRAD, WPACK_IM2COL,
 WPACK_IM2COL_DGRAD, WPACK_IM2COL_FLIP, WPACK_ROWFOLD, WPACK_ROWFOLD_DGRAD, WPACK_ROWPATCH,
 WPACK_ROWPATCH_FLIP) = range(12)


class ConvGeom(Structure):
    _fields_ = [(k, c_int32) for k in
                ("n", "h", "w", "c", "k", "r", "s", "stride", "pad_t", "pad_l", "oh", "ow")]


class Epilogue(Structure):
    _fields_ = [("bias", c_void_p), ("aux", c_void_p), ("aux_mode", c_int32), ("act", c_int32),
                ("alpha", c_float), ("alpha_ptr", c_void_p), ("slope", c_float),
                ("out_layout", c_int32), ("stats_partial", c_void_p), ("stats_z", c_void_p),
                ("ch_scale", c_void_p), ("mask_scale", c_void_p), ("mask_shift", c_void_p), ("stats_rows", c_int32)]


class WpackDesc(Structure):
    _fields_ = [(k, c_int32) for k in ("kind", "o", "i", "r", "s")]


class WpackJob(Structure):
    _fields_ = [("d", WpackDesc), ("oc", c_int32), ("o_off", c_int32), ("src", c_void_p), ("dst", c_void_p),
                ("copy_numel", c_int64)]


class PatchGeom(Structure):
    _fields_ = [(k, c_int32) for k in
                ("n", "c", "h", "w", "r", "s", "stride", "pad_t", "pad_l", "oh", "ow", "reflect",
                 "kpad")]


_P = c_void_p
_SIGS = {
    "msig_init": (c_int, [c_int]),
    "msig_version": (c_int, []),
    "msig_last_error": (c_char_p, []),
    "msig_sm_count": (c_int, []),
    "msig_debug_set_m2_mode": (c_int, [c_int]),
    "msig_debug_set_pdl": (c_int, [c_int]),
    "msig_debug_set_wgrad_mode": (c_int, [c_int]),
    "msig_debug_set_pair_mode": (c_int, [c_int]),
    "msig_debug_set_ring_mode": (c_int, [c_int]),
    "msig_kernel_launches": (c_longlong, []),
    "msig_wpack_elems": (c_size_t, [POINTER(WpackDesc)]),
    "msig_wpack": (c_int, [POINTER(WpackDesc), _P, _P, _P]),
    "msig_wpack_table_bytes": (c_size_t, [c_int32]),
    "msig_wpack_table_build": (c_int, [POINTER(WpackJob), c_int32, _P, POINTER(c_int64), POINTER(c_int64)]),
    "msig_wpack_multi": (c_int, [_P, c_int32, c_int64, c_int64, _P]),
    "msig_wpack_part_elems": (c_size_t, [POINTER(WpackDesc), c_int32]),
    "msig_wpack_part": (c_int, [POINTER(WpackDesc), c_int32, c_int32, _P, _P, _P]),
    "msig_conv2d_fwd": (c_int, [POINTER(ConvGeom), _P, _P, POINTER(Epilogue), _P, _P]),
    "msig_img_pad8": (c_int, [_P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P, _P, _P]),
    "msig_conv_rowpatch_fwd": (c_int, [POINTER(ConvGeom), _P, _P, POINTER(Epilogue), _P, _P]),
    "msig_conv_rowpatch_wgrad_workspace": (c_size_t, [POINTER(ConvGeom)]),
    "msig_conv_rowpatch_wgrad": (c_int, [POINTER(ConvGeom), _P, _P, c_int, _P, c_int, _P, c_size_t, _P]),
    "msig_conv_narrow_fwd": (c_int, [POINTER(ConvGeom), _P, _P, POINTER(Epilogue), _P, _P]),
    "msig_reflect_fold_nchw": (c_int, [_P, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "msig_conv2d_dgrad": (c_int, [POINTER(ConvGeom), _P, _P, POINTER(Epilogue), _P, _P]),
    "msig_conv2d_wgrad_workspace": (c_size_t, [POINTER(ConvGeom)]),
    "msig_conv2d_wgrad": (c_int, [POINTER(ConvGeom), _P, _P, _P, c_int, _P, c_size_t, _P]),
    "msig_convT2d_fwd": (c_int, [POINTER(ConvGeom), _P, _P, POINTER(Epilogue), _P, _P]),
    "msig_convT2d_dgrad": (c_int, [POINTER(ConvGeom), _P, _P, POINTER(Epilogue), _P, _P]),
    "msig_convT2d_wgrad_workspace": (c_size_t, [POINTER(ConvGeom)]),
    "msig_convT2d_wgrad": (c_int, [POINTER(ConvGeom), _P, _P, _P, c_int, _P, c_size_t, _P]),
    "msig_patch_gather": (c_int, [POINTER(PatchGeom), _P, _P, _P, _P, _P]),
    "msig_patch_scatter": (c_int, [POINTER(PatchGeom), _P, _P, _P, c_int, _P]),
    "msig_patch_wgrad_workspace": (c_size_t, [c_int64, c_int32, c_int32]),
    "msig_patch_wgrad": (c_int, [POINTER(WpackDesc), c_int64, _P, _P, _P, c_int, _P, c_size_t, _P]),
    "msig_patch_wgrad_part": (c_int, [POINTER(WpackDesc), c_int32, c_int32, c_int64, _P, c_int32,
                                      _P, c_int32, _P, c_int, _P, c_size_t, _P]),
    "msig_in_stats_workspace": (c_size_t, [c_int32, c_int32, c_int32]),
    "msig_in_stats": (c_int, [_P, c_int32, c_int32, c_int32, c_float, _P, _P, c_int64, _P, _P, _P,
                              _P, _P, c_size_t, _P]),
    "msig_norm_act_fwd": (c_int, [_P, _P, _P, _P, c_int32, c_float, c_int32, c_int32, c_int32, _P,
                                  _P]),
    "msig_norm_act_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_float, c_int32,
                                  c_int32, c_int32, _P, _P, _P, c_int64, c_int, _P, c_size_t, _P]),
    "msig_norm_act_fwd_pad": (c_int, [_P, _P, _P, c_int32, c_float, c_int32, c_int32, c_int32, c_int32, c_int32, _P,
                                      _P]),
    "msig_norm_act_bwd_pad_workspace": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "msig_norm_act_bwd_pad": (c_int, [_P, _P, _P, _P, _P, _P, c_int32, c_float, c_int32, c_int32, c_int32, c_int32,
                                      c_int32, _P, _P, c_size_t, _P]),
    "msig_epilogue_stats_rows": (c_int32, [c_int32, c_int32, c_int32]),
    "msig_ring_stats_rows": (c_int32, [c_int32, _P]),
    "msig_in_stats_from_partials": (c_int, [_P, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, _P, _P,
                                            c_int64, _P, _P, _P, _P, _P]),
    "msig_norm_bwd_from_partials": (c_int, [_P, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P, c_int32,
                                            c_int32, _P, _P, _P, c_int64, c_int, _P, _P]),
    "msig_norm_act_fwd_from_partials": (c_int, [_P, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, _P, _P,
                                                c_int64, _P, _P, _P, _P, _P, _P, c_int32, c_float, _P, _P]),
    "msig_norm_bwd_from_partials_fused": (c_int, [_P, c_int32, c_int32, c_int32, _P, _P, _P, _P, _P, _P, c_int32,
                                                  c_int32, _P, _P, _P, c_int64, c_int, _P, _P]),
    "msig_act_bwd": (c_int, [_P, _P, c_int32, c_float, c_int64, _P, _P]),
    "msig_colsum_workspace": (c_size_t, [c_int64, c_int32]),
    "msig_colsum": (c_int, [_P, c_int64, c_int32, _P, c_int, _P, c_size_t, _P]),
    "msig_colsum_f32": (c_int, [_P, c_int64, c_int32, c_int64, _P, c_int, _P]),
    "msig_nchw_chansum_workspace": (c_size_t, [c_int32, c_int32, c_int64]),
    "msig_nchw_chansum": (c_int, [_P, c_int32, c_int32, c_int64, c_int64, _P, c_int, _P, c_size_t, _P]),
    "msig_gemm_tn_partial": (c_int, [c_int64, _P, c_int32, _P, c_int32, _P, c_size_t, POINTER(c_int32), _P]),
    "msig_multi_linear_grads": (c_int, [_P, c_int32, c_int64, _P, c_int64, c_int64, c_int32, c_int32, c_int32, _P, _P,
                                        _P]),
    "msig_wgrad_unpack": (c_int, [POINTER(WpackDesc), c_int32, c_int32, _P, c_int32, c_int64, _P, c_int, _P]),
    "msig_maxpool2_fwd": (c_int, [_P, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "msig_maxpool2_bwd": (c_int, [_P, _P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "msig_avgpool_fwd": (c_int, [_P, c_int32, c_int32, c_int32, _P, _P]),
    "msig_avgpool_bwd": (c_int, [_P, c_int32, c_int32, c_int32, _P, _P]),
    "msig_head_gather": (c_int, [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "msig_head_scatter": (c_int, [_P, _P, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "msig_f32_to_bf16": (c_int, [_P, c_int64, _P, _P]),
    "msig_bf16_to_f32": (c_int, [_P, c_int64, _P, _P]),
    "msig_tanh_bwd": (c_int, [_P, _P, c_int64, _P, _P]),
    "msig_reduce_workspace": (c_size_t, []),
    "msig_l1_loss_f32_fwd": (c_int, [_P, _P, c_int64, _P, _P, c_size_t, _P]),
    "msig_l1_loss_f32_bwd": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "msig_l1_loss_bf16_fwd": (c_int, [_P, _P, c_int64, _P, _P, c_size_t, _P]),
    "msig_l1_loss_bf16_bwd": (c_int, [_P, _P, c_int64, _P, _P, _P, _P]),
    "msig_mse_const_fwd": (c_int, [_P, c_float, c_int64, _P, _P, c_size_t, _P]),
    "msig_mse_loss_fwd": (c_int, [_P, _P, c_int64, _P, _P, c_size_t, _P]),
    "msig_mse_loss_bwd": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "msig_mse_const_bwd": (c_int, [_P, c_float, c_int64, _P, _P, _P]),
    "msig_gram_workspace": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "msig_gram_fwd": (c_int, [_P, c_int32, c_int32, c_int32, c_int32, _P, _P, c_size_t, _P]),
    "msig_gram_l1_workspace": (c_size_t, [c_int32]),
    "msig_gram_l1": (c_int, [_P, _P, c_int32, _P, c_int, _P, _P, c_size_t, _P]),
    "msig_gram_bwd": (c_int, [_P, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P, _P, c_int, _P, _P]),
    "msig_augment_workspace": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "msig_augment_u8": (c_int, [_P, c_int32, c_int32, c_int32, _P, _P, c_int32, _P, _P, c_size_t, _P]),
    "msig_sumsq": (c_int, [_P, c_int64, _P, c_int, _P, c_size_t, _P]),
    "msig_adam_step": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_float, c_float, c_float, c_float,
                               c_float, c_float, c_int32, c_float, _P]),
    "msig_adam_step_dev": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_float, c_float, c_float, c_float,
                                   c_float, c_float, _P, c_float, _P]),
}

_lib = None
_inited_devices = set()


def load():
    """Load libmsig.so (no device needed). Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). There is no fallback path.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGS.keys())


def last_error():
    return load().msig_last_error().decode()


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"libmsig {what} failed ({rc}): {last_error()}")


def init(device=0):
    """Bind the library to a CUDA device (sm_100a). Raises RuntimeError without a GPU."""
    lib = load()
    if device not in _inited_devices:
        check(lib.msig_init(int(device)), "msig_init")
        _inited_devices.add(device)
    return lib


def call(name, *args):
    """Call an entry point that returns a status code; raise on failure."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")
