"""ctypes binding of the C ABI in include/routeformer_b200.h (the only way Python reaches the kernels).

There is no CPU or PyTorch fallback: if the shared library cannot be loaded every op raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RF_LIB_PATH") or os.path.join(HERE, "_lib", "librouteformer_b200.so")  # RF_LIB_PATH: A/B-test a prebuilt library

c_fp = C.c_void_p  # device pointers travel as plain addresses
c_ll = C.c_longlong


class RfFovCropParams(C.Structure):
    _fields_ = [
        ("frames", c_fp), ("src_dtype", C.c_int), ("frame_ids", c_fp),
        ("n_frames", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("centers", c_fp), ("windows", c_fp),
        ("mean", C.c_float * 3), ("inv_std", C.c_float * 3),
        ("out_size", C.c_int), ("patch", C.c_int), ("out", c_fp), ("out_dtype", C.c_int), ("out_ld", c_ll),
    ]


class RfGemmParams(C.Structure):
    _fields_ = [
        ("A", c_fp), ("lda", c_ll), ("a_mn_major", C.c_int),
        ("B", c_fp), ("ldb", c_ll), ("b_mn_major", C.c_int),
        ("C", c_fp), ("ldc", c_ll),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("bias", c_fp),
        ("rowadd", c_fp), ("rowadd_period", C.c_int), ("ld_rowadd", c_ll),
        ("residual", c_fp), ("ld_res", c_ll),
        ("act", C.c_int),
        ("preact", c_fp), ("ld_pre", c_ll),
        ("dact_aux", c_fp), ("ld_aux", c_ll), ("dact", C.c_int),
        ("accumulate", C.c_int), ("split_k", C.c_int),
        ("out_group_in", C.c_int), ("out_group_out", C.c_int), ("out_row_offset", C.c_int),
        ("round_f16", C.c_int), ("ab_dtype", C.c_int), ("colsum_a", c_fp), ("c_dtype", C.c_int),
    ]


class RfConv3AssembleParams(C.Structure):
    _fields_ = [
        ("z", c_fp), ("ldz", c_ll), ("y", c_fp), ("ldy", c_ll),
        ("n_seq", C.c_int), ("L", C.c_int), ("D", C.c_int), ("pad", C.c_int),
        ("bias", c_fp), ("pe", c_fp), ("ld_pe", c_ll), ("wtime", c_fp),
    ]


class RfConv3AssembleBwdParams(C.Structure):
    _fields_ = [
        ("dy", c_fp), ("ldy", c_ll), ("dz", c_fp), ("ldz", c_ll),
        ("n_seq", C.c_int), ("L", C.c_int), ("D", C.c_int), ("pad", C.c_int),
        ("dbias", c_fp), ("dwtime", c_fp),
    ]


class RfAttnParams(C.Structure):
    _fields_ = [
        ("q", c_fp), ("q_bs", c_ll), ("q_ls", c_ll),
        ("k", c_fp), ("k_bs", c_ll), ("k_ls", c_ll),
        ("v", c_fp), ("v_bs", c_ll), ("v_ls", c_ll),
        ("B", C.c_int), ("H", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int), ("dh", C.c_int),
        ("mode", C.c_int), ("out_layout", C.c_int),
        ("idx", c_fp), ("idx_group", C.c_int), ("U", C.c_int), ("u", C.c_int),
        ("out", c_fp), ("top", c_fp), ("measure", c_fp), ("forced_top", c_fp),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_ulonglong), ("dropout_offset", C.c_ulonglong), ("dropout_offset_base", c_fp),
        ("tail_only", C.c_int),
    ]


class RfAttnBwdParams(C.Structure):
    _fields_ = [("f", RfAttnParams), ("dout", c_fp), ("dq", c_fp), ("dk", c_fp), ("dv", c_fp)]


class RfDistilParams(C.Structure):
    _fields_ = [
        ("z", c_fp), ("B", C.c_int), ("Lz", C.c_int), ("D", C.c_int),
        ("gamma", c_fp), ("beta", c_fp), ("running_mean", c_fp), ("running_var", c_fp),
        ("training", C.c_int), ("momentum", C.c_float), ("eps", C.c_float),
        ("mean", c_fp), ("rstd", c_fp), ("out", c_fp), ("argmax", c_fp),
    ]


class RfDistilBwdParams(C.Structure):
    _fields_ = [
        ("z", c_fp), ("B", C.c_int), ("Lz", C.c_int), ("D", C.c_int),
        ("gamma", c_fp), ("beta", c_fp), ("mean", c_fp), ("rstd", c_fp), ("training", C.c_int),
        ("argmax", c_fp), ("dout", c_fp), ("dz", c_fp), ("dgamma", c_fp), ("dbeta", c_fp), ("scratch", c_fp),
    ]


class RfAreaResizeParams(C.Structure):
    _fields_ = [
        ("src", c_fp), ("src_plane_stride", c_ll), ("src_row_pitch", c_ll),
        ("n_planes", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("dst", c_fp), ("dH", C.c_int), ("dW", C.c_int),
    ]


STRUCTS = {
    0: RfFovCropParams, 1: RfGemmParams, 2: RfConv3AssembleParams, 3: RfConv3AssembleBwdParams,
    4: RfAttnParams, 5: RfAttnBwdParams, 6: RfDistilParams, 7: RfDistilBwdParams, 8: RfAreaResizeParams,
}

_I, _F, _P, _L = C.c_int, C.c_float, c_fp, c_ll
# name -> argtypes (restype is int unless noted)
SIGNATURES = {
    "rf_fov_crop": [C.POINTER(RfFovCropParams), _P],
    "rf_gemm_tf32": [C.POINTER(RfGemmParams), _P],
    "rf_conv3_assemble_fwd": [C.POINTER(RfConv3AssembleParams), _P],
    "rf_conv3_assemble_bwd": [C.POINTER(RfConv3AssembleBwdParams), _P],
    "rf_conv3_pack_weight": [_P, _P, _I, _I, _L, _P],
    "rf_conv3_unpack_grad": [_P, _P, _I, _I, _L, _P],
    "rf_attention_fwd": [C.POINTER(RfAttnParams), _P],
    "rf_attention_bwd": [C.POINTER(RfAttnBwdParams), _P],
    "rf_layernorm_fwd": [_P, _L, _P, _P, _P, _L, _P, _P, _I, _I, _P],
    "rf_layernorm_bwd": [_P, _L, _P, _L, _P, _P, _P, _P, _L, _P, _P, _I, _I, _P],
    "rf_distil_fwd": [C.POINTER(RfDistilParams), _P],
    "rf_distil_bwd": [C.POINTER(RfDistilBwdParams), _P],
    "rf_motion_features": [_P, _P, _L, _P, _L, _P, _I, _I, _I, _I, _I, _F, _F, _I, _P],
    "rf_decoder_input_fwd": [_P, _P, _I, _I, _I, _L, _I, _P],
    "rf_decoder_input_bwd": [_P, _P, _I, _I, _I, _L, _I, _P],
    "rf_stream_tokens_fwd": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P],
    "rf_stream_tokens_bwd": [_P, _P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P],
    "rf_decode_waypoints_fwd": [_P, _L, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P],
    "rf_decode_waypoints_bwd": [_P, _P, _P, _L, _I, _I, _I, _I, _F, _P],
    "rf_median_downsample": [_P, _P, _I, _I, _I, _I, _P],
    "rf_ade_fde": [_P, _P, _I, _I, _P, _P, _P],
    "rf_dropout": [_P, _L, _P, _L, _P, _L, _I, _I, _F, C.c_ulonglong, C.c_ulonglong, _P, _P],
    "rf_eval_samples": [_P, _P, _I, _I, _I, _F, _F, _I, _P, _P, _P],
    "rf_discounted_loss_fwd": [_P, _L, _P, _L, _I, _I, _I, _F, _F, _I, _P, _P],
    "rf_discounted_loss_bwd": [_P, _L, _P, _L, _I, _I, _I, _F, _F, _I, _P, _F, _P, _L, _I, _P],
    "rf_colsum_accumulate": [_P, _L, _I, _I, _P, _P],
    "rf_sumsq_accumulate": [_P, _L, _P, _P],
    "rf_adamw_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P, _F, _P],
    "rf_area_resize_u8": [C.POINTER(RfAreaResizeParams), _P],
    "rf_struct_size": [_I],
    "rf_debug_gemm_stamps": [_P],
    "rf_debug_gemm_probe": [C.c_int],
    "rf_debug_attn_stamps": [_P],
    "rf_stage_frames_h2d": [_P, _P, _I, _I, C.POINTER(C.c_int), _I, _L, _P],
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> C.CDLL:
    """Loads (building in-tree first if needed and nvcc is present) the CUDA library. Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and not os.environ.get("RF_LIB_PATH"):
        try:
            from . import build as _build

            _build.build(verbose=False)
        except Exception as e:  # no nvcc / compile error: fall through to a plain load attempt
            if not os.path.exists(LIB_PATH):
                raise LibraryMissing(f"cannot build {LIB_PATH}: {e}") from e
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(f"{LIB_PATH} is missing: run `python -m routeformer_b200.build` (there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    lib.rf_last_error.restype = C.c_char_p
    lib.rf_last_error.argtypes = []
    lib.rf_abi_version.restype = C.c_int
    lib.rf_abi_version.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


class RfError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rf_last_error().decode(errors="replace")
        raise RfError(f"{what} failed with code {rc}: {msg}")
