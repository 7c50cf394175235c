"""ctypes binding of libmickey_b200.so (the C ABI declared in include/mickey_b200.h).

There is no fallback: if the shared library is missing or does not load, importing the binding
raises with the build command — the product path never routes around the CUDA extension.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libmickey_b200.so")


class MkConfig(C.Structure):
    _fields_ = [
        ("embed_dim", C.c_int), ("depth", C.c_int), ("heads", C.c_int),
        ("down_factor", C.c_int),
        ("block_dims", C.c_int * 4),
        ("desc_dim", C.c_int),
        ("use_softmax", C.c_int), ("depth_sigmoid", C.c_int), ("max_depth", C.c_float),
        ("kp_pos_enc", C.c_int), ("dsc_pos_enc", C.c_int), ("norm_dsc", C.c_int),
        ("temperature", C.c_float), ("use_dustbin", C.c_int),
        ("it_matches", C.c_int), ("it_ransac", C.c_int),
        ("num_sampled", C.c_int), ("num_corr", C.c_int), ("num_refine", C.c_int),
        ("th_inlier", C.c_float), ("th_soft_inlier", C.c_float),
    ]


class MkGemmArgs(C.Structure):
    _fields_ = [
        ("epi", C.c_int), ("impl", C.c_int),
        ("a", C.c_void_p), ("a_rows", C.c_longlong), ("a_cols", C.c_longlong), ("a_ld", C.c_longlong),
        ("b", C.c_void_p), ("b_rows", C.c_longlong), ("b_cols", C.c_longlong), ("b_ld", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("k_chunks", C.c_int), ("chunks_per_tap", C.c_int), ("num_taps", C.c_int),
        ("tap_shift", C.c_int * 9),
        ("groups", C.c_int), ("a_row_group_off", C.c_int), ("a_col_group_off", C.c_int), ("a_col_base", C.c_int),
        ("b_row_group_off", C.c_int),
        ("act", C.c_int),
        ("bias", C.c_void_p), ("bias_group_off", C.c_int),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("ln_group_off", C.c_int),
        ("out_f", C.c_void_p), ("out_f_ld", C.c_longlong), ("out_f_group_off", C.c_longlong),
        ("out_h", C.c_void_p), ("out_h_ld", C.c_longlong), ("out_h_group_off", C.c_longlong),
        ("res_h", C.c_void_p), ("res_h_ld", C.c_longlong), ("res_h_group_off", C.c_longlong),
        ("aux", C.c_void_p), ("aux_group_mask", C.c_int),
        ("pad_h2", C.c_int), ("pad_w2", C.c_int), ("tok_per_img", C.c_int),
        ("eps", C.c_float),
        ("n_valid", C.c_int), ("inv_temp", C.c_float),
        ("dustbin", C.c_void_p),
        ("part_row", C.c_void_p), ("part_col", C.c_void_p), ("part_ld", C.c_int),
        ("lse_r", C.c_void_p), ("lse_c", C.c_void_p), ("scr0", C.c_void_p), ("scr1", C.c_void_p),
        ("scores", C.c_void_p), ("kp_scores", C.c_void_p), ("final_scores", C.c_void_p),
        ("lse_bound", C.c_float),
        ("out_pitch", C.c_longlong),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "mk_create": (C.c_int, [C.c_int, C.POINTER(MkConfig), C.POINTER(C.c_void_p)]),
    "mk_destroy": (C.c_int, [C.c_void_p]),
    "mk_last_error": (C.c_char_p, []),
    "mk_version": (C.c_char_p, []),
    "mk_sizeof_config": (C.c_int, []),
    "mk_sizeof_gemm_args": (C.c_int, []),
    "mk_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_longlong]),
    "mk_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mk_workspace_bytes": (C.c_longlong, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "mk_workspace_offset": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "mk_extract": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "mk_extract_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "mk_match": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p]),
    "mk_solve_pose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "mk_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_ulonglong,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                             C.c_void_p]),
    "mk_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_ulonglong,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                C.c_void_p]),
    "mk_pose_to_submission": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mk_launch_count": (C.c_longlong, [C.c_void_p]),
    "mk_set_seed": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_void_p]),
    "mk_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mk_profile_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "mk_op_gemm": (C.c_int, [C.POINTER(MkGemmArgs), C.c_void_p]),
    "mk_op_patch_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_void_p]),
    "mk_op_ingest_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_void_p]),
    "mk_op_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p]),
    "mk_op_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mk_op_linattn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "mk_op_matcher_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mk_op_sample": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_ulonglong, C.c_void_p, C.c_longlong,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "mk_op_sample_workspace_bytes": (C.c_longlong, [C.c_int, C.c_int]),
}

_lib = None


class MickeyB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every exported signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MickeyB200Error(
            f"{LIB_PATH} not found. Build it with `python -m mickey_b200.build` (needs nvcc; sm_100a only). "
            "mickey_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.mk_sizeof_config() != C.sizeof(MkConfig) or lib.mk_sizeof_gemm_args() != C.sizeof(MkGemmArgs):
        raise MickeyB200Error("ctypes struct layout does not match the library (stale build?): "
                              f"mk_config {lib.mk_sizeof_config()} vs {C.sizeof(MkConfig)}, "
                              f"mk_gemm_args {lib.mk_sizeof_gemm_args()} vs {C.sizeof(MkGemmArgs)}")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().mk_last_error().decode(errors="replace")
        raise MickeyB200Error(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device (or host) address of a torch tensor, or None."""
    return None if t is None else C.c_void_p(t.data_ptr())
