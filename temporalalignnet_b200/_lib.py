"""ctypes binding of libtan_b200.so (the C ABI declared in include/tan_b200.h).

There is no fallback: if the shared library is missing or a call fails, a `TanError` is raised.
Build the library with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C temporalalignnet_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TAN_LIB_PATH") or os.path.join(_HERE, "libtan_b200.so")   # (override: A/B of builds)
CSRC_DIR = os.path.join(_HERE, "csrc")

TAN_OK = 0
ERR_NAMES = {-1: "TAN_ERR_SHAPE", -2: "TAN_ERR_ARCH", -3: "TAN_ERR_WORKSPACE", -4: "TAN_ERR_CUDA",
             -5: "TAN_ERR_ARG"}
ACT_NONE, ACT_QUICKGELU, ACT_RELU = 0, 1, 2
ABI_VERSION = 5


class TanError(RuntimeError):
    pass


class LnArgs(C.Structure):
    """struct tan_ln_args (include/tan_b200.h)."""
    _fields_ = [
        ("in_", C.c_void_p), ("in_is_bf16", C.c_int), ("rows", C.c_int), ("d", C.c_int),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("add", C.c_void_p), ("add_rows", C.c_int),
        ("L_in", C.c_int), ("L_out", C.c_int), ("l_off", C.c_int),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p),
        ("l_split", C.c_int), ("strideA", C.c_int64), ("strideB", C.c_int64),
        ("rawA_f32", C.c_void_p), ("rawB_f32", C.c_void_p),
        ("nrmA_bf16", C.c_void_p), ("nrmB_bf16", C.c_void_p),
        ("nrmA_f32", C.c_void_p), ("nrmB_f32", C.c_void_p),
        ("raw_strideA", C.c_int64), ("raw_strideB", C.c_int64),
    ]


class SimGeom(C.Structure):
    """struct tan_sim_geom (include/tan_b200.h)."""
    _fields_ = [("B_loc", C.c_int), ("S", C.c_int), ("T", C.c_int), ("C", C.c_int), ("N", C.c_int),
                ("d", C.c_int), ("b_off", C.c_int), ("col_off", C.c_void_p)]


class OptimTensor(C.Structure):
    """struct tan_optim_tensor (include/tan_b200.h)."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("ema", C.c_void_p), ("numel", C.c_int64), ("decay", C.c_float), ("neg_step_size", C.c_float)]


# symbol -> (restype, argtypes); every symbol include/tan_b200.h declares
SIGNATURES = {
    "tan_abi_version": (C.c_int, []),
    "tan_last_error_string": (C.c_char_p, []),
    "tan_device_check": (C.c_int, []),
    "tan_debug_set_trace": (C.c_int, [C.c_void_p]),
    "tan_cast_f32_to_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tan_linear_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p]),
    "tan_linear_dual_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tan_linear_gelu_bwd_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                           C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tan_linear_res_ln_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p]),
    "tan_linear_res_ln_stage_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                               C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                               C.c_int64, C.c_void_p]),
    "tan_layernorm": (C.c_int, [C.POINTER(LnArgs), C.c_void_p]),
    "tan_attention_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "tan_pos_from_time": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p]),
    "tan_sim_nce_workspace_bytes": (C.c_size_t, [C.POINTER(SimGeom)]),
    "tan_sim_nce_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(SimGeom), C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    "tan_nce_from_logits": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(SimGeom), C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tan_nce_reduce": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tan_own_clip_sim": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tan_agree_scan_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "tan_agree_scan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tan_agree_targets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    # ---- backward pass -------------------------------------------------------------------------------
    "tan_transpose_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "tan_colsum_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "tan_colsum": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                             C.c_size_t, C.c_void_p]),
    "tan_quickgelu_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tan_quickgelu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tan_layernorm_bwd_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "tan_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "tan_l2norm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64,
                                 C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "tan_batch_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p]),
    "tan_sim_grad_tiles": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(SimGeom), C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "tan_sim_grad_gemm": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(SimGeom),
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tan_gemm_tn_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "tan_gemm_tn_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tan_embed_gather_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "tan_text_pool_fc1": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tan_text_pool_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tan_optim_adamw_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tan_ema_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]),
    "tan_align_stitch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tan_align_argmax": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tan_attention_bwd_bf16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libtan_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise TanError("building libtan_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise TanError(f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU "
                           f"fallback; run `make -C {CSRC_DIR}`")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)          # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if h.tan_abi_version() != ABI_VERSION:
            raise TanError(f"ABI mismatch: library {h.tan_abi_version()} vs binding {ABI_VERSION}")
        _lib = h
    return _lib


def check(code: int, what: str = "") -> None:
    if code != TAN_OK:
        msg = lib().tan_last_error_string().decode(errors="replace")
        raise TanError(f"{what}: {ERR_NAMES.get(code, code)}: {msg}")
