"""ctypes binding of libcmpc_b200.so (the C ABI declared in include/cmpc_b200.h).

There is no fallback: if the library is missing or a call fails, a CmpcError is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libcmpc_b200.so"


class CmpcError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a1", C.c_void_p), ("lda1", C.c_int64), ("k1", C.c_int32),
        ("a2", C.c_void_p), ("lda2", C.c_int64), ("k2", C.c_int32),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("w_batch_stride", C.c_int64), ("w_rows", C.c_int32),
        ("m", C.c_int32), ("n", C.c_int32),
        ("rows_per_sample", C.c_int32),
        ("row_scale", C.c_void_p),
        ("bias", C.c_void_p),
        ("sbias", C.c_void_p), ("ld_sbias", C.c_int64),
        ("gate", C.c_void_p), ("ld_gate", C.c_int64),
        ("act", C.c_int32),
        ("group_width", C.c_int32), ("group_valid", C.c_int32),
        ("peep_i", C.c_void_p), ("peep_f", C.c_void_p), ("ld_peep", C.c_int64),
        ("cprev", C.c_void_p), ("ld_cprev", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_fp32", C.c_int32),
        ("row_sumsq", C.c_void_p),
        ("stats", C.c_void_p),
        ("peep_f16", C.c_int32),
        ("a_row_sumsq", C.c_void_p),
    ]


class MutanArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64), ("k", C.c_int32),
        ("a_row_sumsq", C.c_void_p),
        ("w", C.c_void_p), ("ldw", C.c_int64),
        ("m", C.c_int32), ("c", C.c_int32),
        ("rows_per_sample", C.c_int32),
        ("bias", C.c_void_p), ("ld_bias", C.c_int64),
        ("lang", C.c_void_p), ("ld_lang", C.c_int64), ("lang_batch_stride", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("row_sumsq", C.c_void_p),
        ("out_f16", C.c_int32),
    ]


class ConvLstmBwdArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("y16", "opre", "cnew", "cn", "cprev", "mr_g", "mr_o", "ln_gamma", "ln_beta", "w_ci", "w_cf", "w_co", "dh")] + \
               [("ld_dh", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("dcn_in", "sums", "dcnew", "dy16", "dcprev_out", "dw_ci", "dw_cf", "dw_co", "dgamma", "dbeta",
                                          "ws_sample", "ws_chan")] + \
               [("gw", C.c_int32), ("m", C.c_int32), ("rows_per_sample", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    """Loads the shared library (building is the job of __graft_entry__.build / build.py)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise CmpcError(f"{LIB_PATH} not found: run `python -m cmpc_refseg_b200.build` "
                            "(libcmpc_b200 has no CPU or PyTorch fallback)")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.cmpc_last_error.restype = C.c_char_p
        _lib.cmpc_version.restype = C.c_int
        for name in dir(_Sigs):
            if name.startswith("cmpc_"):
                fn = getattr(_lib, name)
                fn.restype = C.c_int
                fn.argtypes = getattr(_Sigs, name)
        for name in _SIZE_FNS:
            fn = getattr(_lib, name)
            fn.restype = C.c_size_t
            fn.argtypes = _SIZE_FNS[name]
        _lib.cmpc_gemm_set_mode.restype = None
        _lib.cmpc_gemm_set_mode.argtypes = [C.c_int]
        _lib.cmpc_graph_set_mode.restype = None
        _lib.cmpc_graph_set_mode.argtypes = [C.c_int]
        _lib.cmpc_ln_relu_l2norm_set_mode.restype = None
        _lib.cmpc_ln_relu_l2norm_set_mode.argtypes = [C.c_int]
        import os
        _lib.cmpc_set_pdl.restype = None
        _lib.cmpc_set_pdl.argtypes = [C.c_int]
        if os.environ.get("CMPC_PDL"):                # programmatic dependent launch of the kernels that support it
            _lib.cmpc_set_pdl(int(os.environ["CMPC_PDL"]))
        if os.environ.get("CMPC_LN_L2_MODE"):         # measurement knob: 1 = register-file ln_relu_l2norm kernels only
            _lib.cmpc_ln_relu_l2norm_set_mode(int(os.environ["CMPC_LN_L2_MODE"]))
        if os.environ.get("CMPC_GRAPH_MODE"):         # measurement knob: 2 = 2-SM MMA graph kernel variant
            _lib.cmpc_graph_set_mode(int(os.environ["CMPC_GRAPH_MODE"]))
        if os.environ.get("CMPC_GEMM_MODE"):          # measurement knob: 1 = weight-multicast GEMM instead of the 2-SM MMA
            _lib.cmpc_gemm_set_mode(int(os.environ["CMPC_GEMM_MODE"]))
    return _lib


class _Sigs:
    cmpc_gemm_f16 = [C.POINTER(GemmArgs), C.c_void_p]
    cmpc_mutan_f16 = [C.POINTER(MutanArgs), C.c_void_p]
    _p, _i64, _i32, _f, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
    cmpc_affinity_softmax = [_p, _p, _i32, _i32, _i32, _f, _p, _p, _p, _p, _p, _sz, _p]
    cmpc_affinity_softmax_scaled = [_p, _p, _i32, _i32, _i32, _f, _p, _p, _p, _p, _p, _p, _sz, _p]
    cmpc_graph_reason_f16 = [_p, _p, _p, _i64, _i32, _i32, _i32, _f, _p, _i64, _p, _p, _p]
    cmpc_ln_residual_relu_f16 = [_p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _p]
    cmpc_ln_relu_l2norm_f16 = [_p, _i64, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _p, _p]
    cmpc_ln_relu_l2norm_scaled_f16 = [_p, _i64, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _p, _p, _p]
    cmpc_ln_residual_relu_scaled_f16 = [_p, _i64, _p, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _p]
    cmpc_ln_finalize = [_p, _i32, C.c_double, _p, _p]
    cmpc_cast_f32_f16 = [_p, _i64, _p, _i64, _i64, _i32, _p]
    cmpc_transpose_cast_f32_f16 = [_p, _i64, _i32, _i32, _p, _i64, _i32, _i64, _p]
    cmpc_scale_cast_f32_f16 = [_p, _i64, _f, _p, _i64, _i64, _i32, _p]
    cmpc_rownorm_f16 = [_p, _i64, _p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p]
    cmpc_rownorm_h16 = [_p, _i64, _p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p]
    cmpc_spatial_fixup_f16 = [_p, _i64, _p, _i64, _i32, _i32, _i32, _p]
    cmpc_add3_l2norm_f16 = [_p, _p, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _p]
    cmpc_add3_l2norm_ld_f16 = [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _p, _p]
    cmpc_global_pool_f16 = [_p, _p, _p, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, _f, _p, _i64, _p, _p, _sz, _p]
    cmpc_words_prepare = [_p, _i32, _i32, _p, _p, _i64, _p, _p]
    cmpc_lang_parse = [_p, _i64, _i32, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _i64, _p]
    cmpc_small_linear_f32 = [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _p]
    cmpc_gv_gates = [_p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _i64, _p]
    cmpc_gv_gates_ex = [_p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _p, _p, _i64, _i64, _p, _i64, _i64, _i64, _i32, _p, _p]
    cmpc_gv_gates_batch = [_p, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _i64, _i32, _p, _p]
    cmpc_convlstm_gates1 = [_p, _i32, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _p]
    cmpc_convlstm_gates2 = [_p, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _i64, _i32, _p]
    cmpc_convlstm_gates2_y16 = [_p, _i64, _p, _p, _i32, _i32, _p, _p, _p, _p, _p, _i32, _i64, _i32, _p]
    cmpc_convlstm_gates1_h16 = [_p, _i64, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _p]
    cmpc_score_upsample = [_p, _i64, _p, _f, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _sz, _p]
    cmpc_score_from_taps = [_p, _i64, _f, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p]
    cmpc_sigmoid_ce_sums = [_p, _p, _i32, _i64, _p, _p]
    cmpc_iou_counts = [_p, _p, _i32, _i64, _f, _i32, _p, _p]
    cmpc_postprocess_iou = [_p, _i32, _i32, _i32, _f, _p, _p, _p, _i32, _p, _p, _p]
    cmpc_gemm_atb_f16 = [_p, _i64, _i32, _p, _i64, _i32, _i32, _p, _i64, _i32, _p]
    cmpc_mutan_out_bwd = [_p, _p, _p, _p, _i64, _p, _i64, _p, _p, _i64, _i64, _i32, _p]
    cmpc_mutan_bwd_f16 = [C.POINTER(MutanArgs), _p, _i64, _p, _i64, _p, _p]
    cmpc_lateral_bwd = [_p, _i64, _p, _i64, _p, _p, _p, _i32, _i32, _i32, _p]
    cmpc_act_bwd_f32 = [_p, _p, _p, _i64, _i32, _p]
    cmpc_lang_bwd = [_p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p]
    cmpc_l2norm_bwd_f32 = [_p, _p, _p, _i32, _i32, _p, _p]
    cmpc_relu_bwd_f32 = [_p, _p, _p, _i32, _i32, _i64, _p]
    cmpc_adam_f32 = [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _f, _p, _p]
    cmpc_embed_gather_f16 = [_p, _p, _i32, _i32, _i32, _p, _i64, _p]
    cmpc_lstm_step = [_p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _i64, _p, _p]
    cmpc_lstm_step_train = [_p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _i64, _p, _p, _p]
    cmpc_lstm_step_bwd = [_p, _p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _f, _p, _i64, _p, _p]
    cmpc_embed_scatter_add = [_p, _p, _i64, _f, _i32, _i32, _i32, _p, _p]
    cmpc_relu_mask_f16 = [_p, _i64, _p, _i64, _p, _p, _i32, _i32, _i32, _p]
    cmpc_ln_bwd_sums = [_p, _i64, _p, _p, _p, _i64, _p, _p, _p, _i64, _p, _p, _p, _i32, _i32, _i32, _p]
    cmpc_ln_bwd_apply = [_p, _i64, _p, _i64, _p, _p, _p, _p, _p, _i32, _i32, _i32, _p]
    cmpc_affinity_bwd = [_p, _p, _p, _p, _p, _p, _f, _i32, _i32, _p, _p, _p, _p]
    cmpc_transpose_gt_f16 = [_p, _i64, _i32, _i32, _i32, _p, _p]
    cmpc_exg_bwd_rows = [_p, _i64, _p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p]
    cmpc_pool_bwd_rows = [_p, _i64, _p, _i64, _p, _p, _i64, _p, _i64, _f, _p, _p, _i64, _p, _i64, _p, _p, _i64, _i32, _i32, _i32, _p]
    cmpc_gv_gates_bwd = [_p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _i64, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _p]
    cmpc_gv_gates_bwd_batch = [_p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _i64, _i32, _i32, _i32, _i64, _p, _p, _p, _p, _i32, _p, _p, _p]
    cmpc_small_atb_f32 = [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p]
    cmpc_score_bwd_dpred = [_p, _p, _f, _i32, _i32, _i32, _i32, _i32, _p, _p, _p]
    cmpc_score_bwd_taps = [_p, _i32, _i32, _i32, _p, _i32, _p]
    cmpc_gemm_atb_batched_f16 = [_p, _i64, _i32, _p, _i64, _i32, _i32, _i32, _p, _i64, _i64, _p]
    cmpc_convlstm_bwd = [_i32, C.POINTER(ConvLstmBwdArgs), _i32, _p]


_SIZE_FNS = {
    "cmpc_affinity_workspace_bytes": [C.c_int32],
    "cmpc_global_pool_workspace_bytes": [C.c_int32, C.c_int32, C.c_int32],
    "cmpc_score_workspace_bytes": [C.c_int64],
    "cmpc_convlstm_bwd_workspace_floats": [C.c_int32, C.c_int32, C.c_int32],
}


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().cmpc_last_error().decode(errors="replace")
        raise CmpcError(f"{what} failed with status {rc}: {msg}")


def exported_symbols():
    """Every entry point include/cmpc_b200.h declares (used by the CPU-side ABI test)."""
    return ["cmpc_last_error", "cmpc_version", "cmpc_gemm_set_mode", "cmpc_graph_set_mode", "cmpc_ln_relu_l2norm_set_mode", "cmpc_set_pdl"] + [n for n in dir(_Sigs) if n.startswith("cmpc_")] + list(_SIZE_FNS)
