"""``torch.ops.cmpc.*``: the C-ABI kernel families of libcmpc_b200 as PyTorch custom ops ("C-ABI the replacement
exports" / "Callers"; north star: "a thin C-ABI layer exposed as PyTorch custom ops").

The reference has no such layer -- every box below is a handful of TF nodes inside ``build_graph`` -- so each op cites the lines of
CMPC_model.py it replaces.  An op is the C entry point plus output allocation: tensors in, tensors out, launched on the caller's
current CUDA stream, no hidden state.  They are registered for CUDA only: a CPU tensor (or a box without the sm_100a library) raises,
there is no fallback.  ``CMPCHeadB200`` chains the same entry points over persistent buffers (head.py); these ops are what a host
that owns its own tensors (``torch.compile``-d glue, a serving wrapper, a TF<->torch bridge) calls one at a time.

    import cmpc_refseg_b200.ops            # registers the library
    y, stats = torch.ops.cmpc.graph_reason(w16, v16, x16, 1000, 2048.0)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L

_HEADS: Dict[int, object] = {}


def _st(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _need(t: torch.Tensor, dtype, name: str, dim: Optional[int] = None):
    if not t.is_cuda:
        raise L.CmpcError(f"cmpc::{name}: tensors must live on a CUDA device (no CPU path)")
    if t.dtype != dtype:
        raise L.CmpcError(f"cmpc::{name}: expected {dtype}, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise L.CmpcError(f"cmpc::{name}: expected a {dim}-d tensor, got shape {tuple(t.shape)}")
    if t.stride(-1) != 1:
        raise L.CmpcError(f"cmpc::{name}: the last dimension must be contiguous")
    return t


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------------------------------------
# _conv with filter_size 1 (+ the elementwise nodes after each call site): CMPC_model.py:412-417
# ------------------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("cmpc::gemm_bias_act", mutates_args=(), device_types="cuda")
def gemm_bias_act(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], sbias: Optional[torch.Tensor],
                  gate: Optional[torch.Tensor], act: int, rows_per_sample: int, out_fp32: bool) -> torch.Tensor:
    """out[m, n] = act(sum_k a[m, k] w[n, k] + bias[n] + sbias[b(m), n]) * gate[b(m), n]; a fp16 [M, K], w fp16 [N, Kpad]
    (Kpad = K rounded up to 64, zero padded), act 0 none / 1 relu / 2 tanh / 3 sigmoid, b(m) = m // rows_per_sample.
    Returns fp16 (or fp32) [M, ceil8(N)]; columns >= N are zero."""
    _need(a, torch.float16, "gemm_bias_act", 2); _need(w, torch.float16, "gemm_bias_act", 2)
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] < _rup(K, 64) or a.stride(0) % 8 or w.stride(0) % 8:
        raise L.CmpcError("cmpc::gemm_bias_act: w must be [N, >= 64*ceil(K/64)] and row strides multiples of 8 elements")
    ldo = _rup(N, 8)
    out = torch.zeros(M, ldo, dtype=torch.float32 if out_fp32 else torch.float16, device=a.device)
    g = L.GemmArgs()
    g.a1, g.lda1, g.k1 = a.data_ptr(), a.stride(0), K
    g.w, g.ldw = w.data_ptr(), w.stride(0)
    g.m, g.n, g.rows_per_sample = M, N, max(int(rows_per_sample), 1)
    nb = (M + g.rows_per_sample - 1) // g.rows_per_sample
    if bias is not None:
        if _need(bias, torch.float32, "gemm_bias_act", 1).numel() < _rup(N, 4):
            raise L.CmpcError("cmpc::gemm_bias_act: bias must hold ceil4(N) floats")
        g.bias = bias.data_ptr()
    for name, t in (("sbias", sbias), ("gate", gate)):
        if t is not None:
            _need(t, torch.float32, "gemm_bias_act", 2)
            if t.shape[0] != nb or t.shape[1] < _rup(N, 4):
                raise L.CmpcError(f"cmpc::gemm_bias_act: {name} must be [{nb}, >= ceil4(N)]")
            setattr(g, name, t.data_ptr()); setattr(g, "ld_" + name, t.stride(0))
    g.act = int(act)
    g.out, g.ldo, g.out_fp32 = out.data_ptr(), ldo, int(out_fp32)
    L.check(L.lib().cmpc_gemm_f16(C.byref(g), _st(a)), "cmpc_gemm_f16")
    return out


@gemm_bias_act.register_fake
def _(a, w, bias, sbias, gate, act, rows_per_sample, out_fp32):
    return a.new_empty(a.shape[0], _rup(w.shape[0], 8), dtype=torch.float32 if out_fp32 else torch.float16)


# ------------------------------------------------------------------------------------------------------------------------------
# mutan_head x5 + mutan_fusion (entity perception): CMPC_model.py:295-328
# ------------------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("cmpc::mutan_fusion", mutates_args=(), device_types="cuda")
def mutan_fusion(a: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, lang: torch.Tensor, channels: int,
                 rows_per_sample: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """a fp16 [M, ld] = [visual (l2-normalised) | spatial(8) | 0...]; w_packed from weights.pack_mutan_weights; bias fp32 [5, ld_b];
    lang fp32 [B, 5, ld_l] = tanh(lang_trans).  Returns (tanh(sum_k tanh(.)*lang) fp32 [M, ceil8(C)], row sums of squares [M]) --
    the l2_normalize of :324 is rsqrt(max(row_sumsq, 1e-12))."""
    _need(a, torch.float16, "mutan_fusion", 2); _need(w_packed, torch.float16, "mutan_fusion", 2)
    _need(bias, torch.float32, "mutan_fusion", 2); _need(lang, torch.float32, "mutan_fusion", 3)
    M = a.shape[0]
    ldo = _rup(channels, 8)
    out = torch.zeros(M, ldo, dtype=torch.float32, device=a.device)
    rss = torch.zeros(M, dtype=torch.float32, device=a.device)
    ma = L.MutanArgs()
    ma.a, ma.lda, ma.k = a.data_ptr(), a.stride(0), channels + 8
    ma.w, ma.ldw = w_packed.data_ptr(), w_packed.stride(0)
    ma.m, ma.c, ma.rows_per_sample = M, channels, rows_per_sample
    ma.bias, ma.ld_bias = bias.data_ptr(), bias.stride(0)
    ma.lang, ma.ld_lang, ma.lang_batch_stride = lang.data_ptr(), lang.stride(1), lang.stride(0)
    ma.out, ma.ldo, ma.row_sumsq = out.data_ptr(), ldo, rss.data_ptr()
    L.check(L.lib().cmpc_mutan_f16(C.byref(ma), _st(a)), "cmpc_mutan_f16")
    return out, rss


@mutan_fusion.register_fake
def _(a, w_packed, bias, lang, channels, rows_per_sample):
    return a.new_empty(a.shape[0], _rup(channels, 8), dtype=torch.float32), a.new_empty(a.shape[0], dtype=torch.float32)


# ------------------------------------------------------------------------------------------------------------------------------
# build_spa_graph: the two affinity softmaxes (:388-399) and adj . X without the adjacency (:400 + :362)
# ------------------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("cmpc::affinity_softmax", mutates_args=(), device_types="cuda")
def affinity_softmax(affi: torch.Tensor, seq_mask: torch.Tensor, rows_per_sample: int,
                     v_scale: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """affi fp32 [B*N, 32] (columns >= T zero, already times R_t / sqrt(C)), seq_mask fp32 [B, T] ->
    (W fp16 [B*N, 32], V * v_scale fp16 [B*N, 32], gw_w fp32 [B*N, T], gw_v fp32 [B*N, T])."""
    _need(affi, torch.float32, "affinity_softmax", 2); _need(seq_mask, torch.float32, "affinity_softmax", 2)
    if affi.shape[1] != 32 or not affi.is_contiguous() or not seq_mask.is_contiguous():
        raise L.CmpcError("cmpc::affinity_softmax: affi must be contiguous [B*N, 32], seq_mask contiguous [B, T]")
    B, T = seq_mask.shape
    lib = L.lib()
    w16 = torch.zeros(affi.shape[0], 32, dtype=torch.float16, device=affi.device)
    v16 = torch.zeros_like(w16)
    gw_w = torch.zeros(affi.shape[0], T, dtype=torch.float32, device=affi.device)
    gw_v = torch.zeros_like(gw_w)
    nws = int(lib.cmpc_affinity_workspace_bytes(B))
    ws = torch.zeros(max(nws, 8), dtype=torch.uint8, device=affi.device)
    L.check(lib.cmpc_affinity_softmax(affi.data_ptr(), seq_mask.data_ptr(), B, rows_per_sample, T, float(v_scale), w16.data_ptr(),
                                      v16.data_ptr(), gw_w.data_ptr(), gw_v.data_ptr(), ws.data_ptr(), nws, _st(affi)), "cmpc_affinity_softmax")
    return w16, v16, gw_w, gw_v


@affinity_softmax.register_fake
def _(affi, seq_mask, rows_per_sample, v_scale):
    T = seq_mask.shape[1]
    h = affi.new_empty(affi.shape[0], 32, dtype=torch.float16)
    f = affi.new_empty(affi.shape[0], T)
    return h, torch.empty_like(h), f, torch.empty_like(f)


@torch.library.custom_op("cmpc::graph_reason", mutates_args=(), device_types="cuda")
def graph_reason(w16: torch.Tensor, v16: torch.Tensor, x16: torch.Tensor, channels: int, batch: int,
                 v_scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Y = (W V^T) X / v_scale per sample, flash-style on tcgen05 (the N x N adjacency stays in TMEM); w16 / v16 fp16 [B*N, 32],
    x16 fp16 [B*N, ld].  Returns (Y fp16 [B*N, ld], per-sample (sum, sum of squares) of Y as fp64 [B, 2] for the layer norm :364)."""
    for t in (w16, v16, x16):
        _need(t, torch.float16, "graph_reason", 2)
    M, ld = x16.shape[0], x16.stride(0)
    if M % batch or w16.shape != (M, 32) or v16.shape != (M, 32) or not (w16.is_contiguous() and v16.is_contiguous()):
        raise L.CmpcError("cmpc::graph_reason: w16 / v16 must be contiguous [B*N, 32] with B*N the rows of x16")
    y = torch.zeros(M, ld, dtype=torch.float16, device=x16.device)
    stats = torch.zeros(batch, 2, dtype=torch.float64, device=x16.device)
    L.check(L.lib().cmpc_graph_reason_f16(w16.data_ptr(), v16.data_ptr(), x16.data_ptr(), ld, batch, M // batch, channels, float(v_scale),
                                          y.data_ptr(), ld, stats.data_ptr(), None, _st(x16)), "cmpc_graph_reason_f16")
    return y[:, :x16.shape[1]], stats


@graph_reason.register_fake
def _(w16, v16, x16, channels, batch, v_scale):
    return torch.empty_like(x16), x16.new_empty(batch, 2, dtype=torch.float64)


# ------------------------------------------------------------------------------------------------------------------------------
# gated_exchange_module's sum + the l2_normalize of :272-284
# ------------------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("cmpc::exchange_add_norm", mutates_args=(), device_types="cuda")
def exchange_add_norm(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, width: int, normalize: bool) -> torch.Tensor:
    """l2_normalize_C(a + b + c) (normalize) or the plain sum (:258); fp16 [rows, ld] each, same ld, pads zero."""
    for t in (a, b, c):
        _need(t, torch.float16, "exchange_add_norm", 2)
    if not (a.shape == b.shape == c.shape and a.stride(0) == b.stride(0) == c.stride(0)):
        raise L.CmpcError("cmpc::exchange_add_norm: the three maps must share shape and row stride")
    out = torch.zeros(a.shape[0], a.stride(0), dtype=torch.float16, device=a.device)
    L.check(L.lib().cmpc_add3_l2norm_f16(a.data_ptr(), b.data_ptr(), c.data_ptr(), a.stride(0), out.data_ptr(), a.stride(0), a.shape[0],
                                         width, int(normalize), None, _st(a)), "cmpc_add3_l2norm_f16")
    return out[:, :a.shape[1]]


@exchange_add_norm.register_fake
def _(a, b, c, width, normalize):
    return torch.empty_like(a)


# ------------------------------------------------------------------------------------------------------------------------------
# score conv + legacy bilinear upsampling + sigmoid (:138-142), losses (:439-445), mIoU counts (:486-489)
# ------------------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("cmpc::score_upsample", mutates_args=(), device_types="cuda")
def score_upsample(feat16: torch.Tensor, w9: torch.Tensor, bias: float, batch: int, h: int, w: int, width: int, out_h: int,
                   out_w: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """feat16 fp16 [B*h*w, ld]; w9 fp32 [9, ld9] tap-major (DW[dy, dx, :, 0]) -> (pred [B,h,w,1], up [B,H,W,1], sigm [B,H,W,1])."""
    _need(feat16, torch.float16, "score_upsample", 2); _need(w9, torch.float32, "score_upsample", 2)
    if feat16.shape[0] != batch * h * w or w9.shape[0] != 9 or w9.stride(0) != feat16.stride(0):
        raise L.CmpcError("cmpc::score_upsample: feat16 must be [B*h*w, ld] and w9 [9, ld] with the same ld")
    dev = feat16.device
    lib = L.lib()
    pred = torch.zeros(batch, h, w, 1, dtype=torch.float32, device=dev)
    up = torch.zeros(batch, out_h, out_w, 1, dtype=torch.float32, device=dev)
    sigm = torch.zeros_like(up)
    nws = int(lib.cmpc_score_workspace_bytes(batch * h * w))
    ws = torch.zeros(max(nws, 8), dtype=torch.uint8, device=dev)
    L.check(lib.cmpc_score_upsample(feat16.data_ptr(), feat16.stride(0), w9.data_ptr(), float(bias), batch, h, w, width, out_h, out_w,
                                    pred.data_ptr(), up.data_ptr(), sigm.data_ptr(), ws.data_ptr(), nws, _st(feat16)), "cmpc_score_upsample")
    return pred, up, sigm


@score_upsample.register_fake
def _(feat16, w9, bias, batch, h, w, width, out_h, out_w):
    f = feat16.new_empty(batch, out_h, out_w, 1, dtype=torch.float32)
    return feat16.new_empty(batch, h, w, 1, dtype=torch.float32), f, torch.empty_like(f)


@torch.library.custom_op("cmpc::ce_loss", mutates_args=(), device_types="cuda")
def ce_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """per-sample sum over pixels of tf.nn.sigmoid_cross_entropy_with_logits (util/loss.py:6-16): fp64 [B]"""
    _need(logits, torch.float32, "ce_loss"); _need(target, torch.float32, "ce_loss")
    if logits.shape != target.shape or not (logits.is_contiguous() and target.is_contiguous()):
        raise L.CmpcError("cmpc::ce_loss: logits and target must be contiguous and of one shape")
    B = logits.shape[0]
    sums = torch.zeros(B, dtype=torch.float64, device=logits.device)
    L.check(L.lib().cmpc_sigmoid_ce_sums(logits.data_ptr(), target.data_ptr(), B, logits.numel() // B, sums.data_ptr(), _st(logits)),
            "cmpc_sigmoid_ce_sums")
    return sums


@ce_loss.register_fake
def _(logits, target):
    return logits.new_empty(logits.shape[0], dtype=torch.float64)


@torch.library.custom_op("cmpc::iou_counts", mutates_args=(), device_types="cuda")
def iou_counts(up: torch.Tensor, target: torch.Tensor, thresh: float, inclusive: bool) -> torch.Tensor:
    """int64 [B, 2] = (|pred & gt|, |pred | gt|), pred = up > thresh (>= if inclusive, trainval_model.py:244), gt = target != 0"""
    _need(up, torch.float32, "iou_counts"); _need(target, torch.float32, "iou_counts")
    if up.shape != target.shape or not (up.is_contiguous() and target.is_contiguous()):
        raise L.CmpcError("cmpc::iou_counts: up and target must be contiguous and of one shape")
    B = up.shape[0]
    iu = torch.zeros(B, 2, dtype=torch.int64, device=up.device)
    L.check(L.lib().cmpc_iou_counts(up.data_ptr(), target.data_ptr(), B, up.numel() // B, float(thresh), int(inclusive), iu.data_ptr(),
                                    _st(up)), "cmpc_iou_counts")
    return iu


@iou_counts.register_fake
def _(up, target, thresh, inclusive):
    return up.new_empty(up.shape[0], 2, dtype=torch.int64)


# ------------------------------------------------------------------------------------------------------------------------------
# build_graph (:89-142) as one op over a registered head
# ------------------------------------------------------------------------------------------------------------------------------
def register_head(head) -> int:
    """Makes a CMPCHeadB200 (or the drop-in LSTM_model) callable as torch.ops.cmpc.head_forward(handle, ...)"""
    head = getattr(head, "_head", head)
    handle = len(_HEADS) + 1
    _HEADS[handle] = head
    return handle


@torch.library.custom_op("cmpc::head_forward", mutates_args=(), device_types="cuda")
def head_forward(handle: int, c3: torch.Tensor, c4: torch.Tensor, c5: torch.Tensor,
                 lstm_outputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(pred [B,h,w,1], up [B,H,W,1], sigm [B,H,W,1]) of the whole head; the outputs are copies (the head reuses its buffers)."""
    head = _HEADS.get(int(handle))
    if head is None:
        raise L.CmpcError(f"cmpc::head_forward: unknown head handle {handle} (ops.register_head)")
    out = head.forward(c3, c4, c5, lstm_outputs)
    return out["pred"].clone(), out["up"].clone(), out["sigm"].clone()


@head_forward.register_fake
def _(handle, c3, c4, c5, lstm_outputs):
    head = _HEADS[int(handle)]
    B, d = c3.shape[0], head.d
    f = c3.new_empty(B, d.H, d.W, 1, dtype=torch.float32)
    return c3.new_empty(B, d.h, d.w, 1, dtype=torch.float32), f, torch.empty_like(f)


OPS = ("gemm_bias_act", "mutan_fusion", "affinity_softmax", "graph_reason", "exchange_add_norm", "score_upsample", "ce_loss",
       "iou_counts", "head_forward")
