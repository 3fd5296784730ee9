"""CPU oracle for the CMPC head -- TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``cmpc_refseg_b200``) never does and has no CPU
fallback.

PARITY PINNED ON WIRING, RESTATED ON OP ARITHMETIC.  The reference (zigonk/CMPC-Refseg) ships no golden vectors, no
known-answer tests and no seeds, and TensorFlow 1.x cannot be installed in this image.  What CAN run here is the
reference's own source: ``oracle/ref_runner.py`` imports ``/root/reference/CMPC_model.py`` (+ ``util/cell.py``,
``util/loss.py``, ``util/processing_tools.py``) UNMODIFIED and executes ``LSTM_model.__init__ -> build_graph -> train_op``
through an eager ``tensorflow`` stand-in (``oracle/tfshim``).  ``tests/test_reference_pin.py`` holds this file to that
execution: every public output (pred, up, sigm, up_c3/4/5, words_parse, gw_w, gw_v, seq_mask) to <= 1e-10 in float64
(<= 1e-5 against the committed float32 fixtures ``tests/golden/ref_*.npz``), the five losses to 1e-9 relative, all 212
parameter gradients of ``compute_gradients`` to <= 1e-6 relative, the word-LSTM front and its gradients, variable
names / shapes / initialisers and the 67 218 008 parameter count.  So the WIRING (which tensor feeds which op, scopes,
masks, reshapes, gate order, loss weights, bias-gradient doubling) is the reference's.  What remains a restatement is
the arithmetic inside each ``tf.*`` op (the stand-in follows TF-1's published definitions, listed in its docstring);
those semantics are additionally checked against hand-computed values in ``tests/test_oracle_semantics.py``.

What it restates (all paths relative to the reference checkout):
  CMPC_model.py:106-142   build_graph (forward of the head)
  CMPC_model.py:159-163   l2-normalised word features + seq_mask
  CMPC_model.py:166-417   valid_lang, nec_lang, lang_se, global_vec, gated_exchange_*,
                          mutan_head/fusion, build_lang2vis, build_lang_parser, graph_conv,
                          build_spa_graph, _conv
  CMPC_model.py:439-447   4-term sigmoid-CE loss + L2 regulariser
  CMPC_model.py:486-490   in-graph mIoU
  util/cell.py:36-79      ConvLSTMCell.call (1x1 kernel, peepholes, whole-map layer norm)
  util/processing_tools.py:5-17   generate_spatial_batch
  util/loss.py:6-16,28-32 weighed_logistic_loss, l2_regularization_loss
  util/eval_tools.py:31-35 + trainval_model.py:267-294   mask I/U and running IoU statistics

Arithmetic lives in TensorFlow 1.13-1.15 (not vendored, not installable here); the op
semantics restated below are TF-1's: NHWC/HWIO cross-correlation conv with SAME padding,
``l2_normalize`` = x * rsqrt(max(sum(x^2), 1e-12)), ``tf.contrib.layers.layer_norm`` with
begin_norm_axis=1 (statistics over ALL non-batch axes, biased variance, eps 1e-12,
per-channel gamma/beta), legacy ``resize_bilinear`` (align_corners=False, NO half-pixel
centres), max-subtracted softmax, ``sigmoid_cross_entropy_with_logits`` =
max(x,0) - x*z + log1p(exp(-|x|)).

Tensors are channels-last ``[B, h, w, C]`` exactly like the reference.  Parameters are a
flat dict keyed by the TF variable names below scope ``text_objseg/``
(e.g. ``"c5_lateral/DW"`` with HWIO shape ``[1,1,2048,1000]``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

FLT_MIN_TF = float(np.finfo(np.float32).min)  # tf.float32.min = -3.4028235e38  (CMPC_model.py:390)


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class HeadConfig:
    """Hyper-parameters the head reads from the ctor (CMPC_model.py:15-40)."""
    batch_size: int = 1
    num_steps: int = 20
    vf_h: int = 40
    vf_w: int = 40
    H: int = 320
    W: int = 320
    vf_dim: int = 2048          # c5 channels (c4 / c3 are hard-coded 1024 / 512, :110,:112)
    c4_dim: int = 1024
    c3_dim: int = 512
    v_emb_dim: int = 1000
    rnn_size: int = 1000
    mlp_dim: int = 500
    weight_decay: float = 0.0005
    parse_hidden: int = 500     # hard-coded 500 at :349 (independent of mlp_dim)

    @property
    def n_nodes(self) -> int:
        return self.vf_h * self.vf_w


# --------------------------------------------------------------------------------------
# TF-1 op semantics
# --------------------------------------------------------------------------------------
def l2_normalize(x: torch.Tensor, axis=None, epsilon: float = 1e-12) -> torch.Tensor:
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x**2, axis, keepdims), eps)).
    axis=None sums over EVERY axis including batch (CMPC_model.py:241)."""
    sq = x * x
    if axis is None:
        s = sq.sum()
    else:
        s = sq.sum(dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(s, min=epsilon))


def layer_norm_tf(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """tf.contrib.layers.layer_norm defaults: moments over axes [1..rank-1] (biased variance),
    variance_epsilon 1e-12, gamma/beta over the last axis (CMPC_model.py:364,370; util/cell.py:53-66)."""
    axes = tuple(range(1, x.dim()))
    mean = x.mean(dim=axes, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=axes, keepdim=True)
    return (x - mean) * torch.rsqrt(var + 1e-12) * gamma + beta


def conv2d_same(x: torch.Tensor, w_hwio: torch.Tensor, b: Optional[torch.Tensor],
                mm: Optional[Callable] = None) -> torch.Tensor:
    """tf.nn.conv2d(x, w, [1,1,1,1], 'SAME') + b on NHWC x HWIO (cross-correlation)  (:412-417).
    ``mm`` optionally replaces the matmul of 1x1 convs (used only by the precision study)."""
    kh, kw, cin, cout = w_hwio.shape
    if kh == 1 and kw == 1:
        w2 = w_hwio.reshape(cin, cout)
        y = (mm(x.reshape(-1, cin), w2) if mm is not None else x.reshape(-1, cin) @ w2)
        y = y.reshape(*x.shape[:-1], cout)
    else:
        assert kh == kw and kh % 2 == 1
        y = F.conv2d(x.permute(0, 3, 1, 2).contiguous(), w_hwio.permute(3, 2, 0, 1).contiguous(), padding=kh // 2)
        y = y.permute(0, 2, 3, 1)
    if b is not None:
        y = y + b
    return y


def resize_bilinear_legacy(x: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """tf.image.resize_bilinear (TF-1 legacy, align_corners=False, no half-pixel centres):
    src = dst * (in/out); lo = floor(src); hi = min(lo+1, in-1); separable lerp.   x: [B,h,w,C]."""
    B, h, w, C = x.shape

    def axis_weights(n_in, n_out):
        scale = n_in / n_out
        src = torch.arange(n_out, dtype=torch.float64) * scale
        lo = torch.floor(src).to(torch.int64)
        hi = torch.clamp(lo + 1, max=n_in - 1)
        frac = (src - lo.to(torch.float64)).to(x.dtype)
        return lo, hi, frac

    ylo, yhi, yf = axis_weights(h, out_h)
    xlo, xhi, xf = axis_weights(w, out_w)
    top = x[:, ylo]          # [B,out_h,w,C]
    bot = x[:, yhi]
    yf_ = yf.view(1, -1, 1, 1)
    xf_ = xf.view(1, 1, -1, 1)
    tl, tr = top[:, :, xlo], top[:, :, xhi]
    bl, br = bot[:, :, xlo], bot[:, :, xhi]
    t = tl + (tr - tl) * xf_
    b = bl + (br - bl) * xf_
    return t + (b - t) * yf_


def sigmoid_ce_with_logits(x: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|))."""
    return torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x)))


def generate_spatial_batch(n: int, fh: int, fw: int) -> np.ndarray:
    """util/processing_tools.py:5-17 (true division): [n, fh, fw, 8] float32 =
    (xmin, ymin, xmax, ymax, xctr, yctr, 1/fw, 1/fh), coordinates in [-1, 1]."""
    out = np.zeros((n, fh, fw, 8), dtype=np.float32)
    for h in range(fh):
        for w in range(fw):
            xmin = w / fw * 2 - 1
            xmax = (w + 1) / fw * 2 - 1
            xctr = (xmin + xmax) / 2
            ymin = h / fh * 2 - 1
            ymax = (h + 1) / fh * 2 - 1
            yctr = (ymin + ymax) / 2
            out[:, h, w, :] = [xmin, ymin, xmax, ymax, xctr, yctr, 1 / fw, 1 / fh]
    return out


# --------------------------------------------------------------------------------------
# parameters and synthetic inputs
# --------------------------------------------------------------------------------------
LEVELS = ("c5", "c4", "c3")                      # build order (:120-125)
EXG = ("c3", "c4", "c5", "c3_2", "c4_2", "c5_2")  # exchange modules (:271-283)


def param_shapes(cfg: HeadConfig) -> Dict[str, tuple]:
    """Every head variable under text_objseg/ with its TF shape."""
    C, R, M = cfg.v_emb_dim, cfg.rnn_size, cfg.mlp_dim
    s: Dict[str, tuple] = {}

    def conv(name, k, cin, cout):
        s[name + "/DW"] = (k, k, cin, cout)
        s[name + "/biases"] = (cout,)

    conv("c5_lateral", 1, cfg.vf_dim, C)
    conv("c4_lateral", 1, cfg.c4_dim, C)
    conv("c3_lateral", 1, cfg.c3_dim, C)
    conv("words_parse_1", 1, R, cfg.parse_hidden)
    conv("words_parse_2", 1, cfg.parse_hidden, 4)
    for lvl in LEVELS:
        for k in range(1, 6):
            conv(f"vis_trans_{lvl}_head{k}", 1, C + 8, C)
            conv(f"lang_trans_{lvl}_head{k}", 1, R, C)
        conv(f"words_trans_{lvl}", 1, R, R)
        conv(f"spa_graph_trans2_{lvl}", 1, C, C)
        conv(f"gconv_update_spa_graph_{lvl}", 1, C, C)
        for ln in ("gconv_feat_ln_spa_graph", "gconv_update_ln_spa_graph"):
            s[f"{ln}_{lvl}/beta"] = (C,)
            s[f"{ln}_{lvl}/gamma"] = (C,)
        conv(f"fusion_{lvl}", 1, 2 * C + R + 8, M)
        conv(f"score_{lvl}", 3, M, 1)
    conv("score", 3, M, 1)
    for x in EXG:
        conv(f"spa_graph_key_{x}gv_f1", 1, M, M)
        conv(f"lang_query_{x}gv_f1", 1, R, M)
        conv(f"gv_lang_{x}gv_f1", 1, M + R, M)
        for f in ("_f1", "_f2"):
            conv(f"lang_feat_{x}{f}", 1, M, M)
            conv(f"trans_feat_{x}{f}", 1, M, M)
    s["rnn/conv_lstm_cell/kernel"] = (1, 1, 2 * M, 4 * M)
    for p in ("W_ci", "W_cf", "W_co"):
        s[f"rnn/conv_lstm_cell/{p}"] = (cfg.vf_h, cfg.vf_w, M)
    for i in range(5):
        nm = "LayerNorm" if i == 0 else f"LayerNorm_{i}"
        s[f"rnn/conv_lstm_cell/{nm}/beta"] = (M,)
        s[f"rnn/conv_lstm_cell/{nm}/gamma"] = (M,)
    return s


def _glorot_limit(shape) -> float:
    """TF glorot/xavier uniform limit sqrt(6/(fan_in+fan_out)); fans as TF computes them
    (receptive field = prod(shape[:-2]))."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    return math.sqrt(6.0 / (fan_in + fan_out))


def init_params(cfg: HeadConfig, seed: int = 0, dtype=torch.float32, *, sharp: float = 1.0,
                bias_std: float = 0.0, ln_jitter: float = 0.0) -> Dict[str, torch.Tensor]:
    """Reference initialisers: xavier-uniform DW, zero biases (:414-416), glorot-uniform
    ConvLSTM kernel/peepholes (TF default), LN gamma=1 beta=0.
    ``sharp`` multiplies words_trans_* / spa_graph_trans2_* DW so the affinity logits are O(1)
    (SURVEY App. D-7); ``bias_std`` / ``ln_jitter`` perturb biases and LN gamma/beta so tests
    exercise those code paths (they are identically 0/1 at reference init)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "biases":
            t = torch.randn(shape, generator=g, dtype=torch.float64) * bias_std
        elif leaf == "beta":
            t = torch.randn(shape, generator=g, dtype=torch.float64) * ln_jitter
        elif leaf == "gamma":
            t = 1.0 + torch.randn(shape, generator=g, dtype=torch.float64) * ln_jitter
        else:
            lim = _glorot_limit(shape)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
            if sharp != 1.0 and (name.startswith("words_trans_") or name.startswith("spa_graph_trans2_")):
                t = t * sharp
        out[name] = t.to(dtype)
    return out


def make_inputs(cfg: HeadConfig, batch: int, seed: int = 1234, seq_len=None, dtype=torch.float32):
    """Synthetic UNC-shaped inputs: c3/c4/c5 = relu(N(0,1)) (taps are post-ReLU),
    lstm_outputs = tanh(N(0,1))*sigmoid(N(0,1)) zeroed for t >= seq_len, rectangular target mask."""
    g = torch.Generator().manual_seed(seed)
    h, w, T, R = cfg.vf_h, cfg.vf_w, cfg.num_steps, cfg.rnn_size
    c3 = torch.relu(torch.randn(batch, h, w, cfg.c3_dim, generator=g)).to(dtype)
    c4 = torch.relu(torch.randn(batch, h, w, cfg.c4_dim, generator=g)).to(dtype)
    c5 = torch.relu(torch.randn(batch, h, w, cfg.vf_dim, generator=g)).to(dtype)
    lstm = (torch.tanh(torch.randn(batch, T, R, generator=g)) *
            torch.sigmoid(torch.randn(batch, T, R, generator=g))).to(dtype)
    if seq_len is None:
        sl = torch.full((batch,), T, dtype=torch.int32)
    elif isinstance(seq_len, str) and seq_len == "unc":
        lam = torch.full((batch,), 3.5)
        sl = torch.clamp(torch.poisson(lam, generator=g), 1, T).to(torch.int32)
    else:
        sl = torch.as_tensor(seq_len, dtype=torch.int32).reshape(-1).expand(batch).clone()
    tmask = (torch.arange(T).view(1, T) < sl.view(-1, 1)).to(dtype).unsqueeze(-1)
    lstm = lstm * tmask
    target = torch.zeros(batch, cfg.H, cfg.W, 1, dtype=dtype)
    for b in range(batch):
        y0 = int(torch.randint(0, cfg.H // 2, (1,), generator=g))
        x0 = int(torch.randint(0, cfg.W // 2, (1,), generator=g))
        target[b, y0:y0 + cfg.H // 3 + 1, x0:x0 + cfg.W // 3 + 1, 0] = 1.0
    return dict(c3=c3, c4=c4, c5=c5, lstm_outputs=lstm, seq_len=sl, target_fine=target)


# --------------------------------------------------------------------------------------
# the head
# --------------------------------------------------------------------------------------
class OracleHead:
    """Literal restatement of LSTM_model.build_graph after the backbone taps and the word LSTM.

    gv_norm: 'sample' normalises gv_lang per sample (== the reference run at B=1, the way its own
             inference drivers run it); 'batch' is the literal axis=None graph at B>1 (:241).
    dense_adj: True materialises adj = W @ V^T and runs adj @ X like the reference (:400,:362);
             False uses the re-association W @ (V^T @ X) (an independent derivation, used by tests).
    mm: optional replacement for every 1x1-conv / batched matmul (precision study only).
    """

    def __init__(self, params: Dict[str, torch.Tensor], cfg: HeadConfig, *, gv_norm: str = "sample",
                 dense_adj: bool = True, mm: Optional[Callable] = None, keep: bool = False):
        assert gv_norm in ("sample", "batch")
        self.p, self.cfg, self.gv_norm, self.dense_adj, self.mm, self.keep = params, cfg, gv_norm, dense_adj, mm, keep
        self.t: Dict[str, torch.Tensor] = {}     # named intermediates (when keep=True)

    # -- helpers ------------------------------------------------------------------------
    def _conv(self, name, x):                                         # CMPC_model.py:412-417
        return conv2d_same(x, self.p[name + "/DW"], self.p[name + "/biases"], self.mm)

    def _matmul(self, a, b):
        if self.mm is None:
            return a @ b
        if a.dim() == 2:
            return self.mm(a, b)
        return torch.stack([self.mm(a[i], b[i]) for i in range(a.shape[0])])

    def _ln(self, scope, x):
        return layer_norm_tf(x, self.p[scope + "/gamma"], self.p[scope + "/beta"])

    def _save(self, name, v):
        if self.keep:
            self.t[name] = v
        return v

    # -- language side ------------------------------------------------------------------
    def words(self, lstm_outputs):                                    # :159-163
        wf = l2_normalize(lstm_outputs, -1).unsqueeze(1)               # [B,1,T,R]
        seq_mask = (wf.abs().sum(-1, keepdim=True) != 0).to(wf.dtype)  # [B,1,T,1]
        return wf, seq_mask

    def build_lang_parser(self, words_feat, seq_mask):                # :347-357
        x = torch.relu(self._conv("words_parse_1", words_feat))
        x = self._conv("words_parse_2", x)
        return torch.softmax(x, dim=3) * seq_mask                     # [B,1,T,4]  (E,A,R,U)

    def _weighted_lang(self, weights, words_feat):                    # :166-192 shared tail
        B, _, T, R = words_feat.shape
        v = weights @ words_feat.reshape(B, T, R)                      # [B,1,R]
        return l2_normalize(v, 2).reshape(B, 1, 1, R)

    def valid_lang(self, words_parse, words_feat):                    # :166-178  (E + A)
        return self._weighted_lang(words_parse[:, :, :, 0] + words_parse[:, :, :, 1], words_feat)

    def nec_lang(self, words_parse, words_feat):                      # :180-192  (E + A + R)
        return self._weighted_lang(words_parse.sum(3) - words_parse[:, :, :, 3], words_feat)

    # -- entity perception ----------------------------------------------------------------
    def mutan_head(self, lang_feat, spatial, visual, level):          # :295-309
        vis = torch.tanh(self._conv(f"vis_trans_{level}", torch.cat([visual, spatial], 3)))
        lang = torch.tanh(self._conv(f"lang_trans_{level}", lang_feat))
        return vis * lang

    def mutan_fusion(self, lang_feat, spatial, visual, level):        # :311-328
        heads = [self.mutan_head(lang_feat, spatial, visual, f"{level}_head{k}") for k in range(1, 6)]
        fused = torch.stack(heads, 4).sum(4)
        return l2_normalize(torch.tanh(fused), 3)

    # -- relation-aware reasoning -----------------------------------------------------------
    def graph_conv(self, graph_feat, adj_or_wv, level):               # :359-374
        B, _, N, C = graph_feat.shape
        X = graph_feat.reshape(B, N, C)
        if self.dense_adj:
            Y = self._matmul(adj_or_wv, X)                             # adj [B,N,N] @ X
        else:
            Wm, Vm = adj_or_wv
            Y = Wm @ (Vm.transpose(1, 2) @ X)
        Y = self._save(f"gconv_y_{level}", Y.reshape(B, 1, N, C))
        Y = self._ln(f"gconv_feat_ln_spa_graph_{level}", Y)
        Z = torch.relu(graph_feat + Y)
        U = self._conv(f"gconv_update_spa_graph_{level}", Z)
        U = self._ln(f"gconv_update_ln_spa_graph_{level}", U)
        return torch.relu(U)

    def build_spa_graph(self, spa_graph, words_feat, words_parse, seq_mask, level):   # :376-410
        cfg = self.cfg
        B, T, N = spa_graph.shape[0], cfg.num_steps, cfg.n_nodes
        wt = self._conv(f"words_trans_{level}", words_feat).reshape(B, T, cfg.rnn_size)
        xt = self._conv(f"spa_graph_trans2_{level}", spa_graph).reshape(B, N, cfg.v_emb_dim)
        affi = self._matmul(xt, wt.transpose(1, 2)) / (cfg.v_emb_dim ** 0.5)      # [B,N,T]
        affi = words_parse[:, :, :, 2] * affi                                      # relation weight R_t
        mask = seq_mask.reshape(B, 1, T)
        gw_w = torch.softmax(mask * affi + (1 - mask) * FLT_MIN_TF, dim=2)         # over words
        gw_v = mask * torch.softmax(affi, dim=1)                                   # over nodes, then mask
        self._save(f"affi_{level}", affi)
        self._save(f"gw_w_{level}", gw_w)
        self._save(f"gw_v_{level}", gw_v)
        adj = gw_w @ gw_v.transpose(1, 2) if self.dense_adj else (gw_w, gw_v)      # [B,N,N]
        g = self.graph_conv(spa_graph.reshape(B, 1, N, cfg.v_emb_dim), adj, level)
        g = g.reshape(B, cfg.vf_h, cfg.vf_w, cfg.v_emb_dim)
        return l2_normalize(g, 3), gw_w, gw_v

    def build_lang2vis(self, visual, words_feat, words_parse, seq_mask, spatial, level):  # :330-345
        cfg = self.cfg
        vl = self.valid_lang(words_parse, words_feat)
        vis_la_sp = self._save(f"vis_la_sp_{level}", self.mutan_fusion(vl, spatial, visual, level))
        spa, gw_w, gw_v = self.build_spa_graph(vis_la_sp, words_feat, words_parse, seq_mask, level)
        self._save(f"spa_graph_{level}", spa)
        tiled = vl.expand(-1, cfg.vf_h, cfg.vf_w, -1)
        feat_all = torch.cat([vis_la_sp, spa, tiled, spatial], 3)
        return torch.relu(self._conv(f"fusion_{level}", feat_all)), gw_w, gw_v

    # -- text-guided exchange ------------------------------------------------------------
    def global_vec(self, feat, lang_feat, level):                     # :212-243
        cfg = self.cfg
        B, N, M = feat.shape[0], cfg.n_nodes, cfg.mlp_dim
        key = self._conv(f"spa_graph_key_{level}", feat).reshape(B, N, M)
        q = self._conv(f"lang_query_{level}", lang_feat).reshape(B, 1, M)
        attn = torch.softmax((key @ q.transpose(1, 2)) / (M ** 0.5), dim=1)        # [B,N,1]
        pooled = (attn.transpose(1, 2) @ feat.reshape(B, N, M)).reshape(B, 1, 1, M)
        gv = self._conv(f"gv_lang_{level}", torch.cat([pooled, lang_feat], 3))
        if self.gv_norm == "batch":
            return l2_normalize(gv, None)                              # literal axis=None (:241)
        return l2_normalize(gv, (1, 2, 3))                             # == reference at B=1

    def lang_se(self, feat, gv, level):                               # :194-210
        gate = torch.sigmoid(self._conv(f"lang_feat_{level}", gv))
        return torch.relu(self._conv(f"trans_feat_{level}", feat)) * gate

    def gated_exchange_module(self, feat, feat1, feat2, lang_feat, level):   # :245-259
        gv = self.global_vec(feat, lang_feat, level + "gv_f1")
        return feat + self.lang_se(feat1, gv, level + "_f1") + self.lang_se(feat2, gv, level + "_f2")

    def conv_lstm(self, seq):                                         # util/cell.py:36-79 via dynamic_rnn
        p, M = self.p, self.cfg.mlp_dim
        pre = "rnn/conv_lstm_cell/"
        c = torch.zeros_like(seq[0])
        h = torch.zeros_like(seq[0])
        for step, x in enumerate(seq):
            y = conv2d_same(torch.cat([x, h], 3), p[pre + "kernel"], None, self.mm)   # no bias (normalize=True)
            j, i, f, o = torch.split(y, M, dim=3)
            i = i + p[pre + "W_ci"] * c
            f = f + p[pre + "W_cf"] * c
            j = self._ln(pre + "LayerNorm", j)
            i = self._ln(pre + "LayerNorm_1", i)
            f = self._ln(pre + "LayerNorm_2", f)
            f = torch.sigmoid(f + 1.0)                                  # forget_bias after LN (:57)
            i = torch.sigmoid(i)
            c = c * f + i * torch.tanh(j)
            o = o + p[pre + "W_co"] * c                                 # peephole on NEW pre-norm c
            o = self._ln(pre + "LayerNorm_3", o)
            c = self._ln(pre + "LayerNorm_4", c)                        # state stores the normed c
            o = torch.sigmoid(o)
            h = o * torch.tanh(c)
            self._save(f"convlstm_h{step}", h)
        return h

    def gated_exchange_fusion_lstm_2times(self, f3, f4, f5, lang):    # :261-293
        e3 = l2_normalize(self.gated_exchange_module(f3, f4, f5, lang, "c3"), 3)
        e4 = l2_normalize(self.gated_exchange_module(f4, f3, f5, lang, "c4"), 3)
        e5 = l2_normalize(self.gated_exchange_module(f5, f3, f4, lang, "c5"), 3)
        self._save("exg1_c3", e3); self._save("exg1_c4", e4); self._save("exg1_c5", e5)
        g3 = l2_normalize(self.gated_exchange_module(e3, e4, e5, lang, "c3_2"), 3)
        g4 = l2_normalize(self.gated_exchange_module(e4, e3, e5, lang, "c4_2"), 3)
        g5 = l2_normalize(self.gated_exchange_module(e5, e3, e4, lang, "c5_2"), 3)
        self._save("exg2_c3", g3); self._save("exg2_c4", g4); self._save("exg2_c5", g5)
        return self.conv_lstm([g3, g4, g5])

    # -- forward -------------------------------------------------------------------------
    def forward(self, c3, c4, c5, lstm_outputs, seq_len=None) -> Dict[str, torch.Tensor]:
        """seq_len is accepted for signature parity; like the reference the mask is derived from
        the (already zeroed) LSTM outputs (:163)."""
        cfg = self.cfg
        B = c5.shape[0]
        words_feat, seq_mask = self.words(lstm_outputs)
        v5 = l2_normalize(self._conv("c5_lateral", c5), 3)             # :108-113
        v4 = l2_normalize(self._conv("c4_lateral", c4), 3)
        v3 = l2_normalize(self._conv("c3_lateral", c3), 3)
        self._save("lateral_c5", v5); self._save("lateral_c4", v4); self._save("lateral_c3", v3)
        spatial = torch.from_numpy(generate_spatial_batch(B, cfg.vf_h, cfg.vf_w)).to(c5.dtype)
        words_parse = self.build_lang_parser(words_feat, seq_mask)
        f5, _, _ = self.build_lang2vis(v5, words_feat, words_parse, seq_mask, spatial, "c5")
        f4, _, _ = self.build_lang2vis(v4, words_feat, words_parse, seq_mask, spatial, "c4")
        f3, gw_w, gw_v = self.build_lang2vis(v3, words_feat, words_parse, seq_mask, spatial, "c3")
        self._save("fusion_c5", f5); self._save("fusion_c4", f4); self._save("fusion_c3", f3)
        out: Dict[str, torch.Tensor] = {}
        for lvl, f in (("c5", f5), ("c4", f4), ("c3", f3)):            # :128-133
            out["up_" + lvl] = resize_bilinear_legacy(self._conv("score_" + lvl, f), cfg.H, cfg.W)
        nec = self.nec_lang(words_parse, words_feat)                   # :135
        fused = self._save("fused", self.gated_exchange_fusion_lstm_2times(f3, f4, f5, nec))
        pred = self._conv("score", fused)                              # :138-142
        up = resize_bilinear_legacy(pred, cfg.H, cfg.W)
        out.update(pred=pred, up=up, sigm=torch.sigmoid(up), words_parse=words_parse,
                   seq_mask=seq_mask, gw_w=gw_w, gw_v=gw_v,
                   valid_lang=self.valid_lang(words_parse, words_feat), nec_lang=nec)
        return out

    # -- loss / metrics (training config) ----------------------------------------------------
    def losses(self, out, target_fine) -> Dict[str, torch.Tensor]:    # :439-447, util/loss.py
        def wll(scores):
            return sigmoid_ce_with_logits(scores, target_fine).sum(dim=(1, 2, 3)).mean()
        r = dict(cls_loss=wll(out["up"]), cls_loss_c5=wll(out["up_c5"]),
                 cls_loss_c4=wll(out["up_c4"]), cls_loss_c3=wll(out["up_c3"]))
        r["cls_loss_all"] = 0.7 * r["cls_loss"] + 0.1 * r["cls_loss_c5"] + 0.1 * r["cls_loss_c4"] + 0.1 * r["cls_loss_c3"]
        reg = sum((v.double() ** 2).sum() / 2 for k, v in self.p.items() if k.endswith("/DW"))
        r["reg_loss"] = (self.cfg.weight_decay * reg).to(out["up"].dtype)
        r["cost"] = r["cls_loss_all"] + r["reg_loss"]
        return r


def word_lstm(words: torch.Tensor, seq_len: torch.Tensor, embedding: torch.Tensor, kernel: torch.Tensor, bias: torch.Tensor,
              forget_bias: float = 1.0) -> torch.Tensor:
    """The word encoder in front of the head, LSTM_model.lstm() up to `outputs` (CMPC_model.py:144-157):
    embedding_lookup(glove, words) -> tf.nn.rnn_cell.LSTMCell(rnn_size, state_is_tuple=False) -> dynamic_rnn(sequence_length).
    The cell is third-party TensorFlow (not vendored; the code base needs TF 1.13-1.15): its published algorithm
    (rnn_cell_impl.LSTMCell.call without peepholes / projection) is restated here:
        [i, j, f, o] = split([x_t, m_{t-1}] kernel + bias, 4);  c_t = sigmoid(f + forget_bias) c_{t-1} + sigmoid(i) tanh(j);
        m_t = sigmoid(o) tanh(c_t)
    and dynamic_rnn emits zeros and carries the state through unchanged for t >= sequence_length.
    words int64 [B, T]; embedding [V, E]; kernel [E + R, 4R]; bias [4R]  ->  outputs [B, T, R]."""
    B, T = words.shape
    R = kernel.shape[1] // 4
    x = embedding[words]                                               # [B, T, E]
    c = torch.zeros(B, R, dtype=kernel.dtype)
    m = torch.zeros(B, R, dtype=kernel.dtype)
    outs = []
    for t in range(T):
        z = torch.cat([x[:, t], m], 1) @ kernel + bias
        i, j, f, o = torch.split(z, R, dim=1)
        c_new = torch.sigmoid(f + forget_bias) * c + torch.sigmoid(i) * torch.tanh(j)
        m_new = torch.sigmoid(o) * torch.tanh(c_new)
        active = (t < seq_len).to(kernel.dtype).unsqueeze(1)
        outs.append(active * m_new)
        c = active * c_new + (1 - active) * c
        m = active * m_new + (1 - active) * m
    return torch.stack(outs, 1)


def resize_and_crop_mask(pred_raw: np.ndarray, out_h: int, out_w: int, mode: str = "constant") -> np.ndarray:
    """The post-processing step after the head (trainval_model.py:244-245): `im_processing.resize_and_crop(pred_raw, gt_h, gt_w)`
    (util/im_processing.py:25-41) for a 2-D {0,1} float mask.  The resize is third-party scikit-image (`skimage.transform.resize`,
    not vendored; the code base is Python-2.7 era => skimage <= 0.14: order=1, mode=None -> 'constant' with cval 0, no
    anti-aliasing; mode='reflect' is what skimage >= 0.15 defaults to and is offered as a variant).  Its published algorithm,
    restated: output pixel (r, c) samples the input at ((r + 0.5) * in_h / res_h - 0.5, (c + 0.5) * in_w / res_w - 0.5) with
    bilinear weights between floor and ceil neighbours (skimage/_shared/interpolation.pxd: bilinear_interpolation), pixels
    outside the image reading cval = 0 ('constant') or their mirror image ('reflect').  PARITY UNPINNED: skimage is not
    installable here, and the real library derives the coordinates from a least-squares affine estimate whose last-ulp noise
    can turn an exactly-integer coordinate into a two-pixel blend; this restatement uses the exact coordinates.
    Returns float64 [out_h, out_w] (values in [0, 1]; compute_mask_IU treats any non-zero value as foreground)."""
    im_h, im_w = pred_raw.shape
    scale = max(out_h / im_h, out_w / im_w)
    res_h, res_w = int(np.round(im_h * scale)), int(np.round(im_w * scale))     # np.round: half to even, like the reference
    crop_h, crop_w = int(np.floor(res_h - out_h) / 2), int(np.floor(res_w - out_w) / 2)

    def axis(res, n_in, crop, n_out):
        pos = (np.arange(crop, crop + n_out, dtype=np.float64) + 0.5) * (n_in / res) - 0.5
        lo, hi = np.floor(pos).astype(np.int64), np.ceil(pos).astype(np.int64)
        return lo, hi, pos - lo

    def fetch(img, idx, axis_):
        n = img.shape[axis_]
        if mode == "reflect":                       # skimage 'reflect' == numpy 'symmetric': -1 -> 0, n -> n - 1
            j = np.where(idx < 0, -idx - 1, np.where(idx >= n, 2 * n - 1 - idx, idx))
            return np.take(img, np.clip(j, 0, n - 1), axis=axis_)
        ok = (idx >= 0) & (idx < n)
        out = np.take(img, np.clip(idx, 0, n - 1), axis=axis_)
        shape = [1, 1]; shape[axis_] = -1
        return out * ok.reshape(shape)

    rlo, rhi, dr = axis(res_h, im_h, crop_h, out_h)
    clo, chi, dc = axis(res_w, im_w, crop_w, out_w)
    img = pred_raw.astype(np.float64)
    top = (1 - dc)[None, :] * fetch(fetch(img, rlo, 0), clo, 1) + dc[None, :] * fetch(fetch(img, rlo, 0), chi, 1)
    bot = (1 - dc)[None, :] * fetch(fetch(img, rhi, 0), clo, 1) + dc[None, :] * fetch(fetch(img, rhi, 0), chi, 1)
    return (1 - dr)[:, None] * top + dr[:, None] * bot


def postprocess_iu(up_val: np.ndarray, gt_mask: np.ndarray, score_thresh: float = 1e-9, mode: str = "constant"):
    """One iteration of the test loop after sess.run (trainval_model.py:243-245, 266): threshold the upsampled logits, resize
    and crop to the ground-truth size, compute_mask_IU (util/eval_tools.py:31-35: logical and / or of non-zero entries)."""
    pred_raw = (np.squeeze(up_val) >= score_thresh).astype(np.float32)
    predicts = resize_and_crop_mask(pred_raw, gt_mask.shape[0], gt_mask.shape[1], mode)
    I = int(np.sum(np.logical_and(predicts, gt_mask)))
    U = int(np.sum(np.logical_or(predicts, gt_mask)))
    return predicts, I, U


def mask_iu(up: torch.Tensor, target_fine: torch.Tensor, thresh: float = 0.0, strict: bool = True):
    """Per-sample integer intersection / union of (up > 0) vs target (CMPC_model.py:486-489;
    util/eval_tools.py:31-35).  strict=False gives the host-driver variant up >= thresh."""
    pred = (up > thresh) if strict else (up >= thresh)
    lab = target_fine != 0
    I = (pred & lab).sum(dim=(1, 2, 3)).to(torch.int64)
    U = (pred | lab).sum(dim=(1, 2, 3)).to(torch.int64)
    return I, U


def iou_stats(I: torch.Tensor, U: torch.Tensor) -> Dict[str, float]:
    """Running statistics of trainval_model.py:267-294: cumulative I/U, mean IoU, precision@{.5..9}."""
    iou = I.double() / U.double()
    r = dict(cum_I=int(I.sum()), cum_U=int(U.sum()), n=int(I.numel()), sum_iou=float(iou.sum()))
    for k, th in enumerate((0.5, 0.6, 0.7, 0.8, 0.9)):
        r[f"prec@{th}"] = int((iou >= th).sum())
    r["overall_iou"] = r["cum_I"] / max(r["cum_U"], 1)
    r["mean_iou"] = r["sum_iou"] / max(r["n"], 1)
    return r
