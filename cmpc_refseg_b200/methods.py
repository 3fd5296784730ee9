"""The reference's per-method API of ``LSTM_model`` (CMPC_model.py:144-417), one call at a time, on the device.

Each method takes and returns tensors in the reference's own shapes (NHWC fp32 on the model's device), loads them into
the head's fp16 operand buffers and runs the SAME stage kernels ``CMPCHeadB200.forward`` chains (head.py ``_st_*``), so
a maintainer can swap any single ``self.<method>(...)`` of ``build_graph`` and compare it against the TF op by op.
Names, argument order and return values follow the reference; what differs is stated per method.

torch is used here to move arguments in and results out (copies and dtype casts); all arithmetic is in libcmpc_b200.
These calls are for inspection / op-level parity -- the fast path is ``forward`` (one fused pass, no staging copies).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .head import on_device

from . import _lib as L
from .weights import EXG, LEVELS, pack_conv1x1, rup


class ReferenceMethods:
    """Mixin for the drop-in ``LSTM_model`` (needs ``self._head``, ``self.params`` and the hyper-parameter attributes)."""

    # ---- argument plumbing ---------------------------------------------------------------------------------------
    def _arg(self, t: torch.Tensor, shape, name: str) -> torch.Tensor:
        if not isinstance(t, torch.Tensor) or tuple(t.shape) != tuple(shape):
            got = tuple(t.shape) if isinstance(t, torch.Tensor) else type(t)
            raise L.CmpcError(f"{name}: expected a tensor of shape {tuple(shape)}, got {got}")
        if t.device != self._head.device or t.dtype != torch.float32:
            raise L.CmpcError(f"{name}: expected float32 on {self._head.device}, got {t.dtype} on {t.device}")
        return t.contiguous()

    def _level(self, level: str) -> int:
        if level not in LEVELS:
            raise L.CmpcError(f"unknown level {level!r} (one of {LEVELS})")
        return LEVELS.index(level)

    def _module(self, level: str, suffix: str) -> Tuple[str, int]:
        """'c3gv_f1' / 'c4_2_f2' -> (exchange module name, slot in EXG)"""
        if not level.endswith(suffix) or level[:len(level) - len(suffix)] not in EXG:
            raise L.CmpcError(f"unknown level {level!r}: expected <module>{suffix} with module in {EXG}")
        x = level[:len(level) - len(suffix)]
        return x, EXG.index(x)

    def _load_cast(self, x: torch.Tensor, cols: int, dst: torch.Tensor, scale: float = 1.0):
        """fp32 [rows, cols] -> fp16 dst[:, :cols] (dst pads stay zero)"""
        h = self._head
        rows = x.numel() // cols
        h._ck(h.lib.cmpc_scale_cast_f32_f16(x.data_ptr(), cols, scale, dst.data_ptr(), dst.stride(0), rows, cols, h._stream()), "cast")

    def _load_map(self, x: torch.Tensor, dst: torch.Tensor, mode: int):
        """fp32 [M, C] -> fp16 C-wide operand buffer; mode > 0 appends the 8 spatial channels, -1 the homogeneous 1.0"""
        h, d = self._head, self._head.d
        M = h.B * d.N
        h._ck(h.lib.cmpc_rownorm_f16(x.data_ptr(), d.C, h.buf["ones"].data_ptr(), dst.data_ptr(), d.LDC, M, d.C,
                                     d.h if mode > 0 else mode, d.w if mode > 0 else 0, d.N, h._stream()), "rownorm")

    def _load_words(self, words_feat):
        h, d, b = self._head, self._head.d, self._head.buf
        wf = self._arg(words_feat, (h.B, 1, d.T, d.R), "words_feat")
        if self.seq_mask is None:
            raise L.CmpcError("seq_mask is not set: call lstm(lstm_outputs) first (CMPC_model.py:163)")
        b["words32"].copy_(wf.view(h.B * d.T, d.R))
        self._load_cast(wf, d.R, b["words16"])
        b["mask"].copy_(self._arg(self.seq_mask, (h.B, 1, d.T, 1), "seq_mask").view(-1))

    def _load_parse(self, words_parse):
        h, d = self._head, self._head.d
        h.buf["parse"].copy_(self._arg(words_parse, (h.B, 1, d.T, 4), "words_parse").view(h.B, d.T, 4))

    def _load_lang(self, lang_feat, which: str):
        h, d, b = self._head, self._head.d, self._head.buf
        lf = self._arg(lang_feat, (h.B, 1, 1, d.R), "lang_feat")
        b[which + "32"].copy_(lf.view(h.B, d.R))
        self._load_cast(lf, d.R, b[which + "16"])

    def _load_visual(self, i: int, visual_feat, spatial):
        """the (already l2-normalised) lateral map -> MUTAN operand [visual | spatial] with unit row scale"""
        h, d, b = self._head, self._head.d, self._head.buf
        self._check_spatial(spatial)
        self._load_map(self._arg(visual_feat, (h.B, d.h, d.w, d.C), "visual_feat"), b["xlat16"], 1)
        b["rowss"][2 * i].fill_(1.0)

    def _load_feat(self, feat, name: str, dst: str) -> torch.Tensor:
        h, d = self._head, self._head.d
        self._load_cast(self._arg(feat, (h.B, d.h, d.w, d.Mm), name), d.Mm, h.buf[dst])
        return h.buf[dst]

    def _map_out(self, buf: torch.Tensor, cols: int, shape=None) -> torch.Tensor:
        h, d = self._head, self._head.d
        return buf[:, :cols].float().reshape(shape or (h.B, d.h, d.w, cols))

    def _compact(self, t: torch.Tensor) -> torch.Tensor:
        """[B, 3, GW] per-round buffer viewed as the [B, GW] block a single-module launch (nmod = 1) fills"""
        h, d = self._head, self._head.d
        return t.view(-1)[:h.B * d.GW].view(h.B, d.GW)

    @on_device
    def generate_spatial_batch(self) -> torch.Tensor:
        """util/processing_tools.py:5-17 as the kernels evaluate it (CMPC_model.py:104): [B, h, w, 8] fp32"""
        h, d, b = self._head, self._head.d, self._head.buf
        M = h.B * d.N
        h._ck(h.lib.cmpc_spatial_fixup_f16(b["z16"].data_ptr(), d.LDC, b["ones"].data_ptr(), M, d.C, d.h, d.w, h._stream()),
              "spatial_fixup")
        return b["z16"][:, d.C:d.C + 8].float().reshape(h.B, d.h, d.w, 8)

    def _check_spatial(self, spatial):
        """the kernels regenerate the 8 coordinate channels on chip; a caller-supplied tensor must be that map"""
        h, d = self._head, self._head.d
        sp = self._arg(spatial, (h.B, d.h, d.w, 8), "spatial")
        if not torch.allclose(sp, self.generate_spatial_batch(), atol=2e-3):
            raise L.CmpcError("spatial must be generate_spatial_batch(batch_size, vf_h, vf_w) (CMPC_model.py:104)")

    # ---- language side ---------------------------------------------------------------------------------------------
    @on_device
    def lstm(self, lstm_outputs=None):
        """CMPC_model.py:158-164, i.e. lstm() after the dynamic_rnn (:144-157, an upstream producer): returns
        (words_feat [B,1,T,R], lang_feat [B,1,R]) and sets self.seq_mask [B,1,T,1]."""
        h, d, b = self._head, self._head.d, self._head.buf
        lo = lstm_outputs if lstm_outputs is not None else self.lstm_outputs
        if lo is None:
            raise L.CmpcError("lstm(): feed lstm_outputs (the dynamic_rnn outputs, zero past seq_len)")
        lo = self._arg(lo, (h.B, d.T, d.R), "lstm_outputs")
        h._st_words(lo)
        words_feat = b["words32"].view(h.B, 1, d.T, d.R).clone()
        self.seq_mask = b["mask"].view(h.B, 1, d.T, 1).clone()
        # lang_feat = reduce_sum(words_feat, -2) (:161; never consumed downstream): ones[T] x words per sample
        lang_feat = torch.empty(h.B, 1, d.R, dtype=torch.float32, device=h.device)
        h._ck(h.lib.cmpc_small_linear_f32(b["ones"].data_ptr(), d.T, 0, b["words32"].data_ptr(), d.R, d.T * d.R, None, 0,
                                          lang_feat.data_ptr(), d.R, d.R, h.B, 1, d.T, d.R, 0, h._stream()), "small_linear")
        return words_feat, lang_feat

    @on_device
    def build_lang_parser(self, words_feat):
        """:347-357 -> words_parse [B,1,T,4] (Entity, Attribute, Relation, Unnecessary), masked by self.seq_mask"""
        h, d = self._head, self._head.d
        self._load_words(words_feat)
        h._st_parse()
        self.words_parse = h.buf["parse"].view(h.B, 1, d.T, 4).clone()
        return self.words_parse

    def _weighted_lang(self, words_parse, words_feat, which):
        h, d = self._head, self._head.d
        self._load_words(words_feat)
        self._load_parse(words_parse)
        h._st_parse(given_parse=True)
        return h.buf[which + "32"].view(h.B, 1, 1, d.R).clone()

    @on_device
    def valid_lang(self, words_parse, words_feat):
        """:166-178 -> l2_normalize(sum_t (E_t + A_t) words_t)  [B,1,1,R]"""
        return self._weighted_lang(words_parse, words_feat, "valid")

    @on_device
    def nec_lang(self, words_parse, words_feat):
        """:180-192 -> l2_normalize(sum_t (E_t + A_t + R_t) words_t)  [B,1,1,R]"""
        return self._weighted_lang(words_parse, words_feat, "nec")

    # ---- entity perception ---------------------------------------------------------------------------------------
    @on_device
    def mutan_head(self, lang_feat, spatial_feat, visual_feat, level=''):
        """:295-309, level = '<c5|c4|c3>_head<1..5>' -> tanh(vis_trans([visual | spatial])) * tanh(lang_trans(lang))"""
        h, d, b, W = self._head, self._head.d, self._head.buf, self._head.Wt
        lvl, _, hd = level.partition("_head")
        i = self._level(lvl)
        if hd not in ("1", "2", "3", "4", "5"):
            raise L.CmpcError(f"unknown MUTAN head {level!r}")
        k = int(hd) - 1
        h._begin()
        self._load_lang(lang_feat, "valid")
        h._st_valid_derived()
        self._load_visual(i, visual_feat, spatial_feat)
        cache = self.__dict__.setdefault("_head_w", {})
        if level not in cache:                       # the fused kernel interleaves the five heads; un-interleave head k once
            cache[level] = W[f"mutan_w_{lvl}"].view(d.CH, 5, 48, d.LDC)[:, k].reshape(d.CH * 48, d.LDC)[:d.C].contiguous()
        h._gemm(b["xlat16"], d.C + 8, cache[level], d.C, b["tmp32"], bias=W[f"mutan_b_{lvl}"][k], act=2,
                gate=b["lang"][:, (i * 5 + k) * d.C:], rows_per_sample=d.N)
        return self._map_out(b["tmp32"], d.C)

    @on_device
    def mutan_fusion(self, lang_feat, spatial_feat, visual_feat, level=''):
        """:311-328 -> l2_normalize(tanh(sum of the five heads), 3)  [B,h,w,C]"""
        h, d = self._head, self._head.d
        i = self._level(level)
        h._begin()
        self._load_lang(lang_feat, "valid")
        h._st_valid_derived()
        self._load_visual(i, visual_feat, spatial_feat)
        h._st_mutan(i)
        return self._map_out(h.buf["x16"], d.C)

    # ---- relation-aware reasoning --------------------------------------------------------------------------------
    @on_device
    def graph_conv(self, graph_feat, nodes_num, nodes_dim, adj_mat, graph_name="", level=""):
        """:359-374 -> relu(LN(update(relu(X + LN(adj @ X)))))  [B,1,N,C].
        adj_mat is the FACTORED adjacency (gw_w, gw_v), both [B,N,T]: adj = gw_w @ gw_v^T (:400) is never formed on this
        device (it is N x N per sample), so a dense [B,N,N] tensor is rejected."""
        h, d, b = self._head, self._head.d, self._head.buf
        i = self._level(level)
        if graph_name != "spa_graph" or nodes_num != d.N or nodes_dim != d.C:
            raise L.CmpcError(f"graph_conv: only graph_name='spa_graph' with nodes_num={d.N}, nodes_dim={d.C} exists in the reference")
        if not (isinstance(adj_mat, (tuple, list)) and len(adj_mat) == 2):
            raise L.CmpcError("graph_conv: pass adj_mat=(gw_w, gw_v); the dense N x N adjacency is not materialised on the device")
        gw_w = self._arg(adj_mat[0], (h.B, d.N, d.T), "gw_w")
        gw_v = self._arg(adj_mat[1], (h.B, d.N, d.T), "gw_v")
        M = h.B * d.N
        h._begin()
        self._load_map(self._arg(graph_feat, (h.B, 1, d.N, d.C), "graph_feat"), b["x16"], -1)
        for src, stage, dst, sc in ((gw_w, b["affi"], b["w16"], 1.0), (gw_v, b["taps"], b["v16"], h.v_scale)):
            stage.zero_()
            stage[:, :d.T].copy_(src.view(M, d.T))          # [M, T] -> zero-padded [M, 32] operand rows
            self._load_cast(stage, 32, dst, sc)
        h._st_graph_conv(i, normalize=False)
        return self._map_out(b["g16"], d.C, (h.B, 1, d.N, d.C))

    @on_device
    def build_spa_graph(self, spa_graph, words_feat, spatial, words_parse, level=""):
        """:376-410 -> l2_normalize(graph_conv(...), 3)  [B,h,w,C]; sets self.gw_w / self.gw_v [B,N,T] (:389-391)"""
        h, d, b = self._head, self._head.d, self._head.buf
        i = self._level(level)
        self._check_spatial(spatial)
        h._begin()
        self._load_words(words_feat)
        self._load_parse(words_parse)
        h._st_parse(given_parse=True)                       # relation weights R_t / sqrt(C) of the supplied words_parse
        h._st_words_derived()
        self._load_map(self._arg(spa_graph, (h.B, d.h, d.w, d.C), "spa_graph"), b["x16"], -1)
        h._st_affinity(i, True)
        h._st_graph_conv(i)
        self.gw_w, self.gw_v = b["gw_w"].view(h.B, d.N, d.T).clone(), b["gw_v"].view(h.B, d.N, d.T).clone()
        return self._map_out(b["g16"], d.C)

    @on_device
    def build_lang2vis(self, visual_feat, words_feat, lang_feat, words_parse, spatial, level=""):
        """:330-345 -> relu(fusion conv([vis_la_sp | spa_graph | tile(valid_lang) | spatial]))  [B,h,w,mlp_dim].
        lang_feat is accepted and ignored, exactly like the reference (:330 never reads it)."""
        h, d, b = self._head, self._head.d, self._head.buf
        i = self._level(level)
        h._begin()
        self._load_words(words_feat)
        self._load_parse(words_parse)
        h._st_parse(given_parse=True)
        h._st_words_derived()
        h._st_valid_derived()
        self._load_visual(i, visual_feat, spatial)
        h._st_mutan(i)
        h._st_affinity(i, True)
        h._st_graph_conv(i)
        h._st_fusion(i)
        self.gw_w, self.gw_v = b["gw_w"].view(h.B, d.N, d.T).clone(), b["gw_v"].view(h.B, d.N, d.T).clone()
        return self._map_out(b[f"fus16_{level}"], d.Mm)

    # ---- text-guided exchange ------------------------------------------------------------------------------------
    @on_device
    def global_vec(self, feat, lang_feat, level=""):
        """:212-243, level = '<module>gv_f1' -> l2_normalize(gv_lang conv([attention-pooled feat | lang]))  [B,1,1,mlp_dim]
        (per-sample normalisation: the reference at batch 1, see CMPCHeadB200)"""
        h, d, b = self._head, self._head.d, self._head.buf
        _, slot = self._module(level, "gv_f1")
        h._begin()
        self._load_lang(lang_feat, "nec")
        h._st_nec_derived()
        f = self._load_feat(feat, "feat", "e3")
        h._st_global_vec((f,), slot, 1)
        return self._compact(b["gv"])[:, :d.Mm].reshape(h.B, 1, 1, d.Mm).clone()

    @on_device
    def lang_se(self, feat, lang_feat, level=""):
        """:194-210, level = '<module>_f1|_f2' -> relu(trans_feat conv(feat)) * sigmoid(lang_feat conv(lang_feat)).
        feat [B,h,w,mlp_dim]; lang_feat [B,1,1,mlp_dim] (the global vector)."""
        h, d, b, W = self._head, self._head.d, self._head.buf, self._head.Wt
        which = level[-3:]
        x, slot = self._module(level, which)
        if which not in ("_f1", "_f2"):
            raise L.CmpcError(f"unknown lang_se level {level!r}")
        gv = self._arg(lang_feat, (h.B, 1, 1, d.Mm), "lang_feat")
        wf, bf = (W["wf1"], W["bf1"]) if which == "_f1" else (W["wf2"], W["bf2"])
        gate = b["gate1"]                                                      # [B, 3, GW]; row 0 of each sample is used
        h._ck(h.lib.cmpc_small_linear_f32(gv.data_ptr(), d.Mm, d.Mm, wf[slot].data_ptr(), d.Mm, 0, bf[slot].data_ptr(), 0,
                                          gate.data_ptr(), 3 * d.GW, 3 * d.GW, h.B, 1, d.Mm, d.Mm, 3, h._stream()), "small_linear")
        f = self._load_feat(feat, "feat", "e3")
        h._st_lang_se(f, f"{x}{which}", gate[:, 0], b["se1"])
        return self._map_out(b["se1"], d.Mm)

    @on_device
    def gated_exchange_module(self, feat, feat1, feat2, lang_feat, level=""):
        """:245-259 -> feat + lang_se(feat1, gv, _f1) + lang_se(feat2, gv, _f2), gv = global_vec(feat, lang_feat)"""
        h, d, b = self._head, self._head.d, self._head.buf
        if level not in EXG:
            raise L.CmpcError(f"unknown exchange module {level!r} (one of {EXG})")
        slot = EXG.index(level)
        h._begin()
        self._load_lang(lang_feat, "nec")
        h._st_nec_derived()
        f0, f1, f2 = (self._load_feat(t, n, dst) for t, n, dst in ((feat, "feat", "e3"), (feat1, "feat1", "e4"), (feat2, "feat2", "e5")))
        h._st_global_vec((f0,), slot, 1)
        h._st_lang_se(f1, f"{level}_f1", self._compact(b["gate1"]), b["se1"])
        h._st_lang_se(f2, f"{level}_f2", self._compact(b["gate2"]), b["se2"])
        h._ck(h.lib.cmpc_add3_l2norm_f16(f0.data_ptr(), b["se1"].data_ptr(), b["se2"].data_ptr(), d.GW, b["g3"].data_ptr(), d.GW,
                                         h.B * d.N, d.GW, 0, None, h._stream()), "add3_l2norm")
        return self._map_out(b["g3"], d.Mm)

    @on_device
    def gated_exchange_fusion_lstm_2times(self, feat3, feat4, feat5, lang_feat):
        """:261-293 -> two exchange rounds (each l2-normalised) + the ConvLSTM over (c3, c4, c5): last h  [B,h,w,mlp_dim]"""
        h, d, b = self._head, self._head.d, self._head.buf
        h._begin()
        self._load_lang(lang_feat, "nec")
        h._st_nec_derived()
        f3, f4, f5 = (self._load_feat(t, n, f"fus16_{l}") for t, n, l in ((feat3, "feat3", "c3"), (feat4, "feat4", "c4"), (feat5, "feat5", "c5")))
        f3, f4, f5 = h._st_exchange_round(0, f3, f4, f5, ("e3", "e4", "e5"))
        f3, f4, f5 = h._st_exchange_round(1, f3, f4, f5, ("g3", "g4", "g5"))
        return self._map_out(h._st_convlstm((f3, f4, f5)), d.Mm)

    # ---- convolutions --------------------------------------------------------------------------------------------
    def _conv(self, name, x, filter_size, in_filters, out_filters, strides):
        """:412-417 conv2d(x, DW, strides, 'SAME') + biases for the two shapes the head uses: any 1x1 conv (a GEMM over
        the NHWC rows) and the 3x3 one-channel score convs on the [B,h,w,mlp_dim] maps.  strides must be [1,1,1,1]."""
        h, d, b = self._head, self._head.d, self._head.buf
        if list(strides) != [1, 1, 1, 1]:
            raise L.CmpcError("_conv: the head only uses unit strides")
        dw, bias = self.params[name + "/DW"], self.params[name + "/biases"]
        if tuple(dw.shape) != (filter_size, filter_size, in_filters, out_filters):
            raise L.CmpcError(f"_conv: variable {name}/DW has shape {tuple(dw.shape)}")
        if x.shape[-1] != in_filters or x.dtype != torch.float32 or x.device != h.device:
            raise L.CmpcError(f"_conv: x must be float32 [..., {in_filters}] on {h.device}")
        x = x.contiguous()
        if filter_size == 3 and out_filters == 1 and (name == "score" or name[6:] in LEVELS) and tuple(x.shape) == (h.B, d.h, d.w, d.Mm):
            f = self._load_feat(x, "x", "se1")
            pred = torch.empty(h.B, d.h, d.w, 1, dtype=torch.float32, device=h.device)
            h._st_score(f, name, pred, None, None, tag="score_pred")
            return pred
        if filter_size != 1 or in_filters % 4 or out_filters % 4:
            raise L.CmpcError("_conv: only 1x1 convs with channel counts that are multiples of 4, and the 3x3 score convs, exist on the device")
        cache = self.__dict__.setdefault("_conv_w", {})
        if name not in cache:
            cache[name] = (pack_conv1x1(dw.to(h.device, torch.float32), rows_pad=rup(out_filters, 32)),
                           bias.to(h.device, torch.float32).contiguous())
        w16, bias32 = cache[name]
        rows = x.numel() // in_filters
        a16 = torch.zeros(rows, rup(in_filters, 64), dtype=torch.float16, device=h.device)
        self._load_cast(x, in_filters, a16)
        out = torch.zeros(rows, rup(out_filters, 8), dtype=torch.float32, device=h.device)
        h._gemm(a16, in_filters, w16, out_filters, out, bias=bias32, w_rows=out_filters)
        return out[:, :out_filters].reshape(*x.shape[:-1], out_filters).contiguous()
