"""Backward pass of the head (what TF autodiff builds at CMPC_model.py:461), stage by stage, on the
device.  Mirrors head.py: every forward stage `_st_*` has a `bwd_*` here that consumes the activations the forward kept in
a `Saved` store and produces (a) the gradient w.r.t. the stage's inputs and (b) gradients of its parameters, accumulated
in fp32 buffers laid out like the packed weights; `grads_tf()` maps them back to the TF variable names / shapes.

HeadBackward.backward() runs the whole pass: loss / upsample / score convs (:128-142, :439-445), ConvLSTM (:287-290, util/cell.py),
both exchange rounds (:194-284), per level fusion conv / graph_conv / dense aggregation / affinity softmaxes (:330-410), MUTAN and
the lateral convs (:108-113, :295-328), the language side (:159-192, :347-357).
torch is used for buffers and for re-laying-out small parameter tensors; all arithmetic is in libcmpc_b200.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from .head import on_device

from . import _lib as L
from .weights import EXG, LEVELS, rup


class Saved:
    """Activations kept by a training-mode forward (head.saved = Saved(device))."""

    def __init__(self, device):
        self.device = device
        self.t: Dict[str, torch.Tensor] = {}

    def alloc(self, name, shape, dtype):
        t = self.t.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = self.t[name] = torch.zeros(*shape, dtype=dtype, device=self.device)
        return t


class HeadBackward:
    def __init__(self, head):
        self.h = head
        self.device = head.device
        d, dev = head.d, head.device
        Mm, GW, N = d.Mm, d.GW, d.N
        f32 = dict(dtype=torch.float32, device=dev)
        self.score_names = ["score"] + [f"score_{l}" for l in ("c5", "c4", "c3")]
        # consumers of each source map inside an exchange round: (module index, which lang_se); see head._st_exchange_round
        self.exg_consumers = {0: ((1, "_f1"), (2, "_f1")), 1: ((0, "_f1"), (2, "_f2")), 2: ((0, "_f2"), (1, "_f2"))}
        self.pack_weights()
        # parameter gradients (fp32, packed layouts): shapes first, then ONE flat arena laid out in the order the backward pass
        # finishes them (self.buckets), so that a data-parallel trainer can all-reduce a contiguous slice as soon as a stage is done
        spec = {
            "lstm_w": (2 * GW, 4 * GW),                 # rows: [x | h] input channel, cols: gate * GW + cout
            "lstm_W_ci": (N, GW,), "lstm_W_cf": (N, GW,), "lstm_W_co": (N, GW,),
            "lstm_ln_gamma": (5, GW,), "lstm_ln_beta": (5, GW,),
        }
        for name in self.score_names:
            spec[name + "_w9"] = (16, GW,)
            spec[name + "_b"] = (1,)
        # ---- text-guided exchange (:194-259) ----
        kp = rup(Mm, 64)
        R = d.R
        for x in EXG:
            for f in ("_f1", "_f2"):
                spec[f"se_w_{x}{f}"] = (GW, GW,)          # [cin, cout]
                spec[f"se_b_{x}{f}"] = (GW,)
        for nm in ("wf1", "wf2", "wg", "key"):
            spec[nm] = (6, Mm, Mm,)                        # [slot, input, output] ("key": [slot, o, cin])
        for nm in ("bf1", "bf2", "q_b", "gvl_b"):
            spec[nm] = (6, Mm,)
        spec["q_w"] = (6, R, Mm,)                          # lang_query DW [R, Mm]
        spec["gvl_w"] = (6, R, Mm,)                        # language rows of gv_lang DW
        # ---- per level: fusion conv, graph conv, affinity (:330-410) ----
        C_, LDC, LDR, T = d.C, d.LDC, d.LDR, d.T
        for lvl in LEVELS:
            spec[f"fusion_w_{lvl}"] = (2 * LDC, GW)        # rows [0, LDC): vis_la_sp, [LDC, 2 LDC): spa_graph | spatial
            spec[f"fusion_lang_{lvl}"] = (R, GW,)
            spec[f"fusion_b_{lvl}"] = (GW,)
            spec[f"gupd_w_{lvl}"] = (LDC, LDC,)
            spec[f"gupd_b_{lvl}"] = (LDC,)
            for nm in ("gfeat_gamma", "gfeat_beta", "gupdate_gamma", "gupdate_beta"):
                spec[f"{nm}_{lvl}"] = (LDC,)
            spec[f"gt_w_{lvl}"] = (LDC, LDR,)               # rows: cin (row C = bias), cols: o
        # ---- MUTAN (:295-328) + lateral convs (:108-113) ----
        CH = d.CH
        CHP = rup(CH * 240, 64)
        self.CHP = CHP
        for li, lvl in enumerate(LEVELS):
            spec[f"mutan_w_{lvl}"] = (LDC, CHP,)
            spec[f"mutan_b_{lvl}"] = (5, LDC,)
            spec[f"lat_w_{lvl}"] = (d.cin[lvl], LDC,)
            spec[f"lat_b_{lvl}"] = (LDC,)
            spec[f"ltrans_w_{lvl}"] = (R, 5 * C_)
            spec[f"ltrans_b_{lvl}"] = (5 * C_,)
        M = head.B * N
        B = head.B
        f16 = dict(dtype=torch.float16, device=dev)
        # ---- language side (:159-192, :347-357, :378) ----
        HID, HIDP, BT = d.HID, d.HIDP, head.B * T
        spec["parse2_w"], spec["parse2_b"] = (HID, 4), (4,)
        spec["parse1_w"], spec["parse1_b"] = (R, HID), (HID,)
        for lvl in LEVELS:
            spec[f"wtrans_w_{lvl}"], spec[f"wtrans_b_{lvl}"] = (R, R), (R,)
        self._build_grad_arena(spec)
        self.dwords = torch.zeros(BT, R, **f32)
        self.dlogit = torch.zeros(BT, 4, **f32)
        self.dhid = torch.zeros(BT, HIDP, **f32)
        self.d_lstm = torch.zeros(BT, R, **f32)
        self.dpre_m16 = torch.zeros(M, CHP, **f16)
        self.d_lang = torch.zeros(B, 15 * C_, **f32)
        self.dpl = torch.zeros(B, 15 * C_, **f32)
        self.S = torch.zeros(B, GW, **f32)
        self.dpre16 = torch.zeros(M, GW, **f16)
        self.dxg, self.dgg = torch.zeros(M, LDC, **f32), torch.zeros(M, LDC, **f32)
        self.dln2, self.dres, self.dzl, self.dagg, self.daff = (torch.zeros(M, LDC, **f32) for _ in range(5))
        self.du16, self.dyg16 = torch.zeros(M, LDC, **f16), torch.zeros(M, LDC, **f16)
        self.lnsums = torch.zeros(2, B, 2, dtype=torch.float64, device=dev)
        self.t1 = torch.zeros(B, LDC, 32, **f32); self.t1_16 = torch.zeros(B, LDC, 64, **f16)
        self.t23 = torch.zeros(2, B, 32, LDC, **f32); self.t23_16 = torch.zeros(2, B, 32, LDC, **f16)
        self.dwm, self.dvm = torch.zeros(M, 32, **f32), torch.zeros(M, 32, **f32)
        self.draw16 = torch.zeros(M, 32, **f16)
        self.cs_ws = torch.zeros(B, 32, **f32)
        self.drgate = torch.zeros(B, 32, **f32)
        self.gtT16 = torch.zeros(B, LDC, 64, **f16)
        self.dgt = torch.zeros(B * T + 32, LDC, **f32); self.dgt16 = torch.zeros(B * T + 32, LDC, **f16)
        self.d_wt = [torch.zeros(B * T, LDR, **f32) for _ in LEVELS]      # d loss / d words_trans_<level> output
        self.d_valid = torch.zeros(B, R, **f32)
        self.colsum = torch.zeros(B, 3, 4, GW, **f32)
        self.dpre1, self.dpre2, self.dz = (torch.zeros(2, B, 3, GW, **f32) for _ in range(3))
        self.dpool = torch.zeros(B, 3, GW, **f32)
        self.gv_dot = torch.zeros(3, **f32)          # gv_norm='batch': sum over the batch of gv . d gv per module
        self.du = torch.zeros(2, B, 3, GW, **f32)
        self.dq = torch.zeros(B, GW, **f32)
        self.d_nec = torch.zeros(B, R, **f32)
        self.ds = [torch.zeros(M, GW, **f32) for _ in range(3)]
        self.dp16 = [[torch.zeros(M, GW, dtype=torch.float16, device=dev) for _ in range(2)] for _ in range(3)]
        self.dgemm = torch.zeros(M, GW, **f32)
        self.ones = torch.ones(max(B * d.T, 64), **f32)
        self.ws = torch.zeros(head.lib.cmpc_convlstm_bwd_workspace_floats(head.B, N, GW), **f32)
        self.sums = torch.zeros(head.B, 10, **f32)
        self.dy16 = torch.zeros(M, 4 * GW, dtype=torch.float16, device=dev)
        self.dcnew = torch.zeros(M, GW, **f32)
        self.dcn = [torch.zeros(M, GW, **f32) for _ in range(2)]
        self.dxh = torch.zeros(M, 2 * GW, **f32)
        self.dpred = torch.zeros(head.B, d.h, d.w, **f32)
        self.d9 = torch.zeros(M, 64, dtype=torch.float16, device=dev)

    # ---- gradient arena ----------------------------------------------------------------------------------------------------
    _side = None      # side stream of backward_stages
    BUCKETS = ("fuse", "exchange", "c5_graph", "c5_ltrans", "c5_mutan", "c4_graph", "c4_ltrans", "c4_mutan", "c3_graph", "c3_ltrans",
               "c3_mutan", "language")

    @staticmethod
    def bucket_of(name: str) -> str:
        """the backward stage after which the packed gradient buffer `name` is final (see backward_stages)"""
        if name.startswith("lstm_") or name in ("score_w9", "score_b"):
            return "fuse"
        if name in ("key", "q_b", "gvl_b", "q_w", "gvl_w"):      # bwd_exchange_language: two slots under each level's graph stage (side stream)
            return LEVELS[-1] + "_graph"
        if name.startswith(("se_w_", "se_b_", "score_c")) or name in ("wf1", "wf2", "wg", "bf1", "bf2"):
            return "exchange"
        lvl = name.rsplit("_", 1)[-1]
        if lvl in LEVELS:
            if name.startswith(("fusion_", "gupd_", "gfeat_", "gupdate_", "gt_w_")):
                return lvl + "_graph"
            if name.startswith(("ltrans_", "mutan_b_")):      # final after the MUTAN backward kernel, before the big weight-gradient GEMMs
                return lvl + "_ltrans"
            if name.startswith(("mutan_w_", "lat_")):
                return lvl + "_mutan"
        return "language"                            # parse*, wtrans_*

    def _build_grad_arena(self, spec):
        from .parallel import plan_arena
        total, place, self.bucket_range = plan_arena(spec, self.bucket_of, self.BUCKETS)
        self.garena = torch.zeros(total, dtype=torch.float32, device=self.h.device)
        self.g = {k: self.garena[o:o + n].view(*[int(e) for e in spec[k]]) for k, (o, n) in place.items()}

    def bucket_view(self, bname: str) -> torch.Tensor:
        a, b = self.bucket_range[bname]
        return self.garena[a:b]

    @on_device
    def pack_weights(self):
        """fp16 / transposed operand copies of the parameters for the input-gradient GEMMs and the small per-sample maps, from
        head.params (call again after every optimizer step): a 1x1 conv's TF kernel [Cin, Cout] IS the [n_out = cin, k = cout]
        "weight" of its dgrad GEMM."""
        head = self.h
        d, dev = head.d, head.device
        reg = self.__dict__.setdefault("_packed", {})

        def pk(key, t):
            """refresh IN PLACE after the first call: the addresses are baked into launches captured in a CUDA graph"""
            old = reg.get(key)
            if old is None or old.shape != t.shape or old.dtype != t.dtype:
                reg[key] = old = t
            else:
                old.copy_(t)
            return old
        P = {k: v.to(device=dev, dtype=torch.float32) for k, v in head.params.items()}
        Mm, GW, R, C_, LDC = d.Mm, d.GW, d.R, d.C, d.LDC
        f32 = dict(dtype=torch.float32, device=dev)
        kp = rup(Mm, 64)
        kern = P["rnn/conv_lstm_cell/kernel"][0, 0]                      # [2Mm, 4Mm]
        wT = torch.zeros(2 * GW, 4 * GW, **f32)
        for grp in range(2):
            for gate in range(4):
                wT[grp * GW:grp * GW + Mm, gate * GW:gate * GW + Mm] = kern[grp * Mm:(grp + 1) * Mm, gate * Mm:(gate + 1) * Mm]
        self.lstm_wT = pk("lstm_wT", wT.half().contiguous())
        self.score_wT = {}
        for name in self.score_names:
            w = torch.zeros(GW, 64, **f32)
            w[:Mm, :9] = P[name + "/DW"][:, :, :, 0].reshape(9, Mm).t()
            self.score_wT[name] = pk(("score_wT", name), w.half().contiguous())
        self.exg_wT = {}
        for rnd in range(2):
            for src, cons in self.exg_consumers.items():
                w = torch.zeros(GW, 2 * kp, **f32)
                for j, (mi, f) in enumerate(cons):
                    w[:Mm, j * kp:j * kp + Mm] = P[f"trans_feat_{EXG[rnd * 3 + mi]}{f}/DW"][0, 0]      # [cin, cout]
                self.exg_wT[(rnd, src)] = pk(("exg_wT", (rnd, src)), w.half().contiguous())
        self.key_w = pk("key_w", torch.stack([P[f"spa_graph_key_{x}gv_f1/DW"][0, 0] for x in EXG]).contiguous())              # [6, cin, o]
        self.q_wT = pk("q_wT", torch.stack([P[f"lang_query_{x}gv_f1/DW"][0, 0].t() for x in EXG]).contiguous())              # [6, Mm, R]
        self.gvl_wT = pk("gvl_wT", torch.stack([P[f"gv_lang_{x}gv_f1/DW"][0, 0][Mm:].t() for x in EXG]).contiguous())          # [6, Mm, R]
        self.fus_wT, self.fus_lang_wT, self.gupd_wT, self.gt_wT, self.mutan_wT, self.ltrans_wT = {}, {}, {}, {}, {}, {}
        CH = d.CH
        CHP = rup(CH * 240, 64)
        for lvl in LEVELS:
            dw = P[f"fusion_{lvl}/DW"][0, 0]                                   # [2C + R + 8, Mm]
            w = torch.zeros(2 * LDC, kp, **f32)
            w[:C_, :Mm] = dw[:C_]
            w[LDC:LDC + C_, :Mm] = dw[C_:2 * C_]
            self.fus_wT[lvl] = pk(("fus_wT", lvl), w.half().contiguous())
            self.fus_lang_wT[lvl] = pk(("fus_lang_wT", lvl), dw[2 * C_:2 * C_ + R].t().contiguous())      # [Mm, R]
            w = torch.zeros(LDC, LDC, **f32)
            w[:C_, :C_] = P[f"gconv_update_spa_graph_{lvl}/DW"][0, 0]
            self.gupd_wT[lvl] = pk(("gupd_wT", lvl), w.half().contiguous())
            w = torch.zeros(rup(R, 8), LDC, **f32)
            w[:R, :C_] = P[f"spa_graph_trans2_{lvl}/DW"][0, 0].t()              # [o, cin]
            w[:R, C_] = P[f"spa_graph_trans2_{lvl}/biases"]
            self.gt_wT[lvl] = pk(("gt_wT", lvl), w.half().contiguous())
            w = torch.zeros(LDC, CHP, **f32)
            w[:, :CH * 240] = head.Wt[f"mutan_w_{lvl}"].float().t()              # [cin (+ spatial rows), packed (chunk, head, channel)]
            self.mutan_wT[lvl] = pk(("mutan_wT", lvl), w.half().contiguous())
            self.ltrans_wT[lvl] = pk(("ltrans_wT", lvl), torch.cat([P[f"lang_trans_{lvl}_head{k + 1}/DW"][0, 0].t() for k in range(5)], 0).contiguous())   # [5C, R]
        self.parse2_wT = pk("parse2_wT", P["words_parse_2/DW"][0, 0].t().contiguous())            # [4, HID]
        self.parse1_wT = pk("parse1_wT", P["words_parse_1/DW"][0, 0].t().contiguous())            # [HID, R]
        self.wtrans_wT = [pk(("wtrans_wT", lvl), P[f"words_trans_{lvl}/DW"][0, 0].t().contiguous()) for lvl in LEVELS]     # [o, cin]

    @on_device
    def zero_grads(self):
        self.garena.zero_()

    # ---- loss + upsample + score conv -----------------------------------------------------------------------------
    def bwd_score(self, up, target_fine, coef, feat16, name, out):
        """d loss_k / d feat for loss_k = coef * mean_b sum_px CE(up, target) with up = resize(conv3x3(feat) + b):
        writes out fp32 [M, GW]; accumulates the score kernel / bias gradients."""
        h, d, lib = self.h, self.h.d, self.h.lib
        M, st = h.B * d.N, h._stream()
        h._ck(lib.cmpc_score_bwd_dpred(up.data_ptr(), target_fine.data_ptr(), coef / h.B, h.B, d.h, d.w, d.H, d.W, self.dpred.data_ptr(),
                                       self.g[name + "_b"].data_ptr(), st), "score_bwd_dpred")
        h._ck(lib.cmpc_score_bwd_taps(self.dpred.data_ptr(), h.B, d.h, d.w, self.d9.data_ptr(), 64, st), "score_bwd_taps")
        h._gemm(self.d9, 16, self.score_wT[name], d.GW, out, group=(d.GW, d.Mm))                       # dF = D9 . w9
        h._ck(lib.cmpc_gemm_atb_f16(self.d9.data_ptr(), 64, 16, feat16.data_ptr(), d.GW, d.GW, M, self.g[name + "_w9"].data_ptr(), d.GW, 0, st),
              "gemm_atb")                                                                                   # dw9 = D9^T F
        return out

    # ---- ConvLSTM -----------------------------------------------------------------------------------------------------
    def bwd_convlstm(self, dh_last):
        """dh_last fp32 [M, GW] = d loss / d h of the last step.  Returns [d x_0, d x_1, d x_2] (fp32 [M, GW] views, valid until
        the next call) = gradients w.r.t. the three exchanged maps fed to the cell."""
        h, d, lib, W, sv = self.h, self.h.d, self.h.lib, self.h.Wt, self.h.saved.t
        B, N, Mm, GW = h.B, d.N, d.Mm, d.GW
        M, st = B * N, h._stream()
        n_ws = self.ws.numel() - N * 6 * GW
        dx_out = []
        dh, ld_dh, dcn_in = dh_last, GW, None
        for step in (2, 1, 0):
            first = step == 0
            a = L.ConvLstmBwdArgs()
            a.y16, a.opre, a.cnew, a.cn = (sv[f"lstm_{k}{step}"].data_ptr() for k in ("y", "opre", "cnew", "cn"))
            a.cprev = None if first else sv[f"lstm_cn{step - 1}"].data_ptr()
            a.mr_g, a.mr_o = sv[f"lstm_mr_g{step}"].data_ptr(), sv[f"lstm_mr_o{step}"].data_ptr()
            a.ln_gamma, a.ln_beta = W["lstm_ln_gamma"].data_ptr(), W["lstm_ln_beta"].data_ptr()
            a.w_ci, a.w_cf, a.w_co = W["lstm_W_ci"].data_ptr(), W["lstm_W_cf"].data_ptr(), W["lstm_W_co"].data_ptr()
            a.dh, a.ld_dh, a.dcn_in = dh.data_ptr(), ld_dh, None if dcn_in is None else dcn_in.data_ptr()
            dcprev = self.dcn[step & 1]
            a.sums, a.dcnew, a.dy16 = self.sums.data_ptr(), self.dcnew.data_ptr(), self.dy16.data_ptr()
            a.dcprev_out = None if first else dcprev.data_ptr()
            a.dw_ci, a.dw_cf, a.dw_co = (self.g[k].data_ptr() for k in ("lstm_W_ci", "lstm_W_cf", "lstm_W_co"))
            a.dgamma, a.dbeta = self.g["lstm_ln_gamma"].data_ptr(), self.g["lstm_ln_beta"].data_ptr()
            a.ws_sample, a.ws_chan = self.ws.data_ptr(), self.ws[n_ws:].data_ptr()
            a.gw, a.m, a.rows_per_sample = GW, Mm, N
            for phase in (1, 2, 3):
                h._ck(lib.cmpc_convlstm_bwd(phase, C.byref(a), B, st), "convlstm_bwd")
            # the step's 1x1 conv: input gradients d[x | h_prev] = dy . K^T, weight gradient dK += [x | h_prev]^T dy
            x16 = sv[f"lstm_x{step}"]
            out = torch.empty(M, 2 * GW, dtype=torch.float32, device=h.device)
            h._gemm(self.dy16, 4 * GW, self.lstm_wT, GW if first else 2 * GW, out, group=(GW, Mm))
            h._ck(lib.cmpc_gemm_atb_f16(x16.data_ptr(), GW, GW, self.dy16.data_ptr(), 4 * GW, 4 * GW, M, self.g["lstm_w"].data_ptr(), 4 * GW, 0, st),
                  "gemm_atb")
            if not first:
                hp = sv[f"lstm_h{step - 1}"]
                h._ck(lib.cmpc_gemm_atb_f16(hp.data_ptr(), GW, GW, self.dy16.data_ptr(), 4 * GW, 4 * GW, M,
                                            self.g["lstm_w"][GW:].data_ptr(), 4 * GW, 0, st), "gemm_atb")
            dx_out.append(out[:, :GW])
            dh, ld_dh, dcn_in = out[:, GW:], 2 * GW, dcprev
        return dx_out[::-1]

    # ---- one level: fusion conv (:338-344) <- graph_conv (:359-374) <- affinity (:378-400) ------------------------------------------
    def bwd_level(self, i, dfus, ld_dfus):
        """dfus fp32 (row stride ld_dfus) = d loss / d fusion_<level i>.  Returns the four fp32 pieces whose SUM is d loss / d vis_la_sp
        (the MUTAN output): (via the fusion conv, graph residual, graph aggregation, affinity), each [M, LDC].
        Accumulates: parameter gradients of the level, d valid_lang (self.d_valid), d words_trans output (self.d_wt[level]),
        d relation gate (self.drgate)."""
        h, d, lib, W, sv, b = self.h, self.h.d, self.h.lib, self.h.Wt, self.h.saved.t, self.h.buf
        B, N, Mm, GW, C_, LDC, R, T = h.B, d.N, d.Mm, d.GW, d.C, d.LDC, d.R, d.T
        M, st, lvl = B * N, h._stream(), LEVELS[i]
        x16, y16, z16, u16, g16, w16, v16, affi = (sv[f"{nm}_{lvl}"] for nm in ("x16", "y16", "z16", "u16", "g16", "w16", "v16", "affi"))
        fus16 = b[f"fus16_{lvl}"]
        ck = h._ck
        atb = lambda a, lda, ac, c, ldc, cc, m, out, ldo: ck(lib.cmpc_gemm_atb_f16(a.data_ptr(), lda, ac, c.data_ptr(), ldc, cc, m, out.data_ptr(), ldo, 0, st), "gemm_atb")
        atb_b = lambda a, lda, ac, c, ldc, cc, out, ldo, obs: ck(lib.cmpc_gemm_atb_batched_f16(
            a.data_ptr(), lda, ac, c.data_ptr(), ldc, cc, N, B, out.data_ptr(), ldo, obs, st), "gemm_atb")
        # ---- fusion conv ----
        self.S.zero_()
        ck(lib.cmpc_relu_mask_f16(dfus.data_ptr(), ld_dfus, fus16.data_ptr(), GW, self.dpre16.data_ptr(), self.S.data_ptr(), B, N, rup(Mm, 8), st),
           "relu_mask")           # pad channels of the fusion map are exact zeros: the mask keeps them zero
        # (two outputs, two buffers: the GEMM stores whole tiles clipped at the row stride, so an output may not be a column slice)
        h._gemm(self.dpre16, Mm, self.fus_wT[lvl], C_, self.dxg)                          # d vis_la_sp (via the conv)
        h._gemm(self.dpre16, Mm, self.fus_wT[lvl][LDC:], C_, self.dgg)                   # d spa_graph
        atb(x16, LDC, C_, self.dpre16, GW, GW, M, self.g[f"fusion_w_{lvl}"], GW)
        atb(g16, LDC, C_ + 8, self.dpre16, GW, GW, M, self.g[f"fusion_w_{lvl}"][LDC:], GW)
        ck(lib.cmpc_small_atb_f32(b["valid32"].data_ptr(), R, 0, self.S.data_ptr(), GW, 0, self.g[f"fusion_lang_{lvl}"].data_ptr(), GW, 0,
                                  1, B, R, Mm, st), "small_atb")
        ck(lib.cmpc_small_atb_f32(self.ones.data_ptr(), 1, 0, self.S.data_ptr(), GW, 0, self.g[f"fusion_b_{lvl}"].data_ptr(), GW, 0, 1, B, 1, Mm, st),
           "small_atb")
        ck(lib.cmpc_small_linear_f32(self.S.data_ptr(), GW, 0, self.fus_lang_wT[lvl].data_ptr(), R, 0, None, 0, self.d_valid.data_ptr(), R, 0,
                                     1, B, Mm, R, 4, st), "small_linear")
        # ---- graph_conv: l2_normalize <- relu <- LN2 <- update conv <- relu <- (x + LN1(y)) ----
        self.lnsums.zero_()
        ck(lib.cmpc_ln_bwd_sums(self.dgg.data_ptr(), LDC, g16.data_ptr(), sv[f"rss_g_{lvl}"].data_ptr(), u16.data_ptr(), LDC,
                                sv[f"mr_u_{lvl}"].data_ptr(), W[f"gupdate_gamma_{lvl}"].data_ptr(), self.dln2.data_ptr(), LDC,
                                self.lnsums[0].data_ptr(), self.g[f"gupdate_gamma_{lvl}"].data_ptr(), self.g[f"gupdate_beta_{lvl}"].data_ptr(),
                                B, N, C_, st), "ln_bwd_sums")
        ck(lib.cmpc_ln_bwd_apply(self.dln2.data_ptr(), LDC, u16.data_ptr(), LDC, sv[f"mr_u_{lvl}"].data_ptr(), W[f"gupdate_gamma_{lvl}"].data_ptr(),
                                 self.lnsums[0].data_ptr(), self.du16.data_ptr(), self.g[f"gupd_b_{lvl}"].data_ptr(), B, N, C_, st), "ln_bwd_apply")
        h._gemm(self.du16, C_, self.gupd_wT[lvl], C_, self.dzl)
        atb(z16, LDC, C_, self.du16, LDC, C_, M, self.g[f"gupd_w_{lvl}"], LDC)
        ck(lib.cmpc_ln_bwd_sums(self.dzl.data_ptr(), LDC, z16.data_ptr(), None, y16.data_ptr(), LDC, sv[f"mr_y_{lvl}"].data_ptr(),
                                W[f"gfeat_gamma_{lvl}"].data_ptr(), self.dres.data_ptr(), LDC, self.lnsums[1].data_ptr(),
                                self.g[f"gfeat_gamma_{lvl}"].data_ptr(), self.g[f"gfeat_beta_{lvl}"].data_ptr(), B, N, C_, st), "ln_bwd_sums")
        ck(lib.cmpc_ln_bwd_apply(self.dres.data_ptr(), LDC, y16.data_ptr(), LDC, sv[f"mr_y_{lvl}"].data_ptr(), W[f"gfeat_gamma_{lvl}"].data_ptr(),
                                 self.lnsums[1].data_ptr(), self.dyg16.data_ptr(), None, B, N, C_, st), "ln_bwd_apply")
        # ---- y = W V^T x (:400, :362), through the rank-T factors: three skinny per-sample products and their GEMMs ----
        inv_vs = 1.0 / h.v_scale
        cast = lambda src, ldi, sc, dst, ldo, rows, cols: ck(lib.cmpc_scale_cast_f32_f16(src.data_ptr(), ldi, sc, dst.data_ptr(), ldo, rows, cols, st), "cast")
        self.t1.zero_(); self.t23.zero_()
        atb_b(self.dyg16, LDC, C_, w16, 32, 32, self.t1, 32, LDC * 32)                  # t1T[c, t] = sum_n dy[n, c] W[n, t]
        cast(self.t1, 32, inv_vs, self.t1_16, 64, B * LDC, 32)
        h._gemm(v16, 32, self.t1_16, C_, self.dagg, rows_per_sample=N, w_batch_stride=LDC * 64, w_rows=C_)          # d x = V (W^T dy)
        atb_b(v16, 32, 32, x16, LDC, C_, self.t23[0], LDC, 32 * LDC)                     # t2T[t, c] = sum_n V[n, t] x[n, c]
        atb_b(w16, 32, 32, self.dyg16, LDC, C_, self.t23[1], LDC, 32 * LDC)              # t3T[t, c] = sum_n W[n, t] dy[n, c]
        cast(self.t23[0], LDC, inv_vs, self.t23_16[0], LDC, B * 32, LDC)
        cast(self.t23[1], LDC, 1.0, self.t23_16[1], LDC, B * 32, LDC)
        h._gemm(self.dyg16, C_, self.t23_16[0], 32, self.dwm, rows_per_sample=N, w_batch_stride=32 * LDC, w_rows=32)  # d W = dy (x^T V)
        h._gemm(x16, C_, self.t23_16[1], 32, self.dvm, rows_per_sample=N, w_batch_stride=32 * LDC, w_rows=32)         # d V = x (dy^T W)
        # ---- the two softmaxes and the relation gate (:388-399) ----
        ck(lib.cmpc_affinity_bwd(w16.data_ptr(), v16.data_ptr(), self.dwm.data_ptr(), self.dvm.data_ptr(), affi.data_ptr(), b["rgate"].data_ptr(),
                                 h.v_scale, B, N, self.cs_ws.data_ptr(), self.draw16.data_ptr(), self.drgate.data_ptr(), st), "affinity_bwd")
        # ---- affi = x . Gt^T (:384 re-associated): d x = d raw . Gt ;  d Gt = d raw^T [x | 1] ----
        gt16 = b["gt16"][i]
        ck(lib.cmpc_transpose_gt_f16(gt16.data_ptr(), LDC, B, T, LDC, self.gtT16.data_ptr(), st), "transpose_gt")
        h._gemm(self.draw16, 32, self.gtT16, C_, self.daff, rows_per_sample=N, w_batch_stride=LDC * 64, w_rows=C_)
        self.dgt.zero_()
        # rows t >= T of a sample's [32, LDC] block are exact zeros (d raw is zero there), so the blocks may overlap at stride T
        atb_b(self.draw16, 32, 32, x16, LDC, C_ + 8, self.dgt, LDC, T * LDC)
        cast(self.dgt, LDC, 1.0, self.dgt16, LDC, B * T, LDC)
        # Gt = words_trans(words) . [DW2 ; b2]^T: gradients of spa_graph_trans2 and of the words_trans output
        wt16 = b["wt16"][:, i * R:]
        h._gemm(self.dgt16, C_ + 8, self.gt_wT[lvl], R, self.d_wt[i], m=B * T)
        atb(self.dgt16, LDC, C_ + 8, wt16, b["wt16"].stride(0), R, B * T, self.g[f"gt_w_{lvl}"], d.LDR)
        return self.dxg, self.dres, self.dagg, self.daff

    # ---- MUTAN fusion (:295-328) and the lateral conv + l2_normalize in front of it (:108-113) ----------------------------------
    def bwd_mutan(self, i, pieces, part=None):
        """pieces: up to four fp32 [M, LDC] maps whose sum is d loss / d vis_la_sp of level i (bwd_level).  Accumulates the gradients
        of the five vis_trans heads and of the lateral conv of the level, and d loss / d tanh(lang_trans) (self.d_lang).
        part 'a' stops after the MUTAN backward kernel (d_lang and the vis_trans bias gradients are final), part 'b' runs the rest
        (input gradient, the two weight-gradient GEMMs, the lateral conv); None = both."""
        if part != "b":
            self._bwd_mutan_a(i, pieces)
        if part != "a":
            self._bwd_mutan_b(i)

    def _bwd_mutan_a(self, i, pieces):
        h, d, lib, W, sv, b = self.h, self.h.d, self.h.lib, self.h.Wt, self.h.saved.t, self.h.buf
        B, N, C_, LDC = h.B, d.N, d.C, d.LDC
        M, st, lvl, ck = B * N, h._stream(), LEVELS[i], h._ck
        x16, xlat16 = sv[f"x16_{lvl}"], sv[f"xlat16_{lvl}"]
        ss_lat, ss_mut = b["rowss"][2 * i], b["rowss"][2 * i + 1]
        ps = [p.data_ptr() for p in pieces] + [None] * (4 - len(pieces))
        ds = self.dln2
        ck(lib.cmpc_mutan_out_bwd(ps[0], ps[1], ps[2], ps[3], LDC, x16.data_ptr(), LDC, ss_mut.data_ptr(), ds.data_ptr(), LDC, M, C_, st),
           "mutan_out_bwd")
        ma = L.MutanArgs()
        ma.a, ma.lda, ma.k = xlat16.data_ptr(), LDC, C_ + 8
        ma.a_row_sumsq = ss_lat.data_ptr()
        ma.w, ma.ldw = W[f"mutan_w_{lvl}"].data_ptr(), LDC
        ma.m, ma.c, ma.rows_per_sample = M, C_, N
        ma.bias, ma.ld_bias = W[f"mutan_b_{lvl}"].data_ptr(), LDC
        ma.lang, ma.ld_lang, ma.lang_batch_stride = b["lang"][:, i * 5 * C_:].data_ptr(), C_, 15 * C_
        ma.out, ma.ldo = self.dpre_m16.data_ptr(), self.CHP
        ck(lib.cmpc_mutan_bwd_f16(C.byref(ma), ds.data_ptr(), LDC, self.d_lang[:, i * 5 * C_:].data_ptr(), 15 * C_,
                                  self.g[f"mutan_b_{lvl}"].data_ptr(), st), "mutan_bwd")

    def _bwd_mutan_b(self, i):
        h, d, lib, sv, b = self.h, self.h.d, self.h.lib, self.h.saved.t, self.h.buf
        B, N, C_, LDC = h.B, d.N, d.C, d.LDC
        M, st, lvl, ck = B * N, h._stream(), LEVELS[i], h._ck
        xlat16, cin16 = sv[f"xlat16_{lvl}"], sv[f"cin_{lvl}"]
        ss_lat = b["rowss"][2 * i]
        K5 = d.CH * 240
        G = self.dzl
        h._gemm(self.dpre_m16, K5, self.mutan_wT[lvl], C_, G)                         # (d pre * rsc) . Wv^T
        ck(lib.cmpc_gemm_atb_f16(xlat16.data_ptr(), LDC, C_ + 8, self.dpre_m16.data_ptr(), self.CHP, K5, M, self.g[f"mutan_w_{lvl}"].data_ptr(),
                                 self.CHP, 0, st), "gemm_atb")
        dxlat16 = self.du16
        ck(lib.cmpc_lateral_bwd(G.data_ptr(), LDC, xlat16.data_ptr(), LDC, ss_lat.data_ptr(), dxlat16.data_ptr(), self.g[f"lat_b_{lvl}"].data_ptr(),
                                B, N, C_, st), "lateral_bwd")
        kin = d.cin[lvl]
        ck(lib.cmpc_gemm_atb_f16(cin16.data_ptr(), cin16.stride(0), kin, dxlat16.data_ptr(), LDC, C_, M, self.g[f"lat_w_{lvl}"].data_ptr(), LDC, 0, st),
           "gemm_atb")

    def bwd_lang_trans(self, levels=None):
        """tanh(lang_trans(valid_lang)) of the MUTAN heads (:303-306) of `levels` (default: all three): consumes self.d_lang, accumulates
        self.d_valid and the lang_trans gradients.  A level's slice of d_lang is final once its bwd_mutan has run, so the pass calls this
        level by level (the tanh backward is re-run over the whole [B, 15 C] buffer each time: it is tiny)."""
        h, d, lib, b, ck = self.h, self.h.d, self.h.lib, self.h.buf, self.h._ck
        B, C_, R, st = h.B, d.C, d.R, h._stream()
        ck(lib.cmpc_act_bwd_f32(self.d_lang.data_ptr(), b["lang"].data_ptr(), self.dpl.data_ptr(), B * 15 * C_, 2, st), "act_bwd")
        for i, lvl in enumerate(LEVELS):
            if levels is not None and i not in levels:
                continue
            dpl = self.dpl[:, i * 5 * C_:]
            ck(lib.cmpc_small_linear_f32(dpl.data_ptr(), 15 * C_, 0, self.ltrans_wT[lvl].data_ptr(), R, 0, None, 0, self.d_valid.data_ptr(), R, 0,
                                         1, B, 5 * C_, R, 4, st), "small_linear")
            ck(lib.cmpc_small_atb_f32(b["valid32"].data_ptr(), R, 0, dpl.data_ptr(), 15 * C_, 0, self.g[f"ltrans_w_{lvl}"].data_ptr(), 5 * C_, 0,
                                      1, B, R, 5 * C_, st), "small_atb")
            ck(lib.cmpc_small_atb_f32(self.ones.data_ptr(), 1, 0, dpl.data_ptr(), 15 * C_, 0, self.g[f"ltrans_b_{lvl}"].data_ptr(), 5 * C_, 0,
                                      1, B, 1, 5 * C_, st), "small_atb")

    # ---- language side -------------------------------------------------------------------------------------------------
    def bwd_language(self):
        """word-type parser, valid_lang / nec_lang, relation weights, words_trans, l2_normalize of the LSTM outputs (:159-192, :347-357,
        :378): consumes self.d_valid, self.d_nec, self.drgate, self.d_wt; returns d loss / d lstm_outputs [B*T, R] (the boundary of the
        head: what an upstream word LSTM would continue from)."""
        h, d, lib, b, ck, sv = self.h, self.h.d, self.h.lib, self.h.buf, self.h._ck, self.h.saved.t
        B, T, R, HID, HIDP, st = h.B, d.T, d.R, d.HID, d.HIDP, h._stream()
        BT = B * T
        w32 = b["words32"]
        ck(lib.cmpc_lang_bwd(w32.data_ptr(), b["parse"].data_ptr(), b["mask"].data_ptr(), b["valid32"].data_ptr(), b["nec32"].data_ptr(),
                             self.d_valid.data_ptr(), self.d_nec.data_ptr(), self.drgate.data_ptr(), B, T, R, d.C, self.dwords.data_ptr(),
                             self.dlogit.data_ptr(), st), "lang_bwd")
        atb = lambda a, lda, c, ldc, out, ldo, ni, nj: ck(lib.cmpc_small_atb_f32(
            a.data_ptr(), lda, 0, c.data_ptr(), ldc, 0, out.data_ptr(), ldo, 0, 1, BT, ni, nj, st), "small_atb")
        sl = lambda x, ldx, w, ldw, out, ldo, k, n, act: ck(lib.cmpc_small_linear_f32(
            x.data_ptr(), ldx, 0, w.data_ptr(), ldw, 0, None, 0, out.data_ptr(), ldo, 0, 1, BT, k, n, act, st), "small_linear")
        # words_parse_2, relu, words_parse_1
        atb(b["hidden"], HIDP, self.dlogit, 4, self.g["parse2_w"], 4, HID, 4)
        atb(self.ones, 1, self.dlogit, 4, self.g["parse2_b"], 4, 1, 4)
        sl(self.dlogit, 4, self.parse2_wT, HID, self.dhid, HIDP, 4, HID, 0)
        ck(lib.cmpc_relu_bwd_f32(self.dhid.data_ptr(), b["hidden"].data_ptr(), self.dhid.data_ptr(), BT, HID, HIDP, st), "relu_bwd")
        atb(w32, R, self.dhid, HIDP, self.g["parse1_w"], HID, R, HID)
        atb(self.ones, 1, self.dhid, HIDP, self.g["parse1_b"], HID, 1, HID)
        sl(self.dhid, HIDP, self.parse1_wT, R, self.dwords, R, HID, R, 4)
        # words_trans of the three levels
        for i, lvl in enumerate(LEVELS):
            dwt = self.d_wt[i]
            atb(w32, R, dwt, d.LDR, self.g[f"wtrans_w_{lvl}"], R, R, R)
            atb(self.ones, 1, dwt, d.LDR, self.g[f"wtrans_b_{lvl}"], R, 1, R)
            sl(dwt, d.LDR, self.wtrans_wT[i], R, self.dwords, R, R, R, 4)
        ck(lib.cmpc_l2norm_bwd_f32(self.dwords.data_ptr(), w32.data_ptr(), sv["lstm_outputs"].data_ptr(), BT, R, self.d_lstm.data_ptr(), st),
           "l2norm_bwd")
        return self.d_lstm

    # ---- the whole pass ----------------------------------------------------------------------------------------------------
    @on_device
    def backward(self, out, target_fine, weights=(0.7, 0.1, 0.1, 0.1), on_bucket=None):
        """Gradient of cls_loss_all = 0.7 CE(up) + 0.1 CE(up_c5) + 0.1 CE(up_c4) + 0.1 CE(up_c3) (CMPC_model.py:439-445; CE = mean over
        the batch of the per-sample sum over pixels, util/loss.py:6-16) w.r.t. every parameter of the head, after a training-mode
        forward(..., aux=True) of the head whose outputs are `out`.  The L2 regulariser (:446) is a term of the optimizer step.
        on_bucket(name) is called each time the gradient buffers of bucket `name` (self.BUCKETS, self.bucket_view) are final.
        Returns d loss / d lstm_outputs."""
        for bname in self.backward_stages(out, target_fine, weights):
            if on_bucket is not None:
                on_bucket(bname)
        return self.d_lstm

    def backward_stages(self, out, target_fine, weights=(0.7, 0.1, 0.1, 0.1)):
        """The backward pass as a generator: runs up to the point where the next gradient bucket is final and yields its name, in
        the order of self.BUCKETS (reverse order of the forward: ConvLSTM / score first, the language side last)."""
        h, d, b = self.h, self.h.d, self.h.buf
        GW = d.GW
        self.zero_grads()
        for t in (self.d_valid, self.d_lang, self.drgate):
            t.zero_()
        M = h.B * d.N
        h16 = h.saved.t["lstm_h2"]
        dF = torch.zeros(M, GW, dtype=torch.float32, device=h.device)
        self.bwd_score(out["up"], target_fine, weights[0], h16, "score", dF)
        dxs = self.bwd_convlstm(dF)
        yield "fuse"
        d1 = self.bwd_exchange_round(1, dxs, 2 * GW)
        aux = {}
        for lvl, wgt in zip(LEVELS, weights[1:]):                      # (c5, c4, c3) <-> weights[1:] = (c5, c4, c3)
            aux[lvl] = torch.zeros(M, GW, dtype=torch.float32, device=h.device)
            self.bwd_score(out[f"up_{lvl}"], target_fine, wgt, b[f"fus16_{lvl}"], f"score_{lvl}", aux[lvl])
        d0 = self.bwd_exchange_round(0, d1, GW, extra=[aux["c3"], aux["c4"], aux["c5"]])
        self.d_nec.zero_()
        yield "exchange"
        # The language side of the exchange modules (54 launches of B-row products, ~1.1 ms of pure latency at batch 16: key fold,
        # lang_query, gv_lang) needs only du / dz of the two rounds and is consumed by bwd_language at the very end: two of its six
        # slots run on a side stream underneath each level's graph-stage kernels.  Fork and join sit inside ONE stage, so a captured
        # segment (train.py: one CUDA graph per reduce group, or per stage) always contains both.
        main = torch.cuda.current_stream(h.device)
        if self._side is None:
            self._side = torch.cuda.Stream(h.device)
        for i, lvl in enumerate(LEVELS):
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self._side):
                self._side.wait_event(ev)
                self.bwd_exchange_language(slots=(2 * i, 2 * i + 1))
            pieces = self.bwd_level(i, d0[{"c3": 0, "c4": 1, "c5": 2}[lvl]], GW)
            main.wait_stream(self._side)
            yield lvl + "_graph"
            self.bwd_mutan(i, pieces, part="a")
            self.bwd_lang_trans((i,))
            yield lvl + "_ltrans"
            self.bwd_mutan(i, pieces, part="b")
            yield lvl + "_mutan"
        self.bwd_language()
        yield "language"

    # ---- text-guided exchange ------------------------------------------------------------------------------------------
    def bwd_exchange_round(self, rnd, douts, ld_dout, extra=None):
        """douts: three fp32 maps (row stride ld_dout) = d loss / d (e3, e4, e5) of round `rnd` (outputs); extra: optional three
        fp32 [M, GW] maps added to the gradients of the round's INPUTS (other consumers of those maps, e.g. the aux score heads).
        Returns three fp32 [M, GW] maps = d loss / d (inputs of the round); parameter gradients are accumulated."""
        h, d, lib, W, sv, b = self.h, self.h.d, self.h.lib, self.h.Wt, self.h.saved.t, self.h.buf
        B, N, Mm, GW = h.B, d.N, d.Mm, d.GW
        M, st = B * N, h._stream()
        ins = sv[f"exg{rnd}_in"]
        outs = [b[o] for o in (("e3", "e4", "e5") if rnd == 0 else ("g3", "g4", "g5"))]
        g1, g2, gv, pool, pstats = (sv[f"exg{rnd}_{k}"] for k in ("gate1", "gate2", "gv", "pool", "pstats"))
        self.colsum.zero_()
        for mi in range(3):
            x = EXG[rnd * 3 + mi]
            h._ck(lib.cmpc_exg_bwd_rows(douts[mi].data_ptr(), ld_dout, outs[mi].data_ptr(), sv[f"exg_rss_{x}"].data_ptr(),
                                        sv[f"exg_se1_{x}"].data_ptr(), sv[f"exg_se2_{x}"].data_ptr(), g1[:, mi].data_ptr(), g2[:, mi].data_ptr(),
                                        3 * GW, GW, self.ds[mi].data_ptr(), self.dp16[mi][0].data_ptr(), self.dp16[mi][1].data_ptr(),
                                        self.colsum[:, mi].data_ptr(), 3 * 4 * GW, B, N, GW, st), "exg_bwd_rows")
        slot0 = rnd * 3
        gargs = (self.colsum.data_ptr(), g1.data_ptr(), g2.data_ptr(), gv.data_ptr(), pool.data_ptr(),
                 b["gvl"][:, slot0 * GW:].data_ptr(), 6 * GW, GW, W["wg"][slot0:].data_ptr(), W["wf1"][slot0:].data_ptr(),
                 W["wf2"][slot0:].data_ptr(), Mm * Mm, B, 3, Mm, GW, self.dpre1[rnd].data_ptr(), self.dpre2[rnd].data_ptr(),
                 self.dz[rnd].data_ptr(), self.dpool.data_ptr())
        if h.gv_norm == "sample":
            h._ck(lib.cmpc_gv_gates_bwd(*gargs, st), "gv_gates_bwd")
        else:
            # batch-coupled l2_normalize (:241): d z = (d gv - gv * sum_batch(gv . d gv)) / |z|_batch
            gv_ss = sv[f"exg{rnd}_gv_ss"]
            self.gv_dot.zero_()
            h._ck(lib.cmpc_gv_gates_bwd_batch(*gargs, 1, gv_ss.data_ptr(), self.gv_dot.data_ptr(), st), "gv_gates_bwd")
            if h.gv_allreduce is not None:
                h.gv_allreduce(self.gv_dot)
            h._ck(lib.cmpc_gv_gates_bwd_batch(*gargs, 2, gv_ss.data_ptr(), self.gv_dot.data_ptr(), st), "gv_gates_bwd")
        # parameter gradients of the per-sample maps: sums over the batch of outer products
        def atb(a, lda, azs, c, ldc, czs, out, ldo, ozs, nz, ni, nj):
            h._ck(lib.cmpc_small_atb_f32(a.data_ptr(), lda, azs, c.data_ptr(), ldc, czs, out.data_ptr(), ldo, ozs, nz, B, ni, nj, st), "small_atb")
        atb(gv, 3 * GW, GW, self.dpre1[rnd], 3 * GW, GW, self.g["wf1"][slot0:], Mm, Mm * Mm, 3, Mm, Mm)
        atb(gv, 3 * GW, GW, self.dpre2[rnd], 3 * GW, GW, self.g["wf2"][slot0:], Mm, Mm * Mm, 3, Mm, Mm)
        atb(pool, 3 * GW, GW, self.dz[rnd], 3 * GW, GW, self.g["wg"][slot0:], Mm, Mm * Mm, 3, Mm, Mm)
        atb(self.ones, 1, 0, self.dpre1[rnd], 3 * GW, GW, self.g["bf1"][slot0:], Mm, Mm, 3, 1, Mm)
        atb(self.ones, 1, 0, self.dpre2[rnd], 3 * GW, GW, self.g["bf2"][slot0:], Mm, Mm, 3, 1, Mm)
        for mi in range(3):
            x = EXG[rnd * 3 + mi]
            atb(self.ones, 1, 0, self.colsum[:, mi, 2], 12 * GW, 0, self.g[f"se_b_{x}_f1"], GW, 0, 1, 1, GW)
            atb(self.ones, 1, 0, self.colsum[:, mi, 3], 12 * GW, 0, self.g[f"se_b_{x}_f2"], GW, 0, 1, 1, GW)
        # trans_feat convs: weight gradients (source map ^T dP) and the input gradients, one K-concatenated GEMM per source map
        triples = ((0, 1, 2), (1, 0, 2), (2, 0, 1))                        # module mi: (feat, fa, fb) as indices into ins
        for mi in range(3):
            x = EXG[rnd * 3 + mi]
            for j, f in enumerate(("_f1", "_f2")):
                src = ins[triples[mi][1 + j]]
                h._ck(lib.cmpc_gemm_atb_f16(src.data_ptr(), GW, GW, self.dp16[mi][j].data_ptr(), GW, GW, M, self.g[f"se_w_{x}{f}"].data_ptr(),
                                            GW, 0, st), "gemm_atb")
        self.du[rnd].zero_()
        res = []
        for src in range(3):
            (ma, fa), (mb, fb) = self.exg_consumers[src]
            h._gemm(self.dp16[ma][0 if fa == "_f1" else 1], Mm, self.exg_wT[(rnd, src)], GW, self.dgemm,
                    a2=self.dp16[mb][0 if fb == "_f1" else 1], k2=Mm, group=(GW, Mm))
            out = torch.empty(M, GW, dtype=torch.float32, device=h.device)
            ex = None if extra is None else extra[src]
            h._ck(lib.cmpc_pool_bwd_rows(ins[src].data_ptr(), GW, b["u"][:, (slot0 + src) * GW:].data_ptr(), 6 * GW, pool[:, src].data_ptr(),
                                         self.dpool[:, src].data_ptr(), 3 * GW, pstats[:, src].data_ptr(), 6, 1.0 / (Mm ** 0.5),
                                         self.ds[src].data_ptr(), self.dgemm.data_ptr(), GW, None if ex is None else ex.data_ptr(), GW,
                                         out.data_ptr(), self.du[rnd][:, src].data_ptr(), 3 * GW, B, N, GW, st), "pool_bwd_rows")
            res.append(out)
        return res

    def bwd_exchange_language(self, slots=None):
        """Everything between the exchange rounds and nec_lang (:223, :239 and the key conv folded into the query): consumes the du /
        dz of both rounds; accumulates d nec_lang [B, R] (self.d_nec) and the gradients of lang_query, spa_graph_key, gv_lang."""
        h, d, lib, b = self.h, self.h.d, self.h.lib, self.h.buf
        B, Mm, GW, R, st = h.B, d.Mm, d.GW, d.R, h._stream()
        nec = b["nec32"]
        if slots is None:                       # the whole language side at once (stand-alone use): d_nec starts from zero
            self.d_nec.zero_()
            slots = range(6)
        for slot in slots:
            rnd, mi = divmod(slot, 3)
            du, dz, q = self.du[rnd][:, mi], self.dz[rnd][:, mi], b["q"][:, slot * GW:]
            sl = lambda x, ldx, w, ldw, out, ldo, k, n, act: h._ck(lib.cmpc_small_linear_f32(
                x.data_ptr(), ldx, 0, w.data_ptr(), ldw, 0, None, 0, out.data_ptr(), ldo, 0, 1, B, k, n, act, st), "small_linear")
            atb = lambda a, lda, c, ldc, out, ldo, ni, nj: h._ck(lib.cmpc_small_atb_f32(
                a.data_ptr(), lda, 0, c.data_ptr(), ldc, 0, out.data_ptr(), ldo, 0, 1, B, ni, nj, st), "small_atb")
            # u = q . keyT  (keyT[o, cin] = Wk[cin, o]):  dq = du . Wk ;  d keyT[o, cin] += q^T du
            sl(du, 3 * GW, self.key_w[slot], Mm, self.dq, GW, Mm, Mm, 0)
            atb(q, 6 * GW, du, 3 * GW, self.g["key"][slot], Mm, Mm, Mm)
            # q = nec Wq + bq ;  gvl = nec Wgl + bg
            sl(self.dq, GW, self.q_wT[slot], R, self.d_nec, R, Mm, R, 4)
            sl(dz, 3 * GW, self.gvl_wT[slot], R, self.d_nec, R, Mm, R, 4)
            atb(nec, R, self.dq, GW, self.g["q_w"][slot], Mm, R, Mm)
            atb(nec, R, dz, 3 * GW, self.g["gvl_w"][slot], Mm, R, Mm)
            atb(self.ones, 1, self.dq, GW, self.g["q_b"][slot], Mm, 1, Mm)
            atb(self.ones, 1, dz, 3 * GW, self.g["gvl_b"][slot], Mm, 1, Mm)
        return self.d_nec

    # ---- packed gradient buffers -> TF variable names / shapes ---------------------------------------------------------------
    @on_device
    def grads_tf(self, into: Dict[str, torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Packed gradient buffers -> the TF variable names / shapes.  With `into` (name -> contiguous tensor of the TF shape, e.g. the
        views of the trainer's flat gradient buffer) every tensor is written in place with one strided copy; otherwise new tensors."""
        d, dev = self.h.d, self.h.device
        Mm, GW, R, C_, LDC, HID, CH = d.Mm, d.GW, d.R, d.C, d.LDC, d.HID, d.CH
        g, out = self.g, {}

        def dst(name, *shape):
            """the destination of variable `name`, viewed as `shape` (same number of elements as its TF shape)"""
            if into is not None:
                t = into[name]
            else:
                t = torch.empty(tuple(self.h.params[name].shape), dtype=torch.float32, device=dev)
            out[name] = t
            return t.view(*shape)

        k = dst("rnn/conv_lstm_cell/kernel", 2, Mm, 4, Mm)
        k.copy_(g["lstm_w"].view(2, GW, 4, GW)[:, :Mm, :, :Mm])
        for nm in ("W_ci", "W_cf", "W_co"):
            dst(f"rnn/conv_lstm_cell/{nm}", d.N, Mm).copy_(g[f"lstm_{nm}"][:, :Mm])
        for i in range(5):
            nm = "LayerNorm" if i == 0 else f"LayerNorm_{i}"
            dst(f"rnn/conv_lstm_cell/{nm}/gamma", Mm).copy_(g["lstm_ln_gamma"][i, :Mm])
            dst(f"rnn/conv_lstm_cell/{nm}/beta", Mm).copy_(g["lstm_ln_beta"][i, :Mm])
        for name in self.score_names:
            dst(name + "/DW", 9, Mm).copy_(g[name + "_w9"][:9, :Mm])
            dst(name + "/biases", 1).copy_(g[name + "_b"])
        for slot, x in enumerate(EXG):
            for j, f in enumerate(("_f1", "_f2")):
                dst(f"trans_feat_{x}{f}/DW", Mm, Mm).copy_(g[f"se_w_{x}{f}"][:Mm, :Mm])
                dst(f"trans_feat_{x}{f}/biases", Mm).copy_(g[f"se_b_{x}{f}"][:Mm])
                dst(f"lang_feat_{x}{f}/DW", Mm, Mm).copy_(g["wf1" if j == 0 else "wf2"][slot])
                dst(f"lang_feat_{x}{f}/biases", Mm).copy_(g["bf1" if j == 0 else "bf2"][slot])
            dst(f"spa_graph_key_{x}gv_f1/DW", Mm, Mm).copy_(g["key"][slot].t())
            dst(f"spa_graph_key_{x}gv_f1/biases", Mm).zero_()                              # the softmax is shift invariant
            dst(f"lang_query_{x}gv_f1/DW", R, Mm).copy_(g["q_w"][slot])
            dst(f"lang_query_{x}gv_f1/biases", Mm).copy_(g["q_b"][slot])
            gv = dst(f"gv_lang_{x}gv_f1/DW", Mm + R, Mm)
            gv[:Mm].copy_(g["wg"][slot]); gv[Mm:].copy_(g["gvl_w"][slot])
            dst(f"gv_lang_{x}gv_f1/biases", Mm).copy_(g["gvl_b"][slot])
        for lvl in LEVELS:
            fw = g[f"fusion_w_{lvl}"]
            fd = dst(f"fusion_{lvl}/DW", 2 * C_ + R + 8, Mm)
            fd[:C_].copy_(fw[:C_, :Mm]); fd[C_:2 * C_].copy_(fw[LDC:LDC + C_, :Mm])
            fd[2 * C_:2 * C_ + R].copy_(g[f"fusion_lang_{lvl}"][:, :Mm]); fd[2 * C_ + R:].copy_(fw[LDC + C_:LDC + C_ + 8, :Mm])
            dst(f"fusion_{lvl}/biases", Mm).copy_(g[f"fusion_b_{lvl}"][:Mm])
            dst(f"gconv_update_spa_graph_{lvl}/DW", C_, C_).copy_(g[f"gupd_w_{lvl}"][:C_, :C_])
            dst(f"gconv_update_spa_graph_{lvl}/biases", C_).copy_(g[f"gupd_b_{lvl}"][:C_])
            for ln, nm in (("gconv_feat_ln_spa_graph", "gfeat"), ("gconv_update_ln_spa_graph", "gupdate")):
                dst(f"{ln}_{lvl}/gamma", C_).copy_(g[f"{nm}_gamma_{lvl}"][:C_])
                dst(f"{ln}_{lvl}/beta", C_).copy_(g[f"{nm}_beta_{lvl}"][:C_])
            dst(f"spa_graph_trans2_{lvl}/DW", C_, R).copy_(g[f"gt_w_{lvl}"][:C_, :R])
            dst(f"spa_graph_trans2_{lvl}/biases", R).copy_(g[f"gt_w_{lvl}"][C_, :R])
            mw = g[f"mutan_w_{lvl}"][:C_ + 8, :CH * 240].view(C_ + 8, CH, 5, 48)
            tail = C_ - (CH - 1) * 48                                              # valid channels of the last 48-chunk
            for k in range(5):
                vd = dst(f"vis_trans_{lvl}_head{k + 1}/DW", C_ + 8, C_)
                if CH > 1:
                    vd[:, :(CH - 1) * 48].view(C_ + 8, CH - 1, 48).copy_(mw[:, :CH - 1, k, :])
                vd[:, (CH - 1) * 48:].copy_(mw[:, CH - 1, k, :tail])
                dst(f"vis_trans_{lvl}_head{k + 1}/biases", C_).copy_(g[f"mutan_b_{lvl}"][k, :C_])
                dst(f"lang_trans_{lvl}_head{k + 1}/DW", R, C_).copy_(g[f"ltrans_w_{lvl}"][:, k * C_:(k + 1) * C_])
                dst(f"lang_trans_{lvl}_head{k + 1}/biases", C_).copy_(g[f"ltrans_b_{lvl}"][k * C_:(k + 1) * C_])
            dst(f"{lvl}_lateral/DW", d.cin[lvl], C_).copy_(g[f"lat_w_{lvl}"][:, :C_])
            dst(f"{lvl}_lateral/biases", C_).copy_(g[f"lat_b_{lvl}"][:C_])
            dst(f"words_trans_{lvl}/DW", R, R).copy_(g[f"wtrans_w_{lvl}"])
            dst(f"words_trans_{lvl}/biases", R).copy_(g[f"wtrans_b_{lvl}"])
        dst("words_parse_1/DW", R, HID).copy_(g["parse1_w"])
        dst("words_parse_1/biases", HID).copy_(g["parse1_b"])
        dst("words_parse_2/DW", HID, 4).copy_(g["parse2_w"])
        dst("words_parse_2/biases", 4).copy_(g["parse2_b"])
        return out
