"""Backward pass of the head (what TF autodiff builds at CMPC_model.py:461), stage by stage, on the
device.  Mirrors head.py: every forward stage `_st_*` has a `bwd_*` here that consumes the activations the forward kept in
a `Saved` store and produces (a) the gradient w.r.t. the stage's inputs and (b) gradients of its parameters, accumulated
in fp32 buffers laid out like the packed weights; `grads_tf()` maps them back to the TF variable names / shapes.

Status: the tail of the graph is done -- loss + upsample + score conv (:138-142, :439-445) and the ConvLSTM (:287-290,
util/cell.py); the remaining stages (exchange, fusion, graph, affinity, MUTAN, language) follow the same pattern.
torch is used for buffers and for re-laying-out small parameter tensors; all arithmetic is in libcmpc_b200.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _lib as L
from .weights import rup


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
        d, dev = head.d, head.device
        P = {k: v.to(device=dev, dtype=torch.float32) for k, v in head.params.items() if k.startswith(("rnn/", "score"))}
        Mm, GW, N = d.Mm, d.GW, d.N
        f32 = dict(dtype=torch.float32, device=dev)
        # operands of the input-gradient GEMMs: the TF kernels [Cin, Cout] as fp16 "weights" [n_out = cin, k = cout]
        kern = P["rnn/conv_lstm_cell/kernel"][0, 0]                      # [2Mm, 4Mm]
        wT = torch.zeros(2 * GW, 4 * GW, **f32)
        for grp in range(2):
            for gate in range(4):
                wT[grp * GW:grp * GW + Mm, gate * GW:gate * GW + Mm] = kern[grp * Mm:(grp + 1) * Mm, gate * Mm:(gate + 1) * Mm]
        self.lstm_wT = wT.half().contiguous()
        self.score_wT = {}
        for name in ["score"] + [f"score_{l}" for l in ("c5", "c4", "c3")]:
            w = torch.zeros(GW, 64, **f32)
            w[:Mm, :9] = P[name + "/DW"][:, :, :, 0].reshape(9, Mm).t()
            self.score_wT[name] = w.half().contiguous()
        # parameter gradients (fp32, packed layouts)
        self.g = {
            "lstm_w": torch.zeros(2 * GW, 4 * GW, **f32),                 # rows: [x | h] input channel, cols: gate * GW + cout
            "lstm_W_ci": torch.zeros(N, GW, **f32), "lstm_W_cf": torch.zeros(N, GW, **f32), "lstm_W_co": torch.zeros(N, GW, **f32),
            "lstm_ln_gamma": torch.zeros(5, GW, **f32), "lstm_ln_beta": torch.zeros(5, GW, **f32),
        }
        for name in self.score_wT:
            self.g[name + "_w9"] = torch.zeros(16, GW, **f32)
            self.g[name + "_b"] = torch.zeros(1, **f32)
        M = head.B * N
        self.ws = torch.zeros(head.lib.cmpc_convlstm_bwd_workspace_floats(head.B, N, GW), **f32)
        self.sums = torch.zeros(head.B, 10, **f32)
        self.dy16 = torch.zeros(M, 4 * GW, dtype=torch.float16, device=dev)
        self.dcnew = torch.zeros(M, GW, **f32)
        self.dcn = [torch.zeros(M, GW, **f32) for _ in range(2)]
        self.dxh = torch.zeros(M, 2 * GW, **f32)
        self.dpred = torch.zeros(head.B, d.h, d.w, **f32)
        self.d9 = torch.zeros(M, 64, dtype=torch.float16, device=dev)

    def zero_grads(self):
        for v in self.g.values():
            v.zero_()

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

    # ---- packed gradient buffers -> TF variable names / shapes ---------------------------------------------------------------
    def grads_tf(self) -> Dict[str, torch.Tensor]:
        d = self.h.d
        Mm, GW = d.Mm, d.GW
        g, out = self.g, {}
        k = torch.zeros(2 * Mm, 4 * Mm, device=self.h.device)
        for grp in range(2):
            for gate in range(4):
                k[grp * Mm:(grp + 1) * Mm, gate * Mm:(gate + 1) * Mm] = g["lstm_w"][grp * GW:grp * GW + Mm, gate * GW:gate * GW + Mm]
        out["rnn/conv_lstm_cell/kernel"] = k.reshape(1, 1, 2 * Mm, 4 * Mm)
        for nm in ("W_ci", "W_cf", "W_co"):
            out[f"rnn/conv_lstm_cell/{nm}"] = g[f"lstm_{nm}"][:, :Mm].reshape(d.h, d.w, Mm).clone()
        for i in range(5):
            nm = "LayerNorm" if i == 0 else f"LayerNorm_{i}"
            out[f"rnn/conv_lstm_cell/{nm}/gamma"] = g["lstm_ln_gamma"][i, :Mm].clone()
            out[f"rnn/conv_lstm_cell/{nm}/beta"] = g["lstm_ln_beta"][i, :Mm].clone()
        for name in self.score_wT:
            out[name + "/DW"] = g[name + "_w9"][:9, :Mm].reshape(3, 3, Mm, 1).clone()
            out[name + "/biases"] = g[name + "_b"].clone()
        return out
