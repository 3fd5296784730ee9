"""Packs the reference's TF variables (names/shapes of SURVEY App. B, scope text_objseg/) into the device
layouts the sm_100a kernels consume.  Init-time plumbing only (torch is used for device memory and copies).

Layouts (fp16 GEMM operands are [N_out, K] row-major, K contiguous = "K-major B operand"):
  * 1x1 conv DW[1,1,Cin,Cout]      -> W16[Cout, Kpad]            (transpose; K padded to a multiple of 64 with zeros)
  * five MUTAN heads               -> W16[chunks*240, Kpad]      rows (chunk j, head k, cc), channel c = 48 j + cc
  * per-module / per-gate outputs  -> "grouped" rows g*GW + c, GW = multiple of 256 >= mlp_dim (pads are zero rows)
  * K-concatenated inputs (fusion: [vis_la_sp | spa_graph+spatial], ConvLSTM: [x | h]) -> K segments each padded to 64
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch

LEVELS = ("c5", "c4", "c3")                        # build order, CMPC_model.py:120-125
EXG = ("c3", "c4", "c5", "c3_2", "c4_2", "c5_2")    # exchange modules, CMPC_model.py:271-283


def rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class Dims:
    """Derived sizes shared by weights and activation buffers."""
    C: int          # v_emb_dim
    R: int          # rnn_size
    Mm: int         # mlp_dim
    T: int          # num_steps
    HID: int        # words_parse hidden (500 in the reference, CMPC_model.py:349)
    h: int
    w: int
    H: int
    W: int
    cin: Dict[str, int]

    @property
    def N(self): return self.h * self.w
    @property
    def LDC(self): return rup(self.C + 8, 64)     # leading dim of C-wide fp16 buffers (room for 8 spatial channels)
    @property
    def LDR(self): return rup(self.R, 64)
    @property
    def GW(self): return rup(self.Mm, 256)        # group width of mlp_dim-wide buffers
    @property
    def HIDP(self): return rup(self.HID, 64)
    @property
    def CH(self): return (self.C + 47) // 48      # MUTAN column chunks


def _t16(x):
    return x.to(torch.float16).contiguous()


def pack_conv1x1(dw: torch.Tensor, kpad: int | None = None, rows_pad: int | None = None) -> torch.Tensor:
    """DW [1,1,Cin,Cout] -> fp16 [rows_pad or Cout, kpad or rup(Cin,64)]"""
    cin, cout = dw.shape[2], dw.shape[3]
    kp = kpad or rup(cin, 64)
    out = torch.zeros(rows_pad or cout, kp, dtype=torch.float32, device=dw.device)
    out[:cout, :cin] = dw[0, 0].t()
    return _t16(out)


def pack_mutan_weights(dws, C: int, kpad: int) -> torch.Tensor:
    """five DW [1,1,C+8,C] -> fp16 [chunks*240, kpad]; row j*240 + k*48 + cc <- head k, channel 48 j + cc."""
    ch = (C + 47) // 48
    out = torch.zeros(ch, 5, 48, kpad, dtype=torch.float32, device=dws[0].device)
    for k, dw in enumerate(dws):
        wt = dw[0, 0].t()                                  # [C, C+8]
        full = torch.zeros(ch * 48, kpad, device=dw.device)
        full[:C, :wt.shape[1]] = wt
        out[:, k] = full.view(ch, 48, kpad)
    return _t16(out.view(ch * 240, kpad))


def pack_head_weights(params: Dict[str, torch.Tensor], d: Dims, device) -> Dict[str, torch.Tensor]:
    """Returns the dict of packed device tensors used by CMPCHeadB200."""
    P = {k: v.to(device=device, dtype=torch.float32) for k, v in params.items()}
    C, R, Mm, GW, LDC, LDR = d.C, d.R, d.Mm, d.GW, d.LDC, d.LDR
    f32 = dict(dtype=torch.float32, device=device)
    W: Dict[str, torch.Tensor] = {}

    def padvec(v, n):
        o = torch.zeros(n, **f32)
        o[:v.numel()] = v.reshape(-1)
        return o

    # laterals (CMPC_model.py:108-112)
    for lvl in LEVELS:
        W[f"lat_w_{lvl}"] = pack_conv1x1(P[f"{lvl}_lateral/DW"])
        W[f"lat_b_{lvl}"] = padvec(P[f"{lvl}_lateral/biases"], rup(C, 256))
    # language parser (:349-351)
    W["parse1_w"] = pack_conv1x1(P["words_parse_1/DW"], kpad=LDR)
    W["parse1_b"] = padvec(P["words_parse_1/biases"], rup(d.HID, 256))
    W["parse2_w"] = P["words_parse_2/DW"][0, 0].contiguous()          # fp32 [HID, 4]
    W["parse2_b"] = P["words_parse_2/biases"].contiguous()
    # words_trans for the three levels, concatenated along N (:378)
    wt = torch.zeros(3 * R, LDR, **f32)
    wtb = torch.zeros(rup(3 * R, 256), **f32)
    for i, lvl in enumerate(LEVELS):
        wt[i * R:(i + 1) * R, :R] = P[f"words_trans_{lvl}/DW"][0, 0].t()
        wtb[i * R:(i + 1) * R] = P[f"words_trans_{lvl}/biases"]
    W["wtrans_w"], W["wtrans_b"] = _t16(wt), wtb
    # lang_trans of the 15 MUTAN heads, concatenated along N (:303-306), and the MUTAN visual weights (:298-299)
    lt = torch.zeros(15 * C, LDR, **f32)
    ltb = torch.zeros(rup(15 * C, 256), **f32)
    for i, lvl in enumerate(LEVELS):
        for k in range(5):
            o = (i * 5 + k) * C
            lt[o:o + C, :R] = P[f"lang_trans_{lvl}_head{k + 1}/DW"][0, 0].t()
            ltb[o:o + C] = P[f"lang_trans_{lvl}_head{k + 1}/biases"]
        W[f"mutan_w_{lvl}"] = pack_mutan_weights([P[f"vis_trans_{lvl}_head{k + 1}/DW"] for k in range(5)], C, LDC)
        mb = torch.zeros(5, LDC, **f32)
        for k in range(5):
            mb[k, :C] = P[f"vis_trans_{lvl}_head{k + 1}/biases"]
        W[f"mutan_b_{lvl}"] = mb
    W["ltrans_w"], W["ltrans_b"] = _t16(lt), ltb
    # relation-aware reasoning
    fsb_w = torch.zeros(3 * GW, LDR, **f32)
    fsb_b = torch.zeros(3 * GW, **f32)
    for i, lvl in enumerate(LEVELS):
        # affinity re-association: Gt[t, cin] = sum_o wt[t, o] * DW2[cin, o]; extra row C carries the bias term b2 . wt
        g = torch.zeros(C + 8, LDR, **f32)
        g[:C, :R] = P[f"spa_graph_trans2_{lvl}/DW"][0, 0]           # TF layout [Cin, Cout] is already [n=cin, k=o]
        g[C, :R] = P[f"spa_graph_trans2_{lvl}/biases"]
        W[f"gt_w_{lvl}"] = _t16(g)
        W[f"gupd_w_{lvl}"] = pack_conv1x1(P[f"gconv_update_spa_graph_{lvl}/DW"], kpad=LDC)
        W[f"gupd_b_{lvl}"] = padvec(P[f"gconv_update_spa_graph_{lvl}/biases"], rup(C, 256))
        for ln in ("feat", "update"):
            W[f"g{ln}_gamma_{lvl}"] = padvec(P[f"gconv_{ln}_ln_spa_graph_{lvl}/gamma"], LDC)
            W[f"g{ln}_beta_{lvl}"] = padvec(P[f"gconv_{ln}_ln_spa_graph_{lvl}/beta"], LDC)
        # fusion conv over concat[vis_la_sp (C), spa_graph (C), lang (R), spatial (8)]  (:338-343)
        dw = P[f"fusion_{lvl}/DW"][0, 0]                             # [2C+R+8, Mm]
        k1p = rup(C, 64)                                             # second K segment starts at the padded first one
        fw = torch.zeros(rup(Mm, 32), k1p + rup(C + 8, 64), **f32)
        fw[:Mm, :C] = dw[:C].t()
        fw[:Mm, k1p:k1p + C] = dw[C:2 * C].t()
        fw[:Mm, k1p + C:k1p + C + 8] = dw[2 * C + R:2 * C + R + 8].t()
        W[f"fusion_w_{lvl}"] = _t16(fw)
        fsb_w[i * GW:i * GW + Mm, :R] = dw[2 * C:2 * C + R].t()      # tiled-language rows become a per-sample bias
        fsb_b[i * GW:i * GW + Mm] = P[f"fusion_{lvl}/biases"]
        W[f"score_w_{lvl}"] = _pack_score(P[f"score_{lvl}/DW"], GW)
        W[f"score_b_{lvl}"] = P[f"score_{lvl}/biases"].detach().cpu()
    W["fsb_w"], W["fsb_b"] = _t16(fsb_w), fsb_b
    W["score_w"] = _pack_score(P["score/DW"], GW)
    W["score_b"] = P["score/biases"].detach().cpu()
    # text-guided exchange (:194-259): 6 modules
    q_w = torch.zeros(6 * GW, LDR, **f32); q_b = torch.zeros(6 * GW, **f32)
    gvl_w = torch.zeros(6 * GW, LDR, **f32); gvl_b = torch.zeros(6 * GW, **f32)
    keyT = torch.zeros(6, Mm, Mm, **f32)
    wg = torch.zeros(6, Mm, Mm, **f32)
    wf = torch.zeros(2, 6, Mm, Mm, **f32); bf = torch.zeros(2, 6, Mm, **f32)
    for i, x in enumerate(EXG):
        q_w[i * GW:i * GW + Mm, :R] = P[f"lang_query_{x}gv_f1/DW"][0, 0].t()
        q_b[i * GW:i * GW + Mm] = P[f"lang_query_{x}gv_f1/biases"]
        gv = P[f"gv_lang_{x}gv_f1/DW"][0, 0]                         # [Mm + R, Mm]: rows 0..Mm-1 pooled, rest language
        wg[i] = gv[:Mm]
        gvl_w[i * GW:i * GW + Mm, :R] = gv[Mm:].t()
        gvl_b[i * GW:i * GW + Mm] = P[f"gv_lang_{x}gv_f1/biases"]
        # key conv folded into the query: u[cin] = sum_o Wk[cin, o] q[o]  (the key bias only shifts the softmax logits)
        keyT[i] = P[f"spa_graph_key_{x}gv_f1/DW"][0, 0].t()
        for j, f in enumerate(("_f1", "_f2")):
            wf[j, i] = P[f"lang_feat_{x}{f}/DW"][0, 0]
            bf[j, i] = P[f"lang_feat_{x}{f}/biases"]
            W[f"se_w_{x}{f}"] = pack_conv1x1(P[f"trans_feat_{x}{f}/DW"], kpad=rup(Mm, 64), rows_pad=rup(Mm, 32))
            W[f"se_b_{x}{f}"] = padvec(P[f"trans_feat_{x}{f}/biases"], GW)
    W.update(q_w=_t16(q_w), q_b=q_b, gvl_w=_t16(gvl_w), gvl_b=gvl_b, keyT=keyT.contiguous(), wg=wg.contiguous(),
             wf1=wf[0].contiguous(), wf2=wf[1].contiguous(), bf1=bf[0].contiguous(), bf2=bf[1].contiguous())
    # ConvLSTM (util/cell.py:42-66): kernel [1,1,2Mm,4Mm] -> rows g*GW + c, K segments [x | h] each padded to 64
    kp = rup(Mm, 64)
    kern = P["rnn/conv_lstm_cell/kernel"][0, 0]                      # [2Mm, 4Mm]
    kw = torch.zeros(4 * GW, 2 * kp, **f32)
    for g in range(4):
        kw[g * GW:g * GW + Mm, :Mm] = kern[:Mm, g * Mm:(g + 1) * Mm].t()
        kw[g * GW:g * GW + Mm, kp:kp + Mm] = kern[Mm:, g * Mm:(g + 1) * Mm].t()
    W["lstm_w"] = _t16(kw)
    for nm in ("W_ci", "W_cf", "W_co"):
        pw = torch.zeros(d.N, GW, **f32)
        pw[:, :Mm] = P[f"rnn/conv_lstm_cell/{nm}"].reshape(d.N, Mm)
        W[f"lstm_{nm}"] = pw
    lg = torch.zeros(5, GW, **f32); lb = torch.zeros(5, GW, **f32)
    for i in range(5):
        nm = "LayerNorm" if i == 0 else f"LayerNorm_{i}"
        lg[i, :Mm] = P[f"rnn/conv_lstm_cell/{nm}/gamma"]
        lb[i, :Mm] = P[f"rnn/conv_lstm_cell/{nm}/beta"]
    W["lstm_ln_gamma"], W["lstm_ln_beta"] = lg, lb
    return W


def _pack_score(dw: torch.Tensor, gw: int) -> torch.Tensor:
    """DW [3,3,Mm,1] -> fp16 [32, rup(Mm,64)] GEMM weight: row k = tap 3*dy + dx (rows 9..31 zero)."""
    mm = dw.shape[2]
    o = torch.zeros(32, rup(mm, 64), dtype=torch.float32, device=dw.device)
    o[:9, :mm] = dw[:, :, :, 0].reshape(9, mm)
    return _t16(o)
