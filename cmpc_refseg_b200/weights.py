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


def _tcopy(dst: torch.Tensor, src: torch.Tensor):
    """dst[c, r] = src[r, c] (dst fp16 view with unit column stride, src fp32 with unit column stride).  On the device this is the
    tiled transpose-cast kernel (a strided torch copy_ of a transposed 1000 x 1000 kernel is ~10x slower); on the CPU (packing
    tests) a plain copy_."""
    if dst.is_cuda and dst.dtype == torch.float16 and src.dtype == torch.float32 and dst.stride(1) == 1 and src.stride(1) == 1:
        from . import _lib as L
        L.check(L.lib().cmpc_transpose_cast_f32_f16(src.data_ptr(), src.stride(0), src.shape[0], src.shape[1], dst.data_ptr(), dst.stride(0),
                                                    0, 0, torch.cuda.current_stream(dst.device).cuda_stream), "transpose_cast")
    else:
        dst.copy_(src.t())


def _t16(x):
    return x.to(torch.float16).contiguous()


# consumers of each source map (index into (f3, f4, f5)) inside an exchange round, as (module index, which lang_se), in the slot order
# that makes the gate layout linear in the module index: gate1 of module m sits at pair position 2 - m, gate2 at 5 - m
# (position = 2 * source + slot; CMPC_model.py:245-259: module c3 reads (f4, f5), c4 reads (f3, f5), c5 reads (f3, f4))
SE_PAIRS = {0: ((2, "_f1"), (1, "_f1")), 1: ((0, "_f1"), (2, "_f2")), 2: ((1, "_f2"), (0, "_f2"))}


def pack_conv1x1(dw: torch.Tensor, kpad: int | None = None, rows_pad: int | None = None) -> torch.Tensor:
    """DW [1,1,Cin,Cout] -> fp16 [rows_pad or Cout, kpad or rup(Cin,64)]"""
    cin, cout = dw.shape[2], dw.shape[3]
    kp = kpad or rup(cin, 64)
    out = torch.zeros(rows_pad or cout, kp, dtype=torch.float16, device=dw.device)
    out[:cout, :cin].copy_(dw[0, 0].t())
    return out


def pack_mutan_weights(dws, C: int, kpad: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """five DW [1,1,C+8,C] -> fp16 [chunks*240, kpad]; row j*240 + k*48 + cc <- head k, channel 48 j + cc."""
    ch = (C + 47) // 48
    if out is None:
        out = torch.zeros(ch * 240, kpad, dtype=torch.float16, device=dws[0].device)
    o4 = out.view(ch, 5, 48, kpad)
    tail = C - (ch - 1) * 48
    for k, dw in enumerate(dws):
        if out.is_cuda:                                    # one tiled transpose-cast per head into its interleaved rows
            from . import _lib as L
            src = dw[0, 0]                                 # [C+8, C]
            L.check(L.lib().cmpc_transpose_cast_f32_f16(src.data_ptr(), src.stride(0), src.shape[0], src.shape[1], o4[0, k].data_ptr(), kpad,
                                                        48, 240 * kpad, torch.cuda.current_stream(out.device).cuda_stream), "transpose_cast")
            continue
        wt = dw[0, 0].t()                                  # [C, C+8]
        kin = wt.shape[1]
        if ch > 1:
            o4[:ch - 1, k, :, :kin].copy_(wt[:(ch - 1) * 48].view(ch - 1, 48, kin))
        o4[ch - 1, k, :tail, :kin].copy_(wt[(ch - 1) * 48:])
    return out


def pack_head_weights(params: Dict[str, torch.Tensor], d: Dims, device, out: Dict[str, torch.Tensor] | None = None) -> Dict[str, torch.Tensor]:
    """Returns the dict of packed device tensors used by CMPCHeadB200.  With `out` (a dict returned by an earlier call) the
    tensors are refreshed IN PLACE (training: after every optimizer step) -- every entry is one strided, dtype-converting copy
    into its persistent zero-padded buffer."""
    P = {k: v.to(device=device, dtype=torch.float32) for k, v in params.items()}
    C, R, Mm, GW, LDC, LDR = d.C, d.R, d.Mm, d.GW, d.LDC, d.LDR
    W: Dict[str, torch.Tensor] = {} if out is None else out

    def buf(name, shape, dtype=torch.float32):
        """persistent zero-initialised destination (pads stay zero across refreshes)"""
        if name not in W:
            W[name] = torch.zeros(shape, dtype=dtype, device=device)
        return W[name]

    f16 = torch.float16
    # laterals (CMPC_model.py:108-112)
    for lvl in LEVELS:
        dw = P[f"{lvl}_lateral/DW"]
        cin = dw.shape[2]
        _tcopy(buf(f"lat_w_{lvl}", (C, rup(cin, 64)), f16)[:, :cin], dw[0, 0])
        buf(f"lat_b_{lvl}", (rup(C, 256),))[:C].copy_(P[f"{lvl}_lateral/biases"])
    # language parser (:349-351)
    _tcopy(buf("parse1_w", (d.HID, LDR), f16)[:, :R], P["words_parse_1/DW"][0, 0])
    buf("parse1_b", (rup(d.HID, 256),))[:d.HID].copy_(P["words_parse_1/biases"])
    buf("parse2_w", (d.HID, 4)).copy_(P["words_parse_2/DW"][0, 0])           # fp32 [HID, 4]
    buf("parse2_b", (4,)).copy_(P["words_parse_2/biases"])
    # words_trans for the three levels, concatenated along N (:378)
    wt, wtb = buf("wtrans_w", (3 * R, LDR), f16), buf("wtrans_b", (rup(3 * R, 256),))
    # lang_trans of the 15 MUTAN heads, concatenated along N (:303-306), and the MUTAN visual weights (:298-299)
    lt, ltb = buf("ltrans_w", (15 * C, LDR), f16), buf("ltrans_b", (rup(15 * C, 256),))
    for i, lvl in enumerate(LEVELS):
        _tcopy(wt[i * R:(i + 1) * R, :R], P[f"words_trans_{lvl}/DW"][0, 0])
        wtb[i * R:(i + 1) * R].copy_(P[f"words_trans_{lvl}/biases"])
        mb = buf(f"mutan_b_{lvl}", (5, LDC))
        for k in range(5):
            o = (i * 5 + k) * C
            _tcopy(lt[o:o + C, :R], P[f"lang_trans_{lvl}_head{k + 1}/DW"][0, 0])
            ltb[o:o + C].copy_(P[f"lang_trans_{lvl}_head{k + 1}/biases"])
            mb[k, :C].copy_(P[f"vis_trans_{lvl}_head{k + 1}/biases"])
        pack_mutan_weights([P[f"vis_trans_{lvl}_head{k + 1}/DW"] for k in range(5)], C, LDC,
                           out=buf(f"mutan_w_{lvl}", (d.CH * 240, LDC), f16))
    # relation-aware reasoning
    fsb_w, fsb_b = buf("fsb_w", (3 * GW, LDR), f16), buf("fsb_b", (3 * GW,))
    k1p = rup(C, 64)                                                 # second K segment starts at the padded first one
    for i, lvl in enumerate(LEVELS):
        # affinity re-association: Gt[t, cin] = sum_o wt[t, o] * DW2[cin, o]; extra row C carries the bias term b2 . wt
        g = buf(f"gt_w_{lvl}", (C + 8, LDR), f16)
        g[:C, :R].copy_(P[f"spa_graph_trans2_{lvl}/DW"][0, 0])       # TF layout [Cin, Cout] is already [n=cin, k=o]
        g[C, :R].copy_(P[f"spa_graph_trans2_{lvl}/biases"])
        _tcopy(buf(f"gupd_w_{lvl}", (C, LDC), f16)[:, :C], P[f"gconv_update_spa_graph_{lvl}/DW"][0, 0])
        buf(f"gupd_b_{lvl}", (rup(C, 256),))[:C].copy_(P[f"gconv_update_spa_graph_{lvl}/biases"])
        for ln in ("feat", "update"):
            buf(f"g{ln}_gamma_{lvl}", (LDC,))[:C].copy_(P[f"gconv_{ln}_ln_spa_graph_{lvl}/gamma"])
            buf(f"g{ln}_beta_{lvl}", (LDC,))[:C].copy_(P[f"gconv_{ln}_ln_spa_graph_{lvl}/beta"])
        # fusion conv over concat[vis_la_sp (C), spa_graph (C), lang (R), spatial (8)]  (:338-343)
        dw = P[f"fusion_{lvl}/DW"][0, 0]                             # [2C+R+8, Mm]
        fw = buf(f"fusion_w_{lvl}", (rup(Mm, 32), k1p + rup(C + 8, 64)), f16)
        _tcopy(fw[:Mm, :C], dw[:C])
        _tcopy(fw[:Mm, k1p:k1p + C], dw[C:2 * C])
        _tcopy(fw[:Mm, k1p + C:k1p + C + 8], dw[2 * C + R:2 * C + R + 8])
        _tcopy(fsb_w[i * GW:i * GW + Mm, :R], dw[2 * C:2 * C + R])      # tiled-language rows become a per-sample bias
        fsb_b[i * GW:i * GW + Mm].copy_(P[f"fusion_{lvl}/biases"])
        _pack_score(P[f"score_{lvl}/DW"], GW, out=buf(f"score_w_{lvl}", (32, rup(Mm, 64)), f16))
        buf(f"score_b_{lvl}", (1,)).copy_(P[f"score_{lvl}/biases"])            # device scalar (read by the kernel)
    _pack_score(P["score/DW"], GW, out=buf("score_w", (32, rup(Mm, 64)), f16))
    buf("score_b", (1,)).copy_(P["score/biases"])
    # text-guided exchange (:194-259): 6 modules
    q_w, q_b = buf("q_w", (6 * GW, LDR), f16), buf("q_b", (6 * GW,))
    gvl_w, gvl_b = buf("gvl_w", (6 * GW, LDR), f16), buf("gvl_b", (6 * GW,))
    keyT, wg = buf("keyT", (6, Mm, Mm)), buf("wg", (6, Mm, Mm))
    wf = [buf("wf1", (6, Mm, Mm)), buf("wf2", (6, Mm, Mm))]
    bf = [buf("bf1", (6, Mm)), buf("bf2", (6, Mm))]
    for i, x in enumerate(EXG):
        _tcopy(q_w[i * GW:i * GW + Mm, :R], P[f"lang_query_{x}gv_f1/DW"][0, 0])
        q_b[i * GW:i * GW + Mm].copy_(P[f"lang_query_{x}gv_f1/biases"])
        gv = P[f"gv_lang_{x}gv_f1/DW"][0, 0]                         # [Mm + R, Mm]: rows 0..Mm-1 pooled, rest language
        wg[i].copy_(gv[:Mm])
        _tcopy(gvl_w[i * GW:i * GW + Mm, :R], gv[Mm:])
        gvl_b[i * GW:i * GW + Mm].copy_(P[f"gv_lang_{x}gv_f1/biases"])
        # key conv folded into the query: u[cin] = sum_o Wk[cin, o] q[o]  (the key bias only shifts the softmax logits)
        _tcopy(keyT[i], P[f"spa_graph_key_{x}gv_f1/DW"][0, 0])
        for j, f in enumerate(("_f1", "_f2")):
            wf[j][i].copy_(P[f"lang_feat_{x}{f}/DW"][0, 0])
            bf[j][i].copy_(P[f"lang_feat_{x}{f}/biases"])
            _tcopy(buf(f"se_w_{x}{f}", (rup(Mm, 32), rup(Mm, 64)), f16)[:Mm, :Mm], P[f"trans_feat_{x}{f}/DW"][0, 0])
            buf(f"se_b_{x}{f}", (GW,))[:Mm].copy_(P[f"trans_feat_{x}{f}/biases"])
    # the two lang_se convs that read the same SOURCE map of an exchange round, concatenated along N (rows slot * GW + c): one GEMM per
    # source map instead of two (head._st_exchange_round, inference path)
    for rnd in range(2):
        for src, cons in SE_PAIRS.items():
            pw = buf(f"sepair_w_{rnd}_{src}", (2 * GW, rup(Mm, 64)), f16)
            pb = buf(f"sepair_b_{rnd}_{src}", (2 * GW,))
            for slot, (mi, f) in enumerate(cons):
                x = EXG[rnd * 3 + mi]
                _tcopy(pw[slot * GW:slot * GW + Mm, :Mm], P[f"trans_feat_{x}{f}/DW"][0, 0])
                pb[slot * GW:slot * GW + Mm].copy_(P[f"trans_feat_{x}{f}/biases"])
    # ConvLSTM (util/cell.py:42-66): kernel [1,1,2Mm,4Mm] -> rows g*GW + c, K segments [x | h] each padded to 64
    kp = rup(Mm, 64)
    kern = P["rnn/conv_lstm_cell/kernel"][0, 0]                      # [2Mm, 4Mm]
    kw = buf("lstm_w", (4 * GW, 2 * kp), f16)
    kw.view(4, GW, 2, kp)[:, :Mm, :, :Mm].copy_(kern.view(2, Mm, 4, Mm).permute(2, 3, 0, 1))
    for nm in ("W_ci", "W_cf", "W_co"):
        buf(f"lstm_{nm}", (d.N, GW))[:, :Mm].copy_(P[f"rnn/conv_lstm_cell/{nm}"].reshape(d.N, Mm))
    for nm in ("W_ci", "W_cf"):          # fp16 copies for the inference GEMM epilogue (half the per-row peephole loads)
        buf(f"lstm_{nm}16", (d.N, GW), f16)[:, :Mm].copy_(P[f"rnn/conv_lstm_cell/{nm}"].reshape(d.N, Mm))
    lg, lb = buf("lstm_ln_gamma", (5, GW)), buf("lstm_ln_beta", (5, GW))
    for i in range(5):
        nm = "LayerNorm" if i == 0 else f"LayerNorm_{i}"
        lg[i, :Mm].copy_(P[f"rnn/conv_lstm_cell/{nm}/gamma"])
        lb[i, :Mm].copy_(P[f"rnn/conv_lstm_cell/{nm}/beta"])
    return W


def _pack_score(dw: torch.Tensor, gw: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """DW [3,3,Mm,1] -> fp16 [32, rup(Mm,64)] GEMM weight: row k = tap 3*dy + dx (rows 9..31 zero)."""
    mm = dw.shape[2]
    if out is None:
        out = torch.zeros(32, rup(mm, 64), dtype=torch.float16, device=dw.device)
    out[:9, :mm].copy_(dw[:, :, :, 0].reshape(9, mm))
    return out
