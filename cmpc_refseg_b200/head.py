"""Host-side orchestration of the CMPC head on one B200: sequences the sm_100a kernels of libcmpc_b200
through its C ABI (ctypes), on the caller's current CUDA stream.  PyTorch is used for device memory and
streams only -- there is no torch arithmetic on the forward path and no fallback of any kind.

Mirrors LSTM_model.build_graph (CMPC_model.py:89-142) after the backbone taps and the word LSTM:
inputs  c3/c4/c5 [B,h,w,512/1024/2048] (fp32 or fp16), lstm_outputs [B,T,R] fp32 (zero past seq_len)
outputs pred [B,h,w,1], up / sigm [B,H,W,1], words_parse [B,1,T,4], gw_w / gw_v [B,N,T] (level c3), up_c3/4/5.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Dict, Optional

import torch

from . import _lib as L
from .weights import Dims, EXG, LEVELS, SE_PAIRS, pack_head_weights, rup


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def on_device(fn):
    """Runs a method with the object's device current: the C ABI identifies the GPU with cudaGetDevice() and launches on the
    stream it is handed, so a head built for cuda:1 must not run while cuda:0 is the current device."""
    @functools.wraps(fn)
    def wrap(self, *a, **k):
        dev = self.device
        if torch.cuda.current_device() == dev.index or dev.index is None:
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrap


# kernels launched per C-ABI call (for the gpu_launches count of bench.py)
_LAUNCHES = {"affinity_softmax": 2, "global_pool": 2, "score": 2, "score_aux": 2}


class CMPCHeadB200:
    """gv_norm='sample' (default): l2_normalize(gv_lang) per sample == the reference at B=1, the way its own inference drivers
    run it; this is what makes a batch shard independent of the other shards.
    gv_norm='batch': the literal graph -- tf.nn.l2_normalize(gv_lang) has no axis (CMPC_model.py:241), so the sum of squares
    runs over the batch too.  This is what the reference trains with (trainval.sh: -bs 8) and what LSTM_model(mode='train')
    selects.  `gv_allreduce`, when set (a callable taking the [3] fp32 tensor of per-module sums), is called between the two
    phases so that ranks holding shards of one batch reproduce the reference at the GLOBAL batch (parallel.gv_sum_allreduce)."""

    def __init__(self, params: Dict[str, torch.Tensor], *, batch_size=1, num_steps=20, vf_h=40, vf_w=40, H=320, W=320,
                 vf_dim=2048, c4_dim=1024, c3_dim=512, v_emb_dim=1000, rnn_size=1000, mlp_dim=500, parse_hidden=500,
                 device="cuda:0", gv_norm="sample"):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._init(params, batch_size, num_steps, vf_h, vf_w, H, W, vf_dim, c4_dim, c3_dim, v_emb_dim, rnn_size, mlp_dim, parse_hidden, gv_norm)

    @on_device
    def _init(self, params, batch_size, num_steps, vf_h, vf_w, H, W, vf_dim, c4_dim, c3_dim, v_emb_dim, rnn_size, mlp_dim, parse_hidden,
              gv_norm):
        if gv_norm not in ("sample", "batch"):
            raise L.CmpcError("gv_norm must be 'sample' or 'batch'")
        self.gv_norm = gv_norm
        self.gv_allreduce = None
        if self.device.type != "cuda":
            raise L.CmpcError("CMPCHeadB200 needs a CUDA device (sm_100a); there is no CPU path")
        if v_emb_dim != rnn_size:
            raise L.CmpcError("the affinity xt . wt^T (CMPC_model.py:384) needs v_emb_dim == rnn_size")
        if v_emb_dim % 8 or mlp_dim % 4 or num_steps > 32 or W % 4:
            raise L.CmpcError("unsupported dims: v_emb_dim % 8 == 0, mlp_dim % 4 == 0, num_steps <= 32, W % 4 == 0 required")
        if mlp_dim > 512 or v_emb_dim > 1016:
            raise L.CmpcError("unsupported dims: mlp_dim <= 512 and v_emb_dim <= 1016")
        self.lib = L.lib()
        self.B = batch_size
        self.d = Dims(C=v_emb_dim, R=rnn_size, Mm=mlp_dim, T=num_steps, HID=parse_hidden, h=vf_h, w=vf_w, H=H, W=W,
                      cin={"c5": vf_dim, "c4": c4_dim, "c3": c3_dim})
        for k, v in self.d.cin.items():
            if v % 8:
                raise L.CmpcError(f"{k} channel count must be a multiple of 8")
        self.params = params
        self.Wt = pack_head_weights(params, self.d, self.device)
        # v_scale: power of two >= N keeps P = W V^T at O(1) in fp16 (divided out in the graph kernel's epilogue)
        self.v_scale = float(1 << (self.d.N - 1).bit_length())
        self._alloc()
        self.t: Dict[str, torch.Tensor] = {}
        self.launches = 0            # kernels launched so far (all of them ours)
        self.prof = None             # optional {name: [(start_event, end_event), ...]} filled by forward()
        self.prof_names = None       # optional set of the event names to record (each record costs host time in the eager pass)
        self.saved = None            # training: a backward.Saved that keeps per-stage activations (see backward.py)
        # run the three independent chains of the language side on parallel streams (forward()); measured (scripts/overlap_ab.py):
        # -0.5..-1.2 % at batch 32, +3 % at batch 1 where the eager pass is bound by host launch time, hence off for small batches
        self.overlap_lang = batch_size >= 8
        self._side = None
        import os
        # inference: the l2_normalize of the MUTAN map (:324) is NOT a pass of its own -- the map stays un-normalised next to its row sums
        # of squares and the four consumers apply 1 / |x| (affinity GEMM + V, the residual of graph_conv, the fusion GEMM through a
        # row scale of its accumulator); set per forward() (never by the per-method API, keep=True or training).  CMPC_LAZY_NORM=0: off.
        self.lazy_norm = os.environ.get("CMPC_LAZY_NORM", "1") != "0"
        self._lazy = False
        self.lstm_state_f16 = os.environ.get("CMPC_LSTM_STATE", "f16") != "f32"   # inference: ConvLSTM cell state in fp16 between passes (A/B knob)
        self.merge_lang_se = os.environ.get("CMPC_MERGE_LANG_SE", "1") != "0"    # inference: one lang_se GEMM per source map of a round (A/B knob)

    # ------------------------------------------------------------------------------------------
    def _alloc(self):
        d, B, dev = self.d, self.B, self.device
        M, BT = B * d.N, B * d.T
        f16 = dict(dtype=torch.float16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        z16 = lambda *s: torch.zeros(*s, **f16)
        z32 = lambda *s: torch.zeros(*s, **f32)
        b = self.buf = {}
        kin = max(d.cin.values())
        b["cin16"] = z16(M, kin)
        b["tmp32"] = z32(M, d.LDC)
        b["rowss"] = z32(6, M)
        b["xlat16"], b["x16"], b["y16"], b["z16"], b["u16"], b["g16"] = (z16(M, d.LDC) for _ in range(6))
        b["affi"] = z32(M, 32)
        b["taps"] = z32(M, 32)
        b["w16"], b["v16"] = z16(M, 32), z16(M, 32)
        b["gw_w"], b["gw_v"] = z32(M, d.T), z32(M, d.T)
        # fp64 statistics arena: graph 3 x [B,2], gupd 3 x [B,2], lstm 3 x ([B,4,2] + [B,2,2])
        b["stats"] = torch.zeros(3 * B * 2 + 3 * B * 2 + 3 * (B * 8 + B * 4), dtype=torch.float64, device=dev)
        b["mr"] = torch.zeros(b["stats"].numel(), dtype=torch.float32, device=dev)      # (mean, rstd) per (sum, sumsq) pair
        for lvl in LEVELS:
            b[f"fus16_{lvl}"] = z16(M, d.GW)
        for nm in ("se1", "se2", "e3", "e4", "e5", "g3", "g4", "g5", "h16"):
            b[nm] = z16(M, d.GW)
        b["y16g"] = z16(M, 4 * d.GW)          # ConvLSTM gate pre-activations j,i,f,o (statistics come from the fp32 accumulators)
        b["cstate"], b["cnew"], b["opre"] = z32(M, d.GW), z32(M, d.GW), z32(M, d.GW)
        b["cstate16"], b["cnew16"] = z16(M, d.GW), z16(M, d.GW)      # inference: the cell state lives in fp16 between the passes
        # language side
        b["words32"] = z32(BT, d.R)
        b["words16"] = z16(BT, d.LDR)
        b["mask"] = z32(BT)
        b["hidden"] = z32(BT, d.HIDP)
        b["parse"] = z32(B, d.T, 4)
        b["rgate"] = z32(B, 32)
        b["valid32"], b["nec32"] = z32(B, d.R), z32(B, d.R)
        b["valid16"], b["nec16"] = z16(B, d.LDR), z16(B, d.LDR)
        b["wt16"] = z16(BT, rup(3 * d.R, 8))
        b["gt16"] = z16(3, BT, d.LDC)
        b["gtb32"] = z32(3, B, 32)            # lazy_norm: column C of Gt (the bias row of the re-associated affinity) as a per-sample bias
        b["lang"] = z32(B, 15 * d.C)
        b["fsb"] = z32(B, 3 * d.GW)
        b["q"], b["u"], b["gvl"] = z32(B, 6 * d.GW), z32(B, 6 * d.GW), z32(B, 6 * d.GW)
        b["pool"] = z32(B, 3, d.GW)
        b["gv"], b["gate1"], b["gate2"] = z32(B, 3, d.GW), z32(B, 3, d.GW), z32(B, 3, d.GW)
        b["gv_ss"] = z32(2, 3)                # gv_norm='batch': per round, per module sum over the batch of |gv_lang|^2
        b["gatepair"] = z32(B, 6, d.GW)       # inference: the six lang_se gates of a round at position 2 * source + slot (weights.SE_PAIRS)
        b["sepair"] = z16(3, M, 2 * d.GW)     # inference: relu(trans_feat) * gate of both consumers of each source map
        ws = max(self.lib.cmpc_affinity_workspace_bytes(B), self.lib.cmpc_global_pool_workspace_bytes(B, 3, d.GW),
                 self.lib.cmpc_score_workspace_bytes(M))
        b["ws"] = torch.zeros(ws, dtype=torch.uint8, device=dev)
        # outputs
        b["pred"] = z32(B, d.h, d.w, 1)
        b["up"], b["sigm"] = z32(B, d.H, d.W, 1), z32(B, d.H, d.W, 1)
        for lvl in LEVELS:
            b[f"pred_{lvl}"] = z32(B, d.h, d.w, 1)
            b[f"up_{lvl}"] = z32(B, d.H, d.W, 1)
        b["iu"] = torch.zeros(B, 2, dtype=torch.int64, device=dev)
        b["ones"] = torch.ones(max(M, 32), **f32)      # unit row scale / unit word weights for the per-method loaders (methods.py)

    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _gemm(self, a1, k1, w, n, out, *, a2=None, k2=0, bias=None, sbias=None, gate=None, act=0, rows_per_sample=None,
              group=None, stats=None, row_sumsq=None, row_scale=None, peep=None, cprev=None, m=None,
              w_batch_stride=0, w_rows=0, a_row_sumsq=None):
        g = L.GemmArgs()
        g.a1, g.lda1, g.k1 = a1.data_ptr(), a1.stride(0), k1
        if a2 is not None:
            g.a2, g.lda2, g.k2 = a2.data_ptr(), a2.stride(0), k2
        g.w, g.ldw = w.data_ptr(), w.stride(-2)
        g.w_batch_stride, g.w_rows = w_batch_stride, w_rows
        g.m = m if m is not None else a1.shape[0]
        g.n = n
        g.rows_per_sample = rows_per_sample or g.m
        g.row_scale, g.bias = _ptr(row_scale), _ptr(bias)
        if sbias is not None:
            g.sbias, g.ld_sbias = sbias.data_ptr(), sbias.stride(0)
        if gate is not None:
            g.gate, g.ld_gate = gate.data_ptr(), gate.stride(0)
        g.act = act
        if group is not None:
            g.group_width, g.group_valid = group
        if peep is not None:
            g.peep_i, g.peep_f, g.ld_peep = peep[0].data_ptr(), peep[1].data_ptr(), peep[0].stride(0)
            g.cprev, g.ld_cprev = cprev.data_ptr(), cprev.stride(0)
            g.peep_f16 = int(cprev.dtype == torch.float16)
            if g.peep_f16 and not (peep[0].dtype == peep[1].dtype == torch.float16):
                raise L.CmpcError("fp16 cell state needs fp16 peephole weights")
        g.out, g.ldo, g.out_fp32 = out.data_ptr(), out.stride(-2), int(out.dtype == torch.float32)
        g.row_sumsq, g.stats = _ptr(row_sumsq), _ptr(stats)
        g.a_row_sumsq = _ptr(a_row_sumsq)
        L.check(self.lib.cmpc_gemm_f16(C.byref(g), self._stream()), "cmpc_gemm_f16")
        self.launches += 1

    def _ck(self, rc, name):
        L.check(rc, name)
        self.launches += _LAUNCHES.get(name, 1)

    def _ev(self, name):
        """Records a CUDA event on the launching stream when profiling is enabled (bench.py roofline leg)."""
        if self.prof is None or (self.prof_names is not None and name not in self.prof_names):
            return None
        # inside a CUDA-graph capture the event must be an external one (an event-record NODE: it is re-recorded by every replay and can
        # be read with elapsed_time after a synchronize); outside it is an ordinary timing event
        e = torch.cuda.Event(enable_timing=True, external=True) if torch.cuda.is_current_stream_capturing() else torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(self.device))
        self.prof.setdefault(name, []).append(e)
        return e

    def _save(self, keep, name, t, cols=None):
        if keep:
            self.t[name] = (t[..., :cols] if cols else t).detach().float().clone()

    # ------------------------------------------------------------------------------------------
    # Stages.  forward() is these, in the reference's build order, over the head's persistent buffers; methods.py exposes
    # the same stages one reference method at a time.
    def _begin(self):
        self.buf["stats"].zero_()
        self.buf["rowss"].zero_()
        self._so = 0

    def _take(self, n):
        so = self._so
        self._so += n
        if self._so > self.buf["stats"].numel():
            raise L.CmpcError("layer-norm statistics arena exhausted")
        return self.buf["stats"][so:so + n], self.buf["mr"][so:so + n]

    def _finalize(self, sm, count):
        """fp64 (sum, sumsq) pairs -> fp32 (mean, rstd) pairs once per sample (not once per consumer thread)"""
        self._ck(self.lib.cmpc_ln_finalize(sm[0].data_ptr(), sm[0].numel() // 2, float(count), sm[1].data_ptr(), self._stream()),
                 "ln_finalize")

    def _st_words(self, lstm_outputs):
        """words_feat = l2_normalize(outputs, -1), seq_mask (CMPC_model.py:159-163)"""
        b, d = self.buf, self.d
        self._ck(self.lib.cmpc_words_prepare(lstm_outputs.data_ptr(), self.B * d.T, d.R, b["words32"].data_ptr(),
                                             b["words16"].data_ptr(), d.LDR, b["mask"].data_ptr(), self._stream()), "words_prepare")

    def _st_parse(self, given_parse=False):
        """build_lang_parser (:347-357) + valid_lang / nec_lang (:166-192); given_parse: b['parse'] is an input"""
        b, d, W = self.buf, self.d, self.Wt
        if not given_parse:
            self._gemm(b["words16"], d.R, W["parse1_w"], d.HID, b["hidden"], bias=W["parse1_b"], act=1)
        self._ck(self.lib.cmpc_lang_parse(None if given_parse else b["hidden"].data_ptr(), d.HIDP, d.HID, W["parse2_w"].data_ptr(),
                                          W["parse2_b"].data_ptr(), b["words32"].data_ptr(), b["mask"].data_ptr(), self.B, d.T,
                                          d.R, d.C, b["parse"].data_ptr(), b["rgate"].data_ptr(), b["valid32"].data_ptr(),
                                          b["nec32"].data_ptr(), b["valid16"].data_ptr(), b["nec16"].data_ptr(), d.LDR,
                                          self._stream()), "lang_parse")

    def _st_words_derived(self):
        """words_trans x3 (:378) and Gt = wt . DW2^T (+ bias row): the affinity re-association"""
        b, d, W = self.buf, self.d, self.Wt
        self._gemm(b["words16"], d.R, W["wtrans_w"], 3 * d.R, b["wt16"], bias=W["wtrans_b"])
        for i, lvl in enumerate(LEVELS):
            self._gemm(b["wt16"][:, i * d.R:], d.R, W[f"gt_w_{lvl}"], d.C + 8, b["gt16"][i])
        if self._lazy:
            b["gtb32"][:, :, :d.T].copy_(b["gt16"][:, :, d.C].view(3, self.B, d.T))

    def _st_valid_derived(self):
        """everything that consumes valid_lang: tanh(lang_trans) of the 15 MUTAN heads (:303-306), language rows of fusion"""
        b, d, W = self.buf, self.d, self.Wt
        self._gemm(b["valid16"], d.R, W["ltrans_w"], 15 * d.C, b["lang"], bias=W["ltrans_b"], act=2)
        self._gemm(b["valid16"], d.R, W["fsb_w"], 3 * d.GW, b["fsb"], bias=W["fsb_b"], group=(d.GW, d.Mm))

    def _st_nec_derived(self):
        """everything that consumes nec_lang: lang_query (:223), language rows of gv_lang (:239), key conv folded into the query"""
        b, d, W = self.buf, self.d, self.Wt
        GW, Mm = d.GW, d.Mm
        self._gemm(b["nec16"], d.R, W["q_w"], 6 * GW, b["q"], bias=W["q_b"], group=(GW, Mm))
        self._gemm(b["nec16"], d.R, W["gvl_w"], 6 * GW, b["gvl"], bias=W["gvl_b"], group=(GW, Mm))
        # accumulate mode (act 4) on a zeroed buffer: the launcher may then split K over blocks (96 blocks of 500 dependent
        # L2 loads each otherwise: 60 us of pure latency at batch 32)
        b["u"].zero_()
        self._ck(self.lib.cmpc_small_linear_f32(b["q"].data_ptr(), 6 * GW, GW, W["keyT"].data_ptr(), Mm, Mm * Mm, None, 0,
                                                b["u"].data_ptr(), 6 * GW, GW, 6, self.B, Mm, Mm, 4, self._stream()), "key_fold")

    def _lb(self, name, i):
        """level buffer: the shared scratch buffer, or -- training (self.saved) -- this level's own copy kept for backward.py"""
        t = self.buf[name]
        if self.saved is None:
            return t
        return self.saved.alloc(f"{name}_{LEVELS[i]}", tuple(t.shape), t.dtype)

    def _st_lateral(self, i, x, keep=False):
        """lateral conv (:108-112): fp16, NOT yet normalised -- the l2_normalize (:109-113) is folded into the MUTAN GEMM's
        epilogue as a per-row scale of the accumulators; the 8 spatial channels (:297) are written pre-divided by that scale"""
        b, d, W, lvl = self.buf, self.d, self.Wt, LEVELS[i]
        M, kin = self.B * d.N, d.cin[lvl]
        x = x.reshape(M, kin)
        if x.dtype == torch.float32:
            cin16 = (b["cin16"].view(-1)[:M * kin].view(M, kin) if self.saved is None
                     else self.saved.alloc(f"cin16_{lvl}", (M, kin), torch.float16))
            self._ck(self.lib.cmpc_cast_f32_f16(x.data_ptr(), kin, cin16.data_ptr(), kin, M, kin, self._stream()), "cast")
        else:
            cin16 = x
        if self.saved is not None:
            self.saved.t[f"cin_{lvl}"] = cin16
        ss_lat = b["rowss"][2 * i]
        self._gemm(cin16, kin, W[f"lat_w_{lvl}"], d.C, self._lb("xlat16", i), bias=W[f"lat_b_{lvl}"], row_sumsq=ss_lat)
        self._ck(self.lib.cmpc_spatial_fixup_f16(self._lb("xlat16", i).data_ptr(), d.LDC, ss_lat.data_ptr(), M, d.C, d.h, d.w,
                                                 self._stream()), "spatial_fixup")
        if keep:
            self.t[f"lateral_{lvl}"] = (self._lb("xlat16", i)[:, :d.C].float() * torch.rsqrt(ss_lat.clamp_min(1e-12)).unsqueeze(1)).clone()

    def _st_mutan(self, i, keep=False):
        """MUTAN fusion, five heads in one GEMM (:295-328): xlat16 (+ its row sum of squares) -> x16 = vis_la_sp | 1"""
        b, d, W, lvl = self.buf, self.d, self.Wt, LEVELS[i]
        M, C_ = self.B * d.N, d.C
        ss_lat, ss_mut = b["rowss"][2 * i], b["rowss"][2 * i + 1]
        ma = L.MutanArgs()
        ma.a, ma.lda, ma.k = self._lb("xlat16", i).data_ptr(), d.LDC, C_ + 8
        ma.a_row_sumsq = ss_lat.data_ptr()
        ma.w, ma.ldw = W[f"mutan_w_{lvl}"].data_ptr(), d.LDC
        ma.m, ma.c, ma.rows_per_sample = M, C_, d.N
        ma.bias, ma.ld_bias = W[f"mutan_b_{lvl}"].data_ptr(), d.LDC
        ma.lang, ma.ld_lang, ma.lang_batch_stride = b["lang"][:, i * 5 * C_:].data_ptr(), C_, 15 * C_
        # the un-normalised tanh map goes straight into x16 as fp16 (its row sums of squares come from the fp32 values) and is
        # normalised in place: 2 + 2 + 2 bytes per element of HBM traffic instead of 4 + 4 + 2 through an fp32 scratch map
        # TRAINING keeps the fp32 scratch map (one rounding, after the normalisation): rounding the tanh map to fp16 before AND after
        # the l2_normalize costs nothing visible in the logits but raised the gradient error against the reference's train_op by 1.5x
        # (median 0.41 % -> 0.62 %, worst 2.0 % -> 3.0 %, tests/test_reference_parity_gpu.py)
        x16 = self._lb("x16", i)
        via32 = self.saved is not None
        ma.out, ma.ldo, ma.row_sumsq, ma.out_f16 = (b["tmp32"] if via32 else x16).data_ptr(), d.LDC, ss_mut.data_ptr(), 0 if via32 else 1
        self._ev("mutan")
        self._ck(self.lib.cmpc_mutan_f16(C.byref(ma), self._stream()), "mutan")
        self._ev("mutan")
        if via32:
            self._ck(self.lib.cmpc_rownorm_f16(b["tmp32"].data_ptr(), d.LDC, ss_mut.data_ptr(), x16.data_ptr(), d.LDC, M, C_,
                                               -1, 0, d.N, self._stream()), "rownorm_mutan")
        elif self._lazy:
            pass                                 # x16 stays un-normalised; ss_mut travels to the consumers (see lazy_norm)
        else:
            self._ck(self.lib.cmpc_rownorm_h16(x16.data_ptr(), d.LDC, ss_mut.data_ptr(), x16.data_ptr(), d.LDC, M, C_,
                                               -1, 0, d.N, self._stream()), "rownorm_mutan")     # column C := 1 (bias row of Gt)
        self._save(keep, f"vis_la_sp_{lvl}", self._lb("x16", i), C_)

    def _st_affinity(self, i, want_gw, keep=False):
        """affinity (:378-388): affi = X . Gt_b^T * R_t / sqrt(C), per-sample B operand; the two softmaxes (:389-391)"""
        b, d, lvl = self.buf, self.d, LEVELS[i]
        ss_mut = b["rowss"][2 * i + 1] if self._lazy else None
        if self._lazy:      # un-normalised X: accumulators scaled by 1 / |x|, the bias row of Gt as a per-sample bias, V carries 1 / |x_j|
            self._gemm(self._lb("x16", i), d.C, b["gt16"][i], 32, self._lb("affi", i), gate=b["rgate"], sbias=b["gtb32"][i],
                       rows_per_sample=d.N, w_batch_stride=d.T * d.LDC, w_rows=d.T, a_row_sumsq=ss_mut)
        else:
            self._gemm(self._lb("x16", i), d.C + 8, b["gt16"][i], 32, self._lb("affi", i), gate=b["rgate"], rows_per_sample=d.N,
                       w_batch_stride=d.T * d.LDC, w_rows=d.T)
        self._ck(self.lib.cmpc_affinity_softmax_scaled(self._lb("affi", i).data_ptr(), b["mask"].data_ptr(), self.B, d.N, d.T, self.v_scale,
                                                       self._lb("w16", i).data_ptr(), self._lb("v16", i).data_ptr(),
                                                       b["gw_w"].data_ptr() if want_gw or keep else None,
                                                       b["gw_v"].data_ptr() if want_gw or keep else None, _ptr(ss_mut),
                                                       b["ws"].data_ptr(), b["ws"].numel(), self._stream()), "affinity_softmax")
        self._save(keep, f"affi_{lvl}", self._lb("affi", i), d.T)
        self._save(keep, f"gw_w_{lvl}", b["gw_w"]); self._save(keep, f"gw_v_{lvl}", b["gw_v"])

    def _st_graph_conv(self, i, keep=False, normalize=True):
        """graph_conv (:359-374) with the dense aggregation adj @ X, adjacency never in HBM (:400, :362), then the
        l2_normalize of build_spa_graph (:408) unless normalize=False: x16, w16, v16 -> g16 (+ spatial channels)"""
        b, d, W, lvl, lib, st = self.buf, self.d, self.Wt, LEVELS[i], self.lib, self._stream()
        B, N, M, C_ = self.B, d.N, self.B * d.N, d.C
        st_y, st_u = self._take(2 * B), self._take(2 * B)
        self._ev("graph")
        self._ck(lib.cmpc_graph_reason_f16(self._lb("w16", i).data_ptr(), self._lb("v16", i).data_ptr(), self._lb("x16", i).data_ptr(), d.LDC, B, N, C_,
                                           self.v_scale, self._lb("y16", i).data_ptr(), d.LDC, st_y[0].data_ptr(), None, st), "graph_reason")
        self._ev("graph")
        self._save(keep, f"gconv_y_{lvl}", self._lb("y16", i), C_)
        self._finalize(st_y, N * C_)
        ss_mut = b["rowss"][2 * i + 1] if self._lazy else None
        self._ck(lib.cmpc_ln_residual_relu_scaled_f16(self._lb("y16", i).data_ptr(), d.LDC, self._lb("x16", i).data_ptr(), d.LDC, _ptr(ss_mut),
                                                      st_y[1].data_ptr(), W[f"gfeat_gamma_{lvl}"].data_ptr(), W[f"gfeat_beta_{lvl}"].data_ptr(),
                                                      self._lb("z16", i).data_ptr(), d.LDC, M, C_, N, st), "ln_residual_relu")
        self._gemm(self._lb("z16", i), C_, W[f"gupd_w_{lvl}"], C_, self._lb("u16", i), bias=W[f"gupd_b_{lvl}"], rows_per_sample=N, stats=st_u[0])
        self._finalize(st_u, N * C_)
        # lazy_norm: g16 (and its spatial channels) leave multiplied by |x|, the fusion GEMM scales its whole accumulator by 1 / |x|
        self._ck(lib.cmpc_ln_relu_l2norm_scaled_f16(self._lb("u16", i).data_ptr(), d.LDC, st_u[1].data_ptr(), W[f"gupdate_gamma_{lvl}"].data_ptr(),
                                                    W[f"gupdate_beta_{lvl}"].data_ptr(), self._lb("g16", i).data_ptr(), d.LDC, M, C_, d.h, d.w,
                                                    N, int(normalize),
                                                    None if self.saved is None else self.saved.alloc(f"rss_g_{lvl}", (M,), torch.float32).data_ptr(),
                                                    _ptr(ss_mut), st), "ln_relu_l2norm")
        if self.saved is not None:
            self.saved.t[f"mr_y_{lvl}"], self.saved.t[f"mr_u_{lvl}"] = st_y[1], st_u[1]
        self._save(keep, f"spa_graph_{lvl}", self._lb("g16", i), C_)

    def _st_fusion(self, i, keep=False):
        """fusion conv over [vis_la_sp | spa_graph | tile(valid_lang) | spatial] (:338-344)"""
        b, d, W, lvl = self.buf, self.d, self.Wt, LEVELS[i]
        self._gemm(self._lb("x16", i), d.C, W[f"fusion_w_{lvl}"], d.Mm, b[f"fus16_{lvl}"], a2=self._lb("g16", i), k2=d.C + 8,
                   sbias=b["fsb"][:, i * d.GW:], act=1, rows_per_sample=d.N,
                   a_row_sumsq=b["rowss"][2 * i + 1] if self._lazy else None)
        self._save(keep, f"fusion_{lvl}", b[f"fus16_{lvl}"], d.Mm)

    def _st_global_vec(self, feats, slot0, nmod, rnd=None, pair_layout=False):
        """global_vec (:212-243) of nmod exchange modules starting at EXG slot slot0 -> gv, gate1, gate2 [B, nmod(3), GW].
        Training (self.saved and rnd given): the round keeps its own pool / gv / gates and the softmax statistics."""
        b, d, W, lib, st = self.buf, self.d, self.Wt, self.lib, self._stream()
        GW, Mm = d.GW, d.Mm
        sv = self.saved if rnd is not None else None
        if sv is not None:
            pool, gv, g1, g2 = (sv.alloc(f"exg{rnd}_{nm}", (self.B, 3, GW), torch.float32) for nm in ("pool", "gv", "gate1", "gate2"))
            pstats = sv.alloc(f"exg{rnd}_pstats", (self.B, 3, 2), torch.float32)
            gv_ss = sv.alloc(f"exg{rnd}_gv_ss", (3,), torch.float32)
        else:
            pool, gv, g1, g2, pstats, gv_ss = b["pool"], b["gv"], b["gate1"], b["gate2"], None, b["gv_ss"][slot0 // 3]
        fp = [f.data_ptr() for f in feats] + [None] * (3 - len(feats))
        self._ck(lib.cmpc_global_pool_f16(fp[0], fp[1], fp[2], GW, b["u"][:, slot0 * GW:].data_ptr(), GW, 6 * GW, nmod, self.B,
                                          d.N, GW, 1.0 / (Mm ** 0.5), pool.data_ptr(), GW, _ptr(pstats), b["ws"].data_ptr(),
                                          b["ws"].numel(), st), "global_pool")
        head = (pool.data_ptr(), GW, b["gvl"][:, slot0 * GW:].data_ptr(), GW, 6 * GW,
                W["wg"][slot0:].data_ptr(), W["wf1"][slot0:].data_ptr(), W["bf1"][slot0:].data_ptr(),
                W["wf2"][slot0:].data_ptr(), W["bf2"][slot0:].data_ptr(), Mm * Mm, Mm, self.B, nmod, Mm, gv.data_ptr())
        if pair_layout:
            # gate1 of module m at pair position 2 - m, gate2 at 5 - m of b["gatepair"] [B, 6, GW] (weights.SE_PAIRS)
            gp = b["gatepair"]
            gates = (gp[0, 2].data_ptr(), 6 * GW, -GW, gp[0, 5].data_ptr(), 6 * GW, -GW, GW)
        else:
            gates = (g1.data_ptr(), nmod * GW, GW, g2.data_ptr(), nmod * GW, GW, GW)      # [B, nmod, GW] (nmod = 1 from methods.py)
        if self.gv_norm == "sample":
            self._ck(lib.cmpc_gv_gates_ex(*head, *gates, 0, None, st), "gv_gates")
        else:
            # the literal axis-less l2_normalize (:241): sum of squares over the batch between two launches
            gv_ss.zero_()
            self._ck(lib.cmpc_gv_gates_ex(*head, *gates, 1, gv_ss.data_ptr(), st), "gv_gates")
            if self.gv_allreduce is not None:
                self.gv_allreduce(gv_ss)
            self._ck(lib.cmpc_gv_gates_ex(*head, *gates, 2, gv_ss.data_ptr(), st), "gv_gates")
        return g1, g2

    def _st_lang_se(self, feat, name, gate, out):
        """lang_se (:194-210): relu(trans_feat conv) * sigmoid(lang_feat conv) with the gate [B, GW] precomputed"""
        d, W = self.d, self.Wt
        self._gemm(feat, d.Mm, W[f"se_w_{name}"], d.Mm, out, bias=W[f"se_b_{name}"], act=1, gate=gate, rows_per_sample=d.N)

    def _st_exchange_round(self, rnd, f3, f4, f5, outs, keep=False):
        """one round of gated_exchange_module x3 + l2_normalize (:245-259, :271-284).
        Training (self.saved): the two lang_se maps of every module and the row norms of the sums are kept for backward.py."""
        b, d = self.buf, self.d
        M = self.B * d.N
        sv = self.saved
        mods = EXG[rnd * 3:rnd * 3 + 3]
        if sv is None and self.merge_lang_se:
            # inference: ONE lang_se GEMM per source map (both of its consumers side by side, N = 2 GW) instead of two -- six launches
            # per round become three, each source map is read once -- then add + l2-normalise per module from the column halves
            self._st_global_vec((f3, f4, f5), rnd * 3, 3, pair_layout=True)
            feats, W, GW = (f3, f4, f5), self.Wt, d.GW
            for src in range(3):
                self._gemm(feats[src], d.Mm, W[f"sepair_w_{rnd}_{src}"], 2 * GW, b["sepair"][src], bias=W[f"sepair_b_{rnd}_{src}"], act=1,
                           gate=b["gatepair"][:, 2 * src], rows_per_sample=d.N, group=(GW, d.Mm))
            where = {(mi, f): (src, slot) for src, cons in SE_PAIRS.items() for slot, (mi, f) in enumerate(cons)}
            for mi, on in enumerate(outs):
                (sa, la), (sb, lb) = where[(mi, "_f1")], where[(mi, "_f2")]
                se_a, se_b = b["sepair"][sa][:, la * GW:], b["sepair"][sb][:, lb * GW:]
                self._ev("exchange")
                self._ck(self.lib.cmpc_add3_l2norm_ld_f16(feats[mi].data_ptr(), GW, se_a.data_ptr(), 2 * GW, se_b.data_ptr(), 2 * GW,
                                                          b[on].data_ptr(), GW, M, GW, 1, None, self._stream()), "add3_l2norm")
                self._ev("exchange")
            f3, f4, f5 = (b[o] for o in outs)
            self._save(keep, f"exg{rnd + 1}_c3", f3, d.Mm); self._save(keep, f"exg{rnd + 1}_c4", f4, d.Mm)
            self._save(keep, f"exg{rnd + 1}_c5", f5, d.Mm)
            return f3, f4, f5
        g1, g2 = self._st_global_vec((f3, f4, f5), rnd * 3, 3, rnd=rnd if sv is not None else None)
        triples = ((f3, f4, f5), (f4, f3, f5), (f5, f3, f4))
        if sv is not None:
            sv.t[f"exg{rnd}_in"] = (f3, f4, f5)
        for mi, (x, (feat, fa, fb), on) in enumerate(zip(mods, triples, outs)):
            if sv is not None:
                se1, se2 = sv.alloc(f"exg_se1_{x}", (M, d.GW), torch.float16), sv.alloc(f"exg_se2_{x}", (M, d.GW), torch.float16)
                rss = sv.alloc(f"exg_rss_{x}", (M,), torch.float32)
            else:
                se1, se2, rss = b["se1"], b["se2"], None
            self._st_lang_se(fa, f"{x}_f1", g1[:, mi], se1)
            self._st_lang_se(fb, f"{x}_f2", g2[:, mi], se2)
            self._ev("exchange")
            self._ck(self.lib.cmpc_add3_l2norm_f16(feat.data_ptr(), se1.data_ptr(), se2.data_ptr(), d.GW,
                                                   b[on].data_ptr(), d.GW, M, d.GW, 1, _ptr(rss), self._stream()), "add3_l2norm")
            self._ev("exchange")
        f3, f4, f5 = (b[o] for o in outs)
        self._save(keep, f"exg{rnd + 1}_c3", f3, d.Mm); self._save(keep, f"exg{rnd + 1}_c4", f4, d.Mm)
        self._save(keep, f"exg{rnd + 1}_c5", f5, d.Mm)
        return f3, f4, f5

    def _st_convlstm(self, seq, keep=False):
        """ConvLSTM fusion over (c3, c4, c5) (:287-290, util/cell.py:36-79) -> h16.
        With self.saved set (training, backward.py) every step keeps its own y / c' / o' / state / h and statistics."""
        b, d, W, lib, st = self.buf, self.d, self.Wt, self.lib, self._stream()
        B, N, M, Mm, GW = self.B, d.N, self.B * d.N, d.Mm, d.GW
        sv = self.saved
        cprev = hprev = None
        for step, xin in enumerate(seq):
            st_g, st_o = self._take(8 * B), self._take(4 * B)
            first = step == 0
            if sv is not None:
                y16g, cnew, opre, cstate, h16 = (sv.alloc(f"lstm_{nm}{step}", shape, dt) for nm, shape, dt in (
                    ("y", (M, 4 * GW), torch.float16), ("cnew", (M, GW), torch.float32), ("opre", (M, GW), torch.float32),
                    ("cn", (M, GW), torch.float32), ("h", (M, GW), torch.float16)))
                sv.t[f"lstm_x{step}"], sv.t[f"lstm_mr_g{step}"], sv.t[f"lstm_mr_o{step}"] = xin, st_g[1], st_o[1]
            else:
                # inference ("lean"): the cell state (c' before and c after its layer norm) is kept in fp16 between the passes, like
                # every other activation that feeds a GEMM here (its statistics are still taken from the fp32 values); W_ci / W_cf are
                # read as fp16 by the GEMM epilogue: 4 instead of 8 per-row peephole loads per chunk.  CMPC_LSTM_STATE=f32 keeps fp32.
                st16 = self.lstm_state_f16
                y16g, opre, h16 = b["y16g"], b["opre"], b["h16"]
                cnew, cstate = (b["cnew16"], b["cstate16"]) if st16 else (b["cnew"], b["cstate"])
                cprev, hprev = (None, None) if first else (cstate, h16)
            lean = sv is None
            wci, wcf = (W["lstm_W_ci16"], W["lstm_W_cf16"]) if (lean and self.lstm_state_f16) else (W["lstm_W_ci"], W["lstm_W_cf"])
            self._gemm(xin, Mm, W["lstm_w"], 4 * GW, y16g, a2=hprev, k2=0 if first else Mm,
                       group=(GW, Mm), rows_per_sample=N, stats=st_g[0],
                       peep=None if first else (wci, wcf), cprev=cprev)
            self._finalize(st_g, N * Mm)
            # inference: o' = o + W_co c' is not materialised (gates2 recomputes it from the fp16 gate map) and the last step keeps
            # no cell state -- 3 KB less HBM traffic per row and step; training keeps every tensor for backward.py
            if lean and self.lstm_state_f16:
                self._ck(lib.cmpc_convlstm_gates1_h16(y16g.data_ptr(), 4 * GW, GW, Mm, st_g[1].data_ptr(), W["lstm_ln_gamma"].data_ptr(),
                                                      W["lstm_ln_beta"].data_ptr(), _ptr(cprev), W["lstm_W_co"].data_ptr(), cnew.data_ptr(),
                                                      st_o[0].data_ptr(), M, N, st), "convlstm_gates1")
            else:
                self._ck(lib.cmpc_convlstm_gates1(y16g.data_ptr(), 1, 4 * GW, GW, Mm, st_g[1].data_ptr(), W["lstm_ln_gamma"].data_ptr(),
                                                  W["lstm_ln_beta"].data_ptr(), _ptr(cprev),
                                                  W["lstm_W_co"].data_ptr(), cnew.data_ptr(), None if lean else opre.data_ptr(),
                                                  st_o[0].data_ptr(), M, N, st), "convlstm_gates1")
            self._finalize(st_o, N * Mm)
            if lean:
                last = step == len(seq) - 1
                self._ck(lib.cmpc_convlstm_gates2_y16(y16g[:, 3 * GW:].data_ptr(), 4 * GW, W["lstm_W_co"].data_ptr(), cnew.data_ptr(), GW, Mm,
                                                      st_o[1].data_ptr(), W["lstm_ln_gamma"].data_ptr(), W["lstm_ln_beta"].data_ptr(),
                                                      None if last else cstate.data_ptr(), h16.data_ptr(), int(self.lstm_state_f16), M, N, st),
                         "convlstm_gates2")
            else:
                self._ck(lib.cmpc_convlstm_gates2(opre.data_ptr(), cnew.data_ptr(), GW, Mm, st_o[1].data_ptr(),
                                                  W["lstm_ln_gamma"].data_ptr(), W["lstm_ln_beta"].data_ptr(), cstate.data_ptr(),
                                                  h16.data_ptr(), None, M, N, st), "convlstm_gates2")
            self._save(keep, f"convlstm_h{step}", h16, Mm)
            if sv is not None:
                cprev, hprev = cstate, h16
        self._save(keep, "fused", h16, Mm)
        return h16

    def _st_score(self, src, wname, pred, up, sigm, tag="score"):
        """3x3 score conv as a skinny GEMM over the 9 taps + legacy bilinear upsample (+ sigmoid) (:128-142)"""
        b, d, W = self.buf, self.d, self.Wt
        self._gemm(src, d.Mm, W[wname + "_w" if wname == "score" else wname.replace("score_", "score_w_")], 32, b["taps"], w_rows=9)
        bias = W["score_b"] if wname == "score" else W[wname.replace("score_", "score_b_")]
        self._ck(self.lib.cmpc_score_from_taps(b["taps"].data_ptr(), 32, 0.0, bias.data_ptr(), self.B, d.h, d.w, d.H, d.W, pred.data_ptr(),
                                               _ptr(up), _ptr(sigm), self._stream()), tag)

    # ------------------------------------------------------------------------------------------
    def _check_inputs(self, feats, lstm_outputs):
        d, B = self.d, self.B
        for k, v in feats.items():
            if tuple(v.shape) != (B, d.h, d.w, d.cin[k]) or v.device != self.device or not v.is_contiguous():
                raise L.CmpcError(f"{k}: expected contiguous {(B, d.h, d.w, d.cin[k])} on {self.device}, got {tuple(v.shape)} on {v.device}")
            if v.dtype not in (torch.float32, torch.float16):
                raise L.CmpcError(f"{k}: dtype must be float32 or float16")
        if tuple(lstm_outputs.shape) != (B, d.T, d.R) or lstm_outputs.dtype != torch.float32 or lstm_outputs.device != self.device:
            raise L.CmpcError(f"lstm_outputs: expected float32 {(B, d.T, d.R)} on {self.device}")

    @on_device
    def forward(self, c3, c4, c5, lstm_outputs, seq_len=None, *, aux=False, keep=False) -> Dict[str, torch.Tensor]:
        """seq_len is accepted for signature parity with the reference's feed_dict; as in the reference the word
        mask is derived from the (already zeroed) LSTM outputs (CMPC_model.py:163)."""
        d, B, b = self.d, self.B, self.buf
        feats = {"c3": c3, "c4": c4, "c5": c5}
        self._check_inputs(feats, lstm_outputs)
        self._begin()
        self._lazy = self.lazy_norm and self.saved is None and not keep
        try:
            return self._forward(feats, lstm_outputs, aux, keep)
        finally:
            self._lazy = False

    def _forward(self, feats, lstm_outputs, aux, keep):
        d, B, b = self.d, self.B, self.buf
        # ---------------- language side (CMPC_model.py:159-192, 347-357) ----------------
        lstm_outputs = lstm_outputs.contiguous()
        if self.saved is not None:
            self.saved.t["lstm_outputs"] = lstm_outputs
        first_lateral_done = False
        if self.overlap_lang:
            # The language side is twelve launch-bound launches (B*T or B rows each, ~15 us apiece, a handful of CTAs) in three chains:
            # words -> parse -> valid-derived, words_trans -> Gt, and nec-derived.  They run on three side streams (fork / join by
            # events, which a CUDA-graph capture of this pass records as parallel branches) UNDERNEATH the first level's fp32 -> fp16
            # cast and lateral conv, which need nothing from the language side: the ~90 us the chains take end to end would otherwise
            # be 90 us of an almost idle GPU in front of every pass.  Only persistent buffers are touched there.
            main = torch.cuda.current_stream(self.device)
            if self._side is None:
                self._side = tuple(torch.cuda.Stream(self.device) for _ in range(3))
            sa, sb, sc = self._side
            ev0, evw, evp = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            ev0.record(main)                                     # after _begin()'s memsets and whatever produced the inputs
            with torch.cuda.stream(sa):
                sa.wait_event(ev0)
                self._st_words(lstm_outputs)
                evw.record(sa)
                self._st_parse()
                evp.record(sa)
                self._st_valid_derived()
            with torch.cuda.stream(sb):
                sb.wait_event(evw)
                self._st_words_derived()
            with torch.cuda.stream(sc):
                sc.wait_event(evp)
                self._st_nec_derived()
            self._st_lateral(0, feats[LEVELS[0]], keep)
            first_lateral_done = True
            main.wait_stream(sa)
            main.wait_stream(sb)
            main.wait_stream(sc)
        else:
            self._st_words(lstm_outputs)
            self._st_parse()
            self._st_words_derived()
            self._st_valid_derived()
            self._st_nec_derived()
        self._save(keep, "valid_lang", b["valid32"]); self._save(keep, "nec_lang", b["nec32"])
        # ---------------- per level: entity perception + relation-aware reasoning (:120-125) ----------------
        for i, lvl in enumerate(LEVELS):
            if i > 0 or not first_lateral_done:
                self._st_lateral(i, feats[lvl], keep)
            self._st_mutan(i, keep)
            self._st_affinity(i, lvl == "c3", keep)   # the reference's gw_w / gw_v attributes end up pointing at level c3 (App. D-4)
            self._st_graph_conv(i, keep)
            self._st_fusion(i, keep)
        out: Dict[str, torch.Tensor] = {}
        if aux:                                                                           # :128-133
            for lvl in LEVELS:
                self._st_score(b[f"fus16_{lvl}"], f"score_{lvl}", b[f"pred_{lvl}"], b[f"up_{lvl}"], None, tag="score_aux")
                out[f"up_{lvl}"] = b[f"up_{lvl}"]
        # ---------------- text-guided exchange, two rounds (:261-284), ConvLSTM (:287-290), score (:138-142) ----------------
        f3, f4, f5 = b["fus16_c3"], b["fus16_c4"], b["fus16_c5"]
        f3, f4, f5 = self._st_exchange_round(0, f3, f4, f5, ("e3", "e4", "e5"), keep)
        f3, f4, f5 = self._st_exchange_round(1, f3, f4, f5, ("g3", "g4", "g5"), keep)
        h = self._st_convlstm((f3, f4, f5), keep)
        self._st_score(h, "score", b["pred"], b["up"], b["sigm"])
        T, N, R = d.T, d.N, d.R
        out.update(pred=b["pred"], up=b["up"], sigm=b["sigm"], words_parse=b["parse"].view(B, 1, T, 4),
                   seq_mask=b["mask"].view(B, 1, T, 1), gw_w=b["gw_w"].view(B, N, T), gw_v=b["gw_v"].view(B, N, T),
                   valid_lang=b["valid32"].view(B, 1, 1, R), nec_lang=b["nec32"].view(B, 1, 1, R))
        return out

    __call__ = forward

    @on_device
    def forward_graphed(self, c3, c4, c5, lstm_outputs, seq_len=None, *, aux=False) -> Dict[str, torch.Tensor]:
        """Same as forward(), replayed from a CUDA graph: the ~100 launches of one pass are captured once per set of input
        buffers (keyed by their addresses) and then cost a single graph launch -- what matters at batch 1, where the pass
        is launch-bound (1.2 ms eager vs the sum of its kernels).  Outputs are the head's persistent buffers, as in forward()."""
        key = (c3.data_ptr(), c4.data_ptr(), c5.data_ptr(), lstm_outputs.data_ptr(), c3.dtype, bool(aux))
        graphs = self.__dict__.setdefault("_graphs", {})          # one captured graph per set of input buffers (double-buffered callers)
        if key not in graphs:
            if len(graphs) >= 4:
                graphs.pop(next(iter(graphs)))
            prof, self.prof = self.prof, None                    # profiling events: only the captured ones are kept (see _ev)
            # in a graph the language-side chains are parallel branches at no host cost: fork them at every batch size (batch 1:
            # 0.86 -> 0.77 ms per replay; eager, the event traffic of the fork costs more than it hides below batch 8)
            overlap, self.overlap_lang = self.overlap_lang, True
            try:
                for _ in range(2):                               # warm-up outside capture (lazy cudaFuncSetAttribute etc.)
                    self.forward(c3, c4, c5, lstm_outputs, seq_len, aux=aux)
                torch.cuda.synchronize(self.device)
                self.prof = prof
                if self.prof is not None:
                    self.prof.clear()
                l0 = self.launches
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = self.forward(c3, c4, c5, lstm_outputs, seq_len, aux=aux)
            finally:
                self.overlap_lang = overlap
                self.prof = prof
            graphs[key] = (g, out, self.launches - l0)
            self.launches = l0
        g, out, n_launches = graphs[key]
        g.replay()
        self.launches += n_launches                               # kernels of this library inside one replay
        return out

    @on_device
    def ce_sums(self, logits: torch.Tensor, target_fine: torch.Tensor) -> torch.Tensor:
        """Per-sample sum over pixels of sigmoid cross-entropy (util/loss.py:12-14), fp64 [B]."""
        B = logits.shape[0]
        sums = torch.zeros(B, dtype=torch.float64, device=self.device)
        L.check(self.lib.cmpc_sigmoid_ce_sums(logits.contiguous().data_ptr(), target_fine.contiguous().data_ptr(), B,
                                              logits.numel() // B, sums.data_ptr(), self._stream()), "sigmoid_ce_sums")
        self.launches += 1
        return sums

    # ------------------------------------------------------------------------------------------
    @on_device
    def mask_iu(self, up: torch.Tensor, target_fine: torch.Tensor, thresh: float = 0.0, inclusive: bool = False):
        """Integer I/U per sample of (up > thresh) vs target (CMPC_model.py:486-489; util/eval_tools.py:31-35)."""
        B = up.shape[0]
        iu = torch.zeros(B, 2, dtype=torch.int64, device=self.device)
        per = up.numel() // B
        L.check(self.lib.cmpc_iou_counts(up.contiguous().data_ptr(), target_fine.contiguous().data_ptr(), B, per, thresh,
                                         int(inclusive), iu.data_ptr(), self._stream()), "iou_counts")
        return iu[:, 0], iu[:, 1]
