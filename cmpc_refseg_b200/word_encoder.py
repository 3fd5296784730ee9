"""Word encoder in front of the head (CMPC_model.py:144-157): GloVe lookup + LSTMCell(rnn_size) over the
expression, on the device.  Its output is the head's `lstm_outputs` input, so with it `words` / `seq_len` become the entry
signature of the drop-in, as in the reference.  Variables (below text_objseg/): `Variable` [vocab, glove_dim] (the embedding
matrix, :145), `rnn/lstm_cell/kernel` [glove_dim + rnn_size, 4 rnn_size], `rnn/lstm_cell/bias` [4 rnn_size].

Training (the reference's train_op updates all three, :426-431): `forward(..., train=True)` keeps the gate activations and the
state of every step; `backward(d_outputs, grads)` runs back-propagation through time -- per step one pointwise kernel and one
skinny GEMM (dz K_h^T), then three GEMMs over all steps at once (d K_x = X^T dz, d K_h = H_prev^T dz, d x = dz K_x^T) and a
scatter-add into the embedding rows.  fp16 gradient operands carry a power-of-two scale (GRAD_SCALE) against underflow."""
from __future__ import annotations

from typing import Dict

import torch

from .head import on_device

from . import _lib as L
from .weights import rup

EMB, KERNEL, BIAS = "Variable", "rnn/lstm_cell/kernel", "rnn/lstm_cell/bias"
GRAD_SCALE = 1024.0


class WordEncoderB200:
    def __init__(self, head, params: Dict[str, torch.Tensor]):
        self.h = head
        self.device = head.device
        d, dev = head.d, head.device
        self.params = {k: params[k].to(dev, torch.float32) for k in (EMB, KERNEL, BIAS)}
        self.V, self.E = self.params[EMB].shape
        R = d.R
        if tuple(self.params[KERNEL].shape) != (self.E + R, 4 * R) or tuple(self.params[BIAS].shape) != (4 * R,):
            raise L.CmpcError(f"word LSTM variables: kernel {tuple(self.params[KERNEL].shape)}, bias {tuple(self.params[BIAS].shape)} "
                              f"for glove_dim {self.E}, rnn_size {R}")
        if R % 2:
            raise L.CmpcError("rnn_size must be even")
        self.ldx, self.E8, self.ld4 = rup(self.E, 64), rup(self.E, 8), rup(4 * R, 8)
        f16 = dict(dtype=torch.float16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.wx = torch.zeros(4 * R, self.ldx, **f16)            # [n = 4R, k = E]: input half, forward
        self.wh = torch.zeros(4 * R, d.LDR, **f16)               # [n = 4R, k = R]: recurrent half, forward
        self.bias = torch.zeros(rup(4 * R, 256), **f32)
        self.wxb = self.whb = None                               # [n = E, k = 4R], [n = R, k = 4R]: the same blocks as dgrad operands
        B, T = head.B, d.T
        self.x16 = torch.zeros(B * T, self.ldx, **f16)
        self.xg = torch.zeros(B * T, 4 * R, **f32)
        self.hg = torch.zeros(B, 4 * R, **f32)
        self.c = torch.zeros(B, R, **f32)
        self.h16 = torch.zeros(B, d.LDR, **f16)
        self.out = torch.zeros(B, T, R, **f32)
        self.tr = None                                           # training buffers, allocated on first use
        self.pack()

    @on_device
    def pack(self, params: Dict[str, torch.Tensor] = None, train: bool = False):
        """(re)build the fp16 operand copies from the fp32 variables (after every optimizer step when training)"""
        if params is not None:
            self.params = params
        R, E = self.h.d.R, self.E
        kern = self.params[KERNEL]
        self.wx[:, :E].copy_(kern[:E].t())
        self.wh[:, :R].copy_(kern[E:].t())
        self.bias[:4 * R].copy_(self.params[BIAS])
        if train or self.wxb is not None:
            if self.wxb is None:
                self.wxb = torch.zeros(E, self.ld4, dtype=torch.float16, device=self.h.device)
                self.whb = torch.zeros(R, self.ld4, dtype=torch.float16, device=self.h.device)
            self.wxb[:, :4 * R].copy_(kern[:E])
            self.whb[:, :4 * R].copy_(kern[E:])

    def _check(self, words, seq_len):
        h, B, T = self.h, self.h.B, self.h.d.T
        if tuple(words.shape) != (B, T) or tuple(seq_len.shape) != (B,) or words.device != h.device or seq_len.device != h.device:
            raise L.CmpcError(f"words must be [{B}, {T}] and seq_len [{B}] on {h.device}")
        return words.to(torch.int32), seq_len.to(torch.int32).contiguous()

    @on_device
    def forward(self, words: torch.Tensor, seq_len: torch.Tensor, train: bool = False) -> torch.Tensor:
        """words int32 [B, T] (token ids), seq_len int32 [B]  ->  lstm_outputs fp32 [B, T, R] (zero past seq_len)"""
        if train:
            return self._forward_train(words, seq_len)
        h, d, lib = self.h, self.h.d, self.h.lib
        B, T, R, st = h.B, d.T, d.R, h._stream()
        words, seq_len = self._check(words, seq_len)
        words = words.contiguous()
        h._ck(lib.cmpc_embed_gather_f16(words.data_ptr(), self.params[EMB].data_ptr(), self.V, self.E, B * T, self.x16.data_ptr(), self.ldx, st),
              "embed_gather")
        h._gemm(self.x16, self.E, self.wx, 4 * R, self.xg, bias=self.bias)                     # input half of every step at once
        self.c.zero_(); self.h16.zero_()
        for t in range(T):
            if t > 0:
                h._gemm(self.h16, R, self.wh, 4 * R, self.hg)                                  # recurrent half
            h._ck(lib.cmpc_lstm_step(self.xg.data_ptr(), self.hg.data_ptr() if t > 0 else None, seq_len.data_ptr(), t, T, R, B,
                                     self.c.data_ptr(), self.h16.data_ptr(), d.LDR, self.out.data_ptr(), st), "lstm_step")
        return self.out

    # ---- training ---------------------------------------------------------------------------------------------------------
    def _train_buffers(self):
        if self.tr is None:
            h, d = self.h, self.h.d
            B, T, R, dev = h.B, d.T, d.R, h.device
            f16 = dict(dtype=torch.float16, device=dev)
            f32 = dict(dtype=torch.float32, device=dev)
            self.tr = dict(cs=torch.zeros(T + 1, B, R, **f32), hs=torch.zeros(T + 1, B, d.LDR, **f16), gates=torch.zeros(T, B, 4 * R, **f32),
                           dz=torch.zeros(T * B, self.ld4, **f16), dC=torch.zeros(B, R, **f32), G=torch.zeros(B, d.LDR, **f32),
                           dx=torch.zeros(T * B, self.ldx, **f32), dkx=torch.zeros(self.E8, 4 * R, **f32), dkh=torch.zeros(d.LDR, 4 * R, **f32),
                           db=torch.zeros(4 * R, **f32))
            self.pack(train=True)
        return self.tr

    def _forward_train(self, words, seq_len):
        h, d, lib = self.h, self.h.d, self.h.lib
        B, T, R, st = h.B, d.T, d.R, h._stream()
        tr = self._train_buffers()
        words, seq_len = self._check(words, seq_len)
        self._ids_tm = words.t().contiguous()                                                  # time-major rows t * B + b
        self._seq_len = seq_len
        h._ck(lib.cmpc_embed_gather_f16(self._ids_tm.data_ptr(), self.params[EMB].data_ptr(), self.V, self.E, B * T, self.x16.data_ptr(), self.ldx, st),
              "embed_gather")
        h._gemm(self.x16, self.E, self.wx, 4 * R, self.xg, bias=self.bias)
        cs, hs, gates = tr["cs"], tr["hs"], tr["gates"]
        for t in range(T):
            if t > 0:
                h._gemm(hs[t], R, self.wh, 4 * R, self.hg)
            h._ck(lib.cmpc_lstm_step_train(self.xg.data_ptr(), self.hg.data_ptr() if t > 0 else None, seq_len.data_ptr(), t, T, R, B,
                                           cs[t].data_ptr(), cs[t + 1].data_ptr(), hs[t].data_ptr(), hs[t + 1].data_ptr(), d.LDR,
                                           gates[t].data_ptr(), self.out.data_ptr(), st), "lstm_step_train")
        return self.out

    @on_device
    def backward(self, d_out: torch.Tensor, grads: Dict[str, torch.Tensor]) -> None:
        """d_out fp32 [B, T, R] = d loss / d lstm_outputs (HeadBackward.backward); writes d Variable / d kernel / d bias into
        `grads` (tensors of the TF shapes, overwritten)."""
        h, d, lib = self.h, self.h.d, self.h.lib
        B, T, R, E, st = h.B, d.T, d.R, self.E, h._stream()
        tr, S = self.tr, GRAD_SCALE
        if tr is None or not hasattr(self, "_ids_tm"):
            raise L.CmpcError("WordEncoderB200.backward needs forward(..., train=True) first")
        d_out = d_out.to(torch.float32).contiguous()
        cs, gates, dz, G = tr["cs"], tr["gates"], tr["dz"], tr["G"]
        tr["db"].zero_(); tr["dkx"].zero_(); tr["dkh"].zero_(); tr["dC"].zero_()
        for t in range(T - 1, -1, -1):
            if t < T - 1:
                h._gemm(dz[(t + 1) * B:(t + 2) * B], 4 * R, self.whb, R, G)                    # S * dz_{t+1} K_h^T
            h._ck(lib.cmpc_lstm_step_bwd(d_out.data_ptr(), G.data_ptr() if t < T - 1 else None, d.LDR, self._seq_len.data_ptr(), t, T, R, B,
                                         gates[t].data_ptr(), cs[t].data_ptr(), cs[t + 1].data_ptr(), tr["dC"].data_ptr(), S,
                                         dz.data_ptr(), self.ld4, tr["db"].data_ptr(), st), "lstm_step_bwd")
        ck = lambda rc: h._ck(rc, "gemm_atb")
        ck(lib.cmpc_gemm_atb_f16(self.x16.data_ptr(), self.ldx, self.E8, dz.data_ptr(), self.ld4, 4 * R, T * B, tr["dkx"].data_ptr(), 4 * R, 0, st))
        ck(lib.cmpc_gemm_atb_f16(tr["hs"].data_ptr(), d.LDR, rup(R, 8), dz.data_ptr(), self.ld4, 4 * R, T * B, tr["dkh"].data_ptr(), 4 * R, 0, st))
        h._gemm(dz, 4 * R, self.wxb, E, tr["dx"])                                              # S * dz K_x^T  [T*B, E]
        gk = grads[KERNEL]
        torch.mul(tr["dkx"][:E], 1.0 / S, out=gk[:E])
        torch.mul(tr["dkh"][:R], 1.0 / S, out=gk[E:])
        grads[BIAS].copy_(tr["db"])
        ge = grads[EMB]
        ge.zero_()
        h._ck(lib.cmpc_embed_scatter_add(self._ids_tm.data_ptr(), tr["dx"].data_ptr(), self.ldx, 1.0 / S, self.V, E, T * B, ge.data_ptr(), st),
              "embed_scatter_add")
