"""Word encoder in front of the head (CMPC_model.py:144-157): GloVe lookup + LSTMCell(rnn_size) over the
expression, on the device.  Its output is the head's `lstm_outputs` input, so with it `words` / `seq_len` become the entry
signature of the drop-in, as in the reference.  Variables (below text_objseg/): `Variable` [vocab, glove_dim] (the embedding
matrix, :145), `rnn/lstm_cell/kernel` [glove_dim + rnn_size, 4 rnn_size], `rnn/lstm_cell/bias` [4 rnn_size].
Forward only: the backward pass of this package stops at d loss / d lstm_outputs."""
from __future__ import annotations

from typing import Dict

import torch

from . import _lib as L
from .weights import rup

EMB, KERNEL, BIAS = "Variable", "rnn/lstm_cell/kernel", "rnn/lstm_cell/bias"


class WordEncoderB200:
    def __init__(self, head, params: Dict[str, torch.Tensor]):
        self.h = head
        d, dev = head.d, head.device
        emb, kern, bias = (params[k].to(dev, torch.float32) for k in (EMB, KERNEL, BIAS))
        self.V, self.E = emb.shape
        R = d.R
        if tuple(kern.shape) != (self.E + R, 4 * R) or tuple(bias.shape) != (4 * R,):
            raise L.CmpcError(f"word LSTM variables: kernel {tuple(kern.shape)}, bias {tuple(bias.shape)} for glove_dim {self.E}, rnn_size {R}")
        self.emb = emb.contiguous()
        self.ldx = rup(self.E, 64)
        self.wx = torch.zeros(4 * R, self.ldx, dtype=torch.float16, device=dev)
        self.wx[:, :self.E].copy_(kern[:self.E].t())
        self.wh = torch.zeros(4 * R, d.LDR, dtype=torch.float16, device=dev)
        self.wh[:, :R].copy_(kern[self.E:].t())
        self.bias = torch.zeros(rup(4 * R, 256), dtype=torch.float32, device=dev)
        self.bias[:4 * R].copy_(bias)
        B, T = head.B, d.T
        self.x16 = torch.zeros(B * T, self.ldx, dtype=torch.float16, device=dev)
        self.xg = torch.zeros(B * T, 4 * R, dtype=torch.float32, device=dev)
        self.hg = torch.zeros(B, 4 * R, dtype=torch.float32, device=dev)
        self.c = torch.zeros(B, R, dtype=torch.float32, device=dev)
        self.h16 = torch.zeros(B, d.LDR, dtype=torch.float16, device=dev)
        self.out = torch.zeros(B, T, R, dtype=torch.float32, device=dev)

    def forward(self, words: torch.Tensor, seq_len: torch.Tensor) -> torch.Tensor:
        """words int32 [B, T] (token ids), seq_len int32 [B]  ->  lstm_outputs fp32 [B, T, R] (zero past seq_len)"""
        h, d, lib = self.h, self.h.d, self.h.lib
        B, T, R, st = h.B, d.T, d.R, h._stream()
        if tuple(words.shape) != (B, T) or tuple(seq_len.shape) != (B,) or words.device != h.device or seq_len.device != h.device:
            raise L.CmpcError(f"words must be [{B}, {T}] and seq_len [{B}] on {h.device}")
        words = words.to(torch.int32).contiguous()
        seq_len = seq_len.to(torch.int32).contiguous()
        h._ck(lib.cmpc_embed_gather_f16(words.data_ptr(), self.emb.data_ptr(), self.V, self.E, B * T, self.x16.data_ptr(), self.ldx, st), "embed_gather")
        h._gemm(self.x16, self.E, self.wx, 4 * R, self.xg, bias=self.bias)                     # input half of every step at once
        self.c.zero_(); self.h16.zero_()
        for t in range(T):
            if t > 0:
                h._gemm(self.h16, R, self.wh, 4 * R, self.hg)                                  # recurrent half
            h._ck(lib.cmpc_lstm_step(self.xg.data_ptr(), self.hg.data_ptr() if t > 0 else None, seq_len.data_ptr(), t, T, R, B,
                                     self.c.data_ptr(), self.h16.data_ptr(), d.LDR, self.out.data_ptr(), st), "lstm_step")
        return self.out
