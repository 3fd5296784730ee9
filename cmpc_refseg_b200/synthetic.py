"""Synthetic UNC-shaped inputs for benchmarks and smoke runs: the backbone taps are post-ReLU
(relu(N(0,1))), LSTM outputs have the LSTM h-range (tanh * sigmoid) and are zero past seq_len, targets are boxes.
Same generator (and same seeds -> same tensors) as oracle.cmpc_head_ref.make_inputs; kept here so that the product
path never imports the oracle."""
from __future__ import annotations

import torch


def make_inputs(batch: int, *, vf_h=40, vf_w=40, H=320, W=320, c3_dim=512, c4_dim=1024, vf_dim=2048, num_steps=20,
                rnn_size=1000, seed: int = 1234, seq_len=None, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    c3 = torch.relu(torch.randn(batch, vf_h, vf_w, c3_dim, generator=g)).to(dtype)
    c4 = torch.relu(torch.randn(batch, vf_h, vf_w, c4_dim, generator=g)).to(dtype)
    c5 = torch.relu(torch.randn(batch, vf_h, vf_w, vf_dim, generator=g)).to(dtype)
    lstm = (torch.tanh(torch.randn(batch, num_steps, rnn_size, generator=g)) *
            torch.sigmoid(torch.randn(batch, num_steps, rnn_size, generator=g))).to(dtype)
    if seq_len is None:
        sl = torch.full((batch,), num_steps, dtype=torch.int32)
    elif isinstance(seq_len, str) and seq_len == "unc":
        sl = torch.clamp(torch.poisson(torch.full((batch,), 3.5), generator=g), 1, num_steps).to(torch.int32)
    else:
        sl = torch.as_tensor(seq_len, dtype=torch.int32).reshape(-1).expand(batch).clone()
    lstm = lstm * (torch.arange(num_steps).view(1, num_steps) < sl.view(-1, 1)).to(dtype).unsqueeze(-1)
    target = torch.zeros(batch, H, W, 1, dtype=dtype)
    for b in range(batch):
        y0 = int(torch.randint(0, H // 2, (1,), generator=g))
        x0 = int(torch.randint(0, W // 2, (1,), generator=g))
        target[b, y0:y0 + H // 3 + 1, x0:x0 + W // 3 + 1, 0] = 1.0
    return dict(c3=c3, c4=c4, c5=c5, lstm_outputs=lstm, seq_len=sl, target_fine=target)
