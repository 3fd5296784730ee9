"""Generates tests/golden/head_tiny.pt: seeded inputs + outputs of the CPU oracle (oracle/cmpc_head_ref.py) for a small
head configuration.  The reference itself (TF-1.x, Python 2.7) cannot run in this image, so these vectors pin the
ORACLE (regression) and give the GPU tests a fixture that does not need the oracle at run time.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.cmpc_head_ref import HeadConfig, OracleHead, init_params, make_inputs  # noqa: E402

CFG = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, c4_dim=64, c3_dim=32, v_emb_dim=64, rnn_size=64,
           mlp_dim=32, parse_hidden=40)
KEEP = ["lateral_c5", "vis_la_sp_c4", "affi_c3", "gconv_y_c3", "spa_graph_c5", "fusion_c4", "exg2_c4", "convlstm_h1"]


def main():
    torch.set_num_threads(1)
    B = 3
    cfg = HeadConfig(batch_size=B, **CFG)
    params = init_params(cfg, seed=7, sharp=40.0, bias_std=0.05, ln_jitter=0.1)
    inp = make_inputs(cfg, B, seed=4321, seq_len=[20, 9, 2])
    head = OracleHead(params, cfg, keep=True)
    out = head.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
    blob = dict(cfg=CFG, batch=B, param_seed=7, param_kwargs=dict(sharp=40.0, bias_std=0.05, ln_jitter=0.1),
                inputs={k: inp[k] for k in ("c3", "c4", "c5", "lstm_outputs", "target_fine")},
                outputs={k: out[k] for k in ("pred", "words_parse", "gw_w", "gw_v", "up_c3", "up_c4", "up_c5")},
                intermediates={k: head.t[k] for k in KEEP})
    # inputs as fp16-exact values keep the file small and bit-identical across platforms
    path = Path(__file__).with_name("head_tiny.pt")
    torch.save(blob, path)
    print(path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
