"""Golden fixture: the oracle must reproduce tests/golden/head_tiny.pt (made by tests/golden/make_golden.py), and on a
GPU the sm_100a head must match the same fixture without touching the oracle at run time."""
from pathlib import Path

import pytest
import torch

GOLD = Path(__file__).parent / "golden" / "head_tiny.pt"


def _load():
    return torch.load(GOLD, weights_only=False)


def test_oracle_reproduces_golden_vectors():
    from oracle.cmpc_head_ref import HeadConfig, OracleHead, init_params
    g = _load()
    cfg = HeadConfig(batch_size=g["batch"], **g["cfg"])
    params = init_params(cfg, seed=g["param_seed"], **g["param_kwargs"])
    head = OracleHead(params, cfg, keep=True)
    i = g["inputs"]
    out = head.forward(i["c3"], i["c4"], i["c5"], i["lstm_outputs"])
    for k, v in g["outputs"].items():
        assert (out[k] - v).abs().max() < 2e-5, k          # same algorithm, possibly different BLAS threading
    for k, v in g["intermediates"].items():
        assert (head.t[k] - v).abs().max() < 2e-5, k


@pytest.mark.gpu
def test_gpu_head_matches_golden_vectors():
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from oracle.cmpc_head_ref import HeadConfig, init_params      # parameter generator only (weights are not in the fixture)
    g = _load()
    cfg = HeadConfig(batch_size=g["batch"], **g["cfg"])
    params = init_params(cfg, seed=g["param_seed"], **g["param_kwargs"])
    c = g["cfg"]
    dev = torch.device("cuda:0")
    model = LSTM_model(batch_size=g["batch"], num_steps=c["num_steps"], vf_h=c["vf_h"], vf_w=c["vf_w"], H=c["H"], W=c["W"],
                       vf_dim=c["vf_dim"], v_emb_dim=c["v_emb_dim"], rnn_size=c["rnn_size"], mlp_dim=c["mlp_dim"],
                       params=params, device=dev, head_kwargs=dict(c4_dim=c["c4_dim"], c3_dim=c["c3_dim"], parse_hidden=c["parse_hidden"]))
    i = {k: v.to(dev) for k, v in g["inputs"].items()}
    pred, up_c3, parse = model.run(["pred", "up_c3", "words_parse"],
                                   feed_dict=dict(visual_feat_c3=i["c3"], visual_feat_c4=i["c4"], visual_feat_c5=i["c5"],
                                                  lstm_outputs=i["lstm_outputs"]))
    torch.cuda.synchronize()
    o = g["outputs"]
    assert (pred.cpu() - o["pred"]).abs().max() <= 1e-2            # north-star logit tolerance
    assert (up_c3.cpu() - o["up_c3"]).abs().max() <= 1e-2
    assert (parse.cpu() - o["words_parse"]).abs().max() <= 1e-4
    assert (model.gw_w.cpu() - o["gw_w"]).abs().max() <= 5e-3 and (model.gw_v.cpu() - o["gw_v"]).abs().max() <= 5e-3
