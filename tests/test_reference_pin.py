"""The oracle (oracle/cmpc_head_ref.py) against the REFERENCE ITSELF: the reference's own CMPC_model.py / util/cell.py /
util/loss.py / util/processing_tools.py executed unmodified through the eager TensorFlow stand-in (oracle/tfshim).

* fixtures (tests/golden/ref_*.npz, made here by tests/golden/make_ref_golden.py) travel to the GPU box, where the
  reference checkout does not exist;
* the `live` tests re-run the reference in this process when the checkout is present (this container) on inputs the
  fixtures do not contain, in float64 (wiring differences cannot hide behind rounding).

What this pins: every tensor-to-op connection, variable name / scope / shape, mask, reshape, loss weight and the optimizer
recipe of CMPC_model.py:89-142,144-164,166-417,426-492 and util/cell.py:36-79.  What it cannot pin: the arithmetic inside each
tf.* op, which the stand-in restates from TF-1's published definitions (listed in oracle/tfshim/tensorflow/__init__.py).
"""
import numpy as np
import pytest
import torch

from oracle.cmpc_head_ref import HeadConfig, OracleHead, init_params, make_inputs, word_lstm
from oracle.ref_runner import reference_available, run_reference

import refgold

F64 = torch.float64
live = pytest.mark.skipif(not reference_available(), reason="reference checkout not present (GPU box): fixtures only")


def _oracle(params, cfg, inp, B, dtype, **kw):
    gv = "batch" if B > 1 else "sample"        # tf.nn.l2_normalize(gv_lang) has no axis (CMPC_model.py:241)
    head = OracleHead(params, cfg, gv_norm=gv, **kw)
    return head, head.forward(inp["c3"].to(dtype), inp["c4"].to(dtype), inp["c5"].to(dtype), inp["lstm_outputs"].to(dtype))


@pytest.mark.parametrize("name", ["ref_tiny_b1", "ref_tiny_b3", "ref_cfg1_random", "ref_cfg1_sharp"])
def test_oracle_matches_reference_wiring(name):
    """float64 oracle vs the float64-executed reference graph (stored as float32): <= 1e-5 on every public output"""
    kw, B, cfg, params, inp, fix = refgold.forward_case(name, F64)
    _, out = _oracle(params, cfg, inp, B, F64)
    for k in ("pred", "up", "sigm", "up_c3", "up_c4", "up_c5", "words_parse", "gw_w", "gw_v", "seq_mask"):
        d = float((out[k].reshape(fix[k].shape).float() - fix[k]).abs().max())
        assert d <= 1e-5, (name, k, d)
    # the float32 execution of the oracle stays within float32 noise of the float32 execution of the reference
    p32 = {k: v.float() for k, v in params.items()}
    _, o32 = _oracle(p32, cfg, inp, B, torch.float32)
    assert float((o32["pred"] - fix["pred_f32run"]).abs().max()) <= 2e-5


def test_oracle_matches_reference_at_4096_nodes():
    """BASELINE configs[3] geometry (512 x 512 -> 64 x 64 maps, N = 4096), one sample"""
    kw, B, cfg, params, inp, fix = refgold.forward_case("ref_hires_b1", torch.float32)
    _, out = _oracle(params, cfg, inp, B, torch.float32, dense_adj=False)       # (the dense 4096^2 path is what made the fixture)
    assert float((out["pred"] - fix["pred"]).abs().max()) <= 2e-5
    assert float((out["words_parse"] - fix["words_parse"]).abs().max()) <= 1e-6


def test_oracle_matches_reference_at_benchmark_batch():
    """BASELINE configs[1] (batch 32, N = 1600, UNC-shaped sentence lengths), the literal batch-coupled graph"""
    kw, B, cfg, params, inp, fix = refgold.forward_case("ref_cfg2_b32", torch.float32)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    _, out = _oracle(params, cfg, inp, B, torch.float32)
    assert float((out["pred"] - fix["pred"]).abs().max()) <= 2e-5
    assert float((out["words_parse"] - fix["words_parse"]).abs().max()) <= 1e-6
    assert torch.equal(out["seq_mask"], fix["seq_mask"])
    assert (inp["seq_len"].numpy() == fix["seq_len"]).all() and len(set(fix["seq_len"].tolist())) > 3


@pytest.mark.parametrize("name", ["ref_tiny_train", "ref_tiny_train_mild"])
def test_oracle_gradients_match_reference_train_op(name):
    """d cost / d every trainable variable: torch.autograd through the oracle vs compute_gradients of the reference's train_op"""
    kw, B, cfg, params, inp, fix = refgold.train_case(F64, name)
    pp = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    head = OracleHead(pp, cfg, gv_norm="batch")
    out = head.forward(inp["c3"].double(), inp["c4"].double(), inp["c5"].double(), inp["lstm_outputs"].double())
    L = head.losses(out, inp["target_fine"].double())
    for k in ("cls_loss", "cls_loss_c3", "cls_loss_c4", "cls_loss_c5", "cls_loss_all", "reg_loss", "cost"):
        assert abs(float(L[k].detach()) - float(fix[k])) <= 1e-9 * abs(float(fix[k])), k
    names = list(pp)
    grads = torch.autograd.grad(L["cost"], [pp[k] for k in names])
    assert set(names) == {k[5:] for k in fix if k.startswith("grad/")}
    for k, g in zip(names, grads):
        r = torch.from_numpy(fix["grad/" + k]).double()
        if float(r.norm()) < 1e-12:                      # analytically zero (key-conv bias under softmax shift invariance)
            assert float(g.norm()) < 1e-12, k
        else:
            assert float((g - r).norm() / r.norm()) <= 1e-6, k
    # the reference's recipe: DW regularised inside the cost, bias gradients doubled (CMPC_model.py:464-475), then Adam; on the
    # first step Adam moves every coordinate by lr * g / (|g| + eps / sqrt(1 - b2)) ~ lr * sign(g)
    lr = float(fix["learning_rate"])
    assert abs(lr - 0.00025) < 1e-12
    for k in ("c5_lateral/DW", "c5_lateral/biases", "rnn/conv_lstm_cell/W_ci", "gconv_feat_ln_spa_graph_c4/gamma"):
        g, s = fix["grad/" + k].astype(np.float64), fix["step/" + k].astype(np.float64)
        mult = 2.0 if k.endswith("biases") else 1.0
        gg = mult * g
        expect = -lr * gg / (np.abs(gg) + 1e-8 / np.sqrt(1 - 0.999))
        big = np.abs(gg) > 1e-6
        assert np.allclose(s[big], expect[big], rtol=2e-3, atol=1e-9), k


def test_word_encoder_matches_reference_lstm_front():
    """lstm() (CMPC_model.py:144-164) -> head, forward and backward, from token ids"""
    kw, B, cfg, params, inp, fix = refgold.words_case(F64)
    emb, kernel, bias = (torch.from_numpy(fix[k]).double().requires_grad_(True) for k in ("embedding", "kernel", "bias"))
    words, sl = torch.from_numpy(fix["words"]), torch.from_numpy(fix["seq_len"])
    outs = word_lstm(words, sl, emb, kernel, bias)
    pp = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    head = OracleHead(pp, cfg, gv_norm="batch")
    out = head.forward(inp["c3"].double(), inp["c4"].double(), inp["c5"].double(), outs)
    assert float((out["pred"].float() - torch.from_numpy(fix["pred"])).abs().max()) <= 1e-5
    assert torch.equal(out["seq_mask"].float(), torch.from_numpy(fix["seq_mask"]))
    cost = head.losses(out, inp["target_fine"].double())["cost"]
    assert abs(float(cost) - float(fix["cost"])) <= 1e-9 * float(fix["cost"])
    g_emb, g_k, g_b, g_p1, g_c5 = torch.autograd.grad(cost, [emb, kernel, bias, pp["words_parse_1/DW"], pp["c5_lateral/DW"]])
    for g, k in ((g_emb, "Variable"), (g_k, "rnn/lstm_cell/kernel"), (g_b, "rnn/lstm_cell/bias"), (g_p1, "words_parse_1/DW"),
                 (g_c5, "c5_lateral/DW")):
        r = torch.from_numpy(fix["grad/" + k])
        assert float((g - r).norm() / r.norm()) <= 1e-9, k


# ---------------------------------------------------------------------------------------------------------------------
# live: the reference re-executed in this process on fresh inputs
# ---------------------------------------------------------------------------------------------------------------------
@live
@pytest.mark.parametrize("B,seq_len,seed", [(1, [13], 11), (2, [20, 1], 12), (4, "unc", 13)])
def test_live_reference_forward(B, seq_len, seed):
    kw = dict(num_steps=20, vf_h=6, vf_w=10, H=48, W=80, vf_dim=96, v_emb_dim=48, rnn_size=48, mlp_dim=24)   # non-square map
    cfg = HeadConfig(batch_size=B, c4_dim=1024, c3_dim=512, parse_hidden=500, **kw)
    params = init_params(cfg, seed=seed, dtype=F64, sharp=30.0, bias_std=0.1, ln_jitter=0.2)
    inp = make_inputs(cfg, B, seed=seed + 100, seq_len=seq_len)
    ref = run_reference(dict(batch_size=B, mode="eval", **kw), params, inp["c3"], inp["c4"], inp["c5"],
                        lstm_outputs=inp["lstm_outputs"], float64=True)
    assert set(ref["variables"]) - {"Variable"} == set(params)           # same variable names
    for k, v in params.items():
        assert tuple(ref["variables"][k].shape) == tuple(v.shape), k
    for dense in (True, False):
        _, out = _oracle(params, cfg, inp, B, F64, dense_adj=dense)
        for k in ("pred", "up", "sigm", "up_c3", "up_c4", "up_c5", "words_parse", "gw_w", "gw_v", "seq_mask"):
            assert float((out[k].reshape(ref[k].shape) - ref[k]).abs().max()) <= 1e-10, (k, dense)
    if B > 1:     # the per-sample variant the device uses for sharded inference == the reference run one sample at a time
        head = OracleHead(params, cfg, gv_norm="sample")
        o = head.forward(inp["c3"].double(), inp["c4"].double(), inp["c5"].double(), inp["lstm_outputs"].double())
        for b in range(B):
            r1 = run_reference(dict(batch_size=1, mode="eval", **kw), params, inp["c3"][b:b + 1], inp["c4"][b:b + 1], inp["c5"][b:b + 1],
                               lstm_outputs=inp["lstm_outputs"][b:b + 1], float64=True)
            assert float((o["pred"][b:b + 1] - r1["pred"]).abs().max()) <= 1e-10


@live
def test_live_reference_initializers_and_param_count():
    """the reference's own get_variable initialisers: shapes, xavier range, zero biases, LN gamma 1 / beta 0; parameter count"""
    kw = dict(num_steps=20, vf_h=40, vf_w=40, H=320, W=320)
    B = 1
    cfg = HeadConfig(batch_size=B)
    z = lambda c: torch.zeros(B, 2, 2, c)
    # tiny spatial map, full channel widths: the variables (except the [h, w, M] peepholes) do not depend on the map size
    ref = run_reference(dict(batch_size=B, mode="eval", num_steps=20, vf_h=2, vf_w=2, H=16, W=16), None, z(512), z(1024), z(2048),
                        lstm_outputs=torch.zeros(B, 20, 1000))
    shapes = {k: tuple(v.shape) for k, v in ref["variables"].items() if k != "Variable"}
    from oracle.cmpc_head_ref import param_shapes
    want = param_shapes(HeadConfig(batch_size=B, vf_h=2, vf_w=2, H=16, W=16))
    assert shapes == {k: tuple(v) for k, v in want.items()}
    n40 = sum(int(np.prod(s)) for s in param_shapes(cfg).values())
    assert n40 == 67_218_008
    v = ref["variables"]
    lim = (6.0 / (2048 + 1000)) ** 0.5
    assert float(v["c5_lateral/DW"].abs().max()) <= lim and float(v["c5_lateral/DW"].abs().max()) > 0.99 * lim
    assert float(v["c5_lateral/biases"].abs().max()) == 0 and float(v["rnn/conv_lstm_cell/LayerNorm_3/beta"].abs().max()) == 0
    assert float((v["gconv_feat_ln_spa_graph_c3/gamma"] - 1).abs().max()) == 0


@live
def test_live_reference_train_op_gradients():
    kw = dict(num_steps=12, vf_h=5, vf_w=7, H=40, W=56, vf_dim=64, v_emb_dim=32, rnn_size=32, mlp_dim=16)
    B = 2
    cfg = HeadConfig(batch_size=B, c4_dim=1024, c3_dim=512, parse_hidden=500, **kw)
    params = init_params(cfg, seed=5, dtype=F64, sharp=25.0, bias_std=0.1, ln_jitter=0.2)
    inp = make_inputs(cfg, B, seed=77, seq_len=[12, 4], dtype=F64)
    ref = run_reference(dict(batch_size=B, mode="train", weight_decay=0.002, **kw), params, inp["c3"], inp["c4"], inp["c5"],
                        lstm_outputs=inp["lstm_outputs"], target_fine=inp["target_fine"], float64=True)
    cfg.weight_decay = 0.002
    pp = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    head = OracleHead(pp, cfg, gv_norm="batch")
    out = head.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
    L = head.losses(out, inp["target_fine"])
    assert abs(float(L["cost"]) - float(ref["cost"])) <= 1e-10 * float(ref["cost"])
    from oracle.cmpc_head_ref import mask_iu
    I, U = mask_iu(out["up"], inp["target_fine"])
    assert abs(float((I.double() / U.double()).mean()) - float(ref["mIoU"])) < 1e-12         # CMPC_model.py:486-490
    grads = torch.autograd.grad(L["cost"], list(pp.values()))
    for (k, _), g in zip(pp.items(), grads):
        r = ref["raw_grads"][k]
        if float(r.norm()) < 1e-12:
            assert float(g.norm()) < 1e-12, k
        else:
            assert float((g - r).norm() / r.norm()) <= 1e-8, k
        a = ref["applied_grads"][k]
        assert torch.equal(a, r * (2.0 if k.endswith("biases") else 1.0)), k


@live
def test_live_reference_gradients_are_the_finite_differences_of_the_reference_forward():
    """Not autograd-of-a-restatement only: central differences of the REFERENCE's own `cost` (its forward executed from its source,
    float64) on three scalars per stage of the head reproduce what its `compute_gradients` returned."""
    kw = dict(num_steps=8, vf_h=4, vf_w=5, H=32, W=40, vf_dim=48, v_emb_dim=24, rnn_size=24, mlp_dim=12)
    B = 2
    cfg = HeadConfig(batch_size=B, c4_dim=1024, c3_dim=512, parse_hidden=500, **kw)
    params = init_params(cfg, seed=9, dtype=F64, sharp=8.0, bias_std=0.1, ln_jitter=0.2)
    inp = make_inputs(cfg, B, seed=123, seq_len=[8, 3], dtype=F64)

    def run(p):
        return run_reference(dict(batch_size=B, mode="train", **kw), p, inp["c3"], inp["c4"], inp["c5"],
                             lstm_outputs=inp["lstm_outputs"], target_fine=inp["target_fine"], float64=True)
    base = run(params)
    stages = ["c5_lateral/DW", "vis_trans_c4_head2/DW", "lang_trans_c3_head5/biases", "words_trans_c5/DW", "spa_graph_trans2_c4/DW",
              "gconv_feat_ln_spa_graph_c3/gamma", "gconv_update_spa_graph_c5/DW", "fusion_c4/DW", "words_parse_1/DW", "words_parse_2/biases",
              "lang_query_c3gv_f1/DW", "gv_lang_c4_2gv_f1/DW", "lang_feat_c5_f2/DW", "trans_feat_c3_2_f1/DW",
              "rnn/conv_lstm_cell/kernel", "rnn/conv_lstm_cell/W_co", "rnn/conv_lstm_cell/LayerNorm_2/beta", "score/DW", "score_c4/DW"]
    g = torch.Generator().manual_seed(0)
    eps, worst = 1e-5, 0.0
    for k in stages:
        grad = base["raw_grads"][k]
        scale = float(grad.abs().max())
        flat = grad.flatten()
        big = torch.nonzero(flat.abs() > 0.05 * scale).flatten()             # scalars whose derivative is well above the FD noise
        for idx in big[torch.randperm(len(big), generator=g)[:3]].tolist():
            vals = []
            for sgn in (+1, -1):
                p = dict(params)
                t = params[k].clone()
                t.view(-1)[idx] += sgn * eps
                p[k] = t
                vals.append(float(run(p)["cost"]))
            fd = (vals[0] - vals[1]) / (2 * eps)
            rel = abs(fd - float(flat[idx])) / max(abs(float(flat[idx])), 1e-12)
            worst = max(worst, rel)
            assert rel < 2e-4, (k, idx, fd, float(flat[idx]))
    print(f"finite differences of the reference cost vs compute_gradients: worst relative deviation {worst:.2e} over {3 * len(stages)} scalars")
