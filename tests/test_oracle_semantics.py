"""The oracle is the checker for every parity claim, and the reference ships no golden vectors to pin it (TF-1 is not
installable here): these tests pin each TF-1 semantic the oracle restates against hand-computed
values, and the invariants the reference itself states."""
import math

import numpy as np
import pytest
import torch

from oracle import cmpc_head_ref as R

TINY = dict(num_steps=6, vf_h=4, vf_w=5, H=32, W=40, vf_dim=16, c4_dim=12, c3_dim=8, v_emb_dim=24, rnn_size=24,
            mlp_dim=12, parse_hidden=10)


def test_resize_bilinear_legacy_hand_computed():
    # TF-1 legacy (align_corners=False, no half-pixel centres): src = dst * in/out
    x = torch.tensor([[1.0, 3.0], [5.0, 9.0]]).view(1, 2, 2, 1)
    y = R.resize_bilinear_legacy(x, 4, 4)[0, :, :, 0]
    expect = torch.tensor([[1, 2, 3, 3], [3, 4.5, 6, 6], [5, 7, 9, 9], [5, 7, 9, 9]], dtype=torch.float32)
    assert torch.equal(y, expect)
    # and it is NOT torch's half-pixel bilinear
    t = torch.nn.functional.interpolate(x.permute(0, 3, 1, 2), size=(4, 4), mode="bilinear", align_corners=False)[0, 0]
    assert not torch.allclose(t, expect)


def test_resize_x8_samples_are_multiples_of_one_eighth():
    x = torch.arange(20, dtype=torch.float32).view(1, 4, 5, 1)
    y = R.resize_bilinear_legacy(x, 32, 40)[0, :, :, 0]
    assert y[0, 0] == 0 and y[0, 8] == 1 and y[8, 0] == 5
    assert y[0, 4] == pytest.approx(0.5) and y[4, 0] == pytest.approx(2.5)
    assert y[31, 39] == 19          # clamped at the last source pixel


def test_layer_norm_whole_sample_statistics():
    x = torch.tensor([[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], [[0.0, 0.0, 0.0], [6.0, 6.0, 6.0]]])   # [B=2, 2, 3]
    g = torch.tensor([1.0, 2.0, 3.0]); b = torch.tensor([0.0, 0.5, 1.0])
    y = R.layer_norm_tf(x, g, b)
    m0, v0 = 3.5, np.var([1, 2, 3, 4, 5, 6])      # biased variance over ALL non-batch axes
    exp0 = (np.array([[1, 2, 3], [4, 5, 6]]) - m0) / math.sqrt(v0 + 1e-12) * np.array([1, 2, 3]) + np.array([0, .5, 1])
    assert np.allclose(y[0].numpy(), exp0, atol=1e-6)
    assert np.allclose(y[1].numpy(), (np.array([[0, 0, 0], [6, 6, 6]]) - 3) / 3 * np.array([1, 2, 3]) + np.array([0, .5, 1]), atol=1e-6)


def test_l2_normalize_zero_row_and_axis_none():
    x = torch.tensor([[3.0, 4.0], [0.0, 0.0]])
    y = R.l2_normalize(x, -1)
    assert torch.allclose(y[0], torch.tensor([0.6, 0.8])) and torch.equal(y[1], torch.zeros(2))   # zero vector stays zero
    z = R.l2_normalize(x, None)                    # axis=None: every axis, batch included (CMPC_model.py:241)
    assert torch.allclose(z, x / 5.0)


def test_masked_softmax_with_float32_min():
    affi = torch.tensor([[[0.3, -0.2, 0.7]]])                      # [B=1, N=1, T=3]
    mask = torch.tensor([[[1.0, 1.0, 0.0]]])
    w = torch.softmax(mask * affi + (1 - mask) * R.FLT_MIN_TF, dim=2)
    e = np.exp([0.3, -0.2]); e /= e.sum()
    assert np.allclose(w[0, 0].numpy(), [e[0], e[1], 0.0], atol=1e-7)
    assert w[0, 0, 2] == 0.0                                       # exactly zero, not merely small


def test_spatial_batch_values():
    s = R.generate_spatial_batch(2, 4, 5)
    assert s.shape == (2, 4, 5, 8) and s.dtype == np.float32
    # pixel (h=1, w=2): xmin = 2/5*2-1, ymin = 1/4*2-1, ...
    exp = [2 / 5 * 2 - 1, 1 / 4 * 2 - 1, 3 / 5 * 2 - 1, 2 / 4 * 2 - 1, (2 / 5 * 2 - 1 + 3 / 5 * 2 - 1) / 2, (1 / 4 * 2 - 1 + 2 / 4 * 2 - 1) / 2, 1 / 5, 1 / 4]
    assert np.allclose(s[1, 1, 2], np.array(exp, dtype=np.float32))


def test_conv3x3_same_is_cross_correlation_with_zero_pad():
    x = torch.arange(9, dtype=torch.float32).view(1, 3, 3, 1)
    w = torch.zeros(3, 3, 1, 1); w[0, 0, 0, 0] = 1.0              # picks the top-left neighbour (no kernel flip)
    y = R.conv2d_same(x, w, torch.tensor([0.5]))[0, :, :, 0]
    expect = torch.tensor([[0, 0, 0], [0, 0, 1], [0, 3, 4]], dtype=torch.float32) + 0.5
    assert torch.equal(y, expect)


def test_sigmoid_ce_formula():
    x = torch.tensor([-3.0, 0.0, 2.5]); z = torch.tensor([1.0, 0.0, 1.0])
    ref = -(z * torch.log(torch.sigmoid(x)) + (1 - z) * torch.log(1 - torch.sigmoid(x)))
    assert torch.allclose(R.sigmoid_ce_with_logits(x, z), ref, atol=1e-6)


def test_param_count_matches_reference():
    # 67 218 008 head parameters at the default dims
    assert sum(int(np.prod(s)) for s in R.param_shapes(R.HeadConfig()).values()) == 67218008


@pytest.fixture(scope="module")
def tiny_run():
    cfg = R.HeadConfig(batch_size=2, **TINY)
    p = R.init_params(cfg, 0, sharp=30.0, bias_std=0.05, ln_jitter=0.1)
    inp = R.make_inputs(cfg, 2, seq_len=[6, 3])
    head = R.OracleHead(p, cfg, keep=True, gv_norm="sample")
    out = head.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
    return cfg, p, inp, head, out


def test_invariants_stated_by_the_reference(tiny_run):
    cfg, p, inp, head, out = tiny_run
    # words_parse rows sum to seq_mask (CMPC_model.py:352-353)
    assert torch.allclose(out["words_parse"].sum(3, keepdim=True), out["seq_mask"], atol=1e-6)
    # padded words: zero features, zero mask
    assert out["seq_mask"][1, 0, 3:].abs().sum() == 0
    # adjacency rows sum to 1 (the comment at CMPC_model.py:401)
    adj = out["gw_w"] @ out["gw_v"].transpose(1, 2)
    assert torch.allclose(adj.sum(2), torch.ones(2, cfg.n_nodes), atol=1e-5)
    # gw_v is exactly zero on padded words, gw_w too
    assert out["gw_v"][1, :, 3:].abs().sum() == 0 and out["gw_w"][1, :, 3:].abs().sum() == 0


def test_dense_adjacency_equals_low_rank_path(tiny_run):
    cfg, p, inp, head, out = tiny_run
    low = R.OracleHead(p, cfg, dense_adj=False).forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
    assert (low["pred"] - out["pred"]).abs().max() < 1e-5


def test_fp32_agrees_with_fp64(tiny_run):
    cfg, p, inp, head, out = tiny_run
    p64 = {k: v.double() for k, v in p.items()}
    o64 = R.OracleHead(p64, cfg).forward(*(inp[k].double() for k in ("c3", "c4", "c5", "lstm_outputs")))
    assert (o64["pred"].float() - out["pred"]).abs().max() < 1e-4


def test_sample_norm_equals_reference_at_batch_one(tiny_run):
    """gv_norm='sample' on a batch == the literal axis=None graph run per sample at B=1 (how the reference's drivers run)."""
    cfg, p, inp, head, out = tiny_run
    for b in range(2):
        cfg1 = R.HeadConfig(batch_size=1, **TINY)
        o1 = R.OracleHead(p, cfg1, gv_norm="batch").forward(*(inp[k][b:b + 1] for k in ("c3", "c4", "c5", "lstm_outputs")))
        assert (o1["pred"] - out["pred"][b:b + 1]).abs().max() < 1e-5


def test_batch_coupled_norm_differs_at_b2(tiny_run):
    cfg, p, inp, head, out = tiny_run
    lit = R.OracleHead(p, cfg, gv_norm="batch").forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
    assert (lit["pred"] - out["pred"]).abs().max() > 1e-6      # App. D-1: the literal graph couples samples at B > 1


def test_mask_iu_integer_counts():
    up = torch.tensor([[[[0.5], [-1.0]], [[0.0], [2.0]]]])        # [1,2,2,1]
    tgt = torch.tensor([[[[1.0], [1.0]], [[0.0], [0.0]]]])
    I, U = R.mask_iu(up, tgt)
    assert I.tolist() == [1] and U.tolist() == [3]                 # up > 0 is strict: 0.0 is background
    I2, U2 = R.mask_iu(up, tgt, thresh=1e-9, strict=False)         # host driver variant (trainval_model.py:244)
    assert I2.tolist() == [1] and U2.tolist() == [3]
