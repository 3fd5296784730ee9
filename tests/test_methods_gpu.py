"""Op-level parity: every method of the reference's LSTM_model (CMPC_model.py:144-417) that the drop-in exposes, called on
its own with the reference's argument order, against the same-named method of the CPU oracle on the same inputs.

Tolerance: operands are rounded to fp16 (11-bit significand) on the way in, accumulation is fp32; outputs here are O(0.1-1)
so 4e-3 absolute covers the rounding with margin and still catches any layout / weight-slice / normalisation mistake."""
import numpy as np
import pytest
import torch

from oracle.cmpc_head_ref import HeadConfig, OracleHead, generate_spatial_batch, init_params, l2_normalize, make_inputs

pytestmark = pytest.mark.gpu

TOL = 4e-3
TINY = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, c4_dim=64, c3_dim=32, v_emb_dim=64, rnn_size=64,
            mlp_dim=32, parse_hidden=40)
ODD = dict(num_steps=12, vf_h=6, vf_w=10, H=48, W=80, vf_dim=64, c4_dim=64, c3_dim=32, v_emb_dim=72, rnn_size=72,
           mlp_dim=36, parse_hidden=44)            # mlp_dim % 8 != 0 like the reference's 500; h != w; T < 20


class Ctx:
    def __init__(self, cfg_kw, batch=2, seq_len=(9, 4)):
        from cmpc_refseg_b200.CMPC_model import LSTM_model
        self.cfg = HeadConfig(batch_size=batch, **cfg_kw)
        self.params = init_params(self.cfg, 0, sharp=8.0, bias_std=0.05, ln_jitter=0.1)
        self.dev = torch.device("cuda:0")
        hk = {k: cfg_kw[k] for k in ("c4_dim", "c3_dim", "parse_hidden")}
        mk = {k: v for k, v in cfg_kw.items() if k not in hk}
        self.model = LSTM_model(batch_size=batch, params=self.params, device=self.dev, head_kwargs=hk, **mk)
        self.ref = OracleHead(self.params, self.cfg)
        self.inp = make_inputs(self.cfg, batch, seed=7, seq_len=list(seq_len))
        self.B = batch
        g = torch.Generator().manual_seed(11)
        c = self.cfg
        self.rand = lambda *s: torch.randn(*s, generator=g)
        self.spatial = torch.from_numpy(generate_spatial_batch(batch, c.vf_h, c.vf_w)).float()
        self.wf, self.mask = self.ref.words(self.inp["lstm_outputs"])
        self.parse = self.ref.build_lang_parser(self.wf, self.mask)
        wf_d, _ = self.model.lstm(self.g(self.inp["lstm_outputs"]))
        assert (wf_d.cpu() - self.wf).abs().max() < 1e-5
        assert torch.equal(self.model.seq_mask.cpu(), self.mask)

    def g(self, t):
        return t.to(self.dev)

    def vis(self):
        c = self.cfg
        return l2_normalize(self.rand(self.B, c.vf_h, c.vf_w, c.v_emb_dim), 3)

    def feat(self):
        c = self.cfg
        return torch.relu(self.rand(self.B, c.vf_h, c.vf_w, c.mlp_dim)) * 0.3

    def lang(self):
        return l2_normalize(self.rand(self.B, 1, 1, self.cfg.rnn_size), 3)


@pytest.fixture(scope="module", params=["tiny", "odd"])
def ctx(request):
    return Ctx(TINY if request.param == "tiny" else ODD)


def _close(a, b, tol=TOL, what=""):
    a, b = a.float().cpu(), b.float().reshape(a.shape)
    err = float((a - b).abs().max())
    print(f"{what:40s} max-abs {err:.3e}  ref-absmax {float(b.abs().max()):.3e}")
    assert torch.isfinite(a).all() and err <= tol, f"{what}: {err}"


def test_lstm_tail(ctx):
    wf, lang = ctx.model.lstm(ctx.g(ctx.inp["lstm_outputs"]))
    _close(wf, ctx.wf, 1e-5, "words_feat")
    _close(lang, ctx.wf.sum(-2), 1e-4, "lang_feat (reduce_sum)")


def test_build_lang_parser(ctx):
    _close(ctx.model.build_lang_parser(ctx.g(ctx.wf)), ctx.parse, 2e-3, "words_parse")


def test_valid_and_nec_lang(ctx):
    # arbitrary (not parser-produced) word weights: the methods must honour their arguments
    parse = torch.softmax(ctx.rand(ctx.B, 1, ctx.cfg.num_steps, 4), 3) * ctx.mask
    _close(ctx.model.valid_lang(ctx.g(parse), ctx.g(ctx.wf)), ctx.ref.valid_lang(parse, ctx.wf), 1e-4, "valid_lang")
    _close(ctx.model.nec_lang(ctx.g(parse), ctx.g(ctx.wf)), ctx.ref.nec_lang(parse, ctx.wf), 1e-4, "nec_lang")


@pytest.mark.parametrize("level", ["c5_head1", "c4_head3", "c3_head5"])
def test_mutan_head(ctx, level):
    lang, vis = ctx.lang(), ctx.vis()
    out = ctx.model.mutan_head(ctx.g(lang), ctx.g(ctx.spatial), ctx.g(vis), level)
    _close(out, ctx.ref.mutan_head(lang, ctx.spatial, vis, level), TOL, f"mutan_head {level}")


@pytest.mark.parametrize("level", ["c5", "c4", "c3"])
def test_mutan_fusion(ctx, level):
    lang, vis = ctx.lang(), ctx.vis()
    out = ctx.model.mutan_fusion(ctx.g(lang), ctx.g(ctx.spatial), ctx.g(vis), level)
    _close(out, ctx.ref.mutan_fusion(lang, ctx.spatial, vis, level), TOL, f"mutan_fusion {level}")


def test_spatial_is_checked(ctx):
    from cmpc_refseg_b200._lib import CmpcError
    _close(ctx.model.generate_spatial_batch(), ctx.spatial, 1e-3, "generate_spatial_batch")
    with pytest.raises(CmpcError):
        ctx.model.mutan_fusion(ctx.g(ctx.lang()), ctx.g(ctx.spatial * 0.5), ctx.g(ctx.vis()), "c5")


@pytest.mark.parametrize("level", ["c5", "c3"])
def test_build_spa_graph_and_graph_conv(ctx, level):
    c = ctx.cfg
    spa = ctx.vis()
    ref = OracleHead(ctx.params, c, keep=True)
    r_out, r_w, r_v = ref.build_spa_graph(spa, ctx.wf, ctx.parse, ctx.mask, level)
    out = ctx.model.build_spa_graph(ctx.g(spa), ctx.g(ctx.wf), ctx.g(ctx.spatial), ctx.g(ctx.parse), level)
    _close(ctx.model.gw_w, r_w, TOL, f"gw_w {level}")
    _close(ctx.model.gw_v, r_v, TOL, f"gw_v {level}")
    _close(out, r_out, TOL, f"build_spa_graph {level}")
    # graph_conv on its own, with the oracle's own (gw_w, gw_v) as the factored adjacency
    N = c.n_nodes
    gf = spa.reshape(ctx.B, 1, N, c.v_emb_dim)
    r_gc = OracleHead(ctx.params, c).graph_conv(gf, r_w @ r_v.transpose(1, 2), level)
    gc = ctx.model.graph_conv(ctx.g(gf), N, c.v_emb_dim, (ctx.g(r_w), ctx.g(r_v)), graph_name="spa_graph", level=level)
    _close(gc, r_gc, 2e-2, f"graph_conv {level} (un-normalised, O(1) values)")
    from cmpc_refseg_b200._lib import CmpcError
    with pytest.raises(CmpcError):
        ctx.model.graph_conv(ctx.g(gf), N, c.v_emb_dim, ctx.g(r_w @ r_v.transpose(1, 2)), graph_name="spa_graph", level=level)


@pytest.mark.parametrize("level", ["c5", "c4", "c3"])
def test_build_lang2vis(ctx, level):
    vis = ctx.vis()
    r_out, r_w, r_v = ctx.ref.build_lang2vis(vis, ctx.wf, ctx.parse, ctx.mask, ctx.spatial, level)
    out = ctx.model.build_lang2vis(ctx.g(vis), ctx.g(ctx.wf), None, ctx.g(ctx.parse), ctx.g(ctx.spatial), level)
    _close(out, r_out, TOL, f"build_lang2vis {level}")
    _close(ctx.model.gw_w, r_w, TOL, "gw_w")


@pytest.mark.parametrize("module", ["c3", "c5", "c4_2"])
def test_global_vec_lang_se_exchange(ctx, module):
    f0, f1, f2, lang = ctx.feat(), ctx.feat(), ctx.feat(), ctx.lang()
    r_gv = ctx.ref.global_vec(f0, lang, module + "gv_f1")
    gv = ctx.model.global_vec(ctx.g(f0), ctx.g(lang), module + "gv_f1")
    _close(gv, r_gv, TOL, f"global_vec {module}")
    for f in ("_f1", "_f2"):
        _close(ctx.model.lang_se(ctx.g(f1), ctx.g(r_gv), module + f), ctx.ref.lang_se(f1, r_gv, module + f), TOL, f"lang_se {module}{f}")
    out = ctx.model.gated_exchange_module(ctx.g(f0), ctx.g(f1), ctx.g(f2), ctx.g(lang), module)
    _close(out, ctx.ref.gated_exchange_module(f0, f1, f2, lang, module), TOL, f"gated_exchange_module {module}")


def test_gated_exchange_fusion_lstm_2times(ctx):
    f3, f4, f5, lang = ctx.feat(), ctx.feat(), ctx.feat(), ctx.lang()
    out = ctx.model.gated_exchange_fusion_lstm_2times(ctx.g(f3), ctx.g(f4), ctx.g(f5), ctx.g(lang))
    _close(out, ctx.ref.gated_exchange_fusion_lstm_2times(f3, f4, f5, lang), TOL, "gated_exchange_fusion_lstm_2times")


def test_conv(ctx):
    c = ctx.cfg
    x = ctx.vis()
    out = ctx.model._conv("spa_graph_trans2_c4", ctx.g(x), 1, c.v_emb_dim, c.v_emb_dim, [1, 1, 1, 1])
    _close(out, ctx.ref._conv("spa_graph_trans2_c4", x), TOL, "_conv 1x1 CxC")
    hid = torch.relu(ctx.ref._conv("words_parse_1", ctx.wf))
    out = ctx.model._conv("words_parse_2", ctx.g(hid), 1, c.parse_hidden, 4, [1, 1, 1, 1])
    _close(out, ctx.ref._conv("words_parse_2", hid), TOL, "_conv 1x1 -> 4")
    f = ctx.feat()
    out = ctx.model._conv("score_c4", ctx.g(f), 3, c.mlp_dim, 1, [1, 1, 1, 1])
    _close(out, ctx.ref._conv("score_c4", f), TOL, "_conv 3x3 score")
    from cmpc_refseg_b200._lib import CmpcError
    with pytest.raises(CmpcError):
        ctx.model._conv("score_c4", ctx.g(f), 3, c.mlp_dim, 1, [1, 2, 2, 1])


def test_methods_compose_to_build_graph(ctx):
    """build_graph (:89-142) re-assembled from the per-method calls == the fused forward pass (same kernels, same order)"""
    m, c, inp = ctx.model, ctx.cfg, ctx.inp
    fused = {k: v.clone() for k, v in m.forward(ctx.g(inp["c3"]), ctx.g(inp["c4"]), ctx.g(inp["c5"]), ctx.g(inp["lstm_outputs"])).items()}
    words_feat, lang_feat = m.lstm(ctx.g(inp["lstm_outputs"]))
    spatial = m.generate_spatial_batch()
    words_parse = m.build_lang_parser(words_feat)
    fus = {}
    for lvl, cin in (("c5", c.vf_dim), ("c4", c.c4_dim), ("c3", c.c3_dim)):
        lat = m._conv(f"{lvl}_lateral", ctx.g(inp[lvl]), 1, cin, c.v_emb_dim, [1, 1, 1, 1])
        lat = ctx.g(l2_normalize(lat.cpu(), 3))
        fus[lvl] = m.build_lang2vis(lat, words_feat, lang_feat, words_parse, spatial, level=lvl)
    nec = m.nec_lang(words_parse, words_feat)
    fused_feat = m.gated_exchange_fusion_lstm_2times(fus["c3"], fus["c4"], fus["c5"], nec)
    pred = m._conv("score", fused_feat, 3, c.mlp_dim, 1, [1, 1, 1, 1])
    _close(pred, fused["pred"].cpu(), 5e-3, "composed pred vs fused forward")
    _close(m.gw_w, fused["gw_w"].cpu(), 2e-3, "gw_w (level c3 is the last one built)")


@pytest.mark.parametrize("shape", ["tiny", "full"])
def test_word_encoder(shape):
    """embedding lookup + word LSTM of lstm() (CMPC_model.py:144-157) on the device vs the oracle's word_lstm; the recurrent
    product runs with fp16 operands over up to 20 steps, outputs are in (-1, 1): 5e-3 absolute.  Then the same pass fed with
    words / seq_len instead of lstm_outputs."""
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from oracle.cmpc_head_ref import word_lstm
    dev = torch.device("cuda:0")
    if shape == "tiny":
        kw, V, E, seq = TINY, 50, 20, [9, 4]
    else:
        kw = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, c4_dim=64, c3_dim=32, v_emb_dim=1000, rnn_size=1000,
                  mlp_dim=32, parse_hidden=40)                # the reference's encoder: 300-d GloVe, 1000-d cell, 20 steps
        V, E, seq = 2000, 300, [20, 1]
    B = 2
    cfg = HeadConfig(batch_size=B, **kw)
    params = init_params(cfg, 0, sharp=8.0, bias_std=0.05, ln_jitter=0.1)
    g = torch.Generator().manual_seed(5)
    R = cfg.rnn_size
    enc = {"Variable": torch.randn(V, E, generator=g) * 0.5,
           "rnn/lstm_cell/kernel": (torch.rand(E + R, 4 * R, generator=g) * 2 - 1) * (6.0 / (E + 5 * R)) ** 0.5 * 2,
           "rnn/lstm_cell/bias": torch.randn(4 * R, generator=g) * 0.1}
    words = torch.randint(0, V, (B, cfg.num_steps), generator=g)
    seq_len = torch.tensor(seq)
    want = word_lstm(words, seq_len, *(enc[k].double() for k in ("Variable", "rnn/lstm_cell/kernel", "rnn/lstm_cell/bias"))).float()
    hk = {k: kw[k] for k in ("c4_dim", "c3_dim", "parse_hidden")}
    mk = {k: v for k, v in kw.items() if k not in hk}
    model = LSTM_model(batch_size=B, params={**params, **enc}, device=dev, head_kwargs=hk, **mk)
    got = model.encode_words(words.to(dev), seq_len.to(dev)).cpu()
    assert want.abs().max() > 0.05
    _close(got, want, 5e-3, "lstm_outputs")
    for b in range(B):
        assert torch.all(got[b, seq[b]:] == 0)
    inp = make_inputs(cfg, B, seed=7, seq_len=seq)
    c3, c4, c5 = (inp[k].to(dev) for k in ("c3", "c4", "c5"))
    a = model.forward(c3, c4, c5, lstm_outputs=got.to(dev), seq_len=seq_len.to(dev))["up"].clone()
    b_ = model.forward(c3, c4, c5, words=words.to(dev), seq_len=seq_len.to(dev))["up"]
    assert (a - b_).abs().max() < 1e-3            # same inputs; the pass has fp32 atomics (split-K, pooled sums), so not bit-equal
    with pytest.raises(Exception):
        LSTM_model(batch_size=B, params=params, device=dev, head_kwargs=hk, **mk).encode_words(words.to(dev), seq_len.to(dev))
