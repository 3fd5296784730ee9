"""GPU parity tests proper: the sm_100a head (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): logits max-abs <= 1e-2, thresholded-mask agreement >= 99.9 %.
Arithmetic on device: fp16 operands (11-bit significand), fp32 accumulate in TMEM, fp32 epilogues,
fp64 layer-norm statistics.
"""
import pytest
import torch

from oracle.cmpc_head_ref import HeadConfig, OracleHead, init_params, make_inputs, mask_iu

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-2          # north_star: max-abs on logits
MASK_AGREE = 0.999        # north_star: pixel agreement of (logit > 0)

TINY = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, c4_dim=64, c3_dim=32, v_emb_dim=64, rnn_size=64,
            mlp_dim=32, parse_hidden=40)


def _run(cfg_kw, batch, *, sharp=1.0, bias_std=0.0, ln_jitter=0.0, seq_len=None, keep=True, seed=1234, aux=False):
    from cmpc_refseg_b200.head import CMPCHeadB200
    cfg = HeadConfig(batch_size=batch, **cfg_kw)
    params = init_params(cfg, 0, sharp=sharp, bias_std=bias_std, ln_jitter=ln_jitter)
    inp = make_inputs(cfg, batch, seed=seed, seq_len=seq_len)
    # oracle per sample at B=1 (== gv_norm='sample'), SURVEY 8(e)
    ref = OracleHead(params, cfg, keep=keep)
    ro = ref.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
    dev = torch.device("cuda:0")
    head = CMPCHeadB200(params, batch_size=batch, device=dev, **cfg_kw)
    out = head.forward(inp["c3"].to(dev), inp["c4"].to(dev), inp["c5"].to(dev), inp["lstm_outputs"].to(dev), keep=keep, aux=aux)
    torch.cuda.synchronize()
    return cfg, ref, ro, head, {k: v.float().cpu() for k, v in out.items()}, inp


def _report(ref, head, names):
    rows = []
    for k in names:
        if k in head.t and k in ref.t:
            a, b = head.t[k].cpu(), ref.t[k].reshape(head.t[k].shape)
            rows.append(f"{k:16s} max-abs {float((a - b).abs().max()):.3e}   ref-std {float(b.std()):.3e}")
    return "\n".join(rows)


STAGES = ["valid_lang", "nec_lang"] + [f"{s}_{l}" for l in ("c5", "c4", "c3") for s in
          ("lateral", "vis_la_sp", "affi", "gw_w", "gw_v", "gconv_y", "spa_graph", "fusion")] + \
         ["exg1_c3", "exg1_c4", "exg1_c5", "exg2_c3", "exg2_c4", "exg2_c5", "convlstm_h0", "convlstm_h1", "convlstm_h2"]


def _check(ro, out, ref, head, label):
    rep = _report(ref, head, STAGES)
    print(f"\n[{label}]\n{rep}")
    d = (out["pred"] - ro["pred"]).abs().max().item()
    du = (out["up"] - ro["up"]).abs().max().item()
    agree = ((out["up"] > 0) == (ro["up"] > 0)).float().mean().item()
    near = (ro["up"].abs() < LOGIT_TOL).float().mean().item()
    print(f"[{label}] pred max-abs {d:.3e}  up max-abs {du:.3e}  mask agreement {agree:.5f}  (|logit|<1e-2 on {near:.4f} of pixels)")
    assert torch.isfinite(out["pred"]).all()
    assert (out["words_parse"] - ro["words_parse"]).abs().max() < 1e-4
    assert d <= LOGIT_TOL and du <= LOGIT_TOL, rep
    assert (out["sigm"] - ro["sigm"]).abs().max() <= LOGIT_TOL
    return agree


@pytest.mark.parametrize("batch,sharp,seq_len", [(1, 1.0, None), (3, 40.0, [20, 7, 1]), (2, 10.0, [0, 13])])
def test_tiny_head_matches_oracle(batch, sharp, seq_len):
    cfg, ref, ro, head, out, _ = _run(TINY, batch, sharp=sharp, bias_std=0.05, ln_jitter=0.1, seq_len=seq_len)
    _check(ro, out, ref, head, f"tiny B={batch} sharp={sharp}")


def test_cfg1_full_size_b1():
    """BASELINE config 1: batch 1, 320x320 (40x40 maps, N=1600), 20-token expression, reference initialisers."""
    cfg, ref, ro, head, out, _ = _run({}, 1)
    agree = _check(ro, out, ref, head, "cfg1 B=1")
    assert agree >= MASK_AGREE


def test_losses_match_oracle():
    """Forward value of the 4-term sigmoid-CE objective + L2 regulariser (CMPC_model.py:439-447, util/loss.py)."""
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    kw = TINY
    cfg = HeadConfig(batch_size=2, **kw)
    params = init_params(cfg, 3, sharp=10.0, bias_std=0.05, ln_jitter=0.1)
    inp = make_inputs(cfg, 2, seed=99, seq_len=[20, 4])
    oh = OracleHead(params, cfg)
    ro = oh.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
    rl = oh.losses(ro, inp["target_fine"])
    dev = torch.device("cuda:0")
    model = LSTM_model(batch_size=2, num_steps=kw["num_steps"], vf_h=kw["vf_h"], vf_w=kw["vf_w"], H=kw["H"], W=kw["W"],
                       vf_dim=kw["vf_dim"], v_emb_dim=kw["v_emb_dim"], rnn_size=kw["rnn_size"], mlp_dim=kw["mlp_dim"],
                       params=params, device=dev, head_kwargs=dict(c4_dim=kw["c4_dim"], c3_dim=kw["c3_dim"], parse_hidden=kw["parse_hidden"]))
    model.forward(inp["c3"].to(dev), inp["c4"].to(dev), inp["c5"].to(dev), inp["lstm_outputs"].to(dev),
                  target_fine=inp["target_fine"].to(dev), aux=True)
    got = model.losses()
    for k in ("cls_loss", "cls_loss_c5", "cls_loss_c4", "cls_loss_c3", "cls_loss_all", "reg_loss", "cost"):
        a, b = float(got[k]), float(rl[k])
        assert abs(a - b) <= 2e-3 * abs(b) + 1e-6, (k, a, b)      # sums over 4096 pixels of logits that agree to ~1e-3


def test_cfg4_high_resolution_512_b1():
    """BASELINE config 4 geometry: 512x512 input, 64x64 maps, N = 4096 graph nodes (32 key tiles, peepholes 64x64x500)."""
    cfg, ref, ro, head, out, _ = _run(dict(vf_h=64, vf_w=64, H=512, W=512), 1, sharp=40.0, seq_len=[11])
    agree = _check(ro, out, ref, head, "cfg4 512^2 B=1")
    assert agree >= MASK_AGREE


def test_full_size_sharp_ragged_b2():
    """Sharp-affinity regime, ragged sentence lengths, non-trivial biases / LN params, aux heads."""
    cfg, ref, ro, head, out, inp = _run({}, 2, sharp=60.0, bias_std=0.02, ln_jitter=0.05, seq_len=[20, 5], aux=True)
    agree = _check(ro, out, ref, head, "full B=2 sharp")
    assert agree >= MASK_AGREE
    for lvl in ("c5", "c4", "c3"):
        assert (out[f"up_{lvl}"] - ro[f"up_{lvl}"]).abs().max() <= LOGIT_TOL
    # integer I/U must be bit-exact given the same logits
    I, U = mask_iu(out["up"], inp["target_fine"])
    gi, gu = head.mask_iu(out["up"].cuda(), inp["target_fine"].cuda())
    assert torch.equal(I, gi.cpu()) and torch.equal(U, gu.cpu())


@pytest.mark.parametrize("overlap", [False, True], ids=["one-stream", "lang-side-streams"])
def test_cuda_graph_replay_matches_eager(overlap):
    """The graphed pass (one graph launch instead of ~100 kernel launches) must reproduce the eager pass on new data
    written into the same input buffers -- also when the language side forks onto side streams (head.overlap_lang, the default for
    batch >= 8), which the capture has to record as parallel branches that join before the first level."""
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from cmpc_refseg_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    kw = dict(batch_size=1, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, v_emb_dim=64, rnn_size=64, mlp_dim=32, device=dev,
              head_kwargs=dict(c4_dim=64, c3_dim=32, parse_hidden=40))
    eager, graphed = LSTM_model(**kw), LSTM_model(cuda_graph=True, **kw)
    eager._head.overlap_lang = graphed._head.overlap_lang = overlap
    gen = dict(vf_h=8, vf_w=8, H=64, W=64, c3_dim=32, c4_dim=64, vf_dim=128, rnn_size=64)
    bufs = {k: v.to(dev) for k, v in make_inputs(1, seed=1, **gen).items() if k in ("c3", "c4", "c5", "lstm_outputs")}
    for seed in (1, 2, 3):
        new = make_inputs(1, seed=seed, seq_len=[20 - 5 * seed], **gen)
        for k in bufs:
            bufs[k].copy_(new[k])
        a = eager.forward(bufs["c3"], bufs["c4"], bufs["c5"], bufs["lstm_outputs"])["pred"].clone()
        b = graphed.forward(bufs["c3"], bufs["c4"], bufs["c5"], bufs["lstm_outputs"])["pred"].clone()
        torch.cuda.synchronize()
        assert (a - b).abs().max() < 3e-3 and torch.isfinite(b).all()


def test_programmatic_dependent_launch_changes_nothing(lib):
    """cmpc_set_pdl(1): the GEMM kernels are launched with programmatic stream serialization and wait (griddepcontrol.wait) for the
    kernel in front of them after their own prologue -- same results, eager and replayed from a CUDA graph (PDL edges in the capture)."""
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from cmpc_refseg_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    inp = {k: v.to(dev) for k, v in make_inputs(2, seed=5, seq_len=[20, 7]).items() if k in ("c3", "c4", "c5", "lstm_outputs")}
    outs = {}
    try:
        for pdl in (0, 1):
            lib.cmpc_set_pdl(pdl)
            for graph in (False, True):
                model = LSTM_model(batch_size=2, device=dev, cuda_graph=graph, seed=0)
                for _ in range(3):
                    out = model.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
                torch.cuda.synchronize()
                outs[(pdl, graph)] = out["pred"].clone()
                del model
    finally:
        lib.cmpc_set_pdl(0)
    ref = outs[(0, False)]
    for k, v in outs.items():
        assert torch.isfinite(v).all() and float((v - ref).abs().max()) < 3e-3, k      # run-to-run (atomics) noise level


@pytest.mark.parametrize("B,geom", [(3, "tiny"), (2, "cfg1")])
def test_deferred_mutan_normalisation_matches_the_explicit_pass(B, geom):
    """head.lazy_norm (default in inference): the l2_normalize of the MUTAN map (CMPC_model.py:324) is applied by its four consumers
    (affinity GEMM row scale + bias row as per-sample bias, V carrying 1 / |x_j|, the residual of graph_conv, the fusion GEMM's
    accumulator scale with spa_graph pre-multiplied by |x|) instead of by a pass of its own -- same logits as the explicit pass."""
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from cmpc_refseg_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    if geom == "tiny":
        kw = dict(vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, v_emb_dim=64, rnn_size=64, mlp_dim=32, head_kwargs=dict(c4_dim=64, c3_dim=32, parse_hidden=40))
        gen = dict(vf_h=8, vf_w=8, H=64, W=64, c3_dim=32, c4_dim=64, vf_dim=128, rnn_size=64)
    else:
        kw, gen = {}, {}
    model = LSTM_model(batch_size=B, device=dev, seed=3, **kw)
    inp = {k: v.to(dev) for k, v in make_inputs(B, seed=11, seq_len=[20, 3, 9][:B], **gen).items() if k in ("c3", "c4", "c5", "lstm_outputs")}
    outs = {}
    for lazy in (False, True):
        model._head.lazy_norm = lazy
        out = model.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
        torch.cuda.synchronize()
        outs[lazy] = {k: out[k].clone() for k in ("pred", "gw_w", "gw_v")}
    assert (outs[True]["pred"] - outs[False]["pred"]).abs().max() < 3e-3            # two fp16 roundings fewer on the lazy side
    assert (outs[True]["gw_w"] - outs[False]["gw_w"]).abs().max() < 2e-4
    assert (outs[True]["gw_v"] - outs[False]["gw_v"]).abs().max() < 2e-5
    assert torch.isfinite(outs[True]["pred"]).all()
