"""GPU parity against the REFERENCE-EXECUTED fixtures (tests/golden/ref_*.npz: the reference's own CMPC_model.py run
unmodified through oracle/tfshim, float64) -- the oracle is not involved in any comparison here; it only supplies the seeded
input / parameter generators the fixtures were made from (checksummed).

north_star tolerances: logits max-abs <= 1e-2, thresholded-mask agreement >= 99.9 %.
"""
import numpy as np
import pytest
import torch

import refgold

pytestmark = pytest.mark.gpu
LOGIT_TOL, MASK_AGREE = 1e-2, 0.999


def _model(kw, B, params, **extra):
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    hk = {k: kw[k] for k in ("c4_dim", "c3_dim", "parse_hidden")}
    mk = {k: v for k, v in kw.items() if k not in hk}
    return LSTM_model(batch_size=B, params={k: v.float() for k, v in params.items()}, device=torch.device("cuda:0"),
                      head_kwargs=hk, **mk, **extra)


def _feed(inp, dev="cuda:0"):
    return [inp[k].to(dev) for k in ("c3", "c4", "c5", "lstm_outputs")]


@pytest.mark.parametrize("name", ["ref_tiny_b1", "ref_tiny_b3", "ref_cfg1_random", "ref_cfg1_sharp"])
def test_head_matches_reference_execution(name):
    kw, B, cfg, params, inp, fix = refgold.forward_case(name, torch.float32)
    model = _model(kw, B, params, gv_norm="batch" if B > 1 else "sample")
    out = model.forward(*_feed(inp), aux=True)
    torch.cuda.synchronize()
    o = {k: out[k].float().cpu() for k in ("pred", "up", "sigm", "up_c3", "up_c4", "up_c5", "words_parse", "gw_w", "gw_v", "seq_mask")}
    d = {k: float((o[k].reshape(fix[k].shape) - fix[k]).abs().max()) for k in o}
    agree = float(((o["up"] > 0) == (fix["up"] > 0)).float().mean())
    print(f"\n[{name}] vs reference execution: " + "  ".join(f"{k} {v:.2e}" for k, v in d.items()) + f"  mask agreement {agree:.5f}")
    assert torch.equal(o["seq_mask"].reshape(fix["seq_mask"].shape), fix["seq_mask"])
    assert d["pred"] <= LOGIT_TOL and d["up"] <= LOGIT_TOL and d["up_c3"] <= LOGIT_TOL and d["up_c4"] <= LOGIT_TOL and d["up_c5"] <= LOGIT_TOL
    assert d["sigm"] <= LOGIT_TOL / 4 and d["words_parse"] <= 1e-4 and d["gw_w"] <= 5e-3 and d["gw_v"] <= 5e-3
    assert agree >= MASK_AGREE
    # legacy bilinear x8: the stride-8 samples of `up` are the logits themselves (exact property of the reference's resize)
    s = kw["H"] // kw["vf_h"]
    assert torch.equal(o["up"][:, ::s, ::s], o["pred"])


@pytest.mark.parametrize("graphed", [False, True], ids=["eager", "cuda_graph"])
def test_head_matches_reference_at_benchmark_batch(graphed):
    """BASELINE configs[1]: batch 32, N = 1600, UNC-shaped sentence lengths, sharp affinities -- the literal (batch-coupled) graph the
    reference builds at batch 32, through the default forward (side-stream language chains on at this batch) and through CUDA-graph replay."""
    kw, B, cfg, params, inp, fix = refgold.forward_case("ref_cfg2_b32", torch.float32)
    model = _model(kw, B, params, gv_norm="batch", cuda_graph=graphed)
    assert model._head.overlap_lang
    feed = _feed(inp)
    for _ in range(2 if graphed else 1):
        out = model.forward(*feed)
    torch.cuda.synchronize()
    pred = out["pred"].cpu()
    d = (pred - fix["pred"]).abs().amax(dim=(1, 2, 3))
    agree = float(((pred > 0) == (fix["pred"] > 0)).float().mean())
    print(f"\n[cfg2 B=32 {'graph' if graphed else 'eager'}] logits max-abs over the batch {float(d.max()):.3e} (median sample {float(d.median()):.3e}), "
          f"sign agreement {agree:.5f}")
    assert float(d.max()) <= LOGIT_TOL and agree >= MASK_AGREE
    assert float((out["words_parse"].cpu() - fix["words_parse"]).abs().max()) <= 1e-4
    assert torch.equal(out["seq_mask"].cpu().reshape(fix["seq_mask"].shape), fix["seq_mask"])


# measured on B200 (fp16 operands vs the float64 reference execution): mild regime worst/median, near-argmax regime worst/median;
# the bounds are 1.5x the measured figures ("tighten to the measured level")
GRAD_BOUNDS = {"ref_tiny_train_mild": (0.03, 0.0065), "ref_tiny_train": (0.105, 0.005)}     # measured: 0.020 / 0.0041 and 0.070 / 0.0032


def test_head_matches_reference_at_4096_nodes_in_a_batch_of_16():
    """BASELINE configs[3]: 512 x 512 input (64 x 64 maps, a 4096-node graph), batch 16 -- sample 0 of the batch is the input of the
    reference-executed fixture ref_hires_b1 (per-sample gv_lang norm = the reference run one sample at a time), the other 15 are
    different samples; eager and through CUDA-graph replay."""
    from oracle.cmpc_head_ref import make_inputs
    kw, _, cfg, params, inp, fix = refgold.forward_case("ref_hires_b1", torch.float32)
    B = 16
    rest = make_inputs(cfg, B - 1, seed=97, seq_len="unc")
    feed = [torch.cat([inp[k], rest[k]]).to("cuda:0") for k in ("c3", "c4", "c5", "lstm_outputs")]
    for graphed in (False, True):
        model = _model(kw, B, params, gv_norm="sample", cuda_graph=graphed)
        for _ in range(2 if graphed else 1):
            out = model.forward(*feed)
        torch.cuda.synchronize()
        pred = out["pred"][:1].cpu()
        d = float((pred - fix["pred"]).abs().max())
        agree = float(((pred > 0) == (fix["pred"] > 0)).float().mean())
        print(f"\n[hires N=4096, B=16, {'graph' if graphed else 'eager'}] sample 0 vs reference: logits max-abs {d:.3e}, sign agreement {agree:.5f}")
        assert d <= LOGIT_TOL and agree >= MASK_AGREE
        assert float((out["words_parse"][:1].cpu() - fix["words_parse"]).abs().max()) <= 1e-4
        assert torch.isfinite(out["pred"]).all()
        del model
        torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["ref_tiny_train_mild", "ref_tiny_train"])
def test_gradients_match_reference_train_op(name):
    """compute_gradients of the reference's train_op (CMPC_model.py:461), all 212 head variables, float64 reference execution vs the
    device backward (fp16 operands); then the applied Adam step (bias gradients x2, L2 inside the cost, :446-478)."""
    from cmpc_refseg_b200.backward import HeadBackward, Saved
    kw, B, cfg, params, inp, fix = refgold.train_case(torch.float32, name)
    dev = torch.device("cuda:0")
    model = _model(kw, B, params, mode="train")
    assert model.gv_norm == "batch"
    head = model._head
    head.saved = Saved(dev)
    out = head.forward(*_feed(inp), aux=True)
    bw = HeadBackward(head)
    target = inp["target_fine"].to(dev)
    bw.backward(out, target)
    torch.cuda.synchronize()
    gt = bw.grads_tf()
    wd = 0.0005
    rows = []
    for k in params:
        r = torch.from_numpy(fix["grad/" + k]).double()
        if k.endswith("/DW"):
            r = r - wd * params[k].double()              # the fixture differentiates cost = cls_loss_all + reg_loss; the device adds the L2 term in Adam
        a = gt[k].detach().double().cpu().reshape(r.shape)
        assert torch.isfinite(a).all(), k
        if float(r.norm()) < 1e-9:
            assert float(a.abs().max()) <= 1e-6, k
            continue
        rows.append((float((a - r).norm() / r.norm()), k))
    rows.sort(reverse=True)
    print(f"\n[{name}] parameter gradients vs the reference's compute_gradients (relative L2):")
    for l2, k in rows[:10]:
        print(f"   {l2:.3e}  {k}")
    med = rows[len(rows) // 2][0]
    exg = [r for r in rows if any(t in r[1] for t in ("gv_lang", "lang_feat", "lang_query", "spa_graph_key"))]
    print(f"   median {med:.3e} over {len(rows)} tensors; exchange-module tensors (batch-coupled norm) worst {exg[0][0]:.3e} {exg[0][1]}")
    worst_bound, med_bound = GRAD_BOUNDS[name]
    assert rows[0][0] < worst_bound and med < med_bound
    # losses
    L = model_losses(model, head, out, target)
    for k in ("cls_loss", "cls_loss_c3", "cls_loss_c4", "cls_loss_c5", "cls_loss_all"):
        assert abs(L[k] - float(fix[k])) <= 2e-3 * abs(float(fix[k])), (k, L[k], float(fix[k]))


def model_losses(model, head, out, target):
    r = {}
    for k, u in (("cls_loss", "up"), ("cls_loss_c3", "up_c3"), ("cls_loss_c4", "up_c4"), ("cls_loss_c5", "up_c5")):
        r[k] = float(head.ce_sums(out[u], target).mean())
    r["cls_loss_all"] = 0.7 * r["cls_loss"] + 0.1 * (r["cls_loss_c5"] + r["cls_loss_c4"] + r["cls_loss_c3"])
    return r


def test_train_step_matches_reference_adam_update():
    """one full optimizer step through LSTM_model(mode='train').train(...) vs the variables after the reference's apply_gradients"""
    kw, B, cfg, params, inp, fix = refgold.train_case(torch.float32)
    dev = torch.device("cuda:0")
    model = _model(kw, B, params, mode="train")
    tr = model.train_op()
    last = model.train(*_feed(inp), inp["target_fine"].to(dev))
    torch.cuda.synchronize()
    assert abs(last["cls_loss_all"] - float(fix["cls_loss_all"])) <= 2e-3 * float(fix["cls_loss_all"])
    assert abs(float(last["learning_rate"]) - float(fix["learning_rate"])) < 1e-12
    lr = float(fix["learning_rate"])
    agree = tot = 0
    worst = 0.0
    for k, p0 in params.items():
        step_ref = torch.from_numpy(fix["step/" + k]).double()
        step_dev = (tr.params[k].detach().double().cpu() - p0.double()).reshape(step_ref.shape)
        assert float(step_dev.abs().max()) <= lr * 1.001, k
        g = torch.from_numpy(fix["grad/" + k]).double().abs()
        big = g > 1e-2 * g.max()                        # Adam's first step is ~ -lr sign(g): compare where the sign is well defined
        if int(big.sum()) == 0:
            continue
        agree += int((torch.sign(step_dev[big]) == torch.sign(step_ref[big])).sum()); tot += int(big.sum())
        worst = max(worst, float((step_dev[big] - step_ref[big]).abs().max()) / lr)
    print(f"\nAdam step vs reference apply_gradients: {agree}/{tot} coordinates move the same way; worst |delta step| = {worst:.3f} lr")
    assert agree / tot > 0.999


def test_word_encoder_and_head_from_token_ids_match_reference():
    """lstm() front (CMPC_model.py:144-164) on the device, forward: token ids -> embedding -> LSTM -> head logits vs the reference"""
    from cmpc_refseg_b200.word_encoder import BIAS, EMB, KERNEL
    kw, B, cfg, params, inp, fix = refgold.words_case(torch.float32)
    p = dict(params)
    p[EMB], p[KERNEL], p[BIAS] = (torch.from_numpy(fix[k]).float() for k in ("embedding", "kernel", "bias"))
    model = _model(kw, B, p, gv_norm="batch", glove_dim=int(fix["embedding"].shape[1]), vocab_size=int(fix["embedding"].shape[0]))
    dev = torch.device("cuda:0")
    words, sl = torch.from_numpy(fix["words"]).to(dev), torch.from_numpy(fix["seq_len"]).to(dev)
    out = model.forward(inp["c3"].to(dev), inp["c4"].to(dev), inp["c5"].to(dev), None, sl, words=words)
    torch.cuda.synchronize()
    d = float((out["pred"].cpu() - torch.from_numpy(fix["pred"])).abs().max())
    print(f"\n[token ids -> logits] max-abs vs reference {d:.3e}")
    assert d <= LOGIT_TOL
    assert torch.equal(out["seq_mask"].cpu().reshape(fix["seq_mask"].shape), torch.from_numpy(fix["seq_mask"]))
