"""Generates the reference-executed golden fixtures ``tests/golden/ref_*.npz``.

Every value written here comes out of the reference's OWN ``CMPC_model.py`` / ``util/cell.py`` / ``util/loss.py`` /
``util/processing_tools.py``, imported unmodified from the read-only checkout and executed through the eager TensorFlow
stand-in (``oracle/tfshim``, driven by ``oracle/ref_runner.py``) -- NOT out of ``oracle/cmpc_head_ref.py``.  The oracle is only
used for its seeded input / parameter generators (``make_inputs`` / ``init_params``: plain ``torch.Generator`` draws), so that
the fixtures need not carry the 13 MB feature maps; each fixture records float64 checksums of the inputs and parameters it was
made from, and the tests refuse to compare if a regenerated input does not reproduce them.

Main values are the float64 execution of the reference graph cast to float32 (free of summation-order noise, which in a real
TF run depends on Eigen's blocking anyway); ``pred_f32run`` is the same graph executed in float32 (what TF computes in), kept
to show the size of that noise.

    python tests/golden/make_ref_golden.py            # needs /root/reference (or CMPC_REFERENCE_ROOT)
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.cmpc_head_ref import HeadConfig, init_params, make_inputs  # noqa: E402
from oracle.ref_runner import reference_available, run_reference  # noqa: E402

OUT = Path(__file__).resolve().parent
F64 = torch.float64

# the reference hard-codes the c4 / c3 tap widths (CMPC_model.py:110,112) and the parser's hidden width (:349)
REF_FIXED = dict(c4_dim=1024, c3_dim=512, parse_hidden=500)
TINY = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, v_emb_dim=64, rnn_size=64, mlp_dim=32)
FULL = dict(num_steps=20, vf_h=40, vf_w=40, H=320, W=320, vf_dim=2048, v_emb_dim=1000, rnn_size=1000, mlp_dim=500)
HIRES = dict(num_steps=20, vf_h=64, vf_w=64, H=512, W=512, vf_dim=2048, v_emb_dim=1000, rnn_size=1000, mlp_dim=500)
FWD_KEYS = ("pred", "up", "sigm", "up_c3", "up_c4", "up_c5", "words_parse", "gw_w", "gw_v", "seq_mask")

# name -> (model kwargs, batch, param seed, param kwargs, input seed, seq_len, keys to store)
FORWARD_CASES = {
    "ref_tiny_b1": (TINY, 1, 7, dict(sharp=40.0, bias_std=0.05, ln_jitter=0.1), 4321, [7], FWD_KEYS),
    "ref_tiny_b3": (TINY, 3, 7, dict(sharp=40.0, bias_std=0.05, ln_jitter=0.1), 4321, [20, 9, 2], FWD_KEYS),
    "ref_cfg1_random": (FULL, 1, 0, dict(), 1234, None, FWD_KEYS),
    "ref_cfg1_sharp": (FULL, 1, 0, dict(sharp=60.0, bias_std=0.02, ln_jitter=0.1), 1234, [6], FWD_KEYS),
    # BASELINE configs[1] at its full batch, the literal (batch-coupled, CMPC_model.py:241) graph; only the small outputs are kept
    # BASELINE configs[3]: 512 x 512 input, 64 x 64 maps, a 4096-node graph (one sample; the dense 4096 x 4096 adjacency is materialised here)
    "ref_hires_b1": (HIRES, 1, 0, dict(sharp=60.0, bias_std=0.02, ln_jitter=0.1), 2468, [9], ("pred", "words_parse", "seq_mask")),
    "ref_cfg2_b32": (FULL, 32, 0, dict(sharp=60.0, bias_std=0.02, ln_jitter=0.1), 1234, "unc", ("pred", "words_parse", "seq_mask")),
}


def checksums(inp, params):
    cs = {"cs_" + k: float(inp[k].double().sum()) for k in ("c3", "c4", "c5", "lstm_outputs")}
    cs["cs_params"] = float(sum(v.double().abs().sum() for v in params.values()))
    return cs


def gen_inputs(model_kw, B, pseed, pkw, iseed, seq_len):
    cfg = HeadConfig(batch_size=B, **model_kw, **REF_FIXED)
    params = init_params(cfg, seed=pseed, dtype=F64, **pkw)
    inp = make_inputs(cfg, B, seed=iseed, seq_len=seq_len)
    return cfg, params, inp


def forward_case(name, spec):
    model_kw, B, pseed, pkw, iseed, seq_len, keys = spec
    cfg, params, inp = gen_inputs(model_kw, B, pseed, pkw, iseed, seq_len)
    t0 = time.time()
    ref64 = run_reference(dict(batch_size=B, mode="eval", **model_kw), params, inp["c3"], inp["c4"], inp["c5"],
                          lstm_outputs=inp["lstm_outputs"], float64=True)
    p32 = {k: v.float() for k, v in params.items()}
    ref32 = run_reference(dict(batch_size=B, mode="eval", **model_kw), p32, inp["c3"], inp["c4"], inp["c5"],
                          lstm_outputs=inp["lstm_outputs"], float64=False)
    blob = {k: ref64[k].float().numpy() for k in keys}
    blob["pred_f32run"] = ref32["pred"].numpy()
    blob["seq_len"] = inp["seq_len"].numpy()
    blob.update(checksums(inp, params))
    np.savez_compressed(OUT / (name + ".npz"), **blob)
    noise = float((ref32["pred"].double() - ref64["pred"]).abs().max())
    print(f"{name}: B={B} N={cfg.n_nodes} {time.time() - t0:.1f}s  fp32-run vs fp64-run pred max-abs {noise:.2e}  "
          f"{(OUT / (name + '.npz')).stat().st_size / 1e6:.2f} MB")


TRAIN_CASES = {   # name -> parameter kwargs: a near-argmax affinity regime and a mild one
    "ref_tiny_train": dict(sharp=40.0, bias_std=0.05, ln_jitter=0.1),
    "ref_tiny_train_mild": dict(sharp=6.0, bias_std=0.05, ln_jitter=0.2),
}


def train_case(name="ref_tiny_train"):
    """train_op (CMPC_model.py:426-492) on the tiny head: losses, every gradient as compute_gradients returned it, the
    variables after the one Adam step apply_gradients performed (bias gradients doubled, L2 term inside the cost)."""
    B = 3
    pkw = TRAIN_CASES[name]
    cfg, params, inp = gen_inputs(TINY, B, 7, pkw, 4321, [20, 9, 2])
    ref = run_reference(dict(batch_size=B, mode="train", **TINY), params, inp["c3"], inp["c4"], inp["c5"],
                        lstm_outputs=inp["lstm_outputs"], target_fine=inp["target_fine"], float64=True)
    blob = {k: np.float64(ref[k]) for k in ("cls_loss", "cls_loss_c3", "cls_loss_c4", "cls_loss_c5", "cls_loss_all", "reg_loss",
                                            "cost", "learning_rate", "mIoU")}
    blob["target"] = ref["target"].float().numpy()
    for k, g in ref["raw_grads"].items():
        if g is not None:
            blob["grad/" + k] = g.float().numpy()
            blob["step/" + k] = (ref["variables"][k] - params[k]).float().numpy()      # the applied Adam update
    blob["trainable"] = np.array(ref["trainable"])
    blob.update(checksums(inp, params))
    blob["cs_target"] = float(inp["target_fine"].double().sum())
    np.savez_compressed(OUT / (name + ".npz"), **blob)
    print(name + ":", len(ref["raw_grads"]), "variables,", f"{(OUT / (name + '.npz')).stat().st_size / 1e6:.2f} MB")


def words_case():
    """the whole lstm() front (CMPC_model.py:144-164: embedding lookup -> LSTMCell -> dynamic_rnn(sequence_length)) feeding the
    head, in training mode: pins the word encoder, its backward and the gradient at the head's lstm_outputs boundary."""
    B, V, E = 3, 50, 12
    pkw = dict(sharp=40.0, bias_std=0.05, ln_jitter=0.1)
    cfg, params, inp = gen_inputs(TINY, B, 7, pkw, 4321, [20, 9, 2])
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(V, E, generator=g, dtype=F64)
    words = torch.randint(0, V, (B, cfg.num_steps), generator=g)
    kernel = (torch.rand(E + cfg.rnn_size, 4 * cfg.rnn_size, generator=g, dtype=F64) * 2 - 1) * 0.3
    bias = torch.randn(4 * cfg.rnn_size, generator=g, dtype=F64) * 0.1
    params = dict(params)
    params["rnn/lstm_cell/kernel"], params["rnn/lstm_cell/bias"] = kernel, bias
    sl = torch.tensor([20, 9, 2])
    ref = run_reference(dict(batch_size=B, mode="train", glove_dim=E, **TINY), params, inp["c3"], inp["c4"], inp["c5"],
                        words=words, seq_len=sl, embedding=emb, target_fine=inp["target_fine"], float64=True)
    blob = dict(words=words.numpy(), seq_len=sl.numpy(), embedding=emb.numpy(), kernel=kernel.numpy(), bias=bias.numpy(),
                pred=ref["pred"].float().numpy(), seq_mask=ref["seq_mask"].float().numpy(), cost=np.float64(ref["cost"]))
    for k in ("Variable", "rnn/lstm_cell/kernel", "rnn/lstm_cell/bias", "words_parse_1/DW", "c5_lateral/DW"):
        blob["grad/" + k] = ref["raw_grads"][k].numpy()
    blob.update(checksums(inp, {k: v for k, v in params.items() if not k.startswith("rnn/lstm_cell")}))
    np.savez_compressed(OUT / "ref_tiny_words.npz", **blob)
    print("ref_tiny_words:", f"{(OUT / 'ref_tiny_words.npz').stat().st_size / 1e6:.2f} MB")


def main():
    if not reference_available():
        raise SystemExit("the reference checkout is not available (set CMPC_REFERENCE_ROOT)")
    torch.manual_seed(0)
    only = set(sys.argv[1:])
    for name, spec in FORWARD_CASES.items():
        if not only or name in only:
            forward_case(name, spec)
    for name in TRAIN_CASES:
        if not only or name in only:
            train_case(name)
    if not only or "ref_tiny_words" in only:
        words_case()


if __name__ == "__main__":
    main()
