"""Debugging aid (not collected by pytest): per-variable gradient of the eager and the CUDA-graphed gradient phase of HeadTrainer on two
alternating batches, parameters held fixed.  python tests/debug_graph_grad.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from cmpc_refseg_b200.CMPC_model import LSTM_model                       # noqa: E402
from oracle.cmpc_head_ref import HeadConfig, init_params, make_inputs   # noqa: E402

TINY = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, c4_dim=64, c3_dim=32, v_emb_dim=64, rnn_size=64, mlp_dim=32, parse_hidden=40)
dev = torch.device("cuda:0")
kw, B = TINY, 2
cfg = HeadConfig(batch_size=B, **kw)
params = init_params(cfg, 0, sharp=6.0, bias_std=0.05, ln_jitter=0.2)
hk = {k: kw[k] for k in ("c4_dim", "c3_dim", "parse_hidden")}
mk = {k: v for k, v in kw.items() if k not in hk}
batches = []
for s in range(2):
    inp = make_inputs(cfg, B, seed=30 + s, seq_len=[20, 6])
    g = torch.Generator().manual_seed(s)
    batches.append([inp[k] for k in ("c3", "c4", "c5", "lstm_outputs")] + [(torch.rand(B, cfg.H, cfg.W, 1, generator=g) > 0.6).float()])

res = {}
for mode in ("eager", "graph"):
    model = LSTM_model(batch_size=B, params=params, device=dev, head_kwargs=hk, mode='train', start_lr=1e-3, lr_decay_step=10, **mk)
    tr = model.train_op()
    bufs = [t.to(dev).clone() for t in batches[0]]
    args = (bufs[0], bufs[1], bufs[2], bufs[3], bufs[4], None, None)
    if mode == "graph":
        key = tuple((x.data_ptr(), x.dtype) if torch.is_tensor(x) else None for x in args)
        ga, gb = tr._graphs_for(key, args)
    outs = []
    for step in range(4):
        for dst, src in zip(bufs, batches[step % 2]):
            dst.copy_(src)
        if mode == "graph":
            ga.replay()
            out = tr._graph_out
        else:
            out = tr._phase_grad(*args)
        torch.cuda.synchronize()
        if step == 1:
            def scan(prefix, obj):
                for k, v in (obj.items() if isinstance(obj, dict) else vars(obj).items()):
                    vs = v if isinstance(v, (list, tuple)) else [v]
                    for j, t in enumerate(vs):
                        if isinstance(t, (list, tuple)):
                            for jj, tt in enumerate(t):
                                if torch.is_tensor(tt) and tt.is_floating_point() and not torch.isfinite(tt).all():
                                    print(f"   [{mode}] non-finite: {prefix}.{k}[{j}][{jj}] {tuple(tt.shape)} count {int((~torch.isfinite(tt)).sum())}")
                        elif torch.is_tensor(t) and t.is_floating_point() and not torch.isfinite(t).all():
                            print(f"   [{mode}] non-finite: {prefix}.{k}[{j}] {tuple(t.shape)} count {int((~torch.isfinite(t)).sum())}")
            scan("bw", tr.bw); scan("bw.g", tr.bw.g); scan("saved", tr.h.saved.t); scan("buf", tr.h.buf)
        outs.append(({k: v.clone() for k, v in tr.grads.items()}, {k: v.double().clone() for k, v in out.items() if torch.is_tensor(v)},
                     {k: v.clone() for k, v in tr.ce.items()}))
    res[mode] = outs

for step in range(4):
    ge, oe, ce = res["eager"][step]
    gg, og, cg = res["graph"][step]
    print(f"--- step {step}")
    for k in oe:
        dn = float((oe[k] - og[k]).abs().max())
        if dn > 0:
            print(f"   out {k:12s} max-abs diff {dn:.3e}")
    for k in ce:
        print(f"   ce {k}: {ce[k].tolist()} vs {cg[k].tolist()}")
    bad = []
    for k in ge:
        den = float(ge[k].norm()) + 1e-30
        r = float((ge[k] - gg[k]).norm()) / den
        if not (r <= 1e-4):
            bad.append((r, k))
    for r, k in sorted(bad, reverse=True)[:25]:
        print(f"   grad {k:50s} rel diff {r:.3e}  eager {ge[k].flatten()[:4].tolist()} graph {gg[k].flatten()[:4].tolist()}")
    print(f"   {len(bad)} of {len(ge)} gradients differ > 1e-4")
