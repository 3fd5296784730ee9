"""Debug aid: which persistent buffer differs between the first and the second forward of a fresh head (lazy_norm on)."""
import sys; sys.path.insert(0,'/root/repo')
import torch
from cmpc_refseg_b200.head import CMPCHeadB200
from oracle.cmpc_head_ref import HeadConfig, init_params, make_inputs
dev=torch.device('cuda:0')
kw = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, c4_dim=64, c3_dim=32, v_emb_dim=64, rnn_size=64, mlp_dim=32, parse_hidden=40)
cfg = HeadConfig(batch_size=2, **kw)
params = init_params(cfg, 0, sharp=20.0, bias_std=0.05, ln_jitter=0.1)
inp = {k: v.to(dev) for k, v in make_inputs(cfg, 2, seq_len=[20, 6]).items() if torch.is_tensor(v)}
head = CMPCHeadB200(params, batch_size=2, device=dev, **kw)
head.lazy_norm = True
snaps=[]
for i in range(2):
    o = head.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"]); torch.cuda.synchronize()
    snaps.append({k: v.clone() for k, v in head.buf.items() if torch.is_tensor(v)})
for k in snaps[0]:
    a, b = snaps[0][k].float(), snaps[1][k].float()
    dmax = float((a - b).abs().max()) if a.numel() else 0.0
    nd = int(((a - b) != 0).sum()) if a.numel() else 0
    print(f"{k:12s} shape {tuple(a.shape)} max diff {dmax:.3e}  differing {nd} / {a.numel()}")
