"""Not a test (no test_ prefix): prints the stage-by-stage gradient errors of one level against autograd on the oracle.
Run by hand on a GPU box: python tests/debug_level_bwd.py  (lives under tests/ because only tests may import the oracle)."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import torch
from oracle.cmpc_head_ref import (HeadConfig, OracleHead, generate_spatial_batch, init_params, l2_normalize, make_inputs, layer_norm_tf)
from cmpc_refseg_b200.CMPC_model import LSTM_model
from cmpc_refseg_b200.backward import HeadBackward, Saved
from cmpc_refseg_b200.weights import LEVELS
cfg_kw = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, c4_dim=64, c3_dim=32, v_emb_dim=64, rnn_size=64, mlp_dim=32, parse_hidden=40)
level = "c5"
B = 2
cfg = HeadConfig(batch_size=B, **cfg_kw)
params = init_params(cfg, 0, sharp=6.0, bias_std=0.05, ln_jitter=0.2)
g = torch.Generator().manual_seed(13)
C_, Mm, R, T = cfg.v_emb_dim, cfg.mlp_dim, cfg.rnn_size, cfg.num_steps
inp = make_inputs(cfg, B, seed=3, seq_len=[min(T, 9), 4])
x = l2_normalize(torch.randn(B, cfg.vf_h, cfg.vf_w, C_, generator=g), 3)
gup = torch.randn(B, cfg.vf_h, cfg.vf_w, Mm, generator=g) * 0.1
spatial = torch.from_numpy(generate_spatial_batch(B, cfg.vf_h, cfg.vf_w)).float()
P = params
ref = OracleHead(P, cfg)
wf, mask = ref.words(inp["lstm_outputs"])
parse = ref.build_lang_parser(wf, mask)
vl = ref.valid_lang(parse, wf)
N = cfg.n_nodes
xg = x.clone().requires_grad_(True)
wt_in = ref._conv(f"words_trans_{level}", wf)
rel = parse[:, :, :, 2]
xt = ref._conv(f"spa_graph_trans2_{level}", xg).reshape(B, N, C_)
raw = (xt @ wt_in.reshape(B, T, R).transpose(1, 2)) / (C_ ** 0.5)
affi = rel * raw; affi.retain_grad()
m = mask.reshape(B, 1, T)
gw_w = torch.softmax(m * affi + (1 - m) * torch.finfo(torch.float32).min, dim=2); gw_w.retain_grad()
gw_v = m * torch.softmax(affi, dim=1); gw_v.retain_grad()
adj = gw_w @ gw_v.transpose(1, 2)
X = xg.reshape(B, N, C_)
Y = (adj @ X).reshape(B, 1, N, C_); Y.retain_grad()
Yn = layer_norm_tf(Y, P[f"gconv_feat_ln_spa_graph_{level}/gamma"], P[f"gconv_feat_ln_spa_graph_{level}/beta"])
Z = torch.relu(xg.reshape(B, 1, N, C_) + Yn); Z.retain_grad()
U = ref._conv(f"gconv_update_spa_graph_{level}", Z); U.retain_grad()
Un = layer_norm_tf(U, P[f"gconv_update_ln_spa_graph_{level}/gamma"], P[f"gconv_update_ln_spa_graph_{level}/beta"])
G0 = torch.relu(Un)
spa = l2_normalize(G0.reshape(B, cfg.vf_h, cfg.vf_w, C_), 3); spa.retain_grad()
fus = torch.relu(ref._conv(f"fusion_{level}", torch.cat([xg, spa, vl.expand(-1, cfg.vf_h, cfg.vf_w, -1), spatial], 3)))
loss = (fus * gup).sum()
loss.backward()
dev = torch.device("cuda:0")
hk = {k: cfg_kw[k] for k in ("c4_dim", "c3_dim", "parse_hidden")}
mk = {k: v for k, v in cfg_kw.items() if k not in hk}
model = LSTM_model(batch_size=B, params=params, device=dev, head_kwargs=hk, **mk)
head, i = model._head, LEVELS.index(level)
head.saved = Saved(dev)
head._begin()
wfd, _ = model.lstm(inp["lstm_outputs"].to(dev))
model._load_words(wfd); head._st_parse(); head._st_words_derived(); head._st_valid_derived()
model._load_map(x.to(dev).contiguous(), head._lb("x16", i), -1)
head._st_affinity(i, True); head._st_graph_conv(i); head._st_fusion(i)
bw = HeadBackward(head)
dfus = torch.zeros(B * N, head.d.GW, device=dev); dfus[:, :Mm] = gup.reshape(B * N, Mm).to(dev)
dxg, dres, dagg, daff = bw.bwd_level(i, dfus, head.d.GW)
torch.cuda.synchronize()
LDC = head.d.LDC
def cmp(name, a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu().reshape(a.shape)
    print(f"{name:28s} rel-L2 {float((a-b).norm()/b.norm().clamp_min(1e-30)):.3e}  absmax ref {float(b.abs().max()):.3e} dev {float(a.abs().max()):.3e}")
cmp("d spa (dxg[:, LDC:])", dxg[:, LDC:LDC + C_], spa.grad)
cmp("d U (du16)", bw.du16[:, :C_].float(), U.grad)
cmp("d Z*mask (dres)", dres[:, :C_], Z.grad * (Z > 0))
cmp("d Y (dyg16)", bw.dyg16[:, :C_].float(), Y.grad)
cmp("d gw_w (dwm)", bw.dwm[:, :T], gw_w.grad)
cmp("d gw_v (dvm)", bw.dvm[:, :T], gw_v.grad)
cmp("d affi*r (draw16)", bw.draw16[:, :T].float(), affi.grad * rel.reshape(B, 1, T))
xgrad_agg = None
cmp("dagg", dagg[:, :C_], (adj.transpose(1, 2) @ Y.grad.reshape(B, N, C_)))
cmp("d x total", dxg[:, :C_] + dres[:, :C_] + dagg[:, :C_] + daff[:, :C_], xg.grad)
cmp("dz raw (gemm)", bw.dzl[:, :C_], Z.grad)
wT = bw.gupd_wT[level].float()
cmp("dz via torch on device ops", (bw.du16.float() @ wT.t())[:, :C_], Z.grad)
cmp("dz via torch, other orient", (bw.du16.float() @ wT)[:, :C_], Z.grad)
sv = head.saved.t
cmp("z16 vs Z", sv[f"z16_{level}"][:, :C_].float(), Z)
cmp("y16 vs Y", sv[f"y16_{level}"][:, :C_].float(), Y)
cmp("u16 vs U", sv[f"u16_{level}"][:, :C_].float(), U)
dA = affi.grad
DW2 = P[f"spa_graph_trans2_{level}/DW"][0, 0]
daff_ref = ((dA * rel.reshape(B, 1, T) / C_ ** 0.5) @ wt_in.reshape(B, T, R)) @ DW2.t()
cmp("daff", daff[:, :C_], daff_ref)
cmp("draw16 * sqrt(C)", bw.draw16[:, :T].float() * C_ ** 0.5, dA * rel.reshape(B, 1, T))
cmp("dxg[:, :C] (fusion conv)", dxg[:, :C_], xg.grad - dres[:, :C_].cpu().reshape(xg.shape) - dagg[:, :C_].cpu().reshape(xg.shape) - daff_ref.reshape(xg.shape))
gt16 = head.buf["gt16"][i].float().view(B, T, -1)
Gt_ref = wt_in.reshape(B, T, R) @ DW2.t()
cmp("gt16 vs Gt", gt16[:, :, :C_], Gt_ref)
cmp("gtT16", bw.gtT16[:, :C_, :T].float(), Gt_ref.transpose(1, 2))
