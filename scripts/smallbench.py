"""key_fold (u = W_key q per module, head._st_nec_derived): plain vs split-K accumulate mode of cmpc_small_linear_f32."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
import os
if os.environ.get('CMPC_LIB'): L.LIB_PATH = L._PKG / os.environ['CMPC_LIB']
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
B, GW, Mm = 32, 512, 500
q = torch.randn(B, 6 * GW, device=dev); w = torch.randn(6, Mm, Mm, device=dev); u = torch.zeros(B, 6 * GW, device=dev)
def run(act):
    if act == 4: u.zero_()
    L.check(lib.cmpc_small_linear_f32(q.data_ptr(), 6 * GW, GW, w.data_ptr(), Mm, Mm * Mm, None, 0, u.data_ptr(), 6 * GW, GW, 6, B, Mm, Mm, act, st), "sl")
res = {}
for act in (0, 4):
    for _ in range(3): run(act)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): run(act)
    e1.record(); torch.cuda.synchronize()
    res[act] = u.clone()
    print(f"act={act}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
print("max diff", float((res[0] - res[4]).abs().max()))

# gv_gates (global_vec tail: gv = l2norm(g Wg + gvl), two sigmoid gates) at the bench shape, L2 flushed between launches
nmod = 3
pool = torch.randn(B, nmod, GW, device=dev); gvl = torch.randn(B, 6 * GW, device=dev)
wg = torch.randn(nmod, Mm, Mm, device=dev) * 0.05; wf1 = torch.randn(nmod, Mm, Mm, device=dev) * 0.05; wf2 = torch.randn(nmod, Mm, Mm, device=dev) * 0.05
bf1 = torch.randn(nmod, Mm, device=dev); bf2 = torch.randn(nmod, Mm, device=dev)
gv = torch.zeros(B, nmod, GW, device=dev); g1 = torch.zeros_like(gv); g2 = torch.zeros_like(gv)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def gvg():
    L.check(lib.cmpc_gv_gates(pool.data_ptr(), GW, gvl.data_ptr(), GW, 6 * GW, wg.data_ptr(), wf1.data_ptr(), bf1.data_ptr(), wf2.data_ptr(), bf2.data_ptr(),
                              Mm * Mm, Mm, B, nmod, Mm, gv.data_ptr(), g1.data_ptr(), g2.data_ptr(), GW, st), "gv_gates")
for cold in (False, True):
    ts = []
    for _ in range(12):
        if cold: flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); gvg(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"gv_gates {'cold' if cold else 'warm'}: {sorted(ts)[len(ts) // 2]:.1f} us   checksum {float(g1.sum() + g2.sum() + gv.sum()):.4f}")

# lang_parse (word-type softmax + valid_lang / nec_lang) at the bench shape
T, R, HID, HIDP = 20, 1000, 500, 512
hidden = torch.relu(torch.randn(B * T, HIDP, device=dev)); w2 = torch.randn(HID, 4, device=dev) * 0.1; b2 = torch.zeros(4, device=dev)
words = torch.randn(B * T, R, device=dev); mask = torch.ones(B, T, device=dev)
parse = torch.zeros(B, T, 4, device=dev); rgate = torch.zeros(B, 32, device=dev)
v32 = torch.zeros(B, R, device=dev); n32 = torch.zeros(B, R, device=dev)
v16 = torch.zeros(B, 1024, device=dev, dtype=torch.float16); n16 = torch.zeros_like(v16)
def lp():
    L.check(lib.cmpc_lang_parse(hidden.data_ptr(), HIDP, HID, w2.data_ptr(), b2.data_ptr(), words.data_ptr(), mask.data_ptr(), B, T, R, 1000,
                                parse.data_ptr(), rgate.data_ptr(), v32.data_ptr(), n32.data_ptr(), v16.data_ptr(), n16.data_ptr(), 1024, st), "lang_parse")
for cold in (False, True):
    ts = []
    for _ in range(12):
        if cold: flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); lp(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"lang_parse {'cold' if cold else 'warm'}: {sorted(ts)[len(ts) // 2]:.1f} us   checksum {float(parse.sum() + v32.sum() + n32.sum()):.4f}")
