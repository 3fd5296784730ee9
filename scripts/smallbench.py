"""key_fold (u = W_key q per module, head._st_nec_derived): plain vs split-K accumulate mode of cmpc_small_linear_f32."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
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
