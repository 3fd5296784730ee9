import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
L.LIB_PATH = L._PKG / "libcmpc_b200_timing.so"
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
B, N, Cc = 32, 1600, 1000; M = B * N
w16 = torch.rand(M, 32, device=dev).half(); v16 = torch.rand(M, 32, device=dev).half()
x = (torch.randn(M, 1024, device=dev) * 0.03).half(); y = torch.empty_like(x); stt = torch.zeros(B, 2, device=dev, dtype=torch.float64)
ncta = 4 * 14 * B
tl = torch.zeros(ncta, 64, device=dev, dtype=torch.int64)
for _ in range(3):
    L.check(lib.cmpc_graph_reason_f16(w16.data_ptr(), v16.data_ptr(), x.data_ptr(), 1024, B, N, Cc, 2048.0, y.data_ptr(), 1024, stt.data_ptr(), tl.data_ptr(), st))
torch.cuda.synchronize()
t = tl.cpu().double()
t0 = t[:, 0:1]
rel = (t - t0)
def col(i): return rel[:, i]
print("CTAs", ncta)
print(f"w_full wait done      : {col(1).median():9.0f} cyc")
waits = [(col(3 + 2 * j) - col(2 + 2 * j)).median().item() for j in range(13)]
starts = [col(2 + 2 * j).median().item() for j in range(13)]
print("x/p_full wait per j (median cycles):", [int(w) for w in waits])
print("iteration start (median cycles)  :", [int(s) for s in starts])
print(f"all MMAs issued       : {col(40).median():9.0f}")
c = lambda i: (t[:, i] - t[:, 0]).median().item()
print(f"epilogue: wait start {c(41):.0f}  o_full {c(42):.0f}  boxes staged {c(49):.0f}  tma read done {c(51):.0f}  syncthreads {c(52):.0f}  dealloc {c(43):.0f}")
