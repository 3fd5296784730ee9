import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
L.LIB_PATH = L._PKG / "libcmpc_b200_timing.so"
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
B, N, Cc = 32, 1600, 1000; M = B * N
w16 = torch.rand(M, 32, device=dev).half(); v16 = torch.rand(M, 32, device=dev).half()
x = (torch.randn(M, 1024, device=dev) * 0.03).half(); y = torch.empty_like(x); stt = torch.zeros(B, 2, device=dev, dtype=torch.float64)
ncta = 148
tl = torch.zeros(ncta, 64, device=dev, dtype=torch.int64)
for _ in range(3):
    L.check(lib.cmpc_graph_reason_f16(w16.data_ptr(), v16.data_ptr(), x.data_ptr(), 1024, B, N, Cc, 2048.0, y.data_ptr(), 1024, stt.data_ptr(), tl.data_ptr(), st))
torch.cuda.synchronize()
t = tl.cpu().double()
# persistent kernel: every tick is relative to slot 0 = the MMA warp's o_full commit of the CTA's FIRST unit; slots 1-40 and
# 53-57 belong to the SECOND unit (steady state across a unit boundary), 41-52 to the first unit's epilogue
c = lambda i: (t[:, i] - t[:, 0]).median().item()
print("CTAs", ncta)
print(f"unit 1: w_full wait done {c(1):7.0f}")
print("unit 1: MMA2(j) wait start :", [int(c(2 + 2 * j)) for j in range(13)])
print("unit 1: MMA2(j) wait end   :", [int(c(3 + 2 * j)) for j in range(13)])
print(f"unit 1: all MMAs issued {c(40):7.0f}")
print(f"unit 0 epilogue: wait start {c(41):.0f}  o_full seen {c(42):.0f}  O drained (o_empty) {c(46):.0f}  boxes staged {c(49):.0f}  "
      f"stage_free {c(51):.0f}  stats {c(52):.0f}")
print(f"unit 1 converts: P(0) ready {c(53):.0f}  P(1) ready {c(54):.0f}")
print(f"producer: unit 1 tile 0 issued {c(56):.0f}  stage_free seen {c(55):.0f}  tile 2 issued {c(57):.0f}")
print(f"unit 1, tile 6: wait start {c(14):.0f}  x_full seen {c(58):.0f}  p_full seen {c(15):.0f}")
print(f"unit 1 MMA1(0): v_full seen {c(59):.0f} issued {c(60):.0f}   MMA1(1): v_full seen {c(61):.0f} issued {c(62):.0f}")
