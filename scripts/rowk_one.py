"""One launch sequence of ln_relu_l2norm at the bench shape (for ncu): python scripts/rowk_one.py [mode]"""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
B, N, C, LDC = 32, 1600, 1000, 1024
M = B * N
u16 = (torch.randn(M, LDC, device=dev) * 0.1).half(); o16 = torch.empty_like(u16)
gamma, beta = torch.ones(LDC, device=dev), torch.zeros(LDC, device=dev)
mr = torch.zeros(B, 2, device=dev); mr[:, 1] = 1.0
lib.cmpc_ln_relu_l2norm_set_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for _ in range(3):
    L.check(lib.cmpc_ln_relu_l2norm_f16(u16.data_ptr(), LDC, mr.data_ptr(), gamma.data_ptr(), beta.data_ptr(), o16.data_ptr(), LDC, M, C, 40, 40, N, 1, None, st))
torch.cuda.synchronize()
