"""CUDA-event timings of the HBM-bound row kernels at the bench shape (batch 32, N = 1600): python scripts/rowbench.py"""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
B, N, C, Mm, GW, LDC = 32, 1600, 1000, 500, 512, 1024
M = B * N
f16 = lambda *s: (torch.randn(*s, device=dev) * 0.1).half()
x16, y16, u16, o16 = f16(M, LDC), f16(M, LDC), f16(M, LDC), f16(M, LDC)
gamma, beta = torch.ones(LDC, device=dev), torch.zeros(LDC, device=dev)
mr = torch.zeros(B, 2, device=dev); mr[:, 1] = 1.0
fa, fb, fc, fo = f16(M, GW), f16(M, GW), f16(M, GW), f16(M, GW)
u = torch.randn(B, 6 * GW, device=dev); pool = torch.zeros(B, 3, GW, device=dev)
ws = torch.zeros(max(lib.cmpc_global_pool_workspace_bytes(B, 3, GW), 1 << 20), dtype=torch.uint8, device=dev)
big = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)

def timeit(name, fn, nbytes):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        big.zero_()                      # flush L2
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:22s} {t:7.1f} us   {nbytes / t / 1e3:7.0f} GB/s (algorithmic bytes)")

ck = L.check
timeit("ln_residual_relu", lambda: ck(lib.cmpc_ln_residual_relu_f16(y16.data_ptr(), LDC, x16.data_ptr(), LDC, mr.data_ptr(), gamma.data_ptr(),
       beta.data_ptr(), o16.data_ptr(), LDC, M, C, N, st)), M * C * 2 * 3)
lib.cmpc_ln_relu_l2norm_set_mode(1)
timeit("ln_relu_l2norm (regs)", lambda: ck(lib.cmpc_ln_relu_l2norm_f16(u16.data_ptr(), LDC, mr.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
       o16.data_ptr(), LDC, M, C, 40, 40, N, 1, None, st)), M * C * 2 * 2)
for mode in (2, 3, 4):
    lib.cmpc_ln_relu_l2norm_set_mode(mode)
    timeit(f"ln_relu_l2norm (mode {mode})", lambda: ck(lib.cmpc_ln_relu_l2norm_f16(u16.data_ptr(), LDC, mr.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
           o16.data_ptr(), LDC, M, C, 40, 40, N, 1, None, st)), M * C * 2 * 2)
lib.cmpc_ln_relu_l2norm_set_mode(0)
timeit("ln_relu_l2norm", lambda: ck(lib.cmpc_ln_relu_l2norm_f16(u16.data_ptr(), LDC, mr.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
       o16.data_ptr(), LDC, M, C, 40, 40, N, 1, None, st)), M * C * 2 * 2)
timeit("add3_l2norm", lambda: ck(lib.cmpc_add3_l2norm_f16(fa.data_ptr(), fb.data_ptr(), fc.data_ptr(), GW, fo.data_ptr(), GW, M, GW, 1, None, st)),
       M * Mm * 2 * 4)
timeit("global_pool (3 maps)", lambda: ck(lib.cmpc_global_pool_f16(fa.data_ptr(), fb.data_ptr(), fc.data_ptr(), GW, u.data_ptr(), GW, 6 * GW, 3, B, N, GW,
       0.0447, pool.data_ptr(), GW, None, ws.data_ptr(), ws.numel(), st)), M * Mm * 2 * 3)
# ConvLSTM gate passes (inference form: o' recomputed from the fp16 gate map)
y16 = f16(M, 4 * GW); cprev = torch.randn(M, GW, device=dev); cnew = torch.empty(M, GW, device=dev); cst = torch.empty(M, GW, device=dev)
h16 = torch.empty(M, GW, device=dev, dtype=torch.float16); wco = torch.randn(N, GW, device=dev)
g5, b5 = torch.ones(5, GW, device=dev), torch.zeros(5, GW, device=dev)
mr4 = torch.zeros(B, 4, 2, device=dev); mr4[:, :, 1] = 1.0; mr2 = torch.zeros(B, 2, 2, device=dev); mr2[:, :, 1] = 1.0
so = torch.zeros(B, 2, 2, device=dev, dtype=torch.float64)
timeit("convlstm_gates1", lambda: ck(lib.cmpc_convlstm_gates1(y16.data_ptr(), 1, 4 * GW, GW, Mm, mr4.data_ptr(), g5.data_ptr(), b5.data_ptr(), cprev.data_ptr(),
       wco.data_ptr(), cnew.data_ptr(), None, so.data_ptr(), M, N, st)), M * Mm * (8 + 4 + 4))
timeit("convlstm_gates2_y16", lambda: ck(lib.cmpc_convlstm_gates2_y16(y16[:, 3 * GW:].data_ptr(), 4 * GW, wco.data_ptr(), cnew.data_ptr(), GW, Mm, mr2.data_ptr(),
       g5.data_ptr(), b5.data_ptr(), cst.data_ptr(), h16.data_ptr(), 0, M, N, st)), M * Mm * (2 + 4 + 4 + 2))
cprev16, cnew16, cst16 = cprev.half(), torch.empty(M, GW, device=dev, dtype=torch.float16), torch.empty(M, GW, device=dev, dtype=torch.float16)
timeit("convlstm_gates1_h16", lambda: ck(lib.cmpc_convlstm_gates1_h16(y16.data_ptr(), 4 * GW, GW, Mm, mr4.data_ptr(), g5.data_ptr(), b5.data_ptr(), cprev16.data_ptr(),
       wco.data_ptr(), cnew16.data_ptr(), so.data_ptr(), M, N, st)), M * Mm * (8 + 2 + 2))
timeit("convlstm_gates2_y16 (f16 state)", lambda: ck(lib.cmpc_convlstm_gates2_y16(y16[:, 3 * GW:].data_ptr(), 4 * GW, wco.data_ptr(), cnew16.data_ptr(), GW, Mm, mr2.data_ptr(),
       g5.data_ptr(), b5.data_ptr(), cst16.data_ptr(), h16.data_ptr(), 1, M, N, st)), M * Mm * (2 + 2 + 2 + 2))
