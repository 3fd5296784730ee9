"""Stand-alone launches of the hot kernels at bench shapes (for ncu captures and quick timing).
usage: python scripts/kbench.py [lang_se|gupd|lateral|lstm|mutan|graph|all] [iters]"""
import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
B, N = 32, 1600; M = B * N
torch.manual_seed(0)
def timeit(name, fn, flops):
    for _ in range(2): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:10s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s", flush=True)
def gemm_case(name, K, Nn, ldo, *, relu=0, gate=False, fp32=False, rowss=False, stats=False, K2=0, group=None):
    kp = (K + 63) // 64 * 64; kp2 = (K2 + 63) // 64 * 64 if K2 else 0
    a = (torch.randn(M, kp, device=dev) * 0.1).half(); a2 = (torch.randn(M, kp2, device=dev) * 0.1).half() if K2 else None
    w = (torch.randn(Nn, kp + kp2, device=dev) * 0.05).half()
    bias = torch.randn((Nn + 255) // 256 * 256, device=dev); g = torch.rand(B, ldo, device=dev)
    out = torch.empty(M, ldo, device=dev, dtype=torch.float32 if fp32 else torch.float16)
    rs = torch.zeros(M, device=dev); stt = torch.zeros(B * 8, device=dev, dtype=torch.float64)
    ar = L.GemmArgs(); ar.a1 = a.data_ptr(); ar.lda1 = kp; ar.k1 = K
    if K2: ar.a2 = a2.data_ptr(); ar.lda2 = kp2; ar.k2 = K2
    ar.w = w.data_ptr(); ar.ldw = kp + kp2; ar.m = M; ar.n = Nn; ar.rows_per_sample = N
    ar.bias = bias.data_ptr(); ar.act = relu
    if gate: ar.gate = g.data_ptr(); ar.ld_gate = ldo
    if group: ar.group_width, ar.group_valid = group
    ar.out = out.data_ptr(); ar.ldo = ldo; ar.out_fp32 = int(fp32)
    if rowss: ar.row_sumsq = rs.data_ptr()
    if stats: ar.stats = stt.data_ptr()
    nv = Nn if not group else Nn // group[0] * group[1]
    timeit(name, lambda: L.check(lib.cmpc_gemm_f16(C.byref(ar), st)), 2.0 * M * nv * (K + K2))
if which in ('lang_se', 'all'): gemm_case('lang_se', 500, 500, 512, relu=1, gate=True)
if which in ('gupd', 'all'): gemm_case('gupd', 1000, 1000, 1024, stats=True)
if which in ('lateral', 'all'): gemm_case('lateral', 2048, 1000, 1024, fp32=True, rowss=True)
if which in ('fusion', 'all'): gemm_case('fusion', 1000, 500, 512, relu=1, K2=1008)
if which in ('lstm', 'all'): gemm_case('lstm', 500, 2048, 2048, fp32=True, stats=True, K2=500, group=(512, 500))
if which in ('mutan', 'all'):
    a = (torch.randn(M, 1024, device=dev) * 0.03).half(); w = (torch.randn(21 * 240, 1024, device=dev) * 0.03).half()
    bias = torch.zeros(5, 1024, device=dev); lang = torch.rand(B, 5, 1000, device=dev); out = torch.empty(M, 1024, device=dev); rs = torch.zeros(M, device=dev)
    ma = L.MutanArgs(); ma.a = a.data_ptr(); ma.lda = 1024; ma.k = 1008; ma.w = w.data_ptr(); ma.ldw = 1024; ma.m = M; ma.c = 1000
    ma.rows_per_sample = N; ma.bias = bias.data_ptr(); ma.ld_bias = 1024; ma.lang = lang.data_ptr(); ma.ld_lang = 1000
    ma.out = out.data_ptr(); ma.ldo = 1024; ma.row_sumsq = rs.data_ptr()
    timeit('mutan', lambda: L.check(lib.cmpc_mutan_f16(C.byref(ma), st)), 2.0 * M * 5000 * 1008)
if which in ('graph', 'all'):
    w16 = torch.rand(M, 32, device=dev).half(); v16 = torch.rand(M, 32, device=dev).half()
    x = (torch.randn(M, 1024, device=dev) * 0.03).half(); y = torch.empty_like(x); stt = torch.zeros(B, 2, device=dev, dtype=torch.float64)
    timeit('graph', lambda: L.check(lib.cmpc_graph_reason_f16(w16.data_ptr(), v16.data_ptr(), x.data_ptr(), 1024, B, N, 1000, 2048.0, y.data_ptr(), 1024, stt.data_ptr(), None, st)),
           B * (2.0 * N * N * 20 + 2.0 * N * N * 1000))
