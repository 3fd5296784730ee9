import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
L.LIB_PATH = L._PKG / (sys.argv[1] if len(sys.argv) > 1 else "libcmpc_b200_timing.so")
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
dbg = torch.zeros(148, 8, device=dev, dtype=torch.int64)
lib.cmpc_gemm_set_debug.argtypes = [C.c_void_p]; lib.cmpc_gemm_set_debug.restype = None
lib.cmpc_gemm_set_debug(dbg.data_ptr())
B, N = 32, 1600; M = B * N
def case(name, K, Nn, ldo, relu=0, gate=False, fp32=False, K2=0, group=None, mutan=False, stats=False, peep=False):
    kp = (K + 63) // 64 * 64; kp2 = (K2 + 63) // 64 * 64 if K2 else 0
    a = (torch.randn(M, kp, device=dev) * 0.1).half(); a2 = (torch.randn(M, max(kp2, 64), device=dev) * 0.1).half()
    if mutan:
        w = (torch.randn(21 * 240, 1024, device=dev) * 0.03).half(); bias = torch.zeros(5, 1024, device=dev); lang = torch.rand(B, 5, 1000, device=dev)
        out = torch.empty(M, 1024, device=dev); rs = torch.zeros(M, device=dev)
        ma = L.MutanArgs(); ma.a = a.data_ptr(); ma.lda = kp; ma.k = 1008; ma.w = w.data_ptr(); ma.ldw = 1024; ma.m = M; ma.c = 1000
        ma.rows_per_sample = N; ma.bias = bias.data_ptr(); ma.ld_bias = 1024; ma.lang = lang.data_ptr(); ma.ld_lang = 1000
        ma.out = out.data_ptr(); ma.ldo = 1024; ma.row_sumsq = rs.data_ptr()
        fn = lambda: L.check(lib.cmpc_mutan_f16(C.byref(ma), st))
    else:
        w = (torch.randn(Nn, kp + kp2, device=dev) * 0.05).half(); bias = torch.randn((Nn + 255) // 256 * 256, device=dev); g = torch.rand(B, ldo, device=dev)
        out = torch.empty(M, ldo, device=dev, dtype=torch.float32 if fp32 else torch.float16)
        ar = L.GemmArgs(); ar.a1 = a.data_ptr(); ar.lda1 = kp; ar.k1 = K
        if K2: ar.a2 = a2.data_ptr(); ar.lda2 = kp2; ar.k2 = K2
        ar.w = w.data_ptr(); ar.ldw = kp + kp2; ar.m = M; ar.n = Nn; ar.rows_per_sample = N; ar.bias = bias.data_ptr(); ar.act = relu
        if gate: ar.gate = g.data_ptr(); ar.ld_gate = ldo
        if group: ar.group_width, ar.group_valid = group
        ar.out = out.data_ptr(); ar.ldo = ldo; ar.out_fp32 = int(fp32)
        if stats:
            stt = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64); ar.stats = stt.data_ptr()
        if peep:
            pi = torch.randn(N, 512, device=dev); pf = torch.randn(N, 512, device=dev); cp = torch.randn(M, 512, device=dev)
            ar.peep_i, ar.peep_f, ar.ld_peep = pi.data_ptr(), pf.data_ptr(), 512
            ar.cprev, ar.ld_cprev = cp.data_ptr(), 512
        fn = lambda: L.check(lib.cmpc_gemm_f16(C.byref(ar), st))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 200
    d = dbg.cpu().double()
    tiles = d[:, 4].mean().item()
    tot = d[:, 0].max().item(); tiles_l = d[:, 4].max().item()
    print(f"{name:8s} {us:7.1f} us | leader MMA warp: total {tot:9.0f} clk ({tot / max(us, 1e-9) / 1e3:.2f} GHz if it spanned the kernel), tiles {tiles_l:.0f}, "
          f"wait TMA {d[:,1].max():8.0f} wait tmem_empty {d[:,2].max():8.0f} | epilogue warp 4: wait acc {d[:,5].median():8.0f} compute {d[:,6].median():8.0f} "
          f"-> per tile: compute {d[:,6].median()/max(tiles_l,1):6.0f} (tmem ld wait {d[:,3].median()/max(tiles_l,1):5.0f}, wait prev store {d[:,7].median()/max(tiles_l,1):5.0f})")
case('lateral', 2048, 1000, 1024, fp32=True)
case('gupd', 1000, 1000, 1024)
case('fusion', 1000, 500, 512, relu=1, K2=1008)
case('lang_se', 500, 500, 512, relu=1, gate=True)
case('lang_se2', 500, 1024, 1024, relu=1, gate=True, group=(512, 500))
case('lstm', 500, 2048, 2048, fp32=True, K2=500, group=(512, 500))
case('lstm16a', 500, 2048, 2048, group=(512, 500), stats=True)
case('lstm16b', 500, 2048, 2048, K2=500, group=(512, 500), stats=True)
case('lstm16p', 500, 2048, 2048, K2=500, group=(512, 500), stats=True, peep=True)
case('lstm16n', 500, 2048, 2048, K2=500, group=(512, 500))
case('mutan', 1008, 0, 0, mutan=True)
