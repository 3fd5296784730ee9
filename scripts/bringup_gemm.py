"""GPU bring-up: tcgen05 GEMM + MUTAN epilogue vs torch fp32 reference on the same fp16-rounded operands."""
import ctypes as C, sys, time
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L

lib = L.lib()
dev = torch.device('cuda:0')
torch.manual_seed(0)
stream = torch.cuda.current_stream().cuda_stream

def run_gemm(M, N, K1, K2=0, act=0, rows_per_sample=None, with_bias=True, with_sbias=False, with_gate=False,
             out_fp32=False, stats=False, rowss=False):
    rps = rows_per_sample or M
    B = (M + rps - 1) // rps
    kp1 = (K1 + 63) // 64 * 64; kp2 = (K2 + 63) // 64 * 64 if K2 else 0
    lda1 = (K1 + 63) // 64 * 64
    a1 = torch.full((M, lda1), float('nan'), device=dev, dtype=torch.float16); a1[:, :K1] = torch.randn(M, K1, device=dev) * 0.5
    a2 = None
    if K2:
        a2 = torch.full((M, kp2), float('nan'), device=dev, dtype=torch.float16); a2[:, :K2] = torch.randn(M, K2, device=dev) * 0.5
    w = torch.zeros(N, kp1 + kp2, device=dev, dtype=torch.float16)
    w[:, :K1] = torch.randn(N, K1, device=dev) * 0.05
    if K2: w[:, kp1:kp1 + K2] = torch.randn(N, K2, device=dev) * 0.05
    npad = (N + 255) // 256 * 256
    bias = torch.randn(npad, device=dev) if with_bias else None
    sbias = torch.randn(B, npad, device=dev) if with_sbias else None
    gate = torch.rand(B, npad, device=dev) if with_gate else None
    ldo = (N + 31) // 32 * 32
    out = torch.full((M, ldo), 777.0, device=dev, dtype=torch.float32 if out_fp32 else torch.float16)
    rs = torch.zeros(M, device=dev) if rowss else None
    st = torch.zeros(B, 1, 2, device=dev, dtype=torch.float64) if stats else None
    args = L.GemmArgs()
    args.a1 = a1.data_ptr(); args.lda1 = lda1; args.k1 = K1
    args.a2 = a2.data_ptr() if K2 else None; args.lda2 = kp2; args.k2 = K2
    args.w = w.data_ptr(); args.ldw = kp1 + kp2
    args.m = M; args.n = N; args.rows_per_sample = rps
    args.bias = bias.data_ptr() if with_bias else None
    args.sbias = sbias.data_ptr() if with_sbias else None; args.ld_sbias = npad
    args.gate = gate.data_ptr() if with_gate else None; args.ld_gate = npad
    args.act = act
    args.out = out.data_ptr(); args.ldo = ldo; args.out_fp32 = int(out_fp32)
    args.row_sumsq = rs.data_ptr() if rowss else None
    args.stats = st.data_ptr() if stats else None
    L.check(lib.cmpc_gemm_f16(C.byref(args), stream), "gemm")
    torch.cuda.synchronize()
    ref = a1[:, :K1].float() @ w[:, :K1].float().t()
    if K2: ref = ref + a2[:, :K2].float() @ w[:, kp1:kp1 + K2].float().t()
    bidx = torch.arange(M, device=dev) // rps
    if with_bias: ref = ref + bias[:N]
    if with_sbias: ref = ref + sbias[bidx, :N]
    if act == 1: ref = torch.relu(ref)
    if with_gate: ref = ref * gate[bidx, :N]
    got = out[:, :N].float()
    err = (got - ref).abs().max().item()
    msg = f"gemm M={M} N={N} K1={K1} K2={K2} act={act} fp32={out_fp32}: max-abs err {err:.3e} (ref max {ref.abs().max().item():.2f})"
    ok = err < (2e-3 if out_fp32 else 2e-2)
    if ldo > N: ok = ok and bool((out[:, N:] == 0).all())
    if rowss:
        e2 = ((rs - (ref ** 2).sum(1)).abs() / (ref ** 2).sum(1)).max().item(); msg += f" rowss rel {e2:.2e}"; ok = ok and e2 < 1e-3
    if stats:
        s1 = torch.zeros(B, device=dev, dtype=torch.float64).index_add_(0, bidx, ref.double().sum(1))
        s2 = torch.zeros(B, device=dev, dtype=torch.float64).index_add_(0, bidx, (ref.double() ** 2).sum(1))
        e3 = ((st[:, 0, 0] - s1).abs().max() / s1.abs().max()).item(); e4 = ((st[:, 0, 1] - s2).abs() / s2).max().item()
        msg += f" stats rel {e3:.2e} {e4:.2e}"; ok = ok and e3 < 1e-3 and e4 < 1e-3
    print(("PASS " if ok else "FAIL ") + msg, flush=True)
    return ok

def run_mutan(M, Cc, K, rps):
    B = (M + rps - 1) // rps
    kp = (K + 63) // 64 * 64
    chunks = (Cc + 47) // 48
    a = torch.full((M, kp), float('nan'), device=dev, dtype=torch.float16); a[:, :K] = torch.randn(M, K, device=dev) * 0.3
    Wk = torch.randn(5, Cc, K, device=dev) * 0.05            # [head, c, k]
    w = torch.zeros(chunks * 240, kp, device=dev, dtype=torch.float16)
    for j in range(chunks):
        for k in range(5):
            c0 = j * 48; c1 = min(c0 + 48, Cc)
            w[j * 240 + k * 48: j * 240 + k * 48 + (c1 - c0), :K] = Wk[k, c0:c1]
    ldb = (Cc + 63) // 64 * 64
    bias = torch.randn(5, ldb, device=dev) * 0.1
    lang = torch.tanh(torch.randn(B, 5, ldb, device=dev)); 
    ldo = ldb
    out = torch.full((M, ldo), 777.0, device=dev)
    rs = torch.zeros(M, device=dev)
    args = L.MutanArgs()
    args.a = a.data_ptr(); args.lda = kp; args.k = K
    args.w = w.data_ptr(); args.ldw = kp
    args.m = M; args.c = Cc; args.rows_per_sample = rps
    args.bias = bias.data_ptr(); args.ld_bias = ldb
    args.lang = lang.data_ptr(); args.ld_lang = ldb
    args.out = out.data_ptr(); args.ldo = ldo; args.row_sumsq = rs.data_ptr()
    L.check(lib.cmpc_mutan_f16(C.byref(args), stream), "mutan")
    torch.cuda.synchronize()
    bidx = torch.arange(M, device=dev) // rps
    pre = torch.einsum('mk,hck->mhc', a[:, :K].float(), w.new_tensor(0).float() + Wk.half().float())
    ref = torch.tanh((torch.tanh(pre + bias[None, :, :Cc]) * lang[bidx][:, :, :Cc]).sum(1))
    err = (out[:, :Cc] - ref).abs().max().item()
    e2 = ((rs - (ref ** 2).sum(1)).abs() / (ref ** 2).sum(1)).max().item()
    ok = err < 2e-3 and e2 < 1e-3 and bool((out[:, Cc:chunks * 48] == 0).all())
    print(("PASS " if ok else "FAIL ") + f"mutan M={M} C={Cc} K={K}: max-abs err {err:.3e} rowss rel {e2:.2e}", flush=True)
    return ok

def bench_gemm(M, N, K, iters=20):
    a = (torch.randn(M, K, device=dev) * 0.5).half(); w = (torch.randn(N, K, device=dev) * 0.05).half()
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    args = L.GemmArgs(); args.a1 = a.data_ptr(); args.lda1 = K; args.k1 = K; args.w = w.data_ptr(); args.ldw = K
    args.m = M; args.n = N; args.rows_per_sample = M; args.out = out.data_ptr(); args.ldo = N
    for _ in range(3): L.check(lib.cmpc_gemm_f16(C.byref(args), stream))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): lib.cmpc_gemm_f16(C.byref(args), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    for _ in range(3): torch.matmul(a, w.t(), out=out)
    e0.record()
    for _ in range(iters): torch.matmul(a, w.t(), out=out)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    fl = 2.0 * M * N * K
    print(f"bench gemm {M}x{N}x{K}: ours {ms:.3f} ms {fl/ms/1e9:.0f} TF/s | cublas {ms2:.3f} ms {fl/ms2/1e9:.0f} TF/s", flush=True)

ok = True
ok &= run_gemm(128, 256, 64, with_bias=False)
ok &= run_gemm(128, 256, 128)
ok &= run_gemm(256, 512, 256, act=1)
ok &= run_gemm(1000, 500, 1000, act=1, rows_per_sample=250, with_sbias=True, with_gate=True, rowss=True, stats=True)
ok &= run_gemm(3200, 1000, 1000, K2=1008, out_fp32=True, rows_per_sample=1600, stats=True)
ok &= run_gemm(51200, 1000, 2048, rows_per_sample=1600, rowss=True)
ok &= run_mutan(256, 96, 64, 128)
ok &= run_mutan(3200, 1000, 1008, 1600)
if ok:
    bench_gemm(51200, 1000, 2048); bench_gemm(51200, 1024, 1024); bench_gemm(51200, 512, 512); bench_gemm(8192, 8192, 8192)
print("ALL PASS" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
