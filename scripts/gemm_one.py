"""One generic GEMM shape of the head, launched a few times on the production library (for `ncu --set full --import-source on`).
usage: python scripts/gemm_one.py lang_se2|gupd|lstm16p [reps]"""
import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
B, N = 32, 1600; M = B * N
CASES = {  # K, Nn, ldo, relu, gate, K2, group, stats, peep
    "lang_se2": dict(K=500, Nn=1024, ldo=1024, relu=1, gate=True, group=(512, 500)),
    "lang_se": dict(K=500, Nn=500, ldo=512, relu=1, gate=True),
    "gupd": dict(K=1000, Nn=1000, ldo=1024, stats=True),
    "lstm16p": dict(K=500, Nn=2048, ldo=2048, K2=500, group=(512, 500), stats=True, peep=True),
    "lstm16b": dict(K=500, Nn=2048, ldo=2048, K2=500, group=(512, 500), stats=True),
    "lstm16ph": dict(K=500, Nn=2048, ldo=2048, K2=500, group=(512, 500), stats=True, peep=True, peep16=True),
}
def make(K, Nn, ldo, relu=0, gate=False, K2=0, group=None, stats=False, peep=False, peep16=False):
    kp = (K + 63) // 64 * 64; kp2 = (K2 + 63) // 64 * 64 if K2 else 0
    a = (torch.randn(M, kp, device=dev) * 0.1).half(); a2 = (torch.randn(M, max(kp2, 64), device=dev) * 0.1).half()
    w = (torch.randn(Nn, kp + kp2, device=dev) * 0.05).half(); bias = torch.randn((Nn + 255) // 256 * 256, device=dev); g = torch.rand(B, ldo, device=dev)
    out = torch.empty(M, ldo, device=dev, dtype=torch.float16)
    ar = L.GemmArgs(); ar.a1 = a.data_ptr(); ar.lda1 = kp; ar.k1 = K
    if K2: ar.a2 = a2.data_ptr(); ar.lda2 = kp2; ar.k2 = K2
    ar.w = w.data_ptr(); ar.ldw = kp + kp2; ar.m = M; ar.n = Nn; ar.rows_per_sample = N; ar.bias = bias.data_ptr(); ar.act = relu
    if gate: ar.gate = g.data_ptr(); ar.ld_gate = ldo
    if group: ar.group_width, ar.group_valid = group
    ar.out = out.data_ptr(); ar.ldo = ldo; ar.out_fp32 = 0
    keep = [a, a2, w, bias, g, out]
    if stats:
        stt = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64); ar.stats = stt.data_ptr(); keep.append(stt)
    if peep:
        pi = torch.randn(N, 512, device=dev); pf = torch.randn(N, 512, device=dev); cp = torch.randn(M, 512, device=dev)
        if peep16:
            pi, pf, cp = pi.half(), pf.half(), cp.half(); ar.peep_f16 = 1
        ar.peep_i, ar.peep_f, ar.ld_peep = pi.data_ptr(), pf.data_ptr(), 512
        ar.cprev, ar.ld_cprev = cp.data_ptr(), 512
        keep += [pi, pf, cp]
    return (lambda: L.check(lib.cmpc_gemm_f16(C.byref(ar), st))), keep
name = sys.argv[1] if len(sys.argv) > 1 else "lang_se2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
fn, keep = make(**CASES[name])
for _ in range(3): fn()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): fn()
e1.record(); torch.cuda.synchronize()
print(f"{name}: {e0.elapsed_time(e1) * 1000 / reps:.1f} us per launch")
