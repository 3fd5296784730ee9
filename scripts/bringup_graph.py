"""GPU bring-up: graph-reasoning kernel (S=W V^T -> P -> O=P X) vs torch, with the adjacency dump."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200 import _lib as L
lib = L.lib(); dev = torch.device('cuda:0'); st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
ok_all = True
for (B, N, C, T) in [(1, 128, 256, 20), (2, 200, 64, 20), (2, 1600, 1000, 20)]:
    ldx = (C + 8 + 63) // 64 * 64
    w = torch.zeros(B * N, 32, device=dev); v = torch.zeros(B * N, 32, device=dev)
    w[:, :T] = torch.softmax(torch.randn(B * N, T, device=dev) * 2, -1)
    vv = torch.softmax(torch.randn(B, N, T, device=dev) * 2, 1).reshape(B * N, T)
    vs = float(1 << (N - 1).bit_length())
    v[:, :T] = vv * vs
    x = torch.full((B * N, ldx), 3.0, device=dev, dtype=torch.float16); x[:, :C] = (torch.randn(B * N, C, device=dev) * 0.03).half()
    w16, v16 = w.half(), v.half()
    y = torch.full((B * N, ldx), 7.0, device=dev, dtype=torch.float16)
    stats = torch.zeros(B, 2, device=dev, dtype=torch.float64)
    dbg = torch.zeros(B, N, N, device=dev)
    L.check(lib.cmpc_graph_reason_f16(w16.data_ptr(), v16.data_ptr(), x.data_ptr(), ldx, B, N, C, vs, y.data_ptr(), ldx, stats.data_ptr(), dbg.data_ptr(), st), "graph")
    torch.cuda.synchronize()
    Wf, Vf, Xf = w16.float().view(B, N, 32), v16.float().view(B, N, 32), x[:, :C].float().view(B, N, C)
    P = Wf @ Vf.transpose(1, 2)
    ep = ((dbg * vs) - P).abs().max().item()
    Yref = (P.half().float() @ Xf) / vs
    ey = (y[:, :C].float().view(B, N, C) - Yref).abs().max().item()
    s1 = Yref.double().sum((1, 2)); s2 = (Yref.double() ** 2).sum((1, 2))
    es = max(((stats[:, 0] - s1).abs() / s1.abs().clamp_min(1e-9)).max().item(), ((stats[:, 1] - s2).abs() / s2).max().item())
    pad_ok = bool((y[:, C:(C + 7) // 8 * 8] == 0).all())
    ok = ep < 2e-3 * P.abs().max().item() and ey < 2e-3 * Yref.abs().max().item() + 1e-6 and es < 2e-3
    ok_all &= ok
    print(("PASS" if ok else "FAIL"), f"B={B} N={N} C={C}: P err {ep:.3e} (max {P.abs().max().item():.3f})  Y err {ey:.3e} (max {Yref.abs().max().item():.3e}) stats rel {es:.2e}", flush=True)
if ok_all:
    B, N, C = 32, 1600, 1000
    ldx = 1024
    w16 = torch.rand(B * N, 32, device=dev).half(); v16 = torch.rand(B * N, 32, device=dev).half()
    x = (torch.randn(B * N, ldx, device=dev) * 0.03).half(); y = torch.empty_like(x)
    stats = torch.zeros(B, 2, device=dev, dtype=torch.float64)
    for _ in range(3): lib.cmpc_graph_reason_f16(w16.data_ptr(), v16.data_ptr(), x.data_ptr(), ldx, B, N, C, 2048.0, y.data_ptr(), ldx, stats.data_ptr(), None, st)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): lib.cmpc_graph_reason_f16(w16.data_ptr(), v16.data_ptr(), x.data_ptr(), ldx, B, N, C, 2048.0, y.data_ptr(), ldx, stats.data_ptr(), None, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = B * (2.0 * N * N * 20 + 2.0 * N * N * C)
    print(f"graph kernel B=32 N=1600: {ms:.3f} ms  dense F_graph {fl/ms/1e9:.0f} TFLOP/s ({fl/ms/1e9/1671.9*100:.1f}% of measured bf16 peak)")
print("ALL PASS" if ok_all else "SOME FAILED")
sys.exit(0 if ok_all else 1)
