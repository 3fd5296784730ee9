"""Probe: CUDA events recorded INSIDE a captured graph (torch.cuda.Event(enable_timing=True, external=True)) and read after a replay."""
import torch
dev = torch.device("cuda:0")
a = torch.randn(8192, 8192, device=dev, dtype=torch.float16)
b = torch.randn(8192, 8192, device=dev, dtype=torch.float16)
c = torch.empty_like(a)
torch.matmul(a, b, out=c); torch.cuda.synchronize()
s = torch.cuda.Stream()
evs = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(4)]
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    cur = torch.cuda.current_stream()
    evs[0].record(cur)
    torch.matmul(a, b, out=c)
    evs[1].record(cur)
    c.mul_(2.0)
    evs[2].record(cur)
    torch.matmul(a, b, out=c)
    evs[3].record(cur)
for rep in range(3):
    g.replay()
    torch.cuda.synchronize()
    print(rep, [round(evs[i].elapsed_time(evs[i + 1]), 4) for i in range(3)])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record(); torch.cuda.synchronize()
print("10 replays", e0.elapsed_time(e1) / 10, "last in-graph:", [round(evs[i].elapsed_time(evs[i + 1]), 4) for i in range(3)])
