"""Forward (training mode, aux heads) + backward of the head at BASELINE config 5's per-GPU shape (batch 16, 320x320): timings."""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200.CMPC_model import LSTM_model
from cmpc_refseg_b200.backward import HeadBackward, Saved
from cmpc_refseg_b200.synthetic import make_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
model = LSTM_model(batch_size=B, device=dev)
head = model._head
head.saved = Saved(dev)
inp = {k: v.to(dev) for k, v in make_inputs(B, seed=1234).items() if hasattr(v, "to")}
target = (torch.rand(B, 320, 320, 1, device=dev) > 0.7).float()
bw = HeadBackward(head)
torch.cuda.synchronize()
print(f"memory after setup: {torch.cuda.memory_allocated() / 2**30:.2f} GiB")
def step():
    out = head.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"], aux=True)
    bw.backward(out, target)
for _ in range(2): step()
torch.cuda.synchronize()
print(f"memory after warm-up: {torch.cuda.memory_allocated() / 2**30:.2f} GiB, peak {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
n = 5
tf = tb = 0.0
for _ in range(n):
    e[0].record()
    out = head.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"], aux=True)
    e[1].record()
    bw.backward(out, target)
    e[2].record()
    torch.cuda.synchronize()
    tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
print(f"B={B}: forward(train) {tf / n:.2f} ms, backward {tb / n:.2f} ms, total {(tf + tb) / n:.2f} ms -> {B / ((tf + tb) / n) * 1e3:.0f} samples/s")
gn = sum(float((v.double() ** 2).sum()) for v in bw.g.values()) ** 0.5
print("grad norm (packed buffers)", gn)
# coarse breakdown with the torch profiler
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
