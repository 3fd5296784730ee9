"""Where a training step's time goes after the backward: gradient re-layout, Adam, operand re-pack (batch 16, 320x320)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200.CMPC_model import LSTM_model
from cmpc_refseg_b200.synthetic import make_inputs
B = 16
dev = torch.device("cuda:0")
model = LSTM_model(batch_size=B, device=dev, mode='train')
tr = model.train_op()
head = model._head
inp = {k: v.to(dev) for k, v in make_inputs(B, seed=1234).items() if hasattr(v, "to")}
a = (inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"], inp["target_fine"])
for _ in range(2): tr.train_step(*a, report_loss=False)
def t(fn, n=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
out = head.forward(*a[:4], aux=True)
print(f"forward(train)      {t(lambda: head.forward(*a[:4], aux=True)):.2f} ms")
print(f"backward            {t(lambda: tr.bw.backward(out, a[4])):.2f} ms")
print(f"grads -> flat views {t(lambda: tr.bw.grads_tf(into=tr.grads)):.2f} ms")
from cmpc_refseg_b200.weights import pack_head_weights
print(f"repack (head)       {t(lambda: pack_head_weights(tr.params, head.d, head.device, out=head.Wt)):.2f} ms")
print(f"repack (backward)   {t(tr.bw.pack_weights):.2f} ms")
print(f"whole step          {t(lambda: tr.train_step(*a, report_loss=False)):.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.bw.backward(out, a[4]); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=50))
