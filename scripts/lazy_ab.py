"""A/B of head.lazy_norm (deferred l2_normalize of the MUTAN map) on the replayed forward, alternating in one process."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200.CMPC_model import LSTM_model
from cmpc_refseg_b200.synthetic import make_inputs
dev = torch.device("cuda:0")
B = 32
inp = {k: v.to(dev) for k, v in make_inputs(B, seed=1234).items() if hasattr(v, "to")}
models = {}
for lazy in (False, True):
    m = LSTM_model(batch_size=B, device=dev, cuda_graph=True)
    m._head.lazy_norm = lazy
    for _ in range(4):
        out = m.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
    torch.cuda.synchronize()
    models[lazy] = m
for rep in range(4):
    for lazy in (False, True):
        m = models[lazy]
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            out = m.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
        e1.record(); torch.cuda.synchronize()
        print(f"rep {rep} lazy_norm={lazy}: {e0.elapsed_time(e1) / 40:.3f} ms per replayed forward   checksum {float(out['up'].double().abs().mean()):.6f}")
