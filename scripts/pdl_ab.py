"""A/B of programmatic dependent launch (CMPC_PDL) on the forward pass: python scripts/pdl_ab.py  (run once per setting)"""
import os, sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200.CMPC_model import LSTM_model
from cmpc_refseg_b200.synthetic import make_inputs
dev = torch.device("cuda:0")
for B in (32, 1):
    inp = {k: v.to(dev) for k, v in make_inputs(B, seed=1234).items() if hasattr(v, "to")}
    for graph in (False, True):
        model = LSTM_model(batch_size=B, device=dev, cuda_graph=graph)
        for _ in range(4):
            out = model.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
        torch.cuda.synchronize()
        n = 40 if B == 32 else 200
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = model.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
        e1.record(); torch.cuda.synchronize()
        print(f"CMPC_PDL={os.environ.get('CMPC_PDL', '0')} B={B} cuda_graph={graph}: {e0.elapsed_time(e1) / n:.3f} ms per forward   checksum {float(out['up'].double().abs().mean()):.6f}")
        del model
