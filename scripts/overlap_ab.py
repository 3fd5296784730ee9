"""A/B of CMPCHeadB200.overlap_lang (language-side chains on parallel streams): forward time at batch 32 and batch 1."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200.CMPC_model import LSTM_model
from cmpc_refseg_b200.synthetic import make_inputs
dev = torch.device("cuda:0")
for B in (32, 1):
    model = LSTM_model(batch_size=B, device=dev, cuda_graph=('graph' in sys.argv))
    inp = {k: v.to(dev) for k, v in make_inputs(B, seed=1234).items() if hasattr(v, "to")}
    outs = {}
    for rep in range(2):
        for ov in (False, True):
            model._head.overlap_lang = ov
            model._head.__dict__.pop('_graphs', None)
            for _ in range(3):
                out = model.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            n = 20 if B == 32 else 100
            e0.record()
            for _ in range(n):
                out = model.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
            e1.record(); torch.cuda.synchronize()
            outs[ov] = out["up"].clone()
            print(f"B={B} overlap={ov}: {e0.elapsed_time(e1) / n:.3f} ms per forward")
    print("   max diff of up:", float((outs[True] - outs[False]).abs().max()))
