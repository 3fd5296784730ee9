"""Two forward passes of the head at the bench shape (batch 32, 320x320) -- the target of ncu captures."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from cmpc_refseg_b200.CMPC_model import LSTM_model
from cmpc_refseg_b200.synthetic import make_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
model = LSTM_model(batch_size=B, device=dev)
inp = {k: v.to(dev) for k, v in make_inputs(B, seed=1234).items() if hasattr(v, "to")}
for _ in range(2):
    out = model.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])
torch.cuda.synchronize()
print("ok", float(out["up"].abs().mean()))
