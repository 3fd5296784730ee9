"""Data-parallel training step, sharded == unsharded: two ranks with half of the batch each, bucketed
asynchronous gradient all-reduce, must produce the gradient (and the parameters after the step) of one rank with the whole batch.
Runs on ONE GPU (both ranks on cuda:0 over gloo, which reduces CUDA tensors through host staging) so that the single-GPU test tier
covers it; with two or more GPUs visible each rank takes its own device over NCCL."""
import os
import socket
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

TINY = dict(num_steps=20, vf_h=8, vf_w=8, H=64, W=64, vf_dim=128, c4_dim=64, c3_dim=32, v_emb_dim=64, rnn_size=64,
            mlp_dim=32, parse_hidden=40)
B, SEQ = 4, [20, 6, 11, 3]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(batch, dev, lo=0, hi=B):
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from oracle.cmpc_head_ref import HeadConfig, init_params, make_inputs
    cfg = HeadConfig(batch_size=B, **TINY)
    params = init_params(cfg, 0, sharp=6.0, bias_std=0.05, ln_jitter=0.2)
    inp = make_inputs(cfg, B, seed=17, seq_len=SEQ)
    g = torch.Generator().manual_seed(3)
    target = (torch.rand(B, cfg.H, cfg.W, 1, generator=g) > 0.6).float()
    hk = {k: TINY[k] for k in ("c4_dim", "c3_dim", "parse_hidden")}
    mk = {k: v for k, v in TINY.items() if k not in hk}
    model = LSTM_model(batch_size=batch, params=params, device=dev, head_kwargs=hk, mode="train", start_lr=1e-3, **mk)
    args = [inp[k][lo:hi].contiguous().to(dev) for k in ("c3", "c4", "c5", "lstm_outputs")] + [target[lo:hi].contiguous().to(dev)]
    return model, args


def _worker(rank, world, port, path, graph):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    multi = torch.cuda.device_count() >= world
    dev = torch.device("cuda", rank if multi else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if multi else "gloo", rank=rank, world_size=world)
    from cmpc_refseg_b200.parallel import gv_sum_allreduce, shard_range
    lo, hi = shard_range(rank, world, B)
    model, args = _setup(hi - lo, dev, lo, hi)
    tr = model.train_op()
    assert tr.world == world and model.gv_norm == "batch"
    ref = torch.load(path)
    if not graph:
        # the literal batch-coupled l2_normalize (CMPC_model.py:241) over the GLOBAL batch: exchange the per-module sums
        model._head.gv_allreduce = gv_sum_allreduce()
    tr.train_step(*args, graph=graph)
    torch.cuda.synchronize()
    grad = tr.grad.double().cpu() / world
    theta = tr.theta.double().cpu()
    if not graph:
        rel = float((grad - ref["grad"]).norm() / ref["grad"].norm())
        assert rel < 2e-4, f"rank {rank}: sharded gradient differs from the unsharded one: rel-L2 {rel:.3e}"
        moved = float((theta - ref["theta"]).abs().max())
        assert moved <= 2.1e-3                       # Adam's first step is +-lr per coordinate; sign flips only where g ~ 0
        same = float(((theta - ref["theta0"]).sign() == (ref["theta"] - ref["theta0"]).sign()).double().mean())
        assert same > 0.99, same
    # every rank ends the step with identical parameters
    mine = tr.theta.clone()
    other = tr.theta.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(mine, other), f"rank {rank}: parameters diverged across ranks"
    # second step (graph mode: replays the captured segments with a fresh all-reduce between them)
    tr.train_step(*args, graph=graph)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(tr.theta).all())
    other = tr.theta.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(tr.theta, other)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph_segments"])
def test_two_ranks_half_batch_equal_one_rank_whole_batch(graph):
    dev = torch.device("cuda:0")
    model, args = _setup(B, dev)
    tr = model.train_op()
    theta0 = tr.theta.double().cpu()
    tr.train_step(*args)
    torch.cuda.synchronize()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "ref.pt")
        torch.save(dict(grad=tr.grad.double().cpu(), theta=tr.theta.double().cpu(), theta0=theta0), path)
        del tr, model, args
        torch.cuda.empty_cache()
        port = _free_port()
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_worker, args=(r, 2, port, path, graph)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=600)
            assert p.exitcode == 0


def test_head_on_a_non_current_device():
    """a head built for cuda:1 must run while cuda:0 is the current device (the C ABI identifies the GPU with cudaGetDevice
    and keeps per-device function attributes), and must agree with the same head on cuda:0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from oracle.cmpc_head_ref import HeadConfig, init_params, make_inputs
    cfg = HeadConfig(batch_size=2, **TINY)
    params = init_params(cfg, 0, sharp=6.0, bias_std=0.05, ln_jitter=0.2)
    inp = make_inputs(cfg, 2, seed=17, seq_len=[20, 6])
    hk = {k: TINY[k] for k in ("c4_dim", "c3_dim", "parse_hidden")}
    mk = {k: v for k, v in TINY.items() if k not in hk}
    torch.cuda.set_device(0)
    outs = []
    for dev in ("cuda:1", "cuda:0"):
        model = LSTM_model(batch_size=2, params=params, device=torch.device(dev), head_kwargs=hk, **mk)
        assert torch.cuda.current_device() == 0
        out = model.forward(*[inp[k].to(dev) for k in ("c3", "c4", "c5", "lstm_outputs")])
        torch.cuda.synchronize(dev)
        assert out["pred"].device == torch.device(dev) and torch.cuda.current_device() == 0
        outs.append(out["pred"].float().cpu())
    assert float((outs[0] - outs[1]).abs().max()) < 2e-3          # same kernels, atomics-order noise only
