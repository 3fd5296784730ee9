"""world_size-2 gloo test of the N > 1 host path: batch sharding + the single IoU-statistics all-reduce must give
exactly the unsharded result (integers bit-exact)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, I_all, U_all, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cmpc_refseg_b200.parallel import local_iou_stats, reduce_iou_stats, shard_range, summarize
    lo, hi = shard_range(rank, world, I_all.numel())
    stats = reduce_iou_stats(local_iou_stats(I_all[lo:hi], U_all[lo:hi]))
    q.put((rank, summarize(stats)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_iou_reduction_equals_unsharded():
    from cmpc_refseg_b200.parallel import local_iou_stats, summarize
    g = torch.Generator().manual_seed(0)
    U = torch.randint(1000, 50000, (13,), generator=g)
    I = (U.double() * torch.rand(13, generator=g, dtype=torch.float64)).long()
    ref = summarize(local_iou_stats(I, U))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, I, U, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert res[r]["cum_I"] == ref["cum_I"] and res[r]["cum_U"] == ref["cum_U"] and res[r]["n"] == 13
        assert abs(res[r]["mean_iou"] - ref["mean_iou"]) < 1e-12
        for th in (0.5, 0.6, 0.7, 0.8, 0.9):
            assert res[r][f"precision@{th}"] == ref[f"precision@{th}"]
