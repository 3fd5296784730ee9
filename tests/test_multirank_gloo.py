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


# ---- gradient buckets of the training step (HeadBackward's arena layout + the asynchronous per-bucket all-reduce) ----------
def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cmpc_refseg_b200.backward import HeadBackward
    from cmpc_refseg_b200.parallel import BucketReducer, plan_arena
    spec = _toy_spec()
    total, place, ranges = plan_arena(spec, HeadBackward.bucket_of, HeadBackward.BUCKETS)
    arena = torch.zeros(total, dtype=torch.float32)
    views = {k: arena[o:o + n] for k, (o, n) in place.items()}
    red = BucketReducer(arena, ranges)
    for bname in HeadBackward.BUCKETS:                     # "backward": fill the buffers of a bucket, then hand it to the reducer
        for k, v in views.items():
            if HeadBackward.bucket_of(k) == bname:
                v.copy_(torch.arange(v.numel(), dtype=torch.float32) * (rank + 1) + len(k))
        red.reduce(bname)
    red.wait()
    ok = all(torch.equal(v, torch.arange(v.numel(), dtype=torch.float32) * 3 + 2 * len(k)) for k, v in views.items())
    q.put((rank, ok, len(red.works)))
    dist.barrier()
    dist.destroy_process_group()


def _toy_spec():
    names = ["lstm_w", "lstm_W_ci", "score_w9", "score_b", "se_w_c3_f1", "se_b_c4_2_f2", "score_c5_w9", "wf1", "key", "q_w", "gvl_b"]
    for lvl in ("c5", "c4", "c3"):
        names += [f"fusion_w_{lvl}", f"fusion_lang_{lvl}", f"gupd_w_{lvl}", f"gfeat_gamma_{lvl}", f"gupdate_beta_{lvl}", f"gt_w_{lvl}",
                  f"mutan_w_{lvl}", f"mutan_b_{lvl}", f"lat_w_{lvl}", f"lat_b_{lvl}", f"ltrans_w_{lvl}", f"ltrans_b_{lvl}",
                  f"wtrans_w_{lvl}", f"wtrans_b_{lvl}"]
    names += ["parse1_w", "parse1_b", "parse2_w", "parse2_b"]
    return {k: (3 + i % 5, 7 + i % 11) for i, k in enumerate(names)}


def test_gradient_buckets_tile_the_arena_in_backward_order():
    from cmpc_refseg_b200.backward import HeadBackward
    from cmpc_refseg_b200.parallel import plan_arena
    spec = _toy_spec()
    total, place, ranges = plan_arena(spec, HeadBackward.bucket_of, HeadBackward.BUCKETS)
    edges = [ranges[b] for b in HeadBackward.BUCKETS]
    assert edges[0][0] == 0 and edges[-1][1] == total and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert all(b > a for a, b in edges)                                   # every stage of the backward owns something
    for k, (o, n) in place.items():
        a, b = ranges[HeadBackward.bucket_of(k)]
        assert a <= o and o + n <= b and o % 64 == 0
    spans = sorted(place.values())
    assert all(o1 + n1 <= o2 for (o1, n1), (o2, _) in zip(spans, spans[1:]))       # no overlap
    assert HeadBackward.bucket_of("ltrans_w_c4") == "c4_ltrans" and HeadBackward.bucket_of("mutan_w_c4") == "c4_mutan" and HeadBackward.bucket_of("wtrans_w_c4") == "language"
    assert HeadBackward.bucket_of("score_c3_w9") == "exchange" and HeadBackward.bucket_of("score_w9") == "fuse"


def test_bucketed_async_allreduce_sums_every_bucket():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
