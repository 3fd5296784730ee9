"""Batch sharding and the collectives of the path: the IoU-statistics all-reduce of inference, the bucketed gradient all-reduce
of the training step (BASELINE configs[4]) and the optional exchange of the batch-coupled gv_lang norm.

The head is independent per sample, so N GPUs = N ranks each running the whole head on its slice of
the batch, with replicated weights and no activation exchange.  The only cross-rank step is the evaluation
bookkeeping of trainval_model.py:267-294: cumulative I and U, the sum of per-sample IoUs, the five precision@X
counters and the sample count -- nine numbers, summed with one all-reduce (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

EVAL_SEG_IOU = (0.5, 0.6, 0.7, 0.8, 0.9)     # trainval_model.py:159


def shard_range(rank: int, world: int, global_batch: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_iou_stats(I: torch.Tensor, U: torch.Tensor) -> torch.Tensor:
    """[cum_I, cum_U, sum_i I_i/U_i, prec@.5 .. prec@.9 (5), count] as float64 (exact for integers < 2^53)."""
    I = I.to(torch.float64)
    U = U.to(torch.float64)
    iou = I / U                                   # U = 0 -> nan, exactly like the reference's I / U
    vals = [I.sum(), U.sum(), iou.sum()] + [(iou >= th).sum().to(torch.float64) for th in EVAL_SEG_IOU]
    vals.append(torch.tensor(float(I.numel()), dtype=torch.float64, device=I.device))
    return torch.stack(vals)


def reduce_iou_stats(stats: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum over ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize(stats: torch.Tensor) -> Dict[str, float]:
    """The report of trainval_model.py:289-296: precision@X, overall IoU (cum_I / cum_U), mean IoU."""
    s = stats.detach().cpu().tolist()
    n = s[8]
    out = {"cum_I": s[0], "cum_U": s[1], "n": n, "overall_iou": s[0] / s[1] if s[1] else float("nan"),
           "mean_iou": s[2] / n if n else float("nan")}
    for k, th in enumerate(EVAL_SEG_IOU):
        out[f"precision@{th}"] = s[3 + k] / n if n else float("nan")
    return out


# ---------------------------------------------------------------------------------------------------------------------------
# gradient buckets of the training step
# ---------------------------------------------------------------------------------------------------------------------------
def plan_arena(spec: Dict[str, Tuple[int, ...]], bucket_of, buckets, align: int = 64):
    """Lays the buffers of `spec` (name -> shape) out in ONE flat arena, bucket after bucket in the order of `buckets` (the order in
    which the backward pass finishes them), declaration order inside a bucket, every buffer aligned to `align` elements.
    Returns (total elements, {name: (offset, numel)}, {bucket: (start, end)}); the bucket ranges tile [0, total) without gaps."""
    place, ranges, off = {}, {}, 0
    for bname in buckets:
        start = off
        for k, shape in spec.items():
            if bucket_of(k) == bname:
                n = 1
                for e in shape:
                    n *= int(e)
                place[k] = (off, n)
                off += (n + align - 1) // align * align
        ranges[bname] = (start, off)
    missing = set(spec) - set(place)
    if missing:
        raise ValueError(f"buffers outside every bucket: {sorted(missing)}")
    return off, place, ranges


class BucketReducer:
    """Issues one asynchronous all-reduce per finished bucket (`reduce(name)`, in backward order) and waits for all of them
    (`wait()`) before the optimizer reads the arena: the transfer of bucket i runs under the computation of buckets i+1..."""

    def __init__(self, arena: torch.Tensor, ranges: Dict[str, Tuple[int, int]], group: Optional[dist.ProcessGroup] = None):
        self.arena, self.ranges, self.group, self.works = arena, ranges, group, []
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def reduce(self, bname: str):
        a, b = self.ranges[bname]
        if self.world > 1 and b > a:
            self.works.append(dist.all_reduce(self.arena[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []


def gv_sum_allreduce(group: Optional[dist.ProcessGroup] = None):
    """Hook for CMPCHeadB200.gv_allreduce: sums the per-module |gv_lang|^2 (forward) / gv . d gv (backward) over the ranks that hold
    shards of one batch, so that gv_norm='batch' reproduces the reference's axis-less l2_normalize (CMPC_model.py:241) at the GLOBAL batch."""
    def hook(t: torch.Tensor):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return hook
