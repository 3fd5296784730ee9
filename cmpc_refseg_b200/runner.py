"""Host-buffer entry point: runs the head on batches that live in (pinned) HOST memory, the way a driver that owns the
backbone on another device / process would call it (trainval_model.py:232 feeds numpy arrays through feed_dict).

The copies are part of every step: each batch is copied host->device, processed, and its result copied device->host.
Two sets of device input buffers and a dedicated copy stream let the H2D copy of batch i+1 overlap the kernels of batch i
(at batch 32 a step moves 737 MB over PCIe, about twice the kernel time, so without overlap the GPU idles 2/3 of the time).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator

import torch

_IN = ("c3", "c4", "c5", "lstm_outputs")


class HostPipeline:
    def __init__(self, model, fetch: str = "sigm", depth: int = 2):
        self.model, self.fetch, self.depth = model, fetch, depth
        self.dev = model.device
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.out_stream = torch.cuda.Stream(self.dev)
        self.slots = []          # device input buffers, allocated on first use (shapes / dtypes of the host batch)
        self.h2d_bytes = self.d2h_bytes = 0

    def _slot(self, i: int, host: Dict[str, torch.Tensor]):
        while len(self.slots) <= i:
            bufs = {k: torch.empty(host[k].shape, dtype=host[k].dtype, device=self.dev) for k in _IN}
            self.slots.append(dict(bufs=bufs, copied=torch.cuda.Event(), free=torch.cuda.Event(), result=None, dres=None,
                                   done=torch.cuda.Event()))
        return self.slots[i]

    def run(self, batches: Iterable[Dict[str, torch.Tensor]]) -> Iterator[torch.Tensor]:
        """Yields, for every host batch, the fetched output as a pinned host tensor (valid until `depth` more batches have
        been consumed).  Call torch.cuda.synchronize() (or read .done) before touching the last results."""
        main = torch.cuda.current_stream(self.dev)
        it = iter(batches)
        pending = []             # (slot index) in flight on the main stream

        def stage(i, host):
            s = self._slot(i % self.depth, host)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(s["free"])                      # kernels that read these buffers have finished
                for k in _IN:
                    s["bufs"][k].copy_(host[k], non_blocking=True)
                s["copied"].record(self.copy_stream)
            self.h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in _IN)
            return s

        try:
            nxt = stage(0, next(it))
        except StopIteration:
            return
        i = 0
        while nxt is not None:
            cur = nxt
            try:
                nxt = stage(i + 1, next(it))                                # H2D of the next batch overlaps this batch's kernels
            except StopIteration:
                nxt = None
            main.wait_event(cur["copied"])
            b = cur["bufs"]
            out = self.model.forward(b["c3"], b["c4"], b["c5"], b["lstm_outputs"])[self.fetch]
            cur["free"].record(main)
            if cur["result"] is None:
                cur["result"] = torch.empty(out.shape, dtype=out.dtype).pin_memory()
                cur["dres"] = torch.empty_like(out)
            # the head's output buffer is reused by the next forward: park the result in this slot's own device buffer (a 13 MB
            # device copy, microseconds) so that the D2H copy never holds the main stream; the slot's previous D2H finished
            # `depth` batches ago
            main.wait_event(cur["done"])
            cur["dres"].copy_(out, non_blocking=True)
            computed = torch.cuda.Event()
            computed.record(main)
            with torch.cuda.stream(self.out_stream):
                self.out_stream.wait_event(computed)
                cur["result"].copy_(cur["dres"], non_blocking=True)
                cur["done"].record(self.out_stream)
            self.d2h_bytes = out.numel() * out.element_size()
            yield cur["result"]
            i += 1


_TRAIN_IN = ("c3", "c4", "c5", "lstm_outputs", "target_fine")


class TrainPipeline:
    """Training from batches that live in (pinned) HOST memory (trainval_model.py:102-107 feeds numpy arrays through feed_dict):
    each batch is copied host->device into one of two staging buffer sets on a copy stream, so the H2D copy of batch i+1 runs
    under the kernels of step i, and the step's losses are read back (the host waits for them, as `sess.run` does).
    With graph=True the trainer replays one captured CUDA-graph pair per staging set."""

    def __init__(self, trainer, graph: bool = False):
        self.tr, self.graph = trainer, graph
        self.dev = trainer.h.device
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.slots = []
        self.h2d_bytes, self.d2h_bytes = 0, 8 * 4          # four fp64 loss means read back per step

    def _slot(self, i, host):
        while len(self.slots) <= i:
            bufs = {k: torch.empty(host[k].shape, dtype=host[k].dtype, device=self.dev) for k in _TRAIN_IN}
            self.slots.append(dict(bufs=bufs, copied=torch.cuda.Event(), free=torch.cuda.Event()))
        return self.slots[i]

    def run(self, batches: Iterable[Dict[str, torch.Tensor]]) -> Iterator[Dict[str, float]]:
        main = torch.cuda.current_stream(self.dev)
        it = iter(batches)

        def stage(i, host):
            s = self._slot(i % 2, host)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(s["free"])
                for k in _TRAIN_IN:
                    s["bufs"][k].copy_(host[k], non_blocking=True)
                s["copied"].record(self.copy_stream)
            self.h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in _TRAIN_IN)
            return s

        try:
            nxt = stage(0, next(it))
        except StopIteration:
            return
        i = 0
        while nxt is not None:
            cur = nxt
            try:
                nxt = stage(i + 1, next(it))
            except StopIteration:
                nxt = None
            main.wait_event(cur["copied"])
            b = cur["bufs"]
            self.tr.train_step(b["c3"], b["c4"], b["c5"], b["lstm_outputs"], b["target_fine"], report_loss=True, graph=self.graph)
            cur["free"].record(main)
            yield dict(self.tr.last)
            i += 1
