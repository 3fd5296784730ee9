"""Training step of the head: the reference's ``train_op`` (CMPC_model.py:426-478) on the device.

    cost = 0.7 CE(up) + 0.1 CE(up_c5) + 0.1 CE(up_c4) + 0.1 CE(up_c3) + weight_decay * sum_{DW} |w|^2 / 2        (:439-447)
    lr   = polynomial_decay(start_lr, step, lr_decay_step, end 1e-5, power 0.9)                                 (:450-451)
    Adam (TF defaults) on every head variable, the gradients of `biases` doubled                                (:455-478)

One step = training-mode forward (aux heads on) -> HeadBackward.backward_stages, with the all-reduce of every finished gradient
bucket issued asynchronously underneath the remaining stages when there are data-parallel ranks -> fused Adam over three flat
fp32 groups (DW / biases / other) -> re-pack of the fp16 operand copies.  torch is used
for buffers, for re-laying gradients / parameters out between the TF shapes and the packed operand layouts, and for
``torch.distributed.all_reduce`` (NCCL); the arithmetic is in libcmpc_b200.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .head import on_device

from . import _lib as L
from .backward import HeadBackward, Saved
from .weights import pack_head_weights


class HeadTrainer:
    BETA1, BETA2, EPS, END_LR, POWER = 0.9, 0.999, 1e-8, 1e-5, 0.9

    def __init__(self, head, *, start_lr=0.00025, lr_decay_step=800000, weight_decay=0.0005, process_group=None, encoder=None,
                 reduce_groups=None):
        self.h = head
        self.device = head.device
        self.encoder = encoder                     # optional WordEncoderB200: its three variables train with the head's (:426-431)
        self.start_lr, self.lr_decay_step, self.weight_decay = start_lr, lr_decay_step, weight_decay
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (process_group is not None or
                                                                          torch.distributed.is_initialized()) else 1
        dev = head.device
        # flat fp32 master copy in three groups (what differs between them is the Adam call: weight decay on DW, 2x gradient on biases)
        groups = {"dw": [], "bias": [], "other": []}
        source = dict(head.params)
        if encoder is not None:
            source.update(encoder.params)          # `Variable`, `rnn/lstm_cell/kernel`, `rnn/lstm_cell/bias`: no decay, multiplier 1
        for k in sorted(source):
            groups["dw" if k.endswith("/DW") else "bias" if k.endswith("/biases") else "other"].append(k)
        self.layout, off = {}, 0
        self.group_range = {}
        for gname, names in groups.items():
            start = off
            for k in names:
                n = source[k].numel()
                self.layout[k] = (off, n, tuple(source[k].shape))
                off += (n + 3) // 4 * 4
            self.group_range[gname] = (start, off)
        self.theta = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad = torch.zeros_like(self.theta)
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.params: Dict[str, torch.Tensor] = {}
        self.grads: Dict[str, torch.Tensor] = {}
        for k, (o, n, shape) in self.layout.items():
            self.params[k] = self.theta[o:o + n].view(shape)
            self.grads[k] = self.grad[o:o + n].view(shape)
            self.params[k].copy_(source[k].to(dev, torch.float32))
        head.params = {k: v for k, v in self.params.items() if k in head.params}   # the head (and its backward) now pack from the master copy
        if encoder is not None:
            self.enc_names = tuple(encoder.params)
            encoder.pack({k: self.params[k] for k in self.enc_names}, train=True)
        head.saved = Saved(dev)
        self.bw = HeadBackward(head)
        from .parallel import BucketReducer
        self.groups = tuple(tuple(g) for g in (reduce_groups or self.REDUCE_GROUPS))
        flat = [b for g in self.groups for b in g]
        if flat != list(self.bw.BUCKETS):
            raise L.CmpcError("reduce_groups must partition HeadBackward.BUCKETS in order")
        self.group_range_of = {g[-1]: (self.bw.bucket_range[g[0]][0], self.bw.bucket_range[g[-1]][1]) for g in self.groups}
        self.reducer = BucketReducer(self.bw.garena, self.group_range_of, process_group)
        self.step = 0
        self.last: Dict[str, float] = {}
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)      # bias-corrected step size of this step, read by the Adam kernel

    def learning_rate(self, step: Optional[int] = None) -> float:
        s = min(self.step if step is None else step, self.lr_decay_step)
        return (self.start_lr - self.END_LR) * (1.0 - s / self.lr_decay_step) ** self.POWER + self.END_LR

    @on_device
    def repack(self):
        h = self.h
        pack_head_weights(self.params, h.d, h.device, out=h.Wt)        # in place: one strided fp32 -> fp16 copy per operand
        self.bw.pack_weights()
        if self.encoder is not None:
            self.encoder.pack(train=True)

    # The step is a gradient phase and an update phase; neither touches the host, so each can be replayed from CUDA graphs
    # (train_step(..., graph=True)): ~900 launches per step otherwise cost more host time than the GPU needs.
    # Data parallel (world > 1): the gradient phase is cut where a bucket of the backward's gradient arena becomes final
    # (HeadBackward.BUCKETS: ConvLSTM / score first, the language side last) and each bucket's all-reduce is issued right there,
    # asynchronously, so it runs on NCCL's stream underneath the remaining backward stages; only the last bucket is exposed.
    # The packed (padded) gradient buffers are what is reduced -- the re-layout into TF shapes is linear and runs afterwards.
    # Buckets are all-reduced in GROUPS of consecutive backward stages: one NCCL call per group, issued when its last stage is done.
    # Measured on 8 B200: twelve per-stage calls (8-60 MB each) take 1.54 ms on their own against ~0.7 ms for the same 274 MB as one
    # message, and only half of that hides under the backward (NCCL's CTAs get SMs at kernel boundaries only: every SM is held by a
    # persistent kernel otherwise) -- so fewer, larger messages with a small last one.
    REDUCE_GROUPS = (("fuse", "exchange", "c5_graph", "c5_ltrans", "c5_mutan"), ("c4_graph", "c4_ltrans", "c4_mutan"),
                     ("c3_graph", "c3_ltrans"), ("c3_mutan", "language"))

    def stage_names(self):
        """the reduce groups, each named after its last backward stage (+ the word encoder's own group)"""
        return [g[-1] for g in self.groups] + (["encoder"] if self.encoder is not None else [])

    def _grad_stages(self, c3, c4, c5, lstm_outputs, target_fine, seq_len, words):
        """generator: runs forward + losses + the backward up to the next finished gradient bucket, yields the bucket's name"""
        h = self.h
        use_enc = lstm_outputs is None
        if use_enc:
            if self.encoder is None or words is None or seq_len is None:
                raise L.CmpcError("train_step: feed lstm_outputs, or words and seq_len with a word encoder")
            lstm_outputs = self.encoder.forward(words, seq_len, train=True)
        out = self._out = h.forward(c3, c4, c5, lstm_outputs, seq_len, aux=True)
        self.ce = {k: h.ce_sums(out[k], target_fine) for k in ("up", "up_c5", "up_c4", "up_c3")}      # fp64 [B] each, on the device
        last = {g[-1] for g in self.groups}
        for bname in self.bw.backward_stages(out, target_fine):
            if bname in last:
                yield bname                                          # a whole reduce group is final
        if self.encoder is not None:
            if use_enc:
                self.encoder.backward(self.bw.d_lstm, self.grads)   # BPTT through the word LSTM, embedding rows
            else:
                for k in self.enc_names:
                    self.grads[k].zero_()
            yield "encoder"

    def bucket_tensors(self, bname):
        """what a data-parallel step all-reduces for stage `bname`: a slice of the backward's gradient arena, or the encoder's views"""
        if bname == "encoder":
            return [self.grads[k] for k in self.enc_names]
        a, b = self.group_range_of[bname]
        return [self.bw.garena[a:b]]

    def _reduce_async(self, bname):
        """issue the asynchronous all-reduce of a finished bucket (parallel.BucketReducer keeps the handles until wait())"""
        if bname == "encoder":
            self.reducer.works += [torch.distributed.all_reduce(t, group=self.pg, async_op=True) for t in self.bucket_tensors(bname)]
        else:
            self.reducer.reduce(bname)                               # the group's contiguous arena slice, named after its last stage

    def bucket_bytes(self):
        """bytes all-reduced per step, by bucket (reported by bench.py)"""
        return {b: sum(t.numel() * t.element_size() for t in self.bucket_tensors(b)) for b in self.stage_names()}

    def _phase_grad(self, *args):
        """the whole gradient phase; the all-reduces it issued (none on one rank) are outstanding until self.reducer.wait()"""
        for bname in self._grad_stages(*args):
            if self.world > 1:
                self._reduce_async(bname)

    def _phase_update(self):
        h, lib = self.h, self.h.lib
        self.bw.grads_tf(into=self.grads)                           # packed gradient buffers -> the flat TF-shaped views, in place
        scale = 1.0 / self.world                                    # data parallel: mean over the global batch (util/loss.py:12)
        for gname, (a, b) in self.group_range.items():
            if b > a:
                L.check(lib.cmpc_adam_f32(self.theta[a:].data_ptr(), self.grad[a:].data_ptr(), self.m[a:].data_ptr(), self.v[a:].data_ptr(), b - a,
                                          0.0, self.BETA1, self.BETA2, self.EPS, scale * (2.0 if gname == "bias" else 1.0),
                                          self.weight_decay if gname == "dw" else 0.0, self.lr_dev.data_ptr(), h._stream()), "adam")
        self.repack()

    def _graphs_for(self, key, args):
        """capture the phases once per set of input buffers (addresses are baked into the graph); a few sets are kept, so that
        a caller alternating between two staging buffer sets (runner.TrainPipeline) replays instead of re-capturing.
        One rank: the gradient phase is ONE graph.  Data parallel: one graph per gradient bucket (shared memory pool, replayed in
        capture order) so that the all-reduces can be issued between them."""
        cache = self.__dict__.setdefault("_graph_cache", {})
        if key not in cache:
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            dev = self.h.device
            for _ in range(2):                                       # warm-up outside capture (lazy attribute calls, workspace growth)
                self._phase_grad(*args)
                self.reducer.wait()
            torch.cuda.synchronize(dev)
            l0 = self.h.launches
            segs, pool = [], None
            gen = self._grad_stages(*args)
            names = self.stage_names()
            if self.world == 1:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in gen:
                        pass
                segs.append((g, None))
                pool = g.pool()
            else:
                for want in names:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pool):
                        got = next(gen)
                    assert got == want, (got, want)
                    pool = g.pool()
                    segs.append((g, want))
                assert next(gen, None) is None
            out, ce = self._out, self.ce
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb, pool=pool):
                self._phase_update()
            self._graph_launches = self.h.launches - l0              # kernels of this library inside one replay of all graphs
            self.h.launches = l0
            cache[key] = (segs, gb, out, ce)
        return cache[key]

    @on_device
    def train_step(self, c3, c4, c5, lstm_outputs, target_fine, seq_len=None, *, report_loss=True, words=None, graph=False):
        """Either lstm_outputs (the word LSTM then stays outside: its gradient is returned by HeadBackward only) or words + seq_len
        with an encoder, in which case the embedding and the word LSTM are trained too, as in the reference.
        graph=True replays the step from CUDA graphs captured on first use for these input buffers (refill them in place)."""
        args = (c3, c4, c5, lstm_outputs, target_fine, seq_len, words)
        t = self.step + 1
        lr = self.learning_rate(t - 1)
        lr_t = lr * math.sqrt(1.0 - self.BETA2 ** t) / (1.0 - self.BETA1 ** t)
        if graph:
            key = tuple((x.data_ptr(), x.dtype) if torch.is_tensor(x) else None for x in args)
            segs, gb, out, self.ce = self._graphs_for(key, args)
            self.lr_dev.fill_(lr_t)                                  # by value: no host buffer for a later step to overwrite
            for g, bname in segs:
                g.replay()
                if bname is not None:
                    self._reduce_async(bname)                        # NCCL's stream waits for this segment, the next one does not wait for NCCL
            self.reducer.wait()
            gb.replay()
            self.h.launches += self._graph_launches
        else:
            self.lr_dev.fill_(lr_t)
            self._phase_grad(*args)
            out = self._out
            self.reducer.wait()
            self._phase_update()
        self.step = t
        if report_loss:                                             # the only host synchronisation of the step
            ce = {k: float(v.mean()) for k, v in self.ce.items()}
            self.last = dict(cls_loss=ce["up"], cls_loss_c5=ce["up_c5"], cls_loss_c4=ce["up_c4"], cls_loss_c3=ce["up_c3"],
                             cls_loss_all=0.7 * ce["up"] + 0.1 * (ce["up_c5"] + ce["up_c4"] + ce["up_c3"]))
        self.last["learning_rate"] = lr
        return out

    # ---- snapshot / resume (trainval_model.py:135-142 saves one every `snapshot` iterations) ---------------------------------
    def state_dict(self) -> Dict[str, object]:
        return {"step": self.step, "layout": self.layout, "theta": self.theta.detach().cpu(), "m": self.m.detach().cpu(),
                "v": self.v.detach().cpu(), "hparams": (self.start_lr, self.lr_decay_step, self.weight_decay)}

    @on_device
    def load_state_dict(self, sd: Dict[str, object]) -> None:
        if sd["layout"] != self.layout:
            raise L.CmpcError("snapshot was written for a head with different variables / shapes")
        self.step = int(sd["step"])
        for name in ("theta", "m", "v"):
            getattr(self, name).copy_(sd[name])
        self.repack()
