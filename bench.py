#!/usr/bin/env python
"""Benchmark of the CMPC head hot path (BASELINE.json: samples/s at 320^2 on 1/2/4/8 B200; graph-reasoning
% of tensor peak vs the TF-CPU-equivalent head).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (sm_100a kernels through the C ABI)
  python bench.py --impl reference [...]                         reference arm: the CPU restatement of the TF-1 head
                                                                 (oracle/; TF itself is not installable in this image)

One step = one forward pass of the head over one batch of synthetic UNC-shaped inputs (configs[1]: batch 32 per GPU,
320x320 -> 40x40 feature maps, N=1600 graph nodes, 20-token expressions, random-init weights).  N > 1 is launched by
torchrun (one rank per GPU); batches are sharded by sample with no data-path collective; the only exchange is the
9-element IoU-statistics all-reduce (trainval_model.py:267-294), so scaling is weak.  Rank 0 prints ONE JSON line.

The same line carries the other BASELINE configs as extra legs (--legs, default all):
  "strong_256"  configs[2]: global batch 256 split 256/N per GPU (strong scaling), samples/s
  "hires_512"   configs[3]: 512x512 (64x64 maps, N = 4096 nodes), batch 16 per GPU: ms/step, graph kernel TF/s (tensor-bound)
                beside the exchange kernel's GB/s (HBM-bound)
  "train"       configs[4]: training step (forward + backward + bucketed gradient all-reduce + Adam), batch 16 per GPU, replayed
                from CUDA graphs; with N > 1 the all-reduce time exposed / hidden under the backward
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "samples/s at 320^2 (CMPC head forward)"
PER_GPU_BATCH = 32
T_WORDS = 20


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def _cpu_oracle_rate(n_samples: int, runs: int, warmup: int):
    """samples/s of the CPU restatement (oracle/) on all host cores, batch 1 per forward like the reference's drivers."""
    import torch
    from oracle.cmpc_head_ref import HeadConfig, OracleHead, init_params, make_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = HeadConfig(batch_size=1)
    params = init_params(cfg, 0)
    head = OracleHead(params, cfg)
    inps = [make_inputs(cfg, 1, seed=1234 + i) for i in range(n_samples)]
    def one(inp):
        with torch.no_grad():
            return head.forward(inp["c3"], inp["c4"], inp["c5"], inp["lstm_outputs"])["pred"]
    for _ in range(warmup):
        one(inps[0])
    times = []
    for _ in range(runs):
        t0 = time.perf_counter()
        for inp in inps:
            one(inp)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return n_samples / med, med, cores


def run_reference(args):
    """Reference arm: TF-1 cannot be installed here (no wheel for cp312, no network) and the reference has no C/C++
    sources to compile, so the arm times the op-for-op CPU restatement (kind 'port'), dense adjacency included."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2
    rate, med, cores = _cpu_oracle_rate(sample, max(1, args.steps), max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "CMPC head forward, 320x320 (N=1600 nodes), 20-token expressions, random init; "
                               f"each step = {sample} samples at batch 1 (bounded sample of configs[1])"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} samples per step, batch 1, PyTorch-CPU fp32 restatement of the TF-1 head (not TF itself)"},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _init_nccl(dist, dev):
    """NCCL on a high-priority stream: the bucketed gradient all-reduces of the train leg run underneath persistent kernels that
    occupy every SM, so their CTAs must win the block scheduler when SMs free up (CMPC_NCCL_PRIO=0 switches it off for A/B runs)."""
    opts = None
    if os.environ.get("CMPC_NCCL_PRIO", "1") != "0":
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            opts = None
    if opts is not None:
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    else:
        dist.init_process_group("nccl", device_id=dev)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cmpc_refseg_b200 import build as _build
    _build.build()
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from cmpc_refseg_b200.synthetic import make_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        _init_nccl(dist, dev)
    B = PER_GPU_BATCH
    model = LSTM_model(batch_size=B, mode="eval", device=dev, seed=0)     # identical weights on every rank (seed 0)
    head = model._head
    # synthetic inputs: seed 1234 + rank; pinned host copies for the e2e leg
    inp = make_inputs(B, seed=1234 + rank)
    host = {k: inp[k].pin_memory() for k in ("c3", "c4", "c5", "lstm_outputs", "target_fine")}
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()

    def step():
        return model.forward(devin["c3"], devin["c4"], devin["c5"], devin["lstm_outputs"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def avg_ms(prof, name):
        ev = prof.get(name, [])
        d = [ev[i].elapsed_time(ev[i + 1]) for i in range(0, len(ev) - 1, 2)]
        return (sum(d) / len(d), len(d)) if d else (None, 0)

    def timed_region(nsteps, run=None):
        """exactly nsteps steps between barrier + synchronize, CUDA events on the launching stream, max over ranks -> ms per step"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if run is None:
            for _ in range(nsteps):
                step()
        else:
            run(nsteps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / nsteps

    # ---------------- timed region of `value`: K steps through the public API with cuda_graph=True -------------------------
    # LSTM_model(cuda_graph=True) replays the pass from a CUDA graph captured on the first call for these input buffers: the ~94
    # kernels of a forward then cost one graph launch instead of 94 dependent stream launches (~2 us of launch gap each).  The
    # CUDA events around the graph / MUTAN kernels are captured INTO the graph as external event-record nodes (head._ev), so the
    # kernel durations below are measured inside this timed region (they hold the last replay's three launches).
    sampler = ClockSampler(local)
    graph_ok, graph_err = 1.0, ""
    try:
        model.cuda_graph = True
        head.prof, head.prof_names = {}, {"graph", "mutan"}
        for _ in range(max(args.warmup, 3)):
            step()
        torch.cuda.synchronize()
    except Exception as e:                                   # never lose the bench line: fall back to the eager pass
        graph_ok, graph_err = 0.0, f"{type(e).__name__}: {str(e)[:200]}"
        model.cuda_graph = False
        head.prof = {}
    t = torch.tensor([graph_ok], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if float(t.item()) < 1.0 and model.cuda_graph:           # another rank failed: every rank measures the eager pass
        model.cuda_graph = False
        head.prof = {}
    if rank == 0:
        sampler.start()
    l0 = head.launches
    ms_step = timed_region(args.steps)
    launches = head.launches - l0
    clocks_main = sampler.stop() if rank == 0 else None
    prof, head.prof = head.prof, None
    value = world * B / (ms_step * 1e-3)
    g_ms, g_n = avg_ms(prof, "graph")
    m_ms, m_n = avg_ms(prof, "mutan")
    used_graph = bool(model.cuda_graph)

    # ---------------- eager pass (one stream launch per kernel), reported beside `value`; its stream-ordered events time the
    # roofline kernels.  It runs SECOND: on a power-capped part whichever leg runs later sees the lower clocks, and `value` is the headline.
    model.cuda_graph = False
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    head.prof, head.prof_names = {}, {"graph", "mutan"}
    l0 = head.launches
    eager_ms = timed_region(args.steps)
    eager_launches = head.launches - l0
    eager_prof, head.prof = head.prof, None
    eg_ms, eg_n = avg_ms(eager_prof, "graph")
    em_ms, _ = avg_ms(eager_prof, "mutan")
    eager = {"value": world * B / (eager_ms * 1e-3), "unit": "samples/s", "ms_per_step": eager_ms, "gpu_launches": eager_launches,
             "graph_kernel_ms": eg_ms, "graph_kernel_launches_timed": eg_n}
    # ---------------- e2e: host buffers in, result out, through the public host-buffer API ----------------
    # every step copies its inputs from pinned host memory (H2D) and its result back (D2H); HostPipeline overlaps the H2D
    # of step i+1 with the kernels of step i on a copy stream (two sets of device input buffers)
    from cmpc_refseg_b200.runner import HostPipeline
    model.cuda_graph = False                                 # PCIe-bound legs: eager launches under the copies
    pipe = HostPipeline(model, fetch="sigm")
    hb = {k: host[k] for k in ("c3", "c4", "c5", "lstm_outputs")}

    def e2e_run(n):
        for _ in pipe.run(hb for _ in range(n)):
            pass
    e2e_run(3)
    e2e_value = world * B / (timed_region(args.steps, e2e_run) * 1e-3)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    # same leg with the three feature maps staged as fp16 in pinned host memory: the head's first op on them is that cast
    # (head._st_lateral), so the results are bit-identical while a step moves half the bytes over PCIe
    hb16 = {k: (host[k].half().pin_memory() if k != "lstm_outputs" else host[k]) for k in ("c3", "c4", "c5", "lstm_outputs")}
    pipe16 = HostPipeline(model, fetch="sigm")

    def e2e16_run(n):
        for _ in pipe16.run(hb16 for _ in range(n)):
            pass
    e2e16_run(3)
    e2e16_value = world * B / (timed_region(args.steps, e2e16_run) * 1e-3)
    graph_replay = ({"value": value, "unit": "samples/s", "ms_per_step": ms_step} if used_graph
                    else {"error": graph_err or "graph capture failed on a rank; `value` is the eager pass"})
    clocks = clocks_main

    # ---------------- IoU reduction over ranks (the one collective of the inference path) ----------------
    from cmpc_refseg_b200.parallel import local_iou_stats, reduce_iou_stats, summarize
    I, U = model.mIoU_counts(devin["target_fine"])
    iou_report = summarize(reduce_iou_stats(local_iou_stats(I, U)))

    if rank == 0:
        peaks, peak_kind = _peaks()
        N, C, L = model.vf_h * model.vf_w, model.v_emb_dim, 1
        f_graph = B * L * (2.0 * N * N * T_WORDS + 2.0 * N * N * C)          # dense algorithmic FLOPs of ONE launch (one level)
        peak_tf = float(peaks.get("bf16_tflops"))                              # burst: the kernel's launches are 0.2 ms inside a 6 ms step at max SM clock
        peak_sus = float(peaks.get("bf16_tflops_sustained", peak_tf))
        roof = None
        # kernel duration: CUDA events recorded on the launching stream around the kernel's launches in the EAGER timed region above
        # (3 launches x K steps).  The events captured into the replayed graph are reported beside it (`launch_ms_in_graph`, the last
        # replay's three launches): an event-record NODE in front of and behind a kernel node adds two node-to-node scheduling
        # latencies to the interval and the gap-free replay runs at lower (power-capped) clocks: +5-8 % on a 0.2 ms kernel.
        k_ms, k_n, k_src = (eg_ms, eg_n, "eager leg") if eg_ms else (g_ms, g_n, "graph replay (in-graph events)")
        if k_ms:
            ach = f_graph / (k_ms * 1e-3) / 1e12
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed `ncu --set full` capture of this kernel
            # (profiles/r02_ncu_graph_summary.json, written by scripts/ncu_summary.py); null when no capture of this build exists
            traffic, tsrc = None, None
            summ = ROOT / "profiles" / "r02_ncu_graph_summary.json"
            if summ.exists():
                try:
                    j = json.loads(summ.read_text())
                    traffic, tsrc = float(j["dram_bytes_per_launch"]), "profiles/r02_ncu_graph_summary.json"
                except Exception:
                    pass
            roof = {"kernel": "graph_reason_kernel", "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf, "frac_sustained": ach / peak_sus,
                    "traffic": traffic, "traffic_unit": "bytes/launch", "traffic_source": tsrc,
                    # X + W + V read once (109 MB) + Y written (105 MB)
                    "algorithmic_bytes": B * N * (2.0 * (C + 24) * 2 + 2 * 32 * 2),
                    "peak_kind": f"{peak_kind} burst cuBLAS bf16 ({peak_tf:.0f}); sustained figure {peak_sus:.0f} in frac_sustained",
                    "launch_ms": k_ms, "launches_timed": k_n, "timed_in": k_src, "launch_ms_in_graph": g_ms,
                    "flops_per_launch": f_graph, "note": "dense F_graph = B*(2N^2 T + 2N^2 C), T=20, C=1000 (SURVEY 8(d)); fp16 operands, fp32 accumulate"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, med, cores = _cpu_oracle_rate(2, 3, 1)
            cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": "2 samples x 3 runs (median), batch 1, PyTorch-CPU fp32 restatement of the TF-1 head incl. dense adjacency"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": "configs[1]: CMPC head forward, batch 32 per GPU, 320x320 (40x40 maps, N=1600 nodes, three levels), "
                                   "20-token expressions, random init", "global_batch": world * B, "parallelism": f"batch-sharded x{world}",
                       "l2": "inputs (734 MB fp32 features per step) exceed the 126 MB L2", "operands": "fp16 x fp16 -> fp32 (TMEM)"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_f16_features": {"value": e2e16_value, "unit": "samples/s", "h2d_bytes_per_step": pipe16.h2d_bytes,
                                 "d2h_bytes_per_step": pipe16.d2h_bytes,
                                 "note": "c3/c4/c5 staged as fp16 on the host (identical results: the head casts them to fp16 first); "
                                         "`e2e` above is PCIe-bound on the fp32 feed"},
            "graph_replay": graph_replay,
            "eager": eager,
            "value_path": ("LSTM_model.forward(cuda_graph=True): the pass replayed from a CUDA graph captured on the first call" if used_graph
                           else "LSTM_model.forward, eager stream launches (graph capture failed)"),
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "kernels": {"graph_reason_ms": eg_ms or g_ms, "mutan_gemm_ms": em_ms or m_ms,
                        "mutan_tflops": (2.0 * B * N * 1008 * 5040 / ((em_ms or m_ms) * 1e-3) / 1e12) if (em_ms or m_ms) else None,
                        "in_graph": {"graph_reason_ms": g_ms, "mutan_gemm_ms": m_ms}},
            "iou": iou_report,
        }
    # ---------------- the other BASELINE configs, same JSON line ----------------
    del pipe, pipe16, model, head, devin, host, hb, hb16, inp
    torch.cuda.empty_cache()
    legs = {}
    want = [] if args.legs == "none" else [x for x in args.legs.split(",") if x] if args.legs != "all" else ["strong_256", "hires_512", "train"]
    for name in want:
        fn = {"strong_256": leg_strong_256, "hires_512": leg_hires_512, "train": leg_train}[name]
        try:
            legs[name] = fn(args, dev, world, rank)
        except Exception as e:                                  # never lose the bench line over an extra leg
            legs[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
        torch.cuda.empty_cache()
    if rank == 0:
        line.update(legs)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _timed(fn, n, dev, world):
    """K calls bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks -> ms per call"""
    import torch
    import torch.distributed as dist
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / n


def _device_inputs(B, dev, seed, *, vf=40, hw=320):
    """synthetic inputs generated ON the device (big batches: 256 samples are 5.9 GB of features): relu(N(0,1)) taps, tanh*sigmoid
    word features, full 20-token sentences"""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, device=dev)
    c3, c4, c5 = torch.relu(r(B, vf, vf, 512)), torch.relu(r(B, vf, vf, 1024)), torch.relu(r(B, vf, vf, 2048))
    lstm = torch.tanh(r(B, T_WORDS, 1000)) * torch.sigmoid(r(B, T_WORDS, 1000))
    target = torch.zeros(B, hw, hw, 1, device=dev)
    target[:, hw // 4:hw // 4 + hw // 3, hw // 5:hw // 5 + hw // 3] = 1.0
    return dict(c3=c3, c4=c4, c5=c5, lstm_outputs=lstm, target_fine=target)


def _avg_ms(prof, name):
    ev = prof.get(name, [])
    d = [ev[i].elapsed_time(ev[i + 1]) for i in range(0, len(ev) - 1, 2)]
    return (sum(d) / len(d), len(d)) if d else (None, 0)


def leg_strong_256(args, dev, world, rank):
    """BASELINE configs[2]: a GLOBAL batch of 256 sharded by sample over the N GPUs (256 / N each), forward, IoU statistics reduced
    over NCCL -- strong scaling (the default `value` is the weak-scaling line at 32 per GPU)."""
    import torch
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from cmpc_refseg_b200.parallel import local_iou_stats, reduce_iou_stats, shard_range, summarize
    lo, hi = shard_range(rank, world, 256)
    B = hi - lo
    model = LSTM_model(batch_size=B, mode="eval", device=dev, seed=0)
    x = _device_inputs(B, dev, 4321 + rank)
    step = lambda: model.forward(x["c3"], x["c4"], x["c5"], x["lstm_outputs"])
    for _ in range(3):
        step()
    n = max(3, min(args.steps, 8))
    ms = _timed(step, n, dev, world)
    I, U = model.mIoU_counts(x["target_fine"])
    rep = summarize(reduce_iou_stats(local_iou_stats(I, U)))
    return {"value": 256 / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "global_batch": 256, "per_gpu_batch": B, "steps": n,
            "scaling": "strong", "iou_samples_reduced": rep["n"],
            "workload": "configs[2]: CMPC head forward, global batch 256 at 320^2 split by sample over the GPUs, IoU reduced over NCCL"}


def leg_hires_512(args, dev, world, rank):
    """BASELINE configs[3]: 512x512 input (64x64 maps, a 4096-node graph), batch 16 per GPU: the tensor-bound dense graph aggregation
    beside the HBM-bound exchange kernel (add3 + l2-normalise, CMPC_model.py:258,272-284)."""
    import torch
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    B, vf, hw = 16, 64, 512
    model = LSTM_model(batch_size=B, mode="eval", device=dev, seed=0, H=hw, W=hw, vf_h=vf, vf_w=vf)
    head = model._head
    x = _device_inputs(B, dev, 777 + rank, vf=vf, hw=hw)
    step = lambda: model.forward(x["c3"], x["c4"], x["c5"], x["lstm_outputs"])
    for _ in range(3):
        step()
    n = max(3, min(args.steps, 10))
    head.prof, head.prof_names = {}, {"graph", "exchange"}
    ms = _timed(step, n, dev, world)
    prof, head.prof = head.prof, None
    g_ms, g_n = _avg_ms(prof, "graph")
    x_ms, x_n = _avg_ms(prof, "exchange")
    peaks, _ = _peaks()
    N, C, M = vf * vf, 1000, 500
    f_graph = B * (2.0 * N * N * T_WORDS + 2.0 * N * N * C)
    x_bytes = B * N * M * 4 * 2                                  # B*N*M*4*s, fp16 storage
    out = {"value": world * B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "per_gpu_batch": B, "nodes": N, "steps": n,
           "workload": "configs[3]: CMPC head forward at 512x512 (64x64 maps, 4096-node graph), batch 16 per GPU"}
    if g_ms:
        tf = f_graph / (g_ms * 1e-3) / 1e12
        out["graph_kernel"] = {"bound": "tensor", "launch_ms": g_ms, "launches_timed": g_n, "flops_per_launch": f_graph, "achieved": tf,
                               "unit": "TFLOP/s", "frac": tf / float(peaks["bf16_tflops"]),
                               "frac_sustained": tf / float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))}
    if x_ms:
        gbs = x_bytes / (x_ms * 1e-3) / 1e9
        out["exchange_kernel"] = {"bound": "hbm", "kernel": "add3_l2norm", "launch_ms": x_ms, "launches_timed": x_n,
                                  "bytes_per_launch": x_bytes, "achieved": gbs, "unit": "GB/s", "frac": gbs / float(peaks["hbm_gbs"])}
    return out


def leg_train(args, dev, world, rank):
    """BASELINE configs[4]: training step of the head, batch 16 per GPU, replayed from CUDA graphs; data parallel over the GPUs with
    the gradient all-reduce issued bucket by bucket underneath the backward (cmpc_refseg_b200/train.py)."""
    import torch
    import torch.distributed as dist
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    B = 16
    model = LSTM_model(batch_size=B, mode="train", device=dev, seed=0)
    from cmpc_refseg_b200.backward import HeadBackward
    rg = {"stage": tuple((b,) for b in HeadBackward.BUCKETS), "flat": (tuple(HeadBackward.BUCKETS),)}.get(os.environ.get("CMPC_REDUCE_GROUPS", ""))
    tr = model.train_op(reduce_groups=rg)                        # default grouping: HeadTrainer.REDUCE_GROUPS (A/B knob for measurements)
    x = _device_inputs(B, dev, 99 + rank)
    step = lambda: tr.train_step(x["c3"], x["c4"], x["c5"], x["lstm_outputs"], x["target_fine"], report_loss=False, graph=True)
    for _ in range(3):
        step()
    n = max(3, min(args.steps, 10))
    l0 = model._head.launches
    ms = _timed(step, n, dev, world)
    launches = (model._head.launches - l0) // n
    out = {"value": world * B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "per_gpu_batch": B, "steps": n, "scaling": "weak",
           "cuda_graph": True, "gpu_launches_per_step": launches, "gv_norm": model.gv_norm,
           "workload": "configs[4]: CMPC head training step (forward + backward + gradient all-reduce + Adam), batch 16 per GPU, "
                       "67.2 M parameters, fp32 master copy, fp16 operands"}
    nbytes = tr.bucket_bytes()
    out["allreduce"] = {"buckets": len(nbytes), "bytes_per_step": sum(nbytes.values()), "largest_bucket_bytes": max(nbytes.values())}
    if world > 1:
        def alone():
            for b in tr.stage_names():
                tr._reduce_async(b)
            tr.reducer.wait()
        for _ in range(2):
            alone()
        ar_ms = _timed(alone, 5, dev, world)
        tr.world_saved, tr.world = tr.world, 1                   # the same step without any all-reduce (parameters drift apart: measured last)
        tr._graph_cache = {}
        for _ in range(2):
            step()
        ms_no = _timed(step, n, dev, world)
        tr.world = tr.world_saved
        exposed = max(0.0, ms - ms_no)
        out["allreduce"].update({"ms_alone": ar_ms, "ms_exposed": exposed, "ms_hidden": max(0.0, ar_ms - exposed),
                                 "ms_per_step_without_allreduce": ms_no,
                                 "busbw_gbs_alone": 2.0 * (world - 1) / world * sum(nbytes.values()) / (ar_ms * 1e-3) / 1e9})
    return out


def run_train(args):
    """BASELINE configs[4]: training step (forward + backward + gradient all-reduce + Adam) of the head, batch 16 per GPU."""
    import torch
    import torch.distributed as dist
    from cmpc_refseg_b200 import build as _build
    _build.build()
    from cmpc_refseg_b200.CMPC_model import LSTM_model
    from cmpc_refseg_b200.synthetic import make_inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        _init_nccl(dist, dev)
    B = 16
    model = LSTM_model(batch_size=B, mode="train", device=dev, seed=0)
    tr = model.train_op()
    head = model._head
    inp = make_inputs(B, seed=1234 + rank)
    keys = ("c3", "c4", "c5", "lstm_outputs", "target_fine")
    host = {k: inp[k].pin_memory() for k in keys}
    devin = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    torch.cuda.synchronize()

    def step(src):
        return tr.train_step(src["c3"], src["c4"], src["c5"], src["lstm_outputs"], src["target_fine"], report_loss=False, graph=args.graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    for _ in range(max(args.warmup, 3)):
        step(devin)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = head.launches
    ms_step = timed(lambda: step(devin), args.steps)
    launches = head.launches - l0
    # e2e: every step copies its batch from pinned host memory and reads the losses back, through runner.TrainPipeline (two staging
    # buffer sets, H2D of step i+1 on a copy stream under the kernels of step i)
    from cmpc_refseg_b200.runner import TrainPipeline
    pipe = TrainPipeline(tr, graph=args.graph)

    def e2e_run(n):
        for _ in pipe.run(host for _ in range(n)):
            pass
    e2e_run(3)
    n_e2e = max(4, args.steps // 2)
    e2e_ms = timed(lambda: e2e_run(n_e2e), 1) / n_e2e
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        h2d = sum(host[k].numel() * host[k].element_size() for k in keys)
        line = {
            "metric": "samples/s at 320^2 (CMPC head training step: forward + backward + all-reduce + Adam)",
            "value": world * B / (ms_step * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "configs[4]: CMPC head training step, batch 16 per GPU, 320x320 (40x40 maps, N=1600), 20-token expressions, "
                                   "random init; 67.2 M parameters, fp32 master copy, fp16 operands",
                       "global_batch": world * B, "parallelism": f"data parallel x{world}, one flat 269 MB gradient all-reduce per step (NCCL)",
                       "l2": "inputs (367 MB fp32 features per step) exceed the 126 MB L2", "cuda_graph": bool(args.graph)},
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32},
            "gpu_launches": launches, "roofline": None, "cpu_baseline": None, "clocks": clocks,
            "loss": tr.last,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="forward", choices=["forward", "train"],
                    help="forward = the headline configs[1] line (default); train = configs[4], the training step")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graph", action="store_true", help="train workload: replay the step from CUDA graphs (HeadTrainer.train_step(graph=True))")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle timing (used under ncu)")
    ap.add_argument("--legs", default="all", help="extra legs in the forward line: all | none | comma list of strong_256,hires_512,train")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
