"""Host->device copy ceiling of the box, per GPU and aggregate: every rank copies the e2e leg's 736.6 MB of
pinned fp32 features to its GPU repeatedly, no kernels.  Launch with torchrun (one rank per GPU) or alone.
Variants: default affinity / pinned allocation, and the rank's threads bound to the CPUs of the GPU's NUMA node before allocating.
Prints one JSON line on rank 0: aggregate GB/s = what bounds `e2e` at N GPUs whatever the kernels do."""
import json
import os
import sys

import torch
import torch.distributed as dist


def numa_cpus_of_gpu(local):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open(f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node").read())
        if node < 0:
            return None, node
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        return cpus, node
    except Exception as e:
        return None, str(e)[:80]


def measure(dev, nbytes, iters):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(2):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        dst.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes, iters = 736_563_200, 20
    out = {"n_gpus": world, "bytes_per_copy": nbytes, "iters": iters, "cpu_count": os.cpu_count(),
           "affinity_default": len(os.sched_getaffinity(0))}
    res = {}
    res["default"] = measure(dev, nbytes, iters)
    cpus, node = numa_cpus_of_gpu(local)
    out["numa_node_of_gpu0"] = node if rank == 0 else None
    if cpus:
        os.sched_setaffinity(0, cpus)
    res["numa_bound"] = measure(dev, nbytes, iters)
    for k, v in res.items():
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            s, m = t.clone(), t.clone()
            dist.all_reduce(s)
            dist.all_reduce(m, op=dist.ReduceOp.MIN)
            out[k] = {"aggregate_gbs": float(s), "min_per_gpu_gbs": float(m), "mean_per_gpu_gbs": float(s) / world}
        else:
            out[k] = {"aggregate_gbs": v, "min_per_gpu_gbs": v, "mean_per_gpu_gbs": v}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
