"""Extracts the judged metrics from an .ncu-rep (run here, no GPU needed):
    python scripts/ncu_summary.py file.ncu-rep                       # text table of every captured launch
    python scripts/ncu_summary.py file.ncu-rep --json out.json --kernel graph_reason
        -> {"kernel", "launches", "time_us", "dram_bytes_per_launch", "tensor_pipe_pct", ...} averaged over the matching launches
           (what bench.py reads for roofline.traffic)"""
import csv, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "smsp__mem_tensor_reads_op_utcmma_matrix_c.sum.pct_of_peak_sustained_elapsed", "smsp__mem_tensor_writes_op_utcmma.sum.pct_of_peak_sustained_elapsed"]
_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def num(v, u):
    return float(v.replace(",", "")) * _SCALE.get(u, 1.0)


args = sys.argv[1:]
rep = args[0]
jout = args[args.index("--json") + 1] if "--json" in args else None
kfilter = args[args.index("--kernel") + 1] if "--kernel" in args else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
sel = []
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    name = d.get("Kernel Name", "?")
    if kfilter and kfilter not in name:
        continue
    sel.append((d, u))
    if jout is None:
        print("kernel:", name[:110])
        for k in KEYS:
            if k in d: print(f"  {k:100s} {d[k]:>16s} {u[k]}")
        try:
            t = num(d["gpu__time_duration.sum"], u["gpu__time_duration.sum"])
            b = num(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + num(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
            print(f"  {'-> DRAM bytes / time':100s} {b / t / 1e3:16.1f} GB/s")
        except Exception:
            pass
if jout is not None:
    if not sel:
        raise SystemExit(f"no launch matches {kfilter!r}")
    avg = lambda k: sum(num(d[k], u[k]) for d, u in sel) / len(sel)
    j = {"kernel": sel[0][0]["Kernel Name"].split("(")[0], "launches": len(sel), "source": rep.split("/")[-1],
         "time_us": avg("gpu__time_duration.sum"),
         "dram_bytes_per_launch": avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum"),
         "dram_bytes_read": avg("dram__bytes_read.sum"), "dram_bytes_write": avg("dram__bytes_write.sum"),
         "tensor_pipe_pct": avg("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
         "note": "ncu --set full --clock-control none; per-launch averages; times are cold-cache / serialised"}
    json.dump(j, open(jout, "w"), indent=1)
    print(json.dumps(j))
