#!/usr/bin/env python
"""Summaries of the ncu captures brought back in gpurun_out/ (run here, on the CPU box).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv [--last N] > profiles/rNN_ncu_launches_summary.md
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep               > profiles/rNN_ncu_full_summary.md

`launches`: the `--metrics gpu__time_duration.sum` launch list -> per-kernel totals and shares of the last N launches
(default: the last training step, found as the launches after the last adamw_kernel but one).
`full`: one `--set full` report -> per-launch duration, DRAM bytes read / written, DRAM %, tensor-pipe %, grid, registers.
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = name.replace("void ", "").replace("osb::<unnamed>::", "").replace("(anonymous namespace)::", "")
    return re.sub(r"\(.*$", "", name)[:90]


def launches(path: str, last: int | None):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3, r["Grid Size"]))
    if last is None:  # one optimizer step = the launches between the last two adamw_kernel launches
        idx = [i for i, r in enumerate(rows) if "adamw_kernel" in r[0]]
        rows = rows[idx[-2] + 1: idx[-1] + 1] if len(idx) >= 2 else rows
    else:
        rows = rows[-last:]
    agg = OrderedDict()
    for name, us, _ in rows:
        a = agg.setdefault(short(name), [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    lib = sum(a[1] for k, a in agg.items() if not k.startswith("at::") and "nccl" not in k and "cublas" not in k.lower() and "gemv" not in k)
    print(f"kernels in window: {len(rows)}; total device time {total / 1e3:.3f} ms; libosb200 kernels {lib / 1e3:.3f} ms "
          f"({100 * lib / total:.1f}% of device time)\n")
    print("Per-launch times are cold-cache and serialised (ncu): compare SHARES, not absolutes.\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f}% | {us / n:.1f} |")


def full(path: str):
    if path.endswith(".csv"):   # already exported on the GPU box (`ncu -i rep --page raw --csv`): the report itself was too big to travel
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h = r[0]
    idx = {n: i for i, n in enumerate(h)}
    cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("launch__grid_size", "grid"),
            ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]
    cols = [(k, t) for k, t in cols if k in idx]
    print("| kernel | " + " | ".join(t for _, t in cols) + " |\n|---|" + "---:|" * len(cols))
    for row in r[2:]:
        vals = []
        for k, _ in cols:
            v = row[idx[k]]
            try:
                vals.append(f"{float(v):.2f}" if "." in v else v)
            except ValueError:
                vals.append(v)
        print(f"| `{short(row[idx['Kernel Name']])}` | " + " | ".join(vals) + " |")
    if "--json" in sys.argv:  # per C-ABI entry point: mean DRAM bytes (read + write) per launch -> bench.py roofline.traffic
        import json

        abi = {"gemm_nt_kernel": "osb_gemm", "convnext_fused_kernel": "osb_convnext_block_fwd", "gemm_wgrad_kernel": "osb_gemm_wgrad",
               "convnext_bwd_fused_kernel": "osb_convnext_block_bwd", "mha_fwd_kernel": "osb_mha_fwd", "mha_bwd_kernel": "osb_mha_bwd"}
        acc = {}
        for row in r[2:]:
            name = row[idx["Kernel Name"]]
            for k, v in abi.items():
                if k in name:
                    mb = float(row[idx["dram__bytes_read.sum"]]) + float(row[idx["dram__bytes_write.sum"]])
                    unit = r[1][idx["dram__bytes_read.sum"]]
                    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
                    a = acc.setdefault(v, [0, 0.0])
                    a[0] += 1
                    a[1] += mb * scale
        out_path = sys.argv[sys.argv.index("--json") + 1]
        json.dump({k: {"launches_captured": n, "dram_bytes_per_launch": b / n, "source": path} for k, (n, b) in acc.items()},
                  open(out_path, "w"), indent=1)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "launches":
        last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else None
        launches(path, last)
    else:
        full(path)
