#!/usr/bin/env python
"""Summarise a gpurun session for profiles/: per-kernel shares from the ncu launch list (gpu__time_duration)
and the key raw metrics of an `ncu --set full` report.  usage: summarize_ncu.py <tag> [gpurun_out] [profiles]"""
import csv
import os
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
dst = sys.argv[3] if len(sys.argv) > 3 else "profiles"
os.makedirs(dst, exist_ok=True)
out = [f"# ncu summary {tag}\n"]

launches = os.path.join(src, f"{tag}_launches.csv")
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches, errors="ignore")) if len(r) > 5]
    hdr = next((r for r in rows if "Kernel Name" in r), None)
    if hdr:
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows:
            if r is hdr or len(r) <= vi:
                continue
            try:
                v = float(r[vi].replace(",", ""))
            except ValueError:
                continue
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
            name = r[ki].split("(")[0]
            agg[name][0] += 1
            agg[name][1] += v * scale
        ours = {k: v for k, v in agg.items() if any(s in k for s in ("hist_kernel", "encode", "dec_"))}
        tot = sum(v[1] for v in ours.values()) or 1.0
        out.append("## launch list (ncu --metrics gpu__time_duration.sum, cold-cache serialised: compare SHARES)\n")
        out.append("| kernel | launches | total ms | avg ms | share of our kernels |\n|---|---|---|---|---|")
        for k, (c, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            out.append(f"| {k} | {c} | {ms:.3f} | {ms / c:.4f} | {100 * ms / tot:.1f}% |")
        other = sum(v[1] for k, v in agg.items() if k not in ours)
        out.append(f"\nother kernels (torch data generation / checks): {other:.1f} ms in {sum(v[0] for k, v in agg.items() if k not in ours)} launches\n")

rep = os.path.join(src, f"{tag}_prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
            "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    out.append("## ncu --set full (per launch; traffic = dram read + write)\n")
    seen = defaultdict(int)
    traffic = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        seen[name] += 1
        if seen[name] > 1:
            continue
        out.append(f"### {name}\n")
        try:
            unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * unit_scale[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * unit_scale[units[idx["dram__bytes_write.sum"]]]
            traffic[name.replace("void ", "")] = int(rd + wr)
        except Exception:
            pass
        for w in want:
            if w in idx:
                out.append(f"- {w}: {r[idx[w]]} {units[idx[w]]}")
        stalls = []
        for h, i in idx.items():
            if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h and r[i]:
                try:
                    stalls.append((float(r[i].replace(",", "")), h))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        out.append("- top stall reasons (avg warp latency per issued inst): " +
                   ", ".join(f"{h.split('issue_stalled_')[1].replace('.ratio', '')}={v:.2f}" for v, h in stalls[:6]))
        out.append("")
open(os.path.join(dst, f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
if rep and os.path.exists(rep):
    import json
    json.dump({"source": f"{tag}_prof.ncu-rep (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
               "workload": "1 GiB Zipf(1.1), 1 GPU (bench.py default)", "bytes_per_launch": traffic},
              open(os.path.join(dst, f"traffic_{tag}.json"), "w"), indent=1)
print("\n".join(out))
