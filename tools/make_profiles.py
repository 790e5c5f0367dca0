#!/usr/bin/env python
"""Turns a record session (tools/gpu_record.sh <tag>, files in gpurun_out/) into the tracked summaries under profiles/:
  <tag>_bench_<workload>.json   the bench lines
  <tag>_launches.csv            ncu launch list (gpu__time_duration per launch; cold-cache, serialised: compare SHARES)
  <tag>_ncu_summary.md          per kernel: key raw metrics of the ncu --set full capture, stall reasons, DRAM traffic,
                                shares of the launch list, dynamic opcode histogram of the hottest code
  traffic.json                  DRAM bytes per launch + the hash of the kernel sources they were captured for
usage: python tools/make_profiles.py <tag> [commit]"""
import csv
import hashlib
import json
import os
import shutil
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def csrc_sha16():
    d = os.path.join(ROOT, "golden-huffman_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h", ".cc")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


out = [f"# ncu summary {tag} (commit {commit}, kernel sources {csrc_sha16()})\n"]
for wl in ("zipf", "uniform", "text", "skewed"):
    f = os.path.join(SRC, f"{tag}_bench_{wl}.json")
    if os.path.exists(f):
        shutil.copy(f, os.path.join(DST, f"{tag}_bench_{wl}.json"))

launches = os.path.join(SRC, f"{tag}_launches.csv")
if os.path.exists(launches):
    shutil.copy(launches, os.path.join(DST, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(launches, errors="ignore")) if len(r) > 5]
    hdr = next((r for r in rows if "Kernel Name" in r), None)
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
    ours = {k: v for k, v in agg.items() if any(s in k for s in ("hist_kernel", "encode", "dec_", "build_code"))}
    tot = sum(v[1] for v in ours.values()) or 1.0
    out.append("## launch list, Zipf 1 GiB (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: compare SHARES)\n")
    out.append("| kernel | launches | total ms | avg ms | share of our kernels |\n|---|---|---|---|---|")
    for k, (c, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {c} | {ms:.3f} | {ms / c:.4f} | {100 * ms / tot:.1f}% |")
    out.append("")

traffic = {}
for sfx, wl in (("z", "zipf"), ("u", "uniform"), ("t", "text")):
    raw = os.path.join(SRC, f"{tag}{sfx}_raw.csv")
    if not os.path.exists(raw):
        continue
    rows = list(csv.reader(open(raw, errors="ignore")))
    hdr, units = rows[0], rows[1]
    out.append(f"## ncu --set full, workload {wl} (per launch; traffic = dram read + write)\n")
    for vals in rows[2:]:
        if len(vals) < len(hdr):
            continue
        name = vals[hdr.index("Kernel Name")].split("(")[0]
        out.append(f"### {name}\n")
        rd = wr = 0.0
        stalls = []
        for i, h in enumerate(hdr):
            if h in KEYS:
                out.append(f"- {h}: {vals[i]} {units[i]}")
            if h == "dram__bytes_read.sum":
                rd = to_bytes(vals[i], units[i])
            if h == "dram__bytes_write.sum":
                wr = to_bytes(vals[i], units[i])
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                stalls.append((float(vals[i] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        out.append("- top stall reasons (warps per issue): " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:6]))
        out.append(f"- DRAM traffic per launch: {(rd + wr) / 1e9:.3f} GB\n")
        if wl == "zipf":
            traffic[name.replace("void ", "").split("<")[0].replace("gh::", "")] = int(rd + wr)
    # dynamic opcode histogram of the kernels in this capture (executed warp instructions by opcode)
    src = os.path.join(SRC, f"{tag}{sfx}_source.csv")
    if os.path.exists(src):
        rows = list(csv.reader(open(src, errors="ignore")))
        cur, per = None, defaultdict(Counter)
        for r in rows:
            if r and r[0] == "Kernel Name":
                cur = r[1].split("(")[0]
                continue
            if cur is None or len(r) < 6 or not r[0].startswith("0x"):
                continue
            t = r[1].strip().split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            try:
                per[cur][op.split(".")[0]] += int(r[5] or 0)
            except ValueError:
                pass
        for k, c in per.items():
            tot = sum(c.values()) or 1
            out.append(f"### executed warp instructions by opcode: {k} ({tot / 1e6:.1f} M)\n")
            out.append(", ".join(f"{op} {100 * n / tot:.1f}%" for op, n in c.most_common(16)) + "\n")

if traffic:
    json.dump({"session": tag, "commit": commit, "csrc_sha16": csrc_sha16(), "workload": "zipf", "size_mib": 1024,
               "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, bench.py --workload zipf",
               "bytes_per_launch": traffic}, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
open(os.path.join(DST, f"{tag}_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}_ncu_summary.md", "traffic:", traffic)
