#!/usr/bin/env python
"""Tuning aid (not part of the product or the tests): times gh_encode alone with CUDA events for every
lib/libgh_b200*.so, WITHOUT checking the output -- so that "probe" builds which deliberately skip a phase of the
kernel (GH_PROBE_* in gh_encode.cu: wrong output, right amount of the remaining work) can be timed to see which
phase the kernel's time hangs on.  usage: python tools/enc_probe.py [workload] [MiB]"""
import glob
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import golden_huffman_b200 as gh  # noqa: E402
import golden_huffman_b200.workloads as W  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "zipf"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n = mib << 20
x = {"zipf": W.zipf_torch, "text": W.text_torch, "uniform": W.uniform_torch, "skewed": W.skewed_torch}[wl](n, torch.device("cuda"))
out = {}
for path in sorted(glob.glob(os.path.join(ROOT, "golden-huffman_b200", "lib", "libgh_b200*.so"))):
    lib = gh.GhLib(path)
    codec = gh.Codec(lib)
    code = codec.build_code(codec.histogram(x))
    payload, _ = codec.encode(x, code)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for _ in range(8):
        ev[0].record()
        codec.encode(x, code, out=payload)
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]))
    out[os.path.basename(path)] = round(sorted(times)[len(times) // 2], 4)
    print(os.path.basename(path), out[os.path.basename(path)], "ms (median of 8, encode + stitch + memset)", flush=True)
print(json.dumps({"workload": wl, "mib": mib, "encode_ms": out}))
