#!/usr/bin/env python
"""Tuning aid (not part of the product or the tests): times gh_decode alone (all decode kernels + its host syncs)
with CUDA events for every lib/libgh_b200*.so, WITHOUT checking the output, so that probe builds which skip a part
of a kernel (GH_PROBE_W_* in gh_decode.cu: wrong output) can be timed.  The payload is produced once with the
default library.  usage: python tools/dec_probe.py [workload] [MiB]"""
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
base = gh.Codec(gh.GhLib())
code = base.build_code(base.histogram(x))
payload, end_bit = base.encode(x, code)
torch.cuda.synchronize()
nbytes = (int(end_bit.item()) + 7) // 8
out = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
res = {}
for path in sorted(glob.glob(os.path.join(ROOT, "golden-huffman_b200", "lib", "libgh_b200*.so"))):
    codec = gh.Codec(gh.GhLib(path))
    codec.decode(payload, nbytes, code, n + 4096, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for _ in range(8):
        ev[0].record()
        codec.decode(payload, nbytes, code, n + 4096, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]))
    name = os.path.basename(path)
    res[name] = round(sorted(times)[len(times) // 2], 4)
    print(name, res[name], "ms (median of 8, whole gh_decode)", flush=True)
print(json.dumps({"workload": wl, "mib": mib, "decode_ms": res}))
