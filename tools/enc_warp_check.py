#!/usr/bin/env python
"""GH_ENCODE_KERNEL=warp vs the default encoder on the GPU: identical payload bytes, and the time of each
(CUDA events around gh_encode, median of 8).  usage: python tools/enc_warp_check.py [MiB]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import golden_huffman_b200 as gh  # noqa: E402
import golden_huffman_b200.workloads as W  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
codec = gh.Codec(gh.GhLib(os.environ.get("GH_LIB_PATH") or None))
for wl, n in (("zipf", (1 << 20) * mib), ("uniform", (1 << 24) + 12345), ("zipf", 70001), ("zipf", 16384)):
    x = {"zipf": W.zipf_torch, "uniform": W.uniform_torch}[wl](n, torch.device("cuda"))
    code = codec.build_code(codec.histogram(x))
    res = {}
    for which in ("default", "warp"):
        if which == "warp":
            os.environ["GH_ENCODE_KERNEL"] = "warp"
        else:
            os.environ.pop("GH_ENCODE_KERNEL", None)
        payload, end_bit = codec.encode(x, code)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        times = []
        for _ in range(8):
            ev[0].record()
            codec.encode(x, code, out=payload)
            ev[1].record()
            torch.cuda.synchronize()
            times.append(ev[0].elapsed_time(ev[1]))
        nbytes = (int(end_bit.item()) + 7) // 8
        res[which] = (payload[:nbytes].clone(), sorted(times)[4])
    same = res["default"][0].numel() == res["warp"][0].numel() and torch.equal(res["default"][0], res["warp"][0])
    print(wl, n, "identical" if same else "DIFFERENT", "default %.4f ms" % res["default"][1], "warp %.4f ms" % res["warp"][1], flush=True)
os.environ.pop("GH_ENCODE_KERNEL", None)
