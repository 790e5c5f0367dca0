#!/usr/bin/env python
"""Timing probe for EXPERIMENT builds of the decoder (build.py build_variant with GH_W_EXP_* defines, which may
decode wrongly on purpose -- e.g. stores switched off -- to find what bounds a kernel).  Never a bench number.
usage: GH_LIB_PATH=<variant .so> python tools/dec_probe.py [workload] [MiB]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

import golden_huffman_b200 as gh  # noqa: E402  (bench.py installs the importable alias of the package directory)
import golden_huffman_b200.workloads as W  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "zipf"
n = (int(sys.argv[2]) if len(sys.argv) > 2 else 1024) << 20
dev = torch.device("cuda:0")
base = gh.GhLib(None)
lib = gh.GhLib(os.environ.get("GH_LIB_PATH") or None)
x = bench.make_input(W, wl, n, dev, 0, 1)
stream = torch.cuda.current_stream().cuda_stream
ctx0 = base.ctx_create()
base.ctx_set_stream(ctx0, stream)
img = torch.empty(base.compress_bound(n), dtype=torch.uint8, device=dev)
nbytes, _ = base.compress_device(ctx0, x.data_ptr(), n, img.data_ptr(), img.numel())
ctx = lib.ctx_create()
lib.ctx_set_stream(ctx, stream)
out = torch.empty(n + 64, dtype=torch.uint8, device=dev)
for _ in range(3):
    lib.decompress_device(ctx, img.data_ptr(), nbytes, out.data_ptr(), n)
torch.cuda.synchronize()
ok = bool(torch.equal(out[:n], x))
lib.profile_enable(True)
for _ in range(5):
    lib.decompress_device(ctx, img.data_ptr(), nbytes, out.data_ptr(), n)
torch.cuda.synchronize()
prof = lib.profile_fetch()
lib.profile_enable(False)
print(os.path.basename(os.environ.get("GH_LIB_PATH", "main")), wl, "roundtrip_ok" if ok else "OUTPUT DIFFERS (experiment)",
      "  ".join("%s %.4f" % (name.replace("gh::", "").split("_kernel")[0], ms / max(cnt, 1)) for name, (cnt, ms) in
                sorted(prof.items(), key=lambda kv: -kv[1][1])[:4]))
