#!/usr/bin/env python
"""Decode time on near-fixed-length codes: bytes uniform over K values (K = 2^L -> 2^L - 1 codes of L bits, 2 of L + 1).
usage: python tools/nearfixed_probe.py [MiB]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import golden_huffman_b200 as gh  # noqa: E402

n = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
dev = torch.device("cuda:0")
lib = gh.GhLib(os.environ.get("GH_LIB_PATH") or None)
codec = gh.Codec(lib)
lib.ctx_set_stream(codec.ctx, torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device=dev)
g.manual_seed(7)
for K in (256, 128, 130, 136, 64, 66, 32, 16, 200):
    x = torch.randint(0, K, (n,), dtype=torch.uint8, device=dev, generator=g)
    img = torch.empty(lib.compress_bound(n), dtype=torch.uint8, device=dev)
    out = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    nb, _ = lib.compress_device(codec.ctx, x.data_ptr(), n, img.data_ptr(), img.numel())
    code, _ = lib.parse_header(img[:2048].cpu().numpy().tobytes())
    for _ in range(2):
        nd, _ = lib.decompress_device(codec.ctx, img.data_ptr(), nb, out.data_ptr(), n)
    assert nd == n and torch.equal(out[:n], x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.profile_enable(True)
    e0.record()
    for _ in range(3):
        lib.decompress_device(codec.ctx, img.data_ptr(), nb, out.data_ptr(), n)
    e1.record()
    torch.cuda.synchronize()
    prof = lib.profile_fetch()
    lib.profile_enable(False)
    top = "  ".join("%s %dx%.3f" % (k.replace("gh::", "").replace("_kernel", ""), c // 3, ms / max(c, 1)) for k, (c, ms) in
                    sorted(prof.items(), key=lambda kv: -kv[1][1])[:4])
    print("K=%3d  lengths %d..%d first_code[min]=%d  decode %.2f ms   %s" % (K, code.min_len, code.max_len, code.first_code[code.min_len],
                                                                         e0.elapsed_time(e1) / 3, top))
