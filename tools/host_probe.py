#!/usr/bin/env python
"""Wall time of gh_compress_host and gh_decompress_host (pinned buffers) by pipeline chunk size.
usage: python tools/host_probe.py [workload] [MiB]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import golden_huffman_b200 as gh  # noqa: E402
import golden_huffman_b200.workloads as W  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "zipf"
n = (int(sys.argv[2]) if len(sys.argv) > 2 else 1024) << 20
dev = torch.device("cuda:0")
lib = gh.GhLib(os.environ.get("GH_LIB_PATH") or None)
codec = gh.Codec(lib)
x = bench.make_input(W, wl, n, dev, 0, 1)
h_in = x.cpu().pin_memory()
h_img = torch.empty(lib.compress_bound(n), dtype=torch.uint8).pin_memory()
h_out = torch.empty(n + 64, dtype=torch.uint8).pin_memory()


def best(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts) * 1e3


nb = codec.compress_host(h_in, h_img)
print("compress_host %.2f ms  (n %d MiB, image %.1f MiB)" % (best(lambda: codec.compress_host(h_in, h_img)), n >> 20, nb / 2**20))
for chunk_mib in (0, 8, 16, 32, 64, 128, 4096):
    lib.ctx_set_host_chunk(codec.ctx, chunk_mib << 20)
    ms = best(lambda: codec.decompress_host(h_img, nb, h_out))
    assert torch.equal(h_out[:n], h_in)
    print("decompress_host chunk %4d MiB: %.2f ms" % (chunk_mib, ms))
# plain copies for scale
d = torch.empty(n, dtype=torch.uint8, device=dev)
print("H2D n: %.2f ms   D2H n: %.2f ms" % (best(lambda: (d.copy_(h_in, non_blocking=True), torch.cuda.synchronize())),
                                             best(lambda: (h_out[:n].copy_(d, non_blocking=True), torch.cuda.synchronize()))))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(n, dtype=torch.uint8, device=dev)


def duplex():
    with torch.cuda.stream(s1):
        d.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out[:n].copy_(d2, non_blocking=True)
    torch.cuda.synchronize()


print("H2D n and D2H n at once: %.2f ms" % best(duplex))
