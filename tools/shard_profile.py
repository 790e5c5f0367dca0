#!/usr/bin/env python
"""Where a sharded step spends its host time: cProfile of rank 0 over 30 compress_shard + decompress_shard steps.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/shard_profile.py"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import golden_huffman_b200 as gh  # noqa: E402
import golden_huffman_b200.workloads as W  # noqa: E402
from golden_huffman_b200.sharded import ShardedCodec  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = gh.GhLib(None)
codec = gh.Codec(lib)
lib.ctx_set_stream(codec.ctx, torch.cuda.current_stream().cuda_stream)
n = 1 << 30
x = bench.make_input(W, "zipf", n, dev, rank, world)
sc = ShardedCodec(codec, dist.group.WORLD)
st = sc.prepare(n)
for _ in range(5):
    enc = sc.compress_shard(x, st)
    out, nsym = sc.decompress_shard(enc, st)
torch.cuda.synchronize()
dist.barrier()
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    enc = sc.compress_shard(x, st)
    out, nsym = sc.decompress_shard(enc, st)
torch.cuda.synchronize()
pr.disable()
if rank == 0:
    ps = pstats.Stats(pr)
    ps.sort_stats("tottime").print_stats(22)
dist.destroy_process_group()
