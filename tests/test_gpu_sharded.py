"""N > 1 on real GPUs: world_size 2 (or more) over NCCL, one process per GPU. Skipped on a 1-GPU box (the
same orchestration is covered on the CPU by tests/test_sharded_gloo.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, kind, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import golden_huffman_b200 as gh
        import golden_huffman_b200.workloads as W
        from golden_huffman_b200.sharded import ShardedCodec
        from oracle_lib import Oracle
        codec = gh.Codec()
        sc = ShardedCodec(codec)
        full = W.WORKLOADS_NP[kind](n_total, seed=21)
        lo = [0] + [((n_total * (k + 1)) // world) // 16 * 16 for k in range(world - 1)] + [n_total]
        mine = torch.from_numpy(full[lo[rank]:lo[rank + 1]].copy()).cuda()
        state = sc.prepare(max(lo[k + 1] - lo[k] for k in range(world)))
        enc = sc.compress_shard(mine, state)
        torch.cuda.synchronize()
        rc, img = Oracle().compress(full.tobytes())
        assert rc == 0
        hdr = len(img) - enc["total_bytes"]
        if rank == 0:
            assert enc["header"] == img[:hdr]
        a, b = enc["first_byte"], enc["end_byte"]
        got = enc["payload"][a - enc["base_byte"]: b - enc["base_byte"]].cpu().numpy().tobytes()
        assert got == img[hdr + a: hdr + b], f"rank {rank}: payload bytes differ from the reference stream"
        out, nsym = sc.decompress_shard(enc, state)
        torch.cuda.synchronize()
        assert sc.verify_roundtrip(mine, out, nsym)
        off = sc.last_decode["offset"]
        assert out[:nsym].cpu().numpy().tobytes() == full[off:off + nsym].tobytes()
        q.put((rank, "ok", sc.last_decode["rounds"]))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("kind,n_total", [("text", (48 << 20) + 12345), ("uniform", (48 << 20) + 12345),
                                          ("skewed", (48 << 20) + 12345), ("text", 1 << 30)])
def test_sharded_roundtrip_nccl(kind, n_total):
    """every rank's owned payload bytes == the oracle's stream for the whole input (the last case: 1 GiB in total),
    decoded slices == the input"""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=1500) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"
