#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: canonical Huffman encode/decode GB/s on B200 and % of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload zipf|uniform|text|skewed] [--size-mib M]
  python bench.py --impl reference ...     the reference's own CPU implementation on the box's host cores

A step = one pass of the hot path over one batch: compress (histogram -> code -> header -> pack) of the
resident input into a .crs2 image, then decompress of that image (self-synchronising decode -> scatter).
  value      uncompressed bytes through encode+decode per second, inputs resident in HBM, timed with CUDA
             events on the launching stream (max over ranks); whole-job aggregate over all ranks
  e2e        same metric through gh_compress_host / gh_decompress_host with pinned HOST buffers
             (H2D and D2H copies inside the timed region)
  roofline   the dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs the measured copy peak
  cpu_baseline  the reference (oracle/_ref, compiled unmodified) on one host core, on a bounded sample
One JSON line on stdout (rank 0)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "canonical_huffman_encode+decode_throughput"
UNIT = "GB/s"
GIB = 1 << 30


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="zipf", choices=["zipf", "uniform", "text", "skewed"])
    ap.add_argument("--size-mib", type=int, default=1024, help="uncompressed MiB per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-mib", type=int, default=256)
    ap.add_argument("--no-named-config", action="store_true", help="N > 1: skip the BASELINE config named for this N")
    ap.add_argument("--device-code", type=int, default=-1, help="1/0: build the code on the device / on the host (default: the library's default)")
    ap.add_argument("--named-steps", type=int, default=5)
    return ap.parse_args()


# BASELINE.json configs 4 and 5: what is additionally timed at N = 2 / 4 / 8 ("named_config" in the JSON line)
NAMED = {2: ("text", 4096, "config 4: 4 GiB English-like text sharded across 2 B200"),
         4: ("text", 4096, "config 4: 4 GiB English-like text sharded across 4 B200"),
         8: ("skewed", 16384, "config 5: 16 GiB skewed stream (max code length 32) across 8 B200")}


def workload_name(args, n_gpus=None, workload=None, size_mib=None):
    desc = {"zipf": "Zipf(s=1.1) byte stream", "uniform": "uniform random bytes",
            "text": "English-like text (~4.6 bits/symbol)", "skewed": "skewed stream, max code length 32"}[workload or args.workload]
    return f"{size_mib or args.size_mib} MiB/GPU synthetic {desc}, canonical encode+decode"


# ---- clocks during the timed region ------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, name in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference arm --------------------------------------------------------------------------------
def host_cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_roundtrip(data, repeats=1):
    """times the reference's compress() + decompress() (Table decoder, its fastest) on `data`; one thread
    (the reference has no threading). Falls back to the C port of the oracle if oracle/_ref is not there."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Reference
    best = None
    if Reference.available():
        ref = Reference()
        for _ in range(repeats):
            tc, td, size = ref.time_roundtrip(data, "table")
            if best is None or tc + td < best[0] + best[1]:
                best = (tc, td, size)
        kind = "reference"
    else:
        o = Oracle()
        for _ in range(repeats):
            t0 = time.perf_counter()
            rc, img = o.compress(data)
            t1 = time.perf_counter()
            rc2, back = o.decompress(img, len(data))
            t2 = time.perf_counter()
            assert rc == 0 and rc2 == 0 and len(back) == len(data)
            if best is None or t2 - t0 < best[0] + best[1]:
                best = (t1 - t0, t2 - t1, len(img))
        kind = "port"
    return kind, best


def run_reference(args):
    """the reference's own CPU implementation on the box's host cores, on the GPU arm's config: every step is one
    compress() + decompress() of the whole --size-mib input (1 GiB: about 20 s per step on one core -- the reference
    has no threading, so one core is all it can use)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import importlib.util
    spec = importlib.util.spec_from_file_location("ghw", os.path.join(ROOT, "golden-huffman_b200", "workloads.py"))
    w = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(w)
    n = args.size_mib << 20
    data = w.WORKLOADS_NP[args.workload](n)
    times = []
    kind = None
    size = 0
    for i in range(args.warmup + args.steps):
        kind, (tc, td, size) = cpu_reference_roundtrip(data)
        if i >= args.warmup:
            times.append((tc, td))
    tot = sum(a + b for a, b in times)
    value = n * len(times) / tot / 1e9
    enc = n * len(times) / sum(a for a, _ in times) / 1e9
    dec = n * len(times) / sum(b for _, b in times) / 1e9
    sample = (f"the whole {args.size_mib} MiB input per step (same distribution and size as one GPU's share in the GPU arm); reference compress() + "
              f"TableCanonicalHuffDecoder decompress() via files in /dev/shm; 1 thread; {host_cpu_model()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, 1), "bytes_per_gpu": n, "compressed_bytes_total": size,
                   "note": "the reference is a single-process CPU program: it runs the per-GPU workload once, whatever N"},
        "encode_GBps": enc, "decode_GBps": dec, "host_cpu": host_cpu_model(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def csrc_sha16():
    """identifies the kernel sources a committed ncu capture belongs to"""
    import hashlib
    d = os.path.join(ROOT, "golden-huffman_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h", ".cc")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


# ---- the B200 arm ---------------------------------------------------------------------------------------
def make_input(W, workload, n, dev, rank, world):
    """this rank's slice: the shards are samples of ONE distribution (the Zipf permutation is fixed; only the sampling
    stream depends on the rank), and the skewed input's Fibonacci counts are those of the whole input"""
    if workload == "skewed":
        return W.skewed_torch(n, dev, seed=W.SEED + rank, rank=rank, world=world)
    return W.WORKLOADS_TORCH[workload](n, dev, seed=W.SEED + rank)


def oracle_check_shard(x, enc, rank, sample=1 << 20):
    """N > 1 warm-up check: the first `sample` input bytes of this rank, encoded by the ORACLE with the run's code,
    must be the bits this rank's payload holds from its global start bit on"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, GhCode
    o = Oracle()
    m = min(sample, x.numel())
    raw = x[:m].cpu().numpy().tobytes()
    code = GhCode.from_buffer_copy(bytes(enc["code"]))
    lens = np.array(code.length[:256], dtype=np.int64)
    nbits = int(lens[np.frombuffer(raw, dtype=np.uint8)].sum())
    _, want = o.encode_payload(raw, code)
    want_bits = np.unpackbits(np.frombuffer(want, dtype=np.uint8))[:nbits]
    local_bit = enc["start_bit"] - enc["base_byte"] * 8
    lo = local_bit // 8
    got = enc["payload"][lo: lo + (local_bit % 8 + nbits + 7) // 8 + 1].cpu().numpy()
    got_bits = np.unpackbits(got)[local_bit % 8: local_bit % 8 + nbits]
    return bool((got_bits == want_bits).all())


def measure(args, lib, codec, W, dist, dev, rank, world, workload, n, steps, warmup, want_kernels):
    """compress + decompress of this rank's n bytes, `steps` times; returns the numbers of one JSON block"""
    import torch
    x = make_input(W, workload, n, dev, rank, world)
    torch.cuda.synchronize()
    sc = state = None
    if world > 1:
        from golden_huffman_b200.sharded import ShardedCodec
        sc = ShardedCodec(codec, dist.group.WORLD)
        state = sc.prepare(n)

        def step():
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
            e0.record()
            enc = sc.compress_shard(x, state)
            e1.record()
            out, nsym = sc.decompress_shard(enc, state)
            e2.record()
            return (e0, e1, e2), enc, out, nsym
    else:
        img = torch.empty(lib.compress_bound(n), dtype=torch.uint8, device=dev)
        out_buf = torch.empty(n + 64, dtype=torch.uint8, device=dev)

        def step():
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
            e0.record()
            nbytes, _ = lib.compress_device(codec.ctx, x.data_ptr(), n, img.data_ptr(), img.numel())
            e1.record()
            nsym, _ = lib.decompress_device(codec.ctx, img.data_ptr(), nbytes, out_buf.data_ptr(), n)
            e2.record()
            return (e0, e1, e2), nbytes, out_buf, nsym

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (untimed) + a correctness check of what will be timed
    for _ in range(max(warmup, 3)):
        _, enc_res, out, nsym = step()
    torch.cuda.synchronize()
    checks = {}
    if world > 1:
        assert sc.verify_roundtrip(x, out, nsym), "sharded round trip mismatch"
        ok = oracle_check_shard(x, enc_res, rank)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        assert int(flag.item()) == 1, "a rank's payload bits differ from the oracle's encoding of its input"
        checks = {"roundtrip": "all ranks: decoded slices == inputs", "oracle_bits": "all ranks: first 1 MiB of each shard bit-exact vs the oracle"}
        max_len = int(enc_res["code"].max_len)
    else:
        assert nsym == n and torch.equal(out[:n], x), "round trip mismatch"
        checks = {"roundtrip": "decoded image == input"}
        max_len = int(lib.parse_header(img[:2048].cpu().numpy().tobytes())[0].max_len)
    comp_bytes = int(enc_res) if world == 1 else int(enc_res["payload_bytes"])

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    time.sleep(0.3)
    launches0 = lib.launch_count()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    evs = []
    for _ in range(steps):
        evs.append(step()[0])
    t1.record()
    barrier()
    launches = lib.launch_count() - launches0
    total_ms = t0.elapsed_time(t1)
    enc_ms = sum(a.elapsed_time(b) for a, b, _ in evs)
    dec_ms = sum(b.elapsed_time(c) for _, b, c in evs)
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, enc_ms, dec_ms = t.tolist()
        cb = torch.tensor([comp_bytes], dtype=torch.int64, device=dev)
        dist.all_reduce(cb)
        comp_total = int(cb.item())
    else:
        comp_total = comp_bytes
    n_total = n * world
    res = {
        "x": x, "step": step, "barrier": barrier, "comp_bytes": comp_bytes, "comp_total": comp_total, "n_total": n_total,
        "total_ms": total_ms, "enc_ms": enc_ms, "dec_ms": dec_ms, "launches": int(launches), "clocks": clocks,
        "checks": checks, "max_len": max_len,
        "value": n_total * steps / (total_ms * 1e-3) / 1e9,
        "enc_gbps": n_total * steps / (enc_ms * 1e-3) / 1e9, "dec_gbps": n_total * steps / (dec_ms * 1e-3) / 1e9,
    }
    return res


def stage_fracs(res, steps, peak_gbs, world):
    n_total, comp_total = res["n_total"], res["comp_total"]
    out = {
        "encode": {"algorithmic_bytes": 2 * n_total + comp_total, "GBps": (2 * n_total + comp_total) * steps / (res["enc_ms"] * 1e-3) / 1e9},
        "decode": {"algorithmic_bytes": comp_total + n_total, "GBps": (comp_total + n_total) * steps / (res["dec_ms"] * 1e-3) / 1e9},
    }
    for v in out.values():
        v["frac_of_peak"] = v["GBps"] / (peak_gbs * world)
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    import golden_huffman_b200 as gh
    import golden_huffman_b200.workloads as W
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    lib = gh.GhLib(os.environ.get("GH_LIB_PATH") or None)  # GH_LIB_PATH: a tuning build (build.py build_variant)
    codec = gh.Codec(lib)
    stream = torch.cuda.current_stream()
    lib.ctx_set_stream(codec.ctx, stream.cuda_stream)
    if args.device_code >= 0:
        lib.ctx_set_device_code(codec.ctx, bool(args.device_code))

    n = args.size_mib << 20
    res = measure(args, lib, codec, W, dist, dev, rank, world, args.workload, n, args.steps, args.warmup, True)
    x, step, barrier = res["x"], res["step"], res["barrier"]
    comp_bytes, comp_total, n_total = res["comp_bytes"], res["comp_total"], res["n_total"]
    total_ms, enc_ms, dec_ms = res["total_ms"], res["enc_ms"], res["dec_ms"]
    value, enc_gbps, dec_gbps, launches, clocks = res["value"], res["enc_gbps"], res["dec_gbps"], res["launches"], res["clocks"]
    checks_main = res["checks"]

    # ---- roofline pass: per-kernel CUDA-event durations (separate from the timed steps) --------------------
    roofline = None
    kernels = {}
    if world == 1:
        lib.profile_enable(True)
        prof_steps = 3
        for _ in range(prof_steps):
            step()
        torch.cuda.synchronize()
        prof = lib.profile_fetch()
        lib.profile_enable(False)
        C_ = comp_total
        alg = {  # algorithmic bytes per launch (DESIGN.md "roofline accounting")
            "hist_kernel": n, "encode_kernel": n + C_, "encode_stitch_kernel": 0,
            "dec_build_luts_kernel": 0, "dec_speculate_kernel": C_, "dec_sync_kernel": 0, "dec_tile_sum_kernel": 0, "dec_offsets_kernel": 0,
            "dec_write_kernel": C_ + n, "dec_locate_eof_kernel": 0, "build_code_kernel": 0,
            "dec_phase_walk_kernel": C_, "dec_worklist_kernel": 0,  # K5c reads the payload once, like K5a
        }
        def alg_bytes(name):
            return alg.get(name.split("<")[0], 0)

        for name, (cnt, ms) in prof.items():
            avg = ms / cnt
            kernels[name] = {"launches_per_step": cnt / prof_steps, "avg_ms": avg, "ms_per_step": ms / prof_steps,
                             "GBps": (alg_bytes(name) / (avg * 1e-3) / 1e9) if avg > 0 else None}
        # DRAM bytes per launch: from the committed ncu --set full capture of this same command -- used only while the
        # kernel sources are the ones that were captured (their hash is part of the record)
        traffic_by_kernel, traffic_src = {}, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if args.workload == tj.get("workload", "zipf") and args.size_mib == tj.get("size_mib", 1024):
                if tj.get("csrc_sha16") == csrc_sha16():
                    traffic_by_kernel = tj.get("bytes_per_launch", {})
                    traffic_src = {"session": tj.get("session"), "commit": tj.get("commit"), "csrc_sha16": tj.get("csrc_sha16")}
                else:
                    traffic_src = {"stale": f"profiles/traffic.json was captured for csrc {tj.get('csrc_sha16')}, this is {csrc_sha16()}"}
        except Exception:
            pass
        if prof:
            dom = max(prof, key=lambda k: prof[k][1])
            avg = prof[dom][1] / prof[dom][0]
            ach = alg_bytes(dom) / (avg * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                        "frac": ach / peak_gbs, "traffic": traffic_by_kernel.get(dom.split("<")[0]), "traffic_source": traffic_src,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes(dom), "avg_launch_ms": avg}
    stage_roofline = stage_fracs(res, args.steps, peak_gbs, world)

    # ---- e2e: host buffers in and out, copies inside the timed region ------------------------------------
    e2e = None
    if not args.no_e2e:
        h_src = x.cpu().pin_memory()
        h_img = torch.empty(comp_bytes + 4096 if world == 1 else lib.compress_bound(n), dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
        # what the host link allows: the same bytes as plain pinned copies (all ranks at once), no kernels
        d_tmp = torch.empty(n, dtype=torch.uint8, device=dev)
        barrier()
        c0 = time.perf_counter()
        for _ in range(2):
            d_tmp.copy_(h_src, non_blocking=True)
            h_out[:n].copy_(d_tmp, non_blocking=True)
        torch.cuda.synchronize()
        copy_s = (time.perf_counter() - c0) / 2
        del d_tmp
        e2e_steps = max(1, min(args.steps, 5))
        nb = 0
        for i in range(1 + e2e_steps):
            if i == 1:
                barrier()
                w0 = time.perf_counter()
            nb = codec.compress_host(h_src, h_img)
            nd, _ = codec.decompress_host(h_img, nb, h_out)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        assert nd == n and torch.equal(h_out[:n], h_src)
        e2e_s = (w1 - w0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s, copy_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s, copy_s = t.tolist()
        e2e = {"value": n_total / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": n + nb, "d2h_bytes_per_step": nb + n,
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "host_copy_ceiling": {"GBps_per_direction_all_ranks": n_total / (copy_s / 2) / 1e9, "ms_h2d_plus_d2h_of_n": copy_s * 1e3,
                                     "what": "pinned H2D of the input + D2H of as many bytes, every rank at once, no kernels"},
               "api": "gh_compress_host + gh_decompress_host (pinned host buffers, per rank independent images)"}
        del h_src, h_img, h_out

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        m = min(args.cpu_sample_mib << 20, n)
        sample = x[:m].cpu().numpy()
        kind, (tc, td, size) = cpu_reference_roundtrip(sample)
        cpu_baseline = {"value": m / (tc + td) / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                        "encode_GBps": m / tc / 1e9, "decode_GBps": m / td / 1e9,
                        "sample": f"first {m >> 20} MiB of the same input; reference compress() + TableCanonicalHuffDecoder "
                                  f"decompress(), files in /dev/shm, single thread (the reference has no threading); {host_cpu_model()}"}

    # ---- N > 1: the BASELINE config named for this N, timed beside the weak-scaling workload ----------------
    named = None
    if world in NAMED and not args.no_named_config:
        del x, step, res
        torch.cuda.empty_cache()
        wl, total_mib, what = NAMED[world]
        nn = (total_mib // world) << 20
        r2 = measure(args, lib, codec, W, dist, dev, rank, world, wl, nn, args.named_steps, args.warmup, False)
        if wl == "skewed":
            assert r2["max_len"] == 32, f"config 5 must have a maximum code length of 32, the header says {r2['max_len']}"
        named = {"config": what, "workload": workload_name(args, workload=wl, size_mib=total_mib // world), "bytes_total": r2["n_total"],
                 "compressed_bytes_total": r2["comp_total"], "max_code_length": r2["max_len"], "steps": args.named_steps,
                 "value": r2["value"], "unit": UNIT, "ms_per_step": r2["total_ms"] / args.named_steps,
                 "encode_ms": r2["enc_ms"] / args.named_steps, "decode_ms": r2["dec_ms"] / args.named_steps,
                 "encode_GBps": r2["enc_gbps"], "decode_GBps": r2["dec_gbps"],
                 "stage_roofline": stage_fracs(r2, args.named_steps, peak_gbs, world), "checks": r2["checks"], "clocks": r2["clocks"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args), "bytes_per_gpu": n, "compressed_bytes_total": comp_total,
                       "l2": "inputs (>= 1 GiB per pass) exceed the 126 MB L2; no flush needed" if n >= (256 << 20) else "input smaller than 2x L2",
                       "sharding": "contiguous byte slices of ONE stream, one per rank (same distribution on every rank)" if world > 1 else "single GPU",
                       "checks": res_checks(locals())},
            "encode_GBps": enc_gbps, "decode_GBps": dec_gbps, "encode_ms": enc_ms / args.steps, "decode_ms": dec_ms / args.steps,
            "roofline": roofline, "stage_roofline": stage_roofline, "kernels": kernels,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "named_config": named, "host_cpu": host_cpu_model(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def res_checks(scope):
    return scope.get("checks_main")


if __name__ == "__main__":
    sys.exit(main())
