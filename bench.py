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
    return ap.parse_args()


def workload_name(args, n_gpus):
    desc = {"zipf": "Zipf(s=1.1) byte stream", "uniform": "uniform random bytes",
            "text": "English-like text (~4.6 bits/symbol)", "skewed": "skewed stream, max code length 32"}[args.workload]
    return f"{args.size_mib} MiB/GPU synthetic {desc}, canonical encode+decode, {n_gpus} GPU(s)"


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
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import importlib.util
    spec = importlib.util.spec_from_file_location("ghw", os.path.join(ROOT, "golden-huffman_b200", "workloads.py"))
    w = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(w)
    sample_mib = 64
    n = sample_mib << 20
    data = w.WORKLOADS_NP[args.workload](n)
    times = []
    kind = None
    for i in range(args.warmup + args.steps):
        kind, (tc, td, size) = cpu_reference_roundtrip(data)
        if i >= args.warmup:
            times.append((tc, td))
    tot = sum(a + b for a, b in times)
    value = n * len(times) / tot / 1e9
    enc = n * len(times) / sum(a for a, _ in times) / 1e9
    dec = n * len(times) / sum(b for _, b in times) / 1e9
    sample = (f"{sample_mib} MiB prefix-equivalent sample of the workload per step; reference compress() + "
              f"TableCanonicalHuffDecoder decompress() via files in /dev/shm; {host_cpu_model()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "sample_mib": sample_mib},
        "encode_GBps": enc, "decode_GBps": dec,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---- the B200 arm ---------------------------------------------------------------------------------------
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

    n = args.size_mib << 20
    x = W.WORKLOADS_TORCH[args.workload](n, dev, seed=W.SEED + rank)
    torch.cuda.synchronize()

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
    for _ in range(max(args.warmup, 3)):
        _, enc_res, out, nsym = step()
    torch.cuda.synchronize()
    if world > 1:
        assert sc.verify_roundtrip(x, out, nsym), "sharded round trip mismatch"
    else:
        assert nsym == n and torch.equal(out[:n], x), "round trip mismatch"
    comp_bytes = int(enc_res) if world == 1 else int(enc_res["payload_bytes"])

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = lib.launch_count()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    evs = []
    for _ in range(args.steps):
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
    value = n_total * args.steps / (total_ms * 1e-3) / 1e9
    enc_gbps = n_total * args.steps / (enc_ms * 1e-3) / 1e9
    dec_gbps = n_total * args.steps / (dec_ms * 1e-3) / 1e9

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
            "dec_write_kernel": C_ + n, "dec_fine_speculate_kernel": C_, "dec_fine_write_kernel": C_ + n,
            "dec_sub_offsets_kernel": 0, "dec_locate_eof_kernel": 0,
            "dec_phase_walk_kernel": C_, "dec_worklist_kernel": 0,  # K5c reads the payload once, like K5a
        }
        def alg_bytes(name):
            return alg.get(name.split("<")[0], 0)

        for name, (cnt, ms) in prof.items():
            avg = ms / cnt
            kernels[name] = {"launches_per_step": cnt / prof_steps, "avg_ms": avg, "ms_per_step": ms / prof_steps,
                             "GBps": (alg_bytes(name) / (avg * 1e-3) / 1e9) if avg > 0 else None}
        traffic_by_kernel = {}
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this same command
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if args.workload == "zipf" and args.size_mib == 1024:
                traffic_by_kernel = tj.get("bytes_per_launch", {})
        except Exception:
            pass
        if prof:
            dom = max(prof, key=lambda k: prof[k][1])
            avg = prof[dom][1] / prof[dom][0]
            ach = alg_bytes(dom) / (avg * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                        "frac": ach / peak_gbs, "traffic": traffic_by_kernel.get(dom), "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg_bytes(dom), "avg_launch_ms": avg}
    stage_roofline = {
        "encode": {"algorithmic_bytes": 2 * n_total + comp_total, "GBps": (2 * n_total + comp_total) * args.steps / (enc_ms * 1e-3) / 1e9},
        "decode": {"algorithmic_bytes": comp_total + n_total, "GBps": (comp_total + n_total) * args.steps / (dec_ms * 1e-3) / 1e9},
    }
    for v in stage_roofline.values():
        v["frac_of_peak"] = v["GBps"] / (peak_gbs * world)

    # ---- e2e: host buffers in and out, copies inside the timed region ------------------------------------
    e2e = None
    if not args.no_e2e:
        h_src = x.cpu().pin_memory()
        h_img = torch.empty(comp_bytes + 4096 if world == 1 else lib.compress_bound(n), dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n + 64, dtype=torch.uint8).pin_memory()
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
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e = {"value": n_total / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": n + nb, "d2h_bytes_per_step": nb + n,
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "bytes_per_gpu": n, "compressed_bytes_total": comp_total,
                       "l2": "inputs (>= 1 GiB per pass) exceed the 126 MB L2; no flush needed" if n >= (256 << 20) else "input smaller than 2x L2",
                       "sharding": "contiguous byte slices, one per rank" if world > 1 else "single GPU"},
            "encode_GBps": enc_gbps, "decode_GBps": dec_gbps, "encode_ms": enc_ms / args.steps, "decode_ms": dec_ms / args.steps,
            "roofline": roofline, "stage_roofline": stage_roofline, "kernels": kernels,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
