"""Parity tests proper (B200): the CUDA path, called through the C ABI, against the oracle and the committed
golden vectors -- bit-exact (all arithmetic on this path is integer/byte work)."""
import hashlib
import json
import os

import numpy as np
import pytest

from golden_cases import CASES, make_input

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
META = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def _cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _bytes(t):
    return t.cpu().numpy().tobytes()


# ---- K1 histogram -------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 15, 16, 17, 255, 4096, 100003, (1 << 22) + 5])
def test_histogram_sizes(codec, oracle, n):
    rng = np.random.default_rng(n)
    d = rng.integers(0, 256, n, dtype=np.uint8)
    h = codec.histogram(_cuda(d)).cpu().numpy().astype(np.uint64)
    assert (h == oracle.histogram(d.tobytes())).all()


def test_histogram_misaligned_and_skewed(codec, oracle):
    import torch
    rng = np.random.default_rng(0)
    base = np.zeros(1 << 20, dtype=np.uint8)
    base[rng.integers(0, base.size, 1000)] = rng.integers(1, 256, 1000)  # one value ~99.9 %
    t = _cuda(base)
    for off in (0, 1, 7, 13, 16, 31):
        view = t[off:off + 700001]
        h = torch.empty(256, dtype=torch.int64, device="cuda")
        codec.lib.histogram(view.data_ptr(), view.numel(), h.data_ptr(), False, torch.cuda.current_stream().cuda_stream)
        assert (h.cpu().numpy().astype(np.uint64) == oracle.histogram(base[off:off + 700001].tobytes())).all()
    # accumulate across two calls == one call
    h2 = codec.histogram(t[: 1 << 19])
    codec.histogram(t[1 << 19:], out=h2, accumulate=True)
    assert (h2.cpu().numpy().astype(np.uint64) == oracle.histogram(base.tobytes())).all()


# ---- whole images against the golden vectors ------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_images(codec, name):
    data = make_input(name)
    m = META[name]
    x = _cuda(np.frombuffer(data, dtype=np.uint8))
    img = codec.compress(x)
    blob = _bytes(img)
    assert len(blob) == m["crs2_bytes"]
    assert hashlib.sha256(blob).hexdigest() == m["sha256_crs2"]
    small = os.path.join(ROOT, "tests", "golden", name + ".crs2")
    if os.path.exists(small):
        assert blob == open(small, "rb").read()
    out, n, rc = codec.decompress(img, len(data) + 32)
    assert rc == 0 and n == len(data)
    assert _bytes(out) == data


def _gen(kind, n, rng):
    if kind == 0:
        return rng.integers(0, 256, n, dtype=np.uint8)
    if kind == 1:
        return rng.integers(0, int(rng.integers(1, 8)), n, dtype=np.uint8)
    if kind == 2:
        return np.minimum(rng.geometric(0.3, n), 255).astype(np.uint8)
    if kind == 3:
        return (rng.zipf(1.3, n) % 256).astype(np.uint8)
    return np.repeat(rng.integers(0, 256, 8, dtype=np.uint8), rng.integers(1, max(2, n // 4), 8))[: max(1, n)]


def test_random_inputs_vs_oracle(codec, oracle):
    rng = np.random.default_rng(2024)
    sizes = [1, 2, 3, 15, 16, 17, 31, 32, 33, 127, 4095, 4096, 4097, 8191, 12289, 65536, 100001, 1 << 20, 3 * (1 << 20) + 11]
    for trial in range(60):
        n = int(sizes[trial % len(sizes)] if trial < 2 * len(sizes) else rng.integers(1, 1 << 21))
        d = _gen(trial % 5, n, rng)
        data = d.tobytes()
        rc, img = oracle.compress(data)
        assert rc == 0
        x = _cuda(d)
        got = codec.compress(x)
        assert _bytes(got) == img, (trial, n)
        out, nd, rc = codec.decompress(got, len(data) + 7)
        assert rc == 0 and nd == len(data) and _bytes(out) == data, (trial, n)


@pytest.mark.parametrize("workload", ["zipf", "uniform", "text", "skewed"])
def test_baseline_workloads_64mib_vs_oracle(codec, oracle, workload):
    """the four BASELINE.json input shapes at a size the oracle finishes in seconds"""
    import torch
    import golden_huffman_b200.workloads as w
    n = 1 << 26
    x = w.WORKLOADS_TORCH[workload](n, "cuda")
    data = x.cpu().numpy()
    hist = codec.histogram(x).cpu().numpy().astype(np.uint64)
    assert (hist == oracle.histogram(data)).all()
    code = codec.build_code(hist)
    if workload == "skewed":
        assert code.max_len == 32 and code.min_len == 1  # the long-codeword stress really is one
    if workload == "uniform":
        assert code.max_len <= 9
    rc, img = oracle.compress(data)
    assert rc == 0
    got = codec.compress(x)
    assert got.numel() == len(img)
    assert torch.equal(got.cpu(), torch.from_numpy(np.frombuffer(img, dtype=np.uint8).copy()))
    out, nd, rc = codec.decompress(got, n + 64)
    assert rc == 0 and nd == n
    assert torch.equal(out, x)


def test_step_by_step_abi(codec, oracle):
    """the individual entry points (histogram -> build_code -> header -> encode -> decode), not the image wrappers"""
    import torch
    data = make_input("skew_geometric_200k")
    x = _cuda(np.frombuffer(data, dtype=np.uint8))
    hist = codec.histogram(x)
    code = codec.build_code(hist)
    rc, ocode = oracle.build_code(oracle.histogram(data))
    assert codec.lib.write_header(code) == oracle.write_header(ocode)
    payload, end_bit = codec.encode(x, code)
    torch.cuda.synchronize()
    bits = int(end_bit.item())
    assert bits == oracle.payload_bits(ocode, oracle.histogram(data))
    rc, opayload = oracle.encode_payload(data, ocode)
    nbytes = (bits + 7) // 8
    assert _bytes(payload[:nbytes]) == opayload
    out, n, rc = codec.decode(payload, nbytes, code, len(data))
    assert rc == 0 and n == len(data) and _bytes(out[:n]) == data


def test_encode_start_bit_contract(codec, oracle):
    import torch
    data = make_input("lcg_uniform_100k")
    x = _cuda(np.frombuffer(data, dtype=np.uint8))
    rc, ocode = oracle.build_code(oracle.histogram(data))
    code = codec.build_code(oracle.histogram(data))
    _, full = oracle.encode_payload(data, ocode)
    bits = np.unpackbits(np.frombuffer(full, dtype=np.uint8))
    total = oracle.payload_bits(ocode, oracle.histogram(data)) - ocode.length[256]
    for start_bit in (0, 1, 31, 32, 64, 100, 127):
        cap = codec.lib.encode_payload_capacity(len(data), code, start_bit)
        out = torch.full((cap,), 0xAA, dtype=torch.uint8, device="cuda")
        payload, end_bit = codec.encode(x, code, start_bit=start_bit, append_eof=False, out=out)
        torch.cuda.synchronize()
        assert int(end_bit.item()) == start_bit + total
        host = payload.cpu().numpy()
        first_word = start_bit // 32 * 4
        assert (host[:first_word] == 0xAA).all()
        got = np.unpackbits(host[first_word:])
        lead = start_bit - first_word * 8
        assert not got[:lead].any()
        assert (got[lead:lead + total] == bits[:total]).all()


def test_sharded_slices_on_one_gpu(codec, oracle):
    """the multi-GPU decomposition run as slices on one device: per-slice encode at the global bit phase,
    OR-stitch of the boundary bytes, then two-slice decode with exit_bit exchange"""
    import torch
    import golden_huffman_b200.workloads as w
    n = 6 * (1 << 20) + 12345
    data = w.text_np(n, seed=9)
    x = _cuda(data)
    hist = oracle.histogram(data.tobytes())
    code = codec.build_code(hist)
    rc, img = oracle.compress(data.tobytes())
    hdr = codec.lib.header_bytes(code)
    ref_payload = np.frombuffer(img[hdr:], dtype=np.uint8)
    world = 4
    bounds = [n * r // world // 16 * 16 for r in range(world)] + [n]
    stream = np.zeros(ref_payload.size + 64, dtype=np.uint8)
    start = 0
    for r in range(world):
        xs = x[bounds[r]:bounds[r + 1]]
        h = codec.histogram(xs).cpu().numpy().astype(np.uint64)
        bits = codec.lib.payload_bits(code, h, with_eof=(r == world - 1))
        phase = start % 128
        payload, end_bit = codec.encode(xs, code, start_bit=phase, append_eof=(r == world - 1))
        torch.cuda.synchronize()
        assert int(end_bit.item()) == phase + bits
        nbytes = (phase + bits + 7) // 8
        base = (start - phase) // 8
        lo = phase // 8  # bytes before the one holding start_bit are not this shard's (and are not written)
        stream[base + lo:base + nbytes] |= payload[lo:nbytes].cpu().numpy()  # boundary byte: OR of the two shards
        start += bits
    assert (stream[:ref_payload.size] == ref_payload).all()

    # decode in 3 slices
    pay = _cuda(np.concatenate([ref_payload, np.zeros(32, dtype=np.uint8)]))
    total = ref_payload.size
    cuts = [0, total // 3 // 16 * 16, 2 * total // 3 // 16 * 16, total]
    ws = [torch.empty(codec.lib.decode_workspace_bytes(cuts[k + 1] - cuts[k]) + 256, dtype=torch.uint8, device="cuda")
          for k in range(3)]
    st = torch.cuda.current_stream().cuda_stream
    res = []
    for k in range(3):
        nb = cuts[k + 1] - cuts[k]
        readable = min(nb + 8, total - cuts[k])
        res.append(codec.lib.decode_sync(pay.data_ptr() + cuts[k], nb, readable, code, 0, True, ws[k].data_ptr(), ws[k].numel(), st))
    for k in range(1, 3):  # one exchange round suffices unless a slice's exit changes
        nb = cuts[k + 1] - cuts[k]
        readable = min(nb + 8, total - cuts[k])
        res[k] = codec.lib.decode_sync(pay.data_ptr() + cuts[k], nb, readable, code, res[k - 1].exit_bit, False,
                                       ws[k].data_ptr(), ws[k].numel(), st)
    assert res[2].eof_found and sum(r.n_symbols for r in res) == n
    out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    o = 0
    for k in range(3):
        nb = cuts[k + 1] - cuts[k]
        readable = min(nb + 8, total - cuts[k])
        codec.lib.decode_write(pay.data_ptr() + cuts[k], nb, readable, code, out.data_ptr() + o, res[k].n_symbols,
                               ws[k].data_ptr(), ws[k].numel(), st)
        o += res[k].n_symbols
    torch.cuda.synchronize()
    assert torch.equal(out[:n], x)


def test_host_buffer_api(codec, oracle):
    """gh_compress_host / gh_decompress_host: pinned host buffers in and out (the e2e path of bench.py)"""
    import torch
    import golden_huffman_b200.workloads as w
    n = (1 << 23) + 3
    data = w.zipf_np(n, seed=5)
    src = torch.from_numpy(data).pin_memory()
    dst = torch.empty(codec.lib.compress_bound(n), dtype=torch.uint8).pin_memory()
    nb = codec.compress_host(src, dst)
    rc, img = oracle.compress(data.tobytes())
    assert dst[:nb].numpy().tobytes() == img
    back = torch.empty(n + 8, dtype=torch.uint8).pin_memory()
    nd, rc = codec.decompress_host(dst, nb, back)
    assert rc == 0 and nd == n and torch.equal(back[:n], src)


def test_host_decompress_pipeline(codec):
    """gh_decompress_host's chunked pipeline (upload, decode and read-back of different chunks at once) at a size with
    several default chunks, and with small chunks; pinned and pageable host buffers"""
    import torch
    import golden_huffman_b200.workloads as w
    n = (200 << 20) + 12345
    x = w.zipf_torch(n, "cuda", seed=21)
    img = codec.compress(x)
    nb = img.numel()
    src = img.cpu().pin_memory()
    want = x.cpu()
    try:
        for chunk, pinned in ((0, True), (1 << 20, True), (0, False), (5 << 20, False)):
            codec.lib.ctx_set_host_chunk(codec.ctx, chunk)
            back = torch.empty(n + 8, dtype=torch.uint8)
            back = back.pin_memory() if pinned else back
            s = src if pinned else src.clone()
            nd, rc = codec.decompress_host(s, nb, back)
            assert rc == 0 and nd == n and torch.equal(back[:n], want), (chunk, pinned)
    finally:
        codec.lib.ctx_set_host_chunk(codec.ctx, 0)


def test_error_statuses(codec, oracle):
    import torch
    import golden_huffman_b200 as gh
    data = make_input("text_small") * 50
    x = _cuda(np.frombuffer(data, dtype=np.uint8))
    img = codec.compress(x)
    out, n, rc = codec.decompress(img, len(data) - 100, allow=(gh.capi.GH_ERR_SPACE,))
    assert rc == gh.capi.GH_ERR_SPACE and n == len(data)
    assert _bytes(out) == data[: len(data) - 100]
    trunc = img[: img.numel() - 64].clone()
    out, n, rc = codec.decompress(trunc, len(data) + 64, allow=(gh.capi.GH_ERR_NO_EOF,))
    assert rc == gh.capi.GH_ERR_NO_EOF
    bad = img.clone()
    bad[:4] = 0
    with pytest.raises(gh.GhError) as e:
        codec.decompress(bad, len(data))
    assert e.value.status == gh.capi.GH_ERR_FORMAT
    with pytest.raises(gh.GhError) as e:
        codec.compress(torch.empty(0, dtype=torch.uint8, device="cuda"))
    assert e.value.status == gh.capi.GH_ERR_EMPTY


def test_decode_stops_at_first_end_mark(codec, oracle):
    """garbage after the end mark is ignored exactly like the reference's decoders ignore it"""
    import torch
    data = make_input("text_small") * 9
    rc, img = oracle.compress(data)
    rng = np.random.default_rng(1)
    blob = np.concatenate([np.frombuffer(img, dtype=np.uint8), rng.integers(0, 256, 5000, dtype=np.uint8)])
    rc, want = oracle.decompress(blob, len(data) + 10000)
    assert rc == 0 and want == data
    out, n, rc = codec.decompress(_cuda(blob), len(data) + 10000)
    assert rc == 0 and n == len(data) and _bytes(out) == data


@pytest.mark.parametrize("workload", ["zipf", "uniform"])
def test_full_size_1gib(codec, oracle, workload):
    """BASELINE.json configs 2 and 3 at full size: compressed image identical to the oracle's (sha256 of both),
    size equal to the closed form from the histogram, and decode(encode(x)) == x compared on the device"""
    import torch
    import golden_huffman_b200.workloads as w
    n = 1 << 30
    x = w.WORKLOADS_TORCH[workload](n, "cuda")
    hist = codec.histogram(x).cpu().numpy().astype(np.uint64)
    assert int(hist.sum()) == n
    code = codec.build_code(hist)
    img = codec.compress(x)
    hdr = codec.lib.header_bytes(code)
    assert img.numel() == hdr + (codec.lib.payload_bits(code, hist) + 7) // 8
    out, nd, rc = codec.decompress(img, n)
    assert rc == 0 and nd == n
    assert torch.equal(out, x)
    del out
    host = x.cpu().numpy()
    rc, oimg = oracle.compress(host)
    assert rc == 0 and len(oimg) == img.numel()
    assert hashlib.sha256(oimg).digest() == hashlib.sha256(img.cpu().numpy().tobytes()).digest()


@pytest.mark.parametrize("mode", ["coarse", "no_phase_walk"])
def test_forced_decode_pipelines(codec, oracle, mode):
    """the library's test hooks: the thread-per-subsequence pipeline forced, and the phase walk of 8/9-bit codes switched
    off (uniform bytes then take the general re-walk rounds): identical output on fast- and slow-synchronising codes"""
    import torch
    import golden_huffman_b200.workloads as w
    pipeline = mode
    codec.lib.lib.gh_debug_select_writer(1 if mode == "coarse" else 0)
    codec.lib.lib.gh_debug_disable_phase_walk(1 if mode == "no_phase_walk" else 0)
    try:
        for name in ("zipf", "uniform", "skewed", "text"):
            n = (1 << 23) + 77
            x = w.WORKLOADS_TORCH[name](n, "cuda", seed=3)
            img = codec.compress(x)
            out, nd, rc = codec.decompress(img, n + 5)
            assert rc == 0 and nd == n and torch.equal(out, x), (pipeline, name)
        data = make_input("fib24_shuffled")
        x = _cuda(np.frombuffer(data, dtype=np.uint8))
        out, nd, rc = codec.decompress(codec.compress(x), len(data))
        assert rc == 0 and nd == len(data) and _bytes(out) == data
    finally:
        codec.lib.lib.gh_debug_select_writer(0)
        codec.lib.lib.gh_debug_disable_phase_walk(0)


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [0, 16])
@pytest.mark.parametrize("name", ["text", "uniform", "kat3"])
def test_decode_payload_alignment(codec, oracle, name, shift):
    """gh_decode takes a payload that is only 16-byte aligned (two 128-bit loads per unit) as well as a 32-byte
    aligned one (one 256-bit load); `uniform` / `kat3` go through the phase walk of the 8/9-bit code, `text`
    through speculation + synchronisation rounds. Output identical to the input either way."""
    import torch
    import golden_huffman_b200.workloads as w
    if name == "kat3":
        data = make_input("kat3_allbytes512")
        x = _cuda(np.frombuffer(data, dtype=np.uint8))
    else:
        x = w.WORKLOADS_TORCH[name]((1 << 22) + 1234, "cuda", seed=5)
        data = _bytes(x)
    code = codec.build_code(codec.histogram(x))
    payload, end_bit = codec.encode(x, code)
    torch.cuda.synchronize()
    nbytes = (int(end_bit.item()) + 7) // 8
    rc, opayload = oracle.encode_payload(data, oracle.build_code(oracle.histogram(data))[1])
    assert _bytes(payload[:nbytes]) == opayload
    buf = torch.empty(nbytes + 64 + 256, dtype=torch.uint8, device="cuda")
    off = (-buf.data_ptr()) % 32 + shift  # 32-byte aligned, or exactly 16 past such a boundary
    buf[off:off + nbytes].copy_(payload[:nbytes])
    out, n, rc = codec.decode(buf[off:off + nbytes], nbytes, code, len(data))
    assert rc == 0 and n == len(data) and _bytes(out[:n]) == data


@pytest.mark.gpu
def test_device_code_builder_matches_host(codec):
    """build_code_kernel (one warp) == gh_build_code on 400 histograms of every shape (ties, long codes, lengths beyond
    32 rejected the same way), header bytes and payload size included"""
    import ctypes as C
    import torch
    import golden_huffman_b200 as gh
    from test_emul_kernels import _random_histograms, _check_device_build
    rng = np.random.default_rng(77)
    keep = []

    def alloc(h):
        d_h = torch.from_numpy(h.astype(np.int64)).cuda()
        d_code = torch.zeros(C.sizeof(gh.GhDeviceCode) + 64, dtype=torch.uint8, device="cuda")
        d_hdr = torch.zeros(2048, dtype=torch.uint8, device="cuda")
        keep[:] = [d_h, d_code, d_hdr]
        return d_h.data_ptr(), d_code.data_ptr(), d_hdr.data_ptr()

    def to_host(ptr, n):
        torch.cuda.synchronize()
        for t in keep:
            if t.data_ptr() == ptr:
                return t.view(torch.uint8)[:n].cpu().numpy().tobytes()
        raise AssertionError("unknown device pointer")

    _check_device_build(codec.lib, _random_histograms(rng, 400), alloc, to_host)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zipf", "uniform", "text", "skewed"])
def test_compress_with_device_code(codec, oracle, name):
    """gh_compress_device with gh_ctx_set_device_code on (histogram -> code -> header -> packing without the host):
    the image the host-code path produces, which is the oracle's"""
    import torch
    import golden_huffman_b200.workloads as w
    n = (1 << 24) + 1234 if name != "skewed" else (1 << 25)
    x = w.WORKLOADS_TORCH[name](n, "cuda", seed=9)
    ref = codec.compress(x).clone()
    codec.lib.ctx_set_device_code(codec.ctx, True)
    try:
        img = codec.compress(x)
        assert img.numel() == ref.numel() and torch.equal(img, ref)
        out, nd, rc = codec.decompress(img, n)
        assert rc == 0 and nd == n and torch.equal(out, x)
    finally:
        codec.lib.ctx_set_device_code(codec.ctx, False)
    rc, oimg = oracle.compress(x[: 1 << 22].cpu().numpy().tobytes())
    small = codec.compress(x[: 1 << 22].clone())
    assert rc == 0 and small.cpu().numpy().tobytes() == oimg


@pytest.mark.gpu
def test_roundtrip_beyond_4gib(codec):
    """more than 2^32 input / output bytes on one device (the ABI's sizes are 64-bit; tile, subsequence and offset
    arithmetic must be too): 4.25 GiB of Zipf bytes round-trip, and the image's first 64 MiB worth of input is the
    oracle's stream (checked through a second, independent compress of that prefix being a prefix-consistent code is
    not possible -- so the check here is the round trip plus the size the code predicts)"""
    import torch
    import golden_huffman_b200.workloads as w
    free, _ = torch.cuda.mem_get_info()
    n = (17 << 28) + 12345  # 4.25 GiB + a ragged tail
    if free < 8 * n:
        pytest.skip("needs ~36 GB of free device memory")
    x = w.zipf_torch(n, "cuda", seed=12)
    hist = codec.histogram(x)
    torch.cuda.synchronize()
    assert int(hist.sum().item()) == n
    code = codec.build_code(hist)
    bits = codec.lib.payload_bits(code, hist.cpu().numpy().astype(np.uint64), with_eof=True)
    img = codec.compress(x)
    assert img.numel() == codec.lib.header_bytes(code) + (bits + 7) // 8
    out, nd, rc = codec.decompress(img, n)
    assert rc == 0 and nd == n
    # compare in pieces (torch.equal on > 2^32 elements allocates a full-size temporary)
    step = 1 << 30
    for a in range(0, n, step):
        assert torch.equal(out[a:a + step], x[a:a + step]), f"mismatch in [{a}, {a + step})"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zipf", "text", "uniform", "skewed"])
def test_multi_device_host_api(codec, oracle, name):
    """gh_compress_host_multi / gh_decompress_host_multi (one process, one host thread per shard): three shards on the
    devices that are there (all on device 0 on a one-GPU box, which exercises the same stitching and entry logic): the
    image is the oracle's, the output the input"""
    import torch
    import golden_huffman_b200.workloads as w
    ndev = torch.cuda.device_count()
    devices = [k % ndev for k in range(3)]
    n = (48 << 20) + 777
    data = w.WORKLOADS_NP[name](n, seed=17)
    rc, want = oracle.compress(data.tobytes())
    assert rc == 0
    lib = codec.lib
    img = np.zeros(lib.compress_bound(n), dtype=np.uint8)
    nb, _ = lib.compress_host_multi(devices, data.ctypes.data, n, img.ctypes.data, img.size)
    assert nb == len(want) and img[:nb].tobytes() == want
    out = np.zeros(n + 8, dtype=np.uint8)
    nd, _ = lib.decompress_host_multi(devices, img.ctypes.data, nb, out.ctypes.data, n)
    assert nd == n and (out[:n] == data).all()
