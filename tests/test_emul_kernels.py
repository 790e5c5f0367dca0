"""Kernel logic on the CPU: the product's .cu sources compiled against tests/emul/cuda_emul.h (threads of a
block = OS threads, blocks run one after another) and compared bit-for-bit with the oracle. This checks
indexing, barriers, the look-back bookkeeping, bit packing and the self-synchronising decode before any GPU
time is spent; real concurrency is only exercised by the -m gpu tests."""
import ctypes as C

import numpy as np
import pytest

import emul_lib
from emul_lib import aligned
from golden_cases import make_input


@pytest.fixture(scope="module")
def emu():
    return emul_lib.load()


@pytest.fixture(scope="module")
def ctx(emu):
    c = emu.ctx_create()
    yield c
    emu.ctx_destroy(c)


def _roundtrip(emu, ctx, oracle, data):
    n = len(data)
    d = np.frombuffer(data, dtype=np.uint8)
    din = aligned(n + 16)
    din[:n] = d
    hist = aligned(256, np.uint64)
    emu.histogram(din.ctypes.data, n, hist.ctypes.data)
    assert (hist == oracle.histogram(data)).all()
    rc, img = oracle.compress(data)
    assert rc == 0
    cap = emu.compress_bound(n)
    dout = aligned(cap)
    nb, _ = emu.compress_device(ctx, din.ctypes.data, n, dout.ctypes.data, cap)
    assert dout[:nb].tobytes() == img
    dimg = aligned(nb + 32)
    dimg[:nb] = np.frombuffer(img, dtype=np.uint8)
    dde = aligned(n + 64)
    nd, _ = emu.decompress_device(ctx, dimg.ctypes.data, nb, dde.ctypes.data, n + 64)
    assert nd == n and dde[:n].tobytes() == data
    return img


@pytest.mark.parametrize("name", ["kat1_abracadabra", "kat2_a1000", "one_byte", "two_symbols", "text_small",
                                  "tile_exact_4096", "tile_plus1_4097"])
def test_emulated_kernels_on_golden(emu, ctx, oracle, name):
    _roundtrip(emu, ctx, oracle, make_input(name))


def test_emulated_near_fixed_length_code(emu, ctx, oracle):
    """every byte value equally often -> 255 codes of 8 bits + 2 of 9: paths re-synchronise only after thousands of
    symbols, so the decoder takes its re-walk rounds over dense work lists (and coarsens the subsequences)"""
    _roundtrip(emu, ctx, oracle, make_input("kat3_allbytes512"))
    rng = np.random.default_rng(11)
    _roundtrip(emu, ctx, oracle, rng.permutation(np.repeat(np.arange(256, dtype=np.uint8), 300)).tobytes())
    # sorted: the byte with the 9-bit all-zero codeword comes as one run of 300 -> 2700 zero bits in a row, i.e. an
    # "eight zero bits start here" event at every bit position of the run (only every ninth is a codeword start)
    _roundtrip(emu, ctx, oracle, np.repeat(np.arange(256, dtype=np.uint8), 300).tobytes())


def test_emulated_encoder_phases_and_ragged_sizes(emu, ctx, oracle):
    """the single-gather encoder: ragged sizes around the 32-byte slice, the 1 KiB warp row, the 8 KiB sub-tile and the
    16 KiB tile (the slice that holds the last byte takes the codeword-by-codeword path), and every start phase of the
    sharding contract"""
    import golden_huffman_b200 as gh
    rng = np.random.default_rng(17)
    for n in (5, 31, 32, 33, 511, 1023, 1024, 1025, 2049, 8191, 8192, 8193, 16384, 16385, 40000, 70001):
        _roundtrip(emu, ctx, oracle, np.minimum(rng.geometric(0.3, n), 255).astype(np.uint8).tobytes())
    data = make_input("text_small")
    n = len(data)
    rc, code = oracle.build_code(oracle.histogram(data))
    _, full = oracle.encode_payload(data, code)
    bits = np.unpackbits(np.frombuffer(full, dtype=np.uint8))
    total = oracle.payload_bits(code, oracle.histogram(data)) - code.length[256]
    pcode = gh.GhCode.from_buffer_copy(bytes(code))
    for start_bit in (0, 5, 31, 32, 77, 127, 255):
        din = aligned(n + 16)
        din[:n] = np.frombuffer(data, dtype=np.uint8)
        cap = emu.encode_payload_capacity(n, pcode, start_bit)
        out = aligned(cap)
        out[:] = 0xAA
        ws = aligned(emu.encode_workspace_bytes(n) + 256)
        end = aligned(1, np.uint64)
        emu.encode(din.ctypes.data, n, pcode, out.ctypes.data, cap, ws.ctypes.data, ws.size, start_bit=start_bit,
                   append_eof=False, d_end_bit=end.ctypes.data)
        assert int(end[0]) == start_bit + total
        first_word = start_bit // 32 * 4
        got = np.unpackbits(out[first_word:])
        lead = start_bit - first_word * 8
        assert not got[:lead].any()
        assert (got[lead:lead + total] == bits[:total]).all()


def test_emulated_encoder_mixed_long_and_short_rows(emu, ctx, oracle):
    """a code with lengths beyond 16 whose long codewords are rare: most warp rows take the four-codewords-per-chunk
    path, the rows that contain a long codeword the slow one, inside the same tiles (8 KiB tiles for such codes)"""
    f = [1, 2]
    while len(f) < 24:
        f.append(f[-1] + f[-2])
    rng = np.random.default_rng(23)
    body = np.repeat(np.arange(len(f), dtype=np.uint8), f)
    rng.shuffle(body)
    data = np.concatenate([np.full(20000, len(f) - 1, dtype=np.uint8), body, np.full(9000, len(f) - 2, dtype=np.uint8)])
    rc, code = oracle.build_code(oracle.histogram(data.tobytes()))
    assert rc == 0 and code.max_len > 16
    _roundtrip(emu, ctx, oracle, data.tobytes())


def _big_tile_input(seed=29):
    """240 byte values that occur 36 times each below a ladder of nine values whose counts double (9000, 18000, ...):
    the rare values sit nine levels down the code tree and get 16- and 17-bit codewords. All of them are packed into
    the second 8 KiB tile, which therefore needs more than 16 bits per byte, i.e. both staging buffers of its group."""
    rng = np.random.default_rng(seed)
    rare = np.repeat(np.arange(10, 250, dtype=np.uint8), 36)
    rng.shuffle(rare)
    ladder = np.repeat(np.arange(9, dtype=np.uint8), [9000 << i for i in range(9)])
    rng.shuffle(ladder)
    return np.concatenate([ladder[:8192], rare, ladder[8192:]])


def test_emulated_encoder_big_tile(emu, oracle):
    import golden_huffman_b200 as gh
    data = _big_tile_input()
    n = data.size
    raw = data.tobytes()
    rc, code = oracle.build_code(oracle.histogram(raw))
    assert rc == 0 and max(code.length[b] for b in range(10, 250)) > 16
    assert sum(code.length[b] for b in data[8192:16384]) > 8192 * 16
    _, want = oracle.encode_payload(raw, code)
    pcode = gh.GhCode.from_buffer_copy(bytes(code))
    din = aligned(n + 32)
    din[:n] = data
    cap = emu.encode_payload_capacity(n, pcode, 0)
    out = aligned(cap)
    ws = aligned(emu.encode_workspace_bytes(n) + 256)
    end = aligned(1, np.uint64)
    emu.encode(din.ctypes.data, n, pcode, out.ctypes.data, cap, ws.ctypes.data, ws.size, start_bit=0, append_eof=True,
               d_end_bit=end.ctypes.data)
    assert out[:len(want)].tobytes() == want


def test_emulated_kernels_random(emu, ctx, oracle):
    rng = np.random.default_rng(3)
    sizes = [3, 17, 4095, 8193, 30000]
    for i, n in enumerate(sizes):
        if i % 3 == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif i % 3 == 1:
            d = np.minimum(rng.geometric(0.25, n), 255).astype(np.uint8)
        else:
            d = rng.integers(0, 3, n, dtype=np.uint8)
        _roundtrip(emu, ctx, oracle, d.tobytes())


def test_emulated_long_codes(emu, ctx, oracle):
    """Fibonacci counts -> code lengths up to 21 here: exercises the beyond-LUT search and 64-bit packing"""
    f = [1, 2]
    while len(f) < 21:
        f.append(f[-1] + f[-2])
    data = np.concatenate([np.full(c, i * 7 % 256, dtype=np.uint8) for i, c in enumerate(f)])
    np.random.default_rng(5).shuffle(data)
    img = _roundtrip(emu, ctx, oracle, data.tobytes())
    _, code = oracle.parse_header(img)
    assert code.max_len >= 20


def test_emulated_encode_start_bit_and_no_eof(emu, oracle):
    """sharding contract of gh_encode: phase-aligned slice, zero bits before start_bit, no end mark"""
    data = make_input("text_small")
    n = len(data)
    rc, code = oracle.build_code(oracle.histogram(data))
    _, full = oracle.encode_payload(data, code)
    bits = np.unpackbits(np.frombuffer(full, dtype=np.uint8))
    total = oracle.payload_bits(code, oracle.histogram(data)) - code.length[256]
    import golden_huffman_b200 as gh
    pcode = gh.GhCode.from_buffer_copy(bytes(code))
    for start_bit in (0, 5, 31, 32, 77, 127):
        din = aligned(n + 16)
        din[:n] = np.frombuffer(data, dtype=np.uint8)
        cap = emu.encode_payload_capacity(n, pcode, start_bit)
        out = aligned(cap)
        out[:] = 0xAA
        ws = aligned(emu.encode_workspace_bytes(n) + 256)
        end = aligned(1, np.uint64)
        emu.encode(din.ctypes.data, n, pcode, out.ctypes.data, cap, ws.ctypes.data, ws.size, start_bit=start_bit,
                   append_eof=False, d_end_bit=end.ctypes.data)
        assert int(end[0]) == start_bit + total
        first_word = start_bit // 32 * 4
        assert (out[:first_word] == 0xAA).all()  # untouched before the first word
        got = np.unpackbits(out[first_word:])
        lead = start_bit - first_word * 8
        assert not got[:lead].any()
        assert (got[lead:lead + total] == bits[:total]).all()


@pytest.mark.parametrize("case", ["text", "near_fixed"])
def test_emulated_sharded_decode_sync(emu, oracle, case):
    """the two halves of gh_decode on a payload cut in two: slice 1 first assumes entry 0, then is corrected
    with slice 0's exit_bit; concatenated output equals the input. near_fixed: the 8/9-bit code of equally frequent
    bytes, which is synchronised by the phase walk (transfer functions + scan) instead of rounds"""
    import golden_huffman_b200 as gh
    data = make_input("text_small") * 3 if case == "text" else make_input("kat3_allbytes512")
    rc, code = oracle.build_code(oracle.histogram(data))
    _, payload = oracle.encode_payload(data, code)
    pcode = gh.GhCode.from_buffer_copy(bytes(code))
    cut = (len(payload) // 2) // 16 * 16
    buf = aligned(len(payload) + 16)
    buf[:len(payload)] = np.frombuffer(payload, dtype=np.uint8)
    slices = [(0, cut, min(cut + 8, len(payload))), (cut, len(payload) - cut, len(payload) - cut)]
    ws = [aligned(emu.decode_workspace_bytes(s[1]) + 256) for s in slices]
    res = []
    for k, (off, nb, readable) in enumerate(slices):
        res.append(emu.decode_sync(buf.ctypes.data + off, nb, readable, pcode, 0, True, ws[k].ctypes.data, ws[k].size))
    assert not res[0].eof_found
    res[1] = emu.decode_sync(buf.ctypes.data + cut, slices[1][1], slices[1][2], pcode, res[0].exit_bit, False,
                             ws[1].ctypes.data, ws[1].size)
    assert res[1].eof_found and res[0].n_symbols + res[1].n_symbols == len(data)
    out = aligned(len(data) + 16)
    o = 0
    for k, (off, nb, readable) in enumerate(slices):
        emu.decode_write(buf.ctypes.data + off, nb, readable, pcode, out.ctypes.data + o, res[k].n_symbols,
                         ws[k].ctypes.data, ws[k].size)
        o += res[k].n_symbols
    assert out[:len(data)].tobytes() == data


def test_emulated_decode_errors(emu, ctx, oracle):
    import golden_huffman_b200 as gh
    data = make_input("text_small")
    rc, img = oracle.compress(data)
    n = len(data)
    dimg = aligned(len(img) + 32)
    dimg[:len(img)] = np.frombuffer(img, dtype=np.uint8)
    out = aligned(n + 64)
    # output too small: the count is still reported
    nd, rc = emu.decompress_device(ctx, dimg.ctypes.data, len(img), out.ctypes.data, n - 10, allow=(gh.capi.GH_ERR_SPACE,))
    assert rc == gh.capi.GH_ERR_SPACE and nd == n
    # truncated stream: no end mark
    nd, rc = emu.decompress_device(ctx, dimg.ctypes.data, len(img) - 40, out.ctypes.data, n + 64,
                                   allow=(gh.capi.GH_ERR_NO_EOF,))
    assert rc == gh.capi.GH_ERR_NO_EOF


def test_emulated_staged_entry_points(emu, ctx, oracle):
    """the four-step calls the C++ adapters make (gh_stage_input / gh_encode_staged / gh_stage_payload / gh_decode_staged)"""
    import golden_huffman_b200 as gh
    data = make_input("tile_plus1_4097")
    n = len(data)
    src = np.frombuffer(data, dtype=np.uint8).copy()
    hist = np.zeros(256, dtype=np.uint64)
    emu.check(emu.lib.gh_stage_input(ctx, src.ctypes.data, n, hist.ctypes.data), "gh_stage_input")
    assert (hist == oracle.histogram(data)).all()
    code = emu.build_code(hist)
    rc, img = oracle.compress(data)
    hdr = emu.write_header(code)
    out = np.zeros(n * 4 + 64, dtype=np.uint8)
    nb = C.c_uint64(0)
    emu.check(emu.lib.gh_encode_staged(ctx, C.byref(code), out.ctypes.data, out.size, C.byref(nb)), "gh_encode_staged")
    assert hdr + out[: nb.value].tobytes() == img
    payload = np.frombuffer(img[len(hdr):], dtype=np.uint8).copy()
    nsym = C.c_uint64(0)
    emu.check(emu.lib.gh_stage_payload(ctx, payload.ctypes.data, payload.size, C.byref(code), C.byref(nsym)), "gh_stage_payload")
    assert nsym.value == n
    back = np.zeros(n, dtype=np.uint8)
    emu.check(emu.lib.gh_decode_staged(ctx, back.ctypes.data, n), "gh_decode_staged")
    assert back.tobytes() == data


@pytest.mark.parametrize("kind", ["uniform", "skewed", "zipf"])
def test_emulated_decode_sync_rounds(emu, oracle, kind):
    """larger payloads through gh_decode_sync/gh_decode_write: uniform bytes (8/9-bit code whose 9-bit end mark
    shows up on every mis-phased path: paths must run through false end marks), a 1-bit-dominated code (15
    codewords per table lookup) and Zipf; also reports how many synchronisation rounds were needed"""
    import golden_huffman_b200 as gh
    import golden_huffman_b200.workloads as W
    n = 200000
    if kind == "uniform":
        data = W.uniform_np(n, seed=3)
    elif kind == "zipf":
        data = W.zipf_np(n, seed=3)
    else:
        rng = np.random.default_rng(3)
        data = np.where(rng.random(n) < 0.97, 7, rng.integers(0, 40, n)).astype(np.uint8)
    raw = data.tobytes()
    rc, code = oracle.build_code(oracle.histogram(raw))
    _, payload = oracle.encode_payload(raw, code)
    pcode = gh.GhCode.from_buffer_copy(bytes(code))
    buf = aligned(len(payload) + 16)
    buf[:len(payload)] = np.frombuffer(payload, dtype=np.uint8)
    ws = aligned(emu.decode_workspace_bytes(len(payload)) + 256)
    res = emu.decode_sync(buf.ctypes.data, len(payload), len(payload), pcode, 0, True, ws.ctypes.data, ws.size)
    assert res.eof_found and res.n_symbols == n
    assert res.rounds <= 64, f"{kind}: {res.rounds} rounds at {res.sub_bytes}-byte subsequences"
    out = aligned(n + 16)
    emu.decode_write(buf.ctypes.data, len(payload), len(payload), pcode, out.ctypes.data, n, ws.ctypes.data, ws.size)
    assert out[:n].tobytes() == raw
    print(kind, "rounds", res.rounds, "sub_bytes", res.sub_bytes)


def test_emulated_histogram_pipelined_path(emu, oracle):
    """large enough for the software-pipelined main loop of K1 (several batches per thread), misaligned start,
    plus the accumulate flag"""
    rng = np.random.default_rng(11)
    n = 400003
    raw = aligned(n + 64)
    raw[:] = rng.integers(0, 256, n + 64, dtype=np.uint8)
    for off in (0, 3, 16):
        view = raw[off:off + n]
        hist = aligned(256, np.uint64)
        emu.histogram(view.ctypes.data, n, hist.ctypes.data)
        assert (hist == oracle.histogram(view.tobytes())).all()
    hist = aligned(256, np.uint64)
    emu.histogram(raw.ctypes.data, 1000, hist.ctypes.data)
    emu.histogram(raw.ctypes.data + 1000, n - 1000, hist.ctypes.data, accumulate=True)
    assert (hist == oracle.histogram(raw[:n].tobytes())).all()


@pytest.mark.parametrize("mode", ["coarse", "no_phase_walk"])
def test_emulated_forced_pipelines(emu, ctx, oracle, mode):
    """the library's test hooks (thread-per-subsequence pipeline forced; phase walk of 8/9-bit codes switched off, so
    that uniform bytes take the general re-walk rounds) give the oracle's bytes on every kind of code"""
    import golden_huffman_b200.workloads as W
    emu.lib.gh_debug_select_writer(1 if mode == "coarse" else 0)
    emu.lib.gh_debug_disable_phase_walk(1 if mode == "no_phase_walk" else 0)
    try:
        rng = np.random.default_rng(8)
        cases = [make_input("text_small") * 30, W.zipf_np(70001, seed=2).tobytes(),
                 np.where(rng.random(90000) < 0.97, 7, rng.integers(0, 40, 90000)).astype(np.uint8).tobytes(),
                 W.uniform_np(50000, seed=4).tobytes(), b"Z", make_input("kat1_abracadabra"),
                 make_input("fib24_shuffled")]
        for data in cases:
            _roundtrip(emu, ctx, oracle, data)
    finally:
        emu.lib.gh_debug_select_writer(0)
        emu.lib.gh_debug_disable_phase_walk(0)


def _random_histograms(rng, count):
    """histograms of many shapes: flat (every tie goes through the heap's order), geometric, Fibonacci (lengths up to
    32 and beyond), sparse, single-valued"""
    out = []
    for i in range(count):
        kind = i % 6
        k = int(rng.integers(1, 257))
        h = np.zeros(256, dtype=np.uint64)
        idx = rng.permutation(256)[:k]
        if kind == 0:
            h[idx] = rng.integers(1, 4, k)                 # heavy ties
        elif kind == 1:
            h[idx] = rng.integers(1, 1 << 20, k)
        elif kind == 2:
            h[idx] = (rng.geometric(0.02, k)).astype(np.uint64)
        elif kind == 3:
            f = [1, 2]
            while len(f) < k:
                f.append(f[-1] + f[-2])
            h[idx] = np.array(f[:k], dtype=np.float64).clip(max=2 ** 62).astype(np.uint64)  # long codes; may exceed 32
        elif kind == 4:
            h[idx] = int(rng.integers(1, 1000))            # all equal
        else:
            h[idx] = rng.integers(0, 3, k)                 # zeros among them
        out.append(h)
    return out


def _check_device_build(lib, hists, device_alloc, to_host):
    """gh_build_code_device == gh_build_code field by field, header and payload size included"""
    import golden_huffman_b200 as gh
    from golden_huffman_b200.capi import GhError
    for h in hists:
        d_h, d_code, d_hdr = device_alloc(h)
        lib.build_code_device(d_h, 1, d_code, d_hdr)
        got = gh.GhDeviceCode.from_buffer_copy(to_host(d_code, C.sizeof(gh.GhDeviceCode)))
        try:
            want = lib.build_code(h)
            rc = 0
        except GhError as e:
            rc = e.status
        assert got.status == rc, (got.status, rc)
        if rc != 0:
            continue
        assert bytes(got.code) == bytes(want)
        hdr = lib.write_header(want)
        assert got.header_bytes == len(hdr) and to_host(d_hdr, len(hdr)) == hdr
        assert got.payload_bits == lib.payload_bits(want, h, with_eof=True)
        assert got.total_symbols == int(h.sum())
        assert list(got.table.codeword) == list(want.codeword) and list(got.table.length)[:257] == list(want.length)


def test_emulated_device_code_builder(emu):
    """the one-warp code builder (csrc/gh_build.cu) against the host builder: 300 histograms of every shape"""
    import golden_huffman_b200 as gh
    rng = np.random.default_rng(41)
    keep = []

    def alloc(h):
        d_h = aligned(256, np.uint64)
        d_h[:] = h
        d_code = aligned(C.sizeof(gh.GhDeviceCode) + 64)
        d_hdr = aligned(2048)
        keep[:] = [d_h, d_code, d_hdr]
        return d_h.ctypes.data, d_code.ctypes.data, d_hdr.ctypes.data

    def to_host(ptr, n):
        return C.string_at(ptr, n)

    _check_device_build(emu, _random_histograms(rng, 300), alloc, to_host)


def test_emulated_compress_with_device_code(emu, oracle):
    """gh_compress_device with the code built on the device: the oracle's image"""
    c = emu.ctx_create()
    emu.ctx_set_device_code(c, True)
    try:
        rng = np.random.default_rng(5)
        for data in (make_input("kat1_abracadabra"), make_input("text_small"), make_input("kat3_allbytes512"),
                     np.minimum(rng.geometric(0.2, 50001), 255).astype(np.uint8).tobytes(), b"x"):
            _roundtrip(emu, c, oracle, data)
    finally:
        emu.ctx_destroy(c)


def test_emulated_histogram_epochs(emu):
    """K1 counts in 16-bit halves of its shared-memory bins and folds them every 4088 vectors per thread: an input long
    enough for two folds on the shim's two SMs, once with a single byte value (the counters' worst case), once random,
    and from a misaligned address"""
    n = 2 * 448 * 4088 * 16 + 5_000_001
    rng = np.random.default_rng(1)
    for fill in (7, None):
        d = aligned(n + 16)
        d[:n] = fill if fill is not None else rng.integers(0, 256, n, dtype=np.uint8)
        h = aligned(256, np.uint64)
        for off in (0, 3):
            emu.histogram(d.ctypes.data + off, n - off, h.ctypes.data)
            assert (h == np.bincount(d[off:n], minlength=256).astype(np.uint64)).all(), (fill, off)


def test_emulated_host_decompress_in_chunks(emu, oracle):
    """gh_decompress_host cuts the payload into chunks that decode in order (each from the previous one's exit bit): with
    4 KiB and 12 KiB chunks the bytes are those of the one-chunk decode, the symbol count survives a too small output
    buffer, and a truncated image is still reported as such"""
    import golden_huffman_b200 as gh
    rng = np.random.default_rng(11)
    cases = [make_input("text_small") * 9,
             np.minimum(rng.geometric(0.08, 70001), 255).astype(np.uint8).tobytes(),
             bytes(range(256)) * 150 + b"\x07"]                       # {8, 9}-bit code: the phase walk in every chunk
    c = emu.ctx_create()
    try:
        for data in cases:
            n = len(data)
            rc, img = oracle.compress(data)
            assert rc == 0
            src = np.frombuffer(img, dtype=np.uint8).copy()
            for chunk in (4096, 12288, 131072, 0):
                emu.ctx_set_host_chunk(c, chunk)
                out = np.zeros(n + 8, dtype=np.uint8)
                nd, rc = emu.decompress_host(c, src.ctypes.data, len(img), out.ctypes.data, n + 8)
                assert rc == 0 and nd == n and out[:n].tobytes() == data, (chunk, n)
            emu.ctx_set_host_chunk(c, 4096)
            small = np.zeros(n - 5000, dtype=np.uint8)
            nd, rc = emu.decompress_host(c, src.ctypes.data, len(img), small.ctypes.data, n - 5000, allow=(gh.capi.GH_ERR_SPACE,))
            assert rc == gh.capi.GH_ERR_SPACE and nd == n
            out = np.zeros(n + 8, dtype=np.uint8)
            nd, rc = emu.decompress_host(c, src.ctypes.data, len(img) - 4200, out.ctypes.data, n + 8, allow=(gh.capi.GH_ERR_NO_EOF,))
            assert rc == gh.capi.GH_ERR_NO_EOF
    finally:
        emu.ctx_destroy(c)


def _stream_file_roundtrip(lib, oracle, tmp_path, data, chunk, resident):
    """the streaming layer through the C ABI, as the adapters drive it: pass 1, code, header, pass 2; then decode"""
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    libc.fwrite.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    libc.fseek.argtypes = [C.c_void_p, C.c_long, C.c_int]
    src, crs, dec = tmp_path / "in.bin", tmp_path / "out.crs2", tmp_path / "out.de"
    src.write_bytes(data)
    s = C.c_void_p(0)
    lib.check(lib.lib.gh_stream_create(C.byref(s), chunk, resident), "gh_stream_create")
    try:
        fin, fout = libc.fopen(str(src).encode(), b"rb"), libc.fopen(str(crs).encode(), b"wb")
        hist = np.zeros(256, dtype=np.uint64)
        lib.check(lib.lib.gh_stream_histogram(s, fin, hist.ctypes.data), "gh_stream_histogram")
        assert (hist == oracle.histogram(data)).all()
        code = lib.build_code(hist)
        hdr = lib.write_header(code)
        libc.fwrite(hdr, 1, len(hdr), fout)
        nbytes = C.c_uint64(0)
        lib.check(lib.lib.gh_stream_encode(s, fin, fout, C.byref(code), C.byref(nbytes)), "gh_stream_encode")
        libc.fclose(fin)
        libc.fclose(fout)
        rc, want = oracle.compress(data)
        assert rc == 0 and crs.read_bytes() == want and nbytes.value == len(want) - len(hdr)
        fin, fout = libc.fopen(str(crs).encode(), b"rb"), libc.fopen(str(dec).encode(), b"wb")
        code2, hb = lib.parse_header(want[:2048])
        libc.fseek(fin, hb, 0)
        n_out = C.c_uint64(0)
        lib.check(lib.lib.gh_stream_decode(s, fin, fout, C.byref(code2), hb, C.byref(n_out)), "gh_stream_decode")
        libc.fclose(fin)
        libc.fclose(fout)
        assert n_out.value == len(data) and dec.read_bytes() == data
    finally:
        lib.lib.gh_stream_destroy(s)


@pytest.mark.parametrize("resident", [0, 1 << 30])
def test_emulated_streaming_files(emu, oracle, tmp_path, resident):
    """files streamed in 64 KiB / 12 KiB chunks (multi-pass when nothing stays resident): chunk boundaries fall inside
    bytes of the payload (encode) and inside codewords (decode); one-chunk and one-byte files too"""
    import golden_huffman_b200.workloads as W
    for data, chunk in ((make_input("text_small") * 40, 65536), (W.zipf_np(200001, seed=3).tobytes(), 12288),
                        (W.uniform_np(70000, seed=4).tobytes(), 4096), (b"abracadabra", 4096), (b"q", 4096)):
        _stream_file_roundtrip(emu, oracle, tmp_path, data, chunk, resident)


def _multi_roundtrip(lib, oracle, data, shards):
    n = len(data)
    src = np.frombuffer(data, dtype=np.uint8).copy()
    rc, want = oracle.compress(data)
    assert rc == 0
    img = np.zeros(lib.compress_bound(n), dtype=np.uint8)
    nb, _ = lib.compress_host_multi(shards, src.ctypes.data, n, img.ctypes.data, img.size)
    assert img[:nb].tobytes() == want
    out = np.zeros(n + 8, dtype=np.uint8)
    nd, _ = lib.decompress_host_multi(shards, img.ctypes.data, nb, out.ctypes.data, n)
    assert nd == n and out[:n].tobytes() == data


def test_emulated_multi_shard_host_api(emu, oracle):
    """gh_compress_host_multi / gh_decompress_host_multi: 1, 2, 3 and 5 shards (here all on the one emulated device): the
    shards meet inside bytes (encode) and inside codewords (decode: entries found by walking a 4 KiB left halo, checked
    against the neighbours' exits); inputs too small for that many shards fall back to fewer"""
    import golden_huffman_b200.workloads as W
    for data in (make_input("text_small") * 120, W.zipf_np(150001, seed=8).tobytes(), W.uniform_np(90000, seed=9).tobytes(),
                 make_input("kat1_abracadabra")):
        for shards in ([0], [0, 0], [0, 0, 0], [0] * 5):
            _multi_roundtrip(emu, oracle, data, shards)
