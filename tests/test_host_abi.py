"""No-GPU checks of the product library: it loads, exports every symbol include/gh_codec.h declares, and its
HOST entry points (code construction, header) are byte-identical to the oracle. No device call is made."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

from golden_cases import CASES, make_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
META = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gh_codec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gh_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    import golden_huffman_b200 as gh
    declared = _declared_symbols()
    assert declared, "no declarations found"
    assert sorted(gh.SIGNATURES) == declared


def test_library_exports_every_declared_symbol(ghlib):
    raw = C.CDLL(ghlib.path)
    for name in _declared_symbols():
        assert hasattr(raw, name), name


def test_struct_layout(ghlib):
    import golden_huffman_b200 as gh
    from oracle_lib import GhCode as OracleCode
    assert C.sizeof(gh.GhCode) == C.sizeof(OracleCode) == (3 * 257 + 2 + 33 + 33) * 4
    assert C.sizeof(gh.GhShardSync) == 24


def _same_code(a, b):
    for f in ("length", "codeword", "symbol"):
        assert list(getattr(a, f)) == list(getattr(b, f)), f
    assert (a.min_len, a.max_len) == (b.min_len, b.max_len)
    m = a.max_len
    assert list(a.start_pos)[1:m + 1] == list(b.start_pos)[1:m + 1]
    assert list(a.first_code)[1:m + 1] == list(b.first_code)[1:m + 1]


@pytest.mark.parametrize("name", sorted(CASES))
def test_build_code_and_header_match_oracle_on_golden(ghlib, oracle, name):
    data = make_input(name)
    hist = oracle.histogram(data)
    rc, ocode = oracle.build_code(hist)
    assert rc == 0
    code = ghlib.build_code(hist)
    _same_code(code, ocode)
    hdr = ghlib.write_header(code)
    assert hdr == oracle.write_header(ocode)
    # and the header bytes are the reference's: prefix of the golden image
    rc, img = oracle.compress(data)
    assert hashlib.sha256(img).hexdigest() == META[name]["sha256_crs2"]
    assert img[: len(hdr)] == hdr
    assert ghlib.payload_bits(code, hist) == oracle.payload_bits(ocode, hist)
    assert (ghlib.payload_bits(code, hist) + 7) // 8 == len(img) - len(hdr)
    code2, n = ghlib.parse_header(img)
    assert n == len(hdr)
    _same_code(code2, ocode)


def test_build_code_random_histograms(ghlib, oracle):
    rng = np.random.default_rng(7)
    for trial in range(400):
        k = int(rng.integers(1, 257))
        hist = np.zeros(256, dtype=np.uint64)
        idx = rng.choice(256, size=k, replace=False)
        mode = trial % 4
        if mode == 0:
            hist[idx] = rng.integers(1, 1000, k)
        elif mode == 1:
            hist[idx] = rng.integers(1, 4, k)  # many ties
        elif mode == 2:
            hist[idx] = (rng.pareto(0.7, k) * 10 + 1).astype(np.uint64)
        else:
            hist[idx] = 1 << rng.integers(0, 28, k)
        rc, ocode = oracle.build_code(hist)
        if rc != 0:
            with pytest.raises(Exception):
                ghlib.build_code(hist)
            continue
        _same_code(ghlib.build_code(hist), ocode)


def test_rejects_undefined_inputs(ghlib):
    import golden_huffman_b200 as gh
    with pytest.raises(gh.GhError) as e:
        ghlib.build_code(np.zeros(256, dtype=np.uint64))
    assert e.value.status == gh.capi.GH_ERR_EMPTY
    f = [1, 2]
    while len(f) < 33:
        f.append(f[-1] + f[-2])
    hist = np.zeros(256, dtype=np.uint64)
    hist[:33] = f
    with pytest.raises(gh.GhError) as e:
        ghlib.build_code(hist)
    assert e.value.status == gh.capi.GH_ERR_TOO_LONG
    with pytest.raises(gh.GhError):
        ghlib.parse_header(b"\x00" * 2000)
    with pytest.raises(gh.GhError):
        ghlib.parse_header(b"\x00\x00\x01\x01" + b"\x00" * 100)


def test_no_cpu_fallback_without_gpu(ghlib):
    """without a device every device entry point must fail loudly, never compute on the host"""
    import torch
    import golden_huffman_b200 as gh
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    buf = np.zeros(4096, dtype=np.uint8)
    hist = np.zeros(256, dtype=np.uint64)
    with pytest.raises(gh.GhError) as e:
        ghlib.histogram(buf.ctypes.data, buf.size, hist.ctypes.data)
    assert e.value.status == gh.capi.GH_ERR_CUDA
    assert hist.sum() == 0
    with pytest.raises(gh.GhError) as e:
        ghlib.ctx_create()
    assert e.value.status == gh.capi.GH_ERR_CUDA
    with pytest.raises(RuntimeError):
        gh.Codec(ghlib)
