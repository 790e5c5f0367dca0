"""The oracle itself: pinned against the reference's known answers before anything is compared with it."""
import hashlib
import json
import os

import numpy as np
import pytest

from golden_cases import CASES, SMALL, make_input

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
META = json.load(open(os.path.join(GOLDEN, "golden.json")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(oracle, name):
    """compressed bytes identical to what the compiled reference produced (sha256 / stored file), and the
    bit-serial decoder (reference canonical_huff_encoder.cc:377-419) restores the input"""
    data = make_input(name)
    m = META[name]
    assert len(data) == m["n"] and hashlib.sha256(data).hexdigest() == m["sha256_in"]
    rc, img = oracle.compress(data)
    assert rc == 0
    assert len(img) == m["crs2_bytes"]
    assert hashlib.sha256(img).hexdigest() == m["sha256_crs2"]
    if name in SMALL:
        assert img == open(os.path.join(GOLDEN, name + ".crs2"), "rb").read()
        assert data == open(os.path.join(GOLDEN, name + ".in"), "rb").read()
    rc, back = oracle.decompress(img, len(data))
    assert rc == 0 and back == data


def test_kat1_tables(oracle):
    """SURVEY.md 8c KAT-1, hand-verified: codes a=1 b=001 d=010 r=011 c=0000 EOF=0001, payload 97 0a 97 1f"""
    data = b"abracadabra"
    rc, code = oracle.build_code(oracle.histogram(data))
    assert rc == 0
    d = code.as_dict()
    assert [s for s in d["symbol"] if s != 0xFFFFFFFF] == [97, 98, 100, 114, 99, 256]
    assert (d["min_len"], d["max_len"]) == (1, 4)
    assert d["start_pos"][1:] == [0, 1, 1, 4] and d["first_code"][1:] == [1, 2, 1, 0]
    enc = {chr(s) if s < 256 else "EOF": format(d["codeword"][s], "0%db" % d["length"][s]) for s in (97, 98, 100, 114, 99, 256)}
    assert enc == {"a": "1", "b": "001", "d": "010", "r": "011", "c": "0000", "EOF": "0001"}
    rc, payload = oracle.encode_payload(data, code)
    assert rc == 0 and payload == bytes([0x97, 0x0A, 0x97, 0x1F])
    assert len(oracle.write_header(code)) == 1072


def test_kat3_tie_break(oracle):
    """all 256 bytes equally frequent: the end mark pairs with byte 2 (libstdc++ heap order)"""
    rc, code = oracle.build_code(np.full(256, 512, dtype=np.uint64))
    assert rc == 0
    assert [s for s in range(257) if code.length[s] == 9] == [2, 256]
    assert sum(1 for s in range(256) if code.length[s] == 8) == 255
    assert code.first_code[8] == 1 and code.first_code[9] == 0


def test_undefined_inputs_are_rejected(oracle):
    rc, _ = oracle.compress(b"")
    assert rc == 1  # GHO_ERR_EMPTY
    f = [1, 2]
    while len(f) < 33:
        f.append(f[-1] + f[-2])
    hist = np.zeros(256, dtype=np.uint64)
    hist[:33] = f  # 33 Fibonacci bytes + the end mark -> max_len 33, outside the reference's domain
    rc, _ = oracle.build_code(hist)
    assert rc == 2  # GHO_ERR_TOO_LONG


def test_header_roundtrip(oracle):
    data = make_input("text_small")
    rc, code = oracle.build_code(oracle.histogram(data))
    hdr = oracle.write_header(code)
    assert len(hdr) == 1040 + 8 * code.max_len
    n, code2 = oracle.parse_header(hdr)
    assert n == len(hdr)
    for f in ("symbol", "min_len", "max_len", "start_pos", "first_code", "length", "codeword"):
        a, b = getattr(code, f), getattr(code2, f)
        if f in ("start_pos", "first_code"):
            assert list(a)[1:code.max_len + 1] == list(b)[1:code.max_len + 1]
        elif f in ("min_len", "max_len"):
            assert a == b
        else:
            assert list(a) == list(b), f


def test_oracle_vs_compiled_reference_random(oracle, reference):
    """build container only: byte-for-byte against the unmodified reference on randomised inputs, and the
    reference's three decoders agree with the oracle's decoder"""
    rng = np.random.default_rng(42)
    for trial in range(120):
        n = int(rng.integers(1, 20000))
        kind = trial % 5
        if kind == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 1:
            d = rng.integers(0, int(rng.integers(1, 8)), n, dtype=np.uint8)
        elif kind == 2:
            d = np.minimum(rng.geometric(0.3, n), 255).astype(np.uint8)
        elif kind == 3:
            d = (rng.zipf(1.3, n) % 256).astype(np.uint8)
        else:
            d = np.repeat(rng.integers(0, 256, 8, dtype=np.uint8), rng.integers(1, 300, 8))
        data = d.tobytes()
        rc, img = oracle.compress(data)
        assert rc == 0
        assert img == reference.compress(data), (trial, len(data))
        kinds = ("simple", "fast", "table") if len(img) > 1400 else ("simple",)
        for k in kinds:
            assert reference.decompress(img, k) == data
        rc, back = oracle.decompress(img, len(data))
        assert rc == 0 and back == data
