"""Damaged images through the decoder (kernels under the CPU shim): the reference's decoders have undefined behaviour on
malformed input (SURVEY 8b "Errors"); the C ABI has to come back with a status -- any status -- without touching memory
outside its buffers. Run under ASan with  python tests/emul/build_emul.py --asan  +  LD_PRELOAD=libasan  to check reads too;
in the normal suite the output buffer's guard bytes are the check."""
import numpy as np
import pytest

from emul_lib import aligned, load
from golden_cases import make_input
from oracle_lib import Oracle


@pytest.fixture(scope="module")
def emu():
    return load()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


ANY = tuple(range(0, 8))
GUARD = 4096


def _decode_guarded(emu, ctx, img_bytes, cap):
    nb = len(img_bytes)
    dimg = aligned(nb + 64)
    dimg[:nb] = np.frombuffer(img_bytes, dtype=np.uint8)
    out = aligned(cap + 2 * GUARD)
    out[:] = 0xA5
    body = out[GUARD:GUARD + cap]
    nd, rc = emu.decompress_device(ctx, dimg.ctypes.data, nb, body.ctypes.data, cap, allow=ANY)
    assert (out[:GUARD] == 0xA5).all() and (out[GUARD + cap:] == 0xA5).all(), "decoder wrote outside its output buffer"
    return nd, rc, body


def _images(oracle):
    rng = np.random.default_rng(3)
    datas = [make_input("text_small") * 3,
             np.minimum(rng.geometric(0.15, 40001), 255).astype(np.uint8).tobytes(),
             bytes(range(256)) * 40]
    out = []
    for d in datas:
        rc, img = oracle.compress(d)
        assert rc == 0
        out.append((d, img))
    return out


def test_flipped_payload_bits(emu, oracle):
    """1 to 8 flipped payload bits: the decode re-synchronises or runs to a different end; the call returns"""
    import golden_huffman_b200 as gh
    rng = np.random.default_rng(17)
    ctx = emu.ctx_create()
    try:
        for data, img in _images(oracle):
            code, hdr = emu.parse_header(img[:2048])
            for _ in range(12):
                bad = bytearray(img)
                for _ in range(int(rng.integers(1, 9))):
                    at = int(rng.integers(hdr, len(img)))
                    bad[at] ^= 1 << int(rng.integers(0, 8))
                nd, rc, _ = _decode_guarded(emu, ctx, bytes(bad), len(data) + 64)
                assert rc in (gh.capi.GH_OK, gh.capi.GH_ERR_SPACE, gh.capi.GH_ERR_NO_EOF, gh.capi.GH_ERR_FORMAT)
    finally:
        emu.ctx_destroy(ctx)


def test_truncated_and_padded_images(emu, oracle):
    import golden_huffman_b200 as gh
    ctx = emu.ctx_create()
    try:
        for data, img in _images(oracle):
            code, hdr = emu.parse_header(img[:2048])
            for cut in (hdr + 1, hdr + 7, hdr + 33, (hdr + len(img)) // 2, len(img) - 1):
                nd, rc, _ = _decode_guarded(emu, ctx, img[:cut], len(data) + 64)
                assert rc in (gh.capi.GH_ERR_NO_EOF, gh.capi.GH_OK, gh.capi.GH_ERR_SPACE)
            for cut in (0, 3, 1027, hdr - 1, hdr):
                nd, rc, _ = _decode_guarded(emu, ctx, img[:cut] if cut else b"\x00", len(data) + 64)
                assert rc in (gh.capi.GH_ERR_FORMAT, gh.capi.GH_ERR_NO_EOF)
            nd, rc, body = _decode_guarded(emu, ctx, img + b"\x00" * 777, len(data) + 64)
            assert rc == gh.capi.GH_OK and nd == len(data) and body[:nd].tobytes() == data
    finally:
        emu.ctx_destroy(ctx)


def test_damaged_headers(emu, oracle):
    """random damage to the header's tables: rejected as GH_ERR_FORMAT or decoded to something, never out of bounds"""
    import golden_huffman_b200 as gh
    rng = np.random.default_rng(29)
    ctx = emu.ctx_create()
    try:
        for data, img in _images(oracle):
            code, hdr = emu.parse_header(img[:2048])
            for trial in range(40):
                bad = bytearray(img)
                kind = trial % 4
                if kind == 0:    # one byte anywhere in the header
                    bad[int(rng.integers(0, hdr))] = int(rng.integers(0, 256))
                elif kind == 1:  # a whole big-endian word of the symbol table
                    at = 4 + 4 * int(rng.integers(0, 257))
                    bad[at:at + 4] = int(rng.integers(0, 2**32)).to_bytes(4, "big")
                elif kind == 2:  # min_len / max_len
                    at = 4 + 4 * 257 + 4 * int(rng.integers(0, 2))
                    bad[at:at + 4] = int(rng.integers(0, 40)).to_bytes(4, "big")
                else:            # a start_pos / first_code word
                    at = 4 + 4 * 257 + 8 + 4 * int(rng.integers(0, 2 * code.max_len))
                    bad[at:at + 4] = int(rng.integers(0, 2**(int(rng.integers(1, 33))))).to_bytes(4, "big")
                nd, rc, _ = _decode_guarded(emu, ctx, bytes(bad), len(data) + 64)
                assert rc in ANY
    finally:
        emu.ctx_destroy(ctx)
