"""Property tests on the emulated kernels (hypothesis): for arbitrary byte strings the device path is bit-identical
to the oracle, decode(encode(x)) == x, and the closed-form payload size from the histogram holds."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import emul_lib
from emul_lib import aligned


@pytest.fixture(scope="module")
def emu():
    return emul_lib.load()


@pytest.fixture(scope="module")
def ctx(emu):
    c = emu.ctx_create()
    yield c
    emu.ctx_destroy(c)


def _alphabets():
    # few symbols (short codes, 15 codewords per lookup), full byte range, and heavily repeated runs
    return st.one_of(
        st.binary(min_size=1, max_size=6000),
        st.lists(st.sampled_from([0, 1, 2, 255]), min_size=1, max_size=9000).map(bytes),
        st.tuples(st.binary(min_size=1, max_size=40), st.integers(1, 300)).map(lambda t: t[0] * t[1]),
    )


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(data=_alphabets())
def test_device_path_equals_oracle(emu, ctx, oracle, data):
    n = len(data)
    din = aligned(n + 16)
    din[:n] = np.frombuffer(data, dtype=np.uint8)
    rc, img = oracle.compress(data)
    assert rc == 0
    cap = emu.compress_bound(n)
    dout = aligned(cap)
    nb, _ = emu.compress_device(ctx, din.ctypes.data, n, dout.ctypes.data, cap)
    assert dout[:nb].tobytes() == img
    hist = oracle.histogram(data)
    code = emu.build_code(hist)
    assert nb == emu.header_bytes(code) + (emu.payload_bits(code, hist) + 7) // 8
    dde = aligned(n + 32)
    nd, _ = emu.decompress_device(ctx, dout.ctypes.data, nb, dde.ctypes.data, n + 32)
    assert nd == n and dde[:n].tobytes() == data
