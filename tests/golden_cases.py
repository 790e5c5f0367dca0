"""Recipes of the golden inputs (deterministic, no external data). KAT-1..4 are SURVEY.md section 8c."""
import numpy as np

TEXT = (b"It was the best of times, it was the worst of times, it was the age of wisdom, it was the age of "
        b"foolishness, it was the epoch of belief, it was the epoch of incredulity, it was the season of Light, "
        b"it was the season of Darkness, it was the spring of hope, it was the winter of despair.\n")


def _fib(k):
    f = [1, 2]
    while len(f) < k:
        f.append(f[-1] + f[-2])
    return f[:k]


def _lcg_bytes(n, seed):
    """tiny portable generator so the recipe does not depend on a numpy version"""
    out = np.empty(n, dtype=np.uint8)
    x = np.uint64(seed)
    a, c = np.uint64(6364136223846793005), np.uint64(1442695040888963407)
    with np.errstate(over="ignore"):
        for i in range(n):
            x = x * a + c
            out[i] = (x >> np.uint64(33)) & np.uint64(0xFF)
    return out


def _skew_small(n, seed):
    """geometric-ish byte distribution from the LCG: long codes appear, lengths stay < 32"""
    r = _lcg_bytes(2 * n, seed).astype(np.uint32)
    v = r[0::2] * 256 + r[1::2]                      # 16 uniform bits
    sym = 15 - np.floor(np.log2(v + 1)).astype(np.int64)  # P(sym = k) ~ 2^-(k+1)
    return (sym.clip(0, 40) * 5 % 256).astype(np.uint8)


CASES = {
    "kat1_abracadabra": lambda: b"abracadabra",
    "kat2_a1000": lambda: b"a" * 1000,
    "kat3_allbytes512": lambda: bytes((i * 131) & 0xFF for i in range(131072)),
    "kat4_fib32": lambda: b"".join(bytes([i]) * c for i, c in enumerate(_fib(32))),
    "one_byte": lambda: b"Z",
    "two_symbols": lambda: b"ab" * 37 + b"a",
    "text_small": lambda: TEXT * 7,
    "tile_exact_4096": lambda: (TEXT * 20)[:4096],
    "tile_plus1_4097": lambda: (TEXT * 20)[:4097],
    "buffer_exact_131072": lambda: _lcg_bytes(131072, 7).tobytes(),
    "lcg_uniform_100k": lambda: _lcg_bytes(100000, 11).tobytes(),
    "skew_geometric_200k": lambda: _skew_small(200000, 5).tobytes(),
    "fib24_shuffled": lambda: _fib_shuffled(24, 3),
}

SMALL = {"kat1_abracadabra", "kat2_a1000", "one_byte", "two_symbols", "text_small", "tile_exact_4096", "tile_plus1_4097"}


def _fib_shuffled(k, seed):
    f = _fib(k)
    data = np.concatenate([np.full(c, i, dtype=np.uint8) for i, c in enumerate(f)])
    perm = np.argsort(_lcg_bytes(data.size * 4, seed).view(np.uint32), kind="stable")
    return data[perm].tobytes()


def make_input(name):
    return bytes(CASES[name]())
