"""Synthetic inputs of the shapes BASELINE.json names. Seeds are fixed.

numpy versions for CPU-sized cases, torch-on-device versions for the GiB-sized ones (generated in chunks so
the temporaries stay small). Data generation is not part of any timed region."""
import numpy as np

ZIPF_S = 1.1
SEED = 1234

ENGLISH = (" etaoinshrdlcumwfgypbvkjxqz" ".,;:'\"!?-\n" "ETAOINSHRDLCUMWFGYPBVKJXQZ" "0123456789")


def _zipf_pmf(s=ZIPF_S, k=256):
    p = 1.0 / np.arange(1, k + 1, dtype=np.float64) ** s
    return p / p.sum()


def _rank_to_byte():
    """the Zipf rank -> byte value permutation: part of the DISTRIBUTION, so it is fixed (SEED) whatever seed draws
    the samples -- the shards of a multi-GPU run are then samples of one and the same distribution"""
    return np.random.default_rng(SEED).permutation(256).astype(np.uint8)


def _english_pmf():
    """order-0 letter/space/punctuation mix tuned to about 4.5 bits/symbol"""
    p = np.zeros(256, dtype=np.float64)
    letters = "etaoinshrdlcumwfgypbvkjxqz"
    freq = [12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2, 2.0, 2.0, 1.9, 1.5, 1.0,
            0.8, 0.15, 0.15, 0.1, 0.07]
    for ch, f in zip(letters, freq):
        p[ord(ch)] = f
        p[ord(ch.upper())] = f * 0.04
    p[ord(" ")] = 19.0
    for ch, f in zip(".,;:'\"!?-\n", [1.1, 1.2, 0.1, 0.1, 0.4, 0.3, 0.1, 0.1, 0.3, 1.5]):
        p[ord(ch)] = f
    for ch in "0123456789":
        p[ord(ch)] = 0.12
    return p / p.sum()


def entropy_bits(pmf):
    q = pmf[pmf > 0]
    return float(-(q * np.log2(q)).sum())


# ---- numpy ------------------------------------------------------------------------------------------
def _sample_pmf_np(n, pmf, table, seed, chunk=1 << 24):
    rng = np.random.default_rng(seed)
    cdf = np.cumsum(pmf)
    cdf[-1] = 1.0
    out = np.empty(n, dtype=np.uint8)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        idx = np.searchsorted(cdf, rng.random(m, dtype=np.float32), side="left")
        out[s:s + m] = table[np.minimum(idx, len(pmf) - 1)]
    return out


def zipf_np(n, seed=SEED):
    return _sample_pmf_np(n, _zipf_pmf(), _rank_to_byte(), seed)


def uniform_np(n, seed=SEED):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8)


def text_np(n, seed=SEED):
    return _sample_pmf_np(n, _english_pmf(), np.arange(256, dtype=np.uint8), seed)


def fib_counts(k=31):
    f = [1, 2]
    while len(f) < k:
        f.append(f[-1] + f[-2])
    return f[:k]


def shard_fib_counts(rank=0, world=1, k=31):
    """this shard's share of the Fibonacci counts: the counts of the WHOLE input are Fibonacci(i) whatever the
    number of shards (symbol i's occurrences are dealt round-robin, starting at rank i % world)"""
    return [c // world + (1 if (rank - i) % world < c % world else 0) for i, c in enumerate(fib_counts(k))]


def skewed_np(n, seed=SEED, k=31, rank=0, world=1):
    """bytes 0..k-1 occur Fibonacci(i) times, byte k fills the rest -> k+2 symbols with the end mark, and for
    k = 31 a maximum code length of exactly 32; positions are shuffled.
    With world > 1 this is shard `rank` of such an input of world * n bytes."""
    f = shard_fib_counts(rank, world, k)
    total = sum(f)
    assert n > total, f"need n > {total}"
    out = np.full(n, k, dtype=np.uint8)
    pos = np.random.default_rng(seed).choice(n, size=total, replace=False)
    start = 0
    for sym, c in enumerate(f):
        out[pos[start:start + c]] = sym
        start += c
    return out


# ---- torch, on device ---------------------------------------------------------------------------------
def _sample_pmf_torch(n, pmf, table, device, seed, chunk=1 << 26):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cdf = torch.tensor(np.cumsum(pmf), dtype=torch.float32, device=device)
    cdf[-1] = 1.0
    tab = torch.tensor(table, dtype=torch.uint8, device=device)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        u = torch.rand(m, device=device, generator=g)
        idx = torch.searchsorted(cdf, u).clamp_(max=len(pmf) - 1)
        out[s:s + m] = tab[idx]
    return out


def zipf_torch(n, device, seed=SEED):
    return _sample_pmf_torch(n, _zipf_pmf(), _rank_to_byte(), device, seed)


def uniform_torch(n, device, seed=SEED):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    chunk = 1 << 28
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        out[s:s + m] = torch.randint(0, 256, (m,), dtype=torch.uint8, device=device, generator=g)
    return out


def text_torch(n, device, seed=SEED):
    return _sample_pmf_torch(n, _english_pmf(), np.arange(256, dtype=np.uint8), device, seed)


def skewed_torch(n, device, seed=SEED, k=31, rank=0, world=1):
    import torch
    f = shard_fib_counts(rank, world, k)
    total = sum(f)
    assert n > total
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.full((n,), k, dtype=torch.uint8, device=device)
    # distinct positions: a random odd-stride walk modulo n is a permutation prefix
    stride = (int(torch.randint(1, 1 << 30, (1,), generator=g, device=device).item()) * 2 + 1)
    while np.gcd(stride, n) != 1:
        stride += 2
    idx = (torch.arange(total, device=device, dtype=torch.int64) * stride + 12345) % n
    vals = torch.repeat_interleave(torch.arange(k, device=device, dtype=torch.uint8),
                                   torch.tensor(f, device=device))
    out[idx] = vals
    return out


WORKLOADS_NP = {"zipf": zipf_np, "uniform": uniform_np, "text": text_np, "skewed": skewed_np}
WORKLOADS_TORCH = {"zipf": zipf_torch, "uniform": uniform_torch, "text": text_torch, "skewed": skewed_torch}
