"""TEST INFRASTRUCTURE ONLY: loads the product's kernels compiled for the CPU under the pthread CUDA shim
(tests/emul). Used by the no-GPU tests to check kernel logic against the oracle; never by the product."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))


def load():
    import build_emul
    import golden_huffman_b200 as gh
    return gh.GhLib(build_emul.build())


def aligned(n, dtype=np.uint8, align=256):
    """numpy buffer standing in for device memory (the shim's cudaMemcpy is memcpy)"""
    item = np.dtype(dtype).itemsize
    raw = np.zeros(n * item + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * item].view(dtype)
