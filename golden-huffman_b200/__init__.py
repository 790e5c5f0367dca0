"""B200-native canonical Huffman codec: drop-in for the encode/decode hot path of chenghuige/golden-huffman.

Layout
  csrc/      hand-written sm_100a kernels + the C ABI (include/gh_codec.h) + host code construction
  host/      C++ adapters satisfying the reference's Compressor<_Encoder> / Decompressor<_Decoder> contract, CLI
  capi.py    ctypes binding of the C ABI
  codec.py   torch-side plumbing (device buffers, streams) over the C ABI
  sharded.py contiguous-slice sharding across GPUs with torch.distributed (NCCL)
  workloads.py synthetic inputs of BASELINE.json's shapes

PyTorch is used for device memory, streams and torch.distributed only; every byte of codec work happens in
lib/libgh_b200.so, and importing fails loudly if that library has not been built."""
from .capi import GhLib, GhCode, GhDeviceCode, GhError, GhShardSync, DEFAULT_LIB, SIGNATURES  # noqa: F401
from .codec import Codec  # noqa: F401

__all__ = ["GhLib", "GhCode", "GhDeviceCode", "GhError", "GhShardSync", "Codec", "DEFAULT_LIB", "SIGNATURES"]
