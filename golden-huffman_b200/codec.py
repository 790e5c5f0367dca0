"""torch-side plumbing over the C ABI: allocates device buffers, passes raw pointers and the current stream.

Mirrors the reference's call order (include/compressor.h:62-73, 87-92):
  compress   = caculate_frequency -> gen_encode -> write_encode_info -> encode_file
  decompress = get_encode_info -> decode_file
No codec arithmetic happens in Python or torch."""
import ctypes as C

import numpy as np
import torch

from .capi import GhLib, GH_ERR_SPACE


def _on_device(fn):
    """every library call runs with the codec's device current (allocations, kernel attributes and launches are per
    device; a process may hold codecs on several)"""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        if self.device.type != "cuda":  # the tests' emulated codec keeps "device" memory in CPU tensors
            return fn(self, *a, **kw)
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapper


class Codec:
    def __init__(self, lib=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("golden_huffman_b200.Codec needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = lib or GhLib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            self.ctx = self.lib.ctx_create()
        self._ws = None

    def _stream(self):
        """raw cudaStream_t of torch's current stream: kernels are launched where torch would launch them"""
        return torch.cuda.current_stream().cuda_stream

    def _sync(self):
        torch.cuda.current_stream().synchronize()

    def close(self):
        if self.ctx:
            self.lib.ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- buffers -------------------------------------------------------------------------------------
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=self.device)
        return self._ws

    def _check_u8(self, x):
        assert x.dtype == torch.uint8 and x.device == self.device and x.is_contiguous(), \
            "expected a contiguous uint8 tensor on the codec's device"
        assert x.data_ptr() % 16 == 0, "the kernels read and write 128-bit vectors: the tensor's storage must start on a " \
                                       "16-byte boundary (a slice such as x[1:] does not; copy it first)"

    # ---- the path, step by step -------------------------------------------------------------------------
    @_on_device
    def histogram(self, x, out=None, accumulate=False):
        """K1: 256 byte counts (torch.int64 tensor on the device)."""
        self._check_u8(x)
        hist = out if out is not None else torch.empty(256, dtype=torch.int64, device=x.device)
        self.lib.histogram(x.data_ptr(), x.numel(), hist.data_ptr(), accumulate, self._stream())
        return hist

    def build_code(self, hist):
        """Host: code lengths + canonical codewords with the reference's tie-breaking."""
        h = hist.detach().cpu().numpy().astype(np.uint64) if isinstance(hist, torch.Tensor) else np.asarray(hist, np.uint64)
        return self.lib.build_code(h)

    @_on_device
    def encode(self, x, code, start_bit=0, append_eof=True, out=None):
        """K2-K4: returns (payload tensor, end_bit tensor[1] on the device)."""
        self._check_u8(x)
        n = x.numel()
        cap = self.lib.encode_payload_capacity(n, code, start_bit)
        payload = out if out is not None else torch.empty(cap, dtype=torch.uint8, device=x.device)
        assert payload.numel() >= cap
        ws_bytes = self.lib.encode_workspace_bytes(n)
        ws = self._workspace(ws_bytes)
        end_bit = torch.zeros(1, dtype=torch.int64, device=x.device)
        self.lib.encode(x.data_ptr(), n, code, payload.data_ptr(), payload.numel(), ws.data_ptr(), ws.numel(),
                        start_bit=start_bit, append_eof=append_eof, d_end_bit=end_bit.data_ptr(), stream=self._stream())
        return payload, end_bit

    @_on_device
    def decode(self, payload, nbytes, code, out_cap, out=None, allow=()):
        """K5-K7: returns (output tensor, symbols decoded, status)."""
        self._check_u8(payload)
        dst = out if out is not None else torch.empty(max(out_cap, 1), dtype=torch.uint8, device=payload.device)
        ws_bytes = self.lib.decode_workspace_bytes(nbytes)
        ws = self._workspace(ws_bytes)
        n, rc = self.lib.decode(payload.data_ptr(), nbytes, code, dst.data_ptr(), out_cap, ws.data_ptr(), ws.numel(),
                                self._stream(), allow=allow)
        return dst, n, rc

    # ---- whole .crs2 images -----------------------------------------------------------------------------
    @_on_device
    def compress(self, x, out=None):
        """device tensor -> device .crs2 image (header + payload), kernels only."""
        self._check_u8(x)
        cap = self.lib.compress_bound(x.numel()) if out is None else out.numel()
        img = out if out is not None else torch.empty(cap, dtype=torch.uint8, device=x.device)
        self._sync()  # the context runs on its own stream
        nbytes, _ = self.lib.compress_device(self.ctx, x.data_ptr(), x.numel(), img.data_ptr(), cap)
        return img[:nbytes]

    @_on_device
    def decompress(self, img, out_cap, out=None, allow=()):
        self._check_u8(img)
        dst = out if out is not None else torch.empty(max(out_cap, 1), dtype=torch.uint8, device=img.device)
        self._sync()
        n, rc = self.lib.decompress_device(self.ctx, img.data_ptr(), img.numel(), dst.data_ptr(), out_cap, allow=allow)
        return dst[: min(n, out_cap)], n, rc

    @_on_device
    def compress_host(self, src, dst):
        """host (ideally pinned) uint8 tensors/arrays in and out; H2D + kernels + D2H inside the call."""
        sp, sn = _host_ptr(src)
        dp, dn = _host_ptr(dst)
        n, _ = self.lib.compress_host(self.ctx, sp, sn, dp, dn)
        return n

    @_on_device
    def decompress_host(self, src, nbytes, dst, allow=()):
        sp, _ = _host_ptr(src)
        dp, dn = _host_ptr(dst)
        n, rc = self.lib.decompress_host(self.ctx, sp, nbytes, dp, dn, allow=allow)
        return n, rc


def _host_ptr(a):
    if isinstance(a, torch.Tensor):
        assert not a.is_cuda and a.dtype == torch.uint8 and a.is_contiguous()
        return a.data_ptr(), a.numel()
    a = np.asarray(a)
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data, a.size
