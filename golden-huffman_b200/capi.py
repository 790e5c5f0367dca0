"""ctypes binding of the C ABI in include/gh_codec.h (the same stub a reference-side binding would use).

Pointers are passed as plain integers (device pointers come from torch tensors' data_ptr(), host pointers from
numpy). There is no fallback: if the CUDA shared library is missing, loading fails loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "lib", "libgh_b200.so")

NSYM = 257
EOF_SYMBOL = 256

GH_OK, GH_ERR_EMPTY, GH_ERR_TOO_LONG, GH_ERR_SPACE, GH_ERR_FORMAT, GH_ERR_NO_EOF, GH_ERR_ARG, GH_ERR_CUDA = range(8)


class GhCode(C.Structure):
    """struct gh_code (include/gh_codec.h) == the tables of reference include/canonical_huff_encoder.h:107-120"""
    _fields_ = [
        ("length", C.c_uint32 * NSYM),
        ("codeword", C.c_uint32 * NSYM),
        ("symbol", C.c_uint32 * NSYM),
        ("min_len", C.c_uint32),
        ("max_len", C.c_uint32),
        ("start_pos", C.c_uint32 * 33),
        ("first_code", C.c_uint32 * 33),
    ]


class GhEncodeTable(C.Structure):
    _fields_ = [("codeword", C.c_uint32 * NSYM), ("length", C.c_uint8 * (NSYM + 3))]


class GhDeviceCode(C.Structure):
    """struct gh_device_code: what gh_build_code_device leaves in device memory"""
    _fields_ = [
        ("code", GhCode),
        ("status", C.c_uint32),
        ("header_bytes", C.c_uint32),
        ("reserved", C.c_uint32),
        ("payload_bits", C.c_uint64),
        ("total_symbols", C.c_uint64),
        ("table", GhEncodeTable),
    ]


class GhShardSync(C.Structure):
    _fields_ = [
        ("n_symbols", C.c_uint64),
        ("exit_bit", C.c_uint32),
        ("eof_found", C.c_uint32),
        ("rounds", C.c_uint32),
        ("sub_bytes", C.c_uint32),
    ]


class GhError(RuntimeError):
    def __init__(self, status, what, detail=""):
        self.status = status
        super().__init__(f"{what}: status {status} ({detail})")


# every symbol include/gh_codec.h declares: name -> (restype, argtypes)
_VP, _U64, _U32, _SZ, _INT = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t, C.c_int
_CODEP = C.POINTER(GhCode)
SIGNATURES = {
    "gh_strerror": (C.c_char_p, [_INT]),
    "gh_last_cuda_error": (_INT, []),
    "gh_launch_count": (_U64, []),
    "gh_profile_enable": (None, [_INT]),
    "gh_profile_fetch": (_SZ, [C.c_char_p, _SZ]),
    "gh_debug_select_writer": (None, [_INT]),
    "gh_debug_disable_phase_walk": (None, [_INT]),
    "gh_ctx_set_stream": (_INT, [_VP, _VP]),
    "gh_ctx_set_device_code": (_INT, [_VP, _INT]),
    "gh_ctx_set_host_chunk": (_INT, [_VP, _U64]),
    "gh_build_code_device": (_INT, [_VP, _INT, _VP, _VP, _VP]),
    "gh_build_code": (_INT, [_VP, _CODEP]),
    "gh_header_bytes": (_SZ, [_CODEP]),
    "gh_write_header": (_INT, [_CODEP, _VP, _SZ, C.POINTER(_SZ)]),
    "gh_parse_header": (_INT, [_VP, _SZ, _CODEP, C.POINTER(_SZ)]),
    "gh_payload_bits": (_U64, [_CODEP, _VP, _INT]),
    "gh_histogram": (_INT, [_VP, _U64, _VP, _INT, _VP]),
    "gh_encode_workspace_bytes": (_SZ, [_U64]),
    "gh_encode_payload_capacity": (_U64, [_U64, _CODEP, _U64]),
    "gh_encode": (_INT, [_VP, _U64, _CODEP, _U64, _INT, _VP, _U64, _VP, _VP, _SZ, _VP]),
    "gh_decode_workspace_bytes": (_SZ, [_U64]),
    "gh_decode": (_INT, [_VP, _U64, _CODEP, _VP, _U64, C.POINTER(_U64), _VP, _SZ, _VP]),
    "gh_decode_sync": (_INT, [_VP, _U64, _U64, _CODEP, _U32, _INT, C.POINTER(GhShardSync), _VP, _SZ, _VP]),
    "gh_decode_write": (_INT, [_VP, _U64, _U64, _CODEP, _VP, _U64, _VP, _SZ, _VP]),
    "gh_ctx_create": (_INT, [C.POINTER(_VP)]),
    "gh_ctx_destroy": (None, [_VP]),
    "gh_compress_bound": (_U64, [_U64]),
    "gh_compress_host": (_INT, [_VP, _VP, _U64, _VP, _U64, C.POINTER(_U64)]),
    "gh_decompress_host": (_INT, [_VP, _VP, _U64, _VP, _U64, C.POINTER(_U64)]),
    "gh_compress_device": (_INT, [_VP, _VP, _U64, _VP, _U64, C.POINTER(_U64)]),
    "gh_decompress_device": (_INT, [_VP, _VP, _U64, _VP, _U64, C.POINTER(_U64)]),
    "gh_stream_create": (_INT, [C.POINTER(_VP), _U64, _U64]),
    "gh_stream_destroy": (None, [_VP]),
    "gh_stream_histogram": (_INT, [_VP, _VP, _VP]),
    "gh_stream_encode": (_INT, [_VP, _VP, _VP, _CODEP, C.POINTER(_U64)]),
    "gh_stream_decode": (_INT, [_VP, _VP, _VP, _CODEP, _U64, C.POINTER(_U64)]),
    "gh_compress_host_multi": (_INT, [_INT, _VP, _VP, _U64, _VP, _U64, C.POINTER(_U64)]),
    "gh_decompress_host_multi": (_INT, [_INT, _VP, _VP, _U64, _VP, _U64, C.POINTER(_U64)]),
    "gh_stage_input": (_INT, [_VP, _VP, _U64, _VP]),
    "gh_encode_staged": (_INT, [_VP, _CODEP, _VP, _U64, C.POINTER(_U64)]),
    "gh_stage_payload": (_INT, [_VP, _VP, _U64, _CODEP, C.POINTER(_U64)]),
    "gh_decode_staged": (_INT, [_VP, _VP, _U64]),
}


class GhLib:
    """Thin, explicit wrapper: one method per C entry point, raising GhError on non-zero status."""

    def __init__(self, path=None):
        path = path or DEFAULT_LIB
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it with `python golden-huffman_b200/build.py` "
                "(nvcc, sm_100a). This codec has no CPU fallback.")
        self.path = path
        self.lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, name)  # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args

    # -- helpers ---------------------------------------------------------------------------------------
    def strerror(self, status):
        return self.lib.gh_strerror(status).decode()

    def check(self, status, what):
        if status != GH_OK:
            detail = self.strerror(status)
            if status == GH_ERR_CUDA:
                detail += f"; cudaError {self.lib.gh_last_cuda_error()}"
            raise GhError(status, what, detail)

    def launch_count(self):
        return int(self.lib.gh_launch_count())

    def profile_enable(self, on=True):
        self.lib.gh_profile_enable(1 if on else 0)

    def profile_fetch(self):
        """-> {kernel name: (launches, total ms)} since the last fetch"""
        buf = C.create_string_buffer(8192)
        n = self.lib.gh_profile_fetch(buf, 8192)
        out = {}
        for line in buf.raw[:n].decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    def ctx_set_stream(self, ctx, stream):
        self.check(self.lib.gh_ctx_set_stream(ctx, stream), "gh_ctx_set_stream")

    # -- host side ------------------------------------------------------------------------------------
    def build_code(self, hist256):
        import numpy as np
        h = np.ascontiguousarray(hist256, dtype=np.uint64)
        assert h.size == 256
        code = GhCode()
        self.check(self.lib.gh_build_code(h.ctypes.data, C.byref(code)), "gh_build_code")
        return code

    def build_code_device(self, d_hists, n_hists, d_code, d_header=0, stream=0):
        """device pointers: n_hists x 256 u64 counters in, struct gh_device_code (and optionally the header) out"""
        self.check(self.lib.gh_build_code_device(d_hists, n_hists, d_code, d_header, stream), "gh_build_code_device")

    def ctx_set_host_chunk(self, ctx, nbytes=0):
        self.check(self.lib.gh_ctx_set_host_chunk(ctx, nbytes), "gh_ctx_set_host_chunk")

    def ctx_set_device_code(self, ctx, on=True):
        self.check(self.lib.gh_ctx_set_device_code(ctx, 1 if on else 0), "gh_ctx_set_device_code")

    def header_bytes(self, code):
        return int(self.lib.gh_header_bytes(C.byref(code)))

    def write_header(self, code):
        import numpy as np
        buf = np.zeros(self.header_bytes(code), dtype=np.uint8)
        n = C.c_size_t(0)
        self.check(self.lib.gh_write_header(C.byref(code), buf.ctypes.data, buf.size, C.byref(n)), "gh_write_header")
        return buf[: n.value].tobytes()

    def parse_header(self, blob):
        import numpy as np
        a = np.frombuffer(bytes(blob), dtype=np.uint8)
        code, n = GhCode(), C.c_size_t(0)
        self.check(self.lib.gh_parse_header(a.ctypes.data, a.size, C.byref(code), C.byref(n)), "gh_parse_header")
        return code, int(n.value)

    def payload_bits(self, code, hist256, with_eof=True):
        import numpy as np
        h = np.ascontiguousarray(hist256, dtype=np.uint64)
        return int(self.lib.gh_payload_bits(C.byref(code), h.ctypes.data, 1 if with_eof else 0))

    def compress_bound(self, n):
        return int(self.lib.gh_compress_bound(n))

    # -- device side (raw pointers) -------------------------------------------------------------------
    def histogram(self, d_in, n, d_hist, accumulate=False, stream=0):
        self.check(self.lib.gh_histogram(d_in, n, d_hist, 1 if accumulate else 0, stream), "gh_histogram")

    def encode_workspace_bytes(self, n):
        return int(self.lib.gh_encode_workspace_bytes(n))

    def encode_payload_capacity(self, n, code, start_bit=0):
        return int(self.lib.gh_encode_payload_capacity(n, C.byref(code), start_bit))

    def encode(self, d_in, n, code, d_payload, payload_cap, d_ws, ws_bytes, start_bit=0, append_eof=True,
               d_end_bit=0, stream=0):
        self.check(self.lib.gh_encode(d_in, n, C.byref(code), start_bit, 1 if append_eof else 0, d_payload,
                                      payload_cap, d_end_bit, d_ws, ws_bytes, stream), "gh_encode")

    def decode_workspace_bytes(self, payload_bytes):
        return int(self.lib.gh_decode_workspace_bytes(payload_bytes))

    def decode(self, d_payload, nbytes, code, d_out, out_cap, d_ws, ws_bytes, stream=0, allow=()):
        n = C.c_uint64(0)
        rc = self.lib.gh_decode(d_payload, nbytes, C.byref(code), d_out, out_cap, C.byref(n), d_ws, ws_bytes, stream)
        if rc not in allow:
            self.check(rc, "gh_decode")
        return int(n.value), rc

    def decode_sync(self, d_payload, slice_bytes, readable, code, entry_bit, first_call, d_ws, ws_bytes, stream=0):
        res = GhShardSync()
        self.check(self.lib.gh_decode_sync(d_payload, slice_bytes, readable, C.byref(code), entry_bit,
                                           1 if first_call else 0, C.byref(res), d_ws, ws_bytes, stream),
                   "gh_decode_sync")
        return res

    def decode_write(self, d_payload, slice_bytes, readable, code, d_out, out_cap, d_ws, ws_bytes, stream=0):
        self.check(self.lib.gh_decode_write(d_payload, slice_bytes, readable, C.byref(code), d_out, out_cap, d_ws,
                                            ws_bytes, stream), "gh_decode_write")

    # -- whole images ------------------------------------------------------------------------------------
    def ctx_create(self):
        ctx = C.c_void_p(0)
        self.check(self.lib.gh_ctx_create(C.byref(ctx)), "gh_ctx_create")
        return ctx

    def ctx_destroy(self, ctx):
        self.lib.gh_ctx_destroy(ctx)

    def _image_call(self, fn, name, ctx, src, n, dst, cap, allow=()):
        out = C.c_uint64(0)
        rc = fn(ctx, src, n, dst, cap, C.byref(out))
        if rc not in allow:
            self.check(rc, name)
        return int(out.value), rc

    def compress_host(self, ctx, src, n, dst, cap, allow=()):
        return self._image_call(self.lib.gh_compress_host, "gh_compress_host", ctx, src, n, dst, cap, allow)

    def decompress_host(self, ctx, src, n, dst, cap, allow=()):
        return self._image_call(self.lib.gh_decompress_host, "gh_decompress_host", ctx, src, n, dst, cap, allow)

    def compress_host_multi(self, devices, src, n, dst, cap, allow=()):
        """devices: list of device ordinals, one per shard (a device may repeat)"""
        arr = (C.c_int * len(devices))(*devices)
        out = C.c_uint64(0)
        rc = self.lib.gh_compress_host_multi(len(devices), C.cast(arr, C.c_void_p), src, n, dst, cap, C.byref(out))
        if rc not in allow:
            self.check(rc, "gh_compress_host_multi")
        return int(out.value), rc

    def decompress_host_multi(self, devices, src, n, dst, cap, allow=()):
        arr = (C.c_int * len(devices))(*devices)
        out = C.c_uint64(0)
        rc = self.lib.gh_decompress_host_multi(len(devices), C.cast(arr, C.c_void_p), src, n, dst, cap, C.byref(out))
        if rc not in allow:
            self.check(rc, "gh_decompress_host_multi")
        return int(out.value), rc

    def compress_device(self, ctx, src, n, dst, cap, allow=()):
        return self._image_call(self.lib.gh_compress_device, "gh_compress_device", ctx, src, n, dst, cap, allow)

    def decompress_device(self, ctx, src, n, dst, cap, allow=()):
        return self._image_call(self.lib.gh_decompress_device, "gh_decompress_device", ctx, src, n, dst, cap, allow)
