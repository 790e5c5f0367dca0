"""ctypes bindings for the TEST-ONLY CPU oracle (oracle/gh_oracle.c) and, when prebuilt, the compiled
unmodified reference (oracle/_ref/libglzip_ref.so, built from /root/reference by oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libgh_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libglzip_ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "glzip_ref")

NSYM = 257
EOF_SYM = 256


class GhCode(C.Structure):
    """Layout shared by gho_code (oracle) and gh_code (product, include/gh_codec.h)."""
    _fields_ = [
        ("length", C.c_uint32 * NSYM),
        ("codeword", C.c_uint32 * NSYM),
        ("symbol", C.c_uint32 * NSYM),
        ("min_len", C.c_uint32),
        ("max_len", C.c_uint32),
        ("start_pos", C.c_uint32 * 33),
        ("first_code", C.c_uint32 * 33),
    ]

    def as_dict(self):
        return {
            "length": list(self.length), "codeword": list(self.codeword), "symbol": list(self.symbol),
            "min_len": self.min_len, "max_len": self.max_len,
            "start_pos": list(self.start_pos)[: self.max_len + 1],
            "first_code": list(self.first_code)[: self.max_len + 1],
        }


def build_oracle():
    """Compile the C restatement (and the reference, if /root/reference is present). Building the checker
    is not using it."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = self.lib = C.CDLL(ORACLE_SO)
        u8p, u64, u64p = C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint64)
        L.gho_build_code.argtypes = [u64p, C.POINTER(GhCode)]
        L.gho_header_bytes.argtypes = [C.POINTER(GhCode)]
        L.gho_header_bytes.restype = C.c_size_t
        L.gho_write_header.argtypes = [C.POINTER(GhCode), u8p]
        L.gho_write_header.restype = C.c_size_t
        L.gho_parse_header.argtypes = [u8p, C.c_size_t, C.POINTER(GhCode)]
        L.gho_parse_header.restype = C.c_size_t
        L.gho_payload_bits.argtypes = [C.POINTER(GhCode), u64p]
        L.gho_payload_bits.restype = u64
        L.gho_encode_payload.argtypes = [u8p, u64, C.POINTER(GhCode), u8p, u64, u64p]
        L.gho_decode_payload.argtypes = [u8p, u64, C.POINTER(GhCode), u8p, u64, u64p]
        L.gho_compress_bound.argtypes = [u64]
        L.gho_compress_bound.restype = u64
        L.gho_compress.argtypes = [u8p, u64, u8p, u64, u64p]
        L.gho_decompress.argtypes = [u8p, u64, u8p, u64, u64p]

    @staticmethod
    def histogram(data):
        return np.bincount(np.frombuffer(data, dtype=np.uint8), minlength=256).astype(np.uint64)

    def build_code(self, hist256):
        h = np.ascontiguousarray(hist256, dtype=np.uint64)
        code = GhCode()
        rc = self.lib.gho_build_code(h.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(code))
        return rc, code

    def write_header(self, code):
        buf = np.zeros(self.lib.gho_header_bytes(C.byref(code)), dtype=np.uint8)
        n = self.lib.gho_write_header(C.byref(code), _u8p(buf))
        assert n == buf.size
        return buf.tobytes()

    def parse_header(self, blob):
        a = np.frombuffer(blob, dtype=np.uint8)
        code = GhCode()
        n = self.lib.gho_parse_header(_u8p(a), a.size, C.byref(code))
        return n, code

    def payload_bits(self, code, hist256):
        h = np.ascontiguousarray(hist256, dtype=np.uint64)
        return int(self.lib.gho_payload_bits(C.byref(code), h.ctypes.data_as(C.POINTER(C.c_uint64))))

    def encode_payload(self, data, code):
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        out = np.zeros(4 * a.size + 16, dtype=np.uint8)
        n = C.c_uint64(0)
        rc = self.lib.gho_encode_payload(_u8p(a), a.size, C.byref(code), _u8p(out), out.size, C.byref(n))
        return rc, out[: n.value].tobytes()

    def decode_payload(self, payload, code, cap):
        a = np.frombuffer(payload, dtype=np.uint8) if not isinstance(payload, np.ndarray) else payload
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        n = C.c_uint64(0)
        rc = self.lib.gho_decode_payload(_u8p(a), a.size, C.byref(code), _u8p(out), cap, C.byref(n))
        return rc, out[: n.value].tobytes()

    def compress(self, data):
        """bytes/ndarray -> (rc, .crs2 image bytes)"""
        a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        out = np.zeros(int(self.lib.gho_compress_bound(a.size)), dtype=np.uint8)
        n = C.c_uint64(0)
        rc = self.lib.gho_compress(_u8p(a), a.size, _u8p(out), out.size, C.byref(n))
        return rc, out[: n.value].tobytes()

    def decompress(self, blob, cap):
        a = np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        n = C.c_uint64(0)
        rc = self.lib.gho_decompress(_u8p(a), a.size, _u8p(out), cap, C.byref(n))
        return rc, out[: n.value].tobytes()


class Reference:
    """The unmodified reference (FILE*-based), driven through files in /dev/shm (or $TMPDIR)."""

    KINDS = {"simple": 0, "fast": 1, "table": 2}

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        L = self.lib = C.CDLL(REF_SO)
        L.ref_compress_file.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_decompress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_time_compress_file.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_time_compress_file.restype = C.c_double
        L.ref_time_decompress_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_time_decompress_file.restype = C.c_double
        self.tmp = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def _tmpfile(self, suffix):
        fd, p = tempfile.mkstemp(prefix="ghref_", suffix=suffix, dir=self.tmp)
        os.close(fd)
        return p

    def compress(self, data):
        pin, pout = self._tmpfile(".in"), self._tmpfile(".crs2")
        try:
            with open(pin, "wb") as f:
                f.write(bytes(data))
            self.lib.ref_compress_file(pin.encode(), pout.encode())
            with open(pout, "rb") as f:
                return f.read()
        finally:
            os.unlink(pin), os.unlink(pout)

    def decompress(self, blob, kind="simple"):
        pin, pout = self._tmpfile(".crs2"), self._tmpfile(".de")
        try:
            with open(pin, "wb") as f:
                f.write(bytes(blob))
            self.lib.ref_decompress_file(pin.encode(), pout.encode(), self.KINDS[kind])
            with open(pout, "rb") as f:
                return f.read()
        finally:
            os.unlink(pin), os.unlink(pout)

    def time_roundtrip(self, data, kind="table"):
        """-> (compress seconds, decompress seconds, compressed size); wall clock of the reference calls only"""
        pin, pout, pde = self._tmpfile(".in"), self._tmpfile(".crs2"), self._tmpfile(".de")
        try:
            with open(pin, "wb") as f:
                f.write(memoryview(data))
            tc = self.lib.ref_time_compress_file(pin.encode(), pout.encode())
            td = self.lib.ref_time_decompress_file(pout.encode(), pde.encode(), self.KINDS[kind])
            return tc, td, os.path.getsize(pout)
        finally:
            for p in (pin, pout, pde):
                if os.path.exists(p):
                    os.unlink(p)
