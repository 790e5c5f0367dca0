"""BASELINE.json config 1 (the reference's own 'utet' round trip, unit_tests/test.cc:247-280) reproduced with the
GPU adapters: compress a small text file through Compressor<GpuCanonicalHuffEncoder>, decompress through
Decompressor<GpuCanonicalHuffDecoder>, byte-compare; and cross-check both directions with the compiled
reference binary (oracle/_ref/glzip_ref, prebuilt in the build container) when it is present."""
import os
import shutil
import subprocess

import pytest

from golden_cases import TEXT, make_input

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "golden-huffman_b200", "lib")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "glzip_ref")


def _clis():
    out = [os.path.join(LIBDIR, "ghzip")]
    if os.path.exists(os.path.join(LIBDIR, "ghzip_refframe")):
        out.append(os.path.join(LIBDIR, "ghzip_refframe"))  # built against the reference's unmodified compressor.h
    return out


@pytest.mark.parametrize("cli", _clis())
def test_utet_roundtrip_through_adapters(tmp_path, cli, codec):
    src = tmp_path / "big.log"
    src.write_bytes(TEXT * 1500)  # a few hundred KB of English-like text
    subprocess.run([cli, str(src), "3"], check=True)
    crs = str(src) + ".crs2"  # empty outfile name -> in + ".crs2", as the reference does
    assert os.path.exists(crs)
    from oracle_lib import Oracle
    rc, want = Oracle().compress(src.read_bytes())
    assert rc == 0 and open(crs, "rb").read() == want
    for t in ("4", "5", "6"):
        subprocess.run([cli, crs, t], check=True)
        assert open(crs + ".de", "rb").read() == src.read_bytes()
        os.unlink(crs + ".de")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference not shipped")
@pytest.mark.parametrize("name", ["text_small", "kat3_allbytes512", "fib24_shuffled"])
def test_cross_decode_with_reference_binary(tmp_path, name, codec):
    cli = os.path.join(LIBDIR, "ghzip")
    data = make_input(name)
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    # GPU-compressed file decoded by the reference's three decoders
    subprocess.run([cli, str(src), "3", str(tmp_path / "gpu.crs2")], check=True)
    for t in ("4", "5", "6"):
        subprocess.run([REF_BIN, str(tmp_path / "gpu.crs2"), t, str(tmp_path / f"ref{t}.de")], check=True)
        assert (tmp_path / f"ref{t}.de").read_bytes() == data
    # reference-compressed file decoded by the GPU, and the two compressed files are identical
    shutil.copy(src, tmp_path / "in2.bin")
    subprocess.run([REF_BIN, str(tmp_path / "in2.bin"), "3", str(tmp_path / "ref.crs2")], check=True)
    assert (tmp_path / "ref.crs2").read_bytes() == (tmp_path / "gpu.crs2").read_bytes()
    subprocess.run([cli, str(tmp_path / "ref.crs2"), "4", str(tmp_path / "gpu.de")], check=True)
    assert (tmp_path / "gpu.de").read_bytes() == data


def test_cli_rejects_empty_input(tmp_path, codec):
    cli = os.path.join(LIBDIR, "ghzip")
    src = tmp_path / "empty"
    src.write_bytes(b"")
    r = subprocess.run([cli, str(src), "3"], capture_output=True, text=True)
    assert r.returncode == 1 and "empty input" in r.stderr
