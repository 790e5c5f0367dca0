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


def test_streaming_small_chunks_through_cli(tmp_path, codec):
    """the adapters stream: with 1 MiB chunks and nothing resident a 40 MB file takes the multi-pass path (file read
    twice, 40 chunks whose payloads meet inside bytes); the .crs2 is the oracle's and decodes back, chunk by chunk"""
    import golden_huffman_b200.workloads as w
    from oracle_lib import Oracle
    cli = os.path.join(LIBDIR, "ghzip")
    data = w.zipf_np((40 << 20) + 4321, seed=31).tobytes()
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    subprocess.run([cli, str(src), "3", str(tmp_path / "s.crs2"), str(1 << 20), "0"], check=True)
    rc, want = Oracle().compress(data)
    assert rc == 0 and (tmp_path / "s.crs2").read_bytes() == want
    subprocess.run([cli, str(tmp_path / "s.crs2"), "4", str(tmp_path / "s.de"), str(1 << 20), "0"], check=True)
    assert (tmp_path / "s.de").read_bytes() == data
    # default geometry (64 MiB chunks, resident): same bytes
    subprocess.run([cli, str(src), "3", str(tmp_path / "d.crs2")], check=True)
    assert (tmp_path / "d.crs2").read_bytes() == want


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="compiled reference not shipped")
def test_file_beyond_4gib_matches_reference_binary(codec):
    """a 4.5 GiB file through the CLI built against the reference's unmodified compressor.h, streamed in 64 MiB chunks:
    the .crs2 is byte-identical to the one the compiled reference writes, and decodes back to the input"""
    import hashlib
    import tempfile
    import golden_huffman_b200.workloads as w
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    if shutil.disk_usage(base).free < (22 << 30):
        pytest.skip("needs ~22 GB of scratch space")
    cli = os.path.join(LIBDIR, "ghzip_refframe")
    if not os.path.exists(cli):
        cli = os.path.join(LIBDIR, "ghzip")
    d = tempfile.mkdtemp(prefix="gh_big_", dir=base)
    try:
        n = (9 << 29) + 777
        src = os.path.join(d, "big.bin")
        with open(src, "wb") as f:
            for a in range(0, n, 1 << 30):
                m = min(1 << 30, n - a)
                f.write(w.text_torch(m, "cuda", seed=100 + (a >> 30)).cpu().numpy().tobytes())
        subprocess.run([cli, src, "3", os.path.join(d, "gpu.crs2")], check=True)
        subprocess.run([REF_BIN, src, "3", os.path.join(d, "ref.crs2")], check=True)

        def sha(path):
            h = hashlib.sha256()
            with open(path, "rb") as f:
                for blk in iter(lambda: f.read(1 << 24), b""):
                    h.update(blk)
            return h.hexdigest()

        assert os.path.getsize(os.path.join(d, "gpu.crs2")) == os.path.getsize(os.path.join(d, "ref.crs2"))
        assert sha(os.path.join(d, "gpu.crs2")) == sha(os.path.join(d, "ref.crs2"))
        os.unlink(os.path.join(d, "ref.crs2"))
        subprocess.run([cli, os.path.join(d, "gpu.crs2"), "4", os.path.join(d, "gpu.de")], check=True)
        assert os.path.getsize(os.path.join(d, "gpu.de")) == n and sha(os.path.join(d, "gpu.de")) == sha(src)
    finally:
        shutil.rmtree(d, ignore_errors=True)
