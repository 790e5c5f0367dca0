"""TEST INFRASTRUCTURE ONLY: compiles the product's .cu sources with g++ against the pthread CUDA shim
(cuda_emul.h) into tests/emul/_build/libgh_emul.so so kernel logic can be checked on the CPU."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "golden-huffman_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libgh_emul.so")
CU = ["gh_runtime.cu", "gh_hist.cu", "gh_encode.cu", "gh_decode.cu", "gh_build.cu", "gh_stream.cu", "gh_multi.cu", "gh_api.cu"]


def build(force=False, sanitize=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in CU]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "cuda_emul.h"), os.path.join(HERE, "cuda_emul.cc"), os.path.join(ROOT, "include", "gh_codec.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-DGH_EMUL", "-Wall", "-Wno-unused-function",
           "-Wno-unknown-pragmas", "-include", os.path.join(HERE, "cuda_emul.h"),
           "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I" + HERE, "-o", OUT]
    if sanitize:
        cmd += ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
    for s in srcs:
        cmd += ["-x", "c++", s]
    cmd += ["-x", "c++", os.path.join(CSRC, "gh_host.cc"), os.path.join(HERE, "cuda_emul.cc")]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force=True, sanitize="--asan" in sys.argv))
