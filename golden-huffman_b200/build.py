"""Builds the product shared library  golden-huffman_b200/lib/libgh_b200.so  (sm_100a only) and the host CLI.

nvcc cross-compiles without a GPU. The library is built IN-TREE so it travels to the GPU box with the repo
snapshot; it is git-ignored (history stays source-only)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgh_b200.so")
CLI = os.path.join(LIBDIR, "ghzip")

CU_SOURCES = ["gh_runtime.cu", "gh_hist.cu", "gh_encode.cu", "gh_decode.cu", "gh_build.cu", "gh_stream.cu", "gh_multi.cu", "gh_api.cu"]
CC_SOURCES = ["gh_host.cc"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_variant(name, defines):
    """tuning builds: lib/libgh_b200_<name>.so with -D overrides of the kernels' compile-time parameters
    (select with GH_LIB_PATH in bench.py); not used by the product or the tests"""
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in CU_SOURCES + CC_SOURCES]
    out = os.path.join(LIBDIR, f"libgh_b200_{name}.so")
    cmd = [_nvcc(), "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"),
           "-I" + CSRC, "-shared", "-cudart", "static", *["-D" + d for d in defines], "-o", out, *srcs]
    subprocess.run(cmd, check=True)
    return out


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in CU_SOURCES + CC_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "gh_codec.h"))
    if force or _newer(LIB, deps):
        cmd = [_nvcc(), "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-Wall,-Wextra",
               "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-shared", "-cudart", "static",
               "-o", LIB, *srcs]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    cli_src = os.path.join(HOST, "ghzip.cc")
    if os.path.exists(cli_src):
        cli_deps = [cli_src, LIB] + [os.path.join(HOST, f) for f in os.listdir(HOST)]
        if force or _newer(CLI, cli_deps):
            cxx = shutil.which("g++") or "g++"
            subprocess.run([cxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
                            "-I" + HOST, "-o", CLI, cli_src, "-L" + LIBDIR, "-lgh_b200",
                            "-Wl,-rpath,$ORIGIN"], check=True)
        # the same CLI compiled against the reference's UNMODIFIED compressor.h (build container only): the
        # binary travels to the GPU box, where tests/test_gpu_cli.py runs it
        ref_inc = "/root/reference/include"
        if os.path.exists(os.path.join(ref_inc, "compressor.h")):
            cxx = shutil.which("g++") or "g++"
            subprocess.run([cxx, "-O2", "-std=c++17", "-w", "-DGH_USE_REFERENCE_FRAME", "-I" + ref_inc,
                            "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-o", CLI + "_refframe", cli_src,
                            "-L" + LIBDIR, "-lgh_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
