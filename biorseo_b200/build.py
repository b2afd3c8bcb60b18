"""Builds the shared library of the insertion-site search for sm_100a, in-tree.

    biorseo_b200/lib/libbiorseo_b200.so   CUDA kernels + C ABI (csrc/) + C++ host glue (host/)

nvcc cross-compiles without a GPU. The oracle (test infrastructure) is built separately by
oracle/Makefile; nothing of it is linked here.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libbiorseo_b200.so")

CUDA_SOURCES = ["csrc/bss_kernels.cu", "csrc/bss_mask.cu", "csrc/bss_index.cu", "csrc/bss_probe.cu", "csrc/bss_exchange.cu", "csrc/bss_sharded.cu", "csrc/bss_capi.cu", "csrc/bss_blob.cpp"]
HOST_SOURCES = ["host/Motif.cpp", "host/Pool.cpp", "host/mini_json.cpp", "host/module_library.cpp",
                "host/site_search.cpp", "host/fasta.cpp", "host/bsh_capi.cpp"]
HEADERS = ["csrc/bss_device.cuh", "csrc/bss_layout.h", "csrc/bss_window.cuh", "csrc/bss_window_walk.h", "host/Motif.h", "host/Pool.h", "host/mini_json.h", "host/module_library.h",
           "host/site_search.h", "host/fasta.h", "../include/biorseo_b200.h", "../include/biorseo_b200_host.h"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in CUDA_SOURCES + HOST_SOURCES]
    deps = srcs + [os.path.normpath(os.path.join(HERE, h)) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not _stale(LIB_PATH, deps):
        return LIB_PATH
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-Xptxas", "-v" if verbose else "-O3",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "host"), "-I", os.path.join(HERE, "csrc"),
           "-o", LIB_PATH] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
