"""Runs the GPU fuzz/golden tests against the -DBSS_DEBUG_ASSERT build and prints the violation counters."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from biorseo_b200 import build, capi
if len(sys.argv) > 1 and sys.argv[1] == "build":     # here (no GPU needed): python tools/run_assert.py build
    os.makedirs(os.path.join(ROOT, "tools", "_build"), exist_ok=True)
    srcs = [os.path.join(build.HERE, s) for s in build.CUDA_SOURCES + build.HOST_SOURCES]
    subprocess.run([build._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-DBSS_DEBUG_ASSERT", "-DBSS_DEBUG_SYNC",
                    "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(build.HERE, "host"), "-I", os.path.join(build.HERE, "csrc"),
                    "-o", os.path.join(ROOT, "tools", "_build", "libassert.so")] + srcs, check=True)
    sys.exit(0)
capi.LIB_PATH = os.path.join(ROOT, "tools", "_build", "libassert.so")
import pytest
rc = pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_gpu_parity.py")])
lib = capi.load()
buf = (C.c_uint32 * 32)()
lib.bss_debug_assert_counts.argtypes = [C.POINTER(C.c_uint32)]
print("pytest rc", rc, "assert rc", lib.bss_debug_assert_counts(buf), "violations", list(buf)[:12])
