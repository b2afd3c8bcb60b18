"""Runs the GPU fuzz/golden tests against the -DBSS_DEBUG_ASSERT build and prints the violation counters."""
import ctypes as C, os, sys
ROOT = "/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from biorseo_b200 import capi
capi.LIB_PATH = os.path.join(ROOT, "scratch", "libassert.so")
import pytest
rc = pytest.main(["-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", os.path.join(ROOT, "tests", "test_gpu_parity.py")])
lib = capi.load()
buf = (C.c_uint32 * 32)()
lib.bss_debug_assert_counts.argtypes = [C.POINTER(C.c_uint32)]
print("pytest rc", rc, "assert rc", lib.bss_debug_assert_counts(buf), "violations", list(buf)[:12])
