"""Runs one golden case on the GPU through the Python API against the debug build (tools/_build/libassert.so:
index assertions + a synchronisation after every kernel), and prints the assertion counters."""
import ctypes as C, os, sys, tempfile, pathlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from biorseo_b200 import capi
if os.environ.get("BSS_LIB"):
    capi.LIB_PATH = os.path.join(ROOT, "tools", "_build", "lib" + os.environ["BSS_LIB"] + ".so")
import numpy as np
import biorseo_b200 as b
import cases
name = sys.argv[1]
case = [c for c in cases.CASES if c["name"] == name][0]
with tempfile.TemporaryDirectory() as tmp:
    lib = cases.materialise(case, pathlib.Path(tmp))
    host = b.HostLibrary.load(case["source"], lib)
    dev = host.to_device()
    for kv in sys.argv[2:]:
        k, v = kv.split("=")
        dev.set_tuning(k, int(v))
    seq = case["seq"]
    try:
        sites = dev.search(seq, cases.mask_of(case["mask"], seq))
        got = sorted(host.records_text(sites.records()))
        print(name, "equal" if got == case["expected"] else "DIFFERENT", len(got), len(case["expected"]), dev.stats())
    except Exception as exc:
        print(name, "FAILED", exc)
    if os.environ.get("BSS_LIB"):
        l = capi.load()
        buf = (C.c_uint32 * 32)()
        l.bss_debug_assert_counts.argtypes = [C.POINTER(C.c_uint32)]
        print("assert rc", l.bss_debug_assert_counts(buf), "violations", list(buf))
