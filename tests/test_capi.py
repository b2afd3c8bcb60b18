"""The shared library loads on a box without a GPU, exports every symbol the headers declare,
and refuses to search instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import pytest

import biorseo_b200 as b
from biorseo_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    return set(re.findall(r"\b(bs[sh]_[a-z_0-9]+)\s*\(", text))


def test_exports_every_declared_symbol():
    lib = capi.load()
    declared = _declared("biorseo_b200.h") | _declared("biorseo_b200_host.h")
    assert len(declared) >= 35
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared but not exported"
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    assert lib.bss_abi_version() == 1


def test_no_cpu_fallback_without_device():
    lib = capi.load()
    if lib.bss_device_count() > 0:
        pytest.skip("a GPU is present")
    host = b.HostLibrary.from_json_modules(synth.json_modules(10, 1, "c1"))
    with pytest.raises(b.SearchError) as err:
        host.to_device()
    assert err.value.code == capi.BSS_ERR_NO_DEVICE
    assert "no CPU path" in str(err.value)


def test_create_rejects_bad_tables():
    lib = capi.load()
    handle = C.c_void_p()
    assert lib.bss_library_create(None, -1, C.byref(handle)) == capi.BSS_ERR_INVALID
    t = capi.BssTable()
    t.n_modules, t.n_units, t.n_gates = 1, 1, 0
    assert lib.bss_library_create(C.byref(t), -1, C.byref(handle)) == capi.BSS_ERR_INVALID   # null arrays
    assert lib.bss_last_error()


def test_product_sources_do_not_touch_the_oracle():
    """The product path may not include, link or run anything under oracle/."""
    for folder in ("biorseo_b200", "include"):
        for dirpath, _, names in os.walk(os.path.join(ROOT, folder)):
            for name in names:
                if name.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    text = open(os.path.join(dirpath, name)).read()
                    assert "oracle/" not in text.replace("oracle/Makefile", "") or name == "build.py", os.path.join(dirpath, name)
