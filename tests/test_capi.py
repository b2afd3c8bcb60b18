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
    assert lib.bss_abi_version() == 2


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


def test_new_entry_points_validate_their_arguments_before_touching_a_device():
    """Index build, exchange and probe: malformed calls are refused with a status code and a message (no crash, no CUDA)."""
    lib = capi.load()
    out = C.c_void_p()
    assert lib.bss_site_index_build(None, None, 0, None, 0, 1, C.byref(out)) == capi.BSS_ERR_INVALID and not out.value
    assert lib.bss_site_index_count(None, 0) == 0 and not lib.bss_site_index_host(None, 0) and not lib.bss_site_index_device(None, 0)
    lib.bss_site_index_free(None)
    v = C.c_double()
    assert lib.bss_measure_int_issue(None, C.byref(v)) == capi.BSS_ERR_INVALID
    for world, rank, cap, header in ((0, 0, 64, 4), (17, 0, 64, 4), (2, 2, 64, 4), (2, 0, 64, 3), (2, 0, 64, 0), (2, 0, 1 << 41, 4)):
        rc = lib.bss_exchange_create(0, world, rank, cap, header, C.byref(out))
        assert rc in (capi.BSS_ERR_INVALID, capi.BSS_ERR_LIMIT) and not out.value, (world, rank, cap, header)
        assert lib.bss_last_error()
    assert lib.bss_exchange_run_async(None, None, None, None) == capi.BSS_ERR_INVALID
    assert lib.bss_exchange_handle(None, None) == capi.BSS_ERR_INVALID and lib.bss_exchange_open(None, None) == capi.BSS_ERR_INVALID
    assert not lib.bss_exchange_block(None) and lib.bss_exchange_capacity(None) == 0
    lib.bss_exchange_destroy(None)
    text, length = C.c_void_p(), C.c_uint64()
    assert lib.bsh_fasta_text(None, C.byref(text), C.byref(length)) != 0
