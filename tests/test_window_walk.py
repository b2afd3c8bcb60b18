"""CPU emulation of the window pass against the flat-table interpreter (no GPU).

The warp-cooperative enumerator of window_kernel is host/device source (biorseo_b200/csrc/bss_window_walk.h);
tests/cpp/window_host.cpp compiles that very source with g++ and drives it like the kernel does (descriptors
from the product's builder, the 32 lanes as loops, count walk then emit walk). Here its record stream must equal the
interpreter's, bit for bit and in order, on the units the window walk is eligible for: JSON / DESC libraries,
random tables with row checks, overlapping chains, empty components, ends outside their component, narrow and
wide bands. The kernel around it is covered by the GPU parity tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import biorseo_b200 as b
from biorseo_b200 import capi, synth

import flat_interp
import tables
import util

ROOT = util.ROOT
SO = os.path.join(ROOT, "tests", "cpp", "_build", "libwindow_host.so")


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "cpp", "window_host.cpp"), os.path.join(ROOT, "biorseo_b200", "csrc", "bss_blob.cpp")]
    deps = srcs + [os.path.join(ROOT, "biorseo_b200", "csrc", h) for h in ("bss_window_walk.h", "bss_layout.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-shared", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "biorseo_b200", "csrc"), "-o", SO] + srcs, check=True)
    lib = C.CDLL(SO)
    lib.window_host_search.restype = C.c_int
    lib.window_host_search.argtypes = [C.POINTER(capi.BssTable), C.c_char_p, C.c_uint32, C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)),
                                       C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_int)]
    lib.window_host_free.argtypes = [C.POINTER(C.c_uint32)]
    return lib


def _raw(table):
    keep = {k: np.ascontiguousarray(table[k], dtype=dt) for k, dt in
            [("unit_module", np.uint32), ("unit_delay", np.uint32), ("unit_gate", np.uint32), ("unit_comp_begin", np.uint32),
             ("comp_pat_begin", np.uint32), ("pattern", np.uint8), ("unit_check_begin", np.uint32), ("checks", np.int16)]}
    t = capi.BssTable()
    t.n_modules, t.n_units, t.n_gates = int(table["n_modules"]), int(table["n_units"]), int(table["n_gates"])
    for name, arr in keep.items():
        ctype = {np.dtype(np.uint32): C.c_uint32, np.dtype(np.uint8): C.c_uint8, np.dtype(np.int16): C.c_int16}[arr.dtype]
        setattr(t, name, arr.ctypes.data_as(C.POINTER(ctype)) if arr.size else None)
    return t, keep


def window_search(lib, table, seq, mask):
    """(records of the window-walk units, unit_window flags, band)."""
    seq = seq.encode() if isinstance(seq, str) else bytes(seq)
    mask = np.ascontiguousarray(mask, dtype=np.uint32)
    t, keep = _raw(table)
    rec, n_words, band = C.POINTER(C.c_uint32)(), C.c_uint64(), C.c_int()
    flags = np.zeros(max(1, int(table["n_units"])), dtype=np.uint8)
    rc = lib.window_host_search(C.byref(t), seq, len(seq), mask.ctypes.data_as(C.c_void_p), C.byref(rec), C.byref(n_words),
                                flags.ctypes.data_as(C.c_void_p), C.byref(band))
    assert rc == 0, rc
    out = np.ctypeslib.as_array(rec, shape=(max(1, n_words.value),))[:n_words.value].copy()
    lib.window_host_free(rec)
    del keep
    return out, flags[:int(table["n_units"])].astype(bool), band.value


def expected_of_window_units(table, seq, mask, flags):
    """The interpreter's record stream restricted to the units the window walk took."""
    sub = dict(table)
    n_units = int(table["n_units"])
    gate = np.array(table["unit_gate"], dtype=np.uint32, copy=True)
    # units the window walk did not take get a gate that can never open: they emit nothing
    # (simplest way to keep unit numbering, module numbers and order intact)
    out = []
    for u in range(n_units):
        if not flags[u]:
            continue
        one = flat_interp.search_table(_single(table, u), seq, mask)
        out.append(one)
    del sub, gate
    return np.concatenate(out) if out else np.zeros(0, np.uint32)


def _single(table, u):
    """Flat table holding unit u alone (and its gate, if any)."""
    units = [u]
    g = int(table["unit_gate"][u])
    if g != 0xFFFFFFFF:
        units.append(g)
    ucb, cpb, uck = table["unit_comp_begin"], table["comp_pat_begin"], table["unit_check_begin"]
    pats, comp_begin, unit_comp_begin, unit_check_begin, checks = [], [0], [0], [0], []
    for x in units:
        for c in range(int(ucb[x]), int(ucb[x + 1])):
            pats.extend(table["pattern"][int(cpb[c]):int(cpb[c + 1])].tolist())
            comp_begin.append(len(pats))
        unit_comp_begin.append(len(comp_begin) - 1)
        checks.extend(table["checks"][int(uck[x]):int(uck[x + 1])].tolist())
        unit_check_begin.append(len(checks))
    return {
        "n_modules": int(table["n_modules"]), "n_units": 1, "n_gates": len(units) - 1,
        "unit_module": np.array([table["unit_module"][x] for x in units], np.uint32),
        "unit_delay": np.array([table["unit_delay"][x] for x in units], np.uint32),
        "unit_gate": np.array([1 if len(units) > 1 else 0xFFFFFFFF], np.uint32),
        "unit_comp_begin": np.array(unit_comp_begin, np.uint32), "comp_pat_begin": np.array(comp_begin, np.uint32),
        "pattern": np.array(pats, np.uint8) if pats else np.zeros(0, np.uint8), "unit_check_begin": np.array(unit_check_begin, np.uint32),
        "checks": np.array(checks, np.int16).reshape(-1, 4),
    }


def check(lib, table, seq, mask, min_window=1):
    got, flags, band = window_search(lib, table, seq, mask)
    assert int(flags.sum()) >= min_window, (int(flags.sum()), band)
    want = expected_of_window_units(table, seq, mask, flags)
    assert np.array_equal(got, want), (got.size, want.size, band)
    return int(flags.sum()), got.size


def low_complexity(rng, alphabet, n, runs=6):
    body = [alphabet[i] for i in rng.integers(0, len(alphabet), n)]
    for _ in range(runs):
        at, run = int(rng.integers(0, n)), int(rng.integers(3, 14))
        unit = [alphabet[i] for i in rng.integers(0, len(alphabet), int(rng.integers(1, 3)))]
        for j in range(run):
            if at + j < n:
                body[at + j] = unit[j % len(unit)]
    return bytes(body)


@pytest.mark.parametrize("source,n,seed", [("jsonfolder", 120, 1), ("jsonfolder", 300, 2), ("descfolder", 200, 3), ("descfolder", 90, 4),
                                           ("rinfolder", 110, 5)])
def test_libraries(harness, tmp_path, source, n, seed):
    seq = synth.sequence(n, 40 + seed)
    mask = synth.mask_canonical(seq, seed=seed, keep=0.6, span=60)
    if source == "jsonfolder":
        path = synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(300, seed, "c1") + [(f"x{i}", s, ss) for i, (_, s, ss) in enumerate(synth.json_modules(100, seed, "c4b"))])
    elif source == "descfolder":
        path = synth.write_desc_folder(str(tmp_path / "desc"), 300, seed)
    else:
        path = synth.write_rin_folder(str(tmp_path / "rin"), 200, seed, max_components=4, min_len=2)
    host = b.HostLibrary.load(source, path)
    table = host.table()
    n_window, words = check(harness, table, seq, mask, min_window=(0 if source == "rinfolder" else 50))
    if source == "jsonfolder":
        assert n_window == int(table["n_units"])   # every JSON loop module closes its strands by pairs
        assert words > 0


@pytest.mark.parametrize("seed", range(24))
def test_random_window_tables(harness, seed):
    rng = np.random.default_rng(4000 + seed)
    alphabet = [ord(c) for c in ("AG" if seed % 3 == 0 else "ACG" if seed % 3 == 1 else "ACGU")]
    n = int([37, 64, 97, 130, 200, 33, 260, 96][seed % 8])
    table = tables.window_table(rng, 60, alphabet, max_comp=4, max_len=4)
    seq = low_complexity(rng, alphabet, n)
    span = [12, 30, 5, 60, 255, 31, 32, 33][seed % 8]
    mask = tables.banded_random_mask(n, seed, span, [1.0, 0.6, 0.3][seed % 3])
    n_window, _ = check(harness, table, seq, mask, min_window=30)


@pytest.mark.parametrize("seed", range(6))
def test_generic_random_tables(harness, seed):
    """The generic random tables of the GPU tests: few units are eligible, those that are must be exact."""
    rng = np.random.default_rng(900 + seed)
    alphabet = [ord(c) for c in "ACG"]
    n = 150
    table = tables.random_table(rng, 200, alphabet, max_comp=3, max_len=4)
    seq = low_complexity(rng, alphabet, n)
    mask = tables.banded_random_mask(n, seed, 40, 0.5)
    check(harness, table, seq, mask, min_window=5)


def test_wide_bands(harness, monkeypatch):
    """An all-allowed mask on 400 nt (band 399): the walk takes it like any other band — queues of band + 1 entries, a stash
    of one band window per lane — and a harness-side limit (the kernel's is its shared-memory budget) leaves it alone."""
    rng = np.random.default_rng(1)
    table = tables.window_table(rng, 10, [ord(c) for c in "ACGU"], max_comp=3, max_len=4, min_len=2)
    seq = synth.sequence(400, 9)
    mask = synth.mask_all(400)
    got, flags, band = window_search(harness, table, seq, mask)
    assert band > 255 and flags.all()
    assert np.array_equal(got, expected_of_window_units(table, seq, mask, flags)) and got.size > 0
    monkeypatch.setenv("WINDOW_HOST_BAND_LIMIT", "255")
    got, flags, band = window_search(harness, table, seq, mask)
    assert band > 255 and not flags.any() and got.size == 0


def test_empty_mask(harness):
    rng = np.random.default_rng(2)
    table = tables.window_table(rng, 40, [ord(c) for c in "AC"], max_comp=3, max_len=3)
    seq = bytes(rng.choice([65, 67], 80).tolist())
    mask = np.zeros((80, 3), np.uint32)
    check(harness, table, seq, mask, min_window=20)   # band = -1: only units without any check can have sites


@pytest.mark.parametrize("seed", range(6))
def test_scratch_is_never_read_before_written_and_small_queues_give_the_same_stream(harness, monkeypatch, seed):
    """The walk's scratch (level parameters, cursors, queues, stash) is poisoned with two different values and the
    level queues are shrunk to the smallest legal size, which forces the "queue full: place the next level, come
    back" path on every dense item; the record stream may not change."""
    rng = np.random.default_rng(7000 + seed)
    alphabet = [ord(c) for c in ("AG" if seed % 2 else "ACG")]
    n = [150, 220, 96][seed % 3]
    table = tables.window_table(rng, 50, alphabet, max_comp=4, max_len=3)
    seq = low_complexity(rng, alphabet, n)
    mask = tables.banded_random_mask(n, seed, [40, 90, 20][seed % 3], [1.0, 0.8][seed % 2])
    base, flags, _ = window_search(harness, table, seq, mask)
    assert base.size > 0
    for poison, qcap in (("0xFFFFFFFF", "0"), ("0x5A5A5A5A", "1"), ("0", "1")):
        monkeypatch.setenv("WINDOW_HOST_POISON", poison)
        if qcap != "0":
            monkeypatch.setenv("WINDOW_HOST_QCAP", qcap)   # (clamped to max(32, band + 1) by the harness)
        got, flags2, _ = window_search(harness, table, seq, mask)
        assert np.array_equal(flags, flags2) and np.array_equal(got, base), (poison, qcap)
    want = expected_of_window_units(table, seq, mask, flags)
    assert np.array_equal(base, want)
