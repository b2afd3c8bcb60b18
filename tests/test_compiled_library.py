"""Compiled library file (row "next" #3): everything the loaders produce in one binary file.
Round trips for the three sources (flat table, reject log, module ids and link recipes equal),
corrupt files are refused, and loading is much faster than parsing."""
import os
import time

import numpy as np
import pytest

import biorseo_b200 as b
from biorseo_b200 import synth

import cases


def _same(a, c):
    ta, tc = a.table(), c.table()
    assert set(ta) == set(tc)
    for key in ta:
        assert np.array_equal(np.asarray(ta[key]), np.asarray(tc[key])), key
    assert a.kind == c.kind and a.delay == c.delay and a.n_modules == c.n_modules and a.n_seen == c.n_seen
    assert a.rejected() == c.rejected()
    assert [a.module_id(m) for m in range(a.n_modules)] == [c.module_id(m) for m in range(c.n_modules)]
    # the link recipes: reconstruct a site of every module at arbitrary starts
    comps = a.module_components()
    for m in range(a.n_modules):
        rec = np.array([m] + [10 + 20 * j for j in range(int(comps[m]))], dtype=np.uint32)
        assert a.records_text(rec) == c.records_text(rec)


@pytest.mark.parametrize("source", ["jsonfolder", "rinfolder", "descfolder"])
def test_round_trip(tmp_path, source):
    if source == "jsonfolder":
        path = synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(300, 3, "c1") + synth.json_fuzz_modules(120, 4))
    elif source == "rinfolder":
        path = synth.write_rin_folder(str(tmp_path / "rin"), 150, 5)
    else:
        path = synth.write_desc_folder(str(tmp_path / "desc"), 150, 6)
    host = b.HostLibrary.load(source, path)
    assert host.n_modules > 0 and len(host.rejected()) > 0
    out = host.save_compiled(str(tmp_path / "lib.bshlib"))
    again = b.HostLibrary.load_compiled(out)
    _same(host, again)
    again.save_compiled(str(tmp_path / "again.bshlib"))
    assert open(out, "rb").read() == open(str(tmp_path / "again.bshlib"), "rb").read()


@pytest.mark.parametrize("case", [c for c in cases.CASES][:6], ids=[c["name"] for c in cases.CASES][:6])
def test_golden_libraries_round_trip(case, tmp_path):
    host = b.HostLibrary.load(case["source"], cases.materialise(case, tmp_path))
    _same(host, b.HostLibrary.load_compiled(host.save_compiled(str(tmp_path / "g.bshlib"))))


def test_the_table_of_a_loaded_library_lives_in_the_mapping(tmp_path):
    """load_compiled maps the file: the flat table's arrays are addresses inside ONE region of the file's size (no copy), in
    the order and alignment the header promises; a parsed library has no mapping."""
    import ctypes as C
    host = b.HostLibrary.from_json_modules(synth.json_modules(400, 2, "mixed"))
    assert host.mapped_bytes == 0
    path = host.save_compiled(str(tmp_path / "m.bshlib"))
    again = b.HostLibrary.load_compiled(path)
    size = os.path.getsize(path)
    assert again.mapped_bytes == size
    t = again.raw_table()
    addrs = [C.cast(getattr(t, name), C.c_void_p).value for name in
             ("unit_module", "unit_delay", "unit_gate", "unit_comp_begin", "comp_pat_begin", "pattern", "unit_check_begin", "checks")]
    assert all(a % 16 == 0 for a in addrs) and addrs == sorted(addrs)
    assert addrs[-1] - addrs[0] < size
    # ... and the bytes behind the pointers are the file's bytes
    raw = open(path, "rb").read()
    first = bytes((C.c_char * 64).from_address(addrs[0]))
    assert raw.find(first) % 16 == 0


def test_corrupt_files_are_refused(tmp_path):
    host = b.HostLibrary.from_json_modules(synth.json_modules(50, 1, "c1"))
    good = open(host.save_compiled(str(tmp_path / "ok.bshlib")), "rb").read()
    rng = np.random.default_rng(5)
    cases_ = [("empty", b""), ("magic", b"NOTALIB0" + good[8:]), ("cut", good[: len(good) // 2]), ("tail", good + b"x"),
              ("units", good[:20] + b"\xff\xff\xff\x7f" + good[24:]), ("sections", good[:12] + b"\x03\x00\x00\x00" + good[16:]),
              ("offset", good[:56] + b"\x08" + good[57:])]
    for name, data in cases_:
        bad = tmp_path / f"{name}.bshlib"
        bad.write_bytes(data)
        with pytest.raises(b.SearchError):
            b.HostLibrary.load_compiled(str(bad))
    # random damage anywhere: either refused, or loaded into something that can still be walked without reading out of bounds
    for trial in range(60):
        data = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            at = int(rng.integers(8, len(data)))
            data[at] = int(rng.integers(0, 256))
        bad = tmp_path / "fuzz.bshlib"
        bad.write_bytes(bytes(data))
        try:
            lib = b.HostLibrary.load_compiled(str(bad))
        except b.SearchError:
            continue
        comps = lib.module_components()
        for m in range(lib.n_modules):
            lib.records_text(np.array([m] + [5 + 9 * j for j in range(int(comps[m]))], dtype=np.uint32))
        lib.table()


def test_loading_beats_parsing(tmp_path):
    path = synth.write_json(str(tmp_path / "big.json"), synth.json_modules(20000, 7, "mixed"))
    t0 = time.perf_counter()
    host = b.HostLibrary.load("jsonfolder", path)
    parse = time.perf_counter() - t0
    out = host.save_compiled(str(tmp_path / "big.bshlib"))
    t0 = time.perf_counter()
    again = b.HostLibrary.load_compiled(out)
    load = time.perf_counter() - t0
    assert again.n_modules == host.n_modules == 20000
    print(f"parse {parse * 1e3:.1f} ms, compiled load {load * 1e3:.1f} ms, file {os.path.getsize(out) / 1e6:.1f} MB vs json {os.path.getsize(path) / 1e6:.1f} MB")
    assert load < parse


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["jsonfolder", "rinfolder", "descfolder"])
def test_search_from_a_mapped_library_equals_the_parsed_one(tmp_path, source):
    """The device library is built straight from the mapped file (the upload reads the mapping's pages) and the search must
    give the very records of the parsed library — which are the reference's sites."""
    import util
    if source == "jsonfolder":
        path = synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(500, 3, "c1") + synth.json_fuzz_modules(60, 4))
    elif source == "rinfolder":
        path = synth.write_rin_folder(str(tmp_path / "rin"), 150, 5, max_components=4, min_len=2)
    else:
        path = synth.write_desc_folder(str(tmp_path / "desc"), 200, 6)
    parsed = b.HostLibrary.load(source, path)
    mapped = b.HostLibrary.load_compiled(parsed.save_compiled(str(tmp_path / "lib.bshlib")))
    assert mapped.mapped_bytes > 0
    seq = synth.sequence(300, 12)
    mask = synth.mask_canonical(seq, seed=3, keep=0.7)
    dev_p, dev_m = parsed.to_device(), mapped.to_device()
    a, c = dev_p.search(seq, mask), dev_m.search(seq, mask)
    assert a.n_sites == c.n_sites and np.array_equal(a.records(), c.records())
    assert sorted(mapped.records_text(c.records())) == util.best_oracle(source, path, seq, mask)
    for h in (a, c):
        h.close()
    dev_p.close()
    dev_m.close()
