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


def test_corrupt_files_are_refused(tmp_path):
    host = b.HostLibrary.from_json_modules(synth.json_modules(50, 1, "c1"))
    good = open(host.save_compiled(str(tmp_path / "ok.bshlib")), "rb").read()
    for name, data in (("empty", b""), ("magic", b"NOTALIB0" + good[8:]), ("cut", good[: len(good) // 2]), ("tail", good + b"x"),
                       ("count", good[:20] + b"\xff\xff\xff\x7f" + good[24:])):
        bad = tmp_path / f"{name}.bshlib"
        bad.write_bytes(data)
        with pytest.raises(b.SearchError):
            b.HostLibrary.load_compiled(str(bad))


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
