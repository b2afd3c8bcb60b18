"""Insertion-variable index (row "next" #1 of the path: cppsrc/MOIP.cpp:302-328 variables, :507-586
c3 / c4 constraints built from insertion_sites_).

CPU part: the plain-Python restatement (tests/index_ref.py) is pinned against the reference's own
constraint builder, run unmodified against a recording solver shim (oracle/_ref `model` mode), on the
golden libraries and on seeded ones; a committed fixture (tests/golden/site_index.json, written by
tests/golden/make_site_index.py from the reference's output) travels to boxes without the reference.

GPU part: bss_site_index_build through the C ABI must give exactly the restatement's arrays, and the
constraints rebuilt from the device arrays must be the reference's.
"""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import biorseo_b200 as b
from biorseo_b200 import capi, synth

import cases
import index_ref
import util

FIXTURE = os.path.join(util.GOLDEN, "site_index.json")
RULE = {"jsonfolder": 0, "rinfolder": 0, "descfolder": 1}


def run_model(source, library, seq, mask):
    with tempfile.TemporaryDirectory() as tmp:
        seq_path, mask_path = util.write_query(tmp, seq, mask)
        res = subprocess.run([util.REF_BIN, "model", source, str(library), seq_path, mask_path], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    return index_ref.parse_model_dump(res.stdout)


def canonical_sites(host, source, library, seq, mask):
    """Site lines in the product's canonical order (the order of the record stream). On a box without a GPU
    the flat-table interpreter provides the records."""
    import flat_interp
    rec = flat_interp.search_table(host.table(), seq, mask)
    return host.records_text(rec)


GOLDEN_FOR_INDEX = [c for c in cases.CASES if c["name"] in
                    ("A2_json", "A3_json_alignment", "A4_rin_renumber", "A4_rin_canonical", "A5_desc", "desc_all_mask", "seed11_jsonfuzz", "seed11_rin",
                     "seed12_json", "seed12_rin", "seed12_desc")]


@pytest.mark.ref
@pytest.mark.skipif(not util.have_ref(), reason="needs oracle/_ref")
@pytest.mark.parametrize("case", GOLDEN_FOR_INDEX, ids=[c["name"] for c in GOLDEN_FOR_INDEX])
def test_restatement_equals_reference_constraints(case, tmp_path):
    lib = cases.materialise(case, tmp_path)
    seq, mask = case["seq"], cases.mask_of(case["mask"], case["seq"])
    n = len(seq)
    ref_sites, ref_constraints = run_model(case["source"], lib, seq, mask)
    host = b.HostLibrary.load(case["source"], lib)
    lines = canonical_sites(host, case["source"], lib, seq, mask)
    assert sorted(lines) == sorted(ref_sites)
    rule = RULE[case["source"]]
    index = index_ref.build(lines, n, mask, rule)
    mine = index_ref.constraints_from_index(index, lines, n, mask, rule)
    assert len(mine) == len(ref_constraints)
    assert mine == ref_constraints


@pytest.mark.ref
@pytest.mark.skipif(not util.have_ref(), reason="needs oracle/_ref")
@pytest.mark.parametrize("source", ["jsonfolder", "rinfolder", "descfolder"])
@pytest.mark.parametrize("seed,n", [(21, 60), (22, 90), (23, 120)])
def test_restatement_equals_reference_on_seeded_libraries(tmp_path, source, seed, n):
    """Wider pinning of the restatement (the GPU tests compare the device arrays with it at sizes the reference cannot
    reach): seeded libraries and dense RANDOM pair masks (not only canonical ones), constraints against the reference's.
    The JSON modules are cut out of the query itself, so that they do have placements."""
    import numpy as np
    seq = synth.sequence(n, 500 + seed)
    mask = synth.mask_random(n, seed, density=0.6) if seed % 2 else synth.mask_all(n)
    if source == "jsonfolder":
        rng = np.random.default_rng(seed)
        mods = []
        for i in range(25):
            lengths = [int(x) for x in rng.integers(2, 6, int(rng.integers(1, 4)))]
            if sum(lengths) < 4:
                lengths[0] += 4
            at, strands = int(rng.integers(0, 8)), []
            for k in lengths:
                strands.append(seq[at:at + k])
                at += k + int(rng.integers(1, 9))
            if at >= n or any(len(x) < 2 for x in strands):
                continue
            mods.append((f"m{i:02d}", "&".join(strands), synth._strand_struct([len(x) for x in strands])))
        path = synth.write_json(str(tmp_path / "lib.json"), mods)
    elif source == "rinfolder":
        path = synth.write_rin_folder(str(tmp_path / "rin"), 60, seed, max_components=3, min_len=1)
    else:
        path = synth.write_desc_folder(str(tmp_path / "desc"), 150, seed)
    ref_sites, ref_constraints = run_model(source, path, seq, mask)
    host = b.HostLibrary.load(source, path)
    lines = canonical_sites(host, source, path, seq, mask)
    assert len(lines) >= 2 and sorted(lines) == sorted(ref_sites)
    rule = RULE[source]
    index = index_ref.build(lines, n, mask, rule)
    assert index_ref.constraints_from_index(index, lines, n, mask, rule) == ref_constraints


def _fixture():
    with open(FIXTURE) as f:
        return json.load(f)


def _from_json(c):
    return sorted((op, rhs, tuple((name, coef) for name, coef in terms)) for op, rhs, terms in c)


@pytest.mark.parametrize("name", [e["name"] for e in _fixture()] if os.path.exists(FIXTURE) else [])
def test_restatement_equals_committed_fixture(name, tmp_path):
    entry = next(e for e in _fixture() if e["name"] == name)
    case = next(c for c in cases.CASES if c["name"] == name)
    lib = cases.materialise(case, tmp_path)
    seq, mask = case["seq"], cases.mask_of(case["mask"], case["seq"])
    host = b.HostLibrary.load(case["source"], lib)
    lines = canonical_sites(host, case["source"], lib, seq, mask)
    rule = RULE[case["source"]]
    index = index_ref.build(lines, len(seq), mask, rule)
    assert index_ref.constraints_from_index(index, lines, len(seq), mask, rule) == _from_json(entry["constraints"])


def run_port_index(source, library, seq, mask):
    """oracle/port `index` mode: the reference's variable / c3 / c4 loops restated in C++ (sizes + seconds)."""
    with tempfile.TemporaryDirectory() as tmp:
        seq_path, mask_path = util.write_query(tmp, seq, mask)
        res = subprocess.run([util.PORT_BIN, "index", source, str(library), seq_path, mask_path], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not util.have_port(), reason="needs oracle/_build/biorseo_port")
@pytest.mark.parametrize("name", ["A2_json", "A4_rin_canonical", "desc_all_mask", "seed12_rin"])
def test_port_index_loops_equal_restatement(name, tmp_path):
    case = next(c for c in cases.CASES if c["name"] == name)
    lib = cases.materialise(case, tmp_path)
    seq, mask = case["seq"], cases.mask_of(case["mask"], case["seq"])
    host = b.HostLibrary.load(case["source"], lib)
    lines = canonical_sites(host, case["source"], lib, seq, mask)
    index = index_ref.build(lines, len(seq), mask, RULE[case["source"]])
    port = run_port_index(case["source"], lib, seq, mask)
    assert port["sites"] == len(lines) and port["variables"] == index["var_first"].size
    assert port["c3_terms"] == int(index["var_c3_terms"].sum()) and port["c3_constraints"] == int((index["var_c3_terms"] > 0).sum())
    assert port["c4_entries"] == index["overlap_vars"].size
    assert port["c4_constraints"] == int((np.diff(index["overlap_begin"].astype(np.int64)) > 1).sum())


# ---------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------
def _device_index(host, seq, mask, rule=None):
    dev = host.to_device()
    q = dev.query([seq], [mask])
    sites = dev.search_query(q, fetch=True)
    rec = sites.records()
    idx = dev.site_index(q, sites.device_pointer(), 0, rule)
    got = idx.sections()
    assert idx.kernel_launches >= 5
    idx.close()
    sites.close()
    q.close()
    dev.close()
    return rec, got


def _assert_index_equal(got, want):
    for name in capi.IDX_SECTIONS:
        assert np.array_equal(got[name], want[name]), name


@pytest.mark.gpu
@pytest.mark.parametrize("case", GOLDEN_FOR_INDEX, ids=[c["name"] for c in GOLDEN_FOR_INDEX])
def test_gpu_index_golden(case, tmp_path):
    lib = cases.materialise(case, tmp_path)
    seq, mask = case["seq"], cases.mask_of(case["mask"], case["seq"])
    n, rule = len(seq), RULE[case["source"]]
    host = b.HostLibrary.load(case["source"], lib)
    rec, got = _device_index(host, seq, mask)
    lines = host.records_text(rec)
    _assert_index_equal(got, index_ref.build(lines, n, mask, rule))
    mine = index_ref.constraints_from_index(got, lines, n, mask, rule)
    entry = next((e for e in _fixture() if e["name"] == case["name"]), None)
    if entry is not None:
        assert mine == _from_json(entry["constraints"])
    if util.have_ref():
        assert mine == run_model(case["source"], lib, seq, mask)[1]


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["jsonfolder", "rinfolder", "descfolder"])
@pytest.mark.parametrize("seed,n", [(1, 73), (2, 120), (3, 230), (4, 31), (5, 129)])
def test_gpu_index_seeded(tmp_path, source, seed, n):
    seq = synth.sequence(n, 400 + seed)
    mask = synth.mask_all(n) if seed % 2 else synth.mask_canonical(seq, seed=seed, keep=0.7)
    if source == "jsonfolder":
        path = synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(120, seed, "c1"))
    elif source == "rinfolder":
        path = synth.write_rin_folder(str(tmp_path / "rin"), 80, seed, max_components=4 if n <= 130 else 3, min_len=2)
    else:
        path = synth.write_desc_folder(str(tmp_path / "desc"), 120, seed)
    host = b.HostLibrary.load(source, path)
    rec, got = _device_index(host, seq, mask)
    lines = host.records_text(rec)
    rule = RULE[source]
    _assert_index_equal(got, index_ref.build(lines, n, mask, rule))
    if util.have_ref() and len(lines) <= 4000:
        assert index_ref.constraints_from_index(got, lines, n, mask, rule) == run_model(source, path, seq, mask)[1]


@pytest.mark.gpu
def test_gpu_index_many_sites_properties():
    """Output-heavy case (all-allowed mask, short strands): too large for the Python restatement, checked through
    properties that hold for any input — every list sorted, every entry covers its nucleotide, sizes add up."""
    n = 600
    seq = synth.sequence(n, 9)
    mask = synth.mask_all(n)
    host = b.HostLibrary.from_json_modules(synth.json_modules(3000, 31, "c4b"))
    dev = host.to_device()
    q = dev.query([seq], [mask])
    sites = dev.search_query(q, fetch=True)
    rec = sites.records()
    idx = dev.site_index(q, sites.device_pointer(), 0)
    got = idx.sections()
    S = sites.n_sites
    assert S > 100000
    vb, vf, vs = got["var_begin"].astype(np.int64), got["var_first"].astype(np.int64), got["var_second"].astype(np.int64)
    assert vb.size == S + 1 and vb[0] == 0 and vb[-1] == vf.size == rec.size - S
    # records are [module, p0 .. pC-1]: stripping the module word of every site leaves exactly var_first
    keep = np.ones(rec.size, dtype=bool)
    keep[vb[:-1] + np.arange(S)] = False
    assert np.array_equal(rec[keep], got["var_first"])
    k = vs - vf + 1
    assert (k >= 1).all()
    ob, ov = got["overlap_begin"].astype(np.int64), got["overlap_vars"].astype(np.int64)
    assert ob[-1] == ov.size == int(k.sum())
    cover = np.zeros(n + 1, dtype=np.int64)
    np.add.at(cover, vf, 1)
    np.add.at(cover, vs + 1, -1)
    assert np.array_equal(np.diff(ob), np.cumsum(cover)[:n])
    nuc = np.repeat(np.arange(n), np.diff(ob))
    assert ((vf[ov] <= nuc) & (nuc <= vs[ov])).all()
    same = nuc[1:] == nuc[:-1]
    assert (ov[1:][same] > ov[:-1][same]).all()
    # pair side against numpy on the unpacked mask
    bits = np.unpackbits(mask.view(np.uint8), bitorder="little").reshape(n, -1)[:, :n].astype(bool)
    bits = np.triu(bits, 1)
    sym = bits | bits.T
    assert np.array_equal(np.diff(got["pair_begin"].astype(np.int64)), sym.sum(1))
    assert np.array_equal(got["pair_partner"], np.nonzero(sym)[1].astype(np.uint32))
    assert np.array_equal(np.diff(got["row_first_pair"].astype(np.int64)), bits.sum(1))
    # c3 counts: degree sums over the component minus the site's own (allowed, distinct) links; every JSON link is allowed here
    deg = np.concatenate([[0], np.cumsum(sym.sum(1))])
    assert (got["var_c3_terms"].astype(np.int64) <= deg[vs + 1] - deg[vf]).all()
    lines = host.records_text(rec[: int(vb[50] + 50)])   # first 50 sites in full
    small = index_ref.build(lines, n, mask, 0)
    assert np.array_equal(small["var_c3_terms"], got["var_c3_terms"][: int(vb[50])])
    idx.close()
    sites.close()
    q.close()
    dev.close()


@pytest.mark.gpu
@pytest.mark.skipif(not util.have_port(), reason="needs oracle/_build/biorseo_port")
def test_gpu_index_totals_equal_host_loops(tmp_path):
    """20 000 sites: the device index against the reference's host loops (restated in oracle/port, `index` mode)."""
    n = 2000
    seq, mask = synth.sequence(n, 4), synth.mask_all(n)
    mods = synth.json_modules(2000, 104, "c4b")[:20]
    path = synth.write_json(str(tmp_path / "lib.json"), mods)
    host = b.HostLibrary.load("jsonfolder", path)
    rec, got = _device_index(host, seq, mask)
    port = run_port_index("jsonfolder", path, seq, mask)
    assert port["sites"] == got["var_begin"].size - 1 and port["variables"] == got["var_first"].size
    assert port["c3_terms"] == int(got["var_c3_terms"].astype(np.int64).sum())
    assert port["c3_constraints"] == int((got["var_c3_terms"] > 0).sum())
    assert port["c4_entries"] == got["overlap_vars"].size
    assert port["c4_constraints"] == int((np.diff(got["overlap_begin"].astype(np.int64)) > 1).sum())


@pytest.mark.gpu
def test_gpu_index_long_and_short_components():
    """Components longer than a warp (the overlap lists' sequential step) next to short ones (4-, 8-, 16-lane groups)."""
    n = 150
    seq = synth.sequence(n, 77)
    mask = synth.mask_all(n)
    mods = [("long", seq[5:45], "(" + "." * 38 + ")"), ("long2", seq[20:55] + "&" + seq[70:110], "(" + "." * 34 + "&" + "." * 39 + ")"),
            ("mid", seq[30:42] + "&" + seq[60:66], "(" + "." * 11 + "&" + "." * 5 + ")")]
    mods += [(f"s{i}", seq[i:i + 3] + "&" + seq[i + 9:i + 13], "(..&...)") for i in range(0, 120, 7)]
    host = b.HostLibrary.from_json_modules(sorted(mods))
    rec, got = _device_index(host, seq, mask)
    lines = host.records_text(rec)
    assert any(line.startswith("JSONlong\t") for line in lines) and any(line.startswith("JSONlong2\t") for line in lines)
    _assert_index_equal(got, index_ref.build(lines, n, mask, 0))


@pytest.mark.gpu
def test_gpu_index_batch_sequence_and_errors():
    host = b.HostLibrary.from_json_modules(synth.json_modules(200, 5, "c1"))
    dev = host.to_device()
    seqs = [synth.sequence(n, 50 + i) for i, n in enumerate((90, 0, 64, 130))]
    masks = [synth.mask_all(len(s)) for s in seqs]
    q = dev.query(seqs, masks)
    sites = dev.search_query(q, fetch=True)
    rec, begin = sites.records(), sites.seq_word_begin()
    for i, (s, m) in enumerate(zip(seqs, masks)):
        idx = dev.site_index(q, sites.device_pointer(), i)
        lines = host.records_text(rec[begin[i]:begin[i + 1]])
        _assert_index_equal(idx.sections(), index_ref.build(lines, len(s), m, 0))
        idx.close()
    with pytest.raises(b.SearchError):
        dev.site_index(q, sites.device_pointer(), 7)
    other = dev.query([seqs[0]], [masks[0]])
    with pytest.raises(b.SearchError):          # not the query of the last search
        dev.site_index(other, sites.device_pointer(), 0)
    other.close()
    sites.close()
    q.close()
    dev.close()
