"""Host side without a GPU: the C++ loaders (JSON / RIN / DESC -> flat component table) and
the record -> Motif reconstruction, checked against the golden vectors and the oracles by
running the flat table through the pure-Python interpreter (tests/flat_interp.py)."""
import numpy as np
import pytest

import biorseo_b200 as b
from biorseo_b200 import synth

import cases
import flat_interp
import util


def _sites(source, lib, seq, mask):
    host = b.HostLibrary.load(source, lib)
    rec = flat_interp.search_table(host.table(), seq, mask)
    return host, sorted(host.records_text(rec))


@pytest.mark.parametrize("case", cases.CASES, ids=[c["name"] for c in cases.CASES])
def test_loader_plus_interpreter_reproduces_golden(case, tmp_path):
    lib = cases.materialise(case, tmp_path)
    _, got = _sites(case["source"], lib, case["seq"], cases.mask_of(case["mask"], case["seq"]))
    assert got == case["expected"]


@pytest.mark.parametrize("entry", cases.FNO, ids=[f"{e['seq']}:{'+'.join(e['comps'])}" for e in cases.FNO])
def test_interpreter_enumerator_golden(entry):
    n = len(entry["seq"])
    rec = flat_interp.search_table(cases.fno_table(entry["comps"], entry["delay"]), entry["seq"], synth.mask_all(n))
    assert rec.tolist() == cases.fno_expected_records(entry)


@pytest.mark.skipif(not (util.have_ref() or util.have_port()), reason="no oracle binary")
@pytest.mark.parametrize("seed", range(10))
def test_loaders_match_oracle_on_random_libraries(seed, tmp_path):
    n = 60 + 31 * seed
    seq = synth.sequence(n, 900 + seed)
    mask = synth.mask_all(n) if seed % 2 == 0 else synth.mask_canonical(seq, seed=seed, keep=0.8)
    libs = [("jsonfolder", synth.write_json(str(tmp_path / "f.json"), synth.json_fuzz_modules(120, 40 + seed))),
            ("rinfolder", synth.write_rin_folder(str(tmp_path / "rin"), 50, 40 + seed)),
            ("descfolder", synth.write_desc_folder(str(tmp_path / "desc"), 50, 40 + seed))]
    for source, lib in libs:
        _, got = _sites(source, lib, seq, mask)
        assert got == util.best_oracle(source, lib, seq, mask), (source, seed)


def test_json_reject_codes(tmp_path):
    mods = [("ok", "GGA&UCC", "((.&.))"), ("x", "GGA", "(.)."), ("r", "GGNA", "(..)"), ("n", "GGAA", "((.)"), ("n2", "GGAA", ")..("),
            ("l", "GGA", "(.)"), ("l2", "G&G&A&C", "(&.&.&)"), ("nolink", "GGAA", "...."), ("U", "GGAAC", "(&&&)")]
    host = b.HostLibrary.load("jsonfolder", synth.write_json(str(tmp_path / "l.json"), mods))
    assert dict(host.rejected()) == {"x": "x", "r": "r", "n": "n", "n2": "n", "l": "l", "l2": "l", "U": "U"}
    assert host.n_seen == len(mods) and host.n_modules == 2      # "ok" and the link-less "nolink"
    t = host.table()
    assert t["n_units"] == 1                                     # a module without links is never searched
    assert host.module_id(0) == "nolink" and host.module_id(1) == "ok"   # byte-wise key order


def test_json_key_order_is_bytewise(tmp_path):
    mods = [(k, "GGAAC", "(...)") for k in ["10", "9", "1", "100", "b", "B", "a"]]
    host = b.HostLibrary.load("jsonfolder", synth.write_json(str(tmp_path / "k.json"), mods))
    assert [host.module_id(i) for i in range(host.n_modules)] == sorted(k for k, _, _ in mods)


def test_rin_reject_and_ids(tmp_path):
    folder = synth.write_rin_folder(str(tmp_path / "rin"), 30, 3, invalid_fraction=0.5)
    host = b.HostLibrary.load("rinfolder", folder)
    codes = {c for _, c in host.rejected()}
    assert codes <= {"l", "x"} and codes
    assert host.n_seen == 30
    ids = [int(host.module_id(i)) for i in range(host.n_modules)]
    assert all(1 <= i <= 30 for i in ids)


def test_table_shapes_and_limits(tmp_path):
    host = b.HostLibrary.from_json_modules(synth.json_modules(500, 9, "mixed"))
    t = host.table()
    assert t["n_units"] == 500 and t["n_gates"] == 0
    assert t["unit_comp_begin"][-1] == len(t["comp_pat_begin"]) - 1
    assert t["comp_pat_begin"][-1] == len(t["pattern"])
    assert (np.diff(t["unit_comp_begin"].astype(np.int64)) >= 1).all()
    assert set(np.unique(t["pattern"]).tolist()) <= set(b"ACGU")
    # too many strands is a reject, not a crash
    wide = "&".join(["ACGU"] * 17)
    host2 = b.HostLibrary.from_json_modules([("wide", wide, wide.replace("A", "(").replace("U", ")").replace("C", ".").replace("G", "."))])
    assert host2.rejected() == [("wide", "L")]


def test_host_batch_packs_the_flat_buffers_of_the_c_abi():
    """HostBatch = the packed inputs of bss_search_batch: concatenated sequences + offsets, masks + offsets."""
    import numpy as np
    seqs = ["ACGUACGUA", "", "GGGAAACCC", "A"]
    masks = [synth.mask_all(len(s)) for s in seqs]
    hb = b.HostBatch(seqs, masks)
    assert hb.n_seqs == 4 and hb.seq_begin.tolist() == [0, 9, 9, 18, 19]
    assert bytes(hb.seqs[:19]) == "".join(seqs).encode()
    assert hb.mask_begin.tolist() == [0, 9, 9, 18, 19]          # n * ceil(n/32) words each
    for i, m in enumerate(masks):
        assert np.array_equal(hb.masks[int(hb.mask_begin[i]):int(hb.mask_begin[i + 1])], np.asarray(m, dtype=np.uint32).reshape(-1))
    assert hb.h2d_bytes == 19 + 4 * 19 + 16 * 5
