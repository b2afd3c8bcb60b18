"""GPU parity: the CUDA search, called through the C ABI, against the oracles.

Bit-exact is the bar (integer / index work): the record stream must equal the flat-table
interpreter's (same canonical order), and the reconstructed insertion sites must equal the
reference's own `insertion_sites_` as a multiset (the reference's order is a thread race,
cppsrc/MOIP.cpp:1209-1211).
"""
import os

import numpy as np
import pytest

import biorseo_b200 as b
from biorseo_b200 import synth

import flat_interp
import util

pytestmark = pytest.mark.gpu


def _search(host, seq, mask):
    dev = host.to_device()
    sites = dev.search(seq, mask)
    rec = sites.records()
    assert sites.n_words == rec.size
    stats = dev.stats()
    assert stats["kernel_launches"] >= 2
    sites.close()
    dev.close()
    return rec


def _lib(tmp_path, source, seed, n_modules, n):
    if source == "jsonfolder":
        return synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(n_modules, seed, "c1"))
    if source == "rinfolder":   # keep the reference's pre-filter placement count bounded on long queries
        return synth.write_rin_folder(str(tmp_path / "rin"), n_modules, seed, max_components=6 if n <= 130 else 3, min_len=1 if n <= 130 else 2)
    return synth.write_desc_folder(str(tmp_path / "desc"), n_modules, seed)


@pytest.mark.parametrize("source", ["jsonfolder", "rinfolder", "descfolder"])
@pytest.mark.parametrize("seed,n", [(1, 73), (2, 120), (3, 230), (4, 31), (5, 129), (6, 700)])
def test_matches_interpreter_and_reference(tmp_path, source, seed, n):
    seq = synth.sequence(n, 100 + seed)
    mask = synth.mask_all(n) if seed % 2 else synth.mask_canonical(seq, seed=seed, keep=0.7)
    path = _lib(tmp_path, source, seed, 150, n)
    host = b.HostLibrary.load(source, path)
    rec = _search(host, seq, mask)
    expect = flat_interp.search_table(host.table(), seq, mask)
    assert np.array_equal(rec, expect)
    mine = sorted(host.records_text(rec))
    assert mine == util.best_oracle(source, path, seq, mask)


# ---------------------------------------------------------------------------------------
# committed golden vectors (produced by the reference's own code, tests/golden/cases.json)
# ---------------------------------------------------------------------------------------
import cases  # noqa: E402


@pytest.mark.parametrize("case", cases.CASES, ids=[c["name"] for c in cases.CASES])
def test_golden_cases(case, tmp_path):
    lib = cases.materialise(case, tmp_path)
    host = b.HostLibrary.load(case["source"], lib)
    rec = _search(host, case["seq"], cases.mask_of(case["mask"], case["seq"]))
    assert sorted(host.records_text(rec)) == case["expected"]


@pytest.mark.parametrize("entry", cases.FNO, ids=[f"{e['seq']}:{'+'.join(e['comps'])}" for e in cases.FNO])
def test_golden_enumerator(entry):
    """find_next_ones_in vectors: a unit without checks emits every placement."""
    n = len(entry["seq"])
    dev = b.DeviceLibrary.from_table(cases.fno_table(entry["comps"], entry["delay"]))
    sites = dev.search(entry["seq"], synth.mask_all(n))
    assert sites.records().tolist() == cases.fno_expected_records(entry)
    sites.close()
    dev.close()


# ---------------------------------------------------------------------------------------
# batch mode, capacity negotiation, edge cases
# ---------------------------------------------------------------------------------------
def test_batch_equals_one_by_one():
    host = b.HostLibrary.from_json_modules(synth.json_modules(400, 21, "c1") )
    dev = host.to_device()
    lens = [0, 1, 5, 6, 7, 31, 32, 33, 63, 64, 65, 120, 127, 128, 200, 73]
    seqs = [synth.sequence(n, 300 + i) for i, n in enumerate(lens)]
    masks = [synth.mask_all(n) if i % 2 else synth.mask_canonical(s, seed=i, keep=0.8) for i, (n, s) in enumerate(zip(lens, seqs))]
    batch = dev.search_batch(seqs, masks)
    rec, begin = batch.records(), batch.seq_word_begin()
    assert begin[0] == 0 and begin[-1] == rec.size
    for i, (s, m) in enumerate(zip(seqs, masks)):
        one = dev.search(s, m)
        assert np.array_equal(one.records(), rec[begin[i]:begin[i + 1]]), i
        assert np.array_equal(one.records(), flat_interp.search_table(host.table(), s, m)), i
        one.close()
    batch.close()
    dev.close()


def test_capacity_negotiation_and_device_output():
    torch = pytest.importorskip("torch")
    host = b.HostLibrary.from_json_modules(synth.json_modules(300, 22, "c4b"))
    dev = host.to_device()
    seq = synth.sequence(300, 5)
    q = dev.query([seq], [synth.mask_all(300)])
    rc, n_words, n_sites = dev.search_query_into(q, 0, 0)
    assert rc == b.capi.BSS_ERR_OVERFLOW and n_words > 0 and n_sites > 0
    small = torch.full((n_words - 1,), -7, dtype=torch.int32, device="cuda")
    rc, nw2, _ = dev.search_query_into(q, small.data_ptr(), small.numel())
    assert rc == b.capi.BSS_ERR_OVERFLOW and nw2 == n_words
    assert bool((small == -7).all())            # nothing written on overflow
    out = torch.empty(n_words, dtype=torch.int32, device="cuda")
    rc, nw3, ns3 = dev.search_query_into(q, out.data_ptr(), out.numel())
    assert rc == 0 and nw3 == n_words and ns3 == n_sites
    fetched = dev.search_query(q)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), fetched.records())
    fetched.close()
    q.close()
    dev.close()


def test_rejects_line_terminators_and_reports_errors():
    host = b.HostLibrary.from_json_modules(synth.json_modules(10, 23, "c1"))
    dev = host.to_device()
    with pytest.raises(b.SearchError) as err:
        dev.search("ACGU\nACGU", synth.mask_all(9))
    assert err.value.code == b.capi.BSS_ERR_BAD_SEQUENCE
    dev.close()


def test_self_overlapping_and_dense_patterns():
    """Leftmost non-overlapping chains restarted at every suffix: periodic patterns, wildcards,
    one-nucleotide strands and a periodic query."""
    seq = "GGGGGGGGGGAAAAAAAAAAGGGGGGGGGGCCCCCCCCCCACACACACACACGGGGGGGGGG" * 3
    n = len(seq)
    for delay in (0, 1, 5):
        for comps in (["GG", "GG"], ["GGG", "AA", "GG"], ["G.G", "C.C"], ["G", "G", "G"], ["ACA", "CAC"], ["AA", "", "GG"], ["A.", ".G", "C"]):
            table = cases.fno_table(comps, delay)
            dev = b.DeviceLibrary.from_table(table)
            got = dev.search(seq, synth.mask_all(n))
            assert np.array_equal(got.records(), flat_interp.search_table(table, seq, synth.mask_all(n))), (delay, comps)
            got.close()
            dev.close()


# ---------------------------------------------------------------------------------------
# BASELINE sizes: size-independent properties + a sample checked against the reference
# ---------------------------------------------------------------------------------------
def _walk(rec, comps):
    at, out = 0, []
    while at < rec.size:
        w = 1 + int(comps[rec[at]])
        out.append(tuple(int(x) for x in rec[at:at + w]))
        at += w
    return out


@pytest.mark.parametrize("kind,mask_kind", [("mixed", "banded"), ("c4b", "all")])
def test_full_size_c4_properties_and_reference_sample(tmp_path, kind, mask_kind):
    n, n_modules = 2000, 50000 if mask_kind == "banded" else 12000
    seq = synth.sequence(n, 4)
    mask = synth.mask_canonical(seq, seed=204, keep=0.3) if mask_kind == "banded" else synth.mask_all(n)
    modules = synth.json_modules(n_modules, 104, kind)
    host = b.HostLibrary.from_json_modules(modules)
    dev = host.to_device()
    first = dev.search(seq, mask)
    rec = first.records()
    stats = dev.stats()
    assert stats["placements_tested"] == n * n_modules and stats["n_words"] == rec.size
    again = dev.search(seq, mask)
    assert np.array_equal(rec, again.records())                       # deterministic, order included
    comps = host.module_components()
    sites = _walk(rec, comps)
    assert len(sites) == first.n_sites
    assert sites == sorted(sites, key=lambda s: s[0]) or True         # module-major order (stable check below)
    assert all(a[0] <= b_[0] for a, b_ in zip(sites, sites[1:]))      # canonical: modules ascending
    for s in sites[:: max(1, len(sites) // 5000)]:                    # every site is a strictly increasing placement
        assert all(x < y for x, y in zip(s[1:], s[2:]))
    # the first 250 modules, searched on their own, give the same records ...
    sample = 250
    sub_host = b.HostLibrary.from_json_modules(modules[:sample])
    sub_dev = sub_host.to_device()
    sub = sub_dev.search(seq, mask).records()
    cut = next((i for i, s in enumerate(sites) if s[0] >= sample), len(sites))
    assert np.array_equal(sub, rec[: sum(len(s) for s in sites[:cut])])
    # ... and those equal what the reference's own matcher finds
    path = synth.write_json(str(tmp_path / "sample.json"), modules[:sample])
    assert sorted(sub_host.records_text(sub)) == util.best_oracle("jsonfolder", path, seq, mask)
    for h in (first, again):
        h.close()
    sub_dev.close()
    dev.close()


# ---------------------------------------------------------------------------------------
# random flat tables: chains of overlapping matches x band windows x every check shape
# ---------------------------------------------------------------------------------------
from tables import random_table as _random_table, banded_random_mask as _banded_random_mask  # noqa: E402


@pytest.mark.parametrize("seed", range(12))
def test_random_tables_against_interpreter(seed):
    rng = np.random.default_rng(900 + seed)
    alphabet = [ord(c) for c in ("AG" if seed % 3 == 0 else "ACG" if seed % 3 == 1 else "ACGU")]
    n = int([37, 64, 97, 130, 200, 33][seed % 6])
    table = _random_table(rng, 60, alphabet, max_comp=4 if n < 100 else 3, max_len=4)
    # low-complexity queries: long runs and periodic stretches make matches overlap
    body = [alphabet[i] for i in rng.integers(0, len(alphabet), n)]
    for _ in range(4):
        at, run = int(rng.integers(0, n)), int(rng.integers(3, 12))
        unit = [alphabet[i] for i in rng.integers(0, len(alphabet), int(rng.integers(1, 3)))]
        for j in range(run):
            if at + j < n:
                body[at + j] = unit[j % len(unit)]
    seq = bytes(body)
    span = [n, 12, 30, 5][seed % 4]
    mask = _banded_random_mask(n, seed, span, [1.0, 0.6, 0.3][seed % 3])
    dev = b.DeviceLibrary.from_table(table)
    got = dev.search(seq, mask)
    expect = flat_interp.search_table(table, seq, mask)
    assert np.array_equal(got.records(), expect)
    got.close()
    dev.close()


@pytest.mark.parametrize("seed", range(10))
def test_random_tables_long_sequences(seed):
    """Sequences long enough for 32-lane groups: the rank-space walk (match lists, counting
    programme, unranking) and its fall-back to the depth-first walk when the lists are too long."""
    rng = np.random.default_rng(1700 + seed)
    alphabet = [ord(c) for c in ("ACGU" if seed % 2 == 0 else "ACG")]
    n = int([520, 640, 777, 900, 1025][seed % 5])
    table = _random_table(rng, 40, alphabet, max_comp=3, max_len=5, min_len=2 if seed % 3 else 1, short_limit=1)
    body = [alphabet[i] for i in rng.integers(0, len(alphabet), n)]
    for _ in range(10):
        at, run = int(rng.integers(0, n)), int(rng.integers(3, 14))
        unit = [alphabet[i] for i in rng.integers(0, len(alphabet), int(rng.integers(1, 3)))]
        for j in range(run):
            if at + j < n:
                body[at + j] = unit[j % len(unit)]
    seq = bytes(body)
    span = [n, 25, 60, 8][seed % 4]
    mask = _banded_random_mask(n, seed, span, [1.0, 0.5, 0.25][seed % 3])
    dev = b.DeviceLibrary.from_table(table)
    got = dev.search(seq, mask)
    expect = flat_interp.search_table(table, seq, mask)
    assert np.array_equal(got.records(), expect)
    got.close()
    dev.close()


@pytest.mark.parametrize("seed,threshold", [(0, 64), (3, 200), (8, 1000), (9, 5000)])
def test_heavy_items_are_walked_block_by_block(seed, threshold, monkeypatch):
    """Items with many placements are parked by the count pass and walked as (item, block-of-ranks)
    tasks by separate passes; with the threshold lowered, ordinary items take that route."""
    monkeypatch.setattr(b.api, "DEFAULT_TUNING", {"heavy_ranks": threshold, "window": 0})
    test_random_tables_long_sequences(seed)
    test_full_size_like_small(seed)


def test_full_size_like_small(seed=0):
    """c4b-like units on a 700-nt query, all-allowed mask: thousands of sites per unit."""
    seq = synth.sequence(700, 40 + seed)
    modules = synth.json_modules(120, 300 + seed, "c4b")
    host = b.HostLibrary.from_json_modules(modules)
    dev = host.to_device()
    mask = synth.mask_all(700)
    got = dev.search(seq, mask)
    assert np.array_equal(got.records(), flat_interp.search_table(host.table(), seq, mask))
    got.close()
    dev.close()


def test_gather_cut_concatenates_padded_blocks():
    """The cut kernel of the padded all-gather-v: blocks [header | records | padding] -> records back to back."""
    torch = pytest.importorskip("torch")
    dev = b.HostLibrary.from_json_modules(synth.json_modules(5, 1, "c1")).to_device()
    rng = np.random.default_rng(3)
    for world, cap, header in ((1, 64, 4), (2, 1000, 4), (8, 4096, 4), (5, 257 * 4, 22)):
        counts = [int(c) for c in rng.integers(0, cap - header + 1, world)]
        counts[rng.integers(0, world)] = 0
        blocks = rng.integers(0, 1 << 31, (world, cap), dtype=np.int64).astype(np.int32)
        for r, c in enumerate(counts):
            blocks[r, header - 2:header] = np.array([c], dtype=np.int64).view(np.int32)
        d_blocks = torch.from_numpy(blocks).cuda()
        out = torch.full((world * (cap - header),), -1, dtype=torch.int32, device="cuda")
        d_counts = torch.zeros(world, dtype=torch.int64, device="cuda")
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        dev.gather_cut(d_blocks.data_ptr(), world, cap, header, out.data_ptr(), d_counts.data_ptr(), flag.data_ptr())
        torch.cuda.synchronize()
        expect = np.concatenate([blocks[r, header:header + c] for r, c in enumerate(counts)])
        assert d_counts.tolist() == counts and int(flag.item()) == 0
        assert np.array_equal(out.cpu().numpy()[:expect.size], expect)
        blocks[0, header - 2:header] = np.array([cap], dtype=np.int64).view(np.int32)   # a count that cannot fit
        dev.gather_cut(torch.from_numpy(blocks).cuda().data_ptr(), world, cap, header, out.data_ptr(), d_counts.data_ptr(), flag.data_ptr())
        torch.cuda.synchronize()
        assert int(flag.item()) == 1
    dev.close()


@pytest.mark.parametrize("n", [4097, 9000, 16000])
def test_very_long_sequences(n):
    """Up to BSS_MAX_SEQ_LEN: the shared-memory working set grows with the sequence; the search either
    works (and is exact) or refuses with BSS_ERR_LIMIT — never a wrong answer."""
    rng = np.random.default_rng(n)
    table = _random_table(rng, 30, [ord(c) for c in "ACGU"], max_comp=3, max_len=7, min_len=5)
    seq = bytes(rng.choice([ord(c) for c in "ACGU"], n).tolist())
    mask = _banded_random_mask(n, 1, 60, 0.5)
    dev = b.DeviceLibrary.from_table(table)
    got = dev.search(seq, mask)
    assert np.array_equal(got.records(), flat_interp.search_table(table, seq, mask))
    got.close()
    dev.close()


def test_working_set_limit_is_reported():
    rng = np.random.default_rng(5)
    table = _random_table(rng, 4, [ord(c) for c in "ACGU"], max_comp=3, max_len=6, min_len=5)
    wide = cases.fno_table(["ACGUA"] * 16, 1)            # 16 components: 16 masks of 501 words per group
    seq = bytes(rng.choice([ord(c) for c in "ACGU"], 16000).tolist())
    dev = b.DeviceLibrary.from_table(wide)
    with pytest.raises(b.SearchError) as err:
        dev.search(seq, synth.mask_all(16000))
    assert err.value.code == b.capi.BSS_ERR_LIMIT
    dev.close()


@pytest.mark.parametrize("tuning", [{"list_cap": 40, "window": 0}, {"group": 32}, {"group": 32, "window": 0}, {"units_per_chunk": 3},
                                    {"group": 8}, {"window_stage": 1}], ids=str)
def test_rarely_taken_paths_forced_on_small_inputs(tuning, monkeypatch):
    """Tiny list capacity (items fall back to the depth-first walk inside 32-lane groups), 32-lane groups
    on short sequences (with and without the window walk), odd chunk sizes, narrow groups, staged window tables."""
    monkeypatch.setattr(b.api, "DEFAULT_TUNING", tuning)
    test_random_tables_long_sequences(2)
    test_random_tables_against_interpreter(4)
    test_full_size_like_small(1)


def test_full_size_c3_batch_properties_and_reference_sample(tmp_path):
    """BASELINE config 3 at full size: 10 000 queries of 120 nt x the 1 000-module JSON library, one batched
    search. Size-independent properties (offsets, determinism, batch = one-by-one on a sample) and a sample of
    sequences checked against the reference's own matcher."""
    n_seqs = 10000
    seqs = [synth.sequence(120, 3 + i) for i in range(n_seqs)]
    rng = np.random.default_rng(203)
    masks = [synth.mask_canonical(s, seed=int(rng.integers(1 << 30)), keep=0.3) for s in seqs]
    modules = synth.json_modules(1000, 101, "c1")
    host = b.HostLibrary.from_json_modules(modules)
    dev = host.to_device()
    batch = b.HostBatch(seqs, masks)
    first = dev.search_host(batch)
    rec, begin = first.records(), first.seq_word_begin()
    stats = dev.stats()
    assert stats["placements_tested"] == 120 * n_seqs * 1000 and stats["n_words"] == rec.size == begin[-1]
    assert begin[0] == 0 and np.all(np.diff(begin.astype(np.int64)) >= 0)
    again = dev.search_host(batch)
    assert np.array_equal(rec, again.records()) and np.array_equal(begin, again.seq_word_begin())
    path = synth.write_json(str(tmp_path / "lib.json"), modules)
    for i in list(range(0, n_seqs, 997)) + [n_seqs - 1]:
        one = dev.search(seqs[i], masks[i])
        mine = rec[begin[i]:begin[i + 1]]
        assert np.array_equal(one.records(), mine), i
        one.close()
        if i % 3 == 0:
            assert sorted(host.records_text(mine)) == util.best_oracle("jsonfolder", path, seqs[i], masks[i]), i
    first.close()
    again.close()
    dev.close()


# ---------------------------------------------------------------------------------------
# window pass (route_kernel + window_kernel): units whose components are tied by row checks, narrow-band masks
# ---------------------------------------------------------------------------------------
import tables  # noqa: E402


@pytest.mark.parametrize("seed", range(12))
def test_window_tables_against_interpreter(seed):
    """Random window-walk units (row checks from earlier components, ends outside their component, empty and
    self-overlapping patterns, extra checks of every shape) on low-complexity queries, banded masks of several
    widths: window pass == interpreter == ordinary passes; with the tables staged in shared memory too."""
    rng = np.random.default_rng(5000 + seed)
    alphabet = [ord(c) for c in ("AG" if seed % 3 == 0 else "ACG" if seed % 3 == 1 else "ACGU")]
    n = int([37, 64, 97, 130, 200, 33, 700, 1100, 96, 260, 520, 2000][seed])
    if n >= 500:   # keeps the number of placements the Python interpreter enumerates bounded
        alphabet = [ord(c) for c in "ACGU"]
    table = tables.window_table(rng, 60, alphabet, max_comp=4 if n < 300 else 3, max_len=4 if n < 300 else 5, min_len=1 if n < 300 else 3)
    body = [alphabet[i] for i in rng.integers(0, len(alphabet), n)]
    for _ in range(6):
        at, run = int(rng.integers(0, n)), int(rng.integers(3, 14))
        unit = [alphabet[i] for i in rng.integers(0, len(alphabet), int(rng.integers(1, 3)))]
        for j in range(run):
            if at + j < n:
                body[at + j] = unit[j % len(unit)]
    seq = bytes(body)
    span = [12, 30, 5, 60, 255, 31, 32, 33, 100, 64, 20, 99][seed]
    mask = _banded_random_mask(n, seed, span, [1.0, 0.6, 0.3][seed % 3] if n < 300 else 0.15)
    expect = flat_interp.search_table(table, seq, mask)
    for tuning in ({}, {"window": 0}, {"window_stage": 1}):
        dev = b.DeviceLibrary.from_table(table)
        for knob, value in tuning.items():
            dev.set_tuning(knob, value)
        q = dev.query([seq], [mask])
        got = dev.search_query(q)
        st = dev.stats()
        assert np.array_equal(got.records(), expect), (tuning, got.n_words, expect.size)
        if tuning.get("window", 1):
            assert st["window_items"] > 0 and st["window_items"] + st["ordinary_items"] <= 60
        else:
            assert st["window_items"] == 0
        one = dev.search(seq, mask)        # the host-buffer path: the band comes back while the query tables are built
        assert np.array_equal(one.records(), expect), tuning
        for h in (got, one):
            h.close()
        q.close()
        dev.close()


@pytest.mark.parametrize("source", ["jsonfolder", "descfolder"])
def test_window_and_ordinary_passes_agree_on_libraries(tmp_path, source):
    """Whole libraries on a 2000-nt query with the banded mask of the headline workload: the window pass takes
    (nearly) every item; switching it off must not change a word of the output."""
    n = 2000
    seq = synth.sequence(n, 4)
    mask = synth.mask_canonical(seq, seed=204, keep=0.3, span=100)
    if source == "jsonfolder":
        host = b.HostLibrary.from_json_modules(synth.json_modules(3000, 104, "mixed"))
    else:
        host = b.HostLibrary.load(source, synth.write_desc_folder(str(tmp_path / "desc"), 600, 11))
    results = {}
    for name, tuning in (("window", {}), ("ordinary", {"window": 0}), ("staged", {"window_stage": 1})):
        dev = host.to_device()
        for knob, value in tuning.items():
            dev.set_tuning(knob, value)
        q = dev.query([seq], [mask])
        s = dev.search_query(q)
        results[name] = (s.records(), dev.stats())
        if name == "window":
            assert dev.n_window_units > 0.7 * dev.n_units
        s.close()
        q.close()
        dev.close()
    assert results["window"][1]["window_items"] > 0 and results["ordinary"][1]["window_items"] == 0
    assert np.array_equal(results["window"][0], results["ordinary"][0])
    assert np.array_equal(results["window"][0], results["staged"][0])
    assert results["window"][0].size > 0


def _placements_python(table, seq):
    seq = seq.encode() if isinstance(seq, str) else bytes(seq)
    total = 0
    for u in range(int(table["n_units"])):
        gate = int(table["unit_gate"][u])
        if gate != flat_interp.NO_GATE:
            gpats, _, gdelay = flat_interp._unit(table, gate)
            if next(iter(flat_interp._placements(seq, len(seq), gpats, gdelay)), None) is None:
                continue
        pats, _, delay = flat_interp._unit(table, u)
        total += sum(1 for _ in flat_interp._placements(seq, len(seq), pats, delay))
    return total


@pytest.mark.parametrize("seed", range(6))
def test_placements_enumerated_is_the_reference_work_term(seed, tmp_path):
    """bss_placements_enumerated = the number of tuples find_next_ones_in returns, summed over the library:
    against the interpreter's enumeration on random tables (overlapping chains, empty patterns, delay 0) and
    against the reference's own enumerator on a JSON library."""
    rng = np.random.default_rng(7000 + seed)
    alphabet = [ord(c) for c in ("AG" if seed % 2 else "ACGU")]
    n = [60, 90, 150, 33, 200, 64][seed]
    table = _random_table(rng, 40, alphabet, max_comp=3, max_len=4, short_limit=1 if n > 100 else None)
    body = [alphabet[i] for i in rng.integers(0, len(alphabet), n)]
    seq = bytes(body)
    dev = b.DeviceLibrary.from_table(table)
    q = dev.query([seq], [synth.mask_all(n)])
    assert dev.placements_enumerated(q) == _placements_python(table, seq)
    q.close()
    dev.close()
    if seed == 0 and util.have_ref():
        modules = synth.json_modules(200, 9, "c4b")
        path = synth.write_json(str(tmp_path / "lib.json"), modules)
        s = synth.sequence(300, 5)
        host = b.HostLibrary.load("jsonfolder", path)
        dev = host.to_device()
        q = dev.query([s], [synth.mask_all(300)])
        assert dev.placements_enumerated(q) == util.ref_placements("jsonfolder", path, s, synth.mask_all(300))
        q.close()
        dev.close()
