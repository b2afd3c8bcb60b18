"""Parity at the headline sizes, against the reference's own matcher (oracle/_ref/biorseo_ref: the unmodified
cppsrc/Motif.cpp + MOIP.cpp + Pool.cpp), not against a sample:

  * BASELINE config 4 (2000-nt query x 50 000 JSON modules, banded mask): EVERY module's insertion sites;
  * BASELINE config 3 (10 000 x 120-nt queries x 1 000 modules, one batched search): 200 sequences spread over the batch;
  * the queries the reference ships (data/fasta/example.fa:2, data/fasta/applications.fa:2,4,6 — the last-but-one ends in a
    lower-case 'a', which only matches a lower-case pattern byte) against JSON, RIN and DESC libraries.
"""
import os

import numpy as np
import pytest

import biorseo_b200 as b
from biorseo_b200 import synth

import util

pytestmark = pytest.mark.gpu

# data/fasta/example.fa:2 (= applications.fa:2), applications.fa:4, applications.fa:6
SHIPPED = [
    "GGGGUAUCGCCAAGCGGUAAGGCACCGGAUUCUGAUUCCGGAGGUCGAGGUUCGAAUCCUCGUACCCCAGCCA",
    "GGACAUACAAUCGCGUGGAUAUGGCACGCAAGUUUCUGCCGGGCACCGUAAAUGUCCGACUAUGUCCa",
    "GGGCUGUUUUUCUCGCUGACUUUCAGCCCCAAACAAAAAAGUCAGCA",
]


def test_c4_every_module_against_the_reference(tmp_path):
    n, n_modules = 2000, 50000
    seq = synth.sequence(n, 4)
    mask = synth.mask_canonical(seq, seed=204, keep=0.3, span=100)
    modules = synth.json_modules(n_modules, 104, "mixed")
    host = b.HostLibrary.from_json_modules(modules)
    dev = host.to_device()
    found = dev.search(seq, mask)
    mine = sorted(host.records_text(found.records()))
    stats = dev.stats()
    assert stats["window_items"] > 20000 and found.n_sites > 1000
    path = synth.write_json(str(tmp_path / "c4.json"), modules)
    expect = util.best_oracle("jsonfolder", path, seq, mask, timeout=1500, threads=max(2, (os.cpu_count() or 2) - 1))
    assert len(mine) == len(expect) and mine == expect
    # the same library without the pattern table and without the window pass: not a word may change
    for knob in ("pattern_table", "window"):
        other = host.to_device()
        other.set_tuning(knob, 0)
        again = other.search(seq, mask)
        assert np.array_equal(again.records(), found.records()), knob
        again.close()
        other.close()
    found.close()
    dev.close()


def test_c3_two_hundred_sequences_against_the_reference(tmp_path):
    n_seqs = 10000
    seqs = [synth.sequence(120, 3 + i) for i in range(n_seqs)]
    rng = np.random.default_rng(203)
    masks = [synth.mask_canonical(s, seed=int(rng.integers(1 << 30)), keep=0.3) for s in seqs]
    modules = synth.json_modules(1000, 101, "c1")
    host = b.HostLibrary.from_json_modules(modules)
    dev = host.to_device()
    found = dev.search_host(b.HostBatch(seqs, masks))
    rec, begin = found.records(), found.seq_word_begin()
    path = synth.write_json(str(tmp_path / "lib.json"), modules)
    checked = with_sites = 0
    for i in range(7, n_seqs, 50):
        mine = sorted(host.records_text(rec[begin[i]:begin[i + 1]]))
        assert mine == util.best_oracle("jsonfolder", path, seqs[i], masks[i]), i
        checked += 1
        with_sites += bool(mine)
    assert checked == 200 and with_sites > 20
    found.close()
    dev.close()


@pytest.mark.parametrize("source", ["jsonfolder", "rinfolder", "descfolder"])
def test_shipped_queries(tmp_path, source):
    if source == "jsonfolder":
        # the c1 library, plus modules cut from the queries themselves (so that every query has sites, the lower-case one included)
        cut = [("q0", "GGGGUAUC&GGUUCGAAUCCUCGUACCCC", "(......(&)..................)"), ("q1", "UGUCCa", "(....)"), ("q1b", "GUCCGACUAUGUCCa", "(.............)")]
        path = synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(1000, 101, "c1") + cut)
    elif source == "rinfolder":
        path = synth.write_rin_folder(str(tmp_path / "rin"), 200, 3, max_components=4, min_len=2)
    else:
        path = synth.write_desc_folder(str(tmp_path / "desc"), 300, 5)
    host = b.HostLibrary.load(source, path)
    dev = host.to_device()
    total = 0
    for i, seq in enumerate(SHIPPED):
        for mask in (synth.mask_canonical(seq.upper().replace("T", "U"), seed=201 + i, keep=0.6), synth.mask_all(len(seq))):
            found = dev.search(seq, mask)
            mine = sorted(host.records_text(found.records()))
            assert mine == util.best_oracle(source, path, seq, mask), (i, len(mine))
            total += len(mine)
            found.close()
    # the three queries in one batch give the same sites as one by one
    masks = [synth.mask_all(len(s)) for s in SHIPPED]
    batch = dev.search_host(b.HostBatch(SHIPPED, masks))
    rec, begin = batch.records(), batch.seq_word_begin()
    for i, seq in enumerate(SHIPPED):
        one = dev.search(seq, masks[i])
        assert np.array_equal(one.records(), rec[begin[i]:begin[i + 1]])
        one.close()
    batch.close()
    assert total > 0
    dev.close()
