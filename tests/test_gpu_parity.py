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


def _lib(tmp_path, source, seed, n_modules):
    if source == "jsonfolder":
        return synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(n_modules, seed, "c1"))
    if source == "rinfolder":
        return synth.write_rin_folder(str(tmp_path / "rin"), n_modules, seed)
    return synth.write_desc_folder(str(tmp_path / "desc"), n_modules, seed)


@pytest.mark.parametrize("source", ["jsonfolder", "rinfolder", "descfolder"])
@pytest.mark.parametrize("seed,n", [(1, 73), (2, 120), (3, 230), (4, 31), (5, 129), (6, 700)])
def test_matches_interpreter_and_reference(tmp_path, source, seed, n):
    seq = synth.sequence(n, 100 + seed)
    mask = synth.mask_all(n) if seed % 2 else synth.mask_canonical(seq, seed=seed, keep=0.7)
    path = _lib(tmp_path, source, seed, 150)
    host = b.HostLibrary.load(source, path)
    rec = _search(host, seq, mask)
    expect = flat_interp.search_table(host.table(), seq, mask)
    assert np.array_equal(rec, expect)
    mine = sorted(host.records_text(rec))
    assert mine == util.best_oracle(source, path, seq, mask)
