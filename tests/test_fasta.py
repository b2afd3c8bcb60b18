"""Multi-record FASTA front-end of the batch search (row "next" #4): bsh::load_fasta against the reference's own
Fasta::load (cppsrc/fa.cpp:22-60, compiled unmodified into oracle/_ref, `fasta` mode) on files with every quirk
the loader has, and against a committed fixture of the reference's output; on the GPU, the C++ drop-in's batch
mode over such a file against one search per record."""
import json
import os
import subprocess

import pytest

import biorseo_b200 as b
from biorseo_b200 import synth

import util

FIXTURE = os.path.join(util.GOLDEN, "fasta.json")

FILES = {
    "plain": ">a\nACGU\n>b desc ription\nGGGAAACCC\n",
    "wrapped_and_structure": ">rna1\nACGUACGU\nGGCC\n((..))..\n.(.)\n>rna2\nAAAA\n",
    "quirks": "ACGU\n>first\nACGU12ACGU\nacgu tail\n\n>\nGGGG\n>second\r\nAC-GU\r\nx(\n>third\n?..x\nNNNN*\n",
    "no_trailing_newline": ">x\nACGUACGU",
    "empty": "",
    "only_header": ">lonely\n",
    "starts_like_structure": ">s\nleucine\nelement\nACGU\n",
}


def _reference(path):
    out = subprocess.run([util.REF_BIN, "fasta", path], capture_output=True, check=True).stdout.decode()   # bytes: names may hold a '\r'
    return [tuple(line.split("\t")) for line in out.split("\n")[:-1]]


@pytest.mark.parametrize("name", sorted(FILES))
def test_reader_equals_committed_reference_output(name, tmp_path):
    path = tmp_path / f"{name}.fa"
    path.write_bytes(FILES[name].encode())
    with open(FIXTURE) as f:
        want = [tuple(r) for r in json.load(f)[name]]
    assert b.read_fasta(path) == want


@pytest.mark.ref
@pytest.mark.skipif(not util.have_ref(), reason="needs oracle/_ref")
@pytest.mark.parametrize("name", sorted(FILES))
def test_reader_equals_reference(name, tmp_path):
    path = tmp_path / f"{name}.fa"
    path.write_bytes(FILES[name].encode())
    assert b.read_fasta(path) == _reference(str(path))


def test_mismatched_structure_before_another_record_is_an_error(tmp_path):
    """The reference aborts there (assert in cppsrc/fa.cpp:33, assertions are on in its build); the last record is not checked."""
    path = tmp_path / "bad.fa"
    path.write_text(">a\nACGU\n((\n>b\nACGU\n")
    with pytest.raises(b.SearchError):
        b.read_fasta(path)
    if util.have_ref():
        assert subprocess.run([util.REF_BIN, "fasta", str(path)], capture_output=True).returncode != 0
    path.write_text(">a\nACGU\n>b\nACGU\n((\n")
    assert b.read_fasta(path) == [("a", "ACGU", ""), ("b", "ACGU", "((")]


def test_missing_file_is_an_error(tmp_path):
    with pytest.raises(b.SearchError):
        b.read_fasta(tmp_path / "nope.fa")


@pytest.mark.gpu
def test_batch_over_a_fasta_file_through_the_cpp_dropin(tmp_path):
    """dropin batch: every record of a FASTA file searched in one bss_search_batch call, against one search per record."""
    from test_dropin_cpp import build_dropin
    binary = build_dropin()
    seqs = [synth.sequence(n, 900 + i) for i, n in enumerate((73, 120, 8, 64, 200))]
    fasta = tmp_path / "batch.fa"
    fasta.write_text("".join(f">seq{i}\n{s[:50]}\n{s[50:]}\n" for i, s in enumerate(seqs)))
    lib = synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(150, 3, "c1"))
    res = subprocess.run([binary, "jsonfolder", lib, str(fasta), "-", "batch"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    got = {}
    for line in res.stdout.split("\n"):
        if line:
            name, site = line.split("\t", 1)
            got.setdefault(name, []).append(site)
    host = b.HostLibrary.load("jsonfolder", lib)
    dev = host.to_device()
    for i, s in enumerate(seqs):
        one = dev.search(s, synth.mask_canonical(s))
        assert got.get(f"seq{i}", []) == host.records_text(one.records()), i
        one.close()
    dev.close()


@pytest.mark.ref
@pytest.mark.skipif(not util.have_ref(), reason="needs oracle/_ref")
def test_reader_equals_reference_on_random_files(tmp_path):
    """Differential fuzz: random line soups over the characters the loader distinguishes; where the reference aborts
    (its assertion on a record closed by another header) the reader must refuse the file, elsewhere the records agree."""
    import random
    rng = random.Random(20240607)
    pieces = ["ACGU", "acgu", "N", ">", ">id", "(", ")", ".", "[", "]", "?", "x", "l", "e", " ", "\r", "-", "1", "*", "\t", "GG", "((..))", "leu"]
    for i in range(150):
        lines = ["".join(rng.choice(pieces) for _ in range(rng.randint(0, 5))) for _ in range(rng.randint(0, 9))]
        path = tmp_path / f"f{i}.fa"
        path.write_bytes(("\n".join(lines) + ("\n" if rng.random() < 0.7 else "")).encode())
        ref = subprocess.run([util.REF_BIN, "fasta", str(path)], capture_output=True)
        if ref.returncode != 0:
            with pytest.raises(b.SearchError):
                b.read_fasta(path)
            continue
        want = [tuple(line.split("\t")) for line in ref.stdout.decode().split("\n")[:-1]]
        got = b.read_fasta(path)
        assert got == want, (i, lines)
