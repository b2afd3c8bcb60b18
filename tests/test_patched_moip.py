"""The drop-in claim, checked on the reference's own code: oracle/_ref/biorseo_patched is the reference's MOIP
constructor (cppsrc/MOIP.cpp) with the patch of INTEGRATION.md applied — the source branches replaced by
bsh::InsertionSiteSearch, the per-module job functions dropped, everything before and after unchanged — compiled
against the product's Motif.h / Pool.h and linked with libbiorseo_b200.so (oracle/patched/make_patched.py; test
infrastructure, generated at build time from /root/reference, nothing of it is stored in the repository).

  * it compiles and links (the product's host interface really is what the reference's code needs),
  * without a GPU the library branches exit loudly — no CPU fallback,
  * the csvfile branch (host-only) gives the reference's insertion_sites_ AND the reference's constraints,
  * on the GPU the library branches do: same sites, same constraints as the unpatched reference.
"""
import os
import subprocess

import numpy as np
import pytest

from biorseo_b200 import synth

import cases
import index_ref
import util

PATCHED = os.path.join(util.ROOT, "oracle", "_ref", "biorseo_patched")


@pytest.fixture(scope="module")
def patched():
    if os.path.isdir("/root/reference/cppsrc"):
        subprocess.run(["make", "-C", os.path.join(util.ROOT, "oracle"), "patched"], check=True, stdout=subprocess.DEVNULL)
    if not os.access(PATCHED, os.X_OK):
        pytest.skip("oracle/_ref/biorseo_patched is not built (needs the reference sources)")
    return PATCHED


def _run(binary, mode, source, library, seq, mask, tmp):
    seq_path, mask_path = util.write_query(str(tmp), seq, mask)
    return subprocess.run([binary, mode, source, str(library), seq_path, mask_path], capture_output=True, timeout=600)


def _model(binary, source, library, seq, mask, tmp):
    res = _run(binary, "model", source, library, seq, mask, tmp)
    assert res.returncode == 0, res.stderr.decode()[-2000:]
    sites, constraints = index_ref.parse_model_dump(res.stdout.decode())
    return sorted(sites), constraints


def test_patched_constructor_fails_loudly_without_a_gpu(patched, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    case = cases.CASES[0]
    lib = cases.materialise(case, tmp_path)
    res = _run(patched, "ctor", case["source"], lib, case["seq"], cases.mask_of(case["mask"], case["seq"]), tmp_path)
    assert res.returncode != 0 and res.stdout == b""
    assert b"no CUDA device" in res.stderr or b"CUDA" in res.stderr


@pytest.mark.ref
@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref is not built")
@pytest.mark.parametrize("seed,n", [(1, 73), (2, 120)])
def test_patched_constructor_csv_source_equals_reference(patched, seed, n, tmp_path):
    """Host-only branch: the whole patched constructor runs here — sites, C variables and constraints must be the
    unpatched reference's."""
    from test_dropin_cpp import _csv_lines
    rng = np.random.default_rng(seed)
    seq = synth.sequence(n, seed)
    mask = synth.mask_canonical(seq, seed=seed, keep=0.9)
    csv = tmp_path / "placed.csv"
    csv.write_text("\n".join(_csv_lines(rng, n, 120)) + "\n")
    mine = _model(patched, "csvfile", csv, seq, mask, tmp_path)
    want = _model(util.REF_BIN, "csvfile", csv, seq, mask, tmp_path)
    assert len(mine[0]) > 0 and len(mine[1]) > 0
    assert mine == want


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["A2_json", "A4_rin_canonical", "A5_desc", "seed11_rin", "seed12_json", "desc_all_mask"])
def test_patched_constructor_equals_reference_on_the_gpu(patched, name, tmp_path):
    """Library branches: the reference's constructor with the CUDA search inside gives the sites of the golden cases
    and — where the unpatched reference is present — exactly its constraints."""
    case = next(c for c in cases.CASES if c["name"] == name)
    lib = cases.materialise(case, tmp_path)
    seq, mask = case["seq"], cases.mask_of(case["mask"], case["seq"])
    res = _run(patched, "ctor", case["source"], lib, seq, mask, tmp_path)
    assert res.returncode == 0, res.stderr.decode()[-2000:]
    assert sorted(line for line in res.stdout.decode().split("\n") if line) == case["expected"]
    mine = _model(patched, case["source"], lib, seq, mask, tmp_path)
    assert mine[0] == case["expected"]
    if util.have_ref():
        assert mine == _model(util.REF_BIN, case["source"], lib, seq, mask, tmp_path)
