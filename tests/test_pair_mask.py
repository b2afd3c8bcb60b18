"""Pair-mask producer (the row "next" upstream of the search: cppsrc/MOIP.cpp:77-90 + 896-910).

CPU: the numpy restatement of the rule against the reference's own MOIP constructor
(oracle/_ref `pairmask` mode) and against the committed golden fixture.
GPU: the CUDA kernels (row-major, column-major and arbitrary strides) against both, and a
search through a device-built mask against the search through the host mask.
"""
import json
import os
import subprocess

import numpy as np
import pytest

import biorseo_b200 as b
from biorseo_b200 import synth

import util

GOLDEN = os.path.join(util.GOLDEN, "pair_mask.json")


def rule(pij, theta):
    """allowed_basepair as a packed mask: v >= u + 4, u < n - 6, pij(u, v) > theta (float32 compare)."""
    pij = np.asarray(pij, dtype=np.float32)
    n = pij.shape[0]
    if n <= 6:
        return np.zeros((n, (n + 31) // 32), dtype=np.uint32)
    u, v = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    with np.errstate(invalid="ignore"):
        return b.pack_mask((v - u >= 4) & (u < n - 6) & (pij > np.float32(theta)))


def matrix(n, seed, theta):
    """Probabilities scattered around theta, with exact ties, zeros, ones and a banded structure."""
    rng = np.random.default_rng(seed)
    p = (rng.random((n, n)) * 4 * theta).astype(np.float32)
    p[rng.random((n, n)) < 0.1] = np.float32(theta)          # ties are NOT allowed (strict >)
    p[rng.random((n, n)) < 0.1] = 0
    p[rng.random((n, n)) < 0.05] = 1
    u, v = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    p[np.abs(v - u) > max(8, n // 2)] *= np.float32(0.01)
    return p


def reference_mask(seq, pij, theta, tmp):
    seq_path = os.path.join(str(tmp), "s.txt")
    open(seq_path, "w").write(seq + "\n")
    pij_path = os.path.join(str(tmp), "p.bin")
    np.ascontiguousarray(pij, dtype="<f4").tofile(pij_path)
    lib = os.path.join(str(tmp), "empty.json")
    open(lib, "w").write("{}")
    out = subprocess.run([util.REF_BIN, "pairmask", seq_path, pij_path, repr(float(theta)), lib], capture_output=True, text=True, check=True).stdout
    rows = [r for r in out.split("\n") if r]
    W = (len(seq) + 31) // 32
    return np.array([[int(r[8 * w:8 * w + 8], 16) for w in range(W)] for r in rows], dtype=np.uint32).reshape(len(seq), W)


CASES = [(7, 1, 1e-3), (12, 2, 1e-3), (33, 3, 1e-3), (64, 4, 0.05), (73, 5, 1e-3), (130, 6, 0.3)]


@pytest.mark.ref
@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref is not built")
@pytest.mark.parametrize("n,seed,theta", CASES)
def test_rule_matches_the_reference_constructor(n, seed, theta, tmp_path):
    pij = matrix(n, seed, theta)
    assert np.array_equal(rule(pij, theta), reference_mask(synth.sequence(n, seed), pij, theta, tmp_path))


def test_rule_matches_the_golden_fixture():
    for case in json.load(open(GOLDEN))["cases"]:
        pij = matrix(case["n"], case["seed"], case["theta"])
        got = rule(pij, case["theta"])
        assert [f"{int(x):08x}" for x in got.reshape(-1)] == case["mask_hex"], case["n"]


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed,theta", CASES + [(5, 7, 1e-3), (700, 8, 1e-3), (2000, 9, 1e-3)])
def test_device_mask_matches_rule_for_every_layout(n, seed, theta):
    dev = b.HostLibrary.from_json_modules(synth.json_modules(5, 1, "c1")).to_device()
    pij = matrix(n, seed, theta)
    expect = rule(pij, theta)
    assert np.array_equal(dev.pair_mask(pij, theta), expect)                              # row-major
    assert np.array_equal(dev.pair_mask(np.asfortranarray(pij), theta), expect)           # column-major (Eigen)
    padded = np.zeros((n, n + 3), dtype=np.float32)
    padded[:, :n] = pij
    assert np.array_equal(dev.pair_mask(padded[:, :n], theta), expect)                    # row-major with a leading dimension
    wide = np.zeros((2 * n, 2 * n), dtype=np.float32)
    wide[::2, ::2] = pij
    assert np.array_equal(dev.pair_mask(wide[::2, ::2], theta), expect)                   # arbitrary strides
    dev.close()


@pytest.mark.gpu
def test_golden_fixture_on_the_device():
    dev = b.HostLibrary.from_json_modules(synth.json_modules(5, 1, "c1")).to_device()
    for case in json.load(open(GOLDEN))["cases"]:
        got = dev.pair_mask(matrix(case["n"], case["seed"], case["theta"]), case["theta"])
        assert [f"{int(x):08x}" for x in got.reshape(-1)] == case["mask_hex"]
    dev.close()


@pytest.mark.gpu
def test_search_through_a_device_built_mask():
    n, theta = 300, 1e-3
    seq = synth.sequence(n, 12)
    host = b.HostLibrary.from_json_modules(synth.json_modules(300, 13, "c1"))
    dev = host.to_device()
    # probabilities: canonical pairs above theta, the rest below
    allowed = synth.mask_canonical(seq, seed=3, keep=0.7)
    bits = np.unpackbits(allowed.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
    rng = np.random.default_rng(5)
    pij = np.where(bits, 0.5, 0.0005).astype(np.float32) * (0.5 + rng.random((n, n))).astype(np.float32)
    mask = rule(pij, theta)
    q = dev.query_pij(seq, np.asfortranarray(pij), theta)
    via_device = dev.search_query(q).records()
    via_host = dev.search(seq, mask).records()
    assert via_device.size > 0 and np.array_equal(via_device, via_host)
    q.close()
    dev.close()
