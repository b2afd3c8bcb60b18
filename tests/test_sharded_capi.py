"""bss_search_sharded (one process, several GPUs; include/biorseo_b200.h) against the single-GPU search of the same library:
the record stream must be the same, bit for bit and in order — one sequence and batches, JSON / RIN / DESC (gate units are
renumbered per range), ranges without sites, more ranges than GPUs' worth of modules. Runs on one GPU too (every range on
device 0: the cut, the per-range libraries, the displacement of the pieces are the same code); on a box with several GPUs
the ranges go to different devices."""
import numpy as np
import pytest
import torch

import biorseo_b200 as b
from biorseo_b200 import synth

pytestmark = pytest.mark.gpu


def _devices(n_ranges):
    present = torch.cuda.device_count()
    return [r % present for r in range(n_ranges)]


def _library(kind, tmp_path):
    if kind == "json":
        return b.HostLibrary.from_json_modules(synth.json_modules(600, 5, "c1") + [(f"y{i}", s, t) for i, (_, s, t) in enumerate(synth.json_modules(150, 6, "c4b"))])
    if kind == "rin":
        return b.HostLibrary.load("rinfolder", synth.write_rin_folder(str(tmp_path / "rin"), 200, 5, max_components=4, min_len=2))
    return b.HostLibrary.load("descfolder", synth.write_desc_folder(str(tmp_path / "desc"), 300, 6))


@pytest.mark.parametrize("kind", ["json", "rin", "desc"])
@pytest.mark.parametrize("n_ranges", [1, 2, 4])
def test_sharded_equals_single_gpu(tmp_path, kind, n_ranges):
    host = _library(kind, tmp_path)
    seqs = [synth.sequence(n, 30 + i) for i, n in enumerate((160, 97, 230))]
    masks = [synth.mask_canonical(s, seed=i, keep=0.7) for i, s in enumerate(seqs)]
    if kind == "rin":
        masks[0] = synth.mask_all(len(seqs[0]))
    masks[1][:] = 0          # a sequence without any site for the pair-checked units
    single = host.to_device(0)
    sharded = host.to_devices(n_ranges, devices=_devices(n_ranges), typical_len=200)
    ranges = sharded.ranges()
    assert ranges[0][1] == 0 and ranges[-1][2] == host.n_modules and all(a[2] == c[1] for a, c in zip(ranges, ranges[1:]))
    # one sequence
    want = single.search(seqs[0], masks[0])
    got = sharded.search(seqs[0], masks[0])
    assert got.n_sites == want.n_sites and got.n_words == want.n_words and (want.n_sites > 0 or kind == "rin")
    assert np.array_equal(got.records(), want.records())
    # a batch: canonical order is (sequence, module, placement)
    batch = b.HostBatch(seqs, masks)
    want_b = single.search_host(batch)
    got_b = sharded.search_host(batch)
    assert np.array_equal(got_b.seq_word_begin(), want_b.seq_word_begin())
    assert np.array_equal(got_b.records(), want_b.records())
    for s in (want, got, want_b, got_b):
        s.close()
    sharded.close()
    single.close()


def test_sharded_headline_shape_and_reuse():
    """A slice of the headline workload (2000-nt query, banded mask) over 2 ranges, searched twice (buffers recycled)."""
    host = b.HostLibrary.from_json_modules(synth.json_modules(4000, 104, "mixed"))
    seq = synth.sequence(2000, 4)
    mask = synth.mask_canonical(seq, seed=204, keep=0.3, span=100)
    single = host.to_device(0)
    want = single.search(seq, mask)
    sharded = host.to_devices(2, devices=_devices(2), typical_len=2000)
    for _ in range(2):
        got = sharded.search(seq, mask)
        assert np.array_equal(got.records(), want.records()) and got.n_sites == want.n_sites > 0
        got.close()
    per_range = [sharded.shard_stats(r)["n_sites"] for r in range(2)]
    assert sum(per_range) == want.n_sites
    want.close()
    sharded.close()
    single.close()


def test_sharded_rejects_bad_arguments():
    host = b.HostLibrary.from_json_modules(synth.json_modules(20, 1, "c1"))
    with pytest.raises(b.SearchError):
        host.to_devices(2, devices=[0, 99])
    with pytest.raises(b.SearchError):
        host.to_devices(0)
