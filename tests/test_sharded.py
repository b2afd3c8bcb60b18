"""N > 1 host logic on CPU: world_size-2 gloo. Each rank owns a contiguous module range of
the library; its site list (here produced by the flat-table interpreter — the GPU is not
needed to test the plumbing) is exchanged with the same all-gather-v the NCCL path uses, and
the rank-order concatenation must equal the single-rank canonical list."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import biorseo_b200 as b
from biorseo_b200 import sharded, synth

import flat_interp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, seq, mask, modules, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, end = sharded.shard_range(len(modules), world, rank)
        host = b.HostLibrary.from_json_modules(modules[begin:end])
        rec = flat_interp.search_table(host.table(), seq, mask).astype(np.int64)
        # module indices are shard-local on the device; the library's module base makes them global
        comps = host.module_components()
        at = 0
        while at < rec.size:
            width = 1 + int(comps[rec[at]])
            rec[at] += begin
            at += width
        if rank == 1:
            rec = rec[:0] if os.environ.get("BSS_TEST_EMPTY_RANK") else rec
        gathered, counts = sharded.allgatherv(torch.from_numpy(rec.astype(np.int32)))
        assert sum(counts) == gathered.numel()
        # the padded single-collective exchange must agree, from proposals that differ per rank and are too small (it grows)
        padded = sharded.PaddedGatherer("cpu", n_seqs=1, capacity_words=8 + 3 * rank)
        for _ in range(2):
            again, counts2 = padded.gather(torch.from_numpy(rec.astype(np.int32)))
            assert counts2 == counts and torch.equal(again, gathered)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), gathered.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("empty_rank", [False, True])
def test_allgatherv_concatenation_is_canonical(tmp_path, empty_rank):
    seq = synth.sequence(90, 77)
    mask = synth.mask_all(90)
    modules = synth.json_modules(61, 5, "c1")
    if empty_rank:
        os.environ["BSS_TEST_EMPTY_RANK"] = "1"
    try:
        mp.spawn(_worker, args=(2, _free_port(), seq, mask, modules, str(tmp_path)), nprocs=2, join=True)
    finally:
        os.environ.pop("BSS_TEST_EMPTY_RANK", None)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1)
    host = b.HostLibrary.from_json_modules(modules)
    full = flat_interp.search_table(host.table(), seq, mask).astype(np.int32)
    if empty_rank:
        begin, _ = sharded.shard_range(len(modules), 2, 1)
        comps, at, keep = host.module_components(), 0, []
        while at < full.size:
            width = 1 + int(comps[full[at]])
            if full[at] < begin:
                keep.extend(full[at:at + width])
            at += width
        full = np.array(keep, dtype=np.int32)
    assert full.size > 0
    assert np.array_equal(r0, full)


def test_shard_ranges_partition_the_library():
    for n in (0, 1, 7, 50000, 50003):
        for world in (1, 2, 4, 8):
            ranges = [sharded.shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b_[0] for a, b_ in zip(ranges, ranges[1:]))
            sizes = [e - s for s, e in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_cost_balanced_ranges_partition_the_library_and_beat_equal_counts():
    """Library shards balanced by a search-cost estimate instead of by module count: contiguous, complete,
    and on a library whose expensive modules sit at one end, far more even than the equal-count cut."""
    cheap = synth.json_modules(300, 7, "c4a")      # long strands: few expected placements
    heavy = synth.json_modules(100, 8, "c4b")      # three short strands: thousands of expected placements
    modules = sorted(cheap + [("z" + k, s, t) for k, s, t in heavy])      # the heavy ones last in canonical order
    host = b.HostLibrary.from_json_modules(modules)
    costs = sharded.module_costs(host.table(), 2000)
    assert costs.size == host.n_modules and (costs > 0).all()
    for world in (1, 2, 4, 8):
        ranges = sharded.shard_ranges_by_cost(costs, world)
        assert ranges[0][0] == 0 and ranges[-1][1] == costs.size
        assert all(a[1] == b_[0] and a[0] <= a[1] for a, b_ in zip(ranges, ranges[1:]))
        by_cost = max(costs[s:e].sum() for s, e in ranges)
        by_count = max(costs[s:e].sum() for s, e in (sharded.shard_range(costs.size, world, r) for r in range(world)))
        assert by_cost <= by_count + 1e-9
        assert by_cost <= costs.sum() / world + costs.max() + 1e-9
    assert max(costs[s:e].sum() for s, e in sharded.shard_ranges_by_cost(costs, 4)) < 0.6 * max(
        costs[s:e].sum() for s, e in (sharded.shard_range(costs.size, 4, r) for r in range(4)))
    assert sharded.shard_ranges_by_cost(np.zeros(0), 3) == [(0, 0), (0, 0), (0, 0)]
    assert sharded.shard_ranges_by_cost(np.zeros(5), 2) == [(0, 2), (2, 5)]


def test_c_abi_cut_equals_the_python_cut(tmp_path):
    """bss_shard_ranges (the cut of the single-process bss_sharded_create) against sharded.shard_ranges_by_cost (the cut of the
    one-process-per-GPU path): the same contiguous module ranges, for JSON, RIN and DESC (gates) libraries. Host only."""
    cheap = synth.json_modules(300, 7, "c4a")
    heavy = synth.json_modules(100, 8, "c4b")
    libs = [b.HostLibrary.from_json_modules(sorted(cheap + [("z" + k, s, t) for k, s, t in heavy])),
            b.HostLibrary.load("rinfolder", synth.write_rin_folder(str(tmp_path / "rin"), 150, 3, max_components=5, min_len=1)),
            b.HostLibrary.load("descfolder", synth.write_desc_folder(str(tmp_path / "desc"), 150, 4))]
    for host in libs:
        for n in (230, 2000):
            costs = sharded.module_costs(host.table(), n)
            for world in (1, 2, 3, 4, 8):
                assert host.shard_ranges(world, n) == sharded.shard_ranges_by_cost(costs, world), (host.kind, n, world)


def test_sharded_entry_points_fail_loudly_without_a_gpu_and_validate_their_arguments():
    """There is no CPU path behind bss_sharded_create / bss_search_sharded: without a CUDA device the create call
    reports BSS_ERR_NO_DEVICE. Argument validation needs no device."""
    import ctypes as C
    from biorseo_b200 import capi
    lib = capi.load()
    host = b.HostLibrary.from_json_modules(synth.json_modules(30, 2, "c1"))
    t = host.raw_table()
    begin = (C.c_uint32 * 4)()
    assert lib.bss_shard_ranges(None, 2, 100, begin) == capi.BSS_ERR_INVALID
    assert lib.bss_shard_ranges(C.byref(t), 0, 100, begin) == capi.BSS_ERR_INVALID
    assert lib.bss_shard_ranges(C.byref(t), 3, 0, begin) == 0 and begin[0] == 0 and begin[3] == host.n_modules
    h = C.c_void_p()
    assert lib.bss_sharded_create(C.byref(t), None, 0, 100, C.byref(h)) == capi.BSS_ERR_INVALID
    assert lib.bss_sharded_create(C.byref(t), None, 17, 100, C.byref(h)) == capi.BSS_ERR_INVALID
    assert lib.bss_search_sharded(None, 0, None, None, None, None, C.byref(h)) == capi.BSS_ERR_INVALID
    assert lib.bss_sharded_gpus(None) == 0 and lib.bss_sharded_range(None, 0, None, None, None) == capi.BSS_ERR_INVALID
    lib.bss_sharded_destroy(None)
    if lib.bss_device_count() <= 0:
        rc = lib.bss_sharded_create(C.byref(t), None, 2, 100, C.byref(h))
        assert rc == capi.BSS_ERR_NO_DEVICE and not h.value
        assert b"no CUDA device" in lib.bss_last_error() or b"CUDA" in lib.bss_last_error()
