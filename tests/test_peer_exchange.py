"""The one-kernel all-gather-v over NVLink peer memory (bss_exchange_*, sharded.PeerGatherer) against the NCCL
paths, on two GPUs of one box: two processes, one library shard each, several searches in a row (both output
parities), an empty shard, and a stream that does not fit (overflow flag, nothing copied). Skipped on boxes with
fewer than two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import biorseo_b200 as b
from biorseo_b200 import sharded, synth

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        modules = synth.json_modules(400, 5, "c1")
        begin, end = sharded.shard_range(len(modules), world, rank)
        host = b.HostLibrary.from_json_modules(modules[begin:end])
        dev = host.to_device(rank)
        dev.set_module_base(begin)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        dev.set_stream(stream.cuda_stream)
        peer = sharded.PeerGatherer(torch.device("cuda", rank), n_seqs=1, capacity_words=1 << 16)
        head_ptr, rec_ptr, cap = peer.block_pointers()
        results = []
        for step in range(5):
            n = 90 + 13 * step
            seq = synth.sequence(n, 70 + step)
            # step 3: rank 1 searches with an empty mask -> an empty shard stream
            mask = synth.mask_all(n) if not (step == 3 and rank == 1) else np.zeros((n, (n + 31) // 32), np.uint32)
            q = dev.query([seq], [mask])
            dev.search_query_into_async(q, rec_ptr, cap, seq_begin_ptr=head_ptr)
            out, result = peer.exchange_async(stream.cuda_stream)
            rc, n_words, _ = dev.search_wait()
            assert rc == 0
            result = result.tolist()
            assert result[world + 1] == 0, result
            assert result[rank] == n_words and result[world] == sum(result[:world])
            got = out[:result[world]].clone()
            # the NCCL reference: exact-size all-gather-v of the same local stream
            want, counts = sharded.allgatherv(peer.records(n_words).clone())
            assert counts == result[:world]
            assert torch.equal(got, want), step
            results.append(got.cpu().numpy())
            q.close()
        # a stream that does not fit: the flag is raised on every rank and nothing is copied
        small = sharded.PeerGatherer(torch.device("cuda", rank), n_seqs=1, capacity_words=8)
        h2, r2, c2 = small.block_pointers()
        seq = synth.sequence(150, 3)
        q = dev.query([seq], [synth.mask_all(150)])
        # the search refuses to write records into the small block but still leaves its word count in the header:
        # that is what tells every rank about the overflow
        dev.search_query_into_async(q, r2, c2, seq_begin_ptr=h2)
        _, result = small.exchange_async(stream.cuda_stream)
        rc, n_words, _ = dev.search_wait()
        torch.cuda.synchronize()
        assert rc == b.capi.BSS_ERR_OVERFLOW and n_words > c2, (rc, n_words)
        assert result.tolist()[world + 1] == 1
        small.close()
        q.close()
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.concatenate(results))
        peer.close()
        dev.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_exchange_equals_nccl_allgatherv(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert r0.size > 0 and np.array_equal(r0, r1)
