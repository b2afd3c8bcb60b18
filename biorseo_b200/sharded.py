"""Library-sharded search: one process per GPU (torch.distributed), contiguous module ranges.

The motif library is the natural shard axis: modules are independent given the (tiny,
replicated) query. Rank r owns the contiguous module range [base_r, base_r + M_r) in
canonical order and emits its sites already in canonical order, so concatenating the ranks'
record streams in rank order IS the canonical (module, placement) order of the whole
library. The only exchange step of the path is that concatenation: an all-gather of the
per-rank word counts followed by an all-gather-v of the records.

NCCL has no all-gather-v primitive; it is expressed as one group of point-to-point
sends/receives (every rank sends its whole stream to every peer and receives each peer's
stream at that peer's displacement), which NCCL fuses into a single grouped launch over
NVLink / NVSwitch. The same code runs on the gloo backend for the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n_modules, world_size, rank):
    """Contiguous, balanced module range of `rank`: [begin, end)."""
    base, extra = divmod(n_modules, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allgatherv(local, group=None, counts_out=None):
    """Concatenates the ranks' 1-D tensors in rank order on every rank.

    local: 1-D tensor (int32 words of records) on this rank's device. Returns
    (gathered, counts) with counts a python list of per-rank lengths.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts_t = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts_t, n_local, group=group)
    counts = counts_t.tolist()   # one host sync: displacements size the receive buffer
    offsets = [0]
    for c in counts:
        offsets.append(offsets[-1] + c)
    out = torch.empty(offsets[-1], dtype=local.dtype, device=local.device)
    if counts[rank]:
        out[offsets[rank]:offsets[rank + 1]].copy_(local)
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if counts[rank]:
            ops.append(dist.P2POp(dist.isend, local, peer, group=group))
        if counts[peer]:
            ops.append(dist.P2POp(dist.irecv, out[offsets[peer]:offsets[peer + 1]], peer, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if counts_out is not None:
        counts_out[:] = counts
    return out, counts
