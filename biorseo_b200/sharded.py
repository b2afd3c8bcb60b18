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


def module_costs(table, n, alphabet=4):
    """Search-cost estimate per module of a flat component table (HostLibrary.table()) on a query of n nucleotides:
    the match phase (one shift-AND step per pattern byte per 32-position word) plus the expected number of
    placements before the pair checks (product over the components of the expected matches n / alphabet^literals),
    summed over the module's units. Only the ratios between modules matter."""
    import numpy as np
    n_units = int(table["n_units"])
    ucb = table["unit_comp_begin"].astype(np.int64)
    cpb = table["comp_pat_begin"].astype(np.int64)
    pattern = np.asarray(table["pattern"])
    words = (n + 31) // 32
    costs = np.zeros(int(table["n_modules"]), dtype=np.float64)
    for u in range(n_units):
        match, placements = 0.0, 1.0
        for c in range(int(ucb[u]), int(ucb[u + 1])):
            p = pattern[int(cpb[c]):int(cpb[c + 1])]
            literals = int((p != 0x2E).sum())
            match += 2.0 * len(p) * words
            placements *= min(float(n), n / float(alphabet) ** literals)
        costs[int(table["unit_module"][u])] += match + 2.0 * min(placements, 1e9)
    return costs


def shard_ranges_by_cost(costs, world_size):
    """Contiguous module ranges [begin, end) per rank, in canonical module order, whose summed costs are as even as a
    contiguous cut allows: rank r ends at the first module where the running cost reaches (r + 1) / world of the total."""
    import numpy as np
    costs = np.asarray(costs, dtype=np.float64)
    m = costs.size
    prefix = np.concatenate([[0.0], np.cumsum(costs)])
    total = float(prefix[-1])
    cuts = [0]
    for r in range(1, world_size):
        at = int(np.searchsorted(prefix, total * r / world_size, side="left")) if total > 0 else (m * r) // world_size
        # the cut goes before or after module at-1, whichever lands closer to the target
        if total > 0 and at > 0 and at <= m and abs(prefix[at - 1] - total * r / world_size) <= abs(prefix[min(at, m)] - total * r / world_size):
            at -= 1
        cuts.append(min(m, max(cuts[-1], at)))
    cuts.append(m)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


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


class PaddedGatherer:
    """All-gather-v for record streams whose sizes are stable from call to call (the search of a
    query stream against a sharded library): ONE collective, no size exchange ahead of it.

    Every rank contributes a block of `capacity` int32 words: a header of `header_words`
    (the search's own (n_seqs + 1) u64 word offsets, written by bss_search_query_into next to
    the records; its last entry is the number of words that follow), the records, padding.
    The search writes straight into the block (`block()`), one all_gather_into_tensor moves
    the blocks, the counts arrive inside them, and the ranks cut the blocks down to the
    canonical concatenation. When a rank's stream does not fit, every rank sees the same
    counts, grows to the same capacity and the caller searches again into the new block.
    The point-to-point `allgatherv` above is the exact-size path for very large or very
    uneven shards.
    """

    def __init__(self, device, n_seqs=1, capacity_words=1 << 14, slack=1.25, group=None):
        self.group = group
        self.header_words = 2 * (n_seqs + 1)          # (n_seqs + 1) u64 offsets
        self.slack = slack
        cap = torch.tensor([int(capacity_words) + self.header_words], dtype=torch.int64, device=device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)   # blocks must have one size on every rank
        self._allocate(device, int(cap.item()))

    def _allocate(self, device, capacity):
        world = dist.get_world_size(self.group)
        self.capacity = (capacity + 3) & ~3
        self._send = torch.zeros(self.capacity, dtype=torch.int32, device=device)
        self._recv = torch.empty(world * self.capacity, dtype=torch.int32, device=device)
        self._out = None

    def block(self):
        """(header tensor, records tensor): views of this rank's block for the search to fill."""
        return self._send[:self.header_words], self._send[self.header_words:]

    def set_count(self, n_words):
        """Only needed when the search did not fill the header itself (its stream did not fit)."""
        self._send[:self.header_words].view(torch.int64)[-1] = int(n_words)

    def exchange(self):
        """Returns (gathered, counts), or (None, counts) after growing: search again and call again."""
        world = dist.get_world_size(self.group)
        dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        blocks = self._recv.view(world, self.capacity)
        counts = blocks[:, :self.header_words].view(torch.int64)[:, -1].tolist()   # the one host sync of the exchange
        room = self.capacity - self.header_words
        if max(counts) > room:
            self._allocate(self._send.device, int(max(counts) * self.slack) + 16 + self.header_words)
            return None, counts
        h = self.header_words
        out = torch.cat([blocks[r, h:h + c] for r, c in enumerate(counts)]) if sum(counts) else self._send[:0].clone()
        return out, counts

    def exchange_async(self, dev_lib, cuda_stream=None):
        """The exchange without any host round trip (CUDA only): all-gather of the blocks, then the
        library's cut kernel writes the canonical concatenation. Returns (out, counts, overflow):
        device tensors — `out` has room for the worst case, counts[r] words belong to rank r,
        overflow != 0 means a block was too small (check it at the next point the host waits
        anyway, then grow() and repeat that search)."""
        world = dist.get_world_size(self.group)
        dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        room = self.capacity - self.header_words
        if self._out is None or self._out.numel() != world * room:
            self._out = torch.empty(world * room, dtype=torch.int32, device=self._send.device)
            self._counts = torch.zeros(world, dtype=torch.int64, device=self._send.device)
            self._overflow = torch.zeros(1, dtype=torch.int32, device=self._send.device)
        dev_lib.gather_cut(self._recv.data_ptr(), world, self.capacity, self.header_words, self._out.data_ptr(), self._counts.data_ptr(),
                           self._overflow.data_ptr(), cuda_stream)
        return self._out, self._counts, self._overflow

    def grow(self, need_words):
        self._allocate(self._send.device, int(need_words * self.slack) + 16 + self.header_words)
        self._out = None

    def gather(self, local):
        """Convenience for a stream that already exists elsewhere: copies it into the block first."""
        while True:
            n = local.numel()
            self.set_count(n)
            _, rec = self.block()
            m = min(n, rec.numel())
            if m:
                rec[:m].copy_(local[:m])
            out, counts = self.exchange()
            if out is not None:
                return out, counts


class _DevicePointer:
    """Wraps raw device memory owned by the library as a __cuda_array_interface__ object (for torch.as_tensor)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerGatherer:
    """The same exchange as PaddedGatherer.exchange_async, in ONE kernel over NVLink peer memory (bss_exchange_*,
    biorseo_b200/csrc/bss_exchange.cu): every rank's buffers are mapped into its peers through CUDA IPC once; per
    step the ranks trade their counts with peer-to-peer stores, copy their records straight into every peer's
    output at the right displacement and wait for each other's completion flags — no NCCL call, no padding, no cut
    kernel, no host round trip. torch.distributed only carries the 64-byte IPC handles at set-up.

    Raises RuntimeError (on every rank alike) when peer mapping is not possible on this box (devices hidden from
    each other, no peer access): callers then fall back to PaddedGatherer."""

    def __init__(self, device, n_seqs=1, capacity_words=1 << 14, group=None):
        import ctypes as C
        from . import capi
        self._lib = capi.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.header_words = 2 * (n_seqs + 1)
        self.device = torch.device(device)
        cap = torch.tensor([int(capacity_words)], dtype=torch.int64, device=self.device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)       # one geometry on every rank
        self._h = C.c_void_p()
        ok = self.world <= 16 and self._lib.bss_exchange_create(self.device.index, self.world, self.rank, int(cap.item()), self.header_words, C.byref(self._h)) == 0
        handle = torch.zeros(64, dtype=torch.uint8)
        if ok:
            buf = (C.c_uint8 * 64)()
            ok = self._lib.bss_exchange_handle(self._h, buf) == 0
            handle = torch.tensor(list(buf), dtype=torch.uint8)
        ok = self._all_ok(ok)
        if ok:
            every = torch.empty(self.world * 64, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(every, handle.to(self.device), group=group)
            raw = bytes(every.cpu().tolist())
            ok = self._all_ok(self._lib.bss_exchange_open(self._h, raw) == 0)
        if not ok:
            message = (self._lib.bss_last_error() or b"").decode()
            self.close()
            raise RuntimeError(f"peer-memory exchange unavailable: {message}")
        self.capacity = int(self._lib.bss_exchange_capacity(self._h))
        self._block = int(self._lib.bss_exchange_block(self._h))

    def _all_ok(self, ok):
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def block_pointers(self):
        """(header pointer, records pointer, records capacity in words): what the search fills."""
        return self._block, self._block + 4 * self.header_words, self.capacity

    def records(self, n_words):
        """This rank's own records in its block, as a tensor (for checks)."""
        return torch.as_tensor(_DevicePointer(self._block + 4 * self.header_words, n_words, "<i4"), device=self.device)

    def exchange_async(self, cuda_stream):
        """Queues the exchange kernel on `cuda_stream` (the stream of the search). Returns (out, result) device
        tensors: out = the canonical concatenation (world * capacity words of room), result = int64
        [counts of the ranks..., total words, flags (1 overflow, 2 timeout)]."""
        import ctypes as C
        out, res = C.c_void_p(), C.c_void_p()
        rc = self._lib.bss_exchange_run_async(self._h, C.c_void_p(cuda_stream), C.byref(out), C.byref(res))
        if rc != 0:
            raise RuntimeError((self._lib.bss_last_error() or b"").decode())
        # (the two output buffers and the result block never move: their tensor views are made once — wrapping raw device
        #  memory costs tens of microseconds, which is a whole exchange of a small stream)
        views = self.__dict__.setdefault("_views", {})
        key = (out.value, res.value)
        if key not in views:
            views[key] = (torch.as_tensor(_DevicePointer(out.value, self.world * self.capacity, "<i4"), device=self.device),
                          torch.as_tensor(_DevicePointer(res.value, self.world + 2, "<i8"), device=self.device))
        return views[key]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bss_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
