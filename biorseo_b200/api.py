"""Thin Python view of the C ABI: host library (loaders), device library, queries, searches."""
import ctypes as C

import numpy as np

from . import capi


class SearchError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[{code}] {message}")
        self.code = code


def _check(rc, host=False):
    if rc != 0:
        lib = capi.load()
        msg = (lib.bsh_last_error() if host else lib.bss_last_error()) or b""
        raise SearchError(rc, msg.decode("utf-8", "replace"))


def pack_mask(allowed):
    """Boolean n x n matrix (upper triangle used) -> n x ceil(n/32) little-endian u32 words."""
    allowed = np.asarray(allowed, dtype=bool)
    n = allowed.shape[0]
    W = (n + 31) // 32
    bits = np.zeros((n, W * 32), dtype=np.uint8)
    bits[:, :n] = np.triu(allowed, 1)
    return np.packbits(bits, axis=1, bitorder="little").view("<u4").reshape(n, W).copy()


def canonical_pair_mask(seq, span=100, keep=None, rng=None):
    """Synthetic stand-in for `pij > theta` (the mask the reference derives from ViennaRNA,
    cppsrc/MOIP.cpp:77-90): canonical pairs with 4 <= v-u < span and u < n-6, optionally
    thinned by Bernoulli(keep)."""
    s = np.frombuffer(seq.encode() if isinstance(seq, str) else bytes(seq), dtype=np.uint8)
    n = len(s)
    pair = np.zeros((256, 256), dtype=bool)
    for a, b in ("AU", "UA", "GC", "CG", "GU", "UG"):
        pair[ord(a), ord(b)] = True
    m = pair[s[:, None], s[None, :]]
    u, v = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    m &= (v - u >= 4) & (v - u < span) & (u < n - 6)
    if keep is not None:
        m &= rng.random((n, n)) < keep
    return pack_mask(m)


class HostLibrary:
    """Parsed + validated motif library (C++ loaders), owner of the flat component table."""

    def __init__(self, handle):
        self._lib = capi.load()
        self._h = handle

    @classmethod
    def load(cls, source, path):
        lib = capi.load()
        kind = capi.KIND_OF_SOURCE[source] if isinstance(source, str) else source
        h = C.c_void_p()
        _check(lib.bsh_library_load(kind, str(path).encode(), C.byref(h)), host=True)
        return cls(h)

    @classmethod
    def load_compiled(cls, path):
        """Maps a library written by save_compiled(): the flat table is used in place, nothing is parsed."""
        lib = capi.load()
        h = C.c_void_p()
        _check(lib.bsh_library_load_compiled(str(path).encode(), C.byref(h)), host=True)
        return cls(h)

    def save_compiled(self, path):
        _check(self._lib.bsh_library_save_compiled(self._h, str(path).encode()), host=True)
        return path

    @property
    def kind(self):
        return self._lib.bsh_library_kind(self._h)

    @classmethod
    def from_json_modules(cls, modules):
        """modules: iterable of (key, sequence, struct2d), in canonical (key-sorted) order."""
        lib = capi.load()
        h = C.c_void_p()
        _check(lib.bsh_library_new(capi.KIND_JSON, C.byref(h)), host=True)
        for key, seq, ss in modules:
            lib.bsh_library_add_json(h, key.encode(), seq.encode(), ss.encode())
        lib.bsh_library_seal(h)
        return cls(h)

    def close(self):
        if self._h:
            self._lib.bsh_library_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_modules(self):
        return self._lib.bsh_library_modules(self._h)

    @property
    def mapped_bytes(self):
        """Size of the file mapping the flat table lives in (a library from load_compiled), else 0."""
        return int(self._lib.bsh_library_mapped_bytes(self._h))

    @property
    def n_seen(self):
        return self._lib.bsh_library_seen(self._h)

    @property
    def delay(self):
        return self._lib.bsh_library_delay(self._h)

    def rejected(self):
        out = []
        for i in range(self._lib.bsh_library_rejected(self._h)):
            ident, code = C.c_char_p(), C.c_char()
            self._lib.bsh_library_reject(self._h, i, C.byref(ident), C.byref(code))
            out.append((ident.value.decode("utf-8", "replace"), code.value.decode("latin1")))
        return out

    def module_id(self, m):
        return self._lib.bsh_library_module_id(self._h, m).decode("utf-8", "replace")

    def module_components(self):
        return np.array([self._lib.bsh_library_module_components(self._h, m) for m in range(self.n_modules)], dtype=np.int64)

    def raw_table(self):
        t = capi.BssTable()
        _check(self._lib.bsh_library_table(self._h, C.byref(t)), host=True)
        return t

    def table(self):
        """Flat component table as numpy arrays (copies)."""
        t = self.raw_table()
        n_all = t.n_units + t.n_gates

        def arr(ptr, count, dtype):
            if count == 0:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)

        ucb = arr(t.unit_comp_begin, n_all + 1, np.uint32)
        n_comps = int(ucb[-1]) if n_all else 0
        cpb = arr(t.comp_pat_begin, n_comps + 1, np.uint32)
        uck = arr(t.unit_check_begin, n_all + 1, np.uint32)
        n_checks = int(uck[-1]) if n_all else 0
        return {
            "n_modules": t.n_modules, "n_units": t.n_units, "n_gates": t.n_gates,
            "unit_module": arr(t.unit_module, n_all, np.uint32), "unit_delay": arr(t.unit_delay, n_all, np.uint32),
            "unit_gate": arr(t.unit_gate, t.n_units, np.uint32), "unit_comp_begin": ucb, "comp_pat_begin": cpb,
            "pattern": arr(t.pattern, int(cpb[-1]) if n_comps else 0, np.uint8), "unit_check_begin": uck,
            "checks": arr(t.checks, 4 * n_checks, np.int16).reshape(-1, 4),
        }

    def records_text(self, records):
        """u32 records -> list of 'identifier\\tcomponents\\tlinks' lines (the oracle's dump format)."""
        records = np.ascontiguousarray(records, dtype=np.uint32)
        text, length = C.c_void_p(), C.c_uint64()
        _check(self._lib.bsh_records_text(self._h, records.ctypes.data_as(C.c_void_p), records.size, C.byref(text), C.byref(length)), host=True)
        try:
            s = C.string_at(text, length.value).decode("utf-8", "replace")
        finally:
            self._lib.bsh_text_free(text)
        return s.splitlines()

    def to_device(self, device=-1):
        h = C.c_void_p()
        rc = self._lib.bsh_device_library(self._h, device, C.byref(h))
        if rc != 0:
            raise SearchError(rc, (self._lib.bsh_last_error() or b"").decode())
        return DeviceLibrary(h, self)

    def shard_ranges(self, n_shards, typical_len=0):
        """Contiguous module ranges [(begin, end)] balanced by search cost (bss_shard_ranges: host only)."""
        t = self.raw_table()
        begin = (C.c_uint32 * (n_shards + 1))()
        _check(self._lib.bss_shard_ranges(C.byref(t), n_shards, typical_len, begin))
        return [(int(begin[r]), int(begin[r + 1])) for r in range(n_shards)]

    def to_devices(self, n_gpus, devices=None, typical_len=0):
        """The library cut over n_gpus GPUs of this box, driven by ONE process (bss_sharded_create)."""
        t = self.raw_table()
        h = C.c_void_p()
        dev = (C.c_int * n_gpus)(*devices) if devices is not None else None
        _check(self._lib.bss_sharded_create(C.byref(t), dev, n_gpus, typical_len, C.byref(h)))
        return ShardedLibrary(h, self)


class ShardedLibrary:
    """One library over several GPUs in one process (include/biorseo_b200.h, bss_sharded_*): every GPU owns a
    contiguous module range; a search returns the ranges' records as one stream in canonical order."""

    def __init__(self, handle, host):
        self._lib = capi.load()
        self._h = handle
        self.host = host
        self.n_gpus = int(self._lib.bss_sharded_gpus(handle))

    def ranges(self):
        out = []
        for r in range(self.n_gpus):
            d, b, e = C.c_int(), C.c_uint32(), C.c_uint32()
            _check(self._lib.bss_sharded_range(self._h, r, C.byref(d), C.byref(b), C.byref(e)))
            out.append((d.value, b.value, e.value))
        return out

    def search_host(self, batch):
        """Host buffers in (HostBatch), host records out: bss_search_sharded."""
        h = C.c_void_p()
        _check(self._lib.bss_search_sharded(self._h, batch.n_seqs, batch.seqs.ctypes.data_as(C.c_void_p), batch.seq_begin.ctypes.data_as(capi.u64p),
                                            batch.masks.ctypes.data_as(C.c_void_p), batch.mask_begin.ctypes.data_as(capi.u64p), C.byref(h)))
        return Sites(h, batch.n_seqs, owner=self)

    def search(self, seq, mask):
        return self.search_host(HostBatch([seq], [mask]))

    def shard_stats(self, shard):
        st = capi.BssStats()
        _check(self._lib.bss_last_stats(C.c_void_p(self._lib.bss_sharded_library(self._h, shard)), C.byref(st)))
        return {name: getattr(st, name) for name, _ in capi.BssStats._fields_}

    def close(self):
        if self._h:
            self._lib.bss_sharded_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_fasta(path):
    """Records of a (multi-record) FASTA file as (name, sequence, structure) tuples: the C++ reader of the batch
    front-end (biorseo_b200/host/fasta.h), same rules as the reference's Fasta::load."""
    lib = capi.load()
    text, length = C.c_void_p(), C.c_uint64()
    _check(lib.bsh_fasta_text(str(path).encode(), C.byref(text), C.byref(length)), host=True)
    try:
        s = C.string_at(text, length.value).decode("utf-8", "replace")
    finally:
        lib.bsh_text_free(text)
    return [tuple(line.split("\t")) for line in s.split("\n")[:-1]]


class Query:
    """Sequences + pair masks resident in device memory."""

    def __init__(self, dev_lib, seqs, masks):
        lib = capi.load()
        seqs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        self.n_seqs = len(seqs)
        self.lengths = np.array([len(s) for s in seqs], dtype=np.uint64)
        self._seq_begin = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.uint64)
        self._seqs = np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8).copy()
        masks = [np.ascontiguousarray(m, dtype=np.uint32).reshape(-1) for m in masks]
        self._mask_begin = np.concatenate([[0], np.cumsum([m.size for m in masks])]).astype(np.uint64)
        self._masks = np.concatenate(masks + [np.zeros(1, np.uint32)])
        self.h2d_bytes = int(self._seqs.size - 1 + 4 * (self._masks.size - 1) + 16 * (self.n_seqs + 1))
        self._lib = lib
        self._h = C.c_void_p()
        _check(lib.bss_query_create(dev_lib._h, self.n_seqs, self._seqs.ctypes.data_as(C.c_void_p),
                                    self._seq_begin.ctypes.data_as(capi.u64p), self._masks.ctypes.data_as(C.c_void_p),
                                    self._mask_begin.ctypes.data_as(capi.u64p), C.byref(self._h)))

    def close(self):
        if self._h:
            self._lib.bss_query_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostBatch:
    """Query batch packed once into the flat host buffers the C ABI takes (bss_search_batch):
    concatenated sequences + offsets, concatenated pair masks + offsets. With pinned=True the two
    large buffers live in page-locked memory (torch), so every search uploads them by DMA."""

    def __init__(self, seqs, masks, pinned=False):
        seqs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        self.n_seqs = len(seqs)
        self.seq_begin = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
        blob = np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8)
        flat = [np.ascontiguousarray(m, dtype=np.uint32).reshape(-1) for m in masks]
        self.mask_begin = np.concatenate([[0], np.cumsum([m.size for m in flat])]).astype(np.uint64)
        words = int(self.mask_begin[-1])
        if pinned:
            import torch
            self._keep = (torch.empty(blob.size, dtype=torch.uint8, pin_memory=True),
                          torch.empty(words + 1, dtype=torch.int32, pin_memory=True))
            self.seqs = self._keep[0].numpy()
            self.masks = self._keep[1].numpy().view(np.uint32)
        else:
            self.seqs, self.masks = np.empty(blob.size, np.uint8), np.empty(words + 1, np.uint32)
        self.seqs[:] = blob
        at = 0
        for m in flat:
            self.masks[at:at + m.size] = m
            at += m.size
        self.masks[words] = 0
        self.h2d_bytes = int(blob.size - 1 + 4 * words + 16 * (self.n_seqs + 1))


class Sites:
    def __init__(self, handle, n_seqs, owner=None):
        self._lib = capi.load()
        self._h = handle
        self._owner = owner   # keeps the DeviceLibrary alive while its result is
        self.n_sites = int(self._lib.bss_sites_count(handle))
        self.n_words = int(self._lib.bss_sites_words(handle))
        self.n_seqs = n_seqs

    def records(self):
        ptr = self._lib.bss_sites_records(self._h)
        if not ptr or self.n_words == 0:
            return np.zeros(0, dtype=np.uint32)
        return np.ctypeslib.as_array(ptr, shape=(self.n_words,)).copy()

    def device_pointer(self):
        return self._lib.bss_sites_device_records(self._h)

    def seq_word_begin(self):
        ptr = self._lib.bss_sites_seq_word_begin(self._h)
        return np.ctypeslib.as_array(ptr, shape=(self.n_seqs + 1,)).copy()

    def close(self):
        if self._h:
            self._lib.bss_sites_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SiteIndex:
    """Insertion-variable index of one searched sequence (bss_site_index_build): the C_{x,i,p} variables,
    the c3 term counts, the per-nucleotide overlap lists (c4) and the pair adjacency, as u32 arrays."""

    def __init__(self, handle, owner=None):
        self._lib = capi.load()
        self._h = handle
        self._owner = owner
        self.build_ms = float(self._lib.bss_site_index_ms(handle))
        self.kernel_launches = int(self._lib.bss_site_index_launches(handle))

    def count(self, section):
        return int(self._lib.bss_site_index_count(self._h, capi.IDX_SECTIONS.index(section)))

    def section(self, section):
        """Host copy of one section (requires fetch=True at build time)."""
        i = capi.IDX_SECTIONS.index(section)
        n = int(self._lib.bss_site_index_count(self._h, i))
        ptr = self._lib.bss_site_index_host(self._h, i)
        if not ptr:
            raise RuntimeError("the index was built without fetch")
        return np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].copy()

    def sections(self):
        return {name: self.section(name) for name in capi.IDX_SECTIONS}

    def device_pointer(self, section):
        return self._lib.bss_site_index_device(self._h, capi.IDX_SECTIONS.index(section))

    def nbytes(self):
        return 4 * sum(self.count(name) for name in capi.IDX_SECTIONS)

    def close(self):
        if self._h:
            self._lib.bss_site_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# Developer knobs applied to every DeviceLibrary created while it is non-empty (tests force rarely taken paths
# on small inputs this way: {"heavy_ranks": 64}, {"list_cap": 40}, {"window": 0} ...; see bss_library_set_tuning).
DEFAULT_TUNING = {}


class DeviceLibrary:
    """Flat component table resident on one GPU + the search entry points."""

    def __init__(self, handle, host):
        self._lib = capi.load()
        self._h = handle
        self.host = host
        for knob, value in DEFAULT_TUNING.items():
            self.set_tuning(knob, value)

    def set_tuning(self, knob, value):
        """knob: one of capi.TUNE; value < 0 returns it to automatic."""
        _check(self._lib.bss_library_set_tuning(self._h, capi.TUNE[knob], int(value)))

    @property
    def n_window_units(self):
        return self._lib.bss_library_window_units(self._h)

    def placements_enumerated(self, query=None):
        """Exact number of placements the reference's enumerator materialises before its filters, over the whole
        query (None: the query of the last host-buffer search). A diagnostic kernel of its own."""
        v = C.c_uint64()
        _check(self._lib.bss_placements_enumerated(self._h, query._h if query is not None else None, C.byref(v)))
        return v.value

    @classmethod
    def from_table(cls, table, device=-1):
        """Builds the device library straight from flat-table arrays (dict as HostLibrary.table())."""
        lib = capi.load()
        keep = {k: np.ascontiguousarray(table[k], dtype=dt) for k, dt in
                [("unit_module", np.uint32), ("unit_delay", np.uint32), ("unit_gate", np.uint32), ("unit_comp_begin", np.uint32),
                 ("comp_pat_begin", np.uint32), ("pattern", np.uint8), ("unit_check_begin", np.uint32), ("checks", np.int16)]}
        t = capi.BssTable()
        t.n_modules, t.n_units, t.n_gates = int(table["n_modules"]), int(table["n_units"]), int(table["n_gates"])
        for name, arr in keep.items():
            ctype = {np.dtype(np.uint32): C.c_uint32, np.dtype(np.uint8): C.c_uint8, np.dtype(np.int16): C.c_int16}[arr.dtype]
            setattr(t, name, arr.ctypes.data_as(C.POINTER(ctype)) if arr.size else None)
        h = C.c_void_p()
        _check(lib.bss_library_create(C.byref(t), device, C.byref(h)))
        return cls(h, None)

    def close(self):
        if self._h:
            self._lib.bss_library_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_units(self):
        return self._lib.bss_library_units(self._h)

    def set_stream(self, cuda_stream):
        _check(self._lib.bss_library_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_module_base(self, base):
        _check(self._lib.bss_library_set_module_base(self._h, int(base)))

    def query(self, seqs, masks):
        return Query(self, seqs, masks)

    def pair_mask(self, pij, theta):
        """Packed allowed_basepair mask from an n x n float matrix (any strides), built on the device."""
        pij = np.asarray(pij, dtype=np.float32)
        n = pij.shape[0]
        assert pij.ndim == 2 and pij.shape[1] == n
        rs, cs = (st // 4 for st in pij.strides) if n > 1 else (max(n, 1), 1)
        mask = np.zeros((n, (n + 31) // 32), dtype=np.uint32)
        _check(self._lib.bss_pair_mask(self._h, pij.ctypes.data_as(C.c_void_p), n, rs, cs, float(theta), mask.ctypes.data_as(C.c_void_p)))
        return mask

    def pair_mask_device(self, pij_ptr, n, row_stride, col_stride, theta, mask_ptr, cuda_stream=None):
        _check(self._lib.bss_pair_mask_device(self._h, C.c_void_p(pij_ptr), n, row_stride, col_stride, float(theta), C.c_void_p(mask_ptr),
                                              C.c_void_p(cuda_stream) if cuda_stream else None))

    def query_pij(self, seq, pij, theta):
        """One-sequence resident query whose mask is produced on the device from the probability matrix."""
        seq = seq.encode() if isinstance(seq, str) else bytes(seq)
        pij = np.asarray(pij, dtype=np.float32)
        n = len(seq)
        assert pij.shape == (n, n)
        rs, cs = (st // 4 for st in pij.strides) if n > 1 else (max(n, 1), 1)
        q = Query.__new__(Query)
        q._lib, q._h, q.n_seqs = self._lib, C.c_void_p(), 1
        q.lengths = np.array([n], dtype=np.uint64)
        buf = np.frombuffer(seq + b"\0", dtype=np.uint8)
        _check(self._lib.bss_query_create_pij(self._h, buf.ctypes.data_as(C.c_void_p), n, pij.ctypes.data_as(C.c_void_p), rs, cs, float(theta), C.byref(q._h)))
        q.h2d_bytes = int(pij.size * 4 + n + 32)
        return q

    def search_query(self, query, fetch=True):
        h = C.c_void_p()
        _check(self._lib.bss_search_query(self._h, query._h, 1 if fetch else 0, C.byref(h)))
        return Sites(h, query.n_seqs, self)

    def search_query_into(self, query, device_ptr, capacity_words, seq_begin_ptr=None):
        """Returns (rc, n_words, n_sites); rc == BSS_ERR_OVERFLOW leaves the buffer untouched."""
        nw, ns = C.c_uint64(), C.c_uint64()
        rc = self._lib.bss_search_query_into(self._h, query._h, C.c_void_p(device_ptr), capacity_words,
                                             C.c_void_p(seq_begin_ptr) if seq_begin_ptr else None, C.byref(nw), C.byref(ns))
        if rc not in (capi.BSS_OK, capi.BSS_ERR_OVERFLOW):
            _check(rc)
        return rc, nw.value, ns.value

    def search_query_into_async(self, query, device_ptr, capacity_words, seq_begin_ptr=None):
        """Queues the whole search on the library's stream and returns at once (see search_wait)."""
        _check(self._lib.bss_search_query_into_async(self._h, query._h, C.c_void_p(device_ptr), capacity_words,
                                                     C.c_void_p(seq_begin_ptr) if seq_begin_ptr else None))

    def search_wait(self):
        """Blocks until the last queued search is complete. Returns (rc, n_words, n_sites)."""
        nw, ns = C.c_uint64(), C.c_uint64()
        rc = self._lib.bss_search_wait(self._h, C.byref(nw), C.byref(ns))
        if rc not in (capi.BSS_OK, capi.BSS_ERR_OVERFLOW):
            _check(rc)
        return rc, nw.value, ns.value

    def search(self, seq, mask):
        """One-shot: host sequence + mask in, host records out."""
        seq = seq.encode() if isinstance(seq, str) else bytes(seq)
        mask = np.ascontiguousarray(mask, dtype=np.uint32)
        buf = np.frombuffer(seq + b"\0", dtype=np.uint8)
        h = C.c_void_p()
        _check(self._lib.bss_search(self._h, buf.ctypes.data_as(C.c_void_p), len(seq), mask.ctypes.data_as(C.c_void_p), C.byref(h)))
        return Sites(h, 1, self)

    def search_batch(self, seqs, masks):
        if len(seqs) == 1:   # no concatenation of the (possibly large) mask
            return self.search(seqs[0], masks[0])
        seqs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        seq_begin = np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64)
        blob = np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8)
        masks = [np.ascontiguousarray(m, dtype=np.uint32).reshape(-1) for m in masks]
        mask_begin = np.concatenate([[0], np.cumsum([m.size for m in masks])]).astype(np.uint64)
        allm = np.concatenate(masks + [np.zeros(1, np.uint32)])
        h = C.c_void_p()
        _check(self._lib.bss_search_batch(self._h, len(seqs), blob.ctypes.data_as(C.c_void_p), seq_begin.ctypes.data_as(capi.u64p),
                                          allm.ctypes.data_as(C.c_void_p), mask_begin.ctypes.data_as(capi.u64p), C.byref(h)))
        return Sites(h, len(seqs), self)

    def search_host(self, batch):
        """One-shot on a HostBatch: upload, search, download (bss_search_batch)."""
        h = C.c_void_p()
        _check(self._lib.bss_search_batch(self._h, batch.n_seqs, batch.seqs.ctypes.data_as(C.c_void_p), batch.seq_begin.ctypes.data_as(capi.u64p),
                                          batch.masks.ctypes.data_as(C.c_void_p), batch.mask_begin.ctypes.data_as(capi.u64p), C.byref(h)))
        return Sites(h, batch.n_seqs, self)

    def site_index(self, query, records_ptr, seq=0, c3_rule=None, fetch=True):
        """Index of the sites of sequence `seq` of the LAST search on this library (made with `query`;
        None = the last host-buffer search). records_ptr: device pointer of that search's record stream.
        c3_rule defaults to the library's source (descfolder: interior nucleotides)."""
        if c3_rule is None:
            c3_rule = capi.C3_INTERIOR if (self.host is not None and self.host.kind == capi.KIND_DESC) else capi.C3_COMPONENT_LESS_LINKS
        h = C.c_void_p()
        _check(self._lib.bss_site_index_build(self._h, query._h if query is not None else None, int(seq),
                                              C.c_void_p(records_ptr) if records_ptr else None, int(c3_rule), 1 if fetch else 0, C.byref(h)))
        return SiteIndex(h, self)

    def gather_cut(self, blocks_ptr, world, block_words, header_words, out_ptr, counts_ptr, overflow_ptr, cuda_stream=None):
        _check(self._lib.bss_gather_cut(self._h, C.c_void_p(blocks_ptr), world, block_words, header_words, C.c_void_p(out_ptr), C.c_void_p(counts_ptr),
                                        C.c_void_p(overflow_ptr), C.c_void_p(cuda_stream) if cuda_stream else None))

    def measure_int_issue(self):
        """Thread-level 32-bit shift / logic operations per second this GPU sustains (SHF + LOP3 probe)."""
        v = C.c_double()
        _check(self._lib.bss_measure_int_issue(self._h, C.byref(v)))
        return v.value

    def stats(self):
        s = capi.BssStats()
        _check(self._lib.bss_last_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in capi.BssStats._fields_}
