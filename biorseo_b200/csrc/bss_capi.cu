// C ABI of the insertion-site search (include/biorseo_b200.h): library build, query
// upload, search orchestration (count -> scan -> emit), result download. No CPU search
// exists here: without a CUDA device every search entry point fails.
#include "biorseo_b200.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "bss_device.cuh"

namespace {

thread_local std::string g_error;
using bss::guarded;

int fail(int code, const std::string& what)
{
    g_error = what;
    return code;
}

int cuda_fail(cudaError_t e, const char* where)
{
    return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? BSS_ERR_NO_DEVICE : BSS_ERR_CUDA,
                std::string(where) + ": " + cudaGetErrorString(e));
}

#define BSS_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e_ = (call);                                    \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);         \
    } while (0)

// Grow-only device / pinned-host buffers.
struct DeviceBuffer {
    void*  ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

struct PinnedBuffer {
    void*  ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

}   // namespace

int bss::set_error(int code, const std::string& what) { return fail(code, what); }

struct bss_query {
    uint64_t      serial = 0;   // distinguishes queries that reuse an address (CUDA graph cache key)
    int           device = 0;
    bss::DevQuery dev{};
    DeviceBuffer  storage;
    // the occurrence table is built for one alphabet: the creating library's
    uint32_t      n_symbols = 0, pair_rows = 0;
    uint8_t       alphabet[32] = {};
    uint64_t      total_len = 0;
    uint64_t      h2d_bytes = 0;
    std::vector<uint64_t> h_seq_begin, h_mask_begin;   // host copies of the offsets (relative to the first sequence)
    int           known_band = INT32_MIN;   // widest band over the sequences, when it has been read back (resident queries)
    bss_library*  owner = nullptr;          // the library it was created on (null for the library's own one-shot query)
};

struct bss_library {
    int          device     = 0;
    int          sm_count   = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t  ev[4]      = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t  ev_band    = nullptr;   // the bands of the one-shot query have reached h_band
    bss::DevLibrary       dev{};
    DeviceBuffer          d_blob, d_unit_off, d_unit_rlen, d_unit_probe, d_unit_flags, d_prefilter;
    DeviceBuffer          d_pat_blob, d_pat_off, d_pat_tslot, d_unit_slots, d_ptab_m, d_ptab_t, d_pflag;   // pattern table: slots (library), rows (per search)
    uint32_t              n_modules = 0, n_units = 0, n_window_units = 0, n_lean_units = 0;
    bss::Tuning           tuning;
    std::vector<uint32_t> module_components;
    // work buffers, reused across searches
    DeviceBuffer d_counts, d_chunks, d_totals, d_seq_word_begin, d_alive, d_lane_counts, d_heavy, d_wlist, d_alive_w, d_lane_counts_w;
    uint32_t     n_alive = 0, n_heavy = 0, n_alive_w = 0;   // items with sites / parked heavy items / window items with sites of the last count pass
    DeviceBuffer d_placements;               // scratch of bss_placements_enumerated
    // bss_search_query_into replays its launch sequence as a CUDA graph while its arguments stay the same
    struct GraphKey {
        uint64_t query_serial = 0, capacity = 0, dev_serial = 0;
        const void *records = nullptr, *seq_out = nullptr, *stream = nullptr, *counts = nullptr, *alive = nullptr, *lanes = nullptr, *heavy = nullptr,
                   *chunk_sites = nullptr, *chunk_words = nullptr, *seq_begin = nullptr, *prefilter = nullptr, *wlist = nullptr, *alive_w = nullptr,
                   *ptab_m = nullptr, *ptab_t = nullptr, *pflag = nullptr;
        bool operator==(const GraphKey& o) const { return std::memcmp(this, &o, sizeof *this) == 0; }
    } graph_key;
    cudaGraphExec_t graph_exec = nullptr;
    // a search queued by bss_search_query_into_async and not yet waited for
    struct Pending {
        bool            active = false, emitting = false, has_records = false;
        const bss_query* q = nullptr;
        bss::Plan       plan{};
        uint64_t        capacity = 0;
        uint64_t*       seq_out = nullptr;
        bool            ordinary = true;
    } pending;
    // result handles and queries created on this library and still alive: bss_library_destroy detaches them
    std::vector<bss_sites*>      live_sites;
    std::vector<bss_site_index*> live_indexes;
    std::vector<bss_query*>      live_queries;
    bool            graphs_ok  = true;
    uint64_t        dev_serial = 1;   // bumped when launch parameters held by value change (module base, tuning)
    bss_query    oneshot;       // storage of the host-buffer entry points (bss_search, bss_search_batch), reused across calls
    PinnedBuffer h_totals, h_band, h_offsets;
    uint64_t     last_words = 0;         // record words of the previous bss_search_query: the size of the next blind download
    bool         band_pending = false;   // one-shot query: the bands are on their way to h_band (ev_band)
    // one spare set of result buffers, recycled by bss_sites_free
    DeviceBuffer spare_records;
    PinnedBuffer spare_host;
    bss_stats    stats{};
    // insertion-variable index (bss_site_index_build): scratch, and the query of the last search
    DeviceBuffer d_idx_work, d_idx_hist, spare_idx_main, spare_idx_overlap;   // spares: recycled by bss_site_index_free
    PinnedBuffer spare_idx_host_main, spare_idx_host_overlap, h_idx_totals;
    uint64_t     last_query_serial = 0;
};

struct bss_sites {
    bss_library*          lib = nullptr;
    uint64_t              n_sites = 0, n_words = 0;
    DeviceBuffer          d_records;
    PinnedBuffer          h_records;
    bool                  fetched = false;
    bool                  pooled  = false;   // h_records comes from (and returns to) the process-wide pinned pool: a result no library owns
    std::vector<uint64_t> seq_word_begin;
};

struct bss_site_index {
    bss_library* lib = nullptr;
    uint32_t     n = 0, launches = 0;
    float        ms = 0.0f;
    bool         fetched = false;
    DeviceBuffer d_main, d_overlap;
    PinnedBuffer h_main, h_overlap;
    uint64_t     offset[BSS_IDX_SECTIONS] = {}, count[BSS_IDX_SECTIONS] = {};   // in u32 elements; OVERLAP_VARS lives in its own buffer
};

namespace {

template <typename T>
void forget(std::vector<T*>& v, T* x)
{
    v.erase(std::remove(v.begin(), v.end(), x), v.end());
}

void detach_children(bss_library* lib)
{
    for (bss_sites* s : lib->live_sites) s->lib = nullptr;
    for (bss_site_index* x : lib->live_indexes) x->lib = nullptr;
    for (bss_query* q : lib->live_queries) q->owner = nullptr;
    lib->live_sites.clear();
    lib->live_indexes.clear();
    lib->live_queries.clear();
}

}   // namespace

extern "C" {

int bss_abi_version(void) { return BSS_ABI_VERSION; }

const char* bss_last_error(void) { return g_error.c_str(); }

int bss_device_count(void)
{
    return guarded([&]() -> int {
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
    });
}

int bss_library_create(const bss_table* t, int device, bss_library** out)
{
    return guarded([&]() -> int {
    if (!t || !out) return fail(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    bss::BuiltLibrary built;
    {
        std::string why;
        const int   rc = bss::build_library(t, built, why);
        if (rc != BSS_OK) return fail(rc, why);
    }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(BSS_ERR_NO_DEVICE, "no CUDA device: the insertion-site search has no CPU path");
    }
    if (device < 0) BSS_CUDA(cudaGetDevice(&device));
    if (device >= ndev) return fail(BSS_ERR_INVALID, "device index out of range");
    BSS_CUDA(cudaSetDevice(device));

    bss_library* lib = new bss_library();
    lib->device      = device;
    lib->n_modules   = t->n_modules;
    lib->n_units     = t->n_units;
    lib->n_window_units    = built.n_window_units;
    lib->module_components = built.module_components;
    auto bail = [&](int code) { bss_library_destroy(lib); return code; };
#define BSS_CUDA_LIB(call)                                                   \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) return bail(cuda_fail(e_, #call));            \
    } while (0)
    BSS_CUDA_LIB(cudaDeviceGetAttribute(&lib->sm_count, cudaDevAttrMultiProcessorCount, device));
    BSS_CUDA_LIB(cudaStreamCreateWithFlags(&lib->own_stream, cudaStreamNonBlocking));
    lib->stream = lib->own_stream;
    for (auto& ev : lib->ev) BSS_CUDA_LIB(cudaEventCreate(&ev));
    BSS_CUDA_LIB(cudaEventCreateWithFlags(&lib->ev_band, cudaEventDisableTiming));
    auto upload = [&](DeviceBuffer& d, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e_ = d.reserve(std::max<size_t>(4, bytes));
        if (e_ == cudaSuccess && bytes) e_ = cudaMemcpy(d.ptr, src, bytes, cudaMemcpyHostToDevice);
        return e_;
    };
    BSS_CUDA_LIB(upload(lib->d_blob, built.blob.data(), built.blob.size() * 4));
    BSS_CUDA_LIB(upload(lib->d_unit_off, built.unit_off.data(), built.unit_off.size() * 4));
    BSS_CUDA_LIB(upload(lib->d_unit_rlen, built.unit_rlen.data(), built.unit_rlen.size()));
    BSS_CUDA_LIB(upload(lib->d_unit_flags, built.unit_flags.data(), built.unit_flags.size()));
    BSS_CUDA_LIB(upload(lib->d_unit_probe, built.unit_probe.data(), built.unit_probe.size() * 2));
    BSS_CUDA_LIB(upload(lib->d_pat_blob, built.pat_blob.data(), built.pat_blob.size() * 4));
    BSS_CUDA_LIB(upload(lib->d_pat_off, built.pat_off.data(), built.pat_off.size() * 4));
    BSS_CUDA_LIB(upload(lib->d_pat_tslot, built.pat_tslot.data(), built.pat_tslot.size() * 2));
    BSS_CUDA_LIB(upload(lib->d_unit_slots, built.unit_slots.data(), built.unit_slots.size() * 2));
    BSS_CUDA_LIB(lib->h_totals.reserve(8 * bss::kTotalsWords));
    BSS_CUDA_LIB(lib->d_totals.reserve(8 * bss::kTotalsWords));
#undef BSS_CUDA_LIB
    lib->dev.blob       = static_cast<const uint32_t*>(lib->d_blob.ptr);
    lib->dev.unit_off   = static_cast<const uint32_t*>(lib->d_unit_off.ptr);
    lib->dev.unit_rlen  = static_cast<const uint8_t*>(lib->d_unit_rlen.ptr);
    lib->dev.unit_flags = static_cast<const uint8_t*>(lib->d_unit_flags.ptr);
    lib->dev.unit_probe = built.has_probes ? static_cast<const uint16_t*>(lib->d_unit_probe.ptr) : nullptr;
    lib->dev.pat_blob   = static_cast<const uint32_t*>(lib->d_pat_blob.ptr);
    lib->dev.pat_off    = static_cast<const uint32_t*>(lib->d_pat_off.ptr);
    lib->dev.pat_tslot  = static_cast<const uint16_t*>(lib->d_pat_tslot.ptr);
    lib->dev.unit_slots = static_cast<const uint16_t*>(lib->d_unit_slots.ptr);
    lib->dev.n_slots    = built.n_slots;
    lib->dev.n_tslots   = built.n_tslots;
    lib->n_lean_units   = built.n_lean_units;
    lib->dev.n_units    = t->n_units;
    lib->dev.n_symbols  = (uint32_t)built.alphabet.size();
    lib->dev.cmax       = built.cmax;
    lib->dev.pair_rows  = built.alphabet.size() <= bss::kPairAlphabet ? 1u : 0u;
    lib->dev.list_density = built.list_density;
    std::memset(lib->dev.alphabet, 0, sizeof lib->dev.alphabet);
    std::copy(built.alphabet.begin(), built.alphabet.end(), lib->dev.alphabet);
    *out = lib;
    return BSS_OK;
    });
}

void bss_library_destroy(bss_library* lib)
{
    if (!lib) return;
    cudaSetDevice(lib->device);
    if (lib->own_stream) cudaStreamSynchronize(lib->own_stream);
    // Result handles and queries that outlive the library keep their own buffers but lose their way back to it
    // (their free functions then release instead of recycling).
    detach_children(lib);
    if (lib->graph_exec) cudaGraphExecDestroy(lib->graph_exec);
    for (auto& ev : lib->ev)
        if (ev) cudaEventDestroy(ev);
    if (lib->ev_band) cudaEventDestroy(lib->ev_band);
    for (DeviceBuffer* b : {&lib->oneshot.storage, &lib->d_blob, &lib->d_unit_off, &lib->d_unit_rlen, &lib->d_unit_probe, &lib->d_unit_flags, &lib->d_prefilter,
                            &lib->d_pat_blob, &lib->d_pat_off, &lib->d_pat_tslot, &lib->d_unit_slots, &lib->d_ptab_m, &lib->d_ptab_t, &lib->d_pflag,
                            &lib->d_alive, &lib->d_lane_counts, &lib->d_heavy, &lib->d_counts, &lib->d_chunks, &lib->d_wlist, &lib->d_alive_w,
                            &lib->d_lane_counts_w, &lib->d_placements, &lib->d_idx_work, &lib->d_idx_hist, &lib->spare_idx_main, &lib->spare_idx_overlap,
                            &lib->d_totals, &lib->d_seq_word_begin, &lib->spare_records})
        b->release();
    for (PinnedBuffer* b : {&lib->spare_idx_host_main, &lib->spare_idx_host_overlap, &lib->h_idx_totals, &lib->h_totals, &lib->h_band, &lib->h_offsets, &lib->spare_host}) b->release();
    if (lib->own_stream) cudaStreamDestroy(lib->own_stream);
    delete lib;
}

int bss_library_set_stream(bss_library* lib, void* cuda_stream)
{
    return guarded([&]() -> int {
    if (!lib) return fail(BSS_ERR_INVALID, "null library");
    lib->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : lib->own_stream;
    return BSS_OK;
    });
}

int bss_library_set_module_base(bss_library* lib, uint32_t base)
{
    return guarded([&]() -> int {
    if (!lib) return fail(BSS_ERR_INVALID, "null library");
    lib->dev.module_base = base;
    lib->dev_serial++;
    return BSS_OK;
    });
}

int bss_library_set_tuning(bss_library* lib, int knob, int64_t value)
{
    return guarded([&]() -> int {
    if (!lib) return fail(BSS_ERR_INVALID, "null library");
    bss::Tuning& t = lib->tuning;
    switch (knob) {
    case BSS_TUNE_GROUP:
        if (value >= 0 && value != 4 && value != 8 && value != 16 && value != 32) return fail(BSS_ERR_INVALID, "group must be 4, 8, 16 or 32");
        t.group = value;
        break;
    case BSS_TUNE_UNITS_PER_CHUNK:
        if (value == 0) return fail(BSS_ERR_INVALID, "units per chunk must be at least 1");
        t.units_per_chunk = value;
        break;
    case BSS_TUNE_HEAVY_RANKS:
        if (value == 0 || value > 0xFFFFFFFFll) return fail(BSS_ERR_INVALID, "heavy-item threshold out of range");
        t.heavy_ranks = value;
        break;
    case BSS_TUNE_LIST_CAP:
        if (value > 4096) return fail(BSS_ERR_INVALID, "list capacity above 4096");
        t.list_cap = value;
        break;
    case BSS_TUNE_WINDOW: t.window = value; break;
    case BSS_TUNE_WINDOW_STAGE: t.window_stage = value; break;
    case BSS_TUNE_WINDOW_DRAW:
        if (value == 0 || value > 1024) return fail(BSS_ERR_INVALID, "draw block out of range");
        t.window_draw = value;
        break;
    case BSS_TUNE_PATTERN_TABLE: t.dict = value; break;
    case BSS_TUNE_PHASE_EVENTS: t.phase_events = value; break;
    default: return fail(BSS_ERR_INVALID, "unknown tuning knob");
    }
    lib->dev_serial++;   // a cached launch graph was built for the old plan
    return BSS_OK;
    });
}

uint32_t bss_library_window_units(const bss_library* lib) { return lib ? lib->n_window_units : 0; }
uint32_t bss_library_units(const bss_library* lib) { return lib ? lib->n_units : 0; }
uint32_t bss_library_modules(const bss_library* lib) { return lib ? lib->n_modules : 0; }
uint32_t bss_library_module_components(const bss_library* lib, uint32_t module)
{
    return (lib && module < lib->module_components.size()) ? lib->module_components[module] : 0;
}

}   // extern "C"

namespace {

// Resident queries: the host learns the widest band of the query (it waits for the upload anyway), so that a search
// can leave out the passes no item will take and size the window kernel's shared memory.
cudaError_t read_back_band(bss_library* lib, bss_query* q)
{
    std::vector<int32_t> band(q->dev.n_seqs);
    cudaError_t          e = cudaSuccess;
    if (!band.empty()) e = cudaMemcpyAsync(band.data(), q->dev.band, 4 * band.size(), cudaMemcpyDeviceToHost, lib->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(lib->stream);
    if (e == cudaSuccess) {
        int widest = -1;
        for (int32_t b : band) widest = std::max(widest, (int)b);
        q->known_band = widest;
    }
    return e;
}

// Validates the host buffers, (re)uses q's device storage and queues upload + occurrence
// table + band on the library's stream. With `wait` the call returns once the query is resident.
int query_fill(bss_library* lib, bss_query* q, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin, const uint32_t* masks,
               const uint64_t* mask_begin, bool wait, bool device_mask = false)
{
    if (!lib || !seq_begin || !mask_begin || (n_seqs && (!seqs || (!masks && !device_mask)))) return fail(BSS_ERR_INVALID, "null argument");
    uint32_t max_len = 0;
    for (uint32_t q = 0; q < n_seqs; q++) {
        if (seq_begin[q + 1] < seq_begin[q] || mask_begin[q + 1] < mask_begin[q]) return fail(BSS_ERR_INVALID, "non-monotonic offsets");
        const uint64_t n = seq_begin[q + 1] - seq_begin[q];
        if (n > BSS_MAX_SEQ_LEN) return fail(BSS_ERR_LIMIT, "sequence longer than BSS_MAX_SEQ_LEN");
        if (mask_begin[q + 1] - mask_begin[q] != n * ((n + 31) / 32)) return fail(BSS_ERR_INVALID, "pair mask is not n rows of ceil(n/32) words");
        max_len = std::max<uint32_t>(max_len, (uint32_t)n);
    }
    const uint64_t total_len = n_seqs ? seq_begin[n_seqs] - seq_begin[0] : 0, total_mask = n_seqs ? mask_begin[n_seqs] - mask_begin[0] : 0;
    const char* s0 = seqs + (n_seqs ? seq_begin[0] : 0);
    if (std::memchr(s0, '\n', total_len) || std::memchr(s0, '\r', total_len))
        return fail(BSS_ERR_BAD_SEQUENCE, "query contains a line terminator");

    BSS_CUDA(cudaSetDevice(lib->device));
    static uint64_t next_serial = 0;
    q->serial    = ++next_serial;
    q->device    = lib->device;
    // one allocation: [seq_begin | mask_begin | band | occurrence table | banded masks | masks | seqs]
    const uint32_t ns = lib->dev.n_symbols, occ_rows = ns + (lib->dev.pair_rows ? ns * ns : 0u);
    const uint32_t occ_stride = (max_len + 32) / 32 + bss::kOccPad;
    const bool     with_kmers = lib->dev.unit_probe != nullptr;
    const size_t   occ_words  = (size_t)n_seqs * occ_rows * occ_stride + (with_kmers ? (size_t)n_seqs * bss::kKmerWords : 0);
    // banded masks for the window walk: only where whole warps work on an item (make_plan's rule for 32-lane groups)
    const uint64_t n_items_q   = (uint64_t)n_seqs * lib->n_units;
    const bool     with_bmask  = lib->n_window_units > 0 && ((max_len + 32) / 32 > 16 || n_items_q <= (1u << 17));
    const size_t   bmask_words = with_bmask ? (size_t)bss::band_stride(max_len) * total_len : 0;
    const size_t   off_sb = 0, off_mb = off_sb + 8 * (size_t)(n_seqs + 1), off_band = off_mb + 8 * (size_t)(n_seqs + 1),
                 off_occ = off_band + 4 * (((size_t)n_seqs + 3) & ~(size_t)3), off_bmask = off_occ + 4 * ((occ_words + 3) & ~(size_t)3),
                 off_mask = off_bmask + 4 * ((bmask_words + 3) & ~(size_t)3), off_seq = off_mask + 4 * total_mask, bytes = off_seq + total_len + 16;
    cudaError_t e = q->storage.reserve(bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(query)");
    char* d = static_cast<char*>(q->storage.ptr);
    std::vector<uint64_t> sb(n_seqs + 1), mb(n_seqs + 1);
    for (uint32_t i = 0; i <= n_seqs; i++) { sb[i] = seq_begin[i] - seq_begin[0]; mb[i] = mask_begin[i] - mask_begin[0]; }
    cudaStream_t st = lib->stream;
    e = cudaMemcpyAsync(d + off_sb, sb.data(), 8 * (size_t)(n_seqs + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + off_mb, mb.data(), 8 * (size_t)(n_seqs + 1), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && total_mask && !device_mask) e = cudaMemcpyAsync(d + off_mask, masks + mask_begin[0], 4 * total_mask, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && total_len) e = cudaMemcpyAsync(d + off_seq, s0, total_len, cudaMemcpyHostToDevice, st);
    q->dev.seq_begin  = reinterpret_cast<const uint64_t*>(d + off_sb);
    q->dev.mask_begin = reinterpret_cast<const uint64_t*>(d + off_mb);
    q->dev.band       = reinterpret_cast<const int32_t*>(d + off_band);
    q->dev.masks      = reinterpret_cast<const uint32_t*>(d + off_mask);
    q->dev.seqs       = reinterpret_cast<const uint8_t*>(d + off_seq);
    q->dev.occ        = reinterpret_cast<const uint32_t*>(d + off_occ);
    q->dev.kmers      = with_kmers ? q->dev.occ + (size_t)n_seqs * occ_rows * occ_stride : nullptr;
    q->dev.bmask      = with_bmask ? reinterpret_cast<const uint32_t*>(d + off_bmask) : nullptr;
    q->dev.bm_stride  = bss::band_stride(max_len);
    q->dev.occ_rows   = occ_rows;
    q->dev.occ_stride = occ_stride;
    q->dev.n_seqs     = n_seqs;
    q->dev.max_len    = max_len;
    q->n_symbols      = ns;
    q->pair_rows      = lib->dev.pair_rows;
    std::memcpy(q->alphabet, lib->dev.alphabet, sizeof q->alphabet);
    // widest pair the mask allows, per sequence: bounds the window a tied component is searched in
    // The host-buffer entry points (the library's own one-shot query) learn the widest band without an extra wait: the bands
    // are copied to pinned memory right behind the band kernel, and the host reads them while the banded masks and the
    // occurrence table are still being built (search_host_batch). A search then leaves out the passes no item will take.
    int32_t* band_host = nullptr;
    if (q == &lib->oneshot && !wait && !device_mask && n_seqs && lib->h_band.reserve(4 * (size_t)n_seqs) == cudaSuccess) band_host = static_cast<int32_t*>(lib->h_band.ptr);
    lib->band_pending = false;
    if (e == cudaSuccess && !device_mask) {   // (a device-built mask gets its band afterwards)
        e = bss::launch_band(q->dev, reinterpret_cast<int32_t*>(d + off_band), const_cast<uint32_t*>(q->dev.bmask), band_host, lib->ev_band, st);
        lib->band_pending = e == cudaSuccess && band_host != nullptr;
    }
    // per-symbol (and per-symbol-pair) occurrence bit rows of every sequence, matched against by every unit
    if (e == cudaSuccess) e = bss::launch_occ(lib->dev, q->dev, reinterpret_cast<uint32_t*>(d + off_occ), const_cast<uint32_t*>(q->dev.kmers), st);
    // (copies from pageable host memory are staged before cudaMemcpyAsync returns: sb / mb may go out of scope)
    q->known_band = INT32_MIN;
    if (e == cudaSuccess && wait && !device_mask) e = read_back_band(lib, q);
    else if (e == cudaSuccess && wait) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "query upload");
    q->total_len      = total_len;
    q->h2d_bytes      = 16 * (uint64_t)(n_seqs + 1) + 4 * total_mask + total_len;
    q->h_seq_begin    = sb;
    q->h_mask_begin   = mb;
    return BSS_OK;
}

}   // namespace

extern "C" {

namespace {

int check_pij_args(const bss_library* lib, const void* pij, uint32_t n, uint64_t row_stride, uint64_t col_stride)
{
    if (!lib || (n && !pij)) return fail(BSS_ERR_INVALID, "null argument");
    if (n > BSS_MAX_SEQ_LEN) return fail(BSS_ERR_LIMIT, "sequence longer than BSS_MAX_SEQ_LEN");
    if (n > 1 && (row_stride == 0 || col_stride == 0)) return fail(BSS_ERR_INVALID, "zero stride");
    return BSS_OK;
}

// elements spanned by an n x n matrix with these strides
uint64_t pij_extent(uint32_t n, uint64_t row_stride, uint64_t col_stride) { return n ? (uint64_t)(n - 1) * (row_stride + col_stride) + 1 : 0; }

}   // namespace

int bss_pair_mask_device(bss_library* lib, const float* d_pij, uint32_t n, uint64_t row_stride, uint64_t col_stride, float theta,
                         uint32_t* d_mask, void* cuda_stream)
{
    return guarded([&]() -> int {
    int rc = check_pij_args(lib, d_pij, n, row_stride, col_stride);
    if (rc != BSS_OK) return rc;
    if (n && !d_mask) return fail(BSS_ERR_INVALID, "null argument");
    BSS_CUDA(cudaSetDevice(lib->device));
    BSS_CUDA(bss::launch_pair_mask(d_pij, n, row_stride, col_stride, theta, d_mask, lib->sm_count,
                                   cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : lib->stream));
    return BSS_OK;
    });
}

int bss_pair_mask(bss_library* lib, const float* pij, uint32_t n, uint64_t row_stride, uint64_t col_stride, float theta, uint32_t* mask)
{
    return guarded([&]() -> int {
    int rc = check_pij_args(lib, pij, n, row_stride, col_stride);
    if (rc != BSS_OK) return rc;
    if (n == 0) return BSS_OK;
    if (!mask) return fail(BSS_ERR_INVALID, "null argument");
    BSS_CUDA(cudaSetDevice(lib->device));
    const uint64_t extent = pij_extent(n, row_stride, col_stride), words = (uint64_t)n * ((n + 31) / 32);
    DeviceBuffer   d_pij, d_mask;
    cudaError_t    e = d_pij.reserve(extent * 4);
    if (e == cudaSuccess) e = d_mask.reserve(words * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_pij.ptr, pij, extent * 4, cudaMemcpyHostToDevice, lib->stream);
    if (e == cudaSuccess) e = bss::launch_pair_mask(static_cast<const float*>(d_pij.ptr), n, row_stride, col_stride, theta, static_cast<uint32_t*>(d_mask.ptr), lib->sm_count, lib->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mask, d_mask.ptr, words * 4, cudaMemcpyDeviceToHost, lib->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(lib->stream);
    d_pij.release();
    d_mask.release();
    if (e != cudaSuccess) return cuda_fail(e, "bss_pair_mask");
    return BSS_OK;
    });
}

int bss_query_create_pij(bss_library* lib, const char* seq, uint32_t n, const float* pij, uint64_t row_stride, uint64_t col_stride, float theta,
                         bss_query** out)
{
    return guarded([&]() -> int {
    if (!out) return fail(BSS_ERR_INVALID, "null argument");
    *out   = nullptr;
    int rc = check_pij_args(lib, pij, n, row_stride, col_stride);
    if (rc != BSS_OK) return rc;
    if (n && !seq) return fail(BSS_ERR_INVALID, "null argument");
    // the query is laid out as for a host mask (zeroes uploaded in its place), then the mask region is produced on the device
    const uint64_t        seq_begin[2]  = {0, n};
    const uint64_t        words         = (uint64_t)n * ((n + 31) / 32);
    const uint64_t        mask_begin[2] = {0, words};
    bss_query*            q             = new bss_query();
    rc = query_fill(lib, q, 1, seq, seq_begin, nullptr, mask_begin, true, /*device_mask=*/true);
    if (rc == BSS_OK && n) {
        const uint64_t extent = pij_extent(n, row_stride, col_stride);
        DeviceBuffer   d_pij;
        cudaError_t    e = d_pij.reserve(extent * 4);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_pij.ptr, pij, extent * 4, cudaMemcpyHostToDevice, lib->stream);
        if (e == cudaSuccess) e = bss::launch_pair_mask(static_cast<const float*>(d_pij.ptr), n, row_stride, col_stride, theta, const_cast<uint32_t*>(q->dev.masks), lib->sm_count, lib->stream);
        if (e == cudaSuccess) e = bss::launch_band(q->dev, const_cast<int32_t*>(q->dev.band), const_cast<uint32_t*>(q->dev.bmask), nullptr, nullptr, lib->stream);
        if (e == cudaSuccess) e = read_back_band(lib, q);
        d_pij.release();
        if (e != cudaSuccess) rc = cuda_fail(e, "bss_query_create_pij");
        q->h2d_bytes = 16 * 2 + extent * 4 + n;
    }
    if (rc != BSS_OK) {
        q->storage.release();
        delete q;
        return rc;
    }
    q->owner = lib;
    lib->live_queries.push_back(q);
    *out = q;
    return BSS_OK;
    });
}

int bss_query_create(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin, const uint32_t* masks,
                     const uint64_t* mask_begin, bss_query** out)
{
    return guarded([&]() -> int {
    if (!out) return fail(BSS_ERR_INVALID, "null argument");
    *out         = nullptr;
    bss_query* q = new bss_query();
    const int  rc = query_fill(lib, q, n_seqs, seqs, seq_begin, masks, mask_begin, true);
    if (rc != BSS_OK) {
        q->storage.release();
        delete q;
        return rc;
    }
    q->owner = lib;
    lib->live_queries.push_back(q);
    *out = q;
    return BSS_OK;
    });
}

void bss_query_destroy(bss_query* q)
{
    if (!q) return;
    cudaSetDevice(q->device);
    if (q->owner) {
        bss_library* lib = q->owner;
        forget(lib->live_queries, q);
        if (lib->pending.q == q) {   // a search of this query may still be running: let it end before its inputs go
            cudaStreamSynchronize(lib->stream);
            lib->pending.active = false;
            lib->pending.q      = nullptr;
        }
        if (lib->graph_key.query_serial == q->serial && lib->graph_exec) {
            cudaStreamSynchronize(lib->stream);
            cudaGraphExecDestroy(lib->graph_exec);
            lib->graph_exec = nullptr;
        }
    }
    q->storage.release();
    delete q;
}

}   // extern "C"

namespace {

// count + scan, then the totals come back to the host. On success `plan` and `work` describe
// the pending emit.
// Validation, work decomposition and (grow-only) work buffers of a search: no stream operation.
int prepare_search(bss_library* lib, const bss_query* q, bss::Plan& plan, bss::Work& work, bool& ordinary)
{
    if (q->n_symbols != lib->dev.n_symbols || q->pair_rows != lib->dev.pair_rows || std::memcmp(q->alphabet, lib->dev.alphabet, sizeof q->alphabet))
        return fail(BSS_ERR_INVALID, "the query was created for a library with a different pattern alphabet");
    BSS_CUDA(cudaSetDevice(lib->device));
    plan = bss::make_plan(lib->dev, q->dev, lib->sm_count, lib->tuning, lib->n_window_units, q->known_band);
    if (plan.smem_bytes_emit > 200 * 1024) return fail(BSS_ERR_LIMIT, "query too long for the shared-memory working set");
    // The ordinary passes can be left out when the host knows that every item takes the window walk (or ends in the
    // route kernel): a library of window-walk units on a query whose widest band it has read back.
    ordinary = !(plan.window && lib->n_window_units == lib->n_units && q->known_band != INT32_MIN && q->known_band <= plan.window_band);
    const uint64_t n_items = (uint64_t)q->dev.n_seqs * lib->n_units;
    if (n_items >= 0xFFFFFFFFull) return fail(BSS_ERR_LIMIT, "more than 2^32 (sequence, unit) items in one search: split the batch");
    BSS_CUDA(lib->d_counts.reserve(std::max<uint64_t>(4, n_items * 4)));
    BSS_CUDA(lib->d_alive.reserve(std::max<uint64_t>(4, n_items * 4)));
    // one row of 32 counts per alive item of the rank-space walk; beyond the cap the emit pass counts again
    const uint64_t lane_cap = plan.list_cap ? std::min<uint64_t>(n_items, 4u << 20) : 0;
    BSS_CUDA(lib->d_lane_counts.reserve(std::max<uint64_t>(4, lane_cap * 128)));
    work.alive       = static_cast<uint32_t*>(lib->d_alive.ptr);
    work.lane_counts = static_cast<uint32_t*>(lib->d_lane_counts.ptr);
    work.lane_cap    = (uint32_t)lane_cap;
    // heavy items: 1024 items x 256 blocks; [items | ranks | task sites | task lane counts]
    const size_t heavy_words = plan.list_cap ? 2 * 1024 + 1024 * 256 * (size_t)33 : 4;
    BSS_CUDA(lib->d_heavy.reserve(heavy_words * 4));
    work.heavy_items      = static_cast<uint32_t*>(lib->d_heavy.ptr);
    work.heavy_ranks      = work.heavy_items + 1024;
    work.task_sites       = work.heavy_ranks + 1024;
    work.task_lane_counts = work.task_sites + 1024 * 256;
    // chunk totals: [sites | words], zeroed per search (the passes add to them)
    BSS_CUDA(lib->d_chunks.reserve(16 * (size_t)(plan.n_chunks + 1)));
    BSS_CUDA(lib->d_seq_word_begin.reserve(8 * (size_t)(q->dev.n_seqs + 1)));
    work.prefilter = nullptr;
    if (plan.window || (lib->dev.unit_probe && q->dev.kmers)) {
        BSS_CUDA(lib->d_prefilter.reserve(4 * ((n_items + 31) / 32) + 4));
        work.prefilter = static_cast<uint32_t*>(lib->d_prefilter.ptr);
    }
    work.wlist = work.alive_w = nullptr;
    if (plan.window) {
        BSS_CUDA(lib->d_wlist.reserve(std::max<uint64_t>(4, n_items * 4)));
        BSS_CUDA(lib->d_alive_w.reserve(std::max<uint64_t>(4, n_items * 4)));
        work.wlist         = static_cast<uint32_t*>(lib->d_wlist.ptr);
        work.alive_w       = static_cast<uint32_t*>(lib->d_alive_w.ptr);
    }
    work.ptab_m = work.ptab_t = nullptr;
    work.pflag  = nullptr;
    if (plan.dict) {
        const uint64_t ns = q->dev.n_seqs;
        BSS_CUDA(lib->d_ptab_m.reserve(std::max<uint64_t>(4, 4 * ns * lib->dev.n_slots * (plan.W + plan.WS))));
        BSS_CUDA(lib->d_ptab_t.reserve(std::max<uint64_t>(4, 4 * ns * lib->dev.n_tslots * plan.W)));
        BSS_CUDA(lib->d_pflag.reserve(std::max<uint64_t>(4, ns * lib->dev.n_slots)));
        work.ptab_m = static_cast<uint32_t*>(lib->d_ptab_m.ptr);
        work.ptab_t = static_cast<uint32_t*>(lib->d_ptab_t.ptr);
        work.pflag  = static_cast<uint8_t*>(lib->d_pflag.ptr);
    }
    work.counts         = static_cast<uint32_t*>(lib->d_counts.ptr);
    work.chunk_sites    = static_cast<uint64_t*>(lib->d_chunks.ptr);
    work.chunk_words    = work.chunk_sites + plan.n_chunks + 1;
    work.totals         = static_cast<uint64_t*>(lib->d_totals.ptr);
    work.seq_word_begin = static_cast<uint64_t*>(lib->d_seq_word_begin.ptr);
    lib->stats          = bss_stats{};
    lib->last_query_serial = q->serial;
    return BSS_OK;
}

// Timing events: inside a stream capture they must be recorded as external events to stay usable.
cudaError_t record_event(bss_library* lib, int i, bool capturing)
{
    if ((i == 1 || i == 2) && lib->tuning.phase_events == 0) return cudaSuccess;   // (only the whole search is timed)
    return capturing ? cudaEventRecordWithFlags(lib->ev[i], lib->stream, cudaEventRecordExternal) : cudaEventRecord(lib->ev[i], lib->stream);
}

// Queues the count pass (+ heavy count) and the scan.
int enqueue_count(bss_library* lib, const bss_query* q, const bss::Plan& plan, const bss::Work& work, bool ordinary, bool capturing = false,
                  uint64_t* d_seq_out = nullptr)
{
    cudaStream_t st = lib->stream;
    if (!bss::reset_is_folded(lib->dev, q->dev, plan)) BSS_CUDA(bss::launch_reset(plan, work, q->dev.n_seqs, st));
    BSS_CUDA(record_event(lib, 0, capturing));
    BSS_CUDA(bss::launch_count(lib->dev, q->dev, plan, work, ordinary, lib->sm_count, st));
    BSS_CUDA(record_event(lib, 1, capturing));
    BSS_CUDA(bss::launch_scan(q->dev, plan, work, plan.n_chunks ? d_seq_out : nullptr, st));
    BSS_CUDA(record_event(lib, 2, capturing));
    return BSS_OK;
}

// What the count pass found, once the copy queued by enqueue_totals has landed.
constexpr size_t kTotalsFetch = 64;   // bytes of the totals the host reads: [0] .. [7]

// kernels of the count side of a search: route / prefilter, ordinary count, heavy count, window count, scan
uint32_t count_launches(const bss_library* lib, const bss_query* q, const bss::Plan& plan, bool ordinary)
{
    if (!plan.n_chunks) return 2u;   // reset + the scan alone
    uint32_t n = bss::reset_is_folded(lib->dev, q->dev, plan) ? 1u : 2u;   // [reset], scan
    if (plan.window || (lib->dev.unit_probe && q->dev.kmers)) n++;
    if (ordinary) n += 1u + (plan.group == 32 && plan.list_cap ? 1u : 0u);
    if (plan.window) n++;
    if (plan.dict && lib->dev.n_slots && q->dev.n_seqs) n++;   // pattern_kernel
    return n;
}

// emit, heavy emit, window emit when they are queued without knowing the sizes
uint32_t blind_emit_launches(const bss::Plan& plan, bool ordinary)
{
    return (ordinary ? 1u + (plan.group == 32 && plan.list_cap ? 1u : 0u) : 0u) + (plan.window ? 1u : 0u);
}

int parse_totals(bss_library* lib, const bss_query* q, const bss::Plan& plan, bool ordinary, uint64_t& n_words, uint64_t& n_sites)
{
    const uint64_t* tot = static_cast<const uint64_t*>(lib->h_totals.ptr);
    n_words             = tot[0];
    n_sites             = tot[1];
    lib->n_alive        = (uint32_t)tot[3];
    lib->n_heavy        = (uint32_t)std::min<uint64_t>(tot[4], 1024);
    lib->n_alive_w      = (uint32_t)tot[5];
    lib->stats.kernel_launches   = count_launches(lib, q, plan, ordinary);
    lib->stats.placements_tested = q->total_len * lib->n_units;
    lib->stats.n_sites           = n_sites;
    lib->stats.n_words           = n_words;
    lib->stats.d2h_bytes         = kTotalsFetch;
    lib->stats.window_items      = tot[6];
    lib->stats.ordinary_items    = plan.window ? tot[7] : (uint64_t)q->dev.n_seqs * lib->n_units;
    lib->stats.heavy_items       = tot[4];
    if (tot[2] & 1) return fail(BSS_ERR_COUNT, "a unit has 2^32 or more sites on one sequence");
    return BSS_OK;
}

int run_emit(bss_library* lib, const bss_query* q, const bss::Plan& plan, const bss::Work& work, bool ordinary, uint32_t* d_records, uint64_t n_words)
{
    cudaStream_t st = lib->stream;
    if (n_words && plan.n_chunks) {
        // sizes known on the host: only the passes that have something to write
        bss::Plan p = plan;
        if (!lib->n_alive_w) p.window = 0;
        const bool ord = ordinary && (lib->n_alive || lib->n_heavy);
        BSS_CUDA(bss::launch_emit(lib->dev, q->dev, p, work, d_records, lib->n_alive, lib->n_heavy, ~0ull, ord, lib->sm_count, st));
        lib->stats.kernel_launches += (ord ? 1u + (lib->n_heavy ? 1u : 0u) : 0u) + (p.window ? 1u : 0u);
    }
    BSS_CUDA(cudaEventRecord(lib->ev[3], st));
    return BSS_OK;
}

int finish_stats(bss_library* lib)
{
    BSS_CUDA(cudaEventSynchronize(lib->ev[3]));
    if (lib->tuning.phase_events != 0) {
        cudaEventElapsedTime(&lib->stats.count_ms, lib->ev[0], lib->ev[1]);
        cudaEventElapsedTime(&lib->stats.scan_ms, lib->ev[1], lib->ev[2]);
        cudaEventElapsedTime(&lib->stats.emit_ms, lib->ev[2], lib->ev[3]);
    }
    cudaEventElapsedTime(&lib->stats.total_ms, lib->ev[0], lib->ev[3]);
    return BSS_OK;
}

}   // namespace

extern "C" {

int bss_search_query(bss_library* lib, const bss_query* q, int fetch, bss_sites** out)
{
    return guarded([&]() -> int {
    if (!lib || !q || !out) return fail(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (q->device != lib->device) return fail(BSS_ERR_INVALID, "query and library live on different devices");
    bss::Plan plan;
    bss::Work work;
    uint64_t  n_words = 0, n_sites = 0;
    bool      ordinary = true;
    int       rc = prepare_search(lib, q, plan, work, ordinary);
    if (rc != BSS_OK) return rc;

    bss_sites* s = new bss_sites();
    s->lib       = lib;
    lib->live_sites.push_back(s);
    std::swap(s->d_records, lib->spare_records);
    std::swap(s->h_records, lib->spare_host);
    auto bail = [&](int code) { bss_sites_free(s); return code; };
    cudaStream_t st = lib->stream;
    cudaError_t  e  = cudaSuccess;
    // The record buffer kept from the previous search is a guess of this one's size: the emit
    // passes are queued right behind the count pass and write into it when the records fit (they
    // take their sizes from device memory), so the host's wait for the totals overlaps them.
    const uint64_t guess_words = s->d_records.cap / 4;
    bool           emitted     = false;
    rc = enqueue_count(lib, q, plan, work, ordinary);
    if (rc != BSS_OK) return bail(rc);
    if (guess_words >= 16 && plan.n_chunks) {
        e = bss::launch_emit(lib->dev, q->dev, plan, work, static_cast<uint32_t*>(s->d_records.ptr), 0xFFFFFFFFu, 0, guess_words, ordinary, lib->sm_count, st);
        if (e == cudaSuccess) e = cudaEventRecord(lib->ev[3], st);
        if (e != cudaSuccess) return bail(cuda_fail(e, "emit launch"));
        emitted = true;
    }
    // A stream of similar searches (the same library on query after query) downloads blind as well: the records of as many
    // words as the previous search produced and the per-sequence offsets are queued behind the blind emit, so that the host
    // waits ONCE for totals, offsets and records; only what is missing (a larger result) takes a second copy.
    uint64_t copied_words = 0;
    bool     offsets_copied = false;
    s->seq_word_begin.assign(q->dev.n_seqs + 1, 0);
    const size_t offsets_bytes = 8 * (size_t)(q->dev.n_seqs + 1);
    if (emitted && fetch && lib->last_words && lib->last_words <= guess_words &&
        s->h_records.reserve(std::max<uint64_t>(16, lib->last_words * 4)) == cudaSuccess && lib->h_offsets.reserve(offsets_bytes) == cudaSuccess) {
        e = cudaMemcpyAsync(s->h_records.ptr, s->d_records.ptr, lib->last_words * 4, cudaMemcpyDeviceToHost, st);
        // (into pinned memory: a copy to the pageable vector would make this call wait for the whole search right here)
        if (e == cudaSuccess) e = cudaMemcpyAsync(lib->h_offsets.ptr, work.seq_word_begin, offsets_bytes, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return bail(cuda_fail(e, "records download"));
        copied_words   = lib->last_words;
        offsets_copied = true;
    }
    e = cudaMemcpyAsync(lib->h_totals.ptr, work.totals, kTotalsFetch, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(cuda_fail(e, "count pass"));
    if (offsets_copied) std::memcpy(s->seq_word_begin.data(), lib->h_offsets.ptr, offsets_bytes);
    rc = parse_totals(lib, q, plan, ordinary, n_words, n_sites);
    if (rc != BSS_OK) return bail(rc);
    s->n_words = n_words;
    s->n_sites = n_sites;
    lib->stats.d2h_bytes += copied_words * 4;
    if (emitted && n_words <= guess_words) {
        lib->stats.kernel_launches += blind_emit_launches(plan, ordinary);   // emit, heavy emit, window emit (queued blind)
    } else {   // first search of this size: allocate, then emit with the sizes now known on the host
        copied_words = 0;
        e = s->d_records.reserve(std::max<uint64_t>(16, n_words * 4));
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(records)"));
        rc = run_emit(lib, q, plan, work, ordinary, static_cast<uint32_t*>(s->d_records.ptr), n_words);
        if (rc != BSS_OK) return bail(rc);
    }
    bool more = false;
    if (fetch) {
        if (n_words * 4 > s->h_records.cap) copied_words = 0;   // (growing the pinned buffer drops what it holds)
        e = s->h_records.reserve(std::max<uint64_t>(16, n_words * 4));
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMallocHost(records)"));
        if (n_words > copied_words) {
            e = cudaMemcpyAsync(static_cast<uint32_t*>(s->h_records.ptr) + copied_words, static_cast<const uint32_t*>(s->d_records.ptr) + copied_words,
                                (n_words - copied_words) * 4, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return bail(cuda_fail(e, "records download"));
            lib->stats.d2h_bytes += (n_words - copied_words) * 4;
            more = true;
        }
        s->fetched = true;
    }
    if (!offsets_copied) {
        e = cudaMemcpyAsync(s->seq_word_begin.data(), work.seq_word_begin, 8 * (size_t)(q->dev.n_seqs + 1), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return bail(cuda_fail(e, "result download"));
        more = true;
    }
    if (more || !emitted) {
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return bail(cuda_fail(e, "result download"));
    }
    lib->stats.d2h_bytes += 8 * (uint64_t)(q->dev.n_seqs + 1);
    lib->last_words = n_words;
    rc = finish_stats(lib);
    if (rc != BSS_OK) return bail(rc);
    *out = s;
    return BSS_OK;
    });
}

int bss_search_query_into(bss_library* lib, const bss_query* q, uint32_t* d_records, uint64_t capacity_words,
                          uint64_t* d_seq_word_begin, uint64_t* n_words_out, uint64_t* n_sites_out)
{
    return guarded([&]() -> int {
    if (!n_words_out || !n_sites_out) return fail(BSS_ERR_INVALID, "null argument");
    const int rc = bss_search_query_into_async(lib, q, d_records, capacity_words, d_seq_word_begin);
    if (rc != BSS_OK) return rc;
    return bss_search_wait(lib, n_words_out, n_sites_out);
    });
}

int bss_search_query_into_async(bss_library* lib, const bss_query* q, uint32_t* d_records, uint64_t capacity_words, uint64_t* d_seq_word_begin)
{
    return guarded([&]() -> int {
    if (!lib || !q) return fail(BSS_ERR_INVALID, "null argument");
    if (q->device != lib->device) return fail(BSS_ERR_INVALID, "query and library live on different devices");
    bss::Plan plan;
    bss::Work work;
    // Everything is queued at once — count, scan, emit — and the host reads the totals only at the
    // end: the emit passes take their sizes from device memory and write nothing when the records
    // would not fit `capacity_words`. While the arguments stay the same (a query stream searched
    // into the same buffer, as in a benchmark loop or a sharded exchange) the whole sequence is
    // replayed as one CUDA graph: one launch call instead of a dozen.
    bool ordinary = true;
    int  rc = prepare_search(lib, q, plan, work, ordinary);
    if (rc != BSS_OK) return rc;
    cudaStream_t st      = lib->stream;
    const bool   emitting = plan.n_chunks && d_records && capacity_words;
    auto enqueue = [&](bool capturing) -> int {
        // The per-sequence offsets (their last entry is the number of record words) always reach the caller's header — the scan
        // writes them there — also when the records do not fit: a sharded exchange behind this search then sees the true count
        // on every rank and raises its overflow flag instead of forwarding a stale one.
        int r = enqueue_count(lib, q, plan, work, ordinary, capturing, d_seq_word_begin);
        if (r != BSS_OK) return r;
        if (emitting) BSS_CUDA(bss::launch_emit(lib->dev, q->dev, plan, work, d_records, 0xFFFFFFFFu, 0, capacity_words, ordinary, lib->sm_count, st));
        BSS_CUDA(record_event(lib, 3, capturing));
        BSS_CUDA(cudaMemcpyAsync(lib->h_totals.ptr, work.totals, kTotalsFetch, cudaMemcpyDeviceToHost, st));
        return BSS_OK;
    };
    bool launched = false;
    if (lib->graphs_ok) {
        bss_library::GraphKey key;
        key.query_serial = q->serial; key.capacity = capacity_words; key.dev_serial = lib->dev_serial;
        key.records = d_records; key.seq_out = d_seq_word_begin; key.stream = st; key.counts = work.counts; key.alive = work.alive;
        key.lanes = work.lane_counts; key.heavy = work.heavy_items; key.chunk_sites = work.chunk_sites; key.chunk_words = work.chunk_words;
        key.seq_begin = work.seq_word_begin; key.prefilter = work.prefilter; key.wlist = work.wlist; key.alive_w = work.alive_w;
        key.ptab_m = work.ptab_m; key.ptab_t = work.ptab_t; key.pflag = work.pflag;
        if (!lib->graph_exec || !(key == lib->graph_key)) {
            if (lib->graph_exec) { cudaGraphExecDestroy(lib->graph_exec); lib->graph_exec = nullptr; }
            cudaGraph_t graph = nullptr;
            cudaError_t e     = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                const int   r  = enqueue(true);
                cudaError_t e2 = cudaStreamEndCapture(st, &graph);
                e              = r != BSS_OK ? cudaErrorUnknown : e2;
            }
            if (e == cudaSuccess) e = cudaGraphInstantiate(&lib->graph_exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (e != cudaSuccess) {   // no graphs on this stream / driver: plain launches from now on
                cudaGetLastError();
                lib->graph_exec = nullptr;
                lib->graphs_ok  = false;
            } else {
                lib->graph_key = key;
            }
        }
        if (lib->graph_exec) {
            BSS_CUDA(cudaGraphLaunch(lib->graph_exec, st));
            launched = true;
        }
    }
    if (!launched) {
        rc = enqueue(false);
        if (rc != BSS_OK) return rc;
    }
    lib->pending.active      = true;
    lib->pending.emitting    = emitting;
    lib->pending.has_records = d_records != nullptr;
    lib->pending.q           = q;
    lib->pending.plan        = plan;
    lib->pending.capacity    = capacity_words;
    lib->pending.seq_out     = d_seq_word_begin;
    lib->pending.ordinary    = ordinary;
    return BSS_OK;
    });
}

int bss_search_wait(bss_library* lib, uint64_t* n_words_out, uint64_t* n_sites_out)
{
    return guarded([&]() -> int {
    if (!lib) return fail(BSS_ERR_INVALID, "null argument");
    if (!lib->pending.active) return fail(BSS_ERR_INVALID, "no search is pending on this library");
    lib->pending.active = false;
    const bss_library::Pending& p  = lib->pending;
    cudaStream_t                st = lib->stream;
    uint64_t                    n_words = 0, n_sites = 0;
    BSS_CUDA(cudaSetDevice(lib->device));
    BSS_CUDA(cudaStreamSynchronize(st));
    int rc = parse_totals(lib, p.q, p.plan, p.ordinary, n_words, n_sites);
    if (n_words_out) *n_words_out = n_words;
    if (n_sites_out) *n_sites_out = n_sites;
    if (rc != BSS_OK) return rc;
    if (n_words > p.capacity || (n_words && !p.has_records)) return fail(BSS_ERR_OVERFLOW, "output buffer too small");
    if (p.seq_out && !p.plan.n_chunks) {   // an empty library or query: nothing was queued, the offsets are all zero
        BSS_CUDA(cudaMemsetAsync(p.seq_out, 0, 8 * (size_t)(p.q->dev.n_seqs + 1), st));
        BSS_CUDA(cudaStreamSynchronize(st));
    }
    if (p.emitting) lib->stats.kernel_launches += blind_emit_launches(p.plan, p.ordinary);   // emit, heavy emit, window emit
    if (p.seq_out && p.plan.n_chunks) lib->stats.kernel_launches++;                          // offsets copy
    return finish_stats(lib);
    });
}

int bss_gather_cut(bss_library* lib, const uint32_t* d_blocks, uint32_t world, uint64_t block_words, uint32_t header_words, uint32_t* d_out,
                   uint64_t* d_counts, uint32_t* d_overflow, void* cuda_stream)
{
    return guarded([&]() -> int {
    if (!lib || !d_blocks || !d_out || !d_counts || !d_overflow) return fail(BSS_ERR_INVALID, "null argument");
    if (header_words < 2 || (header_words & 1) || (block_words & 1) || block_words < header_words)
        return fail(BSS_ERR_INVALID, "bad block layout (header and block sizes must be even: the counts are 8-byte words)");
    BSS_CUDA(cudaSetDevice(lib->device));
    BSS_CUDA(bss::launch_gather_cut(d_blocks, world, block_words, header_words, d_out, d_counts, d_overflow,
                                    cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : lib->stream));
    return BSS_OK;
    });
}

}   // extern "C"

namespace {

int search_host_batch(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin, const uint32_t* masks,
                      const uint64_t* mask_begin, int fetch, bss_sites** out)
{
    if (!out) return fail(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!lib) return fail(BSS_ERR_INVALID, "null argument");
    bss_query* q  = &lib->oneshot;   // device storage kept between calls: no allocation once it has grown
    int        rc = query_fill(lib, q, n_seqs, seqs, seq_begin, masks, mask_begin, false);
    if (rc != BSS_OK) return rc;
    if (lib->band_pending) {   // the bands have been on their way since the band kernel ended
        BSS_CUDA(cudaEventSynchronize(lib->ev_band));
        const int32_t* band = static_cast<const int32_t*>(lib->h_band.ptr);
        int            widest = -1;
        for (uint32_t i = 0; i < n_seqs; i++) widest = std::max(widest, (int)band[i]);
        q->known_band     = widest;
        lib->band_pending = false;
    }
    rc = bss_search_query(lib, q, fetch, out);
    if (rc == BSS_OK) lib->stats.h2d_bytes = q->h2d_bytes;
    return rc;
}

// A few pinned host buffers kept for the results that no library owns (bss_search_sharded): page-locking memory
// costs far more than a small search.
struct PinnedPool {
    std::mutex                mu;
    std::vector<PinnedBuffer> spare;
    PinnedBuffer take(size_t bytes)
    {
        PinnedBuffer b;
        {
            std::lock_guard<std::mutex> lock(mu);
            size_t best = spare.size();
            for (size_t i = 0; i < spare.size(); i++)
                if (spare[i].cap >= bytes && (best == spare.size() || spare[i].cap < spare[best].cap)) best = i;
            if (best < spare.size()) {
                b = spare[best];
                spare.erase(spare.begin() + (long)best);
            }
        }
        if (b.reserve(bytes) != cudaSuccess) b.release();
        return b;
    }
    void give(PinnedBuffer b)
    {
        if (!b.ptr) return;
        {
            std::lock_guard<std::mutex> lock(mu);
            if (spare.size() < 4) { spare.push_back(b); return; }
            size_t smallest = 0;
            for (size_t i = 1; i < spare.size(); i++)
                if (spare[i].cap < spare[smallest].cap) smallest = i;
            if (spare[smallest].cap < b.cap) std::swap(spare[smallest], b);
        }
        b.release();
    }
} g_pinned_pool;

}   // namespace

bss_sites* bss::new_host_sites(uint64_t n_sites, uint64_t n_words, const std::vector<uint64_t>& seq_word_begin, uint32_t** records)
{
    bss_sites* s = new bss_sites();
    s->h_records = g_pinned_pool.take(std::max<uint64_t>(16, n_words * 4));
    if (!s->h_records.ptr) {
        delete s;
        return nullptr;
    }
    s->pooled         = true;
    s->fetched        = true;
    s->n_sites        = n_sites;
    s->n_words        = n_words;
    s->seq_word_begin = seq_word_begin;
    *records          = static_cast<uint32_t*>(s->h_records.ptr);
    return s;
}

extern "C" {

int bss_search_batch(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin, const uint32_t* masks,
                     const uint64_t* mask_begin, bss_sites** out)
{
    return guarded([&]() -> int { return search_host_batch(lib, n_seqs, seqs, seq_begin, masks, mask_begin, 1, out); });
}

int bss_search_batch_resident(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin, const uint32_t* masks,
                              const uint64_t* mask_begin, bss_sites** out)
{
    return guarded([&]() -> int { return search_host_batch(lib, n_seqs, seqs, seq_begin, masks, mask_begin, 0, out); });
}

int bss_sites_copy_out(const bss_sites* s, uint64_t word_begin, uint64_t n_words, uint32_t* host_dst)
{
    return guarded([&]() -> int {
    if (!s || (!host_dst && n_words)) return fail(BSS_ERR_INVALID, "null argument");
    if (word_begin > s->n_words || n_words > s->n_words - word_begin) return fail(BSS_ERR_INVALID, "range outside the record stream");
    if (!n_words) return BSS_OK;
    if (s->fetched && s->h_records.ptr) {
        std::memcpy(host_dst, static_cast<const uint32_t*>(s->h_records.ptr) + word_begin, n_words * 4);
        return BSS_OK;
    }
    if (!s->lib || !s->d_records.ptr) return fail(BSS_ERR_INVALID, "the records are gone (library destroyed)");
    BSS_CUDA(cudaSetDevice(s->lib->device));
    BSS_CUDA(cudaMemcpyAsync(host_dst, static_cast<const uint32_t*>(s->d_records.ptr) + word_begin, n_words * 4, cudaMemcpyDeviceToHost, s->lib->stream));
    BSS_CUDA(cudaStreamSynchronize(s->lib->stream));
    s->lib->stats.d2h_bytes += n_words * 4;
    return BSS_OK;
    });
}

int bss_search(bss_library* lib, const char* seq, uint32_t n, const uint32_t* mask, bss_sites** out)
{
    return guarded([&]() -> int {
    const uint64_t seq_begin[2]  = {0, n};
    const uint64_t mask_begin[2] = {0, (uint64_t)n * ((n + 31) / 32)};
    return bss_search_batch(lib, 1, seq, seq_begin, mask, mask_begin, out);
    });
}

uint64_t        bss_sites_count(const bss_sites* s) { return s ? s->n_sites : 0; }
uint64_t        bss_sites_words(const bss_sites* s) { return s ? s->n_words : 0; }
const uint32_t* bss_sites_records(const bss_sites* s) { return (s && s->fetched) ? static_cast<const uint32_t*>(s->h_records.ptr) : nullptr; }
const uint32_t* bss_sites_device_records(const bss_sites* s) { return s ? static_cast<const uint32_t*>(s->d_records.ptr) : nullptr; }
const uint64_t* bss_sites_seq_word_begin(const bss_sites* s) { return s ? s->seq_word_begin.data() : nullptr; }

void bss_sites_free(bss_sites* s)
{
    if (!s) return;
    if (s->lib) {   // (null once the library has been destroyed: the handle then only releases its own buffers)
        cudaSetDevice(s->lib->device);
        forget(s->lib->live_sites, s);
        // keep the largest buffers around for the next search
        if (s->d_records.cap > s->lib->spare_records.cap) std::swap(s->d_records, s->lib->spare_records);
        if (s->h_records.cap > s->lib->spare_host.cap) std::swap(s->h_records, s->lib->spare_host);
    }
    s->d_records.release();
    if (s->pooled) { g_pinned_pool.give(s->h_records); s->h_records = PinnedBuffer{}; }
    s->h_records.release();
    delete s;
}

int bss_site_index_build(bss_library* lib, const bss_query* q, uint32_t seq, const uint32_t* d_records, int c3_rule, int fetch,
                         bss_site_index** out)
{
    return guarded([&]() -> int {
    if (!lib || !out) return fail(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!q) q = &lib->oneshot;   // the query of the host-buffer entry points (bss_search, bss_search_batch)
    if (q->device != lib->device) return fail(BSS_ERR_INVALID, "query and library live on different devices");
    if (seq >= q->dev.n_seqs) return fail(BSS_ERR_INVALID, "sequence index out of range");
    if (c3_rule != BSS_C3_COMPONENT_LESS_LINKS && c3_rule != BSS_C3_INTERIOR) return fail(BSS_ERR_INVALID, "unknown c3 rule");
    if (lib->pending.active) return fail(BSS_ERR_INVALID, "a search is pending on this library: call bss_search_wait first");
    if (lib->last_query_serial != q->serial || !lib->d_counts.ptr || !lib->d_seq_word_begin.ptr)
        return fail(BSS_ERR_INVALID, "the last search on this library was not made with this query");
    BSS_CUDA(cudaSetDevice(lib->device));
    cudaStream_t st = lib->stream;

    // where the sequence and its mask start (host copies of the query's offsets)
    if (q->h_seq_begin.size() != (size_t)q->dev.n_seqs + 1) return fail(BSS_ERR_INVALID, "query without host offsets");
    const uint32_t  n = (uint32_t)(q->h_seq_begin[seq + 1] - q->h_seq_begin[seq]), NU = lib->n_units;
    const uint32_t* mask = q->dev.masks + q->h_mask_begin[seq];
    if (lib->stats.n_words && !d_records) return fail(BSS_ERR_INVALID, "null record stream");
    BSS_CUDA(lib->h_idx_totals.reserve(128));
    uint64_t* tot = static_cast<uint64_t*>(lib->h_idx_totals.ptr);   // pinned: [0..7] totals, [8] first record word of the sequence

    // scratch: [unit_site_begin | unit_var_begin | totals] u64, then [row_pairs | degree | row_first_pair | pair_begin | overlap_total] u32
    const size_t n64 = 2 * (size_t)(NU + 1) + 8, n1 = (size_t)n + 1;
    BSS_CUDA(lib->d_idx_work.reserve(8 * n64 + 4 * 5 * n1));
    uint64_t* unit_site_begin = static_cast<uint64_t*>(lib->d_idx_work.ptr);
    uint64_t* unit_var_begin  = unit_site_begin + NU + 1;
    uint64_t* totals          = unit_var_begin + NU + 1;
    uint32_t* row_pairs       = reinterpret_cast<uint32_t*>(totals + 8);
    uint32_t* degree          = row_pairs + n1;
    uint32_t* row_first_tmp   = degree + n1;
    uint32_t* pair_begin_tmp  = row_first_tmp + n1;
    uint32_t* overlap_total   = pair_begin_tmp + n1;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (cudaError_t ee = cudaEventCreate(&ev0); ee != cudaSuccess) return cuda_fail(ee, "cudaEventCreate");
    if (cudaError_t ee = cudaEventCreate(&ev1); ee != cudaSuccess) {
        cudaEventDestroy(ev0);
        return cuda_fail(ee, "cudaEventCreate");
    }
    bss_site_index* x = new bss_site_index();
    x->lib = lib;
    lib->live_indexes.push_back(x);
    x->n   = n;
    std::swap(x->d_main, lib->spare_idx_main);
    std::swap(x->d_overlap, lib->spare_idx_overlap);
    std::swap(x->h_main, lib->spare_idx_host_main);
    std::swap(x->h_overlap, lib->spare_idx_host_overlap);
    auto bail = [&](int code) {
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        bss_site_index_free(x);
        return code;
    };
    cudaError_t e = cudaMemsetAsync(totals, 0, 64, st);
    if (e == cudaSuccess) e = cudaEventRecord(ev0, st);
    if (e == cudaSuccess)
        e = bss::launch_index_scan(lib->dev, static_cast<const uint32_t*>(lib->d_counts.ptr) + (size_t)seq * NU, unit_site_begin, unit_var_begin, mask, n,
                                   row_pairs, degree, row_first_tmp, pair_begin_tmp, totals, lib->sm_count, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tot, totals, 64, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tot + 8, static_cast<const uint64_t*>(lib->d_seq_word_begin.ptr) + seq, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(cuda_fail(e, "index scan"));
    const uint64_t word_at = tot[8];
    x->launches += n ? 4 : 2;
    const uint64_t S = tot[0], V = tot[1], P2 = tot[2];
    if (S >= 0xFFFFFFFFull || V >= 0xFFFFFFFFull) return bail(fail(BSS_ERR_LIMIT, "more than 2^32 sites or variables in one index"));

    uint64_t at = 0;
    auto place = [&](int section, uint64_t count) {
        x->offset[section] = at;
        x->count[section]  = count;
        at += count;
    };
    place(BSS_IDX_VAR_BEGIN, S + 1);
    place(BSS_IDX_VAR_FIRST, V);
    place(BSS_IDX_VAR_SECOND, V);
    place(BSS_IDX_VAR_C3_TERMS, V);
    place(BSS_IDX_OVERLAP_BEGIN, n1);
    place(BSS_IDX_PAIR_BEGIN, n1);
    place(BSS_IDX_ROW_FIRST_PAIR, n1);
    place(BSS_IDX_PAIR_PARTNER, P2);
    e = x->d_main.reserve(4 * at + 16);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(index)"));
    uint32_t* base = static_cast<uint32_t*>(x->d_main.ptr);
    auto dev_at = [&](int section) { return base + x->offset[section]; };
    const uint32_t n_blocks = bss::index_overlap_blocks(V, n, lib->sm_count);
    e = lib->d_idx_hist.reserve(4 * (size_t)n_blocks * n + 16);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(index histogram)"));
    uint32_t* hist = static_cast<uint32_t*>(lib->d_idx_hist.ptr);
    e = cudaMemcpyAsync(dev_at(BSS_IDX_PAIR_BEGIN), pair_begin_tmp, 4 * n1, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dev_at(BSS_IDX_ROW_FIRST_PAIR), row_first_tmp, 4 * n1, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess)
        e = bss::launch_index_vars(lib->dev, unit_site_begin, unit_var_begin, d_records ? d_records + word_at : nullptr, (uint32_t)S, (uint32_t)V, mask, n,
                                   dev_at(BSS_IDX_PAIR_BEGIN), c3_rule, dev_at(BSS_IDX_PAIR_PARTNER), dev_at(BSS_IDX_VAR_BEGIN), dev_at(BSS_IDX_VAR_FIRST),
                                   dev_at(BSS_IDX_VAR_SECOND), dev_at(BSS_IDX_VAR_C3_TERMS), hist, n_blocks, overlap_total, dev_at(BSS_IDX_OVERLAP_BEGIN),
                                   totals, lib->sm_count, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tot, totals, 64, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(cuda_fail(e, "index variables"));
    x->launches += n ? 5 : 3;
    if (tot[4] & 1) return bail(fail(BSS_ERR_INVALID, "the records are not those of the last search on this library"));
    const uint64_t E = tot[3];
    if (E >= 0xFFFFFFFFull) return bail(fail(BSS_ERR_LIMIT, "more than 2^32 overlap entries in one index"));
    x->offset[BSS_IDX_OVERLAP_VARS] = 0;
    x->count[BSS_IDX_OVERLAP_VARS]  = E;
    e = x->d_overlap.reserve(4 * E + 16);
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(overlap lists)"));
    if (E) {
        e = bss::launch_index_fill(dev_at(BSS_IDX_VAR_FIRST), dev_at(BSS_IDX_VAR_SECOND), (uint32_t)V, n, hist, n_blocks, dev_at(BSS_IDX_OVERLAP_BEGIN),
                                   static_cast<uint32_t*>(x->d_overlap.ptr), st);
        x->launches++;
    }
    if (e == cudaSuccess) e = cudaEventRecord(ev1, st);
    if (e == cudaSuccess && fetch) {
        e = x->h_main.reserve(4 * at + 16);
        if (e == cudaSuccess) e = x->h_overlap.reserve(4 * E + 16);
        if (e == cudaSuccess) e = cudaMemcpyAsync(x->h_main.ptr, x->d_main.ptr, 4 * at, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && E) e = cudaMemcpyAsync(x->h_overlap.ptr, x->d_overlap.ptr, 4 * E, cudaMemcpyDeviceToHost, st);
        x->fetched = e == cudaSuccess;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(cuda_fail(e, "index overlap lists"));
    cudaEventElapsedTime(&x->ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    *out = x;
    return BSS_OK;
    });
}

uint64_t bss_site_index_count(const bss_site_index* x, int section)
{
    return (x && section >= 0 && section < BSS_IDX_SECTIONS) ? x->count[section] : 0;
}

const uint32_t* bss_site_index_host(const bss_site_index* x, int section)
{
    if (!x || !x->fetched || section < 0 || section >= BSS_IDX_SECTIONS) return nullptr;
    const void* b = section == BSS_IDX_OVERLAP_VARS ? x->h_overlap.ptr : x->h_main.ptr;
    return static_cast<const uint32_t*>(b) + x->offset[section];
}

const uint32_t* bss_site_index_device(const bss_site_index* x, int section)
{
    if (!x || section < 0 || section >= BSS_IDX_SECTIONS) return nullptr;
    const void* b = section == BSS_IDX_OVERLAP_VARS ? x->d_overlap.ptr : x->d_main.ptr;
    return static_cast<const uint32_t*>(b) + x->offset[section];
}

float    bss_site_index_ms(const bss_site_index* x) { return x ? x->ms : 0.0f; }
uint32_t bss_site_index_launches(const bss_site_index* x) { return x ? x->launches : 0; }

void bss_site_index_free(bss_site_index* x)
{
    if (!x) return;
    if (x->lib) {   // (null once the library has been destroyed)
        cudaSetDevice(x->lib->device);
        forget(x->lib->live_indexes, x);
        // keep the largest buffers around for the next build
        if (x->d_main.cap > x->lib->spare_idx_main.cap) std::swap(x->d_main, x->lib->spare_idx_main);
        if (x->d_overlap.cap > x->lib->spare_idx_overlap.cap) std::swap(x->d_overlap, x->lib->spare_idx_overlap);
        if (x->h_main.cap > x->lib->spare_idx_host_main.cap) std::swap(x->h_main, x->lib->spare_idx_host_main);
        if (x->h_overlap.cap > x->lib->spare_idx_host_overlap.cap) std::swap(x->h_overlap, x->lib->spare_idx_host_overlap);
    }
    x->d_main.release();
    x->d_overlap.release();
    x->h_main.release();
    x->h_overlap.release();
    delete x;
}

int bss_placements_enumerated(bss_library* lib, const bss_query* q, uint64_t* placements)
{
    return guarded([&]() -> int {
    if (!lib || !placements) return fail(BSS_ERR_INVALID, "null argument");
    if (!q) q = &lib->oneshot;
    if (q->device != lib->device) return fail(BSS_ERR_INVALID, "query and library live on different devices");
    if (q->n_symbols != lib->dev.n_symbols || q->pair_rows != lib->dev.pair_rows || std::memcmp(q->alphabet, lib->dev.alphabet, sizeof q->alphabet))
        return fail(BSS_ERR_INVALID, "the query was created for a library with a different pattern alphabet");
    if (lib->pending.active) return fail(BSS_ERR_INVALID, "a search is pending on this library: call bss_search_wait first");
    BSS_CUDA(cudaSetDevice(lib->device));
    *placements = 0;
    const bss::Plan plan = bss::make_plan(lib->dev, q->dev, lib->sm_count, lib->tuning, 0, INT32_MIN);
    if (4u * (bss::kThreads / 32) * (bss::kBlobStage + 2 * lib->dev.cmax * (plan.W + plan.WS)) > 200u * 1024u)
        return fail(BSS_ERR_LIMIT, "query too long for the shared-memory working set");
    // per-warp scratch in global memory: as many warps as 256 MB allow, at most 8 per SM
    const size_t per_warp = bss::placements_words_per_warp(q->dev.max_len);
    uint32_t     ctas     = (uint32_t)lib->sm_count * 2u;
    while (ctas > 1 && (size_t)ctas * (bss::kThreads / 32) * per_warp * 8 > ((size_t)256 << 20)) ctas /= 2;
    BSS_CUDA(lib->d_placements.reserve(8 + (size_t)ctas * (bss::kThreads / 32) * per_warp * 8));
    unsigned long long* total   = static_cast<unsigned long long*>(lib->d_placements.ptr);
    uint64_t*           scratch = static_cast<uint64_t*>(lib->d_placements.ptr) + 1;
    BSS_CUDA(cudaMemsetAsync(total, 0, 8, lib->stream));
    BSS_CUDA(bss::launch_placements(lib->dev, q->dev, plan, scratch, per_warp, ctas, total, lib->stream));
    BSS_CUDA(cudaMemcpyAsync(placements, total, 8, cudaMemcpyDeviceToHost, lib->stream));
    BSS_CUDA(cudaStreamSynchronize(lib->stream));
    lib->stats.placements_enumerated = *placements;
    return BSS_OK;
    });
}

#ifdef BSS_DEBUG_ASSERT
// debug build only: the per-item site counts and the first window lane-count rows of the last search
int bss_debug_counts(bss_library* lib, uint32_t* counts, uint32_t n, uint32_t* lanes_w, uint32_t n_rows, uint32_t* alive_w)
{
    cudaStreamSynchronize(lib->stream);
    if (n && lib->d_counts.ptr) cudaMemcpy(counts, lib->d_counts.ptr, 4ull * n, cudaMemcpyDeviceToHost);
    (void)lanes_w;   // (the window walk no longer keeps per-lane counts)
    if (n_rows && lib->d_alive_w.ptr) cudaMemcpy(alive_w, lib->d_alive_w.ptr, 4ull * n_rows, cudaMemcpyDeviceToHost);
    return (int)cudaGetLastError();
}
#endif

int bss_measure_int_issue(bss_library* lib, double* thread_ops_per_s)
{
    return guarded([&]() -> int {
    if (!lib || !thread_ops_per_s) return fail(BSS_ERR_INVALID, "null argument");
    BSS_CUDA(cudaSetDevice(lib->device));
    BSS_CUDA(lib->d_idx_work.reserve(64));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    BSS_CUDA(cudaEventCreate(&e0));
    BSS_CUDA(cudaEventCreate(&e1));
    double   best = 0.0;
    uint64_t ops  = 0;
    for (int rep = 0; rep < 4; rep++) {   // the first launch warms the clocks up
        cudaError_t e = cudaEventRecord(e0, lib->stream);
        if (e == cudaSuccess) e = bss::launch_int_issue_probe(rep == 0 ? 2000 : 20000, lib->sm_count, static_cast<uint32_t*>(lib->d_idx_work.ptr), &ops, lib->stream);
        if (e == cudaSuccess) e = cudaEventRecord(e1, lib->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        float ms = 0.0f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            return cuda_fail(e, "integer-issue probe");
        }
        if (rep && ms > 0.0f) best = std::max(best, (double)ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *thread_ops_per_s = best;
    return BSS_OK;
    });
}

int bss_last_stats(const bss_library* lib, bss_stats* out)
{
    return guarded([&]() -> int {
    if (!lib || !out) return fail(BSS_ERR_INVALID, "null argument");
    *out = lib->stats;
    return BSS_OK;
    });
}

}   // extern "C"
