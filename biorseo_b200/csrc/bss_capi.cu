// C ABI of the insertion-site search (include/biorseo_b200.h): library build, query
// upload, search orchestration (count -> scan -> emit), result download. No CPU search
// exists here: without a CUDA device every search entry point fails.
#include "biorseo_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bss_device.cuh"

namespace {

thread_local std::string g_error;

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

bool compatible(uint8_t a, uint8_t b) { return a == '.' || b == '.' || a == b; }

// Can two matches of this pattern overlap, i.e. is the pattern compatible with itself at
// some shift 1..k-1 ?
bool can_self_overlap(const uint8_t* p, uint32_t k)
{
    for (uint32_t s = 1; s < k; s++) {
        bool ok = true;
        for (uint32_t j = 0; ok && j + s < k; j++) ok = compatible(p[j], p[j + s]);
        if (ok) return true;
    }
    return false;
}

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
};

struct bss_library {
    int          device     = 0;
    int          sm_count   = 148;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t  ev[4]      = {nullptr, nullptr, nullptr, nullptr};
    bss::DevLibrary       dev{};
    DeviceBuffer          d_blob, d_unit_off, d_unit_rlen, d_unit_probe, d_prefilter;
    uint32_t              n_modules = 0, n_units = 0;
    std::vector<uint32_t> module_components;
    // work buffers, reused across searches
    DeviceBuffer d_counts, d_chunk_sites, d_chunk_words, d_totals, d_seq_word_begin, d_alive, d_lane_counts, d_heavy;
    uint32_t     n_alive = 0, n_heavy = 0;   // items with sites / parked heavy items of the last count pass
    // bss_search_query_into replays its launch sequence as a CUDA graph while its arguments stay the same
    struct GraphKey {
        uint64_t query_serial = 0, capacity = 0, dev_serial = 0;
        const void *records = nullptr, *seq_out = nullptr, *stream = nullptr, *counts = nullptr, *alive = nullptr, *lanes = nullptr, *heavy = nullptr,
                   *chunk_sites = nullptr, *chunk_words = nullptr, *seq_begin = nullptr, *prefilter = nullptr;
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
    } pending;
    bool            graphs_ok  = true;
    uint64_t        dev_serial = 1;   // bumped when launch parameters held by value change (module base)
    bss_query    oneshot;       // storage of the host-buffer entry points (bss_search, bss_search_batch), reused across calls
    PinnedBuffer h_totals;
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
    std::vector<uint64_t> seq_word_begin;
};

extern "C" {

int bss_abi_version(void) { return BSS_ABI_VERSION; }

const char* bss_last_error(void) { return g_error.c_str(); }

int bss_device_count(void)
{
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int bss_library_create(const bss_table* t, int device, bss_library** out)
{
    if (!t || !out) return fail(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    const uint32_t n_all = t->n_units + t->n_gates;
    if (n_all && (!t->unit_module || !t->unit_delay || !t->unit_comp_begin || !t->comp_pat_begin || !t->unit_check_begin))
        return fail(BSS_ERR_INVALID, "null table array");
    if (t->n_units && !t->unit_gate) return fail(BSS_ERR_INVALID, "null unit_gate");

    // alphabet: distinct literal pattern bytes
    int code_of[256];
    std::fill(code_of, code_of + 256, -1);
    std::vector<uint8_t> alphabet;
    const uint32_t n_comps = n_all ? t->unit_comp_begin[n_all] : 0;
    const uint32_t n_pat   = n_comps ? t->comp_pat_begin[n_comps] : 0;
    if (n_pat && !t->pattern) return fail(BSS_ERR_INVALID, "null pattern array");
    for (uint32_t i = 0; i < n_pat; i++) {
        const uint8_t b = t->pattern[i];
        if (b == '.' || code_of[b] >= 0) continue;
        if (b == '\n' || b == '\r') return fail(BSS_ERR_INVALID, "pattern contains a line terminator");
        code_of[b] = (int)alphabet.size();
        alphabet.push_back(b);
    }
    if (alphabet.size() > BSS_MAX_ALPHABET) return fail(BSS_ERR_LIMIT, "more than BSS_MAX_ALPHABET distinct pattern bytes");

    // unit descriptors
    std::vector<uint32_t> blob, unit_off(n_all + 1, 0);
    std::vector<uint32_t> module_components(t->n_modules, 0);
    std::vector<float>    density;   // per searched unit: expected matches per nucleotide, all components
    std::vector<uint16_t> unit_probe(4 * (size_t)std::max<uint32_t>(1, t->n_units), 0xFFFFu);   // k-mer probes of up to four components
    uint32_t              cmax = 1;
    for (uint32_t u = 0; u < n_all; u++) {
        unit_off[u] = (uint32_t)blob.size();
        const uint32_t c0 = t->unit_comp_begin[u], c1 = t->unit_comp_begin[u + 1];
        const uint32_t k0 = t->unit_check_begin[u], k1 = t->unit_check_begin[u + 1];
        if (c1 < c0 || k1 < k0) return fail(BSS_ERR_INVALID, "non-monotonic table offsets");
        const uint32_t C = c1 - c0, NK = k1 - k0;
        if (C < 1 || C > BSS_MAX_COMPONENTS) return fail(BSS_ERR_LIMIT, "unit with 0 or more than BSS_MAX_COMPONENTS components");
        if (NK > BSS_MAX_CHECKS) return fail(BSS_ERR_LIMIT, "unit with more than BSS_MAX_CHECKS checks");
        if (NK && !t->checks) return fail(BSS_ERR_INVALID, "null checks array");
        if (t->unit_delay[u] > 0xFFFFu) return fail(BSS_ERR_LIMIT, "delay above 65535");
        if (t->unit_module[u] >= t->n_modules) return fail(BSS_ERR_INVALID, "unit_module out of range");
        uint32_t gate = bss::kNoGate;
        if (u < t->n_units) {
            gate = t->unit_gate[u];
            if (gate != bss::kNoGate && (gate < t->n_units || gate >= n_all)) return fail(BSS_ERR_INVALID, "unit_gate out of range");
            uint32_t& mc = module_components[t->unit_module[u]];
            if (mc != 0 && mc != C) return fail(BSS_ERR_INVALID, "units of one module disagree on the component count");
            mc = C;
        }
        cmax = std::max(cmax, C);

        // order checks by the level at which both ends are known
        std::vector<uint32_t> order(NK);
        std::vector<int>      level(NK);
        for (uint32_t i = 0; i < NK; i++) {
            const int16_t* c = t->checks + 4 * (size_t)(k0 + i);
            if (c[0] < -1 || c[0] >= (int)C || c[2] < -1 || c[2] >= (int)C) return fail(BSS_ERR_INVALID, "check anchor out of range");
            level[i] = std::max(0, std::max((int)c[0], (int)c[2]));
            order[i] = i;
        }
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return level[a] < level[b]; });

        // match programs: one (occurrence row, shift) step per literal symbol, or per pair of
        // adjacent literal symbols when the alphabet is small enough for pair rows
        const bool     pairs = alphabet.size() <= bss::kPairAlphabet;
        const uint32_t ns    = (uint32_t)alphabet.size();
        std::vector<uint16_t>              program;
        std::vector<std::pair<uint32_t, uint32_t>> prog_range(C);   // first step, number of steps
        for (uint32_t c = c0; c < c1; c++) {
            const uint32_t pb = t->comp_pat_begin[c], k = t->comp_pat_begin[c + 1] - pb;
            if (t->comp_pat_begin[c + 1] < pb || k > BSS_MAX_PATTERN) return fail(BSS_ERR_LIMIT, "component pattern longer than BSS_MAX_PATTERN");
            const uint8_t* p = t->pattern + pb;
            const uint32_t first_step = (uint32_t)program.size();
            for (uint32_t j = 0; j < k;) {
                if (p[j] == '.') { j++; continue; }
                uint32_t row = (uint32_t)code_of[p[j]], shift = j;
                if (pairs && j + 1 < k && p[j + 1] != '.') {
                    row = ns + row * ns + (uint32_t)code_of[p[j + 1]];
                    j += 2;
                } else {
                    j += 1;
                }
                program.push_back((uint16_t)(row | ((shift & 31u) << 7) | ((shift >> 5) << 12)));
            }
            prog_range[c - c0] = {first_step, (uint32_t)program.size() - first_step};
        }
        if (u < t->n_units) {   // uniform symbols: a pattern with l literal symbols matches one position in |alphabet|^l
            double d = 0.0;
            for (uint32_t c = c0; c < c1; c++) {
                const uint32_t pb = t->comp_pat_begin[c], k = t->comp_pat_begin[c + 1] - pb;
                double         f  = 1.0;
                for (uint32_t j = 0; j < k; j++)
                    if (t->pattern[pb + j] != '.') f /= std::max<double>(2.0, (double)alphabet.size());
                d += f;
            }
            density.push_back((float)d);
        }
        if (u < t->n_units && alphabet.size() <= 4) {   // longest window (4..6) of consecutive literal symbols of the (up to four) probed patterns
            uint32_t slot = 0;
            for (uint32_t c = c0; c < c1 && slot < 4; c++) {
                const uint32_t pb = t->comp_pat_begin[c], k = t->comp_pat_begin[c + 1] - pb;
                uint32_t       run = 0, best = 0, best_end = 0;
                for (uint32_t j = 0; j < k && best < bss::kKmerMax; j++) {
                    run = t->pattern[pb + j] == '.' ? 0 : run + 1;
                    if (std::min(run, bss::kKmerMax) > best) { best = std::min(run, bss::kKmerMax); best_end = j + 1; }
                }
                if (best < bss::kKmerMin) continue;
                uint32_t code = 0;
                for (uint32_t i = 0; i < best; i++) code |= (uint32_t)code_of[t->pattern[pb + best_end - best + i]] << (2 * i);
                unit_probe[4 * (size_t)u + slot++] = (uint16_t)(code | ((best - bss::kKmerMin) << 12));
            }
        }
        const uint32_t pat_words = ((uint32_t)program.size() + 1) / 2;
        const uint32_t head      = 4 + 2 * C;
        const size_t   base      = blob.size();
        blob.resize(base + head + pat_words + 2 * (size_t)NK, 0);
        uint32_t* b  = blob.data() + base;
        b[0] = t->unit_module[u];
        b[1] = C | (NK << 8) | (t->unit_delay[u] << 16);
        b[2] = gate;
        b[3] = (head + pat_words) << 16;
        if (!program.empty()) std::memcpy(b + head, program.data(), program.size() * sizeof(uint16_t));
        uint32_t next_check = 0;
        for (uint32_t c = 0; c < C; c++) {
            const uint32_t pb = t->comp_pat_begin[c0 + c], k = t->comp_pat_begin[c0 + c + 1] - pb;
            b[4 + 2 * c] = k | ((head * 2 + prog_range[c].first) << 8) | (can_self_overlap(t->pattern + pb, k) ? bss::kCompOverlap : 0u);
            const uint32_t begin = next_check;
            while (next_check < NK && level[order[next_check]] == (int)c) next_check++;
            b[5 + 2 * c] = begin | (next_check << 8) | (prog_range[c].second << 16);
        }
        uint32_t* chk = b + head + pat_words;
        for (uint32_t i = 0; i < NK; i++) {
            const int16_t* c = t->checks + 4 * (size_t)(k0 + order[i]);
            const int      lv = level[order[i]];
            // word 0: the end known before level lv is entered; word 1: the end on component lv
            const bool a_on = c[0] == lv, b_on = c[2] == lv;
            const int16_t *first = c, *second = c + 2;
            uint32_t       flags = 0;
            if (a_on != b_on) {
                if (a_on) std::swap(first, second);
            } else {
                flags = bss::kCheckFixed;
            }
            if (!flags && first[0] >= 0 && first[0] < second[0] && t->unit_delay[u] >= 1) {
                const uint32_t ka = t->comp_pat_begin[c0 + first[0] + 1] - t->comp_pat_begin[c0 + first[0]];
                const uint32_t kb = t->comp_pat_begin[c0 + second[0] + 1] - t->comp_pat_begin[c0 + second[0]];
                if (first[1] >= 0 && (uint32_t)first[1] < ka && second[1] >= 0 && (uint32_t)second[1] < kb) flags |= bss::kCheckOrdered;
            }
            chk[2 * i]     = ((uint32_t)(uint8_t)(int8_t)first[0]) | (((uint32_t)(uint16_t)first[1]) << 8) | flags;
            chk[2 * i + 1] = ((uint32_t)(uint8_t)(int8_t)second[0]) | (((uint32_t)(uint16_t)second[1]) << 8);
        }
        if (blob.size() > 0x7FFFFFFFu) return fail(BSS_ERR_LIMIT, "library too large");
    }
    unit_off[n_all] = (uint32_t)blob.size();

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
    lib->module_components = module_components;
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
    BSS_CUDA_LIB(lib->d_blob.reserve(std::max<size_t>(4, blob.size() * 4)));
    BSS_CUDA_LIB(lib->d_unit_off.reserve(unit_off.size() * 4));
    if (!blob.empty()) BSS_CUDA_LIB(cudaMemcpy(lib->d_blob.ptr, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
    BSS_CUDA_LIB(cudaMemcpy(lib->d_unit_off.ptr, unit_off.data(), unit_off.size() * 4, cudaMemcpyHostToDevice));
    {
        std::vector<uint8_t> rlen(std::max<uint32_t>(1, t->n_units), 0);
        for (uint32_t u = 0; u < t->n_units; u++) rlen[u] = (uint8_t)(t->unit_comp_begin[u + 1] - t->unit_comp_begin[u] + 1);
        BSS_CUDA_LIB(lib->d_unit_rlen.reserve(rlen.size()));
        BSS_CUDA_LIB(cudaMemcpy(lib->d_unit_rlen.ptr, rlen.data(), rlen.size(), cudaMemcpyHostToDevice));
        BSS_CUDA_LIB(lib->d_unit_probe.reserve(unit_probe.size() * 2));
        BSS_CUDA_LIB(cudaMemcpy(lib->d_unit_probe.ptr, unit_probe.data(), unit_probe.size() * 2, cudaMemcpyHostToDevice));
    }
    BSS_CUDA_LIB(lib->h_totals.reserve(64));
    BSS_CUDA_LIB(lib->d_totals.reserve(64));
#undef BSS_CUDA_LIB
    lib->dev.blob      = static_cast<const uint32_t*>(lib->d_blob.ptr);
    lib->dev.unit_off  = static_cast<const uint32_t*>(lib->d_unit_off.ptr);
    lib->dev.unit_rlen = static_cast<const uint8_t*>(lib->d_unit_rlen.ptr);
    lib->dev.unit_probe = nullptr;
    for (uint16_t p : unit_probe)
        if (p != 0xFFFFu) { lib->dev.unit_probe = static_cast<const uint16_t*>(lib->d_unit_probe.ptr); break; }
    lib->dev.n_units   = t->n_units;
    lib->dev.n_symbols = (uint32_t)alphabet.size();
    lib->dev.cmax      = cmax;
    lib->dev.pair_rows = alphabet.size() <= bss::kPairAlphabet ? 1u : 0u;
    lib->dev.list_density = 0.f;
    if (!density.empty()) {
        const size_t at = (density.size() - 1) * 98 / 100;
        std::nth_element(density.begin(), density.begin() + at, density.end());
        lib->dev.list_density = density[at];
    }
    std::memset(lib->dev.alphabet, 0, sizeof lib->dev.alphabet);
    std::copy(alphabet.begin(), alphabet.end(), lib->dev.alphabet);
    *out = lib;
    return BSS_OK;
}

void bss_library_destroy(bss_library* lib)
{
    if (!lib) return;
    cudaSetDevice(lib->device);
    if (lib->own_stream) cudaStreamSynchronize(lib->own_stream);
    if (lib->graph_exec) cudaGraphExecDestroy(lib->graph_exec);
    for (auto& ev : lib->ev)
        if (ev) cudaEventDestroy(ev);
    lib->oneshot.storage.release(); lib->d_blob.release(); lib->d_unit_off.release(); lib->d_unit_rlen.release(); lib->d_unit_probe.release(); lib->d_prefilter.release(); lib->d_alive.release(); lib->d_lane_counts.release(); lib->d_heavy.release(); lib->d_counts.release(); lib->d_chunk_sites.release();
    lib->d_idx_work.release(); lib->d_idx_hist.release(); lib->spare_idx_main.release(); lib->spare_idx_overlap.release();
    lib->spare_idx_host_main.release(); lib->spare_idx_host_overlap.release(); lib->h_idx_totals.release();
    lib->d_chunk_words.release(); lib->d_totals.release(); lib->d_seq_word_begin.release(); lib->h_totals.release();
    lib->spare_records.release(); lib->spare_host.release();
    if (lib->own_stream) cudaStreamDestroy(lib->own_stream);
    delete lib;
}

int bss_library_set_stream(bss_library* lib, void* cuda_stream)
{
    if (!lib) return fail(BSS_ERR_INVALID, "null library");
    lib->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : lib->own_stream;
    return BSS_OK;
}

int bss_library_set_module_base(bss_library* lib, uint32_t base)
{
    if (!lib) return fail(BSS_ERR_INVALID, "null library");
    lib->dev.module_base = base;
    lib->dev_serial++;
    return BSS_OK;
}

uint32_t bss_library_units(const bss_library* lib) { return lib ? lib->n_units : 0; }
uint32_t bss_library_modules(const bss_library* lib) { return lib ? lib->n_modules : 0; }
uint32_t bss_library_module_components(const bss_library* lib, uint32_t module)
{
    return (lib && module < lib->module_components.size()) ? lib->module_components[module] : 0;
}

}   // extern "C"

namespace {

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
    // one allocation: [seq_begin | mask_begin | band | occurrence table | masks | seqs]
    const uint32_t ns = lib->dev.n_symbols, occ_rows = ns + (lib->dev.pair_rows ? ns * ns : 0u);
    const uint32_t occ_stride = (max_len + 32) / 32 + bss::kOccPad;
    const bool     with_kmers = lib->dev.unit_probe != nullptr;
    const size_t   occ_words  = (size_t)n_seqs * occ_rows * occ_stride + (with_kmers ? (size_t)n_seqs * bss::kKmerWords : 0);
    const size_t off_sb = 0, off_mb = off_sb + 8 * (size_t)(n_seqs + 1), off_band = off_mb + 8 * (size_t)(n_seqs + 1),
                 off_occ = off_band + 4 * (((size_t)n_seqs + 3) & ~(size_t)3), off_mask = off_occ + 4 * ((occ_words + 3) & ~(size_t)3),
                 off_seq = off_mask + 4 * total_mask, bytes = off_seq + total_len + 16;
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
    q->dev.occ_rows   = occ_rows;
    q->dev.occ_stride = occ_stride;
    q->dev.n_seqs     = n_seqs;
    q->dev.max_len    = max_len;
    q->n_symbols      = ns;
    q->pair_rows      = lib->dev.pair_rows;
    std::memcpy(q->alphabet, lib->dev.alphabet, sizeof q->alphabet);
    // per-symbol (and per-symbol-pair) occurrence bit rows of every sequence, matched against by every unit
    if (e == cudaSuccess) e = bss::launch_occ(lib->dev, q->dev, reinterpret_cast<uint32_t*>(d + off_occ), const_cast<uint32_t*>(q->dev.kmers), st);
    // widest pair the mask allows, per sequence: bounds the window a tied component is searched in
    if (e == cudaSuccess && !device_mask) e = bss::launch_band(q->dev, reinterpret_cast<int32_t*>(d + off_band), st);   // (a device-built mask gets its band afterwards)
    // (copies from pageable host memory are staged before cudaMemcpyAsync returns: sb / mb may go out of scope)
    if (e == cudaSuccess && wait) e = cudaStreamSynchronize(st);
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
    int rc = check_pij_args(lib, d_pij, n, row_stride, col_stride);
    if (rc != BSS_OK) return rc;
    if (n && !d_mask) return fail(BSS_ERR_INVALID, "null argument");
    BSS_CUDA(cudaSetDevice(lib->device));
    BSS_CUDA(bss::launch_pair_mask(d_pij, n, row_stride, col_stride, theta, d_mask, lib->sm_count,
                                   cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : lib->stream));
    return BSS_OK;
}

int bss_pair_mask(bss_library* lib, const float* pij, uint32_t n, uint64_t row_stride, uint64_t col_stride, float theta, uint32_t* mask)
{
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
}

int bss_query_create_pij(bss_library* lib, const char* seq, uint32_t n, const float* pij, uint64_t row_stride, uint64_t col_stride, float theta,
                         bss_query** out)
{
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
        if (e == cudaSuccess) e = bss::launch_band(q->dev, const_cast<int32_t*>(q->dev.band), lib->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(lib->stream);
        d_pij.release();
        if (e != cudaSuccess) rc = cuda_fail(e, "bss_query_create_pij");
        q->h2d_bytes = 16 * 2 + extent * 4 + n;
    }
    if (rc != BSS_OK) {
        q->storage.release();
        delete q;
        return rc;
    }
    *out = q;
    return BSS_OK;
}

int bss_query_create(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin, const uint32_t* masks,
                     const uint64_t* mask_begin, bss_query** out)
{
    if (!out) return fail(BSS_ERR_INVALID, "null argument");
    *out         = nullptr;
    bss_query* q = new bss_query();
    const int  rc = query_fill(lib, q, n_seqs, seqs, seq_begin, masks, mask_begin, true);
    if (rc != BSS_OK) {
        q->storage.release();
        delete q;
        return rc;
    }
    *out = q;
    return BSS_OK;
}

void bss_query_destroy(bss_query* q)
{
    if (!q) return;
    cudaSetDevice(q->device);
    q->storage.release();
    delete q;
}

}   // extern "C"

namespace {

// count + scan, then the totals come back to the host. On success `plan` and `work` describe
// the pending emit.
// Validation, work decomposition and (grow-only) work buffers of a search: no stream operation.
int prepare_search(bss_library* lib, const bss_query* q, bss::Plan& plan, bss::Work& work)
{
    if (q->n_symbols != lib->dev.n_symbols || q->pair_rows != lib->dev.pair_rows || std::memcmp(q->alphabet, lib->dev.alphabet, sizeof q->alphabet))
        return fail(BSS_ERR_INVALID, "the query was created for a library with a different pattern alphabet");
    BSS_CUDA(cudaSetDevice(lib->device));
    plan = bss::make_plan(lib->dev, q->dev.n_seqs, q->dev.max_len, lib->sm_count);
    if (plan.smem_bytes_emit > 200 * 1024) return fail(BSS_ERR_LIMIT, "query too long for the shared-memory working set");
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
    BSS_CUDA(lib->d_chunk_sites.reserve(8 * (size_t)(plan.n_chunks + 1)));
    BSS_CUDA(lib->d_chunk_words.reserve(8 * (size_t)(plan.n_chunks + 1)));
    BSS_CUDA(lib->d_seq_word_begin.reserve(8 * (size_t)(q->dev.n_seqs + 1)));
    work.prefilter = nullptr;
    if (lib->dev.unit_probe && q->dev.kmers) {
        BSS_CUDA(lib->d_prefilter.reserve(4 * ((n_items + 31) / 32) + 4));
        work.prefilter = static_cast<uint32_t*>(lib->d_prefilter.ptr);
    }
    work.counts         = static_cast<uint32_t*>(lib->d_counts.ptr);
    work.chunk_sites    = static_cast<uint64_t*>(lib->d_chunk_sites.ptr);
    work.chunk_words    = static_cast<uint64_t*>(lib->d_chunk_words.ptr);
    work.totals         = static_cast<uint64_t*>(lib->d_totals.ptr);
    work.seq_word_begin = static_cast<uint64_t*>(lib->d_seq_word_begin.ptr);
    lib->stats          = bss_stats{};
    lib->last_query_serial = q->serial;
    return BSS_OK;
}

// Timing events: inside a stream capture they must be recorded as external events to stay usable.
cudaError_t record_event(bss_library* lib, int i, bool capturing)
{
    return capturing ? cudaEventRecordWithFlags(lib->ev[i], lib->stream, cudaEventRecordExternal) : cudaEventRecord(lib->ev[i], lib->stream);
}

// Queues the count pass (+ heavy count) and the scan.
int enqueue_count(bss_library* lib, const bss_query* q, const bss::Plan& plan, const bss::Work& work, bool capturing = false)
{
    cudaStream_t st = lib->stream;
    BSS_CUDA(cudaMemsetAsync(work.totals, 0, 48, st));
    BSS_CUDA(cudaMemsetAsync(work.seq_word_begin, 0, 8 * (size_t)(q->dev.n_seqs + 1), st));
    BSS_CUDA(record_event(lib, 0, capturing));
    BSS_CUDA(bss::launch_count(lib->dev, q->dev, plan, work, lib->sm_count, st));
    BSS_CUDA(record_event(lib, 1, capturing));
    BSS_CUDA(bss::launch_scan(q->dev, plan, work, st));
    BSS_CUDA(record_event(lib, 2, capturing));
    return BSS_OK;
}

// What the count pass found, once the copy queued by enqueue_totals has landed.
int parse_totals(bss_library* lib, const bss_query* q, const bss::Plan& plan, uint64_t& n_words, uint64_t& n_sites)
{
    const uint64_t* tot = static_cast<const uint64_t*>(lib->h_totals.ptr);
    n_words             = tot[0];
    n_sites             = tot[1];
    lib->n_alive        = (uint32_t)tot[3];
    lib->n_heavy        = (uint32_t)std::min<uint64_t>(tot[4], 1024);
    // prefilter (when the library has k-mer probes), count, heavy count (32-lane groups), scan
    lib->stats.kernel_launches   = (plan.n_chunks ? 1u : 0u) + 1u + (plan.group == 32 && plan.list_cap && plan.n_chunks ? 1u : 0u) +
                                 (lib->dev.unit_probe && q->dev.kmers && plan.n_chunks ? 1u : 0u);
    lib->stats.placements_tested = q->total_len * lib->n_units;
    lib->stats.n_sites           = n_sites;
    lib->stats.n_words           = n_words;
    lib->stats.d2h_bytes         = 40;
    if (tot[2] & 1) return fail(BSS_ERR_COUNT, "a unit has 2^32 or more sites on one sequence");
    return BSS_OK;
}

int run_emit(bss_library* lib, const bss_query* q, const bss::Plan& plan, const bss::Work& work, uint32_t* d_records, uint64_t n_words)
{
    cudaStream_t st = lib->stream;
    if (n_words && plan.n_chunks) {
        BSS_CUDA(bss::launch_emit(lib->dev, q->dev, plan, work, d_records, lib->n_alive, lib->n_heavy, ~0ull, lib->sm_count, st));
        if (lib->n_heavy) lib->stats.kernel_launches++;
        lib->stats.kernel_launches++;
    }
    BSS_CUDA(cudaEventRecord(lib->ev[3], st));
    return BSS_OK;
}

int finish_stats(bss_library* lib)
{
    BSS_CUDA(cudaEventSynchronize(lib->ev[3]));
    cudaEventElapsedTime(&lib->stats.count_ms, lib->ev[0], lib->ev[1]);
    cudaEventElapsedTime(&lib->stats.scan_ms, lib->ev[1], lib->ev[2]);
    cudaEventElapsedTime(&lib->stats.emit_ms, lib->ev[2], lib->ev[3]);
    cudaEventElapsedTime(&lib->stats.total_ms, lib->ev[0], lib->ev[3]);
    return BSS_OK;
}

}   // namespace

extern "C" {

int bss_search_query(bss_library* lib, const bss_query* q, int fetch, bss_sites** out)
{
    if (!lib || !q || !out) return fail(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (q->device != lib->device) return fail(BSS_ERR_INVALID, "query and library live on different devices");
    bss::Plan plan;
    bss::Work work;
    uint64_t  n_words = 0, n_sites = 0;
    int       rc = prepare_search(lib, q, plan, work);
    if (rc != BSS_OK) return rc;

    bss_sites* s = new bss_sites();
    s->lib       = lib;
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
    rc = enqueue_count(lib, q, plan, work);
    if (rc != BSS_OK) return bail(rc);
    if (guess_words >= 16 && plan.n_chunks) {
        e = bss::launch_emit(lib->dev, q->dev, plan, work, static_cast<uint32_t*>(s->d_records.ptr), 0xFFFFFFFFu, 0, guess_words, lib->sm_count, st);
        if (e == cudaSuccess) e = cudaEventRecord(lib->ev[3], st);
        if (e != cudaSuccess) return bail(cuda_fail(e, "emit launch"));
        emitted = true;
    }
    e = cudaMemcpyAsync(lib->h_totals.ptr, work.totals, 40, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(cuda_fail(e, "count pass"));
    rc = parse_totals(lib, q, plan, n_words, n_sites);
    if (rc != BSS_OK) return bail(rc);
    s->n_words = n_words;
    s->n_sites = n_sites;
    if (emitted && n_words <= guess_words) {
        lib->stats.kernel_launches += 1 + (plan.group == 32 && plan.list_cap ? 1 : 0);   // emit, heavy emit (queued blind)
    } else {   // first search of this size: allocate, then emit with the sizes now known on the host
        e = s->d_records.reserve(std::max<uint64_t>(16, n_words * 4));
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc(records)"));
        rc = run_emit(lib, q, plan, work, static_cast<uint32_t*>(s->d_records.ptr), n_words);
        if (rc != BSS_OK) return bail(rc);
    }
    s->seq_word_begin.assign(q->dev.n_seqs + 1, 0);
    if (fetch) {
        e = s->h_records.reserve(std::max<uint64_t>(16, n_words * 4));
        if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMallocHost(records)"));
        if (n_words) e = cudaMemcpyAsync(s->h_records.ptr, s->d_records.ptr, n_words * 4, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return bail(cuda_fail(e, "records download"));
        s->fetched = true;
        lib->stats.d2h_bytes += n_words * 4;
    }
    e = cudaMemcpyAsync(s->seq_word_begin.data(), work.seq_word_begin, 8 * (size_t)(q->dev.n_seqs + 1), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(cuda_fail(e, "result download"));
    lib->stats.d2h_bytes += 8 * (uint64_t)(q->dev.n_seqs + 1);
    rc = finish_stats(lib);
    if (rc != BSS_OK) return bail(rc);
    *out = s;
    return BSS_OK;
}

int bss_search_query_into(bss_library* lib, const bss_query* q, uint32_t* d_records, uint64_t capacity_words,
                          uint64_t* d_seq_word_begin, uint64_t* n_words_out, uint64_t* n_sites_out)
{
    if (!n_words_out || !n_sites_out) return fail(BSS_ERR_INVALID, "null argument");
    const int rc = bss_search_query_into_async(lib, q, d_records, capacity_words, d_seq_word_begin);
    if (rc != BSS_OK) return rc;
    return bss_search_wait(lib, n_words_out, n_sites_out);
}

int bss_search_query_into_async(bss_library* lib, const bss_query* q, uint32_t* d_records, uint64_t capacity_words, uint64_t* d_seq_word_begin)
{
    if (!lib || !q) return fail(BSS_ERR_INVALID, "null argument");
    if (q->device != lib->device) return fail(BSS_ERR_INVALID, "query and library live on different devices");
    bss::Plan plan;
    bss::Work work;
    // Everything is queued at once — count, scan, emit — and the host reads the totals only at the
    // end: the emit passes take their sizes from device memory and write nothing when the records
    // would not fit `capacity_words`. While the arguments stay the same (a query stream searched
    // into the same buffer, as in a benchmark loop or a sharded exchange) the whole sequence is
    // replayed as one CUDA graph: one launch call instead of a dozen.
    int rc = prepare_search(lib, q, plan, work);
    if (rc != BSS_OK) return rc;
    cudaStream_t st      = lib->stream;
    const bool   emitting = plan.n_chunks && d_records && capacity_words;
    auto enqueue = [&](bool capturing) -> int {
        int r = enqueue_count(lib, q, plan, work, capturing);
        if (r != BSS_OK) return r;
        if (emitting) {
            BSS_CUDA(bss::launch_emit(lib->dev, q->dev, plan, work, d_records, 0xFFFFFFFFu, 0, capacity_words, lib->sm_count, st));
            if (d_seq_word_begin) BSS_CUDA(bss::launch_guarded_copy(work, capacity_words, d_seq_word_begin, q->dev.n_seqs + 1, st));
        }
        BSS_CUDA(record_event(lib, 3, capturing));
        BSS_CUDA(cudaMemcpyAsync(lib->h_totals.ptr, work.totals, 40, cudaMemcpyDeviceToHost, st));
        return BSS_OK;
    };
    bool launched = false;
    if (lib->graphs_ok) {
        bss_library::GraphKey key;
        key.query_serial = q->serial; key.capacity = capacity_words; key.dev_serial = lib->dev_serial;
        key.records = d_records; key.seq_out = d_seq_word_begin; key.stream = st; key.counts = work.counts; key.alive = work.alive;
        key.lanes = work.lane_counts; key.heavy = work.heavy_items; key.chunk_sites = work.chunk_sites; key.chunk_words = work.chunk_words;
        key.seq_begin = work.seq_word_begin; key.prefilter = work.prefilter;
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
    return BSS_OK;
}

int bss_search_wait(bss_library* lib, uint64_t* n_words_out, uint64_t* n_sites_out)
{
    if (!lib) return fail(BSS_ERR_INVALID, "null argument");
    if (!lib->pending.active) return fail(BSS_ERR_INVALID, "no search is pending on this library");
    lib->pending.active = false;
    const bss_library::Pending& p  = lib->pending;
    cudaStream_t                st = lib->stream;
    uint64_t                    n_words = 0, n_sites = 0;
    BSS_CUDA(cudaSetDevice(lib->device));
    BSS_CUDA(cudaStreamSynchronize(st));
    int rc = parse_totals(lib, p.q, p.plan, n_words, n_sites);
    if (n_words_out) *n_words_out = n_words;
    if (n_sites_out) *n_sites_out = n_sites;
    if (rc != BSS_OK) return rc;
    if (n_words > p.capacity || (n_words && !p.has_records)) return fail(BSS_ERR_OVERFLOW, "output buffer too small");
    if (p.seq_out && !p.emitting) {   // nothing was queued for the output: the offsets are all zero
        BSS_CUDA(cudaMemsetAsync(p.seq_out, 0, 8 * (size_t)(p.q->dev.n_seqs + 1), st));
        BSS_CUDA(cudaStreamSynchronize(st));
    }
    if (p.emitting) lib->stats.kernel_launches += 1 + (p.plan.group == 32 && p.plan.list_cap ? 1 : 0) + (p.seq_out ? 1 : 0);   // emit, heavy emit, offsets copy
    return finish_stats(lib);
}

int bss_gather_cut(bss_library* lib, const uint32_t* d_blocks, uint32_t world, uint64_t block_words, uint32_t header_words, uint32_t* d_out,
                   uint64_t* d_counts, uint32_t* d_overflow, void* cuda_stream)
{
    if (!lib || !d_blocks || !d_out || !d_counts || !d_overflow) return fail(BSS_ERR_INVALID, "null argument");
    if (header_words < 2 || (header_words & 1) || block_words < header_words) return fail(BSS_ERR_INVALID, "bad block layout");
    BSS_CUDA(cudaSetDevice(lib->device));
    BSS_CUDA(bss::launch_gather_cut(d_blocks, world, block_words, header_words, d_out, d_counts, d_overflow,
                                    cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : lib->stream));
    return BSS_OK;
}

int bss_search_batch(bss_library* lib, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin, const uint32_t* masks,
                     const uint64_t* mask_begin, bss_sites** out)
{
    if (!out) return fail(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!lib) return fail(BSS_ERR_INVALID, "null argument");
    bss_query* q  = &lib->oneshot;   // device storage kept between calls: no allocation once it has grown
    int        rc = query_fill(lib, q, n_seqs, seqs, seq_begin, masks, mask_begin, false);
    if (rc != BSS_OK) return rc;
    rc = bss_search_query(lib, q, 1, out);
    if (rc == BSS_OK) lib->stats.h2d_bytes = q->h2d_bytes;
    return rc;
}

int bss_search(bss_library* lib, const char* seq, uint32_t n, const uint32_t* mask, bss_sites** out)
{
    const uint64_t seq_begin[2]  = {0, n};
    const uint64_t mask_begin[2] = {0, (uint64_t)n * ((n + 31) / 32)};
    return bss_search_batch(lib, 1, seq, seq_begin, mask, mask_begin, out);
}

uint64_t        bss_sites_count(const bss_sites* s) { return s ? s->n_sites : 0; }
uint64_t        bss_sites_words(const bss_sites* s) { return s ? s->n_words : 0; }
const uint32_t* bss_sites_records(const bss_sites* s) { return (s && s->fetched) ? static_cast<const uint32_t*>(s->h_records.ptr) : nullptr; }
const uint32_t* bss_sites_device_records(const bss_sites* s) { return s ? static_cast<const uint32_t*>(s->d_records.ptr) : nullptr; }
const uint64_t* bss_sites_seq_word_begin(const bss_sites* s) { return s ? s->seq_word_begin.data() : nullptr; }

void bss_sites_free(bss_sites* s)
{
    if (!s) return;
    if (s->lib) {
        cudaSetDevice(s->lib->device);
        // keep the largest buffers around for the next search
        if (s->d_records.cap > s->lib->spare_records.cap) std::swap(s->d_records, s->lib->spare_records);
        if (s->h_records.cap > s->lib->spare_host.cap) std::swap(s->h_records, s->lib->spare_host);
    }
    s->d_records.release();
    s->h_records.release();
    delete s;
}

}   // extern "C"

struct bss_site_index {
    bss_library* lib = nullptr;
    uint32_t     n = 0, launches = 0;
    float        ms = 0.0f;
    bool         fetched = false;
    DeviceBuffer d_main, d_overlap;
    PinnedBuffer h_main, h_overlap;
    uint64_t     offset[BSS_IDX_SECTIONS] = {}, count[BSS_IDX_SECTIONS] = {};   // in u32 elements; OVERLAP_VARS lives in its own buffer
};

extern "C" {

int bss_site_index_build(bss_library* lib, const bss_query* q, uint32_t seq, const uint32_t* d_records, int c3_rule, int fetch,
                         bss_site_index** out)
{
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
    if (x->lib) {
        cudaSetDevice(x->lib->device);
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

int bss_measure_int_issue(bss_library* lib, double* thread_ops_per_s)
{
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
}

int bss_last_stats(const bss_library* lib, bss_stats* out)
{
    if (!lib || !out) return fail(BSS_ERR_INVALID, "null argument");
    *out = lib->stats;
    return BSS_OK;
}

}   // extern "C"
