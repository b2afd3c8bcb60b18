// Multi-GPU search in ONE process (row (e) of the path, SURVEY 8b "X_search_sharded"): the library is cut into
// contiguous module ranges balanced by a search-cost estimate, one range per GPU; every GPU searches its range on
// the (replicated) query, and the records of all ranges come back as ONE stream in canonical order in pinned host
// memory. What it replaces (reference, CPU): the Pool of one job per module of cppsrc/MOIP.cpp:246-293 (and the
// RIN / DESC twins :159-243) — the reference's unit of parallelism is the module, so is the shard axis here.
//
// Host-side orchestration only: the searches run through the public entry points of bss_capi.cu, one persistent
// host thread per GPU (a CUDA context switch per call is avoided and the GPUs work at the same time). The ranges'
// streams are placed by displacement straight into the result buffer (device -> pinned host copies, one per GPU,
// concurrent); no device-side exchange is needed for a host consumer. For device consumers the peer-memory
// exchange kernel (bss_exchange.cu, bss_exchange_open_local) serves the same ranges.
#include "biorseo_b200.h"
#include "bss_device.cuh"

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

// Search-cost estimate per module on a query of n nucleotides: the match phase (one shift-AND step per pattern
// byte per 32-position word) plus the expected number of placements before the pair checks (product over the
// components of the expected matches n / 4^literals), summed over the module's units. Only the ratios matter.
// (The same estimate as biorseo_b200/sharded.py module_costs, which the multi-process path uses.)
std::vector<double> module_costs(const bss_table* t, uint32_t n)
{
    std::vector<double> costs(t->n_modules, 0.0);
    const double        words = (double)((n + 31) / 32);
    for (uint32_t u = 0; u < t->n_units; u++) {
        double match = 0.0, placements = 1.0;
        for (uint32_t c = t->unit_comp_begin[u]; c < t->unit_comp_begin[u + 1]; c++) {
            const uint32_t pb = t->comp_pat_begin[c], k = t->comp_pat_begin[c + 1] - pb;
            uint32_t       literals = 0;
            for (uint32_t j = 0; j < k; j++) literals += t->pattern[pb + j] != '.';
            match += 2.0 * k * words;
            placements *= std::min((double)n, (double)n / std::pow(4.0, (double)literals));
        }
        costs[t->unit_module[u]] += match + 2.0 * std::min(placements, 1e9);
    }
    return costs;
}

// Contiguous module ranges whose summed costs are as even as a contiguous cut allows: range r ends at the module
// where the running cost comes closest to (r + 1) / shards of the total. begin[] has shards + 1 entries.
void ranges_by_cost(const std::vector<double>& costs, uint32_t shards, uint32_t* begin)
{
    const uint32_t      m = (uint32_t)costs.size();
    std::vector<double> prefix(m + 1, 0.0);
    for (uint32_t i = 0; i < m; i++) prefix[i + 1] = prefix[i] + costs[i];
    const double total = prefix[m];
    begin[0] = 0;
    for (uint32_t r = 1; r < shards; r++) {
        uint32_t at;
        if (total > 0) {
            const double target = total * r / shards;
            at = (uint32_t)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
            if (at > 0 && at <= m && std::fabs(prefix[at - 1] - target) <= std::fabs(prefix[std::min(at, m)] - target)) at--;
        } else {
            at = (uint32_t)(((uint64_t)m * r) / shards);
        }
        begin[r] = std::min(m, std::max(begin[r - 1], at));
    }
    begin[shards] = m;
}

// The flat table of the modules [m0, m1): the units of those modules (module indices rebased to m0) and the gates
// they name, renumbered.
struct SubTable {
    std::vector<uint32_t> unit_module, unit_delay, unit_gate, unit_comp_begin, comp_pat_begin, unit_check_begin;
    std::vector<uint8_t>  pattern;
    std::vector<int16_t>  checks;
    bss_table             view{};

    void append_unit(const bss_table* t, uint32_t u, uint32_t module)
    {
        unit_module.push_back(module);
        unit_delay.push_back(t->unit_delay[u]);
        for (uint32_t c = t->unit_comp_begin[u]; c < t->unit_comp_begin[u + 1]; c++) {
            pattern.insert(pattern.end(), t->pattern + t->comp_pat_begin[c], t->pattern + t->comp_pat_begin[c + 1]);
            comp_pat_begin.push_back((uint32_t)pattern.size());
        }
        unit_comp_begin.push_back((uint32_t)comp_pat_begin.size() - 1);
        const uint32_t k0 = t->unit_check_begin[u], k1 = t->unit_check_begin[u + 1];
        if (k1 > k0) checks.insert(checks.end(), t->checks + 4 * (size_t)k0, t->checks + 4 * (size_t)k1);
        unit_check_begin.push_back((uint32_t)(checks.size() / 4));
    }

    bool cut(const bss_table* t, uint32_t u0, uint32_t u1, uint32_t m0, uint32_t m1)
    {
        comp_pat_begin.assign(1, 0);
        unit_comp_begin.assign(1, 0);
        unit_check_begin.assign(1, 0);
        std::vector<uint32_t>                  gates;   // gate units of the whole table that this range names, in order of first use
        std::unordered_map<uint32_t, uint32_t> place;   // gate unit -> its index in `gates`
        for (uint32_t u = u0; u < u1; u++) {
            append_unit(t, u, t->unit_module[u] - m0);
            uint32_t g = t->unit_gate[u];
            if (g != bss::kNoGate) {
                if (g < t->n_units || g >= t->n_units + t->n_gates) return false;
                auto it = place.find(g);
                if (it == place.end()) {
                    it = place.emplace(g, (uint32_t)gates.size()).first;
                    gates.push_back(g);
                }
                g = (u1 - u0) + it->second;
            }
            unit_gate.push_back(g);
        }
        for (uint32_t g : gates) append_unit(t, g, 0);
        view.n_modules = m1 - m0;
        view.n_units   = u1 - u0;
        view.n_gates   = (uint32_t)gates.size();
        view.unit_module = unit_module.data(); view.unit_delay = unit_delay.data(); view.unit_gate = unit_gate.data();
        view.unit_comp_begin = unit_comp_begin.data(); view.comp_pat_begin = comp_pat_begin.data(); view.pattern = pattern.data();
        view.unit_check_begin = unit_check_begin.data(); view.checks = checks.data();
        return true;
    }
};

bool table_is_sane(const bss_table* t)
{
    if (!t) return false;
    const uint32_t n_all = t->n_units + t->n_gates;
    if (n_all && (!t->unit_module || !t->unit_delay || !t->unit_comp_begin || !t->comp_pat_begin || !t->unit_check_begin)) return false;
    if (t->n_units && !t->unit_gate) return false;
    for (uint32_t u = 0; u < t->n_units; u++) {
        if (t->unit_module[u] >= t->n_modules) return false;
        if (u && t->unit_module[u] < t->unit_module[u - 1]) return false;   // units come module by module
        if (t->unit_comp_begin[u + 1] < t->unit_comp_begin[u] || t->unit_check_begin[u + 1] < t->unit_check_begin[u]) return false;
    }
    return true;
}

struct Shard {
    int          device = 0;
    uint32_t     m0 = 0, m1 = 0;
    bss_library* lib = nullptr;
    bss_sites*   sites = nullptr;      // the range's records of the search in hand (device memory)
    uint32_t*    staging = nullptr;    // pinned host copy of a range's stream (batches of several sequences)
    size_t       staging_words = 0;
    int          rc = BSS_OK;
    std::string  error;
    std::thread  thread;
};

}   // namespace

struct bss_sharded {
    std::vector<Shard> shards;
    // one persistent host thread per GPU; a job is a function every thread runs on its shard
    std::mutex                           mu;
    std::condition_variable              wake, done;
    std::function<void(Shard&, uint32_t)> job;
    uint64_t                             generation = 0;
    uint32_t                             running = 0;
    bool                                 quit = false;

    void worker(uint32_t r)
    {
        uint64_t seen = 0;
        cudaSetDevice(shards[r].device);
        for (;;) {
            std::function<void(Shard&, uint32_t)> f;
            {
                std::unique_lock<std::mutex> lock(mu);
                wake.wait(lock, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
                f    = job;
            }
            try {
                f(shards[r], r);
            } catch (...) {
                shards[r].rc    = BSS_ERR_INTERNAL;
                shards[r].error = "unexpected exception in a shard thread";
            }
            {
                std::lock_guard<std::mutex> lock(mu);
                if (--running == 0) done.notify_all();
            }
        }
    }

    // Runs f on every shard at the same time; returns the first failure (its message in bss_last_error()).
    int all(std::function<void(Shard&, uint32_t)> f)
    {
        {
            std::lock_guard<std::mutex> lock(mu);
            for (Shard& s : shards) { s.rc = BSS_OK; s.error.clear(); }
            job     = std::move(f);
            running = (uint32_t)shards.size();
            generation++;
        }
        wake.notify_all();
        {
            std::unique_lock<std::mutex> lock(mu);
            done.wait(lock, [&] { return running == 0; });
        }
        for (Shard& s : shards)
            if (s.rc != BSS_OK) return bss::set_error(s.rc, "GPU " + std::to_string(s.device) + ": " + s.error);
        return BSS_OK;
    }
};

extern "C" {

int bss_shard_ranges(const bss_table* table, uint32_t n_shards, uint32_t typical_len, uint32_t* module_begin)
{
    return bss::guarded([&]() -> int {
    if (!table || !module_begin || n_shards < 1) return bss::set_error(BSS_ERR_INVALID, "null argument");
    if (!table_is_sane(table)) return bss::set_error(BSS_ERR_INVALID, "malformed table (units must come module by module)");
    ranges_by_cost(module_costs(table, typical_len ? typical_len : 1000u), n_shards, module_begin);
    return BSS_OK;
    });
}

int bss_sharded_create(const bss_table* table, const int* devices, uint32_t n_gpus, uint32_t typical_len, bss_sharded** out)
{
    return bss::guarded([&]() -> int {
    if (!out) return bss::set_error(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!table || n_gpus < 1 || n_gpus > 16) return bss::set_error(BSS_ERR_INVALID, "bad argument (1 to 16 GPUs)");
    const int present = bss_device_count();
    if (present <= 0) return bss::set_error(BSS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    std::vector<uint32_t> begin(n_gpus + 1);
    int rc = bss_shard_ranges(table, n_gpus, typical_len, begin.data());
    if (rc != BSS_OK) return rc;
    bss_sharded* s = new bss_sharded();
    s->shards.resize(n_gpus);
    uint32_t u = 0;
    for (uint32_t r = 0; r < n_gpus && rc == BSS_OK; r++) {
        Shard& sh = s->shards[r];
        sh.device = devices ? devices[r] : (int)r;
        sh.m0 = begin[r];
        sh.m1 = begin[r + 1];
        if (sh.device < 0 || sh.device >= present) { rc = bss::set_error(BSS_ERR_NO_DEVICE, "device " + std::to_string(sh.device) + " is not present"); break; }
        const uint32_t u0 = u;
        while (u < table->n_units && table->unit_module[u] < sh.m1) u++;
        SubTable sub;
        if (!sub.cut(table, u0, u, sh.m0, sh.m1)) { rc = bss::set_error(BSS_ERR_INVALID, "unit_gate out of range"); break; }
        rc = bss_library_create(&sub.view, sh.device, &sh.lib);
        if (rc == BSS_OK) rc = bss_library_set_module_base(sh.lib, sh.m0);
    }
    if (rc != BSS_OK) {
        const std::string why = bss_last_error();
        for (Shard& sh : s->shards)
            if (sh.lib) bss_library_destroy(sh.lib);
        delete s;
        return bss::set_error(rc, why);
    }
    for (uint32_t r = 0; r < n_gpus; r++) s->shards[r].thread = std::thread([s, r] { s->worker(r); });
    *out = s;
    return BSS_OK;
    });
}

void bss_sharded_destroy(bss_sharded* s)
{
    if (!s) return;
    {
        std::lock_guard<std::mutex> lock(s->mu);
        s->quit = true;
    }
    s->wake.notify_all();
    for (Shard& sh : s->shards)
        if (sh.thread.joinable()) sh.thread.join();
    for (Shard& sh : s->shards) {
        cudaSetDevice(sh.device);
        if (sh.sites) bss_sites_free(sh.sites);
        if (sh.lib) bss_library_destroy(sh.lib);
        if (sh.staging) cudaFreeHost(sh.staging);
    }
    delete s;
}

uint32_t bss_sharded_gpus(const bss_sharded* s) { return s ? (uint32_t)s->shards.size() : 0; }

int bss_sharded_range(const bss_sharded* s, uint32_t shard, int* device, uint32_t* module_begin, uint32_t* module_end)
{
    if (!s || shard >= s->shards.size()) return bss::set_error(BSS_ERR_INVALID, "bad argument");
    if (device) *device = s->shards[shard].device;
    if (module_begin) *module_begin = s->shards[shard].m0;
    if (module_end) *module_end = s->shards[shard].m1;
    return BSS_OK;
}

bss_library* bss_sharded_library(bss_sharded* s, uint32_t shard) { return (s && shard < s->shards.size()) ? s->shards[shard].lib : nullptr; }

int bss_search_sharded(bss_sharded* s, uint32_t n_seqs, const char* seqs, const uint64_t* seq_begin, const uint32_t* masks,
                       const uint64_t* mask_begin, bss_sites** out)
{
    return bss::guarded([&]() -> int {
    if (!out) return bss::set_error(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (!s) return bss::set_error(BSS_ERR_INVALID, "null argument");
    const uint32_t world = (uint32_t)s->shards.size();
    // 1. every GPU searches its module range; the records stay on the device, the sizes come back
    int rc = s->all([&](Shard& sh, uint32_t) {
        if (sh.sites) { bss_sites_free(sh.sites); sh.sites = nullptr; }
        sh.rc = bss_search_batch_resident(sh.lib, n_seqs, seqs, seq_begin, masks, mask_begin, &sh.sites);
        if (sh.rc != BSS_OK) sh.error = bss_last_error();
    });
    if (rc != BSS_OK) return rc;
    // 2. canonical order = (sequence, module, placement): the piece (sequence q, range r) goes behind the pieces of
    //    the same sequence of the ranges before r
    std::vector<uint64_t> begin(n_seqs + 1, 0);
    uint64_t              n_sites = 0, n_words = 0;
    for (const Shard& sh : s->shards) {
        const uint64_t* b = bss_sites_seq_word_begin(sh.sites);
        for (uint32_t q = 0; q < n_seqs; q++) begin[q + 1] += b[q + 1] - b[q];
        n_sites += bss_sites_count(sh.sites);
        n_words += bss_sites_words(sh.sites);
    }
    for (uint32_t q = 0; q < n_seqs; q++) begin[q + 1] += begin[q];
    uint32_t*  records = nullptr;
    bss_sites* result  = bss::new_host_sites(n_sites, n_words, begin, &records);
    if (!result) return bss::set_error(BSS_ERR_NOMEM, "pinned host memory for the records");
    // 3. every GPU delivers its pieces: one device -> host copy straight to the displacement (one sequence), or its
    //    whole stream to a pinned staging buffer and a host scatter of the per-sequence pieces (batches)
    rc = s->all([&](Shard& sh, uint32_t r) {
        const uint64_t words = bss_sites_words(sh.sites);
        if (!words) return;
        if (n_seqs == 1) {
            uint64_t at = 0;
            for (uint32_t p = 0; p < r; p++) at += bss_sites_words(s->shards[p].sites);
            sh.rc = bss_sites_copy_out(sh.sites, 0, words, records + at);
        } else {
            if (sh.staging_words < words) {
                if (sh.staging) cudaFreeHost(sh.staging);
                sh.staging       = nullptr;
                sh.staging_words = 0;
                if (cudaMallocHost(reinterpret_cast<void**>(&sh.staging), (words + words / 4 + 64) * 4) != cudaSuccess) {
                    sh.rc    = BSS_ERR_NOMEM;
                    sh.error = "pinned staging buffer";
                    return;
                }
                sh.staging_words = words + words / 4 + 64;
            }
            sh.rc = bss_sites_copy_out(sh.sites, 0, words, sh.staging);
            if (sh.rc == BSS_OK) {
                const uint64_t* mine = bss_sites_seq_word_begin(sh.sites);
                for (uint32_t q = 0; q < n_seqs; q++) {
                    const uint64_t len = mine[q + 1] - mine[q];
                    if (!len) continue;
                    uint64_t at = begin[q];
                    for (uint32_t p = 0; p < r; p++) {
                        const uint64_t* b = bss_sites_seq_word_begin(s->shards[p].sites);
                        at += b[q + 1] - b[q];
                    }
                    std::memcpy(records + at, sh.staging + mine[q], len * 4);
                }
            }
        }
        if (sh.rc != BSS_OK && sh.error.empty()) sh.error = bss_last_error();
    });
    if (rc != BSS_OK) {
        bss_sites_free(result);
        return rc;
    }
    (void)world;
    *out = result;
    return BSS_OK;
    });
}

}   // extern "C"
