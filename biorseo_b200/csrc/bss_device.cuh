// Device-side data layout and kernel launch interface of the insertion-site search.
// Internal to the library; the public boundary is include/biorseo_b200.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include <new>
#include <stdexcept>
#include <string>

#include <vector>

#include "bss_layout.h"

struct bss_sites;

namespace bss {

// A result handle that no library owns (bss_search_sharded): pinned host records of n_words words, to be filled by
// the caller through *records. Null when the pinned memory cannot be had.
bss_sites* new_host_sites(uint64_t n_sites, uint64_t n_words, const std::vector<uint64_t>& seq_word_begin, uint32_t** records);

// Sets the thread-local message behind bss_last_error() and returns `code` (bss_capi.cu).
int set_error(int code, const std::string& what);

// Body of an extern "C" entry point: no C++ exception crosses the boundary (allocation failures of the host-side
// containers, anything a callee throws) — it becomes a status code + bss_last_error().
template <typename F>
int guarded(F&& body)
{
    try {
        return body();
    } catch (const std::bad_alloc&) {
        return set_error(-8 /* BSS_ERR_NOMEM */, "out of host memory");
    } catch (const std::exception& e) {
        return set_error(-9 /* BSS_ERR_INTERNAL */, std::string("unexpected exception: ") + e.what());
    } catch (...) {
        return set_error(-9 /* BSS_ERR_INTERNAL */, "unexpected exception");
    }
}

constexpr int      kThreads    = 128;          // threads per CTA (4 warps)
constexpr int      kOccPad     = 10;           // zero words after each occurrence row (patterns up to 255+32 bytes)
constexpr int      kBlobStage  = (int)kBlobStageWords;   // unit descriptor words staged in shared memory per group
constexpr uint32_t kMaxChunkUnits = 2048;      // units per chunk (bounds the per-CTA offset table)
__host__ __device__ constexpr uint32_t kmer_table_word(uint32_t table) { return table == 0 ? 0u : table == 1 ? 8u : 40u; }
// words reserved per position of a sequence of at most max_len nucleotides in the banded pair masks: any band fits
__host__ __device__ constexpr uint32_t band_stride(uint32_t max_len) { return (max_len + 30) / 32 + 1; }

// (the unit descriptor layout and its flags are in bss_layout.h)
struct DevLibrary {
    const uint32_t* blob;
    const uint32_t* unit_off;    // [n_units + n_gates + 1] word offsets into blob
    const uint8_t*  unit_rlen;   // [n_units] words per record of the unit: components + 1
    const uint16_t* unit_probe;  // [n_units][4] k-mer probes (longest window of 4..6 consecutive literal symbols) of up to four components, 0xFFFF = none; null: no filter
    const uint8_t*  unit_flags;  // [n_units] kUnitWindow, kUnitLean
    const uint32_t* pat_blob;    // pattern table: one-component descriptors of the slots (bss_layout.h)
    const uint32_t* pat_off;     // [n_slots + 1] word offsets into pat_blob
    const uint16_t* pat_tslot;   // [n_slots] row of the slot's thinned set, or kNoSlot
    const uint16_t* unit_slots;  // [n_units][4] slots of the first four components, kNoSlot = none
    uint32_t        n_slots, n_tslots;
    uint32_t        n_units;     // searched units
    uint32_t        n_symbols;   // alphabet size
    uint32_t        cmax;        // max components per unit (gates included)
    uint32_t        module_base; // added to every module index written into a record (library shards)
    float           list_density; // expected matches per nucleotide of a unit (all components), 98th percentile over units
    uint32_t        pair_rows;   // 1: the match programs use pair rows (alphabet of at most kPairAlphabet symbols)
    uint8_t         alphabet[32];
};

struct DevQuery {
    const uint8_t*  seqs;
    const uint64_t* seq_begin;    // [n_seqs + 1]
    const uint32_t* masks;
    const uint64_t* mask_begin;   // [n_seqs + 1], in words
    const int32_t*  band;         // [n_seqs] max v-u over the set bits above the diagonal, -1 when there is none
    const uint32_t* occ;          // [n_seqs][occ_rows][occ_stride] occurrence table (see occ_kernel)
    const uint32_t* kmers;        // [n_seqs][kKmerWords] k-mer presence bits, or null (alphabets of more than 4 symbols)
    const uint32_t* bmask;        // banded pair masks, or null: sequence q's starts at word bm_stride * seq_begin[q]; row u = K words
                                  // (K = band_words(band[q])) holding mask words u >> 5 .. of row u, bits at or below the diagonal and
                                  // beyond the sequence cleared.
    uint32_t        bm_stride;    // band_stride(max_len)
    uint32_t        occ_rows;     // n_symbols, plus n_symbols^2 pair rows for small alphabets
    uint32_t        occ_stride;   // words per row: ceil((max_len + 1) / 32) + kOccPad
    uint32_t        n_seqs;
    uint32_t        max_len;
};

// Work decomposition: chunk = (sequence, contiguous unit range); one CTA per chunk, the
// groups of a CTA draw the units of their chunk from a shared counter.
struct Plan {
    uint32_t units_per_chunk;
    uint32_t chunks_per_seq;
    uint32_t n_chunks;
    uint32_t group;        // lanes cooperating on one (sequence, unit) item: 4, 8, 16 or 32
    uint32_t W;            // match-mask words per component, ceil((max_len + 1) / 32)
    uint32_t WS;           // summary words per component, ceil(W / 32)
    uint32_t heavy_ranks;  // items with this many placements before the checks are walked block by block
    uint32_t heavy_block;  // ranks per block of a heavy item (at most 256 blocks per item)
    uint32_t list_cap;     // rank-space walk: match-list entries per item (0 = depth-first walk only)
    uint32_t window;       // 1: items of window-walk units on narrow-band sequences go through route_kernel + window_kernel
    uint32_t window_draw;  // list entries a warp of the window count pass draws per atomic
    int32_t  window_band;  // widest band the window kernel's working set is sized for: sequences with a wider one stay with the ordinary passes
    uint32_t window_qcap;  // window walk: entries per level queue (at least the widest band + 1)
    uint32_t window_kw;    // window walk: stash words per lane = words a band window can span
    uint32_t dict;         // 1: the pattern table of this (query, library) is built at the start of the search and the window pass reads it
    uint32_t window_stage; // 1: the window kernel stages the occurrence table and the banded mask in shared memory (one sequence)
    uint32_t smem_bytes_window;   // dynamic shared memory of the window kernel
    uint32_t smem_bytes;        // dynamic shared memory of the count kernel
    uint32_t smem_bytes_emit;   // ... of the emit kernel (adds the record staging)
};

struct Work {
    uint32_t* counts;        // [n_seqs * n_units] sites per item
    uint32_t* prefilter;     // [ceil(n_items / 32)] bit per item: not ruled out by the k-mer presence filter; null: no filter
    uint64_t* chunk_sites;   // [n_chunks]
    uint64_t* chunk_words;   // [n_chunks] ; after the scan: exclusive prefix, [n_chunks] = total
    uint64_t* totals;        // [0] words, [1] sites, [2] error flags, [3] items with sites (length of `alive`), [4] heavy items,
                             // [5] window items with sites (length of `alive_w`), [6] window items (length of `wlist`),
                             // [7] items left to the ordinary passes, [8] draw cursor of the window count pass (kTotalsWords in all)
    uint32_t* alive;         // [n_seqs * n_units] items with sites, in completion order
    uint32_t* lane_counts;   // [lane_cap][32] sites of every lane's run of ranks (rank-space walk)
    uint32_t  lane_cap;      // alive-list slots that have a lane_counts row
    uint32_t* heavy_items;   // [1024] items parked by the count pass (rank-space walk, too many placements for one warp)
    uint32_t* heavy_ranks;   // [1024] their placements before the checks
    uint32_t* task_sites;    // [1024 * 256] sites of every (heavy item, block) task
    uint32_t* task_lane_counts;   // [1024 * 256][32] sites of every lane's run inside a task
    uint64_t* seq_word_begin;   // [n_seqs + 1]
    // window walk (route_kernel, window_kernel)
    uint32_t* wlist;         // [n_seqs * n_units] items routed to the window pass, in arrival order
    uint32_t* alive_w;       // [n_seqs * n_units] window items with sites, in completion order
    // pattern table of this search (pattern_kernel): per (sequence, slot) the match set + its word summary, the thinned set
    // of the slots that can overlap themselves, and flags (bit 0: the pattern occurs, bit 1: its matches overlap on this sequence)
    uint32_t* ptab_m;        // [n_seqs][n_slots][W + WS]
    uint32_t* ptab_t;        // [n_seqs][n_tslots][W]
    uint8_t*  pflag;         // [n_seqs][n_slots]
};
constexpr uint32_t kTotalsWords = 16;

// Developer knobs (bss_library_set_tuning): force rarely taken paths on small inputs (tests), compare variants
// (benchmarks). A negative value means automatic.
struct Tuning {
    int64_t group = -1, units_per_chunk = -1, heavy_ranks = -1, list_cap = -1, window = -1, window_stage = -1, window_draw = -1, dict = -1, phase_events = -1;
};

__host__ __device__ inline int band_words(int band) { return band < 0 ? 0 : ((31 + band) >> 5) + 1; }

// known_band: the widest band over the query's sequences when the host knows it (>= -1), else INT32_MIN
Plan make_plan(const DevLibrary& lib, const DevQuery& q, int sm_count, const Tuning& tuning, uint32_t n_window_units, int known_band);

cudaError_t launch_band(const DevQuery& q, int32_t* band, uint32_t* bmask, int32_t* band_host, cudaEvent_t copied, cudaStream_t s);
// bit v of row u (u < v) = v >= u + 4 && u + 6 < n && pij[u * row_stride + v * col_stride] > theta
cudaError_t launch_pair_mask(const float* pij, uint32_t n, uint64_t row_stride, uint64_t col_stride, float theta, uint32_t* mask,
                             int sm_count, cudaStream_t s);
cudaError_t launch_gather_cut(const uint32_t* blocks, uint32_t world, uint64_t block_words, uint32_t header_words, uint32_t* out,
                              uint64_t* counts, uint32_t* overflow, cudaStream_t s);
cudaError_t launch_occ(const DevLibrary& lib, const DevQuery& q, uint32_t* occ, uint32_t* kmers, cudaStream_t s);
// `ordinary`: the ordinary count passes are needed (some item may not take the window walk)
cudaError_t launch_count(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, bool ordinary, int sm_count, cudaStream_t s);
// seq_out (may be null): device copy of the per-sequence word offsets for the caller, written by the scan itself
cudaError_t launch_scan(const DevQuery& q, const Plan& plan, const Work& work, uint64_t* seq_out, cudaStream_t s);
// totals, chunk totals and per-sequence offsets zeroed by one kernel — or by the pattern kernel when it is the first kernel of the search
cudaError_t launch_reset(const Plan& plan, const Work& work, uint32_t n_seqs, cudaStream_t s);
inline bool reset_is_folded(const DevLibrary& lib, const DevQuery& q, const Plan& plan) { return plan.dict && lib.n_slots > 0 && q.n_seqs > 0; }

cudaError_t launch_emit(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records,
                        uint32_t n_alive, uint32_t n_heavy, uint64_t capacity_words, bool ordinary, int sm_count, cudaStream_t s);
// Exact number of placements the reference's enumerator materialises before its filters (cppsrc/Motif.cpp:466-474),
// summed over all items into *total (u64, device, added to). Diagnostic: not part of a search.
// scratch: n_ctas * (kThreads / 32) * per_warp u64 words, per_warp = placements_words_per_warp(max_len).
size_t      placements_words_per_warp(uint32_t max_len);
cudaError_t launch_placements(const DevLibrary& lib, const DevQuery& q, const Plan& plan, uint64_t* scratch, size_t per_warp, uint32_t n_ctas,
                              unsigned long long* total, cudaStream_t s);

// Integer-issue probe (bss_probe.cu): SHF + LOP3 only, all SMs; *thread_ops = operations issued by the launch.
cudaError_t launch_int_issue_probe(uint32_t iters, int sm_count, uint32_t* sink, uint64_t* thread_ops, cudaStream_t s);

// Insertion-variable index (bss_index.cu). totals: [0] sites, [1] variables, [2] pair ends (2 x allowed pairs),
// [3] overlap entries, [4] error flags (bit 0: the records are not those of the last search).
uint32_t    index_overlap_blocks(uint64_t n_vars, uint32_t n, int sm_count);
cudaError_t launch_index_scan(const DevLibrary& lib, const uint32_t* counts, uint64_t* unit_site_begin, uint64_t* unit_var_begin,
                              const uint32_t* mask, uint32_t n, uint32_t* row_pairs, uint32_t* degree, uint32_t* row_first_pair,
                              uint32_t* pair_begin, uint64_t* totals, int sm_count, cudaStream_t s);
cudaError_t launch_index_vars(const DevLibrary& lib, const uint64_t* unit_site_begin, const uint64_t* unit_var_begin, const uint32_t* records,
                              uint32_t n_sites, uint32_t n_vars, const uint32_t* mask, uint32_t n, const uint32_t* pair_begin, int c3_rule,
                              uint32_t* pair_partner, uint32_t* var_begin, uint32_t* var_first, uint32_t* var_second, uint32_t* var_c3,
                              uint32_t* hist, uint32_t n_blocks, uint32_t* overlap_total, uint32_t* overlap_begin, uint64_t* totals,
                              int sm_count, cudaStream_t s);
cudaError_t launch_index_fill(const uint32_t* var_first, const uint32_t* var_second, uint32_t n_vars, uint32_t n, const uint32_t* hist,
                              uint32_t n_blocks, const uint32_t* overlap_begin, uint32_t* overlap_vars, cudaStream_t s);

}   // namespace bss
