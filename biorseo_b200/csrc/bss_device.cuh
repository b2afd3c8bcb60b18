// Device-side data layout and kernel launch interface of the insertion-site search.
// Internal to the library; the public boundary is include/biorseo_b200.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include <string>

namespace bss {

// Sets the thread-local message behind bss_last_error() and returns `code` (bss_capi.cu).
int set_error(int code, const std::string& what);

constexpr int      kThreads    = 128;          // threads per CTA (4 warps)
constexpr uint32_t kWildcard   = 0xFFu;        // pattern symbol code of '.'
constexpr int      kOccPad     = 10;           // zero words after each occurrence row (patterns up to 255+32 bytes)
constexpr int      kBlobStage  = 64;           // unit descriptor words staged in shared memory per group
constexpr uint32_t kNoGate     = 0xFFFFFFFFu;
constexpr uint32_t kPairAlphabet = 8;          // largest alphabet that gets pair rows (row index must stay below 128)
constexpr uint32_t kMaxChunkUnits = 2048;      // units per chunk (bounds the per-CTA offset table)
// k-mer presence filter (alphabets of at most 4 symbols, 2 bits per symbol): per sequence one bit per
// possible 4-, 5- and 6-mer; a probe is code | table << 12 (table 0, 1, 2 = k 4, 5, 6), 0xFFFF = none
constexpr uint32_t kKmerMin     = 4, kKmerMax = 6;
constexpr uint32_t kKmerWords   = (256 + 1024 + 4096) / 32;   // presence bits per sequence, the three tables back to back
__host__ __device__ constexpr uint32_t kmer_table_word(uint32_t table) { return table == 0 ? 0u : table == 1 ? 8u : 40u; }
constexpr uint32_t kCompOverlap = 1u << 24;    // component flag: the pattern is compatible with a shifted copy of itself
constexpr uint32_t kCheckFixed  = 1u << 24;    // check flag: no end moves at the check's level, or both do
constexpr uint32_t kCheckOrdered = 1u << 25;   // check flag: 0 <= u < v < n holds for every placement (u = word 0's end)

// Unit descriptor ("blob"), u32 words, one per unit, addressed by unit_off[unit]:
//   [0] module index
//   [1] C | n_checks << 8 | delay << 16
//   [2] gate unit or kNoGate
//   [3] word offset of the checks << 16
//   [4 + 2c]   k_c | offset of the match program << 8 (16 bits, in u16 units from the start of the blob) | kCompOverlap
//   [5 + 2c]   first check ready at level c | one past the last << 8 | steps of the match program << 16
//   match programs, one u16 per step: occurrence row (7 bits) | shift & 31 << 7 | shift >> 5 << 12;
//       the match mask of the component is the AND of its rows shifted down by their shifts
//   checks, 2 words each, ordered by the level at which both ends are placed. Word 0 is the
//   end that is known BEFORE the check's level is entered (an earlier component or an absolute
//   position), word 1 the end on the component placed at the check's level:
//       anchor(int8) | off(int16) << 8 | flags << 24
//   kCheckFixed (on word 0) marks checks without that shape (both ends on the level's own
//   component, or both absolute): they are evaluated per candidate but give no band window.
//   kCheckOrdered (on word 0): both ends sit inside their components, word 0's component comes
//   first and delay >= 1, so u < v and both are inside the sequence without testing.
struct DevLibrary {
    const uint32_t* blob;
    const uint32_t* unit_off;    // [n_units + n_gates + 1] word offsets into blob
    const uint8_t*  unit_rlen;   // [n_units] words per record of the unit: components + 1
    const uint16_t* unit_probe;  // [n_units][4] k-mer probes (longest window of 4..6 consecutive literal symbols) of up to four components, 0xFFFF = none; null: no filter
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
    uint32_t smem_bytes;        // dynamic shared memory of the count kernel
    uint32_t smem_bytes_emit;   // ... of the emit kernel (adds the record staging)
};

struct Work {
    uint32_t* counts;        // [n_seqs * n_units] sites per item
    uint32_t* prefilter;     // [ceil(n_items / 32)] bit per item: not ruled out by the k-mer presence filter; null: no filter
    uint64_t* chunk_sites;   // [n_chunks]
    uint64_t* chunk_words;   // [n_chunks] ; after the scan: exclusive prefix, [n_chunks] = total
    uint64_t* totals;        // [0] words, [1] sites, [2] error flags, [3] items with sites (length of `alive`), [4] heavy items
    uint32_t* alive;         // [n_seqs * n_units] items with sites, in completion order
    uint32_t* lane_counts;   // [lane_cap][32] sites of every lane's run of ranks (rank-space walk)
    uint32_t  lane_cap;      // alive-list slots that have a lane_counts row
    uint32_t* heavy_items;   // [1024] items parked by the count pass (rank-space walk, too many placements for one warp)
    uint32_t* heavy_ranks;   // [1024] their placements before the checks
    uint32_t* task_sites;    // [1024 * 256] sites of every (heavy item, block) task
    uint32_t* task_lane_counts;   // [1024 * 256][32] sites of every lane's run inside a task
    uint64_t* seq_word_begin;   // [n_seqs + 1]
};

Plan make_plan(const DevLibrary& lib, uint32_t n_seqs, uint32_t max_len, int sm_count);

cudaError_t launch_band(const DevQuery& q, int32_t* band, cudaStream_t s);
// bit v of row u (u < v) = v >= u + 4 && u + 6 < n && pij[u * row_stride + v * col_stride] > theta
cudaError_t launch_pair_mask(const float* pij, uint32_t n, uint64_t row_stride, uint64_t col_stride, float theta, uint32_t* mask,
                             int sm_count, cudaStream_t s);
cudaError_t launch_gather_cut(const uint32_t* blocks, uint32_t world, uint64_t block_words, uint32_t header_words, uint32_t* out,
                              uint64_t* counts, uint32_t* overflow, cudaStream_t s);
cudaError_t launch_occ(const DevLibrary& lib, const DevQuery& q, uint32_t* occ, uint32_t* kmers, cudaStream_t s);
cudaError_t launch_count(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, int sm_count, cudaStream_t s);
cudaError_t launch_scan(const DevQuery& q, const Plan& plan, const Work& work, cudaStream_t s);
cudaError_t launch_emit(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records,
                        uint32_t n_alive, uint32_t n_heavy, uint64_t capacity_words, int sm_count, cudaStream_t s);
cudaError_t launch_guarded_copy(const Work& work, uint64_t capacity_words, uint64_t* dst, uint32_t n, cudaStream_t s);

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
