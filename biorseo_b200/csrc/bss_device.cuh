// Device-side data layout and kernel launch interface of the insertion-site search.
// Internal to the library; the public boundary is include/biorseo_b200.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bss {

constexpr int      kThreads    = 128;          // threads per CTA (4 warps)
constexpr uint32_t kWildcard   = 0xFFu;        // pattern symbol code of '.'
constexpr int      kOccPad     = 10;           // zero words after each occurrence row (patterns up to 255+32 bytes)
constexpr int      kBlobStage  = 64;           // unit descriptor words staged in shared memory per group
constexpr uint32_t kNoGate     = 0xFFFFFFFFu;
constexpr uint32_t kMaxChunkUnits = 2048;      // units per chunk (bounds the per-CTA offset table)
constexpr uint32_t kCompOverlap = 1u << 24;    // component flag: the pattern is compatible with a shifted copy of itself
constexpr uint32_t kCheckFixed  = 1u << 24;    // check flag: no end moves at the check's level, or both do

// Unit descriptor ("blob"), u32 words, one per unit, addressed by unit_off[unit]:
//   [0] module index
//   [1] C | n_checks << 8 | delay << 16
//   [2] gate unit or kNoGate
//   [3] word offset of the checks << 16
//   [4 + 2c]   k_c | pattern byte offset << 8 (16 bits, from the start of the blob) | kCompOverlap
//   [5 + 2c]   first check ready at level c | one past the last << 8
//   pattern symbol codes (1 byte each, kWildcard = '.'), padded to a word
//   checks, 2 words each, ordered by the level at which both ends are placed. Word 0 is the
//   end that is known BEFORE the check's level is entered (an earlier component or an absolute
//   position), word 1 the end on the component placed at the check's level:
//       anchor(int8) | off(int16) << 8 | flags << 24
//   kCheckFixed (on word 0) marks checks without that shape (both ends on the level's own
//   component, or both absolute): they are evaluated per candidate but give no band window.
struct DevLibrary {
    const uint32_t* blob;
    const uint32_t* unit_off;    // [n_units + n_gates + 1] word offsets into blob
    uint32_t        n_units;     // searched units
    uint32_t        n_symbols;   // alphabet size
    uint32_t        cmax;        // max components per unit (gates included)
    uint32_t        module_base; // added to every module index written into a record (library shards)
    uint8_t         alphabet[32];
};

struct DevQuery {
    const uint8_t*  seqs;
    const uint64_t* seq_begin;    // [n_seqs + 1]
    const uint32_t* masks;
    const uint64_t* mask_begin;   // [n_seqs + 1], in words
    const int32_t*  band;         // [n_seqs] max v-u over the set bits above the diagonal, -1 when there is none
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
    uint32_t list_cap;     // rank-space walk: match-list entries per item (0 = depth-first walk only)
    uint32_t smem_bytes;
};

struct Work {
    uint32_t* counts;        // [n_seqs * n_units] sites per item
    uint64_t* chunk_sites;   // [n_chunks]
    uint64_t* chunk_words;   // [n_chunks] ; after the scan: exclusive prefix, [n_chunks] = total
    uint64_t* totals;        // [0] words, [1] sites, [2] error flags
    uint64_t* seq_word_begin;   // [n_seqs + 1]
};

Plan make_plan(const DevLibrary& lib, uint32_t n_seqs, uint32_t max_len, int sm_count);

cudaError_t launch_band(const DevQuery& q, int32_t* band, cudaStream_t s);
cudaError_t launch_count(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, cudaStream_t s);
cudaError_t launch_scan(const DevQuery& q, const Plan& plan, const Work& work, cudaStream_t s);
cudaError_t launch_emit(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records,
                        cudaStream_t s);

}   // namespace bss
