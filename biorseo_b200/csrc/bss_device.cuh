// Device-side data layout and kernel launch interface of the insertion-site search.
// Internal to the library; the public boundary is include/biorseo_b200.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bss {

constexpr int      kThreads    = 128;          // threads per CTA (4 warps)
constexpr uint32_t kWildcard   = 0xFFu;        // pattern symbol code of '.'
constexpr int      kOccPad     = 10;           // zero words after each occurrence row (patterns up to 255+32 bytes)
constexpr int      kBlobStage  = 64;           // unit descriptor words staged in shared memory per group (cooperative kernel)
constexpr uint32_t kNoGate     = 0xFFFFFFFFu;
constexpr uint32_t kMaxChunkUnits = 2048;      // units per chunk (bounds the per-CTA offset table)
constexpr int      kFastMaxComponents = 8;     // units with more components take the cooperative kernel
constexpr uint32_t kCheckStatic = 1u << 24;    // check flag: both ends move together (same component)

// Unit descriptor ("blob"), u32 words, one per unit, addressed by unit_off[unit]:
//   [0] module index
//   [1] C | n_checks << 8 | delay << 16
//   [2] gate unit or kNoGate
//   [3] flags | word offset of the checks << 16 ; flag bit0 = component 0 can overlap itself
//       (its match chain from position 0 must be thinned to leftmost non-overlapping matches)
//   [4 + 2c]   k_c | pattern byte offset << 8   (offset from the start of the blob, in bytes)
//   [5 + 2c]   first check ready at level c | one past the last << 8
//   pattern symbol codes (1 byte each, kWildcard = '.'), padded to a word
//   checks, 2 words each, ordered by the level at which both ends are placed. Word 0 is the
//   end that is placed FIRST (lower level, or absolute), word 1 the end placed at the check's
//   level:   anchor(int8) | off(int16) << 8 | flags << 24
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
    const int32_t*  band;         // [n_seqs] max v-u over the set mask bits, -1 when the mask is empty
    uint32_t        n_seqs;
    uint32_t        max_len;
};

// Work decomposition: chunk = (sequence, contiguous unit range); one CTA per chunk.
struct Plan {
    uint32_t units_per_chunk;
    uint32_t chunks_per_seq;
    uint32_t n_chunks;
    uint32_t group;        // cooperative kernel: lanes per (sequence, unit) item: 4, 8, 16 or 32
    uint32_t W;            // match-mask words per component, ceil((max_len + 1) / 32)
    uint32_t WS;           // summary words per component, ceil(W / 32)
    uint32_t smem_bytes;   // cooperative kernel
    uint32_t max_entries;  // lane-per-item kernel: match-list entries per lane (all components together)
    uint32_t mask_words;   // lane-per-item kernel: pair-mask words staged in shared memory (0 = read from global)
    uint32_t fast_smem_bytes;
};

struct Work {
    uint32_t* counts;        // [n_seqs * n_units] sites per item
    uint32_t* slow_bits;     // [ceil(n_items / 32)] items left to the cooperative kernel
    uint64_t* chunk_sites;   // [n_chunks]
    uint64_t* chunk_words;   // [n_chunks] ; after the scan: exclusive prefix, [n_chunks] = total
    uint64_t* totals;        // [0] words, [1] sites, [2] error flags, [3] number of slow items
    uint64_t* seq_word_begin;   // [n_seqs + 1]
};

Plan make_plan(const DevLibrary& lib, uint32_t n_seqs, uint32_t max_len, int sm_count);
cudaError_t configure_kernels();   // one-time function attributes (dynamic shared memory limits)

cudaError_t launch_band(const DevQuery& q, int32_t* band, cudaStream_t s);
cudaError_t launch_count(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, cudaStream_t s);
cudaError_t launch_scan(const DevQuery& q, const Plan& plan, const Work& work, cudaStream_t s);
cudaError_t launch_emit(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records,
                        cudaStream_t s);

}   // namespace bss
