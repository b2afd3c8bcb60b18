// Unit-descriptor layout of the flat device library: constants and the host-side builder.
// Plain C++ (no CUDA headers): shared by the library build (bss_capi.cu), the kernels
// (through bss_device.cuh) and the CPU emulation of the window walk used by the tests.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct bss_table;

namespace bss {

constexpr uint32_t kWildcard     = 0xFFu;      // pattern symbol code of '.'
constexpr uint32_t kNoGate       = 0xFFFFFFFFu;
constexpr uint32_t kPairAlphabet = 8;          // largest alphabet that gets pair rows (row index must stay below 128)
// k-mer presence filter (alphabets of at most 4 symbols, 2 bits per symbol): per sequence one bit per
// possible 4-, 5- and 6-mer; a probe is code | table << 12 (table 0, 1, 2 = k 4, 5, 6), 0xFFFF = none
constexpr uint32_t kKmerMin = 4, kKmerMax = 6;
constexpr uint32_t kKmerWords = (256 + 1024 + 4096) / 32;   // presence bits per sequence, the three tables back to back

constexpr uint32_t kCompOverlap = 1u << 24;    // component flag: the pattern is compatible with a shifted copy of itself
constexpr uint32_t kCompGeneric = 1u << 25;    // component flag: some check of this level is not a row check (window walk: tested per candidate)
constexpr uint32_t kCheckFixed  = 1u << 24;    // check flag: no end moves at the check's level, or both do
constexpr uint32_t kCheckOrdered = 1u << 25;   // check flag: 0 <= u < v < n holds for every placement (u = word 0's end)
constexpr uint32_t kCheckRow    = 1u << 26;    // check flag: u < v holds for every placement, u on an earlier component, v = candidate + offset >= candidate:
                                               //             the candidates that pass are the bits of mask row u shifted down by the offset (window walk)
constexpr uint32_t kUnitWindow  = 1u;          // blob[3] / unit_flags bit: every level >= 1 has a row check and the descriptor fits the staging area (window walk eligible)
constexpr uint32_t kUnitLean    = 2u;          // unit_flags bit: window-walk unit without a gate, at most four components, every pattern in the pattern table
constexpr uint32_t kNoSlot      = 0xFFFFu;     // unit_slots entry: the component's pattern has no slot in the pattern table
constexpr uint32_t kSlotMinUses = 3;           // a pattern gets a slot when at least this many components of searched units share it
constexpr uint32_t kSlotMax     = 0xFFF0u;     // slots per library
constexpr uint32_t kBlobStageWords = 64;       // unit descriptor words staged in shared memory per warp; larger descriptors are read in place
constexpr int      kWindowBandMax = 255;       // widest pair (v - u) a sequence's mask may allow for the window walk (beyond, the rank-space passes win)

// Unit descriptor ("blob"), u32 words, one per unit, addressed by unit_off[unit]:
//   [0] module index
//   [1] C | n_checks << 8 | delay << 16
//   [2] gate unit or kNoGate
//   [3] word offset of the checks << 16 | unit flags (kUnitWindow)
//   [4 + 2c]   k_c | offset of the match program << 8 (16 bits, in u16 units from the start of the blob) | kCompOverlap | kCompGeneric
//   [5 + 2c]   first check ready at level c | one past the last << 8 | steps of the match program << 16 | row checks of the level << 24
//              (the row checks are the first of their level)
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
//   kCheckRow (on word 0; every ordered check is one): word 0's end sits on an earlier component at or before
//   its last nucleotide (it may lie before its first: u < 0 is tested), word 1's offset is >= 0 and delay >= 1.
struct BuiltLibrary {
    std::vector<uint32_t> blob, unit_off;          // unit_off: [n_units + n_gates + 1]
    std::vector<uint8_t>  unit_rlen, unit_flags;   // [max(1, n_units)]
    std::vector<uint16_t> unit_probe;              // [4 * max(1, n_units)]
    // Pattern table: the distinct component patterns that several units share get a slot; per query the match set (and the
    // thinned set) of every slot is computed ONCE (pattern_kernel) instead of once per unit that uses the pattern.
    std::vector<uint32_t> pat_blob, pat_off;       // one-component descriptors of the slots (same layout as a unit descriptor), pat_off: [n_slots + 1]
    std::vector<uint16_t> pat_tslot;               // [n_slots] row of the slot's thinned set, kNoSlot for patterns that cannot overlap themselves
    std::vector<uint16_t> unit_slots;              // [4 * max(1, n_units)] slots of the first four components, kNoSlot = none
    uint32_t              n_slots = 0, n_tslots = 0, n_lean_units = 0;
    std::vector<uint32_t> module_components;       // [n_modules]
    std::vector<uint8_t>  alphabet;
    uint32_t              cmax = 1, n_window_units = 0;
    float                 list_density = 0.f;      // expected matches per nucleotide of a unit (all components), 98th percentile
    bool                  has_probes = false;
};

// Validates the table and builds the descriptors. Returns 0 or a BSS_ERR_* code with `error` set.
int build_library(const bss_table* t, BuiltLibrary& out, std::string& error);

}   // namespace bss
