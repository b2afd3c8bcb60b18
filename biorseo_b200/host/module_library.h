// Motif-library loaders: JSON (RNA 3D Motif Atlas / Rna3Dmotifs style), CaRNAval RIN files
// and Rna3Dmotif .desc files  ->  one flat component table for the device (bss_table,
// include/biorseo_b200.h) plus what is needed to turn search records back into Motif objects.
//
// Each loader accepts and rejects exactly the modules the reference accepts and rejects
// (cppsrc/Motif.cpp:218-385), extracts the same component patterns
// (cppsrc/MOIP.cpp:932-1066, 1108-1122, 1158-1172) and encodes the reference's link
// coordinates (cppsrc/Motif.cpp:55-107, 154-177) and pair filters
// (cppsrc/MOIP.cpp:1072-1079, 1131-1137, 1188-1205) as position-independent "checks".
// Inputs on which the reference's behaviour is undefined (out-of-bounds reads, uncaught
// exceptions, endless loops) are rejected with code 'U' instead of being imitated.
#ifndef BIORSEO_B200_MODULE_LIBRARY_H_
#define BIORSEO_B200_MODULE_LIBRARY_H_

#include <cstdint>
#include <string>
#include <vector>

#include "Motif.h"
#include "biorseo_b200.h"

namespace bsh {

enum LibraryKind { KIND_JSON = 0, KIND_RIN = 1, KIND_DESC = 2 };

// One end of a base pair, relative to a placement: position = start of component `anchor`
// + off, or just `off` when anchor == BSS_ANCHOR_ABSOLUTE.
struct Endpoint {
    int16_t anchor;
    int16_t off;
};

struct LinkSpec {
    Endpoint a, b;
    bool     long_range;
};

struct ModuleInfo {
    std::string           id;       // Motif::id_ (JSON key, RIN number, .desc stem)
    source_type           source;
    std::vector<uint32_t> k;        // component lengths of every unit of this module
    std::vector<LinkSpec> links;    // Motif::links_ recipe (empty for DESC)
    uint32_t              first_unit = 0, n_units = 0;
};

struct Rejected {
    std::string id;
    char        code;   // reference codes ('l','x','e','r','n','b','-', a nucleotide) or 'U' / 'P' / 'L' / '0'
};

class ModuleLibrary
{
  public:
    ModuleLibrary() {}
    ~ModuleLibrary();
    ModuleLibrary(const ModuleLibrary&) = delete;              // (may own a file mapping)
    ModuleLibrary& operator=(const ModuleLibrary&) = delete;

    // Loads `path` (a .json file, or a folder of RIN / .desc files). Returns false with
    // `err` set when the library itself cannot be read; bad modules are only rejected.
    bool load(LibraryKind kind, const std::string& path, std::string& err);

    LibraryKind kind() const { return kind_; }
    uint32_t    delay() const { return kind_ == KIND_DESC ? 5u : 1u; }   // cppsrc/biorseo.cpp:151,156
    const std::vector<ModuleInfo>& modules() const { return modules_; }
    const std::vector<Rejected>&   rejected() const { return rejected_; }
    size_t n_seen() const { return n_seen_; }   // modules examined, valid or not

    // View of the flat table (pointers into this object).
    bss_table table() const;

    // Number of u32 words of a record of `module`.
    uint32_t record_words(uint32_t module) const { return 1u + (uint32_t)modules_[module].k.size(); }
    // Record [module, p_0 .. p_{C-1}] -> the Motif the reference would have pushed.
    Motif rebuild(const uint32_t* record) const;

    // Compiled form: everything load() produces (accepted modules with their link recipes, the flat
    // table of searchable units, the reject log), in one little-endian binary file of plain aligned
    // arrays. load_compiled MAPS it: table() then points into the mapping (no parse, no copy); only the
    // module identifiers and link recipes are unpacked. It replaces the per-run JSON parse / directory
    // walk / per-module validation of cppsrc/MOIP.cpp:159-188, 211-234, 259-288, and the RIN re-read
    // of every placement (cppsrc/Motif.cpp:115-151).
    bool save_compiled(const std::string& path, std::string& err) const;
    bool load_compiled(const std::string& path, std::string& err);
    // Size of the file mapping behind table() (0: the table lives in this object's own arrays).
    size_t mapped_bytes() const { return map_size_; }

    // Direct construction (synthetic libraries in benchmarks and tests).
    void clear(LibraryKind kind);
    // Adds one JSON-style module; returns the reject code or 0.
    char add_json_module(const std::string& key, const std::string& sequence, const std::string& struct2d);
    // Rebuilds the flat table after add_json_module calls.
    void seal() { finish(); }

  private:
    char add_rin_file(const std::string& path);
    char add_desc_file(const std::string& path);
    // Appends a searched unit; returns false when it breaks a BSS_MAX_* limit or uses
    // pattern bytes outside [A-Za-z0-9.].
    char push_unit(uint32_t module, uint32_t delay, uint32_t gate, const std::vector<std::string>& patterns,
                   const std::vector<LinkSpec>& checks, bool is_gate);
    void finish();
    void unmap();

    // a compiled library mapped from disk: the flat arrays live in the mapping
    void*     map_base_ = nullptr;
    size_t    map_size_ = 0;
    bss_table mapped_table_{};

    LibraryKind             kind_ = KIND_JSON;
    std::vector<ModuleInfo> modules_;
    std::vector<Rejected>   rejected_;
    size_t                  n_seen_ = 0;

    struct Unit {
        uint32_t                 module, delay, gate;
        std::vector<std::string> patterns;
        std::vector<LinkSpec>    checks;
    };
    std::vector<Unit> units_, gates_;

    // flat arrays
    std::vector<uint32_t> unit_module_, unit_delay_, unit_gate_, unit_comp_begin_, comp_pat_begin_, unit_check_begin_;
    std::vector<uint8_t>  pattern_;
    std::vector<int16_t>  checks_;
};

// Helpers shared with tests.
bool json_sequence_ok(const std::string& seq);   // cppsrc/Motif.cpp:322-329
bool json_struct_ok(const std::string& ss);      // cppsrc/Motif.cpp:332-385

}   // namespace bsh
#endif
