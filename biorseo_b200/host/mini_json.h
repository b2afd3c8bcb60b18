// Minimal JSON reader for motif libraries of the form
//   { "<module id>": { "sequence": "...", "struct2d": "...", <anything else> }, ... }
// (format: reference SOURCES.md:14-23). It keeps what the reference's parser keeps for this
// path: members ordered by byte-wise key comparison and last-one-wins on duplicate keys
// (the reference iterates a std::map-backed object, cppsrc/MOIP.cpp:260-262).
#ifndef BIORSEO_B200_MINI_JSON_H_
#define BIORSEO_B200_MINI_JSON_H_

#include <map>
#include <string>

namespace bsh {

struct JsonModule {
    bool        is_object    = false;
    bool        has_sequence = false;   // present and a string
    bool        has_struct2d = false;
    std::string sequence;
    std::string struct2d;
};

// Returns false and sets `err` on malformed JSON or a non-object top level.
bool parse_module_json(const std::string& text, std::map<std::string, JsonModule>& out, std::string& err);

}   // namespace bsh
#endif
