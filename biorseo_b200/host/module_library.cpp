#include "module_library.h"

#include <algorithm>
#include <cctype>
#include <climits>
#include <filesystem>
#include <cstring>
#include <fstream>
#include <sstream>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "mini_json.h"

namespace bsh {
namespace {

const uint32_t NO_GATE = 0xFFFFFFFFu;
const size_t   NPOS    = std::string::npos;

// What std::getline hands out line by line: split at '\n'; a final fragment without a
// newline is a line, a trailing newline does not open an empty one.
bool read_lines(const std::string& path, std::vector<std::string>& lines)
{
    lines.clear();
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string all = ss.str();
    size_t            a   = 0;
    while (a < all.size()) {
        size_t b = all.find('\n', a);
        if (b == NPOS) b = all.size();
        lines.push_back(all.substr(a, b - a));
        a = b + 1;
    }
    return true;
}

const std::string& line_or_empty(const std::vector<std::string>& lines, size_t i)
{
    static const std::string empty;
    return i < lines.size() ? lines[i] : empty;
}

// std::stoi on a prefix: optional blanks, optional sign, at least one digit, int range.
bool leading_int(const std::string& s, long& v)
{
    size_t i = 0;
    while (i < s.size() && std::isspace((unsigned char)s[i])) i++;
    bool neg = false;
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) neg = (s[i++] == '-');
    if (i >= s.size() || !std::isdigit((unsigned char)s[i])) return false;
    long x = 0;
    while (i < s.size() && std::isdigit((unsigned char)s[i])) {
        x = x * 10 + (s[i++] - '0');
        if (x > (long)INT_MAX + 1) return false;
    }
    v = neg ? -x : x;
    return v >= INT_MIN && v <= INT_MAX;
}

// substr with the reference's "npos + 1 wraps to 0" arithmetic; `from` beyond the end would
// throw in the reference, which callers treat as undefined input.
bool substr_from(const std::string& s, size_t from, std::string& out, size_t len = NPOS)
{
    if (from > s.size()) return false;
    out = s.substr(from, len);
    return true;
}

bool fits_i16(long v) { return v >= -32768 && v <= 32767; }

std::vector<std::string> sorted_regular_files(const std::string& dir, std::string& err)
{
    std::vector<std::string> files;
    std::error_code          ec;
    for (std::filesystem::recursive_directory_iterator it(dir, ec), end; !ec && it != end; it.increment(ec))
        if (it->is_regular_file()) files.push_back(it->path().string());
    if (ec) err = "cannot walk " + dir + ": " + ec.message();
    std::sort(files.begin(), files.end());
    return files;
}

}   // namespace

// A module sequence may only hold A, C, G, U, T (any case) and the strand separator '&'
// (reference behaviour: cppsrc/Motif.cpp:322-329).
bool json_sequence_ok(const std::string& seq)
{
    for (char ch : seq) {
        char c = (char)std::toupper((unsigned char)ch);
        if (!(c == 'A' || c == 'U' || c == 'T' || c == '&' || c == 'G' || c == 'C')) return false;
    }
    return true;
}

// Dots, '&' and four bracket families, each family balanced on its own
// (net effect of cppsrc/Motif.cpp:332-385: the repeated inner passes change nothing once
// one pass is balanced, and an unbalanced pass either underflows or leaves a non-empty stack).
bool json_struct_ok(const std::string& ss)
{
    int depth[4] = {0, 0, 0, 0};
    for (char c : ss) {
        switch (c) {
        case '.': case '&': break;
        case '(': depth[0]++; break;
        case '[': depth[1]++; break;
        case '{': depth[2]++; break;
        case '<': depth[3]++; break;
        case ')': if (--depth[0] < 0) return false; break;
        case ']': if (--depth[1] < 0) return false; break;
        case '}': if (--depth[2] < 0) return false; break;
        case '>': if (--depth[3] < 0) return false; break;
        default: return false;
        }
    }
    return depth[0] == 0 && depth[1] == 0 && depth[2] == 0 && depth[3] == 0;
}

void ModuleLibrary::clear(LibraryKind kind)
{
    unmap();
    mapped_table_ = bss_table{};
    modules_.clear(); rejected_.clear(); units_.clear(); gates_.clear();
    unit_module_.clear(); unit_delay_.clear(); unit_gate_.clear(); unit_comp_begin_.clear(); comp_pat_begin_.clear(); unit_check_begin_.clear();
    pattern_.clear(); checks_.clear();
    n_seen_ = 0;
    kind_   = kind;
}

ModuleLibrary::~ModuleLibrary() { unmap(); }

char ModuleLibrary::push_unit(uint32_t module, uint32_t delay, uint32_t gate, const std::vector<std::string>& patterns,
                              const std::vector<LinkSpec>& checks, bool is_gate)
{
    if (patterns.empty() || patterns.size() > BSS_MAX_COMPONENTS || checks.size() > BSS_MAX_CHECKS) return 'L';
    for (const std::string& p : patterns) {
        if (p.size() > BSS_MAX_PATTERN) return 'L';
        for (unsigned char c : p)
            if (!(std::isalnum(c) || c == '.')) return 'P';   // would be a regex metacharacter in the reference
    }
    Unit u{module, delay, gate, patterns, checks};
    (is_gate ? gates_ : units_).push_back(u);
    return 0;
}

// ---------------------------------------------------------------------------------------
// JSON   validity: cppsrc/Motif.cpp:291-319   components: cppsrc/MOIP.cpp:1158-1172
//        links: cppsrc/Motif.cpp:55-107        filter: cppsrc/MOIP.cpp:1188-1205
// ---------------------------------------------------------------------------------------
char ModuleLibrary::add_json_module(const std::string& key, const std::string& sequence, const std::string& struct2d)
{
    n_seen_++;
    auto reject = [&](char code) { rejected_.push_back({key, code}); return code; };

    if (struct2d.size() != sequence.size()) return reject('x');
    if (!json_sequence_ok(sequence)) return reject('r');
    if (!json_struct_ok(struct2d)) return reject('n');
    if (sequence.empty()) return reject('e');
    const size_t n_amp = (size_t)std::count(sequence.begin(), sequence.end(), '&');
    if (sequence.size() - n_amp < 4) return reject('l');
    {   // strands of at least two nucleotides must add up to more than three
        size_t total = 0, a = 0;
        for (;;) {
            size_t b   = sequence.find('&', a);
            size_t len = (b == NPOS ? sequence.size() : b) - a;
            if (len >= 2) total += len;
            if (b == NPOS) break;
            a = b + 1;
        }
        if (total <= 3) return reject('l');
    }

    // Strands: upper-cased, one per '&' (empty ones included), the remainder only if non-empty.
    std::vector<std::string> strands;
    {
        std::string up = sequence;
        for (char& c : up) c = (char)std::toupper((unsigned char)c);
        size_t a = 0;
        for (;;) {
            size_t b = up.find('&', a);
            if (b == NPOS) break;
            strands.push_back(up.substr(a, b - a));
            a = b + 1;
        }
        if (a < up.size()) strands.push_back(up.substr(a));
    }
    if (strands.empty()) return reject('U');   // cannot happen after the length checks; the reference would read vc[0] of an empty vector
    std::vector<long> start_of(strands.size() + 1, 0);
    for (size_t c = 0; c < strands.size(); c++) start_of[c + 1] = start_of[c] + (long)strands[c].size();

    // Links. Position of structure character i, with `amp` separators seen before it, is
    // first[amp] + (i - amp) - start_of[amp]: the reference accumulates the inter-strand gaps
    // (Motif.cpp:104), which telescopes to the start of strand `amp` whatever the alignment
    // of '&' between sequence and struct2d.
    std::vector<LinkSpec> links;
    std::vector<Endpoint> open[4];
    size_t                amp = 0;
    for (size_t i = 0; i < struct2d.size(); i++) {
        const char c = struct2d[i];
        if (c == '&') {
            if (++amp >= strands.size()) return reject('U');   // the reference reads v[count] out of bounds
            continue;
        }
        if (c == '.') continue;
        const long off = (long)(i - amp) - start_of[amp];
        if (!fits_i16(off)) return reject('L');
        const Endpoint here{(int16_t)amp, (int16_t)off};
        int            family = (c == '(' || c == ')') ? 0 : (c == '[' || c == ']') ? 1 : (c == '{' || c == '}') ? 2 : 3;
        if (c == '(' || c == '[' || c == '{' || c == '<') open[family].push_back(here);
        else {
            links.push_back({open[family].back(), here, false});
            open[family].pop_back();
        }
    }

    ModuleInfo m;
    m.id     = key;
    m.source = JSON;
    for (const std::string& s : strands) m.k.push_back((uint32_t)s.size());
    m.links      = links;
    m.first_unit = (uint32_t)units_.size();
    // A module without links can never produce a site (cppsrc/MOIP.cpp:1189): it is accepted, but owns no unit.
    if (!links.empty()) {
        if (char e = push_unit((uint32_t)modules_.size(), 1, NO_GATE, strands, links, false)) return reject(e);
        m.n_units = 1;
    }
    modules_.push_back(m);
    return 0;
}

// ---------------------------------------------------------------------------------------
// RIN    validity: cppsrc/Motif.cpp:263-289   components + id: cppsrc/MOIP.cpp:1105-1122
//        links: cppsrc/Motif.cpp:121-177       filter: cppsrc/MOIP.cpp:1131-1137
// ---------------------------------------------------------------------------------------
char ModuleLibrary::add_rin_file(const std::string& path)
{
    n_seen_++;
    const std::string shown = std::filesystem::path(path).stem().string();
    auto reject = [&](char code) { rejected_.push_back({shown, code}); return code; };

    std::vector<std::string> lines;
    read_lines(path, lines);   // an unreadable file behaves like an empty one
    const std::string& link_line = line_or_empty(lines, 1);

    // validity
    long motif_length = 0;
    for (size_t i = 3; i < lines.size(); i++) {
        const std::string& line = lines[i];
        std::string        tail;
        long               k;
        if (!substr_from(line, line.find(';') + 1, tail) || !leading_int(tail, k)) return reject('U');
        motif_length += k;
    }
    if (motif_length < 5) return reject('l');
    if (std::count(link_line.begin(), link_line.end(), ';') == 0) return reject('x');

    // module number: the digits after "Subfiles/", plus one
    long         number;
    const size_t sub = path.find("Subfiles/");
    if (sub == NPOS || !leading_int(path.substr(sub + 9, path.find(".txt")), number)) return reject('U');

    // component patterns: everything after the second ';'
    std::vector<std::string> strands;
    for (size_t i = 3; i < lines.size(); i++) {
        const std::string& line   = lines[i];
        const size_t       second = line.find(';', line.find(';') + 1);
        strands.push_back(line.substr(second + 1));
    }
    if (strands.empty()) return reject('U');

    // "a,b,Bool;" entries
    struct RawLink { long a, b; bool long_range; };
    std::vector<RawLink> raw;
    {
        size_t a = 0;
        while (a < link_line.size()) {
            const size_t semi = link_line.find(';', a);
            if (semi == NPOS) return reject('U');   // the reference never consumes such a tail
            const std::string entry = link_line.substr(a, semi - a);
            a                       = semi + 1;
            RawLink      l;
            const size_t c1 = entry.find(',');
            if (c1 == NPOS || !leading_int(entry.substr(0, c1), l.a)) return reject('U');
            const std::string rest = entry.substr(c1 + 1);
            const size_t      c2   = rest.find(',');
            if (c2 == NPOS || !leading_int(rest.substr(0, c2), l.b)) return reject('U');
            if (l.a < 0 || l.b < 0) return reject('U');
            l.long_range = (rest.substr(c2 + 1) == "True");
            raw.push_back(l);
        }
    }

    // Renumbering. `a` lands in the strand jA that contains motif index a. `b` is looked up
    // from strand jA onwards in windows shifted by k[jA] and anchored on the strand's LAST
    // nucleotide; when no window holds it, it stays the raw motif index.
    std::vector<long> k(strands.size()), start_of(strands.size() + 1, 0);
    for (size_t c = 0; c < strands.size(); c++) {
        k[c]            = (long)strands[c].size();
        start_of[c + 1] = start_of[c] + k[c];
    }
    std::vector<LinkSpec> links;
    for (const RawLink& l : raw) {
        LinkSpec spec{{BSS_ANCHOR_ABSOLUTE, 0}, {BSS_ANCHOR_ABSOLUTE, 0}, l.long_range};
        long     a_off = l.a, b_off = l.b;
        size_t   jA = strands.size();
        for (size_t j = 0; j < strands.size(); j++)
            if (l.a >= start_of[j] && l.a < start_of[j] + k[j]) { jA = j; break; }
        if (jA < strands.size()) {
            spec.a.anchor = (int16_t)jA;
            a_off         = l.a - start_of[jA];
            for (size_t j = jA; j < strands.size(); j++) {
                const long lo = start_of[j] + k[jA];
                if (l.b >= lo && l.b < lo + k[j]) {
                    spec.b.anchor = (int16_t)j;
                    b_off         = (k[j] - 1) + (l.b - lo);
                    break;
                }
            }
        }
        if (!fits_i16(a_off) || !fits_i16(b_off)) return reject('L');
        spec.a.off = (int16_t)a_off;
        spec.b.off = (int16_t)b_off;
        links.push_back(spec);
    }

    ModuleInfo m;
    m.id     = std::to_string(1 + number);
    m.source = CARNAVAL;
    for (long kk : k) m.k.push_back((uint32_t)kk);
    m.links      = links;
    m.first_unit = (uint32_t)units_.size();
    if (char e = push_unit((uint32_t)modules_.size(), 1, NO_GATE, strands, links, false)) return reject(e);
    m.n_units = 1;
    modules_.push_back(m);
    return 0;
}

// ---------------------------------------------------------------------------------------
// DESC   validity: cppsrc/Motif.cpp:218-261   pre-filter: cppsrc/Motif.cpp:387-430
//        components + variants: cppsrc/MOIP.cpp:932-1066   filter: cppsrc/MOIP.cpp:1072-1079
// ---------------------------------------------------------------------------------------
char ModuleLibrary::add_desc_file(const std::string& path)
{
    n_seen_++;
    const std::string stem = std::filesystem::path(path).stem().string();
    auto reject = [&](char code) { rejected_.push_back({stem, code}); return code; };

    std::vector<std::string> lines;
    read_lines(path, lines);

    // "Bases: 866_G  867_G ..." is cut at every blank that follows a blank or a colon.
    std::vector<std::string> tokens(1);
    {
        char prev = 'a';
        for (char c : line_or_empty(lines, 1)) {
            const bool cut = (c == ' ') && (prev == ' ' || prev == ':');
            prev           = c;
            if (cut) tokens.emplace_back();
            else tokens.back().push_back(c);
        }
    }
    if (tokens.size() < 3) return reject('U');   // no base between the label and the trailing token

    // The reference walks tokens [1, size-1) and may step back onto token 0 ("Bases:").
    struct Base { char nt; long pos; bool ok; };
    std::vector<Base> base(tokens.size());
    for (size_t t = 0; t < tokens.size(); t++) {
        const std::string& tok = tokens[t];
        const size_t       us  = tok.find('_');
        std::string        one;
        base[t].ok = substr_from(tok, us + 1, one, 1) && !one.empty() && leading_int(tok.substr(0, us), base[t].pos);
        base[t].nt = one.empty() ? '\0' : one.back();
        if (t == 0) base[t].ok = !one.empty();   // only its "nucleotide" is ever read
    }
    const size_t last_tok = tokens.size() - 1;   // exclusive end of the bases
    for (size_t t = 1; t < last_tok; t++) {
        if (!base[t].ok) return reject('U');
        if (std::string("ACGU").find(base[t].nt) == NPOS) return reject(base[t].nt);
        if (base[t].pos <= 0) return reject('-');
    }

    // interaction lines
    for (size_t i = 2; i < lines.size(); i++) {
        const std::string& line  = lines[i];
        const size_t       slash = line.find('/');
        if (slash == NPOS || slash == 0) return reject('U');
        const std::string interaction = line.substr(slash - 1, 3);
        std::string       b1 = line.substr(line.find('(') + 1, 8);
        const std::string temp = line.substr(slash + 1);
        std::string       b2 = temp.substr(temp.find('(') + 1, 8);
        b1.erase(std::remove(b1.begin(), b1.end(), ' '), b1.end());
        b2.erase(std::remove(b2.begin(), b2.end(), ' '), b2.end());
        long p1, p2;
        if (!leading_int(b1.substr(0, b1.find('_')), p1) || !leading_int(b2.substr(0, b2.find('_')), p2)) return reject('U');
        if (p2 - p1 != 1 && interaction == "C/C") return reject('b');
        if (p2 - p1 < 4 && (interaction == "+/+" || interaction == "-/-")) return reject('l');
    }

    // Strands: a gap of more than 5 in the numbering opens a new one, gaps of 2..5 become dots.
    auto filler = [](long gap) { return (gap >= 2 && gap <= 5) ? std::string((size_t)gap - 1, '.') : std::string(); };
    std::vector<std::string> strands;
    {
        std::string seq;
        long        last = base[1].pos;
        for (size_t t = 1; t < last_tok; t++) {
            const long gap = base[t].pos - last;
            if (gap > 5) { strands.push_back(seq); seq.clear(); }
            else seq += filler(gap);
            seq += base[t].nt;
            last = base[t].pos;
        }
        strands.push_back(seq);
    }

    // Variants: strands of one or two nucleotides are widened to three in every possible way.
    // This follows the reference's token walk literally, including its side effects: the
    // widening reads the nucleotide of the PREVIOUS token (the "Bases:" label for a leading
    // short strand), the strand in progress is dropped, and the strand still open at the end
    // of the walk is never appended (cppsrc/MOIP.cpp:984-1052).
    std::vector<std::vector<std::string>> variants;
    std::vector<size_t>                   short1, short2;
    for (size_t p = 0; p < strands.size(); p++) {
        if (strands[p].size() == 1) short1.push_back(p);
        if (strands[p].size() == 2) short2.push_back(p);
    }
    if (short1.empty() && short2.empty()) variants.push_back(strands);
    else {
        variants.emplace_back();
        size_t      current = 0;
        std::string seq;
        long        last = base[1].pos;
        for (size_t t = 1; t < last_tok; t++) {
            const long pos = base[t].pos;
            if (!short1.empty() && current == short1.front()) {
                if (!base[t - 1].ok) return reject('U');
                const char        nt = base[t - 1].nt;
                const std::string s1 = std::string(1, nt) + "..", s2 = "." + std::string(1, nt) + ".", s3 = ".." + std::string(1, nt);
                const size_t      before = variants.size();
                for (size_t u = 0; u < before; u++) {
                    variants.push_back(variants[u]); variants.back().push_back(s2);
                    variants.push_back(variants[u]); variants.back().push_back(s3);
                    variants[u].push_back(s1);
                }
                seq.clear();
                current++;
                short1.erase(short1.begin());
                last = pos;
                t--;   // the same token is looked at again
            } else if (!short2.empty() && current == short2.front()) {
                if (!base[t - 1].ok) return reject('U');
                const char        nt = base[t - 1].nt, next = base[t].nt;
                last                 = base[t].pos;
                const std::string s1 = std::string(1, nt) + std::string(1, next) + ".", s2 = "." + std::string(1, nt) + std::string(1, next);
                const size_t      before = variants.size();
                for (size_t u = 0; u < before; u++) {
                    variants.push_back(variants[u]); variants.back().push_back(s2);
                    variants[u].push_back(s1);
                }
                seq.clear();
                current++;
                short2.erase(short2.begin());
            } else {
                const long gap = pos - last;
                if (gap > 5) {
                    current++;
                    for (auto& v : variants) v.push_back(seq);
                    seq.clear();
                } else seq += filler(gap);
                seq += base[t].nt;
                last = pos;
            }
        }
    }

    ModuleInfo m;
    m.id     = stem;
    m.source = RNA3DMOTIF;
    for (const std::string& s : variants[0]) m.k.push_back((uint32_t)s.size());
    m.first_unit = (uint32_t)units_.size();

    // Gate = the regex pre-filter: original strands, at least five bases between them.
    const uint32_t gate      = (uint32_t)gates_.size();
    const size_t   units_was = units_.size();
    if (char e = push_unit((uint32_t)modules_.size(), 6, NO_GATE, strands, {}, true)) return reject(e);
    for (const auto& v : variants) {
        if (v.empty()) { units_.resize(units_was); gates_.resize(gate); return reject('U'); }
        // closing pairs: first nt of the first strand with last nt of the last one, and the
        // last nt of each strand with the first nt of the next
        std::vector<LinkSpec> checks;
        const size_t          C = v.size();
        checks.push_back({{0, 0}, {(int16_t)(C - 1), (int16_t)((long)v[C - 1].size() - 1)}, false});
        for (size_t c = 0; c + 1 < C; c++) checks.push_back({{(int16_t)c, (int16_t)((long)v[c].size() - 1)}, {(int16_t)(c + 1), 0}, false});
        if (char e = push_unit((uint32_t)modules_.size(), 5, gate, v, checks, false)) {
            units_.resize(units_was);
            gates_.resize(gate);
            return reject(e);
        }
        m.n_units++;
    }
    modules_.push_back(m);
    return 0;
}

bool ModuleLibrary::load(LibraryKind kind, const std::string& path, std::string& err)
{
    clear(kind);
    if (!std::filesystem::exists(path)) { err = "cannot find " + path; return false; }
    if (kind == KIND_JSON) {
        std::ifstream f(path, std::ios::binary);
        if (!f) { err = "cannot open " + path; return false; }
        std::stringstream ss;
        ss << f.rdbuf();
        std::map<std::string, JsonModule> entries;
        if (!parse_module_json(ss.str(), entries, err)) return false;
        for (const auto& e : entries) {
            if (!e.second.is_object || !e.second.has_sequence || !e.second.has_struct2d) {
                n_seen_++;
                rejected_.push_back({e.first, 'U'});   // the reference dies on a type error here
                continue;
            }
            add_json_module(e.first, e.second.sequence, e.second.struct2d);
        }
    } else {
        for (const std::string& file : sorted_regular_files(path, err)) {
            if (kind == KIND_RIN) add_rin_file(file);
            else add_desc_file(file);
        }
        if (!err.empty()) return false;
    }
    finish();
    return true;
}

// ---- compiled library file -------------------------------------------------------------------
// Everything load() produces, laid out so that the file can be MAPPED and used in place: a fixed little-endian header,
// an offset table, then 16-byte aligned sections holding plain arrays. The flat component table the device library is
// built from (bss_table) is eight of those sections: after load_compiled its pointers point straight into the mapping —
// nothing is parsed, nothing is copied (the pages are read by the upload itself). Module identifiers, link recipes and
// the reject log (what turns records back into Motif objects) are sections too and are unpacked into modules_ / rejected_.
//   header   "BSHLIB2\0", u32 kind, u32 sections, u32 modules, u32 units, u32 gates, u32 0, u64 modules seen, u64 file size
//   table    sections x { u32 id, u32 bytes per element, u64 byte offset, u64 elements }
// Every offset, count and index is checked against the file before use; a damaged file is refused, never trusted.
namespace {

const char kMagic[8] = {'B', 'S', 'H', 'L', 'I', 'B', '2', 0};

enum SectionId : uint32_t {
    S_UNIT_MODULE = 1, S_UNIT_DELAY, S_UNIT_GATE, S_UNIT_COMP_BEGIN, S_COMP_PAT_BEGIN, S_PATTERN, S_UNIT_CHECK_BEGIN, S_CHECKS,
    S_MOD_FIRST_UNIT, S_MOD_N_UNITS, S_MOD_SOURCE, S_MOD_K_BEGIN, S_MOD_K, S_MOD_LINK_BEGIN, S_MOD_LINKS, S_MOD_ID_BEGIN, S_MOD_ID,
    S_REJ_ID_BEGIN, S_REJ_ID, S_REJ_CODE, S_END
};
constexpr uint32_t kSections = S_END - 1;
const uint32_t     kElemSize[kSections] = {4, 4, 4, 4, 4, 1, 4, 2, 4, 4, 4, 4, 4, 4, 2, 4, 1, 4, 1, 1};

struct FileHeader {
    char     magic[8];
    uint32_t kind, n_sections, n_modules, n_units, n_gates, reserved;
    uint64_t n_seen, file_size;
};
struct SectionEntry {
    uint32_t id, elem_size;
    uint64_t offset, count;
};
static_assert(sizeof(FileHeader) == 48 && sizeof(SectionEntry) == 24, "the file layout is fixed");

template <typename T>
const T* section(const char* base, const SectionEntry* table, SectionId id) { return reinterpret_cast<const T*>(base + table[id - 1].offset); }

// begin[] has n + 1 entries, starts at 0, never decreases and ends at `total`
bool offsets_ok(const uint32_t* begin, uint64_t entries, uint64_t n, uint64_t total)
{
    if (entries != n + 1 || begin[0] != 0 || begin[n] != total) return false;
    for (uint64_t i = 0; i < n; i++)
        if (begin[i] > begin[i + 1]) return false;
    return true;
}

}   // namespace

bool ModuleLibrary::save_compiled(const std::string& path, std::string& err) const
{
    if (map_base_) {   // a mapped library is its file
        std::ofstream f(path, std::ios::binary | std::ios::trunc);
        if (!f || !f.write(static_cast<const char*>(map_base_), (std::streamsize)map_size_)) { err = "cannot write " + path; return false; }
        return true;
    }
    // the module side as arrays
    std::vector<uint32_t> first_unit, n_units, source, k_begin(1, 0), k, link_begin(1, 0), id_begin(1, 0), rej_begin(1, 0);
    std::vector<int16_t>  links;
    std::vector<uint8_t>  ids, rej_ids, rej_codes;
    for (const ModuleInfo& m : modules_) {
        first_unit.push_back(m.first_unit);
        n_units.push_back(m.n_units);
        source.push_back((uint32_t)m.source);
        k.insert(k.end(), m.k.begin(), m.k.end());
        k_begin.push_back((uint32_t)k.size());
        for (const LinkSpec& l : m.links) {
            const int16_t packed[5] = {l.a.anchor, l.a.off, l.b.anchor, l.b.off, (int16_t)(l.long_range ? 1 : 0)};
            links.insert(links.end(), packed, packed + 5);
        }
        link_begin.push_back((uint32_t)(links.size() / 5));
        ids.insert(ids.end(), m.id.begin(), m.id.end());
        id_begin.push_back((uint32_t)ids.size());
    }
    for (const Rejected& r : rejected_) {
        rej_ids.insert(rej_ids.end(), r.id.begin(), r.id.end());
        rej_begin.push_back((uint32_t)rej_ids.size());
        rej_codes.push_back((uint8_t)r.code);
    }
    const bss_table t = table();
    const uint32_t  n_all = t.n_units + t.n_gates, n_comps = unit_comp_begin_.back(), n_checks = unit_check_begin_.back();
    struct Piece { const void* data; uint64_t count; };
    const Piece pieces[kSections] = {
        {t.unit_module, n_all}, {t.unit_delay, n_all}, {t.unit_gate, t.n_units}, {t.unit_comp_begin, (uint64_t)n_all + 1}, {t.comp_pat_begin, (uint64_t)n_comps + 1},
        {t.pattern, pattern_.size()}, {t.unit_check_begin, (uint64_t)n_all + 1}, {t.checks, 4ull * n_checks},
        {first_unit.data(), first_unit.size()}, {n_units.data(), n_units.size()}, {source.data(), source.size()}, {k_begin.data(), k_begin.size()}, {k.data(), k.size()},
        {link_begin.data(), link_begin.size()}, {links.data(), links.size()}, {id_begin.data(), id_begin.size()}, {ids.data(), ids.size()},
        {rej_begin.data(), rej_begin.size()}, {rej_ids.data(), rej_ids.size()}, {rej_codes.data(), rej_codes.size()}};
    FileHeader   h{};
    SectionEntry entries[kSections];
    std::memcpy(h.magic, kMagic, 8);
    h.kind = (uint32_t)kind_; h.n_sections = kSections; h.n_modules = t.n_modules; h.n_units = t.n_units; h.n_gates = t.n_gates; h.n_seen = n_seen_;
    uint64_t at = (sizeof h + sizeof entries + 15) & ~15ull;
    for (uint32_t i = 0; i < kSections; i++) {
        entries[i] = {i + 1, kElemSize[i], at, pieces[i].count};
        at         = (at + pieces[i].count * kElemSize[i] + 15) & ~15ull;
    }
    h.file_size = at;
    std::string out((size_t)at, '\0');
    std::memcpy(&out[0], &h, sizeof h);
    std::memcpy(&out[sizeof h], entries, sizeof entries);
    for (uint32_t i = 0; i < kSections; i++)
        if (pieces[i].count) std::memcpy(&out[(size_t)entries[i].offset], pieces[i].data, (size_t)(pieces[i].count * kElemSize[i]));
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f || !f.write(out.data(), (std::streamsize)out.size())) { err = "cannot write " + path; return false; }
    return true;
}

void ModuleLibrary::unmap()
{
    if (map_base_) munmap(map_base_, map_size_);
    map_base_ = nullptr;
    map_size_ = 0;
}

bool ModuleLibrary::load_compiled(const std::string& path, std::string& err)
{
    clear(KIND_JSON);
    const int fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) { err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < (off_t)(sizeof(FileHeader) + kSections * sizeof(SectionEntry))) {
        ::close(fd);
        err = path + " is not a compiled module library";
        return false;
    }
    void* base = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);   // (the mapping keeps the file)
    if (base == MAP_FAILED) { err = "cannot map " + path; return false; }
    map_base_ = base;
    map_size_ = (size_t)st.st_size;
    auto refuse = [&](const std::string& why) { clear(KIND_JSON); err = path + ": " + why; return false; };
    const char*       bytes = static_cast<const char*>(base);
    const FileHeader& h     = *reinterpret_cast<const FileHeader*>(bytes);
    if (std::memcmp(h.magic, kMagic, 8) != 0) return refuse("not a compiled module library");
    if (h.kind > 2 || h.n_sections != kSections || h.file_size != map_size_) return refuse("truncated or corrupt compiled library");
    const SectionEntry* e = reinterpret_cast<const SectionEntry*>(bytes + sizeof(FileHeader));
    uint64_t            floor = sizeof(FileHeader) + kSections * sizeof(SectionEntry);
    for (uint32_t i = 0; i < kSections; i++) {
        if (e[i].id != i + 1 || e[i].elem_size != kElemSize[i] || (e[i].offset & 15) || e[i].offset < floor || e[i].offset > map_size_ ||
            e[i].count > (map_size_ - e[i].offset) / kElemSize[i])
            return refuse("truncated or corrupt compiled library");
        floor = e[i].offset + e[i].count * kElemSize[i];
    }
    // the flat table, in place
    const uint64_t  n_all = (uint64_t)h.n_units + h.n_gates;
    const uint32_t *ucb = section<uint32_t>(bytes, e, S_UNIT_COMP_BEGIN), *cpb = section<uint32_t>(bytes, e, S_COMP_PAT_BEGIN),
                   *uck = section<uint32_t>(bytes, e, S_UNIT_CHECK_BEGIN), *umod = section<uint32_t>(bytes, e, S_UNIT_MODULE),
                   *ugate = section<uint32_t>(bytes, e, S_UNIT_GATE);
    bool sane = e[S_UNIT_MODULE - 1].count == n_all && e[S_UNIT_DELAY - 1].count == n_all && e[S_UNIT_GATE - 1].count == h.n_units &&
                e[S_UNIT_COMP_BEGIN - 1].count == n_all + 1 && e[S_UNIT_CHECK_BEGIN - 1].count == n_all + 1 && e[S_COMP_PAT_BEGIN - 1].count >= 1;
    sane = sane && offsets_ok(ucb, e[S_UNIT_COMP_BEGIN - 1].count, n_all, e[S_COMP_PAT_BEGIN - 1].count - 1);
    sane = sane && offsets_ok(cpb, e[S_COMP_PAT_BEGIN - 1].count, e[S_COMP_PAT_BEGIN - 1].count - 1, e[S_PATTERN - 1].count);
    sane = sane && offsets_ok(uck, e[S_UNIT_CHECK_BEGIN - 1].count, n_all, e[S_CHECKS - 1].count / 4) && e[S_CHECKS - 1].count % 4 == 0;
    for (uint64_t u = 0; sane && u < n_all; u++) sane = umod[u] < h.n_modules && ucb[u + 1] > ucb[u];
    for (uint64_t u = 0; sane && u < h.n_units; u++) sane = ugate[u] == NO_GATE || (ugate[u] >= h.n_units && ugate[u] < n_all);
    // the module side
    const uint32_t *first_unit = section<uint32_t>(bytes, e, S_MOD_FIRST_UNIT), *n_units = section<uint32_t>(bytes, e, S_MOD_N_UNITS),
                   *source = section<uint32_t>(bytes, e, S_MOD_SOURCE), *k_begin = section<uint32_t>(bytes, e, S_MOD_K_BEGIN),
                   *k = section<uint32_t>(bytes, e, S_MOD_K), *link_begin = section<uint32_t>(bytes, e, S_MOD_LINK_BEGIN),
                   *id_begin = section<uint32_t>(bytes, e, S_MOD_ID_BEGIN), *rej_begin = section<uint32_t>(bytes, e, S_REJ_ID_BEGIN);
    const int16_t* links = section<int16_t>(bytes, e, S_MOD_LINKS);
    const uint64_t n_rej = e[S_REJ_CODE - 1].count;
    sane = sane && e[S_MOD_FIRST_UNIT - 1].count == h.n_modules && e[S_MOD_N_UNITS - 1].count == h.n_modules && e[S_MOD_SOURCE - 1].count == h.n_modules &&
           e[S_MOD_LINKS - 1].count % 5 == 0;
    sane = sane && offsets_ok(k_begin, e[S_MOD_K_BEGIN - 1].count, h.n_modules, e[S_MOD_K - 1].count);
    sane = sane && offsets_ok(link_begin, e[S_MOD_LINK_BEGIN - 1].count, h.n_modules, e[S_MOD_LINKS - 1].count / 5);
    sane = sane && offsets_ok(id_begin, e[S_MOD_ID_BEGIN - 1].count, h.n_modules, e[S_MOD_ID - 1].count);
    sane = sane && offsets_ok(rej_begin, e[S_REJ_ID_BEGIN - 1].count, n_rej, e[S_REJ_ID - 1].count);
    for (uint32_t m = 0; sane && m < h.n_modules; m++) {
        const uint32_t C = k_begin[m + 1] - k_begin[m];
        sane = C >= 1 && (uint64_t)first_unit[m] + n_units[m] <= h.n_units;
        // every unit of the module places C components (rebuild() reads p[0 .. C-1] of its records), every link end names one of them
        for (uint32_t u = first_unit[m]; sane && u < first_unit[m] + n_units[m]; u++) sane = umod[u] == m && ucb[u + 1] - ucb[u] == C;
        for (uint32_t l = link_begin[m]; sane && l < link_begin[m + 1]; l++)
            sane = links[5 * l] >= -1 && links[5 * l] < (int)C && links[5 * l + 2] >= -1 && links[5 * l + 2] < (int)C;
    }
    if (!sane) return refuse("truncated or corrupt compiled library");
    kind_   = (LibraryKind)h.kind;
    n_seen_ = (size_t)h.n_seen;
    modules_.resize(h.n_modules);
    const char* ids = section<char>(bytes, e, S_MOD_ID);
    for (uint32_t m = 0; m < h.n_modules; m++) {
        ModuleInfo& info = modules_[m];
        info.id.assign(ids + id_begin[m], ids + id_begin[m + 1]);
        info.source = (source_type)source[m];
        info.k.assign(k + k_begin[m], k + k_begin[m + 1]);
        info.links.resize(link_begin[m + 1] - link_begin[m]);
        for (size_t l = 0; l < info.links.size(); l++) {
            const int16_t* p = links + 5 * (link_begin[m] + l);
            info.links[l]    = LinkSpec{{p[0], p[1]}, {p[2], p[3]}, p[4] != 0};
        }
        info.first_unit = first_unit[m];
        info.n_units    = n_units[m];
    }
    const char*    rej_ids   = section<char>(bytes, e, S_REJ_ID);
    const uint8_t* rej_codes = section<uint8_t>(bytes, e, S_REJ_CODE);
    rejected_.resize(n_rej);
    for (uint64_t i = 0; i < n_rej; i++) rejected_[i] = Rejected{std::string(rej_ids + rej_begin[i], rej_ids + rej_begin[i + 1]), (char)rej_codes[i]};
    mapped_table_.n_modules        = h.n_modules;
    mapped_table_.n_units          = h.n_units;
    mapped_table_.n_gates          = h.n_gates;
    mapped_table_.unit_module      = umod;
    mapped_table_.unit_delay       = section<uint32_t>(bytes, e, S_UNIT_DELAY);
    mapped_table_.unit_gate        = ugate;
    mapped_table_.unit_comp_begin  = ucb;
    mapped_table_.comp_pat_begin   = cpb;
    mapped_table_.pattern          = section<uint8_t>(bytes, e, S_PATTERN);
    mapped_table_.unit_check_begin = uck;
    mapped_table_.checks           = section<int16_t>(bytes, e, S_CHECKS);
    return true;
}

void ModuleLibrary::finish()
{
    unit_module_.clear(); unit_delay_.clear(); unit_gate_.clear(); unit_comp_begin_.assign(1, 0);
    comp_pat_begin_.assign(1, 0); unit_check_begin_.assign(1, 0); pattern_.clear(); checks_.clear();
    auto emit = [&](const Unit& u, bool searched) {
        unit_module_.push_back(u.module);
        unit_delay_.push_back(u.delay);
        if (searched) unit_gate_.push_back(u.gate == NO_GATE ? NO_GATE : (uint32_t)units_.size() + u.gate);
        for (const std::string& p : u.patterns) {
            pattern_.insert(pattern_.end(), p.begin(), p.end());
            comp_pat_begin_.push_back((uint32_t)pattern_.size());
        }
        unit_comp_begin_.push_back((uint32_t)comp_pat_begin_.size() - 1);
        for (const LinkSpec& c : u.checks) {
            checks_.push_back(c.a.anchor); checks_.push_back(c.a.off);
            checks_.push_back(c.b.anchor); checks_.push_back(c.b.off);
        }
        unit_check_begin_.push_back((uint32_t)(checks_.size() / 4));
    };
    for (const Unit& u : units_) emit(u, true);
    for (const Unit& u : gates_) emit(u, false);
}

bss_table ModuleLibrary::table() const
{
    if (map_base_) return mapped_table_;   // the arrays of the mapped file, in place
    bss_table t;
    t.n_modules        = (uint32_t)modules_.size();
    t.n_units          = (uint32_t)units_.size();
    t.n_gates          = (uint32_t)gates_.size();
    t.unit_module      = unit_module_.data();
    t.unit_delay       = unit_delay_.data();
    t.unit_gate        = unit_gate_.data();
    t.unit_comp_begin  = unit_comp_begin_.data();
    t.comp_pat_begin   = comp_pat_begin_.data();
    t.pattern          = pattern_.data();
    t.unit_check_begin = unit_check_begin_.data();
    t.checks           = checks_.data();
    return t;
}

Motif ModuleLibrary::rebuild(const uint32_t* record) const
{
    const ModuleInfo&      m = modules_[record[0]];
    const uint32_t*        p = record + 1;
    std::vector<Component> comps;
    for (size_t c = 0; c < m.k.size(); c++) {
        const uint first = p[c], second = p[c] + m.k[c] - 1u;   // wraps exactly like the reference's uint arithmetic
        comps.push_back(Component(std::make_pair((int)first, (int)second)));
    }
    Motif site(comps, m.id, m.source);
    auto  where = [&](const Endpoint& e) -> uint { return (e.anchor == BSS_ANCHOR_ABSOLUTE ? 0u : p[e.anchor]) + (uint)(int)e.off; };
    for (const LinkSpec& l : m.links) {
        Link link;
        link.nts.first  = where(l.a);
        link.nts.second = where(l.b);
        link.long_range = l.long_range;
        site.links_.push_back(link);
    }
    return site;
}

}   // namespace bsh
