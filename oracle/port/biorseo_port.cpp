// ORACLE — test infrastructure, not product code. Nothing in the shipped library includes,
// links or executes this file.
//
// Plain C++ restatement of Biorseo's insertion-site search (reference v2.1), written
// without std::regex, Boost, nlohmann/json, CPLEX or threads, so that it can be built and
// run anywhere (the GPU box has no /root/reference). Every function cites the reference
// lines it restates. PARITY PINNED: tests/test_oracle.py checks this program against the
// reference's own compiled code (oracle/_ref/biorseo_ref) on the golden vectors of
// SURVEY.md appendix A, on tests/golden/*, and on seeded random libraries.
//
//   biorseo_port jobs|ctor <jsonfolder|rinfolder|descfolder> <library> <seqfile> <maskfile> [ignored...]
//
// Same arguments and same output as oracle/harness/ref_cli.cpp: one insertion site per
// line, "identifier \t first-second ... \t a-b ...". Inputs on which the reference's
// behaviour is undefined make the program exit with status 3.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

using std::string;
using std::vector;
typedef unsigned int uint;

static const size_t npos = string::npos;

[[noreturn]] static void undefined_in_reference(const string& why)
{
    std::fprintf(stderr, "undefined behaviour in the reference: %s\n", why.c_str());
    std::exit(3);
}

struct Comp { uint first, second; size_t k; };          // cppsrc/Motif.h:22-32
struct Link { uint a, b; };                             // cppsrc/Motif.h:35-39
struct Site { string identifier; vector<Comp> comp; vector<Link> links; };

// ---------------------------------------------------------------------------------------
// Query: sequence + pij > theta table
// ---------------------------------------------------------------------------------------
struct Query {
    string           seq;
    size_t           n = 0, W = 0;
    vector<uint32_t> mask;   // bit v of row u: pij(u, v) > theta
    bool pij_above_theta(size_t u, size_t v) const { return (mask[u * W + (v >> 5)] >> (v & 31)) & 1u; }
    // cppsrc/MOIP.cpp:896-910, with index_of_yuv_ (cppsrc/MOIP.cpp:77-90) folded in
    bool allowed_basepair(size_t u, size_t v) const
    {
        size_t a = (v > u) ? u : v, b = (v > u) ? v : u;
        if (b - a < 4) return false;
        if (a >= n - 6) return false;   // size_t arithmetic like the reference; n >= 6 is required there too
        if (b >= n) return false;
        return pij_above_theta(a, b);
    }
};

// ---------------------------------------------------------------------------------------
// Pattern matching without regex. A pattern is literals and '.', cppsrc/Motif.cpp:438.
// ---------------------------------------------------------------------------------------
static bool matches_at(const string& text, size_t at, const string& pat)
{
    if (at + pat.size() > text.size()) return false;
    for (size_t j = 0; j < pat.size(); j++) {
        if (pat[j] == '.') {
            if (text[at + j] == '\n' || text[at + j] == '\r') return false;   // ECMAScript '.' skips line terminators
        } else if (pat[j] != text[at + j]) return false;
    }
    return true;
}

// What std::sregex_iterator yields for such a pattern: leftmost match, then the leftmost
// match starting at or after its end (one further for an empty match), and so on.
static vector<size_t> iterate_matches(const string& text, const string& pat)
{
    vector<size_t> at;
    size_t         from = 0;
    while (from <= text.size()) {
        size_t q = from;
        while (q <= text.size() && !matches_at(text, q, pat)) q++;
        if (q > text.size()) break;
        at.push_back(q);
        from = q + (pat.empty() ? 1 : pat.size());
    }
    return at;
}

// cppsrc/Motif.cpp:432-494 find_next_ones_in(rna, offset, vc)
static vector<vector<Comp>> find_next_ones_in(const string& rna, uint offset, const vector<string>& vc, uint delay)
{
    vector<vector<Comp>> results;
    if (vc.empty()) undefined_in_reference("find_next_ones_in with no component");
    const string& pat = vc[0];
    if (vc.size() > 1) {
        vector<string> next_seqs(vc.begin() + 1, vc.end());
        for (size_t m : iterate_matches(rna, pat)) {
            uint first = (uint)m + offset, second = first + (uint)pat.size() - 1;
            if ((uint)(second - offset + delay) >= rna.length()) break;   // Motif.cpp:461 (uint arithmetic)
            vector<vector<Comp>> next_ones =
                find_next_ones_in(rna.substr(second - offset + delay), second + delay, next_seqs, delay);
            if (next_ones.empty()) break;                                 // Motif.cpp:463
            for (const vector<Comp>& v : next_ones) {
                vector<Comp> r;
                r.push_back(Comp{first, second, (size_t)(1u + second - first)});
                r.insert(r.end(), v.begin(), v.end());
                results.push_back(r);
            }
        }
    } else {
        for (size_t m : iterate_matches(rna, pat)) {
            uint first = (uint)m + offset, second = first + (uint)pat.size() - 1;
            results.push_back(vector<Comp>{Comp{first, second, (size_t)(1u + second - first)}});
        }
    }
    return results;
}

// ---------------------------------------------------------------------------------------
// Small parsing helpers with the standard library's semantics
// ---------------------------------------------------------------------------------------
static vector<string> getlines(const string& path)   // successive std::getline calls
{
    std::ifstream f(path, std::ios::binary);
    vector<string> lines;
    string         line;
    while (std::getline(f, line)) lines.push_back(line);
    return lines;
}
static string line_at(const vector<string>& lines, size_t i) { return i < lines.size() ? lines[i] : string(); }

static int stoi_or_die(const string& s)   // std::stoi; throwing = uncaught exception in the reference
{
    try {
        return std::stoi(s);
    } catch (...) {
        undefined_in_reference("std::stoi(\"" + s + "\") throws");
    }
}
static string substr_or_die(const string& s, size_t from, size_t len = npos)
{
    if (from > s.size()) undefined_in_reference("substr beyond the end of \"" + s + "\"");
    return s.substr(from, len);
}

// ---------------------------------------------------------------------------------------
// JSON modules
// ---------------------------------------------------------------------------------------
// Tiny reader for {"key": {"sequence": "...", "struct2d": "...", ...}, ...}. Members come
// back ordered by key like the reference's std::map-backed object (cppsrc/MOIP.cpp:262).
struct JsonReader {
    const string& s;
    size_t        i = 0;
    explicit JsonReader(const string& text) : s(text) {}
    void ws() { while (i < s.size() && std::strchr(" \t\r\n", s[i])) i++; }
    void expect(char c)
    {
        ws();
        if (i >= s.size() || s[i] != c) { std::fprintf(stderr, "json: expected '%c' at %zu\n", c, i); std::exit(2); }
        i++;
    }
    bool peek(char c) { ws(); return i < s.size() && s[i] == c; }
    string str()
    {
        expect('"');
        string out;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\') {
                i++;
                switch (s[i]) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': {
                    unsigned cp = (unsigned)std::stoul(s.substr(i + 1, 4), nullptr, 16);
                    i += 4;
                    if (cp < 0x80) out += (char)cp;
                    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                    else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                    break;
                }
                default: out += s[i];
                }
                i++;
            } else out += s[i++];
        }
        expect('"');
        return out;
    }
    void skip_value()
    {
        ws();
        if (peek('"')) { str(); return; }
        if (peek('{')) {
            expect('{');
            if (!peek('}')) do { str(); expect(':'); skip_value(); } while (peek(',') && (i++, true));
            expect('}');
            return;
        }
        if (peek('[')) {
            expect('[');
            if (!peek(']')) do { skip_value(); } while (peek(',') && (i++, true));
            expect(']');
            return;
        }
        while (i < s.size() && !std::strchr(",}] \t\r\n", s[i])) i++;
    }
};

struct JsonEntry { bool has_seq = false, has_ss = false; string seq, ss; };

static std::map<string, JsonEntry> read_json_library(const string& path)
{
    std::ifstream     f(path, std::ios::binary);
    std::stringstream buf;
    buf << f.rdbuf();
    const string text = buf.str();
    JsonReader   r(text);
    std::map<string, JsonEntry> lib;
    r.expect('{');
    if (!r.peek('}')) do {
        string    key = r.str();
        JsonEntry e;
        r.expect(':');
        if (!r.peek('{')) undefined_in_reference("JSON module that is not an object");
        r.expect('{');
        if (!r.peek('}')) do {
            string field = r.str();
            r.expect(':');
            if ((field == "sequence" || field == "struct2d") && r.peek('"')) {
                string v = r.str();
                if (field == "sequence") { e.seq = v; e.has_seq = true; } else { e.ss = v; e.has_ss = true; }
            } else {
                if (field == "sequence") e.has_seq = false;
                if (field == "struct2d") e.has_ss = false;
                r.skip_value();
            }
        } while (r.peek(',') && (r.i++, true));
        r.expect('}');
        lib[key] = e;
    } while (r.peek(',') && (r.i++, true));
    r.expect('}');
    return lib;
}

// cppsrc/Motif.cpp:322-329
static bool check_motif_sequence(string seq)
{
    for (char& c : seq) c = (char)::toupper((unsigned char)c);
    for (char c : seq)
        if (!(c == 'A' || c == 'U' || c == 'T' || c == '&' || c == 'G' || c == 'C')) return false;
    return true;
}

// cppsrc/Motif.cpp:332-385, kept literal: for every position, a full pass over the string
// on stacks that are NOT reset between passes.
static bool check_motif_ss(const string& struc)
{
    vector<uint> round_, square, curly, angle;
    for (size_t i = 0; i < struc.size(); i++) {
        if (!std::strchr("().&[]{}<>", struc[i]) || struc[i] == '\0') return false;
        for (size_t j = 0; j < struc.size(); j++) {
            switch (struc[j]) {
            case '(': round_.push_back((uint)j); break;
            case ')': if (round_.empty()) return false; round_.pop_back(); break;
            case '[': square.push_back((uint)j); break;
            case ']': if (square.empty()) return false; square.pop_back(); break;
            case '{': curly.push_back((uint)j); break;
            case '}': if (curly.empty()) return false; curly.pop_back(); break;
            case '<': angle.push_back((uint)j); break;
            case '>': if (angle.empty()) return false; angle.pop_back(); break;
            default: break;
            }
        }
    }
    return round_.empty() && square.empty() && curly.empty() && angle.empty();
}

// cppsrc/Motif.cpp:291-319
static char is_valid_JSON(const JsonEntry& e)
{
    string seq = e.seq;
    const string& ss = e.ss;
    if (ss.size() != seq.size()) return 'x';
    if (!check_motif_sequence(seq)) return 'r';
    if (!check_motif_ss(ss)) return 'n';
    if (seq.empty()) return 'e';
    if (seq.size() - (size_t)std::count(seq.begin(), seq.end(), '&') < 4) return 'l';
    vector<string> components;
    while (seq.find('&') != npos) {
        size_t end    = seq.find('&');
        string subseq = seq.substr(0, end);
        seq           = seq.substr(end + 1);
        if (subseq.size() >= 2) components.push_back(subseq);
    }
    if (seq.size() >= 2) components.push_back(seq);
    size_t total = 0;
    for (const string& c : components) total += c.size();
    if (total <= 3) return 'l';
    return 0;
}

// cppsrc/Motif.cpp:55-107: links of a placed JSON module from its struct2d
static vector<Link> json_links(const vector<Comp>& v, const string& struc)
{
    vector<Link> links;
    vector<uint> stacks[4];
    uint         count = 0, debut = v[0].first, gap = 0;
    for (uint i = 0; i < struc.size(); i++) {
        const char  c     = struc[i];
        const char* open  = std::strchr("([{<", c);
        const char* close = std::strchr(")]}>", c);
        if (c != '\0' && open) stacks[open - "([{<"].push_back(i + debut + gap - count);
        else if (c != '\0' && close) {
            vector<uint>& st = stacks[close - ")]}>"];
            if (st.empty()) undefined_in_reference("closing bracket on an empty stack");
            links.push_back(Link{st.back(), i + debut + gap - count});
            st.pop_back();
        } else if (c == '&') {
            count++;
            if (count >= v.size()) undefined_in_reference("struct2d has more '&' than the sequence has components");
            gap += v[count].first - v[count - 1].second - 1;
        }
    }
    return links;
}

// cppsrc/MOIP.cpp:1146-1213
static void allowed_motifs_from_json(const Query& q, const string& key, const JsonEntry& e, vector<Site>& sites)
{
    string seq = e.seq;
    for (char& c : seq) c = (char)::toupper((unsigned char)c);
    vector<string> component_sequences;
    while (seq.find('&') != npos) {
        size_t end = seq.find('&');
        component_sequences.push_back(seq.substr(0, end));
        seq = seq.substr(end + 1);
    }
    if (!seq.empty()) component_sequences.push_back(seq);
    for (const vector<Comp>& v : find_next_ones_in(q.seq, 0, component_sequences, 1)) {
        Site s{"JSON" + key, v, json_links(v, e.ss)};
        if (s.links.empty()) continue;
        bool unprobable = false;
        for (const Link& l : s.links)
            if (!q.allowed_basepair(l.a, l.b)) unprobable = true;
        if (!unprobable) sites.push_back(s);
    }
}

// ---------------------------------------------------------------------------------------
// CaRNAval RIN modules
// ---------------------------------------------------------------------------------------
// cppsrc/Motif.cpp:263-289
static char is_valid_RIN(const string& rinfile)
{
    vector<string> lines = getlines(rinfile);
    const string   links = line_at(lines, 1);
    size_t n_basepairs   = (size_t)std::count(links.begin(), links.end(), ';');
    uint   motif_length  = 0;
    for (size_t i = 3; i < lines.size(); i++) {
        const string& line = lines[i];
        uint          index = (uint)line.find(";", line.find(";") + 1);
        motif_length += (uint)stoi_or_die(substr_or_die(line, line.find(";") + 1, index));
    }
    if (motif_length < 5) return 'l';
    if (!n_basepairs) return 'x';
    return 0;
}

// cppsrc/Motif.cpp:109-180: links of a placed RIN, renumbered onto the sequence
static vector<Link> rin_links(const vector<Comp>& v, const string& rinfile)
{
    vector<Link>   links;
    vector<string> lines = getlines(rinfile);
    string         line  = line_at(lines, 1);
    while (line != "") {
        size_t index = line.find(";");
        if (index == npos) undefined_in_reference("link line without a final ';' never ends");
        string link_str = line.substr(0, index);
        line.erase(0, index + 1);
        size_t sub = link_str.find(",");
        uint   a   = (uint)stoi_or_die(link_str.substr(0, sub));
        link_str.erase(0, sub + 1);
        sub    = link_str.find(",");
        uint b = (uint)stoi_or_die(link_str.substr(0, sub));
        links.push_back(Link{a, b});
    }
    for (Link& l : links) {
        size_t sum_comp = 0, j;
        for (j = 0; j < v.size(); j++) {
            if (l.a >= sum_comp && l.a < sum_comp + v[j].k) {
                l.a += v[j].first - (uint)sum_comp;
                sum_comp += v[j].k;
                break;
            }
            sum_comp += v[j].k;
        }
        for (; j < v.size(); j++) {
            if (l.b >= sum_comp && l.b < sum_comp + v[j].k) {
                l.b += v[j].second - (uint)sum_comp;
                break;
            }
            sum_comp += v[j].k;
        }
    }
    return links;
}

// cppsrc/MOIP.cpp:1088-1144
static void allowed_motifs_from_rin(const Query& q, const string& filepath, vector<Site>& sites)
{
    size_t sub = filepath.find("Subfiles/");
    if (sub == npos) undefined_in_reference("RIN path without Subfiles/");
    uint carnaval_id = 1 + (uint)stoi_or_die(filepath.substr(sub + 9, filepath.find(".txt")));
    vector<string> lines = getlines(filepath), component_sequences;
    for (size_t i = 3; i < lines.size(); i++) {
        size_t index = lines[i].find(';', lines[i].find(';') + 1);
        component_sequences.push_back(lines[i].substr(index + 1, npos));
    }
    for (const vector<Comp>& v : find_next_ones_in(q.seq, 0, component_sequences, 1)) {
        Site s{"RIN" + std::to_string(carnaval_id), v, rin_links(v, filepath)};
        bool unprobable = false;
        for (const Link& l : s.links)
            if (!q.allowed_basepair(l.a, l.b)) unprobable = true;
        if (!unprobable) sites.push_back(s);
    }
}

// ---------------------------------------------------------------------------------------
// Rna3Dmotif .desc modules
// ---------------------------------------------------------------------------------------
// The reference's boost::split with a stateful predicate (cppsrc/Motif.cpp:231-235): cut at
// each blank that follows a blank or a colon.
static vector<string> split_bases(const string& line)
{
    vector<string> tokens(1);
    char           prev = 'a';
    for (char c : line) {
        bool res = (prev == ' ' || prev == ':');
        prev     = c;
        if (c == ' ' && res) tokens.emplace_back();
        else tokens.back() += c;
    }
    return tokens;
}
static char token_nt(const string& tok)
{
    string one = substr_or_die(tok, tok.find('_') + 1, 1);
    if (one.empty()) undefined_in_reference("back() of an empty string");
    return one.back();
}
static int token_pos(const string& tok) { return stoi_or_die(tok.substr(0, tok.find('_'))); }

// cppsrc/Motif.cpp:218-261
static char is_valid_DESC(const string& descfile)
{
    vector<string> lines = getlines(descfile);
    vector<string> bases = split_bases(line_at(lines, 1));
    if (bases.size() < 2) undefined_in_reference("iterator walk past the end of the bases");
    for (size_t b = 1; b + 1 < bases.size(); b++) {
        char nt  = token_nt(bases[b]);
        int  pos = token_pos(bases[b]);
        if (string(1, nt).find_first_not_of("ACGU") != npos) return nt;
        if (pos <= 0) return '-';
    }
    for (size_t i = 2; i < lines.size(); i++) {
        const string& line  = lines[i];
        uint          slash = (uint)line.find('/');
        string interaction  = substr_or_die(line, (size_t)(uint)(slash - 1), 3);
        string b1           = line.substr(line.find('(') + 1, 8);
        string temp         = substr_or_die(line, (size_t)slash + 1);
        string b2           = temp.substr(temp.find('(') + 1, 8);
        b1.erase(std::remove(b1.begin(), b1.end(), ' '), b1.end());
        b2.erase(std::remove(b2.begin(), b2.end(), ' '), b2.end());
        int p1 = stoi_or_die(b1.substr(0, b1.find('_')));
        int p2 = stoi_or_die(b2.substr(0, b2.find('_')));
        if ((p2 - p1 != 1) && interaction == "C/C") return 'b';
        if ((p2 - p1 < 4) && (interaction == "+/+" || interaction == "-/-")) return 'l';
    }
    return 0;
}

static string dots_for_gap(int gap)   // cppsrc/MOIP.cpp:950-957
{
    if (gap == 2) return ".";
    if (gap == 3) return "..";
    if (gap == 4) return "...";
    if (gap == 5) return "....";
    return "";
}

// cppsrc/Motif.cpp:387-430. The regex built there is strand (.{5,} strand)*, so a search
// succeeds iff the strands can be laid out left to right with at least five bytes between
// them; taking the earliest possible match of each strand decides that.
static bool is_desc_insertible(const string& descfile, const string& rna)
{
    vector<string> bases = split_bases(line_at(getlines(descfile), 1));
    if (bases.size() < 2) undefined_in_reference("bases[1] of a one-token line");
    vector<string> strands(1);
    int            last = token_pos(bases[1]);
    for (size_t b = 1; b + 1 < bases.size(); b++) {
        char nt  = token_nt(bases[b]);
        int  pos = token_pos(bases[b]);
        if (pos - last > 5) strands.emplace_back();
        else strands.back() += dots_for_gap(pos - last);
        strands.back() += nt;
        last = pos;
    }
    size_t from = 0;
    for (size_t c = 0; c < strands.size(); c++) {
        size_t q = from;
        while (q <= rna.size() && !matches_at(rna, q, strands[c])) q++;
        if (q > rna.size()) return false;
        from = q + strands[c].size() + 5;
    }
    return true;
}

// cppsrc/MOIP.cpp:912-1086
static void allowed_motifs_from_desc(const Query& q, const string& descfile, vector<Site>& sites)
{
    vector<string> bases = split_bases(line_at(getlines(descfile), 1));
    vector<string> component_sequences;
    string         seq;
    int            last = token_pos(bases[1]);
    for (size_t b = 1; b + 1 < bases.size(); b++) {
        char nt  = token_nt(bases[b]);
        int  pos = token_pos(bases[b]);
        if (pos - last > 5) { component_sequences.push_back(seq); seq = ""; }
        else seq += dots_for_gap(pos - last);
        seq += nt;
        last = pos;
    }
    component_sequences.push_back(seq);

    vector<uint> comp_of_size_1, comp_of_size_2;
    for (uint p = 0; p < component_sequences.size(); ++p) {
        if (component_sequences[p].length() == 1) comp_of_size_1.push_back(p);
        if (component_sequences[p].length() == 2) comp_of_size_2.push_back(p);
    }
    vector<vector<Comp>> vresults;
    if (comp_of_size_1.size() || comp_of_size_2.size()) {
        vector<vector<string>> motif_variants(1);
        uint actual_comp = 0;
        seq  = "";
        last = token_pos(bases[1]);
        // `b` is the reference's iterator; it is decremented and re-incremented inside the body.
        for (long b = 1; b < (long)bases.size() - 1; b++) {
            int  pos = token_pos(bases[b]);
            char nt  = token_nt(bases[b]);
            if (comp_of_size_1.size() && actual_comp == comp_of_size_1[0]) {
                b--;
                nt = token_nt(bases[b]);
                string seq1 = string(1, nt) + "..", seq2 = "." + string(1, nt) + ".", seq3 = ".." + string(1, nt);
                uint   end  = (uint)motif_variants.size();
                for (uint u = 0; u < end; ++u) {
                    motif_variants.push_back(motif_variants[u]); motif_variants.back().push_back(seq2);
                    motif_variants.push_back(motif_variants[u]); motif_variants.back().push_back(seq3);
                    motif_variants[u].push_back(seq1);
                }
                seq = "";
                actual_comp++;
                comp_of_size_1.erase(comp_of_size_1.begin());
                last = pos;
            } else if (comp_of_size_2.size() && actual_comp == comp_of_size_2[0]) {
                b--;
                nt = token_nt(bases[b]);
                b++;
                char next   = token_nt(bases[b]);
                last        = token_pos(bases[b]);
                string seq1 = string(1, nt) + string(1, next) + ".", seq2 = "." + string(1, nt) + string(1, next);
                uint   end  = (uint)motif_variants.size();
                for (uint u = 0; u < end; ++u) {
                    motif_variants.push_back(motif_variants[u]); motif_variants.back().push_back(seq2);
                    motif_variants[u].push_back(seq1);
                }
                seq = "";
                actual_comp++;
                comp_of_size_2.erase(comp_of_size_2.begin());
            } else {
                if (pos - last > 5) {
                    actual_comp++;
                    for (vector<string>& c_s : motif_variants) c_s.push_back(seq);
                    seq = "";
                } else seq += dots_for_gap(pos - last);
                seq += nt;
                last = pos;
            }
        }
        // cppsrc/MOIP.cpp:1051-1052 appends the open strand to COPIES of the variants: no effect.
        for (const vector<string>& c_s : motif_variants) {
            vector<vector<Comp>> r = find_next_ones_in(q.seq, 0, c_s, 5);
            vresults.insert(vresults.end(), r.begin(), r.end());
        }
    } else vresults = find_next_ones_in(q.seq, 0, component_sequences, 5);

    const string name = std::filesystem::path(descfile).stem().string();
    for (const vector<Comp>& v : vresults) {
        bool unprobable = false;
        if (!q.allowed_basepair(v[0].first, v.back().second)) unprobable = true;
        for (size_t j = 0; j + 1 < v.size(); j++)
            if (!q.allowed_basepair(v[j].second, v[j + 1].first)) unprobable = true;
        if (!unprobable) sites.push_back(Site{name, v, {}});
    }
}

// ---------------------------------------------------------------------------------------
int main(int argc, char** argv)
{
    if (argc >= 5 && string(argv[1]) == "fno") {   // biorseo_port fno <delay> <seq> <component>...
        vector<string> comps(argv + 4, argv + argc);
        for (const vector<Comp>& v : find_next_ones_in(argv[3], 0, comps, (uint)std::atoi(argv[2]))) {
            string line;
            for (size_t j = 0; j < v.size(); j++) line += (j ? " " : "") + std::to_string(v[j].first) + "-" + std::to_string(v[j].second);
            std::puts(line.c_str());
        }
        return 0;
    }
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s jobs|ctor|index <source> <library> <seqfile> <maskfile>\n", argv[0]);
        return 2;
    }
    const string source = argv[2], lib = argv[3];
    Query        q;
    {
        std::ifstream     f(argv[4], std::ios::binary);
        std::stringstream ss;
        ss << f.rdbuf();
        q.seq = ss.str();
        if (!q.seq.empty() && q.seq.back() == '\n') q.seq.pop_back();
        q.n = q.seq.size();
        q.W = (q.n + 31) / 32;
        std::ifstream m(argv[5], std::ios::binary);
        q.mask.resize(q.n * q.W);
        m.read(reinterpret_cast<char*>(q.mask.data()), (std::streamsize)(q.mask.size() * 4));
        if ((size_t)m.gcount() != q.mask.size() * 4) { std::fprintf(stderr, "short mask file\n"); return 2; }
    }
    if (q.n < 6) undefined_in_reference("sequences shorter than 6 (index_of_yuv_ has n - 6 rows)");

    vector<Site> sites;
    if (source == "jsonfolder") {
        for (const auto& kv : read_json_library(lib)) {           // cppsrc/MOIP.cpp:262-288
            if (!kv.second.has_seq || !kv.second.has_ss) undefined_in_reference("module without string fields");
            if (is_valid_JSON(kv.second)) continue;
            allowed_motifs_from_json(q, kv.first, kv.second, sites);
        }
    } else {
        vector<string> files;
        for (auto& e : std::filesystem::recursive_directory_iterator(lib))
            if (e.is_regular_file()) files.push_back(e.path().string());
        std::sort(files.begin(), files.end());
        for (const string& p : files) {
            if (source == "rinfolder") {                          // cppsrc/MOIP.cpp:213-234
                if (is_valid_RIN(p)) continue;
                allowed_motifs_from_rin(q, p, sites);
            } else {                                              // cppsrc/MOIP.cpp:161-188
                if (is_valid_DESC(p)) continue;
                if (!is_desc_insertible(p, q.seq)) continue;
                allowed_motifs_from_desc(q, p, sites);
            }
        }
    }

    if (string(argv[1]) == "index") {
        // The loops of the reference that turn insertion_sites_ into variables and constraints, restated without the
        // solver objects and timed: cppsrc/MOIP.cpp:307-328 (one C_{x,i,p} per component), :507-566 (c3: the y terms of
        // every variable), :569-586 (c4: per nucleotide, the variables covering it). Prints the sizes (to be compared
        // with the device-built index) and the seconds the host loops took.
        const bool desc = source == "descfolder";
        const auto t0   = std::chrono::steady_clock::now();
        vector<size_t>         index_of_first_components;
        vector<vector<size_t>> index_of_Cxip;
        size_t                 i = 0;
        for (const Site& m : sites) {
            index_of_first_components.push_back(i);
            index_of_Cxip.push_back(vector<size_t>());
            for (size_t j = 0; j < m.comp.size(); j++) index_of_Cxip.back().push_back(i++);
        }
        unsigned long long c3_terms = 0, c3_constraints = 0, c4_entries = 0, c4_constraints = 0, checksum = 0;
        for (size_t x = 0; x < sites.size(); x++) {
            const Site& m = sites[x];
            for (size_t j = 0; j < m.comp.size(); j++) {
                const Comp& c = m.comp[j];
                uint count = 0;
                if (!desc) {
                    for (uint u = c.first; u <= c.second; u++)
                        for (uint v = 0; v < q.n; v++)
                            if (q.allowed_basepair(u, v)) {
                                bool is_link = false;
                                for (const Link& l : m.links)
                                    if ((u == l.a && v == l.b) || (u == l.b && v == l.a)) { is_link = true; break; }
                                if (!is_link) { count++; checksum += (unsigned long long)u * 31 + v; }
                            }
                } else {
                    for (uint u = c.first + 1; u < c.second; u++)
                        for (uint v = 0; v < q.n; v++)
                            if (q.allowed_basepair(u, v)) { count++; checksum += (unsigned long long)u * 31 + v; }
                }
                c3_terms += count;
                if (count > 0) c3_constraints++;
            }
        }
        for (uint u = 0; u < q.n; u++) {
            uint nterms = 0;
            for (size_t x = 0; x < sites.size(); x++)
                for (size_t j = 0; j < sites[x].comp.size(); j++)
                    if (u >= sites[x].comp[j].first && u <= sites[x].comp[j].second) { nterms++; checksum += index_of_Cxip[x][j]; }
            c4_entries += nterms;
            if (nterms > 1) c4_constraints++;
        }
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("{\"sites\": %zu, \"variables\": %zu, \"c3_terms\": %llu, \"c3_constraints\": %llu, \"c4_entries\": %llu, "
                    "\"c4_constraints\": %llu, \"checksum\": %llu, \"seconds\": %.6f}\n",
                    sites.size(), i, c3_terms, c3_constraints, c4_entries, c4_constraints, checksum, seconds);
        return 0;
    }

    string out;
    for (const Site& s : sites) {
        out += s.identifier + "\t";
        for (size_t j = 0; j < s.comp.size(); j++) out += (j ? " " : "") + std::to_string(s.comp[j].first) + "-" + std::to_string(s.comp[j].second);
        out += "\t";
        for (size_t j = 0; j < s.links.size(); j++) out += (j ? " " : "") + std::to_string(s.links[j].a) + "-" + std::to_string(s.links[j].b);
        out += "\n";
    }
    std::fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}
