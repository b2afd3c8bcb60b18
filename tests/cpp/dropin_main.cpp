// Test driver of the C++ drop-in (bsh::InsertionSiteSearch, biorseo_b200/host/site_search.h):
// does what the patched MOIP constructor of INTEGRATION.md does — open the library, run the
// search with an allowed_basepair predicate, walk the resulting vector<Motif> — and dumps the
// sites in the oracle's text format (identifier \t first-second ... \t a-b ...).
//
//   dropin <jsonfolder|rinfolder|descfolder> <library path> <seq.txt> <mask.bin> [index]
//
// With `index` the sites are followed by the variable index the patched constructor would use instead of its
// loops over insertion_sites_ (cppsrc/MOIP.cpp:302-328, 507-586): one "V x j first second c3terms" line per
// C_{x,i,p} variable and one "O u v v ..." line per nucleotide covered by more than one variable (the c4 rows).
//
//   dropin <source> <library path> <seq.txt> <mask.bin> constraints
//
// With `constraints` the program goes one step further down the MOIP constructor: from the sites and the variable
// index it builds every constraint of define_problem_constraints that involves an insertion variable — c3
// (cppsrc/MOIP.cpp:507-566), c4 (:569-586), c5 (:588-603), c6 (:605-675) — WITHOUT the reference's loops over
// insertion_sites_ x n, and prints them in the format of the oracle's `biorseo_ref model` ("S" site lines in variable
// order, "K <op> <rhs> <coef>*<name> ..." lines), so that the two dumps can be compared as multisets.
//
//   dropin <source> <library path> <file.fa> - batch
//
// Batch mode: every record of a (multi-record) FASTA file (bsh::load_fasta) searched in ONE call
// (bsh::InsertionSiteSearch::run_batch -> bss_search_batch); the pair mask of each record is the canonical one
// (AU / GC / GU pairs, 4 <= v - u < 100, u < n - 6). Output lines: record name \t site.
//
//   dropin <source> <library path> <seq.txt> <mask.bin> gpus=N        dropin <source> <library path> <file.fa> - batch-gpus=N
//
// The same searches with the library cut over N GPUs of the box in this one process
// (bsh::InsertionSiteSearch::open_sharded -> bss_sharded_create / bss_search_sharded): the output must not change.
//
// Exit codes: 0 ok, 3 search failed (e.g. no GPU: there is no CPU fallback), 2 usage / input.
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

#include "site_search.h"

static void print_site(const Motif& m)
{
    std::cout << m.get_identifier() << '\t';
    for (size_t j = 0; j < m.comp.size(); j++) std::cout << (j ? " " : "") << m.comp[j].pos.first << '-' << m.comp[j].pos.second;
    std::cout << '\t';
    for (size_t j = 0; j < m.links_.size(); j++) std::cout << (j ? " " : "") << m.links_[j].nts.first << '-' << m.links_[j].nts.second;
    std::cout << '\n';
}

// The constraints with an insertion variable, from the variable index (see the header comment). `allowed` is the
// caller's allowed_basepair predicate; it is only needed for the c6 rows, which test single pairs.
template <typename Allowed>
static void print_constraints(bool desc, size_t n, const Allowed& allowed, const std::vector<Motif>& sites, const bsh::VariableIndex& ix)
{
    char buf[96];
    auto c_name = [&](size_t x, size_t j) {
        const uint32_t v = ix.var_begin[x] + (uint32_t)j;
        std::snprintf(buf, sizeof buf, "C%zu,%zu[%d,%d]", x, j, (int)ix.var_first[v], (int)ix.var_second[v]);
        return std::string(buf);
    };
    auto y_name = [&](size_t u, size_t v) {
        std::snprintf(buf, sizeof buf, "y%zu,%zu", u < v ? u : v, u < v ? v : u);
        return std::string(buf);
    };
    auto num = [&](double c) {
        std::snprintf(buf, sizeof buf, "%g", c);
        return std::string(buf);
    };
    for (const Motif& m : sites) {
        std::cout << "S\t";
        print_site(m);
    }
    std::vector<uint32_t> site_of(ix.var_first.size());
    for (size_t x = 0; x + 1 < ix.var_begin.size(); x++)
        for (uint32_t v = ix.var_begin[x]; v < ix.var_begin[x + 1]; v++) site_of[v] = (uint32_t)x;
    // c3: k * C + the y of every allowed pair touching the component (json / rin: the site's links left out) <= k
    for (size_t x = 0; x < sites.size(); x++) {
        const Motif& m = sites[x];
        for (size_t j = 0; j < m.comp.size(); j++) {
            const uint32_t v = ix.var_begin[x] + (uint32_t)j;
            if (ix.var_c3_terms[v] == 0) continue;   // the index says: nothing to add, no scan over the sequence
            const double coef = desc ? (double)m.comp[j].k - 2.0 : (double)m.comp[j].k;
            std::string  line = "K < " + num(coef) + " " + num(coef) + "*" + c_name(x, j);
            const size_t u0 = desc ? m.comp[j].pos.first + 1 : m.comp[j].pos.first, u1 = desc ? m.comp[j].pos.second : m.comp[j].pos.second + 1;
            for (size_t u = u0; u < u1; u++)
                for (uint32_t i = ix.pair_begin[u]; i < ix.pair_begin[u + 1]; i++) {
                    const size_t w = ix.pair_partner[i];
                    bool         is_link = false;
                    if (!desc)
                        for (const Link& l : m.links_)
                            if ((u == l.nts.first && w == l.nts.second) || (u == l.nts.second && w == l.nts.first)) { is_link = true; break; }
                    if (!is_link) line += " 1*" + y_name(u, w);
                }
            std::cout << line << '\n';
        }
    }
    // c4: the variables covering a nucleotide, at most one of them
    for (size_t u = 0; u < n; u++) {
        if (ix.overlap_begin[u + 1] - ix.overlap_begin[u] < 2) continue;
        std::string line = "K < 1";
        for (uint32_t i = ix.overlap_begin[u]; i < ix.overlap_begin[u + 1]; i++) {
            const uint32_t v = ix.overlap_vars[i], x = site_of[v];
            line += " 1*" + c_name(x, v - ix.var_begin[x]);
        }
        std::cout << line << '\n';
    }
    for (size_t x = 0; x < sites.size(); x++) {
        const Motif& m = sites[x];
        // c5: all components of a site or none
        if (m.comp.size() > 1) {
            std::string line = "K = 0";
            for (size_t j = 1; j < m.comp.size(); j++) line += " 1*" + c_name(x, j);
            line += " " + num(-(double)(m.comp.size() - 1)) + "*" + c_name(x, 0);
            std::cout << line << '\n';
        }
        // c6: the base pairs a site imposes
        if (!desc) {
            for (size_t j = 0; j < m.comp.size(); j++) {
                unsigned    ax = 0;
                std::string terms;
                for (const Link& l : m.links_) {
                    if (m.comp[j].pos.first <= l.nts.first && l.nts.first <= m.comp[j].pos.second) {
                        ax++;
                        if (!allowed(l.nts.first, l.nts.second)) break;
                        terms += " 1*" + y_name(l.nts.first, l.nts.second);
                    }
                }
                if (ax) std::cout << "K > 0" << terms << " " << num(-(double)ax) << "*" << c_name(x, j) << '\n';
            }
        } else {
            std::string line = "K < 0 1*" + c_name(x, 0);
            if (allowed(m.comp[0].pos.first, m.comp.back().pos.second)) line += " -1*" + y_name(m.comp[0].pos.first, m.comp.back().pos.second);
            std::cout << line << '\n';
            for (size_t j = 0; j + 1 < m.comp.size(); j++) {
                line = "K < 0 1*" + c_name(x, j);
                if (allowed(m.comp[j].pos.second, m.comp[j + 1].pos.first)) line += " -1*" + y_name(m.comp[j].pos.second, m.comp[j + 1].pos.first);
                std::cout << line << '\n';
            }
        }
    }
}

static int run_batch_mode(char** argv)
{
    std::string                   err;
    std::vector<bsh::FastaRecord> records;
    if (!bsh::load_fasta(argv[3], records, err)) {
        std::cerr << "ERR: " << err << "\n";
        return 2;
    }
    std::vector<std::string>           seqs;
    std::vector<std::vector<uint32_t>> masks;
    for (const bsh::FastaRecord& r : records) {
        const std::string& s = r.seq;
        const size_t       n = s.size();
        auto canonical = [&](size_t u, size_t v) {
            if (v < u + 4 || v - u >= 100 || u + 6 >= n) return false;
            const char a = s[u], c = s[v];
            return (a == 'A' && c == 'U') || (a == 'U' && c == 'A') || (a == 'G' && c == 'C') || (a == 'C' && c == 'G') || (a == 'G' && c == 'U') ||
                   (a == 'U' && c == 'G');
        };
        seqs.push_back(s);
        masks.push_back(bsh::build_pair_mask(n, canonical));
    }
    bsh::InsertionSiteSearch search;
    const std::string mode = argv[5];
    const unsigned    gpus = mode.rfind("batch-gpus=", 0) == 0 ? (unsigned)std::stoul(mode.substr(11)) : 0u;
    if (!(gpus ? search.open_sharded(argv[1], argv[2], gpus, 120, err) : search.open(argv[1], argv[2], -1, err))) {
        std::cerr << "ERR: " << err << "\n";
        return 3;
    }
    std::vector<std::vector<Motif>> sites;
    if (!search.run_batch(seqs, masks, sites, err)) {
        std::cerr << "ERR: " << err << "\n";
        return 3;
    }
    size_t total = 0;
    for (size_t q = 0; q < sites.size(); q++) {
        total += sites[q].size();
        for (const Motif& m : sites[q]) {
            std::cout << records[q].name << '\t';
            print_site(m);
        }
    }
    std::cerr << total << " candidate motifs over " << records.size() << " sequences\n";
    return 0;
}

int main(int argc, char** argv)
{
    const bool with_constraints = argc == 6 && std::string(argv[5]) == "constraints";
    const bool with_index = (argc == 6 && std::string(argv[5]) == "index") || with_constraints;
    if (argc == 6 && std::string(argv[5]).rfind("batch", 0) == 0) return run_batch_mode(argv);
    const unsigned gpus = argc == 6 && std::string(argv[5]).rfind("gpus=", 0) == 0 ? (unsigned)std::stoul(std::string(argv[5]).substr(5)) : 0u;
    if (argc != 5 && !with_index && !gpus) {
        std::cerr << "usage: dropin <source> <library> <seq.txt> <mask.bin>\n";
        return 2;
    }
    std::ifstream sf(argv[3]);
    std::string   seq;
    std::getline(sf, seq);
    const size_t n = seq.size(), W = (n + 31) / 32;
    std::ifstream mf(argv[4], std::ios::binary);
    std::vector<uint32_t> mask(n * W, 0u);
    mf.read(reinterpret_cast<char*>(mask.data()), (std::streamsize)(mask.size() * 4));
    if (!sf || (size_t)mf.gcount() != mask.size() * 4) {
        std::cerr << "cannot read the query\n";
        return 2;
    }
    // the caller's predicate (MOIP::allowed_basepair, cppsrc/MOIP.cpp:896-910) — here backed by the mask file
    auto allowed = [&](size_t u, size_t v) {
        const size_t a = u < v ? u : v, b = u < v ? v : u;
        return a != b && b < n && ((mask[a * W + (b >> 5)] >> (b & 31)) & 1u) != 0;
    };

    std::string err;
    if (std::string(argv[1]) == "csvfile") {   // pre-placed motifs: host-only filter, no library, no GPU
        std::vector<Motif>       placed;
        std::vector<std::string> skipped;
        if (!bsh::InsertionSiteSearch::run_csv(argv[2], allowed, placed, skipped, err)) {
            std::cerr << "ERR: " << err << "\n";
            return 3;
        }
        for (const Motif& m : placed) {
            std::cout << m.get_identifier() << '\t';
            for (size_t j = 0; j < m.comp.size(); j++) std::cout << (j ? " " : "") << m.comp[j].pos.first << '-' << m.comp[j].pos.second;
            std::cout << "\t\n";
        }
        std::cerr << placed.size() << " pre-placed motifs kept, " << skipped.size() << " lines skipped\n";
        return 0;
    }
    bsh::InsertionSiteSearch search;
    if (!(gpus ? search.open_sharded(argv[1], argv[2], gpus, (unsigned)n, err) : search.open(argv[1], argv[2], -1, err))) {
        std::cerr << "ERR: " << err << "\n";
        return 3;
    }
    std::vector<Motif> insertion_sites_;
    bsh::VariableIndex index;
    if (!search.run(seq, bsh::build_pair_mask(n, allowed), insertion_sites_, with_index ? &index : nullptr, err)) {
        std::cerr << "ERR: " << err << "\n";
        return 3;
    }
    if (with_constraints) {
        print_constraints(std::string(argv[1]) == "descfolder", n, allowed, insertion_sites_, index);
        return 0;
    }
    for (const Motif& m : insertion_sites_) print_site(m);
    if (with_index) {
        for (size_t x = 0; x + 1 < index.var_begin.size(); x++)
            for (uint32_t v = index.var_begin[x]; v < index.var_begin[x + 1]; v++)
                std::cout << "V " << x << ' ' << v - index.var_begin[x] << ' ' << index.var_first[v] << ' ' << (int)index.var_second[v] << ' '
                          << index.var_c3_terms[v] << '\n';
        for (size_t u = 0; u + 1 < index.overlap_begin.size(); u++) {
            if (index.overlap_begin[u + 1] - index.overlap_begin[u] < 2) continue;
            std::cout << "O " << u;
            for (uint32_t i = index.overlap_begin[u]; i < index.overlap_begin[u + 1]; i++) std::cout << ' ' << index.overlap_vars[i];
            std::cout << '\n';
        }
    }
    std::cerr << insertion_sites_.size() << " candidate motifs on " << search.library().n_seen() << " (" << search.library().rejected().size()
              << " ignored motifs), Motif::delay = " << Motif::delay << "\n";
    return 0;
}
