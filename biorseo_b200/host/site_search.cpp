#include "site_search.h"

#include <cstdlib>
#include <fstream>

namespace bsh {

std::vector<uint32_t> build_pair_mask(size_t n, const std::function<bool(size_t, size_t)>& allowed)
{
    const size_t          W = (n + 31) / 32;
    std::vector<uint32_t> mask(n * W, 0u);
    for (size_t u = 0; u < n; u++)
        for (size_t v = u + 1; v < n; v++)
            if (allowed(u, v)) mask[u * W + (v >> 5)] |= 1u << (v & 31);
    return mask;
}

InsertionSiteSearch::~InsertionSiteSearch() { release(); }

void InsertionSiteSearch::release()
{
    if (device_library_) bss_library_destroy(device_library_);
    if (sharded_) bss_sharded_destroy(sharded_);
    device_library_ = nullptr;
    sharded_        = nullptr;
}

bool InsertionSiteSearch::load(const std::string& source, const std::string& source_path, std::string& err)
{
    LibraryKind kind;
    if (source == "jsonfolder") kind = KIND_JSON;
    else if (source == "rinfolder") kind = KIND_RIN;
    else if (source == "descfolder") kind = KIND_DESC;
    else { err = "Unknown module source."; return false; }
    if (!library_.load(kind, source_path, err)) return false;
    Motif::delay = library_.delay();
    release();
    return true;
}

bool InsertionSiteSearch::open(const std::string& source, const std::string& source_path, int device, std::string& err)
{
    if (!load(source, source_path, err)) return false;
    const bss_table table = library_.table();
    if (bss_library_create(&table, device, &device_library_) != BSS_OK) { err = bss_last_error(); return false; }
    return true;
}

bool InsertionSiteSearch::open_sharded(const std::string& source, const std::string& source_path, unsigned n_gpus, unsigned typical_len, std::string& err)
{
    if (!load(source, source_path, err)) return false;
    const bss_table table = library_.table();
    if (bss_sharded_create(&table, nullptr, n_gpus, typical_len, &sharded_) != BSS_OK) { err = bss_last_error(); return false; }
    return true;
}

bool InsertionSiteSearch::run(const std::string& seq, const std::vector<uint32_t>& pair_mask, std::vector<Motif>& sites, VariableIndex* index,
                              std::string& err)
{
    if (!device_library_ && !sharded_) { err = "no library open"; return false; }
    if (index && sharded_) { err = "the variable index needs the records on one device: open() the library on one GPU"; return false; }
    const size_t n = seq.size();
    if (pair_mask.size() != n * ((n + 31) / 32)) { err = "pair mask has the wrong size"; return false; }
    bss_sites*     found         = nullptr;
    const uint64_t seq_begin[2]  = {0, n};
    const uint64_t mask_begin[2] = {0, pair_mask.size()};
    const int      rc = sharded_ ? bss_search_sharded(sharded_, 1, seq.data(), seq_begin, pair_mask.data(), mask_begin, &found)
                                 : bss_search(device_library_, seq.data(), (uint32_t)n, pair_mask.data(), &found);
    if (rc != BSS_OK) { err = bss_last_error(); return false; }
    const uint32_t* rec = bss_sites_records(found);
    const uint64_t  end = bss_sites_words(found);
    sites.reserve(sites.size() + bss_sites_count(found));
    for (uint64_t at = 0; at < end; at += library_.record_words(rec[at])) sites.push_back(library_.rebuild(rec + at));
    if (index) {
        bss_site_index* x    = nullptr;
        const int       rule = library_.kind() == KIND_DESC ? BSS_C3_INTERIOR : BSS_C3_COMPONENT_LESS_LINKS;
        if (bss_site_index_build(device_library_, nullptr, 0, bss_sites_device_records(found), rule, 1, &x) != BSS_OK) {
            err = bss_last_error();
            bss_sites_free(found);
            return false;
        }
        auto take = [&](int section, std::vector<uint32_t>& out) {
            const uint32_t* p = bss_site_index_host(x, section);
            out.assign(p, p + bss_site_index_count(x, section));
        };
        take(BSS_IDX_VAR_BEGIN, index->var_begin);
        take(BSS_IDX_VAR_FIRST, index->var_first);
        take(BSS_IDX_VAR_SECOND, index->var_second);
        take(BSS_IDX_VAR_C3_TERMS, index->var_c3_terms);
        take(BSS_IDX_OVERLAP_BEGIN, index->overlap_begin);
        take(BSS_IDX_OVERLAP_VARS, index->overlap_vars);
        take(BSS_IDX_PAIR_BEGIN, index->pair_begin);
        take(BSS_IDX_PAIR_PARTNER, index->pair_partner);
        take(BSS_IDX_ROW_FIRST_PAIR, index->row_first_pair);
        bss_site_index_free(x);
    }
    bss_sites_free(found);
    return true;
}

bool InsertionSiteSearch::run_batch(const std::vector<std::string>& seqs, const std::vector<std::vector<uint32_t>>& masks,
                                    std::vector<std::vector<Motif>>& sites, std::string& err)
{
    if (!device_library_ && !sharded_) { err = "no library open"; return false; }
    if (masks.size() != seqs.size()) { err = "one pair mask per sequence is needed"; return false; }
    std::string           all;
    std::vector<uint32_t> all_masks;
    std::vector<uint64_t> seq_begin(1, 0), mask_begin(1, 0);
    for (size_t q = 0; q < seqs.size(); q++) {
        const size_t n = seqs[q].size();
        if (masks[q].size() != n * ((n + 31) / 32)) { err = "pair mask " + std::to_string(q) + " has the wrong size"; return false; }
        all += seqs[q];
        all_masks.insert(all_masks.end(), masks[q].begin(), masks[q].end());
        seq_begin.push_back(all.size());
        mask_begin.push_back(all_masks.size());
    }
    all_masks.push_back(0u);   // never read: keeps data() valid for an all-empty batch
    bss_sites* found = nullptr;
    const int rc = sharded_ ? bss_search_sharded(sharded_, (uint32_t)seqs.size(), all.c_str(), seq_begin.data(), all_masks.data(), mask_begin.data(), &found)
                            : bss_search_batch(device_library_, (uint32_t)seqs.size(), all.c_str(), seq_begin.data(), all_masks.data(), mask_begin.data(), &found);
    if (rc != BSS_OK) {
        err = bss_last_error();
        return false;
    }
    const uint32_t* rec   = bss_sites_records(found);
    const uint64_t* begin = bss_sites_seq_word_begin(found);
    sites.assign(seqs.size(), std::vector<Motif>());
    for (size_t q = 0; q < seqs.size(); q++)
        for (uint64_t at = begin[q]; at < begin[q + 1]; at += library_.record_words(rec[at])) sites[q].push_back(library_.rebuild(rec + at));
    bss_sites_free(found);
    return true;
}

namespace {

// std::stoi as the reference uses it: optional blanks and sign, digits, trailing garbage ignored;
// anything else throws there (= undefined for us).
bool stoi_like(const std::string& s, int& out)
{
    const char* p   = s.c_str();
    char*       end = nullptr;
    const long  v   = std::strtol(p, &end, 10);
    if (end == p || v < -2147483647L - 1 || v > 2147483647L) return false;
    out = (int)v;
    return true;
}

}   // namespace

bool InsertionSiteSearch::run_csv(const std::string& path, const std::function<bool(size_t, size_t)>& allowed, std::vector<Motif>& sites,
                                  std::vector<std::string>& skipped, std::string& err)
{
    std::ifstream in(path);
    if (!in) { err = "cannot open " + path; return false; }
    std::string line;
    std::getline(in, line);   // header
    while (std::getline(in, line)) {
        std::vector<std::string> tokens;   // boost::split(is_any_of(",")): no token compression, n commas -> n + 1 tokens
        size_t                   at = 0;
        for (;;) {
            const size_t comma = line.find(',', at);
            tokens.push_back(line.substr(at, comma == std::string::npos ? std::string::npos : comma - at));
            if (comma == std::string::npos) break;
            at = comma + 1;
        }
        int  score = 0;
        bool ok    = tokens.size() >= 2 && stoi_like(tokens[1], score) && line.find("True") == std::string::npos &&
                  line.find("False") == std::string::npos;
        std::vector<Component> comps;
        for (size_t i = 2; ok && i + 1 < tokens.size(); i += 2) {   // pairs (start, end); kept when start < end (cppsrc/Motif.cpp:36-44)
            int a = 0, b = 0;
            ok = stoi_like(tokens[i], a) && stoi_like(tokens[i + 1], b);
            if (ok && a < b) comps.push_back(Component(std::make_pair(a, b)));
        }
        if (!ok || comps.empty()) {
            skipped.push_back(line);
            continue;
        }
        if (!allowed(comps.front().pos.first, comps.back().pos.second)) continue;
        bool keep = true;
        for (size_t j = 0; keep && j + 1 < comps.size(); j++) keep = allowed(comps[j].pos.second, comps[j + 1].pos.first);
        if (!keep) continue;
        Motif m(comps, tokens[0], CSV);
        m.score_ = score;
        sites.push_back(m);
    }
    return true;
}

}   // namespace bsh
