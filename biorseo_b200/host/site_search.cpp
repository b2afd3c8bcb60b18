#include "site_search.h"

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

InsertionSiteSearch::~InsertionSiteSearch()
{
    if (device_library_) bss_library_destroy(device_library_);
}

bool InsertionSiteSearch::open(const std::string& source, const std::string& source_path, int device, std::string& err)
{
    LibraryKind kind;
    if (source == "jsonfolder") kind = KIND_JSON;
    else if (source == "rinfolder") kind = KIND_RIN;
    else if (source == "descfolder") kind = KIND_DESC;
    else { err = "Unknown module source."; return false; }
    if (!library_.load(kind, source_path, err)) return false;
    Motif::delay = library_.delay();
    if (device_library_) { bss_library_destroy(device_library_); device_library_ = nullptr; }
    const bss_table table = library_.table();
    if (bss_library_create(&table, device, &device_library_) != BSS_OK) { err = bss_last_error(); return false; }
    return true;
}

bool InsertionSiteSearch::run(const std::string& seq, const std::vector<uint32_t>& pair_mask, std::vector<Motif>& sites, std::string& err)
{
    if (!device_library_) { err = "no library open"; return false; }
    const size_t n = seq.size();
    if (pair_mask.size() != n * ((n + 31) / 32)) { err = "pair mask has the wrong size"; return false; }
    bss_sites* found = nullptr;
    if (bss_search(device_library_, seq.data(), (uint32_t)n, pair_mask.data(), &found) != BSS_OK) { err = bss_last_error(); return false; }
    const uint32_t* rec = bss_sites_records(found);
    const uint64_t  end = bss_sites_words(found);
    sites.reserve(sites.size() + bss_sites_count(found));
    for (uint64_t at = 0; at < end; at += library_.record_words(rec[at])) sites.push_back(library_.rebuild(rec + at));
    bss_sites_free(found);
    return true;
}

}   // namespace bsh
