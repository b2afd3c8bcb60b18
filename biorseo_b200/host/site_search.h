// Drop-in for the library branches of the reference's MOIP constructor
// (cppsrc/MOIP.cpp:146-296): "read the query, the allowed_basepair predicate, the library
// path and Motif::delay; fill insertion_sites_". The module library is parsed once, compiled
// into a flat device table, and every (module, placement) is searched on the GPU through
// the C ABI in include/biorseo_b200.h. See INTEGRATION.md for the patch to MOIP.cpp.
#ifndef BIORSEO_B200_SITE_SEARCH_H_
#define BIORSEO_B200_SITE_SEARCH_H_

#include <functional>
#include <string>
#include <vector>

#include "Motif.h"
#include "fasta.h"
#include "module_library.h"

namespace bsh {

// Packs allowed(u, v), u < v, into n rows of ceil(n/32) words (bit v of row u). The
// predicate is the reference's MOIP::allowed_basepair (cppsrc/MOIP.cpp:896-910).
std::vector<uint32_t> build_pair_mask(size_t n, const std::function<bool(size_t, size_t)>& allowed);

// What the MOIP constructor derives from insertion_sites_ with host loops (cppsrc/MOIP.cpp:302-328 C_{x,i,p}
// variables, :507-566 c3, :569-586 c4, :74-90 / :472-487 the pair side), built on the device behind the search
// (bss_site_index_build, include/biorseo_b200.h). Site x of the returned list owns the variables
// [var_begin[x], var_begin[x+1]), i.e. index_of_first_components[x] = var_begin[x] and
// index_of_Cxip_[x][j] = var_begin[x] + j.
struct VariableIndex {
    std::vector<uint32_t> var_begin, var_first, var_second, var_c3_terms;   // variables table
    std::vector<uint32_t> overlap_begin, overlap_vars;                       // per nucleotide: variables covering it (c4)
    std::vector<uint32_t> pair_begin, pair_partner, row_first_pair;          // per nucleotide: allowed partners; y numbering
};

class InsertionSiteSearch
{
  public:
    InsertionSiteSearch() {}
    ~InsertionSiteSearch();
    InsertionSiteSearch(const InsertionSiteSearch&) = delete;
    InsertionSiteSearch& operator=(const InsertionSiteSearch&) = delete;

    // `source` is the reference's "jsonfolder" | "rinfolder" | "descfolder"
    // (cppsrc/biorseo.cpp:150-165). Also sets Motif::delay like the reference's main().
    bool open(const std::string& source, const std::string& source_path, int device, std::string& err);
    // The same over n_gpus GPUs of this box (devices 0 .. n_gpus - 1) in ONE process: the library is cut into contiguous
    // module ranges balanced by search cost, one per GPU (bss_sharded_create) — the reference's Pool of one job per module
    // (cppsrc/MOIP.cpp:246-293) spread over the GPUs. run / run_batch then return the very list the one-GPU search returns.
    // typical_len: expected query length, for the cost estimate (0: 1000). The variable index (run with `index`) needs
    // the records on one device and is not available in this mode.
    bool open_sharded(const std::string& source, const std::string& source_path, unsigned n_gpus, unsigned typical_len, std::string& err);

    // Appends every insertion site of the library on `seq` to `sites`, in canonical
    // (module, placement) order. There is no CPU fallback: fails when no GPU is usable.
    bool run(const std::string& seq, const std::vector<uint32_t>& pair_mask, std::vector<Motif>& sites, std::string& err)
    {
        return run(seq, pair_mask, sites, nullptr, err);
    }
    // The same, also filling the variable index of the sites found (`sites` must come in empty for the site numbers to line up).
    bool run(const std::string& seq, const std::vector<uint32_t>& pair_mask, std::vector<Motif>& sites, VariableIndex* index, std::string& err);
    bool run(const std::string& seq, const std::function<bool(size_t, size_t)>& allowed, std::vector<Motif>& sites, std::string& err)
    {
        return run(seq, build_pair_mask(seq.size(), allowed), sites, err);
    }

    // Batch mode (BASELINE config "10k sequences x JSON library"): every sequence of `seqs` (e.g. the records of a
    // multi-record FASTA file, fasta.h) searched in one call; sites[q] receives the sites of sequence q in canonical
    // order. masks[q] is the packed allowed_basepair mask of sequence q (build_pair_mask).
    bool run_batch(const std::vector<std::string>& seqs, const std::vector<std::vector<uint32_t>>& masks, std::vector<std::vector<Motif>>& sites,
                   std::string& err);

    // The `--pre-placed` CSV source (cppsrc/MOIP.cpp:114-145, cppsrc/Motif.cpp:17-47): motifs already
    // placed by BayesPairing, one per line after a header ("id,score,start,end,start,end,..."); no
    // search, only the closing-pair filter (first nucleotide of the first component with the last of the
    // last one, and every component's end with the next one's start). Host-only: needs no GPU.
    // Lines on which the reference's behaviour is undefined (non-numeric fields, no component: it reads
    // comp[0] of an empty vector; jar3d lines: it calls stoi("True")) are skipped and reported in `skipped`.
    static bool run_csv(const std::string& path, const std::function<bool(size_t, size_t)>& allowed, std::vector<Motif>& sites,
                        std::vector<std::string>& skipped, std::string& err);

    const ModuleLibrary& library() const { return library_; }

  private:
    ModuleLibrary library_;
    bss_library*  device_library_ = nullptr;
    bss_sharded*  sharded_ = nullptr;     // set instead of device_library_ by open_sharded
    bool load(const std::string& source, const std::string& source_path, std::string& err);
    void release();
};

}   // namespace bsh
#endif
