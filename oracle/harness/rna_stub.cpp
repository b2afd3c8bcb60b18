// Oracle harness (test infrastructure, NOT product code).
// Stands in for /root/reference/cppsrc/rna.cpp, which cannot be built here (ViennaRNA is
// absent). It keeps the reference's rna.h unmodified and fills pij_ from a caller-supplied
// pair mask instead of vrna_pfl_fold (cppsrc/rna.cpp:49-61): pij(u,v) = 1 where the mask
// bit (u,v), u < v, is set and 0 elsewhere, so `pij > theta` (cppsrc/MOIP.cpp:80) reproduces
// the mask for any theta in [0,1).
#include "rna.h"
#include <cstdint>
#include <iostream>

// Set by the harness before an RNA is constructed: row-major n x ceil(n/32) u32 words.
const uint32_t* g_oracle_mask_words = nullptr;
size_t          g_oracle_mask_n     = 0;
// Or a full probability matrix (row-major n x n floats), for the pair-mask producer tests.
const float*    g_oracle_pij        = nullptr;

RNA::RNA(void) {}

RNA::RNA(std::string name, std::string seq, bool verbose)
: verbose_{verbose}, name_(name), seq_(seq), n_(seq.size()), pij_(Eigen::MatrixXf::Zero(seq.size(), seq.size()))
{
    if (g_oracle_pij != nullptr && g_oracle_mask_n == n_) {
        for (size_t u = 0; u < n_; u++)
            for (size_t v = 0; v < n_; v++) pij_(u, v) = g_oracle_pij[u * n_ + v];
        return;
    }
    if (g_oracle_mask_words == nullptr || g_oracle_mask_n != n_) return;
    const size_t W = (n_ + 31) / 32;
    for (size_t u = 0; u < n_; u++)
        for (size_t v = u + 1; v < n_; v++)
            if ((g_oracle_mask_words[u * W + (v >> 5)] >> (v & 31)) & 1u) pij_(u, v) = 1.0f;
}

void RNA::print_basepair_p_matrix(float) const {}

base_t RNA::base_type(char x) const
{
    switch (x) {
    case 'A': case 'a': return BASE_A;
    case 'C': case 'c': return BASE_C;
    case 'G': case 'g': return BASE_G;
    case 'U': case 'u': case 'T': case 't': return BASE_U;
    default: return BASE_N;
    }
}
