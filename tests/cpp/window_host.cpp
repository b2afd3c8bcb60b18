// CPU emulation of the window pass (TEST infrastructure, not product code, never shipped in the library).
//
// The warp-cooperative enumerator of window_kernel lives in a host/device header (biorseo_b200/csrc/bss_window_walk.h).
// This harness compiles THAT source with g++ and drives it the way the kernel does — unit descriptors from the
// product's own builder (bss_blob.cpp), match masks M and thinned sets T computed naively from the pattern
// bytes, the banded mask re-laid like bandmask_kernel does, the 32 lanes of the warp emulated by loops between the
// walk's phases, count walk then emit walk — so that the CPU test-suite can compare the walk with the
// flat-table interpreter on thousands of random units without a GPU. The kernel's warp-level parts (occurrence
// table, match programs, thinning) are not exercised here; they are covered by the GPU tests.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "biorseo_b200.h"
#include "bss_layout.h"
#include "bss_window_walk.h"

namespace {

struct UnitView {
    const uint32_t* blob;
    int             C, delay;
};

int band_words(int band) { return band < 0 ? 0 : ((31 + band) >> 5) + 1; }

// bit q of M[c] = pattern c matches at q (q in [0, n]; '.' matches any byte)
void match_masks(const bss_table* t, uint32_t unit, const uint8_t* seq, int n, int W, std::vector<uint32_t>& M)
{
    const uint32_t c0 = t->unit_comp_begin[unit], c1 = t->unit_comp_begin[unit + 1];
    M.assign((size_t)(c1 - c0) * W, 0u);
    for (uint32_t c = c0; c < c1; c++) {
        const uint32_t pb = t->comp_pat_begin[c];
        const int      k  = (int)(t->comp_pat_begin[c + 1] - pb);
        for (int q = 0; q + k <= n; q++) {
            bool ok = true;
            for (int j = 0; ok && j < k; j++) ok = t->pattern[pb + j] == '.' || t->pattern[pb + j] == seq[q + j];
            if (ok) M[(size_t)(c - c0) * W + (q >> 5)] |= 1u << (q & 31);
        }
    }
}

void summaries(const std::vector<uint32_t>& B, int C, int W, int WS, std::vector<uint32_t>& S)
{
    S.assign((size_t)C * WS, 0u);
    for (int c = 0; c < C; c++)
        for (int w = 0; w < W; w++)
            if (B[(size_t)c * W + w]) S[(size_t)c * WS + (w >> 5)] |= 1u << (w & 31);
}

// T[c] = leftmost non-overlapping matches from position 0 (what sregex_iterator yields on the whole sequence);
// returns the mask of components for which T differs from M
uint32_t thin(const std::vector<uint32_t>& M, const UnitView& uv, int W, int n, std::vector<uint32_t>& T)
{
    T.assign(M.size(), 0u);
    uint32_t ovl = 0;
    for (int c = 0; c < uv.C; c++) {
        const int k = std::max(1, (int)(uv.blob[4 + 2 * c] & 0xFFu));
        int       next = 0;
        for (int q = 0; q <= n; q++) {
            if (!((M[(size_t)c * W + (q >> 5)] >> (q & 31)) & 1u)) continue;
            if (q >= next) {
                T[(size_t)c * W + (q >> 5)] |= 1u << (q & 31);
                next = q + k;
            } else {
                ovl |= 1u << c;
            }
        }
    }
    return ovl;
}

bool gate_open(const bss_table* t, uint32_t gate, const uint32_t* gblob, const uint8_t* seq, int n, int W)
{
    std::vector<uint32_t> M;
    match_masks(t, gate, seq, n, W, M);
    const int C = (int)(gblob[1] & 0xFFu), delay = (int)(gblob[1] >> 16);
    int       at = 0;
    for (int c = 0; c < C; c++) {
        int q = -1;
        for (int p = std::max(at, 0); p <= n; p++)
            if ((M[(size_t)c * W + (p >> 5)] >> (p & 31)) & 1u) { q = p; break; }
        if (q < 0) return false;
        at = q + (int)(gblob[4 + 2 * c] & 0xFFu) - 1 + delay;
    }
    return true;
}

}   // namespace

extern "C" {

// Searches every window-walk unit of the table on one sequence. *records (malloc'ed, free with window_host_free)
// receives the record stream of those units in canonical order, unit_window[u] = 1 for the units that were
// searched (the others are not window-walk units, or the band is too wide: *band_out tells).
int window_host_search(const bss_table* t, const char* seq_bytes, uint32_t n_u, const uint32_t* mask, uint32_t** records, uint64_t* n_words,
                       uint8_t* unit_window, int* band_out)
{
    bss::BuiltLibrary lib;
    std::string       why;
    if (bss::build_library(t, lib, why) != 0) return -1;
    const int      n   = (int)n_u;
    const uint8_t* seq = reinterpret_cast<const uint8_t*>(seq_bytes);
    const int      Wm = (n + 31) >> 5, W = (n + 32) / 32, WS = (W + 31) / 32, Wq = (n + 32) >> 5;
    // band (band_kernel) and banded mask (bandmask_kernel)
    int band = -1;
    for (int u = 0; u < n; u++)
        for (int v = u + 1; v < n; v++)
            if ((mask[(size_t)u * Wm + (v >> 5)] >> (v & 31)) & 1u) band = std::max(band, v - u);
    *band_out = band;
    std::vector<uint32_t> out;
    const char*           limit_env = std::getenv("WINDOW_HOST_BAND_LIMIT");   // (the kernel's limit is its shared-memory budget; tests may set one)
    const bool            narrow = band <= (limit_env ? std::atoi(limit_env) : 1 << 20);
    const int             K = narrow ? band_words(band) : 0;
    std::vector<uint32_t> bm((size_t)std::max(1, n * K), 0u);
    if (narrow)
        for (int u = 0; u < n; u++)
            for (int j = 0; j < K; j++) {
                const int w = (u >> 5) + j;
                uint32_t  x = w < Wm ? mask[(size_t)u * Wm + w] : 0u;
                if (j == 0) x &= (u & 31) == 31 ? 0u : (0xFFFFFFFFu << ((u & 31) + 1));
                if (w == Wm - 1 && (n & 31)) x &= (1u << (n & 31)) - 1u;
                bm[(size_t)u * K + j] = x;
            }
    // the walk's scratch, as the kernel carves it per warp (bss_window.cuh): level parameters, cursors, level queues, stash.
    // It starts from a poison value taken from the environment: the walk must never read an entry it has not written for
    // the current item, so the result may not depend on it. WINDOW_HOST_QCAP shrinks the queues to force the "queue full,
    // descend and come back" path on small inputs.
    const char*           poison_env = std::getenv("WINDOW_HOST_POISON");
    const uint32_t        poison = poison_env ? (uint32_t)std::strtoul(poison_env, nullptr, 0) : 0u;
    const char*           qcap_env = std::getenv("WINDOW_HOST_QCAP");
    const int             widest = std::max(band, 0);
    const int             qcap = qcap_env ? std::max(std::atoi(qcap_env), std::max(32, widest + 1)) : std::max(64, (widest + 1 + 31) & ~31);
    const int             qlevels = std::max<int>(1, (int)lib.cmax - 1);
    std::vector<int>      lp((size_t)bss::win::kLevelWords * lib.cmax);
    std::vector<uint32_t> ctl((size_t)2 * lib.cmax), Q((size_t)qlevels * qcap), stash((size_t)32 * std::max(1, K));
    for (uint32_t u = 0; u < t->n_units; u++) {
        unit_window[u] = 0;
        if (!narrow || !(lib.unit_flags[u] & bss::kUnitWindow)) continue;
        unit_window[u]       = 1;
        const uint32_t* blob = lib.blob.data() + lib.unit_off[u];
        UnitView        uv{blob, (int)(blob[1] & 0xFFu), (int)(blob[1] >> 16)};
        if (blob[2] != bss::kNoGate && !gate_open(t, blob[2], lib.blob.data() + lib.unit_off[blob[2]], seq, n, W)) continue;
        std::vector<uint32_t> M, T, SM, ST;
        match_masks(t, u, seq, n, W, M);
        bool every = true;   // match_components ends the item when a component has no match at all
        for (int c = 0; c < uv.C; c++) {
            bool any = false;
            for (int w = 0; w < W; w++) any = any || M[(size_t)c * W + w];
            every = every && any;
        }
        if (!every) continue;
        const uint32_t ovl = thin(M, uv, W, n, T);
        summaries(M, uv.C, W, WS, SM);
        bss::win::Walk cx;
        std::vector<const uint32_t*> rows;
        for (int c = 0; c < uv.C; c++) {
            rows.push_back(M.data() + (size_t)c * W);
            rows.push_back(T.data() + (size_t)c * W);
            rows.push_back(SM.data() + (size_t)c * WS);
        }
        cx.blob = blob; cx.rows = rows.data(); cx.bm = bm.data();
        cx.lp = lp.data(); cx.ctl = ctl.data(); cx.Q = Q.data(); cx.stash = stash.data(); cx.qcap = qcap;
        cx.ovl = ovl; cx.W = W; cx.WS = WS; cx.n = n; cx.K = K; cx.C = uv.C; cx.delay = uv.delay; cx.band = band;
        std::fill(lp.begin(), lp.end(), (int)poison);
        std::fill(ctl.begin(), ctl.end(), poison);
        std::fill(Q.begin(), Q.end(), poison);
        std::fill(stash.begin(), stash.end(), poison);
        const uint64_t total = bss::win::walk<false>(cx, Wq, blob[0], nullptr);
        const size_t   R = (size_t)uv.C + 1, base = out.size();
        out.resize(base + total * R, 0xDEADBEEFu);
        std::fill(lp.begin(), lp.end(), (int)~poison);
        std::fill(ctl.begin(), ctl.end(), ~poison);
        std::fill(Q.begin(), Q.end(), ~poison);
        std::fill(stash.begin(), stash.end(), ~poison);
        const uint64_t wrote = bss::win::walk<true>(cx, Wq, blob[0], out.data() + base);
        if (wrote != total) return -2;
    }
    *n_words = out.size();
    *records = static_cast<uint32_t*>(std::malloc(std::max<size_t>(4, out.size() * 4)));
    if (!*records) return -3;
    if (!out.empty()) std::memcpy(*records, out.data(), out.size() * 4);
    return 0;
}

void window_host_free(uint32_t* p) { std::free(p); }

}   // extern "C"
