// Window walk: the per-lane placement enumerator of window_kernel (bss_kernels.cu).
//
// What it replaces (reference, CPU): the recursion of find_next_ones_in (cppsrc/Motif.cpp:432-494)
// followed by the pair filters (cppsrc/MOIP.cpp:1072-1079, 1131-1137, 1194-1205), for units in
// which every component after the first is tied to an earlier one by a "row check" (kCheckRow: a
// base pair (u, v) with u on an earlier component and v on the component being placed — every JSON
// loop closing pair, every DESC closing pair, most RIN links).
//
// Idea: for a placed upper part, the candidates of the next component that satisfy such a check are
// exactly the bits of   chain set of the component  AND  (mask row u >> offset)  — a few word-wide
// ANDs instead of a list walk — and they live in the window (u, u + band], `band` being the widest
// pair the sequence's mask allows (the reference's masks come from a windowed fold, span <= 100).
// A lane owns a contiguous range of the first component's candidates and runs a depth-first walk
// whose every step is "next set bit of an AND of a handful of words": only partial placements whose
// pairs are all allowed are ever visited, so in the banded regime the walk ends after a few words
// per candidate. The chain semantics of the reference (leftmost non-overlapping matches of the
// suffix, sregex_iterator) are kept: for a component whose matches overlap on this sequence the
// candidates are a few "greedy" matches followed by the thinned set T (see thin_components).
//
// Plain C++ in host/device functions: the same source runs in the kernel and, lane by lane, in the
// CPU emulation the tests compare with the flat-table interpreter (tests/cpp/window_host.cpp).
#pragma once
#include <cstdint>

#include "bss_layout.h"

#if defined(__CUDACC__) && defined(BSS_WALK_NOINLINE)
#define BSS_HD __host__ __device__ __noinline__
#elif defined(__CUDACC__)
#define BSS_HD __host__ __device__ __forceinline__
#else
#define BSS_HD inline
#endif

// Optional index checks of the debug build (bss_kernels.cu defines it with -DBSS_DEBUG_ASSERT): a violated
// condition is counted and the access is skipped.
#ifndef BSS_WIN_CHECK
#define BSS_WIN_CHECK(cond, id) true
#endif

#if defined(__CUDACC__)
#define BSS_HD_NI __host__ __device__ __noinline__
#else
#define BSS_HD_NI inline
#endif
#ifdef BSS_NI_ROW
#define BSS_HD_ROW BSS_HD_NI
#else
#define BSS_HD_ROW BSS_HD
#endif
#ifdef BSS_NI_LEVEL
#define BSS_HD_LEVEL BSS_HD_NI
#else
#define BSS_HD_LEVEL BSS_HD
#endif
#ifdef BSS_NI_WALK
#define BSS_HD_WALK BSS_HD_NI
#else
#define BSS_HD_WALK BSS_HD
#endif
#ifdef BSS_NI_MISC
#define BSS_HD_MISC BSS_HD_NI
#else
#define BSS_HD_MISC BSS_HD
#endif

namespace bss {
namespace win {

#if defined(__CUDA_ARCH__)
BSS_HD int      ffs32(uint32_t x) { return __ffs((int)x); }
BSS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
#else
BSS_HD int      ffs32(uint32_t x) { return __builtin_ffs((int)x); }
BSS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (lo >> s) | (hi << (32u - s)) : lo; }
#endif

BSS_HD int imax(int a, int b) { return a > b ? a : b; }
BSS_HD int imin(int a, int b) { return a < b ? a : b; }
BSS_HD int anchor_of(uint32_t w) { return (int)(int8_t)(w & 0xFFu); }
BSS_HD int off_of(uint32_t w) { return (int)(int16_t)((w >> 8) & 0xFFFFu); }

struct Ctx {
    const uint32_t* blob;   // unit descriptor
    const uint32_t* M;      // [C][W] every match of every component
    const uint32_t* T;      // [C][W] thinned matches, valid for the components in `ovl`
    const uint32_t* SM;     // [C][WS] one-bit-per-word summaries of M
    const uint32_t* ST;     // [C][WS] ... of T
    const uint32_t* bm;     // banded pair mask of the sequence: row u = K words, the mask words (u >> 5) .. of row u
    uint32_t*       pos;    // this lane's column (stride cs): position | last position of the level's window << 16
    uint32_t*       bits;   // this lane's column: candidates left in the word of the current position
    uint32_t*       grd;    // this lane's column: 1 while the level's chain is still a greedy walk over all matches
    uint32_t        ovl;    // bit c: matches of component c overlap on this sequence
    int             W, WS, n, K, C, delay, band, cs;
};

// First set bit at or after position t (t >= 0) of a bitset with a one-bit-per-word summary, or -1.
BSS_HD_MISC int next_set(const uint32_t* B, const uint32_t* S, int W, int WS, int t)
{
    if (!BSS_WIN_CHECK(t >= 0, 18)) return -1;
    int w = t >> 5;
    if (w >= W) return -1;
    uint32_t x = B[w] & (0xFFFFFFFFu << (t & 31));
    if (x) return (w << 5) + ffs32(x) - 1;
    if (++w >= W) return -1;
    int      sw = w >> 5;
    uint32_t s  = S[sw] & (0xFFFFFFFFu << (w & 31));
    while (!s) {
        if (++sw >= WS) return -1;
        s = S[sw];
    }
    w = (sw << 5) + ffs32(s) - 1;
    if (!BSS_WIN_CHECK(w < W && B[w] != 0, 19)) return -1;
    return (w << 5) + ffs32(B[w]) - 1;
}

BSS_HD bool bit_of(const uint32_t* B, int q)
{
    if (!BSS_WIN_CHECK(q >= 0, 20)) return false;
    return (B[q >> 5] >> (q & 31)) & 1u;
}

BSS_HD int placed(const Ctx& cx, int level)
{
    if (!BSS_WIN_CHECK(level >= 0 && level < cx.C, 17)) return 0;
    return (int)(cx.pos[level * cx.cs] & 0xFFFFu);
}

// Word w (absolute index into the mask row) of row u of the pair mask; zero outside the band.
BSS_HD uint32_t row_word(const Ctx& cx, int u, int w)
{
    const unsigned j = (unsigned)(w - (u >> 5));
    if (!BSS_WIN_CHECK(u >= 0 && u < cx.n, 16)) return 0u;
    return j < (unsigned)cx.K ? cx.bm[u * cx.K + (int)j] : 0u;
}

BSS_HD bool pair_allowed(const Ctx& cx, int u, int v)
{
    const int a = imin(u, v), b = imax(u, v);
    if (a < 0 || b >= cx.n || a == b) return false;
    return (row_word(cx, a, b >> 5) >> (b & 31)) & 1u;
}

// The checks of level `level` that are not row checks, one candidate at a time (rare: pairs inside one
// strand, absolute positions, ends outside their component).
BSS_HD_MISC bool generic_pass(const Ctx& cx, int level)
{
    const uint32_t  range = cx.blob[5 + 2 * level];
    const int       cb = (int)(range & 0xFFu) + (int)(range >> 24), ce = (int)((range >> 8) & 0xFFu);   // (the row checks come first)
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16);
    for (int i = cb; i < ce; i++) {
        const uint32_t ea = chk[2 * i], eb = chk[2 * i + 1];
        const int aa = anchor_of(ea), ba = anchor_of(eb);
        const int u  = (aa < 0 ? 0 : placed(cx, aa)) + off_of(ea);
        const int v  = (ba < 0 ? 0 : placed(cx, ba)) + off_of(eb);
        if (!pair_allowed(cx, u, v)) return false;
    }
    return true;
}

// The row checks of a level, decoded once per step: u = the end on the earlier component (fixed while the level is
// scanned), ob = offset of the other end on the component being placed. The first two live in registers; units with
// more of them per level (rare) take the rest from the descriptor (row_filter_rest).
struct Rows {
    int n, u0, ob0, u1, ob1;
};

// Decodes the row checks of level c (the first `n` checks of the level). False when one of them cannot hold
// (its earlier end lies before the sequence).
BSS_HD bool decode_rows(const Ctx& cx, int c, Rows& r)
{
    const uint32_t  range = cx.blob[5 + 2 * c];
    const int       cb = (int)(range & 0xFFu);
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16) + 2 * cb;
    r.n  = (int)(range >> 24);
    r.u0 = r.u1 = 0;
    r.ob0 = r.ob1 = 0;
    if (r.n > 0) {
        const uint32_t ea = chk[0];
        r.u0  = placed(cx, anchor_of(ea)) + off_of(ea);
        r.ob0 = off_of(chk[1]);
        if (r.u0 < 0) return false;
    }
    if (r.n > 1) {
        const uint32_t ea = chk[2];
        r.u1  = placed(cx, anchor_of(ea)) + off_of(ea);
        r.ob1 = off_of(chk[3]);
        if (r.u1 < 0) return false;
    }
    for (int i = 2; i < r.n; i++)
        if (placed(cx, anchor_of(chk[2 * i])) + off_of(chk[2 * i]) < 0) return false;
    return true;
}

// bit q of the result = bit q + ob of row u of the pair mask, for the positions of word w
BSS_HD uint32_t row_bits(const Ctx& cx, int u, int ob, int w)
{
    const int ws = w + (ob >> 5);
    return funnel_r(row_word(cx, u, ws), row_word(cx, u, ws + 1), (uint32_t)ob & 31u);
}

// x = bits of word w of the chain set of component c; keeps the positions that satisfy every row check of the level.
BSS_HD_ROW uint32_t row_filter(const Ctx& cx, int c, const Rows& r, int w, uint32_t x)
{
    if (r.n > 0) x &= row_bits(cx, r.u0, r.ob0, w);
    if (r.n > 1 && x) x &= row_bits(cx, r.u1, r.ob1, w);
    if (r.n > 2) {
        const uint32_t* chk = cx.blob + (cx.blob[3] >> 16) + 2 * (int)(cx.blob[5 + 2 * c] & 0xFFu);
        for (int i = 2; i < r.n && x; i++) x &= row_bits(cx, placed(cx, anchor_of(chk[2 * i])) + off_of(chk[2 * i]), off_of(chk[2 * i + 1]), w);
    }
    return x;
}

BSS_HD_ROW bool row_pass(const Ctx& cx, int c, const Rows& r, int q)
{
    return (row_filter(cx, c, r, q >> 5, 1u << (q & 31)) >> (q & 31)) & 1u;
}

// Window of component c below the current placement of components 0 .. c-1: the candidates lie in [start, hi].
// False when it is empty. `r` = the decoded row checks of the level.
BSS_HD bool level_window(const Ctx& cx, int c, const Rows& r, int& start, int& hi)
{
    const int kprev = (int)(cx.blob[4 + 2 * (c - 1)] & 0xFFu), kc = (int)(cx.blob[4 + 2 * c] & 0xFFu);
    start = placed(cx, c - 1) + kprev - 1 + cx.delay;    // start of the suffix the component is searched in (inside the sequence)
    hi    = cx.n - kc - (c < cx.C - 1 ? cx.delay : 0);   // the pattern fits and leaves room for the next component
    if (cx.band < 0) return false;                       // no pair at all is allowed: the level's row check cannot hold
    // v = candidate + ob must lie in (u, u + band] and inside the sequence
    if (r.n > 0) hi = imin(hi, imin(r.u0 + cx.band, cx.n - 1) - r.ob0);
    if (r.n > 1) hi = imin(hi, imin(r.u1 + cx.band, cx.n - 1) - r.ob1);
    if (r.n > 2) {
        const uint32_t* chk = cx.blob + (cx.blob[3] >> 16) + 2 * (int)(cx.blob[5 + 2 * c] & 0xFFu);
        for (int i = 2; i < r.n; i++)
            hi = imin(hi, imin(placed(cx, anchor_of(chk[2 * i])) + off_of(chk[2 * i]) + cx.band, cx.n - 1) - off_of(chk[2 * i + 1]));
    }
    return start <= hi;
}

// One step of the walk at level c (>= 1): with `enter`, the first candidate of component c below the current
// placement of components 0 .. c-1; otherwise the candidate after the current one. Returns its position or -1.
// A single routine (one copy of the scan loop in the kernel) instead of separate first / next functions.
BSS_HD_LEVEL int level_step(const Ctx& cx, int c, bool enter)
{
    const int       kc = (int)(cx.blob[4 + 2 * c] & 0xFFu), ke = imax(1, kc);
    const bool      thinned = (cx.ovl >> c) & 1u;
    const uint32_t *M = cx.M + c * cx.W, *SM = cx.SM + c * cx.WS;
    const uint32_t* S = thinned ? cx.T + c * cx.W : M;
    int             start, hi;
    bool            greedy;
    Rows r;
    if (!decode_rows(cx, c, r)) return -1;
    if (enter) {
        if (!level_window(cx, c, r, start, hi)) return -1;
        greedy = thinned;   // the reference's iterator restarts at `start`: a greedy walk over all matches until it meets the thinned set
    } else {
        const uint32_t e = cx.pos[c * cx.cs];
        const int      p = (int)(e & 0xFFFFu);
        hi     = (int)(e >> 16);
        greedy = cx.grd[c * cx.cs] != 0;
        if (greedy) {
            start = p + ke;
        } else {
            const uint32_t x = cx.bits[c * cx.cs];   // candidates left in the word of the current one
            if (x) {
                const int q = (p & ~31) + ffs32(x) - 1;
                if (q > hi) return -1;
                cx.bits[c * cx.cs] = x & (x - 1u);
                cx.pos[c * cx.cs]  = (uint32_t)q | ((uint32_t)hi << 16);
                return q;
            }
            start = (p | 31) + 1;
        }
    }
    if (greedy) {
        int q = next_set(M, SM, cx.W, cx.WS, start);
        while (q >= 0 && q <= hi && !bit_of(S, q)) {
            if (row_pass(cx, c, r, q)) {
                cx.grd[c * cx.cs]  = 1;
                cx.bits[c * cx.cs] = 0;
                cx.pos[c * cx.cs]  = (uint32_t)q | ((uint32_t)hi << 16);
                return q;
            }
            q = next_set(M, SM, cx.W, cx.WS, q + ke);
        }
        if (q < 0 || q > hi) return -1;
        start = q;   // the chain has met the thinned set: it is the thinned set from here on
    }
    cx.grd[c * cx.cs] = 0;
    // the thinned part of the chain: a plain loop over the few words of the window (the band bounds it)
    const int wa = start >> 5, wb = hi >> 5;
    for (int w = wa; w <= wb; w++) {
        uint32_t x = S[w];
        if (w == wa) x &= 0xFFFFFFFFu << (start & 31);
        if (w == wb) x &= 0xFFFFFFFFu >> (31 - (hi & 31));
        if (x) x = row_filter(cx, c, r, w, x);
        if (x) {
            const int p = (w << 5) + ffs32(x) - 1;
            cx.bits[c * cx.cs] = x & (x - 1u);
            cx.pos[c * cx.cs]  = (uint32_t)p | ((uint32_t)hi << 16);
            return p;
        }
    }
    return -1;
}

// Every site whose first component lies in the words [w0, w1) of its candidate set. Returns their number;
// when EMIT, writes their records [module, p_0 .. p_{C-1}] from `out` on, in depth-first order.
template <bool EMIT>
BSS_HD_WALK uint64_t walk_words(const Ctx& cx, int w0, int w1, uint32_t module, uint32_t* out)
{
    const int       C = cx.C, R = C + 1;
    const uint32_t* first = (cx.ovl & 1u) ? cx.T : cx.M;   // the chain from position 0 starts on a run start: it is the thinned set
    const int       k0  = (int)(cx.blob[4] & 0xFFu);
    const int       hi0 = cx.n - k0 - (C > 1 ? cx.delay : 0);
    uint64_t        found = 0;
    for (int w = w0; w < w1; w++) {
        for (uint32_t x = first[w]; x; x &= x - 1u) {
            const int p0 = (w << 5) + ffs32(x) - 1;
            if (p0 > hi0) return found;   // later candidates leave even less room
            cx.pos[0] = (uint32_t)p0;
            if ((cx.blob[4] & kCompGeneric) && !generic_pass(cx, 0)) continue;
            if (C == 1) {
                if (EMIT) { out[found * R] = module; out[found * R + 1] = (uint32_t)p0; }
                found++;
                continue;
            }
            int  level = 1;
            bool enter = true;
            for (;;) {
                const int q = level_step(cx, level, enter);
                enter       = false;
                if (q < 0) {   // this level is exhausted: on to the next candidate of its parent
                    if (--level == 0) break;
                    continue;
                }
                if ((cx.blob[4 + 2 * level] & kCompGeneric) && !generic_pass(cx, level)) continue;
                if (level == C - 1) {
                    if (EMIT) {
                        uint32_t* rec = out + found * (uint64_t)R;
                        rec[0]        = module;
                        for (int c = 0; c < C; c++) rec[1 + c] = (uint32_t)placed(cx, c);
                    }
                    found++;
                    continue;
                }
                level++;
                enter = true;
            }
        }
    }
    return found;
}

}   // namespace win
}   // namespace bss
