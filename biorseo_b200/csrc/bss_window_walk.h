// Window walk: the warp-cooperative placement enumerator of window_kernel (bss_window.cuh).
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
//
// The walk is breadth-first by level with a bounded queue per level, and one WARP walks one item:
//   level 0   a lane takes one 32-position word of the first component's candidates;
//   level c   a lane takes one queue entry of level c - 1 (a partial placement p_0 .. p_{c-1}, stored as
//             position | index of its parent entry << 16) and produces all its children at once: the K
//             words of the window, ANDed with the mask rows of the level's row checks, go to the lane's
//             column of a small stash; a warp scan of the child counts gives every lane the place of its
//             children in the queue of level c (or, at the last level, in the output);
//   the deepest level that has entries is always drained first, so a queue never holds more than one
//   round of children and the entries come out in depth-first (= lexicographic = canonical) order.
// The last level is never stored: its children are counted (popcount, count pass) or written straight
// into the record stream (emit pass). Only partial placements whose pairs are all allowed are ever
// visited, and all 32 lanes work whatever the shape of the placement tree.
// The chain semantics of the reference (leftmost non-overlapping matches of the suffix,
// sregex_iterator) are kept: for a component whose matches overlap on this sequence the candidates are
// a few "greedy" matches followed by the thinned set T (see thin_components).
//
// Plain C++ in host/device functions: the same source runs in the kernel (one thread = one lane, warp
// shuffles between the phases) and in the CPU emulation the tests compare with the flat-table
// interpreter (tests/cpp/window_host.cpp: the lanes are a loop, the shuffles plain sums).
#pragma once
#include <cstdint>

#include "bss_layout.h"

#if defined(__CUDACC__)
#define BSS_HD __host__ __device__ __forceinline__
#define BSS_HD_NI __host__ __device__ __noinline__
#else
#define BSS_HD inline
#define BSS_HD_NI inline
#endif

// Lanes: on the device the calling thread is the lane and per-lane state is one struct; on the host the
// lanes are a loop over 32 structs.
#if defined(__CUDA_ARCH__)
#define BSS_LANES_BEGIN { const int lane = (int)(threadIdx.x & 31u); LaneState& L = ls[0];
#define BSS_LANES_END }
#define BSS_LANE_SLOTS 1
#else
#define BSS_LANES_BEGIN for (int lane = 0; lane < 32; lane++) { LaneState& L = ls[lane];
#define BSS_LANES_END }
#define BSS_LANE_SLOTS 32
#endif

// Optional index checks of the debug build (bss_kernels.cu defines it with -DBSS_DEBUG_ASSERT): a violated
// condition is counted and the access is skipped.
#ifndef BSS_WIN_CHECK
#define BSS_WIN_CHECK(cond, id) true
#endif

namespace bss {
namespace win {

#if defined(__CUDA_ARCH__)
BSS_HD int      ffs32(uint32_t x) { return __ffs((int)x); }
BSS_HD int      popc32(uint32_t x) { return __popc(x); }
BSS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
#else
BSS_HD int      ffs32(uint32_t x) { return __builtin_ffs((int)x); }
BSS_HD int      popc32(uint32_t x) { return __builtin_popcount(x); }
BSS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31u; return s ? (lo >> s) | (hi << (32u - s)) : lo; }
#endif

BSS_HD int imax(int a, int b) { return a > b ? a : b; }
BSS_HD int imin(int a, int b) { return a < b ? a : b; }
BSS_HD int anchor_of(uint32_t w) { return (int)(int8_t)(w & 0xFFu); }
BSS_HD int off_of(uint32_t w) { return (int)(int16_t)((w >> 8) & 0xFFFFu); }

constexpr int kLevelWords = 8;   // decoded parameters per level: k, row checks | generic << 8, then (anchor, off, ob) of the first two row checks

struct Walk {
    const uint32_t* blob;   // unit descriptor
    const uint32_t* const* rows;   // [3 * C] per component: every match (M, W words), the thinned matches (T, valid for the components in
                                   // `ovl`), the one-bit-per-word summary of M (WS words) — rows of the warp's scratch or of the pattern table
    const uint32_t* bm;     // banded pair mask of the sequence: row u = K words, the mask words (u >> 5) .. of row u
    int*            lp;     // [C][kLevelWords] decoded level parameters (written by walk())
    uint32_t*       ctl;    // [2 * C] saved cursor / queue length per level
    uint32_t*       Q;      // [max(1, C - 1)][qcap] level queues: position | parent entry << 16
    uint32_t*       stash;  // [K][32] children of the lane's entry, one column per lane
    uint32_t        ovl;    // bit c: matches of component c overlap on this sequence
    int             W, WS, n, K, C, delay, band, qcap;
};

// First set bit at or after position t (t >= 0) of a bitset with a one-bit-per-word summary, or -1.
BSS_HD int next_set(const uint32_t* B, const uint32_t* S, int W, int WS, int t)
{
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
    return (w << 5) + ffs32(B[w]) - 1;
}

BSS_HD bool bit_of(const uint32_t* B, int q) { return (B[q >> 5] >> (q & 31)) & 1u; }

// Word w (absolute index into the mask row) of row u of the pair mask; zero outside the band.
BSS_HD uint32_t row_word(const Walk& cx, int u, int w)
{
    const unsigned j = (unsigned)(w - (u >> 5));
    return j < (unsigned)cx.K ? cx.bm[u * cx.K + (int)j] : 0u;
}

BSS_HD bool pair_allowed(const Walk& cx, int u, int v)
{
    const int a = imin(u, v), b = imax(u, v);
    if (a < 0 || b >= cx.n || a == b) return false;
    return (row_word(cx, a, b >> 5) >> (b & 31)) & 1u;
}

// bit q of the result = bit q + ob of row u of the pair mask, for the positions of word w
BSS_HD uint32_t row_bits(const Walk& cx, int u, int ob, int w)
{
    const int ws = w + (ob >> 5);
    return funnel_r(row_word(cx, u, ws), row_word(cx, u, ws + 1), (uint32_t)ob & 31u);
}

// Position of component `a` in the partial placement that ends with entry `ent` of level `level` (a <= level).
BSS_HD int prefix_pos(const Walk& cx, int level, uint32_t ent, int a)
{
    for (int l = level; l > a; l--) ent = cx.Q[(l - 1) * cx.qcap + (int)(ent >> 16)];
    return (int)(ent & 0xFFFFu);
}

// The checks of level c that are not row checks, for candidate q of component c below the partial placement that
// ends with `parent` (an entry of level c - 1; unused at level 0). Rare: pairs inside one strand, absolute
// positions, ends outside their component.
BSS_HD_NI bool generic_pass(const Walk& cx, int c, uint32_t parent, int q)
{
    const uint32_t  range = cx.blob[5 + 2 * c];
    const int       cb = (int)(range & 0xFFu) + (int)(range >> 24), ce = (int)((range >> 8) & 0xFFu);   // (the row checks come first)
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16);
    for (int i = cb; i < ce; i++) {
        const uint32_t ea = chk[2 * i], eb = chk[2 * i + 1];
        const int aa = anchor_of(ea), ba = anchor_of(eb);
        const int u  = (aa < 0 ? 0 : aa == c ? q : prefix_pos(cx, c - 1, parent, aa)) + off_of(ea);
        const int v  = (ba < 0 ? 0 : ba == c ? q : prefix_pos(cx, c - 1, parent, ba)) + off_of(eb);
        if (!pair_allowed(cx, u, v)) return false;
    }
    return true;
}

BSS_HD_NI uint32_t generic_filter(const Walk& cx, int c, uint32_t parent, int w, uint32_t x)
{
    uint32_t keep = 0;
    for (uint32_t y = x; y; y &= y - 1u) {
        const int b = ffs32(y) - 1;
        if (generic_pass(cx, c, parent, (w << 5) + b)) keep |= 1u << b;
    }
    return keep;
}

// Decoded parameters of level c, by one lane per level at the start of an item.
BSS_HD void decode_level(const Walk& cx, int c)
{
    int*            lp    = cx.lp + kLevelWords * c;
    const uint32_t  desc  = cx.blob[4 + 2 * c];
    const uint32_t  range = cx.blob[5 + 2 * c];
    const uint32_t* chk   = cx.blob + (cx.blob[3] >> 16) + 2 * (int)(range & 0xFFu);
    const int       rows  = (int)(range >> 24);
    lp[0] = (int)(desc & 0xFFu);
    lp[1] = rows | ((desc & kCompGeneric) ? 0x100 : 0);
    for (int i = 0; i < 2; i++) {
        lp[2 + 3 * i] = i < rows ? anchor_of(chk[2 * i]) : 0;
        lp[3 + 3 * i] = i < rows ? off_of(chk[2 * i]) : 0;
        lp[4 + 3 * i] = i < rows ? off_of(chk[2 * i + 1]) : 0;
    }
}

// Level 0: the candidates of the first component in word w (a bit mask).
BSS_HD uint32_t root_word(const Walk& cx, int w, int Wq)
{
    if (w >= Wq) return 0u;
    const uint32_t* first = cx.rows[(cx.ovl & 1u) ? 1 : 0];   // the chain from position 0 starts on a run start: it is the thinned set
    const int       hi0   = cx.n - cx.lp[0] - (cx.C > 1 ? cx.delay : 0);   // the pattern fits and leaves room for the next component
    const int       left  = hi0 - (w << 5);
    if (left < 0) return 0u;
    uint32_t x = first[w];
    if (left < 31) x &= (2u << left) - 1u;
    if (x && (cx.lp[1] & 0x100)) x = generic_filter(cx, 0, 0u, w, x);
    return x;
}

// Level c >= 1: every child of entry e of level c - 1, as bit masks in the lane's stash column (words wa .. wa + nw - 1
// of the sequence). Returns their number.
BSS_HD uint32_t expand(const Walk& cx, int c, uint32_t e, int lane, int& wa, int& nw)
{
    nw = 0;
    wa = 0;
    if (cx.band < 0) return 0u;   // no pair at all is allowed: the level's row check cannot hold
    const uint32_t ent  = cx.Q[(c - 1) * cx.qcap + (int)e];
    const int*     lp   = cx.lp + kLevelWords * c;
    const int      kc   = lp[0], rows = lp[1] & 0xFF;
    int            start = (int)(ent & 0xFFFFu) + cx.lp[kLevelWords * (c - 1)] - 1 + cx.delay;   // start of the suffix the component is searched in
    int            hi    = cx.n - kc - (c < cx.C - 1 ? cx.delay : 0);                            // the pattern fits and leaves room for the next one
    // v = candidate + ob must lie in (u, u + band] and inside the sequence
    int u0 = 0, u1 = 0;
    if (rows > 0) {
        u0 = prefix_pos(cx, c - 1, ent, lp[2]) + lp[3];
        if (u0 < 0) return 0u;
        hi = imin(hi, imin(u0 + cx.band, cx.n - 1) - lp[4]);
    }
    if (rows > 1) {
        u1 = prefix_pos(cx, c - 1, ent, lp[5]) + lp[6];
        if (u1 < 0) return 0u;
        hi = imin(hi, imin(u1 + cx.band, cx.n - 1) - lp[7]);
    }
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16) + 2 * (int)(cx.blob[5 + 2 * c] & 0xFFu);
    for (int i = 2; i < rows; i++) {   // (rare: more than two row checks on one level)
        const int u = prefix_pos(cx, c - 1, ent, anchor_of(chk[2 * i])) + off_of(chk[2 * i]);
        if (u < 0) return 0u;
        hi = imin(hi, imin(u + cx.band, cx.n - 1) - off_of(chk[2 * i + 1]));
    }
    if (start > hi) return 0u;
    const bool      generic = (lp[1] & 0x100) != 0;
    const bool      thinned = (cx.ovl >> c) & 1u;
    const uint32_t* Mc      = cx.rows[3 * c];
    const uint32_t* S       = thinned ? cx.rows[3 * c + 1] : Mc;
    uint32_t*       col     = cx.stash + lane;
    uint32_t        found   = 0;
    wa = start >> 5;
    nw = (hi >> 5) - wa + 1;   // at most K words: the window lies inside (u, u + band]
    if (nw > cx.K) nw = cx.K;  // (never taken; keeps a corrupt descriptor inside the stash)
    if (thinned) {
        // The reference's iterator restarts at `start`: a greedy walk over all matches until it meets the thinned set.
        for (int j = 0; j < nw; j++) col[32 * j] = 0u;
        const uint32_t* SMc = cx.rows[3 * c + 2];
        const int       ke  = imax(1, kc);
        int             q   = next_set(Mc, SMc, cx.W, cx.WS, start);
        while (q >= 0 && q <= hi && !bit_of(S, q)) {
            const int w = q >> 5;
            uint32_t  x = 1u << (q & 31);
            if (rows > 0) x &= row_bits(cx, u0, lp[4], w);
            if (rows > 1 && x) x &= row_bits(cx, u1, lp[7], w);
            for (int i = 2; i < rows && x; i++) x &= row_bits(cx, prefix_pos(cx, c - 1, ent, anchor_of(chk[2 * i])) + off_of(chk[2 * i]), off_of(chk[2 * i + 1]), w);
            if (x && generic) x = generic_filter(cx, c, ent, w, x);
            if (x && w - wa < nw) {
                col[32 * (w - wa)] |= x;
                found++;
            }
            q = next_set(Mc, SMc, cx.W, cx.WS, q + ke);
        }
        if (q < 0 || q > hi) return found;
        start = q;   // the chain has met the thinned set: it is the thinned set from here on
    }
    const int wb = hi >> 5;
    for (int j = (start >> 5) - wa; j < nw; j++) {
        const int w = wa + j;
        uint32_t  x = S[w];
        if (w == (start >> 5)) x &= 0xFFFFFFFFu << (start & 31);
        if (w == wb) x &= 0xFFFFFFFFu >> (31 - (hi & 31));
        if (x) {
            if (rows > 0) x &= row_bits(cx, u0, lp[4], w);
            if (rows > 1 && x) x &= row_bits(cx, u1, lp[7], w);
            for (int i = 2; i < rows && x; i++) x &= row_bits(cx, prefix_pos(cx, c - 1, ent, anchor_of(chk[2 * i])) + off_of(chk[2 * i]), off_of(chk[2 * i + 1]), w);
            if (x && generic) x = generic_filter(cx, c, ent, w, x);
        }
        if (thinned) x |= col[32 * j];
        col[32 * j] = x;
        found += (uint32_t)popc32(x);
    }
    if (thinned) {   // (the greedy matches were counted one by one and again with their words)
        found = 0;
        for (int j = 0; j < nw; j++) found += (uint32_t)popc32(col[32 * j]);
    }
    return found;
}

struct LaneState {
    uint32_t cnt, excl, x;
    int      wa, nw;
};

// Exclusive prefix of the lanes' counts; returns the total.
BSS_HD uint32_t scan_lanes(LaneState* ls)
{
#if defined(__CUDA_ARCH__)
    const int lane = (int)(threadIdx.x & 31u);
    uint32_t  incl = ls[0].cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += up;
    }
    ls[0].excl = incl - ls[0].cnt;
    return __shfl_sync(0xFFFFFFFFu, incl, 31);
#else
    uint32_t run = 0;
    for (int l = 0; l < 32; l++) {
        ls[l].excl = run;
        run += ls[l].cnt;
    }
    return run;
#endif
}

// Number of leading lanes whose children all fit into `room` queue slots (the lanes' inclusive sums are monotone).
BSS_HD int lanes_that_fit(const LaneState* ls, uint32_t room)
{
#if defined(__CUDA_ARCH__)
    return __popc(__ballot_sync(0xFFFFFFFFu, ls[0].excl + ls[0].cnt <= room));
#else
    int take = 0;
    while (take < 32 && ls[take].excl + ls[take].cnt <= room) take++;
    return take;
#endif
}

BSS_HD void lanes_sync()
{
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

// Every site of the item. Returns their number (the same on every lane); when EMIT, writes their records
// [module, p_0 .. p_{C-1}] from `out` on, in depth-first order. Called by all 32 lanes of a warp (device) or once
// (host emulation).
template <bool EMIT>
BSS_HD uint64_t walk(const Walk& cx, int Wq, uint32_t module, uint32_t* out)
{
    const int C = cx.C, R = C + 1;
    LaneState ls[BSS_LANE_SLOTS];
    BSS_LANES_BEGIN
        (void)L;
        if (lane < C) decode_level(cx, lane);
    BSS_LANES_END
    lanes_sync();
    uint64_t found = 0;
    int      c     = 0;   // the level being placed
    uint32_t cur   = 0;   // next task of that level: a word of the sequence (level 0) or an entry of the queue of level c - 1
    uint32_t tail  = 0;   // entries in the queue of level c
    for (;;) {
        const uint32_t tasks = c == 0 ? (uint32_t)Wq : cx.ctl[C + c - 1];
        if (cur >= tasks) {
            if (c < C - 1 && tail > 0) {   // this level's queue holds the last children: place the next level below them
                cx.ctl[c]     = cur;
                cx.ctl[C + c] = tail;
                lanes_sync();
                c++;
                cur = tail = 0;
                continue;
            }
            if (c == 0) break;
            c--;                 // the queue of level c has been consumed
            cur  = cx.ctl[c];
            tail = 0;
            continue;
        }
        // one round: 32 tasks of level c
        BSS_LANES_BEGIN
            const uint32_t task = cur + (uint32_t)lane;
            L.cnt = 0;
            L.x   = 0;
            L.wa = L.nw = 0;
            if (task < tasks) {
                if (c == 0) {
                    L.x   = root_word(cx, (int)task, Wq);
                    L.cnt = (uint32_t)popc32(L.x);
                } else {
                    L.cnt = expand(cx, c, task, lane, L.wa, L.nw);
                }
            }
        BSS_LANES_END
        const bool leaf = c == C - 1;
        if (leaf && !EMIT) {   // count pass: the children of the last level are only counted
            uint32_t total = 0;
#if defined(__CUDA_ARCH__)
            total = ls[0].cnt;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, d);
#else
            for (int l = 0; l < 32; l++) total += ls[l].cnt;
#endif
            found += total;
            cur += 32;
            continue;
        }
        const uint32_t total = scan_lanes(ls);
        if (leaf) {
            BSS_LANES_BEGIN
                if (L.cnt) {
                    uint32_t*      rec    = out + (found + L.excl) * (uint64_t)R;
                    const uint32_t parent = c == 0 ? 0u : cx.Q[(c - 1) * cx.qcap + (int)(cur + (uint32_t)lane)];
                    for (int j = 0; j < (c == 0 ? 1 : L.nw); j++) {
                        const uint32_t x  = c == 0 ? L.x : cx.stash[32 * j + lane];
                        const int      w0 = c == 0 ? (int)(cur + (uint32_t)lane) : L.wa + j;
                        for (uint32_t y = x; y; y &= y - 1u) {
                            rec[0] = module;
                            rec[C] = (uint32_t)((w0 << 5) + ffs32(y) - 1);
                            uint32_t ent = parent;
                            for (int a = C - 2; a >= 0; a--) {
                                rec[1 + a] = ent & 0xFFFFu;
                                if (a > 0) ent = cx.Q[(a - 1) * cx.qcap + (int)(ent >> 16)];
                            }
                            rec += R;
                        }
                    }
                }
            BSS_LANES_END
            found += total;
            cur += 32;
            continue;
        }
        // children into the queue of level c, in lane order; lanes whose children do not fit are taken again later
        const uint32_t room = (uint32_t)cx.qcap - tail;
        const int      take = total <= room ? 32 : lanes_that_fit(ls, room);
        uint32_t       added = 0;
        BSS_LANES_BEGIN
            if (lane < take) {
                if (L.cnt) {
                    uint32_t*      q      = cx.Q + c * cx.qcap + (int)(tail + L.excl);
                    const uint32_t parent = c == 0 ? 0u : (cur + (uint32_t)lane) << 16;
                    for (int j = 0; j < (c == 0 ? 1 : L.nw); j++) {
                        const uint32_t x  = c == 0 ? L.x : cx.stash[32 * j + lane];
                        const int      w0 = c == 0 ? (int)(cur + (uint32_t)lane) : L.wa + j;
                        for (uint32_t y = x; y; y &= y - 1u) *q++ = (uint32_t)((w0 << 5) + ffs32(y) - 1) | parent;
                    }
                }
#if !defined(__CUDA_ARCH__)
                added += L.cnt;
#endif
            }
        BSS_LANES_END
#if defined(__CUDA_ARCH__)
        added = take == 32 ? total : take == 0 ? 0u : __shfl_sync(0xFFFFFFFFu, ls[0].excl + ls[0].cnt, take - 1);
#endif
        tail += added;
        cur += (uint32_t)take;
        if (take == 0 && tail == 0) cur++;   // (cannot happen: one lane's children always fit an empty queue; never spin on a corrupt descriptor)
        lanes_sync();
        if (take < 32 && tail > 0) {   // the queue is full: place the next level below what it holds, then come back
            cx.ctl[c]     = cur;
            cx.ctl[C + c] = tail;
            lanes_sync();
            c++;
            cur = tail = 0;
        }
    }
    return found;
}

}   // namespace win
}   // namespace bss
