// Hand-written sm_100a kernels of the insertion-site search.
//
// What they replace (reference, CPU): one Pool job per module running std::regex over the
// query and materialising every placement before filtering it
// (cppsrc/Motif.cpp:432-494, cppsrc/MOIP.cpp:1069-1085, 1126-1143, 1178-1212).
//
// Shape of the computation (DESIGN.md section 3 has the long version).
// Once per query: occ_kernel turns every sequence into per-symbol (and per-symbol-pair)
// occurrence bit rows, band_kernel finds the widest pair its mask allows.
// Per search, search_kernel<G, MODE>; a group of G lanes (32, or 4..16 for large batches of short
// sequences) owns one (sequence, unit) item at a time:
//   count pass   one CTA per chunk = (sequence, contiguous unit range), groups draw units from the
//                chunk's counter. Per item:
//     1. match   M_c = AND over the component's match program of shifted occurrence rows, 32
//                positions per lane and step, with a one-bit-per-word summary;
//     2. thin    components whose matches overlap on this sequence get T_c, the leftmost
//                non-overlapping matches of every run. The reference's candidates in the suffix
//                starting at t (sregex_iterator) are: a greedy walk over M_c from t until it
//                lands on a bit of T_c, and T_c itself from there on;
//     3. walk    32-lane groups: sorted match lists, a backward counting programme (placements
//                below every match before the checks; children are list ranges restricted by
//                the band windows of adjacent checks), then every lane walks one of 32 equal
//                runs of consecutive ranks with an odometer, testing the checks (rank-space
//                walk). Narrow groups / oversized lists: per-lane depth-first walk with the same
//                chain semantics and windows. Pruning is result-neutral;
//                per-item counts, per-lane counts, the alive list and chunk totals are recorded;
//   scan         chunk totals -> 64-bit offsets;
//   emit pass    the groups stride over the alive list; every item re-derives its lists, takes
//                its place from the scanned offsets and writes [module, p_0..p_{C-1}] records in
//                canonical (sequence, unit, depth-first) order. No sort, no output atomics;
//   heavy passes items with very many placements are parked by the count pass and walked as
//                (item, block of ranks) tasks spread over a GPU-wide grid, with subtree pruning.
// Not a contraction: integer/bit work only, no tensor cores.
#include "bss_device.cuh"

#ifdef BSS_DEBUG_ASSERT
namespace bss { __device__ unsigned int g_assert_fail[32]; }
#ifdef __CUDA_ARCH__
#define BSS_WIN_CHECK(cond, id) ((cond) ? true : (atomicAdd(&bss::g_assert_fail[id], 1u), false))
#endif
#endif
#include "bss_window_walk.h"

#include <algorithm>


#ifndef BSS_MIN_CTAS
#define BSS_MIN_CTAS 8   // resident CTAs per SM the register allocation is held to
#endif

#ifdef BSS_DEBUG_SYNC
#include <cstdio>
#define BSS_SYNC_CHECK(name, stream)                                                                   \
    do {                                                                                               \
        cudaError_t e_ = cudaStreamSynchronize(stream);                                                \
        if (e_ != cudaSuccess) { std::fprintf(stderr, "[bss debug] %s: %s\n", name, cudaGetErrorString(e_)); return e_; } \
    } while (0)
#else
#define BSS_SYNC_CHECK(name, stream) do { } while (0)
#endif

namespace bss {

// Optional per-phase cycle accounting (debug builds only: -DBSS_PHASE_TIMING). Sums, over all
// groups, the cycles between phase boundaries; read back with bss_debug_phase_cycles().
#ifdef BSS_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define BSS_PHASE(id)                                                                  \
    do {                                                                               \
        const long long now_ = clock64();                                              \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase_cycles[id], (unsigned long long)(now_ - phase_t0_)); \
        phase_t0_ = now_;                                                              \
    } while (0)
#define BSS_PHASE_BEGIN() long long phase_t0_ = clock64()
#else
#define BSS_PHASE(id) do { } while (0)
#define BSS_PHASE_BEGIN() do { } while (0)
#endif

// Optional index checks (debug builds only: -DBSS_DEBUG_ASSERT; compute-sanitizer is not
// available on the GPU pool). Violations are counted per site and read back with
// bss_debug_assert_counts(); the fuzz tests are run against such a build.
#ifdef BSS_DEBUG_ASSERT
#define BSS_ASSERT(cond, id) do { if (!(cond)) atomicAdd(&g_assert_fail[id], 1u); } while (0)
#else
#define BSS_ASSERT(cond, id) do { } while (0)
#endif

namespace {

__device__ __forceinline__ int unpack_anchor(uint32_t w) { return (int)(int8_t)(w & 0xFFu); }
__device__ __forceinline__ int unpack_off(uint32_t w) { return (int)(int16_t)((w >> 8) & 0xFFFFu); }

// First set bit at or after position t (t >= 0) of a bitset with a one-bit-per-word summary, or -1.
__device__ __forceinline__ int next_bit(const uint32_t* __restrict__ B, const uint32_t* __restrict__ S, int W, int WS, int t)
{
    BSS_ASSERT(t >= 0, 0);
    int w = t >> 5;
    if (w >= W) return -1;
    uint32_t x = B[w] & (0xFFFFFFFFu << (t & 31));
    if (x) return (w << 5) + __ffs(x) - 1;
    if (++w >= W) return -1;
    int      sw = w >> 5;
    uint32_t s  = S[sw] & (0xFFFFFFFFu << (w & 31));
    while (!s) {
        if (++sw >= WS) return -1;
        s = S[sw];
    }
    w = (sw << 5) + __ffs(s) - 1;
    return (w << 5) + __ffs(B[w]) - 1;
}

__device__ __forceinline__ bool test_bit(const uint32_t* __restrict__ B, int q)
{
    BSS_ASSERT(q >= 0, 1);
    return (B[q >> 5] >> (q & 31)) & 1u;
}

// Word w of (B << s): bit q of the result is bit q - s of B.
__device__ __forceinline__ uint32_t shifted_up(const uint32_t* __restrict__ B, int w, int s)
{
    const int      i  = w - (s >> 5);
    const uint32_t hi = i >= 0 ? B[i] : 0u, lo = i >= 1 ? B[i - 1] : 0u;
    return __funnelshift_l(lo, hi, s & 31);
}

struct ItemCtx {
    const uint32_t* blob;     // unit descriptor (shared or global memory)
    const uint32_t* M;        // [C][W] every match of every component
    const uint32_t* T;        // [C][W] thinned matches, valid for the components in `ovl`
    const uint32_t* SM;       // [C][WS] summaries of M
    const uint32_t* ST;       // [C][WS] summaries of T
    const uint32_t* mask;     // pair mask of the sequence, row stride mstride words
    uint32_t*       stack;    // this thread's column (stride kThreads): position | last position of the window << 16
    uint32_t        ovl;      // bit c: matches of component c overlap on this sequence
    int             W, WS, n, mstride, C, delay, band;
};

__device__ __forceinline__ int placed(const ItemCtx& cx, int level)
{
    BSS_ASSERT(level >= 0 && level < cx.C, 2);
    return (int)(cx.stack[level * kThreads] & 0xFFFFu);
}

// All checks that become decidable once component `level` is placed.
__device__ __forceinline__ bool checks_pass(const ItemCtx& cx, int level)
{
    const uint32_t range = cx.blob[5 + 2 * level];
    const int      cb = (int)(range & 0xFFu), ce = (int)((range >> 8) & 0xFFu);
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16);
    for (int i = cb; i < ce; i++) {
        const uint32_t ea = chk[2 * i], eb = chk[2 * i + 1];
        const int      aa = unpack_anchor(ea), ba = unpack_anchor(eb);
        const int      u  = (aa < 0 ? 0 : placed(cx, aa)) + unpack_off(ea);
        const int      v  = (ba < 0 ? 0 : placed(cx, ba)) + unpack_off(eb);
        int            a = u, b = v;
        if (!(ea & kCheckOrdered)) {   // ordered checks have 0 <= u < v < n by construction
            a = min(u, v);
            b = max(u, v);
            if (a < 0 || b >= cx.n || a == b) return false;
        }
        BSS_ASSERT(a >= 0 && a < b && b < cx.n, 3);
        if (!((__ldg(cx.mask + (size_t)a * cx.mstride + (b >> 5)) >> (b & 31)) & 1u)) return false;
    }
    return true;
}

// Next candidate of component `level` at or after max(after, lo). `after` is where the
// reference's iterator would resume (suffix start, or one pattern length after the previous
// match). While the level's bit is set in `greedy` the chain is still a greedy walk over all
// matches that has not met the thinned set yet.
__device__ __forceinline__ int chain_next(const ItemCtx& cx, int level, int after, int lo, uint32_t& greedy)
{
    const uint32_t bit = 1u << level;
    const int      W = cx.W, WS = cx.WS;
    if (greedy & bit) {
        const uint32_t *M = cx.M + level * W, *SM = cx.SM + level * WS, *T = cx.T + level * W;
        const int       k = max(1, (int)(cx.blob[4 + 2 * level] & 0xFFu));
        int             q = next_bit(M, SM, W, WS, after);
        while (q >= 0) {
            if (test_bit(T, q)) {   // from here on the chain is the thinned set
                greedy &= ~bit;
                return q >= lo ? q : next_bit(T, cx.ST + level * WS, W, WS, lo);
            }
            if (q >= lo) return q;
            q = next_bit(M, SM, W, WS, q + k);
        }
        return -1;
    }
    const bool thinned = (cx.ovl >> level) & 1u;
    return next_bit((thinned ? cx.T : cx.M) + level * W, (thinned ? cx.ST : cx.SM) + level * WS, W, WS, max(after, lo));
}

// First candidate of component `level` in the suffix starting at t. Checks that tie this
// component to an already known position u restrict it to |v - u| <= band.
__device__ __forceinline__ int enter_level(const ItemCtx& cx, int level, int t, uint32_t& greedy)
{
    int            lo = t, hi = cx.n;
    const uint32_t range = cx.blob[5 + 2 * level];
    const int      cb = (int)(range & 0xFFu), ce = (int)((range >> 8) & 0xFFu);
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16);
    for (int i = cb; i < ce; i++) {
        const uint32_t ea = chk[2 * i];
        if (ea & kCheckFixed) continue;
        const int aa = unpack_anchor(ea);
        const int u  = (aa < 0 ? 0 : placed(cx, aa)) + unpack_off(ea);
        if (u < 0 || u >= cx.n || cx.band < 0) return -1;
        const int ob = unpack_off(chk[2 * i + 1]);
        lo = max(lo, u - cx.band - ob);
        hi = min(hi, u + cx.band - ob);
    }
    if (lo > hi) return -1;
    const uint32_t bit = 1u << level;
    greedy             = (greedy & ~bit) | (cx.ovl & bit);
    const int q        = chain_next(cx, level, t, lo, greedy);
    if (q < 0 || q > hi) return -1;
    cx.stack[level * kThreads] = (uint32_t)q | ((uint32_t)hi << 16);
    return q;
}

__device__ __forceinline__ int advance_level(const ItemCtx& cx, int level, uint32_t& greedy)
{
    const uint32_t e = cx.stack[level * kThreads];
    const int      q = (int)(e & 0xFFFFu), hi = (int)(e >> 16);
    const int      k = max(1, (int)(cx.blob[4 + 2 * level] & 0xFFu));   // matches of one pattern never overlap
    const int      q2 = chain_next(cx, level, q + k, 0, greedy);
    if (q2 < 0 || q2 > hi) return -1;
    cx.stack[level * kThreads] = (uint32_t)q2 | ((uint32_t)hi << 16);
    return q2;
}

// Enumerates every site whose component 0 sits at p0. Returns the number of sites; when
// EMIT, also writes their records starting at out (C + 1 words each).
template <bool EMIT>
__device__ __forceinline__ uint64_t walk_subtree(const ItemCtx& cx, int p0, uint32_t module, uint32_t* __restrict__ out)
{
    const int C = cx.C;
    cx.stack[0] = (uint32_t)p0;
    if (!checks_pass(cx, 0)) return 0;
    if (C == 1) {
        if (EMIT) { out[0] = module; out[1] = (uint32_t)p0; }
        return 1;
    }
    uint32_t greedy = 0;
    int      level  = 1, q;
    {
        const int t = p0 + (int)(cx.blob[4] & 0xFFu) - 1 + cx.delay;
        if (t < 0 || t >= cx.n) return 0;   // no room for the next component
        q = enter_level(cx, 1, t, greedy);
    }
    uint64_t found = 0;
    for (;;) {
        if (q < 0) {   // this level is exhausted: back to the next candidate of its parent
            if (--level == 0) break;
            q = advance_level(cx, level, greedy);
            continue;
        }
        if (checks_pass(cx, level)) {
            if (level == C - 1) {
                if (EMIT) {
                    uint32_t* rec = out + found * (uint64_t)(C + 1);
                    rec[0] = module;
                    for (int c = 0; c < C; c++) rec[1 + c] = (uint32_t)placed(cx, c);
                }
                found++;
            } else {
                const int t = q + (int)(cx.blob[4 + 2 * level] & 0xFFu) - 1 + cx.delay;
                if (t < 0 || t >= cx.n) {   // later matches of this level leave even less room
                    if (--level == 0) break;
                    q = advance_level(cx, level, greedy);
                    continue;
                }
                level++;
                q = enter_level(cx, level, t, greedy);
                continue;
            }
        }
        q = advance_level(cx, level, greedy);
    }
    return found;
}

template <int G>
struct Group {
    static __device__ __forceinline__ uint32_t lane() { return threadIdx.x & (G - 1); }
    static __device__ __forceinline__ uint32_t shift() { return (threadIdx.x & 31u) & ~(uint32_t)(G - 1); }
    static __device__ __forceinline__ uint32_t mask() { return G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << shift()); }
    static __device__ __forceinline__ uint32_t ballot(bool p)
    {
        const uint32_t b = __ballot_sync(mask(), p);
        return G == 32 ? b : ((b >> shift()) & ((1u << G) - 1u));
    }
    static __device__ __forceinline__ void sync() { __syncwarp(mask()); }
    static __device__ __forceinline__ uint32_t first(uint32_t v) { return __shfl_sync(mask(), v, 0, G); }
    template <typename T>
    static __device__ __forceinline__ T sum(T v)
    {
#pragma unroll
        for (int d = G / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(mask(), v, d);
        return v;
    }
    // Exclusive prefix over the group's lanes; `total` receives the group sum.
    static __device__ __forceinline__ uint64_t exclusive(uint64_t v, uint64_t& total)
    {
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const uint64_t up = __shfl_up_sync(mask(), incl, d, G);
            if ((int)lane() >= d) incl += up;
        }
        total = __shfl_sync(mask(), incl, G - 1, G);
        return incl - v;
    }
};

// Match masks of the components of one unit: M[c][w] bit b = pattern c matches at 32w + b.
// Every component carries a match program (row of the query's occurrence table, shift): the
// mask is the AND of the shifted rows. Rows are single symbols or, for small alphabets, pairs
// of adjacent symbols (half the steps). The table lives in global memory (a few KB per
// sequence, built once per query) and is read through L1.
// Returns false as soon as one component has no match at all (the unit cannot be placed).
template <int G>
__device__ __forceinline__ bool match_components(const uint32_t* blob, const uint32_t* __restrict__ occ, int WP, uint32_t* M,
                                                 uint32_t* S, int W, int WS, int n)
{
    using Gr       = Group<G>;
    const int  C   = (int)(blob[1] & 0xFFu);
    const int  Wq  = (n + 32) >> 5;   // ceil((n + 1) / 32): positions 0..n
    const uint16_t* halves = reinterpret_cast<const uint16_t*>(blob);
    for (int c = 0; c < C; c++) {
        const uint32_t  desc  = blob[4 + 2 * c];
        const int       k     = (int)(desc & 0xFFu);
        const int       steps = (int)((blob[5 + 2 * c] >> 16) & 0xFFu);
        const uint16_t* prog  = halves + ((desc >> 8) & 0xFFFFu);
        const int       last  = n - k;   // last position where the pattern still fits (k = 0: position n)
        bool            any   = false;
        for (int w0 = 0; w0 < W; w0 += G) {
            const int w = w0 + (int)Gr::lane();
            uint32_t  m = 0;
            if (w < Wq && last >= (w << 5)) {
                const int hi = last - (w << 5);
                m            = hi >= 31 ? 0xFFFFFFFFu : ((2u << hi) - 1u);
#pragma unroll 4
                for (int i = 0; i < steps; i++) {
                    const uint32_t  st  = prog[i];
                    const uint32_t* row = occ + (st & 0x7Fu) * WP + w + (st >> 12);
                    m &= __funnelshift_r(__ldg(row), __ldg(row + 1), (st >> 7) & 31u);
                }
            }
            if (w < W) M[c * W + w] = m;
            const uint32_t bits = Gr::ballot(m != 0);
            if (Gr::lane() == 0) S[c * WS + (w0 >> 5)] = bits;
            any |= bits != 0;
        }
        if (!any) return false;
    }
    Gr::sync();
    return true;
}

// Thinned match sets. A run is a maximal series of matches in which consecutive matches are
// less than k apart; T_c holds, for every run, its leftmost non-overlapping matches counted
// from the start of the run. Only components whose pattern can overlap a shifted copy of
// itself (kCompOverlap) and whose matches do overlap on this sequence get one; their bit is
// set in the returned mask.
template <int G>
__device__ __forceinline__ uint32_t thin_components(const uint32_t* blob, const uint32_t* M, const uint32_t* SM, uint32_t* T,
                                                    uint32_t* ST, int W, int WS)
{
    using Gr     = Group<G>;
    const int C  = (int)(blob[1] & 0xFFu);
    uint32_t ovl = 0;
    for (int c = 0; c < C; c++) {
        const uint32_t desc = blob[4 + 2 * c];
        if (!(desc & kCompOverlap)) continue;
        const int       k  = (int)(desc & 0xFFu);
        const uint32_t* Mc = M + c * W;
        const uint32_t* Sc = SM + c * WS;
        uint32_t*       Tc = T + c * W;
        bool            any = false;
        for (int w0 = 0; w0 < W; w0 += G) {   // followers: matches with another match in the k-1 positions before them
            const int w = w0 + (int)Gr::lane();
            uint32_t  near = 0;
            if (w < W) {
                const uint32_t m = Mc[w];
                if (m)
                    for (int s = 1; s < k; s++) near |= shifted_up(Mc, w, s);
                near &= m;
                Tc[w] = near;
            }
            any |= Gr::ballot(near != 0) != 0;
        }
        if (!any) continue;   // no two matches overlap on this sequence: the chain from any start is M itself
        ovl |= 1u << c;
        Gr::sync();
        // A run needs a greedy walk only if it holds a follower of a follower (three matches or
        // more): with two, the thinned set is just the run starts.
        bool deep = false;
        for (int w0 = 0; w0 < W; w0 += G) {
            const int w = w0 + (int)Gr::lane();
            uint32_t  f2 = 0;
            if (w < W && Tc[w])
                for (int s = 1; s < k; s++) f2 |= shifted_up(Tc, w, s);
            deep |= Gr::ballot(w < W && (f2 & Tc[w]) != 0) != 0;
        }
        Gr::sync();
        for (int w = (int)Gr::lane(); w < W; w += G) Tc[w] = Mc[w] & ~Tc[w];   // run starts
        if (deep) {
            Gr::sync();
            // Greedy picks inside every run. A lane that reads a word after another lane has
            // added picks to it restarts from a pick, which reproduces the same chain.
            for (int w = (int)Gr::lane(); w < W; w += G) {
                for (uint32_t x = Tc[w]; x; x &= x - 1) {
                    int prev = (w << 5) + __ffs(x) - 1, last = prev;
                    for (;;) {
                        const int m = next_bit(Mc, Sc, W, WS, prev + 1);
                        if (m < 0 || m - prev >= k) break;
                        if (m - last >= k) {
                            atomicOr(&Tc[m >> 5], 1u << (m & 31));
                            last = m;
                        }
                        prev = m;
                    }
                }
            }
        }
        Gr::sync();
        for (int w0 = 0; w0 < W; w0 += G) {
            const int      w    = w0 + (int)Gr::lane();
            const uint32_t bits = Gr::ballot(w < W && Tc[w] != 0);
            if (Gr::lane() == 0) ST[c * WS + (w0 >> 5)] = bits;
        }
    }
    Gr::sync();
    return ovl;
}

// Greedy earliest placement: does the gate unit have any placement on this sequence?
template <int G>
__device__ __forceinline__ bool gate_open(const uint32_t* gblob, const uint32_t* occ, int WP, uint32_t* M, uint32_t* S, int W,
                                          int WS, int n)
{
    if (!match_components<G>(gblob, occ, WP, M, S, W, WS, n)) return false;
    const int C = (int)(gblob[1] & 0xFFu), delay = (int)(gblob[1] >> 16);
    int       t = 0;
    for (int c = 0; c < C; c++) {
        const int q = next_bit(M + c * W, S + c * WS, W, WS, t);
        if (q < 0) return false;
        t = q + (int)(gblob[4 + 2 * c] & 0xFFu) - 1 + delay;
    }
    return true;
}

// ----------------------------------------------------------------------------------------
// Rank-space walk (long sequences, one warp per item).
//
// The placements of a unit, before the base-pair checks, form a nested structure: the
// candidates of component c+1 below a candidate of component c are a few "greedy" matches
// followed by a contiguous range of the sorted match list of component c+1 (restricted to the
// band window of the checks that tie c+1 to c). Counting them is a backward dynamic programme
// over the match lists (N = placements below each match, PS = prefix sums of N), and the r-th
// placement in depth-first order is found by one binary search per component. A warp therefore
// enumerates 32 consecutive ranks at a time, one per lane, tests the checks, and compacts the
// survivors with ballot/popc: no divergence, no per-lane stacks, and the records of one batch
// are written as one contiguous, coalesced block in canonical order.
// ----------------------------------------------------------------------------------------
struct RankCtx {
    ItemCtx   cx;
    uint16_t* mpos;    // [cap] match positions, lists of the components back to back
    uint32_t* N;       // [cap] placements below each match (components c+1.. only); kHasGreedy: some of its
                       //       children are greedy matches (rare: re-derived with kids_of)
    uint32_t* KID;     // [cap] list range [ja, jb) of the children of each match: ja | jb << 16
    uint32_t* PS;      // [cap + C] per component a_c + 1 prefix sums of N over the thinned matches
    uint16_t* MWP;     // [C][W] matches before each mask word
    uint16_t* lbeg;    // [C + 1] first list entry of each component
    uint32_t* idx;     // this thread's odometer column (stride kThreads): list index | end of the sibling range << 16
    uint32_t* aux;     // this thread's column: start of the sibling range | greedy siblings left (this one included) << 16
};

__device__ __forceinline__ int list_len(const RankCtx& rc, int c) { return (int)rc.lbeg[c + 1] - (int)rc.lbeg[c]; }

// Number of matches of component c at positions < x.
__device__ __forceinline__ int match_rank(const RankCtx& rc, int c, int x)
{
    if (x <= 0) return 0;
    const int w = x >> 5;
    if (w >= rc.cx.W) return list_len(rc, c);
    BSS_ASSERT(c >= 0 && c < rc.cx.C, 4);
    return (int)rc.MWP[c * rc.cx.W + w] + __popc(rc.cx.M[c * rc.cx.W + w] & ((1u << (x & 31)) - 1u));
}

__device__ __forceinline__ bool is_thinned(const RankCtx& rc, int c, int pos)
{
    return !((rc.cx.ovl >> c) & 1u) || test_bit(rc.cx.T + c * rc.cx.W, pos);
}

// Candidates of component c below a candidate of component c-1 placed at p: `nex` greedy
// matches starting at position `efirst` (successive ones one pattern length apart at least),
// then the thinned matches among list entries [ja, jb).
struct Kids {
    int efirst, nex, ja, jb;
};
constexpr uint32_t kHasGreedy  = 0x80000000u;   // N flag: some children of this match are greedy matches
constexpr uint32_t kGreedyOnly = 0x40000000u;   // N flag: not in the thinned set, only ever reached as a greedy match
constexpr uint32_t kCountMask  = 0x3FFFFFFFu;

// A list entry that a parent's range [ja, jb) really contains: thinned, with placements below it.
__device__ __forceinline__ bool in_range_live(uint32_t n) { return ((n & ~kHasGreedy) - 1u) < kCountMask; }

__device__ __forceinline__ Kids kids_of(const RankCtx& rc, int c, int p)
{
    const ItemCtx& cx = rc.cx;
    Kids           kd{-1, 0, 0, 0};
    const int      t = p + (int)(cx.blob[4 + 2 * (c - 1)] & 0xFFu) - 1 + cx.delay;
    if (t < 0 || t >= cx.n) return kd;
    const int kc = (int)(cx.blob[4 + 2 * c] & 0xFFu);
    // an empty pattern at position 0 leaves "no room" (unsigned wrap in the reference) and ends the level
    if (c < cx.C - 1 && kc == 0 && cx.delay == 0 && t == 0) return kd;
    int            lo = t, hi = cx.n;
    const uint32_t range = cx.blob[5 + 2 * c];
    const int      cb = (int)(range & 0xFFu), ce = (int)((range >> 8) & 0xFFu);
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16);
    for (int i = cb; i < ce; i++) {   // windows of the checks that tie c to c-1 or to a fixed position
        const uint32_t ea = chk[2 * i];
        if (ea & kCheckFixed) continue;
        const int aa = unpack_anchor(ea);
        if (aa >= 0 && aa != c - 1) continue;
        const int u = (aa < 0 ? 0 : p) + unpack_off(ea);
        if (u < 0 || u >= cx.n || cx.band < 0) return kd;
        const int ob = unpack_off(chk[2 * i + 1]);
        lo = max(lo, u - cx.band - ob);
        hi = min(hi, u + cx.band - ob);
    }
    if (lo > hi) return kd;
    if ((cx.ovl >> c) & 1u) {
        const uint32_t *M = cx.M + c * cx.W, *SM = cx.SM + c * cx.WS, *T = cx.T + c * cx.W;
        const int       ke = max(1, kc);
        int             q  = next_bit(M, SM, cx.W, cx.WS, t);
        while (q >= 0 && !test_bit(T, q)) {
            if (q > hi) return kd;
            if (q >= lo) {
                if (!kd.nex) kd.efirst = q;
                kd.nex++;
            }
            q = next_bit(M, SM, cx.W, cx.WS, q + ke);
        }
        if (q < 0) return kd;
        lo = max(lo, q);
    }
    kd.ja = match_rank(rc, c, lo);
    kd.jb = max(kd.ja, match_rank(rc, c, hi + 1));
    return kd;
}

// Largest j in [ja, jb) with PS[j] <= target (PS is non-decreasing; entries of weight 0 are never chosen
// because target < PS[jb]).
__device__ __forceinline__ int rank_search(const uint32_t* __restrict__ PS, int ja, int jb, uint32_t target)
{
    BSS_ASSERT(ja < jb && PS[ja] <= target && target < PS[jb], 5);
    int lo = ja, hi = jb;   // invariant: PS[lo] <= target < PS[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (PS[mid] <= target) lo = mid; else hi = mid;
    }
    return lo;
}

// Match lists of the components + word prefix counts. False when they do not fit.
__device__ __forceinline__ bool rank_build_lists(const RankCtx& rc, int cap)
{
    const ItemCtx& cx   = rc.cx;
    const int      lane = threadIdx.x & 31;
    int            base = 0;
    for (int c = 0; c < cx.C; c++) {
        if (lane == 0) rc.lbeg[c] = (uint16_t)base;
        int running = 0;
        for (int w0 = 0; w0 < cx.W; w0 += 32) {
            const int      w   = w0 + lane;
            const uint32_t m   = w < cx.W ? cx.M[c * cx.W + w] : 0u;
            const int      cnt = __popc(m);
            int            inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += up;
            }
            const int total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            if (base + running + total > cap) return false;
            int at = running + inc - cnt;
            if (w < cx.W) rc.MWP[c * cx.W + w] = (uint16_t)at;
            for (uint32_t x = m; x; x &= x - 1) rc.mpos[base + at++] = (uint16_t)((w << 5) + __ffs(x) - 1);
            running += total;
        }
        base += running;
    }
    if (lane == 0) rc.lbeg[cx.C] = (uint16_t)base;
    __syncwarp();
    return true;
}

// Backward dynamic programme. Returns the number of placements before the checks, or
// UINT64_MAX when a count does not fit 30 bits (the caller falls back to the depth-first walk).
__device__ __forceinline__ uint64_t rank_count(const RankCtx& rc)
{
    const ItemCtx& cx   = rc.cx;
    const int      lane = threadIdx.x & 31;
    bool           overflow = false;
    uint64_t       total    = 0;
    for (int c = cx.C - 1; c >= 0; c--) {
        const int       b  = rc.lbeg[c], a = list_len(rc, c);
        uint32_t*       PS = rc.PS + b + c;
        const uint32_t* PSn = rc.PS + rc.lbeg[c + 1] + c + 1;   // next component's
        uint64_t        running = 0;
        if (lane == 0) PS[0] = 0;
        for (int j0 = 0; j0 < a; j0 += 32) {
            const int j = j0 + lane;
            uint64_t  n = 0;
            uint32_t  greedy = 0;
            int       p = 0;
            if (j < a) {
                p = rc.mpos[b + j];
                if (c == cx.C - 1) {
                    n = 1;
                } else {
                    const Kids kd = kids_of(rc, c + 1, p);
                    const int  ke = max(1, (int)(cx.blob[4 + 2 * (c + 1)] & 0xFFu));
                    int        e  = kd.efirst;
                    for (int x = 0; x < kd.nex; x++) {
                        n += rc.N[rc.lbeg[c + 1] + match_rank(rc, c + 1, e)] & kCountMask;
                        e = next_bit(cx.M + (c + 1) * cx.W, cx.SM + (c + 1) * cx.WS, cx.W, cx.WS, e + ke);
                    }
                    n += PSn[kd.jb] - PSn[kd.ja];
                    rc.KID[b + j] = (uint32_t)kd.ja | ((uint32_t)kd.jb << 16);
                    if (kd.nex) greedy = kHasGreedy;
                }
                if (n > kCountMask) overflow = true;
                const bool thinned = is_thinned(rc, c, p);
                rc.N[b + j] = (uint32_t)n | greedy | (thinned ? 0u : kGreedyOnly);
                if (!thinned) n = 0;   // only ever reached as a greedy match, never through a range
            }
            uint64_t inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += up;
            }
            if (j < a) PS[j + 1] = (uint32_t)(running + inc);
            running += __shfl_sync(0xFFFFFFFFu, inc, 31);
            if (running > kCountMask) overflow = true;
        }
        total = running;
        __syncwarp();
    }
    if (__any_sync(0xFFFFFFFFu, overflow)) return ~0ull;
    return total;
}

// First child of list entry gj (component c-1) that has placements below it: sets the odometer of level c.
__device__ __forceinline__ void rank_first_child(const RankCtx& rc, int c, int gj)
{
    const ItemCtx& cx  = rc.cx;
    const int      b   = rc.lbeg[c];
    const uint32_t kid = rc.KID[gj];
    const int      ja = (int)(kid & 0xFFFFu), jb = (int)(kid >> 16);
    if (rc.N[gj] & kHasGreedy) {   // greedy matches come first, in position order
        const Kids kd = kids_of(rc, c, (int)rc.mpos[gj]);
        const int  ke = max(1, (int)(cx.blob[4 + 2 * c] & 0xFFu));
        int        e  = kd.efirst;
        for (int x = kd.nex; x > 0; x--) {
            const int je = match_rank(rc, c, e);
            if (rc.N[b + je] & kCountMask) {
                rc.idx[c * kThreads]   = (uint32_t)je | ((uint32_t)jb << 16);
                rc.aux[c * kThreads]   = (uint32_t)ja | ((uint32_t)x << 16);
                cx.stack[c * kThreads] = (uint32_t)e;
                return;
            }
            e = next_bit(cx.M + c * cx.W, cx.SM + c * cx.WS, cx.W, cx.WS, e + ke);
        }
    }
    int j = ja;
    while (j < jb && !in_range_live(rc.N[b + j])) j++;
    BSS_ASSERT(j < jb && b + j < (int)rc.lbeg[c + 1], 6);
    rc.idx[c * kThreads]   = (uint32_t)j | ((uint32_t)jb << 16);
    rc.aux[c * kThreads]   = (uint32_t)ja;
    cx.stack[c * kThreads] = rc.mpos[b + j];
}

// Odometer state (list indices + positions of every component) of the placement of depth-first rank r.
__device__ __forceinline__ void rank_unrank(const RankCtx& rc, uint32_t r)
{
    const ItemCtx& cx = rc.cx;
    int            gj;   // list entry (all components back to back) of the component placed last
    {
        const int a0 = list_len(rc, 0);
        const int j  = rank_search(rc.PS, 0, a0, r);
        r -= rc.PS[j];
        gj          = j;
        rc.idx[0]   = (uint32_t)j | ((uint32_t)a0 << 16);
        rc.aux[0]   = 0;
        cx.stack[0] = rc.mpos[j];
    }
    for (int c = 1; c < cx.C; c++) {
        const int       b   = rc.lbeg[c];
        const uint32_t* PS  = rc.PS + b + c;
        const uint32_t  kid = rc.KID[gj];
        const int       ja = (int)(kid & 0xFFFFu), jb = (int)(kid >> 16);
        int             j = -1;
        uint32_t        aux = (uint32_t)ja;
        if (rc.N[gj] & kHasGreedy) {
            const Kids kd = kids_of(rc, c, (int)rc.mpos[gj]);
            const int  ke = max(1, (int)(cx.blob[4 + 2 * c] & 0xFFu));
            int        e  = kd.efirst;
            for (int x = kd.nex; x > 0; x--) {
                const int      je = match_rank(rc, c, e);
                const uint32_t n  = rc.N[b + je] & kCountMask;
                if (r < n) { j = je; aux |= (uint32_t)x << 16; break; }
                r -= n;
                e = next_bit(cx.M + c * cx.W, cx.SM + c * cx.WS, cx.W, cx.WS, e + ke);
            }
        }
        if (j < 0) {
            const uint32_t base = PS[ja];
            j = rank_search(PS, ja, jb, base + r);
            r -= PS[j] - base;
        }
        gj                     = b + j;
        BSS_ASSERT(j >= 0 && gj < (int)rc.lbeg[c + 1], 9);
        rc.idx[c * kThreads]   = (uint32_t)j | ((uint32_t)jb << 16);
        rc.aux[c * kThreads]   = aux;
        cx.stack[c * kThreads] = rc.mpos[gj];
    }
    BSS_ASSERT(r == 0, 10);
}

// Steps the odometer to the next placement in depth-first order that differs at level `from` or
// above (the caller knows there is one). Returns the first level whose component moved.
__device__ __forceinline__ int rank_advance(const RankCtx& rc, int from)
{
    const ItemCtx& cx = rc.cx;
    int            c  = from, j = 0;
    for (;; c--) {
        BSS_ASSERT(c >= 0, 7);
        const int      b   = rc.lbeg[c];
        const uint32_t idx = rc.idx[c * kThreads], aux = rc.aux[c * kThreads];
        const int      jb  = (int)(idx >> 16);
        int            xs  = (int)(aux >> 16);
        j                  = (int)(idx & 0xFFFFu);
        if (xs > 1) {   // further greedy siblings
            const int ke = max(1, (int)(cx.blob[4 + 2 * c] & 0xFFu));
            int       e  = (int)cx.stack[c * kThreads];
            bool      hit = false;
            while (xs > 1) {
                e = next_bit(cx.M + c * cx.W, cx.SM + c * cx.WS, cx.W, cx.WS, e + ke);
                xs--;
                j = match_rank(rc, c, e);
                if (rc.N[b + j] & kCountMask) { hit = true; break; }
            }
            if (hit) {
                rc.idx[c * kThreads]   = (uint32_t)j | ((uint32_t)jb << 16);
                rc.aux[c * kThreads]   = (aux & 0xFFFFu) | ((uint32_t)xs << 16);
                cx.stack[c * kThreads] = (uint32_t)e;
                break;
            }
        }
        if (xs >= 1) {   // the greedy siblings are used up: on to the range
            j = (int)(aux & 0xFFFFu);
            rc.aux[c * kThreads] = aux & 0xFFFFu;
        } else {
            j++;
        }
        while (j < jb && !in_range_live(rc.N[b + j])) j++;
        if (j < jb) {
            rc.idx[c * kThreads]   = (uint32_t)j | ((uint32_t)jb << 16);
            cx.stack[c * kThreads] = rc.mpos[b + j];
            break;
        }
    }
    for (int d = c + 1; d < cx.C; d++) {
        rank_first_child(rc, d, (int)rc.lbeg[d - 1] + j);
        j = (int)(rc.idx[d * kThreads] & 0xFFFFu);
    }
    return c;
}

constexpr int kHoist = 4;   // leaf-level checks kept in registers by the fast leaf loop

// Leaves (placements of the last component) below the current upper placement: list entries
// [j, jb), at most `quota` of them. Every leaf check is an ordered pair (u fixed by the upper
// placement, v = leaf position + offset): the word index of row u is hoisted, the test is one
// load and one funnel shift. Returns the number of leaves visited; `fill` counts the sites.
template <int NL, bool EMIT>
__device__ __forceinline__ uint32_t leaf_run(const RankCtx& rc, int& j, int jb, uint32_t quota, uint32_t module, uint32_t* slot_base,
                                             uint32_t& fill)
{
    const ItemCtx&  cx = rc.cx;
    const int       C = cx.C, L = C - 1, R = C + 1;
    const int       bL = rc.lbeg[L];
    const uint32_t  range = cx.blob[5 + 2 * L];
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16) + 2 * (range & 0xFFu);
    uint32_t        row[NL > 0 ? NL : 1], off[NL > 0 ? NL : 1];
#pragma unroll
    for (int i = 0; i < NL; i++) {
        const uint32_t ea = chk[2 * i];
        row[i] = (uint32_t)(placed(cx, unpack_anchor(ea)) + unpack_off(ea)) * (uint32_t)cx.mstride;
        off[i] = (uint32_t)unpack_off(chk[2 * i + 1]);
    }
    const bool sparse = (cx.ovl >> L) & 1u;   // some list entries are not in the thinned set
    uint32_t   done   = 0;
    // Three-component units write 16-byte records: when the lane's output range starts on a 16-byte boundary (always, in a library
    // of such units) a record is ONE vector store of registers — the upper placement is fixed for the whole run of leaves.
    const bool vec4 = EMIT && R == 4 && (reinterpret_cast<uintptr_t>(slot_base) & 15u) == 0;
    uint32_t   up0 = 0, up1 = 0;
    if (vec4) {
        up0 = cx.stack[0];
        up1 = cx.stack[kThreads];
    }
    for (; j < jb && done < quota; j++) {
        if (sparse && !in_range_live(rc.N[bL + j])) continue;
        const uint32_t p  = rc.mpos[bL + j];
        uint32_t       ok = 1u;
#pragma unroll
        for (int i = 0; i < NL; i++) {
            const uint32_t v = p + off[i];
            BSS_ASSERT(v < (uint32_t)cx.n && row[i] / (uint32_t)cx.mstride < v, 8);
            ok &= __funnelshift_r(__ldg(cx.mask + row[i] + (v >> 5)), 0u, v);
        }
        done++;
        if (ok) {
            if (EMIT) {
                if (vec4) {
                    reinterpret_cast<uint4*>(slot_base)[fill] = make_uint4(module, up0, up1, p);
                } else {
                    uint32_t* rec = slot_base + fill * R;
                    rec[0]        = module;
                    for (int c = 0; c < L; c++) rec[1 + c] = cx.stack[c * kThreads];
                    rec[C] = p;
                }
            }
            fill++;
        }
    }
    if (sparse)
        while (j < jb && !in_range_live(rc.N[bL + j])) j++;
    return done;
}

// Ranks left in the subtree of the current entry of level `bl` (the current placement included),
// or 0 when greedy matches are involved on the way down (their counts are not in the prefix sums).
// The subtree holds N ranks; those already behind are the placements below the earlier siblings
// of every level below `bl`. Leaves the levels in between on their last sibling, so that an
// advance from the level above the leaves carries up to `bl`.
__device__ __forceinline__ uint32_t subtree_skip(uint32_t* idx, const uint32_t* aux, const uint32_t* N, const uint32_t* PS, const uint16_t* lbeg,
                                              int bl, int L)
{
    uint32_t used = 0;
    for (int c = bl + 1; c <= L; c++) {
        const uint32_t a      = aux[c * kThreads];
        const int      parent = (int)lbeg[c - 1] + (int)(idx[(c - 1) * kThreads] & 0xFFFFu);
        if ((a >> 16) != 0 || (N[parent] & kHasGreedy)) return 0;
        const uint32_t* PSc = PS + lbeg[c] + c;
        used += PSc[idx[c * kThreads] & 0xFFFFu] - PSc[a & 0xFFFFu];
    }
    const uint32_t below = N[(int)lbeg[bl] + (int)(idx[bl * kThreads] & 0xFFFFu)] & kCountMask;
    BSS_ASSERT(used < below, 12);
    for (int c = bl + 1; c < L; c++) {   // last sibling of every level in between
        const uint32_t end = idx[c * kThreads] >> 16;
        idx[c * kThreads]  = (end - 1u) | (end << 16);
    }
    return below - used;
}

// Walks `cnt` consecutive ranks from `first` with this lane's odometer and returns how many
// placements pass their checks. The odometer only steps the upper components; under each of
// their placements the leaves are a run of consecutive list entries (leaf_run). When EMIT, the
// records go straight to `dst`, this lane's part of the output: every lane streams through its
// own contiguous range, and L2 merges the partial sectors before they reach HBM.
template <bool EMIT, bool PRUNE>
__device__ __forceinline__ uint32_t rank_chunk(const RankCtx& rc, uint32_t first, uint32_t cnt, int leaf_checks, uint32_t module,
                                               uint32_t* __restrict__ dst)
{
    const ItemCtx& cx = rc.cx;
    const int      C = cx.C, L = C - 1, R = C + 1;
    const int      bL = rc.lbeg[L];
    uint32_t       npass = 0, remaining = cnt;
    int            ok_upper = 0;   // checks of levels [0, ok_upper) hold for the current upper placement
    if (cnt) rank_unrank(rc, first);
    while (remaining > 0) {
        while (ok_upper < L && checks_pass(cx, ok_upper)) ok_upper++;
        const uint32_t idx = rc.idx[L * kThreads];
        int            j = (int)(idx & 0xFFFFu);
        const int      jb = (int)(idx >> 16);
        const bool     in_greedy = (rc.aux[L * kThreads] >> 16) != 0;
        if (leaf_checks < 0 || in_greedy) {
            // one placement at a time, every check through the generic test
            if (ok_upper == L && checks_pass(cx, L)) {
                if (EMIT) {
                    uint32_t* rec = dst + (size_t)npass * R;
                    rec[0]        = module;
                    for (int c = 0; c < C; c++) rec[1 + c] = cx.stack[c * kThreads];
                }
                npass++;
            }
            if (--remaining) ok_upper = min(ok_upper, rank_advance(rc, L));
            continue;
        }
        if (ok_upper < L) {
            // A check of level ok_upper fails: no placement below the current entry of that level is a
            // site. Skip the rest of its subtree in one step (subtree_skip).
            const int       bl = ok_upper;
            const uint32_t* PSL = rc.PS + bL + L;
            uint32_t        skip = PSL[jb] - PSL[j];   // at least the leaves below the current upper placement
            if (PRUNE && bl < L - 1) {   // (heavy items only: the ordinary passes are sensitive to the extra code)
                const uint32_t whole = subtree_skip(rc.idx, rc.aux, rc.N, rc.PS, rc.lbeg, bl, L);
                if (whole) skip = whole;
            }
            remaining -= min(remaining, skip);
            j = jb;
        } else {
            uint32_t* base = EMIT ? dst + (size_t)npass * R : nullptr;
            uint32_t  done, fill = 0;
            switch (leaf_checks) {
            case 0: done = leaf_run<0, EMIT>(rc, j, jb, remaining, module, base, fill); break;
            case 1: done = leaf_run<1, EMIT>(rc, j, jb, remaining, module, base, fill); break;
            case 2: done = leaf_run<2, EMIT>(rc, j, jb, remaining, module, base, fill); break;
            case 3: done = leaf_run<3, EMIT>(rc, j, jb, remaining, module, base, fill); break;
            default: done = leaf_run<4, EMIT>(rc, j, jb, remaining, module, base, fill); break;
            }
            remaining -= done;
            npass += fill;
        }
        if (remaining) {
            if (j >= jb) {
                ok_upper = min(ok_upper, rank_advance(rc, L - 1));
            } else {
                rc.idx[L * kThreads]   = (uint32_t)j | ((uint32_t)jb << 16);
                cx.stack[L * kThreads] = rc.mpos[bL + j];
            }
        }
    }
    return npass;
}

// Enumerates ranks [first, first + total): lane l takes the l-th of 32 equal runs of consecutive ranks, so
// the lanes of a warp carry the same load whatever the shape of the placement tree. Returns the
// number of sites and, in `lane_sites`, those of this lane's run. The count pass records the
// latter; their prefix gives every lane its place in the output, so the emit pass is a single
// walk that writes the records where they belong (when `known` is missing it counts first):
// canonical order, contiguous block per item.
template <bool EMIT, bool PRUNE>
__device__ __forceinline__ uint64_t rank_walk(const RankCtx& rc, uint32_t first, uint32_t total, uint32_t module, uint32_t* __restrict__ out,
                                              const uint32_t* __restrict__ known, uint32_t& lane_sites)
{
    const ItemCtx& cx   = rc.cx;
    const uint32_t lane = threadIdx.x & 31;
    const int      R    = cx.C + 1;
    const uint32_t per  = (total + 31u) / 32u;
    const uint32_t skip = lane * per;
    const uint32_t mine = first + skip;
    const uint32_t cnt  = skip < total ? min(per, total - skip) : 0u;
    // the fast leaf loop needs every leaf-level check to be an ordered pair with an earlier component
    int leaf_checks = -1;
    if (cx.C >= 2) {
        const uint32_t  range = cx.blob[5 + 2 * (cx.C - 1)];
        const int       cb = (int)(range & 0xFFu), ce = (int)((range >> 8) & 0xFFu);
        const uint32_t* chk = cx.blob + (cx.blob[3] >> 16);
        leaf_checks = ce - cb <= kHoist ? ce - cb : -1;
        for (int i = cb; i < ce; i++)
            if (!(chk[2 * i] & kCheckOrdered)) leaf_checks = -1;
    }
    BSS_PHASE_BEGIN();
    // sites of this lane's run: recorded by the count pass, or counted now
    lane_sites    = (EMIT && known) ? known[lane] : rank_chunk<false, PRUNE>(rc, mine, cnt, leaf_checks, module, nullptr);
    uint32_t incl = lane_sites;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += up;
    }
    const uint32_t sites = __shfl_sync(0xFFFFFFFFu, incl, 31);
    BSS_PHASE(9);
    if (EMIT && sites) {
        rank_chunk<true, PRUNE>(rc, mine, cnt, leaf_checks, module, out + (size_t)(incl - lane_sites) * R);
        __syncwarp();
        BSS_PHASE(10);
    }
    return sites;
}

struct Smem {
    uint32_t* blob;      // [groups][kBlobStage]
    uint32_t* M;         // [groups][cmax][W]
    uint32_t* T;         // [groups][cmax][W]
    uint32_t* SM;        // [groups][cmax][WS]
    uint32_t* ST;        // [groups][cmax][WS]
    uint32_t* stack;     // [cmax][kThreads]
    uint32_t* idx;       // [2][cmax][kThreads] odometer columns of the rank-space walk
    uint32_t* rank;      // [groups][rank_words] rank-space scratch (32-lane groups only)
    uint16_t* surv;      // [units_per_chunk] count pass: units of the chunk the k-mer prefilter left standing
};

// Rank-space scratch of one group, in words: N, PS, staging, then the u16 arrays.
__host__ __device__ inline uint32_t rank_words(uint32_t cmax, uint32_t W, uint32_t cap, bool emit)
{
    if (cap == 0) return 0;
    (void)emit;
    return 2 * cap + (cap + cmax + 1) + (cap + 1) / 2 + (cmax * W + 1) / 2 + (cmax + 2) / 2;
}

__host__ __device__ inline uint32_t smem_words(uint32_t cmax, uint32_t W, uint32_t WS, uint32_t groups,
                                               uint32_t units_per_chunk, uint32_t cap, bool emit)
{
    uint32_t words = groups * (kBlobStage + 2 * cmax * (W + WS) + rank_words(cmax, W, cap, emit));
    words += (cap ? 3 : 1) * cmax * kThreads;
    if (!emit && groups > kThreads / 32) words += (units_per_chunk + 1) / 2;   // survivors of the chunk (narrow groups)
    return words;
}

template <int G, bool EMIT>
__device__ __forceinline__ Smem carve(uint32_t* base, const DevLibrary& lib, const Plan& plan)
{
    const uint32_t groups = kThreads / G;
    Smem           s;
    s.blob  = base;               base += groups * kBlobStage;
    s.M     = base;               base += groups * lib.cmax * plan.W;
    s.T     = base;               base += groups * lib.cmax * plan.W;
    s.SM    = base;               base += groups * lib.cmax * plan.WS;
    s.ST    = base;               base += groups * lib.cmax * plan.WS;
    s.stack = base;               base += lib.cmax * kThreads;
    s.idx   = base;               base += (plan.list_cap ? 2 : 0) * lib.cmax * kThreads;
    s.rank  = base;               base += groups * rank_words(lib.cmax, plan.W, plan.list_cap, EMIT);
    s.surv  = reinterpret_cast<uint16_t*>(base);
    return s;
}

// The sequence-dependent inputs of an item.
struct SeqCtx {
    const uint32_t* pair_mask;
    const uint32_t* occ;
    uint64_t        item0;   // index of (this sequence, unit 0)
    int             n, band, Wq;
};

__device__ __forceinline__ SeqCtx load_seq(const DevLibrary& lib, const DevQuery& query, uint32_t qi)
{
    SeqCtx sq;
    sq.n         = (int)(query.seq_begin[qi + 1] - query.seq_begin[qi]);
    sq.pair_mask = query.masks + query.mask_begin[qi];
    sq.band      = query.band[qi];
    sq.occ       = query.occ + (size_t)qi * query.occ_rows * query.occ_stride;
    sq.item0     = (uint64_t)qi * lib.n_units;
    sq.Wq        = (sq.n + 32) >> 5;
    return sq;
}

// Block-wide exclusive prefix of one u64 per thread (any whole number of warps up to 32); `total` = block sum.
__device__ __forceinline__ uint64_t block_exclusive(uint64_t v, uint64_t* warp_sums, uint64_t& total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    uint64_t       wt;
    const uint64_t in_warp = Group<32>::exclusive(v, wt);
    __syncthreads();   // warp_sums may still be read from the previous call
    if (lane == 31) warp_sums[warp] = wt;
    __syncthreads();
    if (warp == 0) {
        uint64_t       all;
        const uint64_t ex = Group<32>::exclusive(lane < n_warps ? warp_sums[lane] : 0ull, all);
        warp_sums[lane]   = ex;
        if (lane == 0) warp_sums[32] = all;
    }
    __syncthreads();
    total = warp_sums[32];
    return warp_sums[warp] + in_warp;
}

// Exclusive prefix of the per-chunk word counts, totals, and per-sequence word offsets: one CTA. The chunk totals were
// accumulated with atomics by other SMs: they are read past L1.
// `seq_out` (may be null): the caller's copy of the per-sequence offsets (its last entry is the record-word count an
// exchange behind the search reads), written here instead of by a copy kernel of its own.
__device__ __forceinline__ void scan_chunks(const Plan& plan, const Work& work, uint32_t n_seqs, uint64_t* warp_sums, uint64_t* __restrict__ seq_out)
{
    uint64_t carry_words = 0, carry_sites = 0;   // identical in every thread
    for (uint32_t base = 0; base < plan.n_chunks; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t w = i < plan.n_chunks ? __ldcg(work.chunk_words + i) : 0;
        const uint64_t s = i < plan.n_chunks ? __ldcg(work.chunk_sites + i) : 0;
        uint64_t       tw, ts;
        const uint64_t ew = block_exclusive(w, warp_sums, tw);
        block_exclusive(s, warp_sums, ts);
        if (i < plan.n_chunks) {
            work.chunk_words[i] = carry_words + ew;
            if (i % plan.chunks_per_seq == 0) {
                work.seq_word_begin[i / plan.chunks_per_seq] = carry_words + ew;
                if (seq_out) seq_out[i / plan.chunks_per_seq] = carry_words + ew;
            }
        }
        carry_words += tw;
        carry_sites += ts;
    }
    if (threadIdx.x == 0) {
        work.chunk_words[plan.n_chunks] = carry_words;
        work.seq_word_begin[n_seqs]     = carry_words;
        if (seq_out) seq_out[n_seqs] = carry_words;
        work.totals[0]                  = carry_words;
        work.totals[1]                  = carry_sites;
    }
}

__global__ void __launch_bounds__(1024) scan_kernel(const Plan plan, const Work work, uint32_t n_seqs, uint64_t* __restrict__ seq_out)
{
    __shared__ uint64_t warp_sums[33];
    scan_chunks(plan, work, n_seqs, warp_sums, seq_out);
}

// The head of a search: totals, chunk totals and per-sequence offsets zeroed by the threads [0, stride) of a grid — one kernel
// instead of three memsets, or no kernel of its own at all when the pattern table is built first (pattern_kernel does it).
__device__ __forceinline__ void reset_work(const Work& work, uint32_t n_chunks, uint32_t n_seqs, uint32_t i, uint32_t stride)
{
    if (i < kTotalsWords) work.totals[i] = 0;
    for (uint32_t j = i; j < 2 * (n_chunks + 1); j += stride) work.chunk_sites[j] = 0;   // (chunk_words follows chunk_sites)
    for (uint32_t j = i; j <= n_seqs; j += stride) work.seq_word_begin[j] = 0;
}

__global__ void __launch_bounds__(256) reset_kernel(const Work work, uint32_t n_chunks, uint32_t n_seqs)
{
    reset_work(work, n_chunks, n_seqs, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

#include "bss_window.cuh"

// Count pass (EMIT = false): one CTA per chunk = (sequence, contiguous unit range); the groups
// of the CTA draw units from the chunk's counter, write the site count of every item, add the
// items that have sites to the alive list, and the CTA leaves the chunk's totals behind.
// Emit pass (EMIT = true): the groups walk the alive list (any order: every item knows its own
// place in the output from the scanned chunk totals and the counts of the chunk's earlier items).
// Heavy items (MODE 2, 3; 32-lane groups): an item whose placements before the checks number
// plan.heavy_ranks or more is not walked by the warp that finds it in the count pass but
// parked on the heavy list; its rank space is cut into blocks and every (item, block) task is
// taken by a warp of these passes, which re-derives the item's lists and counts and walks its
// block only. One monster unit is thereby spread over the whole GPU instead of one warp.
constexpr int kModeCount = 0, kModeEmit = 1, kModeHeavyCount = 2, kModeHeavyEmit = 3;
constexpr uint32_t kHeavyCap = 1024, kHeavyMaxBlocks = 256;

__device__ __forceinline__ uint32_t heavy_blocks(uint32_t ranks, uint32_t block)
{
    const uint32_t b = (ranks + block - 1) / block;
    return b > kHeavyMaxBlocks ? kHeavyMaxBlocks : b;
}

template <int G, int MODE>
__global__ void __launch_bounds__(kThreads, BSS_MIN_CTAS) search_kernel(const DevLibrary lib, const DevQuery query, const Plan plan,
                                                             const Work work, uint32_t* __restrict__ records, uint32_t n_alive,
                                                             uint64_t capacity_words)
{
    constexpr bool EMIT = (MODE & 1) != 0, HEAVY = MODE >= 2;
    if (HEAVY && work.totals[4] == 0) return;   // the usual case: nothing was parked
    if (MODE == kModeCount && plan.window && work.totals[7] == 0) return;   // every item went to the window pass or ended in route_kernel
    if (EMIT) {
        // Emit passes queued behind the count pass without a host round trip: the sizes come from
        // device memory, and nothing is written when the records would not fit the caller's buffer.
        if (work.totals[0] > capacity_words) return;
        if (n_alive == 0xFFFFFFFFu) n_alive = (uint32_t)work.totals[3];
    }
    static_assert(!HEAVY || G == 32, "heavy items are a 32-lane-group matter");
    using Gr = Group<G>;
    extern __shared__ __align__(16) uint32_t smem_raw[];
    __shared__ uint32_t next_unit, n_survivors;
    constexpr uint32_t groups = kThreads / G;
    const int          W = (int)plan.W, WS = (int)plan.WS, WP = (int)query.occ_stride;

    Smem           sm     = carve<G, EMIT>(smem_raw, lib, plan);
    const uint32_t gid    = threadIdx.x / G;
    uint32_t*      g_blob = sm.blob + gid * kBlobStage;
    uint32_t*      g_M    = sm.M + gid * lib.cmax * W;
    uint32_t*      g_T    = sm.T + gid * lib.cmax * W;
    uint32_t*      g_SM   = sm.SM + gid * lib.cmax * WS;
    uint32_t*      g_ST   = sm.ST + gid * lib.cmax * WS;

    // count pass: this CTA's chunk
    const uint32_t chunk = blockIdx.x;
    uint32_t       u1 = 0, u0 = 0, n_draw = 0;
    bool           compact = false;
    SeqCtx         sq{};
    if (MODE == kModeCount) {
        u0 = (chunk % plan.chunks_per_seq) * plan.units_per_chunk;
        u1                = min(lib.n_units, u0 + plan.units_per_chunk);
        sq                = load_seq(lib, query, chunk / plan.chunks_per_seq);
        if constexpr (G < 32) {
            // Narrow groups (batches of short sequences, large chunks): only the units the k-mer prefilter
            // left standing are drawn. (Kept out of the 32-lane instantiation, which is sensitive to extra code.)
            n_draw  = u1 - u0;
            compact = work.prefilter != nullptr;
            if (compact) {
                if (threadIdx.x < 32) {
                    uint32_t n = 0;
                    for (uint32_t b = u0; b < u1; b += 32) {
                        const uint32_t v = b + threadIdx.x;
                        const uint64_t item = sq.item0 + v;
                        const bool     on = v < u1 && ((__ldg(work.prefilter + (item >> 5)) >> (item & 31)) & 1u);
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, on);
                        if (on) sm.surv[n + __popc(bal & ((1u << threadIdx.x) - 1u))] = (uint16_t)(v - u0);
                        n += __popc(bal);
                    }
                    if (threadIdx.x == 0) n_survivors = n;
                }
                __syncthreads();
                n_draw = n_survivors;
            }
            if (threadIdx.x == 0) next_unit = 0;
        } else {
            if (threadIdx.x == 0) next_unit = u0;
        }
        __syncthreads();
    }
    uint64_t my_sites = 0, my_words = 0;   // accumulated by lane 0 of each group (count pass)
    uint32_t slot = blockIdx.x * groups + gid;   // emit pass: position in the alive list; heavy passes: task

    // heavy passes: tasks = blocks of the heavy items' rank spaces, numbered item after item
    __shared__ uint32_t task_begin[HEAVY ? kHeavyCap + 1 : 1];
    uint32_t            n_heavy = 0, n_tasks = 0;
    if (HEAVY) {
        n_heavy = (uint32_t)min((unsigned long long)work.totals[4], (unsigned long long)kHeavyCap);
        if (threadIdx.x < 32) {
            uint32_t carry = 0;
            for (uint32_t b = 0; b < n_heavy; b += 32) {
                const uint32_t h = b + threadIdx.x;
                uint64_t       total;
                const uint64_t ex = Group<32>::exclusive(h < n_heavy ? heavy_blocks(work.heavy_ranks[h], plan.heavy_block) : 0u, total);
                if (h < n_heavy) task_begin[h] = carry + (uint32_t)ex;
                carry += (uint32_t)total;
            }
            if (threadIdx.x == 0) task_begin[n_heavy] = carry;
        }
        __syncthreads();
        n_tasks = task_begin[n_heavy];
    }

    BSS_PHASE_BEGIN();
    for (;; slot += gridDim.x * groups) {
        BSS_PHASE(7);
        uint32_t u = 0;
        uint64_t out_words = 0;   // emit pass: where the item's records start
        uint32_t rank0 = 0, rank_n = 0, chunk_of_item = 0, heavy_r = 0;   // heavy passes: this task's block of ranks
        if (HEAVY) {
            if (slot >= n_tasks) break;
            int lo = 0, hi = (int)n_heavy;   // item whose tasks contain task `slot`
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (task_begin[mid] <= slot) lo = mid; else hi = mid;
            }
            const uint32_t item = work.heavy_items[lo], ranks = work.heavy_ranks[lo];
            const uint32_t nb = heavy_blocks(ranks, plan.heavy_block), blk = slot - task_begin[lo], per = (ranks + nb - 1) / nb;
            const uint32_t qi = item / lib.n_units;
            u                 = item - qi * lib.n_units;
            sq                = load_seq(lib, query, qi);
            heavy_r           = ranks;
            rank0             = blk * per;
            rank_n            = rank0 < ranks ? min(per, ranks - rank0) : 0u;
            chunk_of_item     = qi * plan.chunks_per_seq + u / plan.units_per_chunk;
            if (EMIT) {
                const uint32_t v0 = (u / plan.units_per_chunk) * plan.units_per_chunk;
                uint64_t       before = 0;
                for (uint32_t v = v0 + Gr::lane(); v < u; v += G)
                    before += (uint64_t)work.counts[sq.item0 + v] * __ldg(lib.unit_rlen + v);
                for (uint32_t t = Gr::lane(); t < blk; t += G)   // sites of the item's earlier blocks
                    before += (uint64_t)work.task_sites[task_begin[lo] + t] * __ldg(lib.unit_rlen + u);
                out_words = work.chunk_words[chunk_of_item] + Gr::sum(before);
            }
        } else if (!EMIT) {
            if (Gr::lane() == 0) u = atomicAdd(&next_unit, 1u);
            u = Gr::first(u);
            if constexpr (G < 32) {
                if (u >= n_draw) break;
                u = u0 + (compact ? (uint32_t)sm.surv[u] : u);
            } else {
                if (u >= u1) break;
                if (work.prefilter) {   // ruled out by the k-mer presence filter (its count is already written)
                    const uint64_t item = sq.item0 + u;
                    if (!((__ldg(work.prefilter + (item >> 5)) >> (item & 31)) & 1u)) continue;
                }
            }
        } else {
            if (slot >= n_alive) break;
            const uint32_t item = work.alive[slot];
            const uint32_t qi   = item / lib.n_units;
            u                   = item - qi * lib.n_units;
            sq                  = load_seq(lib, query, qi);
            // records of the chunk's earlier items
            const uint32_t c0 = u / plan.units_per_chunk, v0 = c0 * plan.units_per_chunk;
            uint64_t       before = 0;
            for (uint32_t v = v0 + Gr::lane(); v < u; v += G)
                before += (uint64_t)work.counts[sq.item0 + v] * __ldg(lib.unit_rlen + v);
            out_words = work.chunk_words[qi * plan.chunks_per_seq + c0] + Gr::sum(before);
        }
        const int       n = sq.n;
        const uint32_t* occ = sq.occ;

        // stage the unit descriptor
        const uint32_t  boff  = __ldg(lib.unit_off + u);
        const uint32_t  blen  = __ldg(lib.unit_off + u + 1) - boff;
        const uint32_t* gblob = lib.blob + boff;
        const uint32_t* blob  = gblob;
        if (G == 32 && blen <= kBlobStage) {   // narrow groups read the descriptor in place (L1): staging it with 4-16 lanes costs more than it saves
            for (uint32_t i = Gr::lane(); i < blen; i += G) g_blob[i] = __ldg(gblob + i);
            blob = g_blob;
            Gr::sync();
        }
        const uint32_t module = blob[0] + lib.module_base;
        const int      C      = (int)(blob[1] & 0xFFu);
        const int      delay  = (int)(blob[1] >> 16);
        const uint32_t gate   = blob[2];

        BSS_PHASE(0);
        bool alive = true;
        if (MODE == kModeCount && gate != kNoGate) {   // items of the later passes have already passed their gate
            const uint32_t* gb = lib.blob + __ldg(lib.unit_off + gate);
            alive              = gate_open<G>(gb, occ, WP, g_M, g_SM, W, WS, n);
            Gr::sync();
        }
        // An empty first pattern at position 0 with delay 0 leaves "no room" (unsigned wrap in
        // the reference, cppsrc/Motif.cpp:461) and ends the enumeration before it starts.
        if (C > 1 && (blob[4] & 0xFFu) == 0 && delay == 0) alive = false;
        if (alive) alive = match_components<G>(blob, occ, WP, g_M, g_SM, W, WS, n);

        BSS_PHASE(1);
        uint64_t item_sites = 0;
        uint32_t lane_sites = 0;
        bool     ranked = false, parked = false;
        if (alive) {
            ItemCtx cx;
            cx.blob = blob; cx.M = g_M; cx.T = g_T; cx.SM = g_SM; cx.ST = g_ST; cx.mask = sq.pair_mask;
            cx.stack = sm.stack + threadIdx.x;
            cx.ovl   = thin_components<G>(blob, g_M, g_SM, g_T, g_ST, W, WS);
            cx.W = W; cx.WS = WS; cx.n = n; cx.mstride = (n + 31) >> 5; cx.C = C; cx.delay = delay; cx.band = sq.band;

            BSS_PHASE(2);
            if constexpr (G == 32) {
                if (plan.list_cap) {
                    const uint32_t cap = plan.list_cap;
                    uint32_t*      rw  = sm.rank + gid * rank_words(lib.cmax, plan.W, cap, EMIT);
                    RankCtx        rc;
                    rc.cx   = cx;
                    rc.N    = rw;                        rw += cap;
                    rc.KID  = rw;                        rw += cap;
                    rc.PS   = rw;                        rw += cap + lib.cmax + 1;
                    rc.mpos = reinterpret_cast<uint16_t*>(rw);   rw += (cap + 1) / 2;
                    rc.MWP  = reinterpret_cast<uint16_t*>(rw);   rw += (lib.cmax * plan.W + 1) / 2;
                    rc.lbeg = reinterpret_cast<uint16_t*>(rw);
                    rc.idx  = sm.idx + threadIdx.x;
                    rc.aux  = sm.idx + lib.cmax * kThreads + threadIdx.x;
                    const bool listed = rank_build_lists(rc, (int)cap);
                    BSS_PHASE(3);
                    if (listed) {
                        const uint64_t total = rank_count(rc);
                        BSS_PHASE(4);
                        if (total != ~0ull) {
                            ranked = true;
                            if (HEAVY) {
                                BSS_ASSERT(total == heavy_r, 11);
                                const uint32_t* known = EMIT ? work.task_lane_counts + (size_t)slot * 32 : nullptr;
                                if (rank_n) item_sites = rank_walk<EMIT, true>(rc, rank0, rank_n, module, EMIT ? records + out_words : nullptr, known, lane_sites);
                            } else if (MODE == kModeCount && total >= plan.heavy_ranks) {
                                uint32_t h = 0;   // park it: the heavy passes walk it block by block
                                if (Gr::lane() == 0) h = (uint32_t)atomicAdd((unsigned long long*)(work.totals + 4), 1ull);
                                h = Gr::first(h);
                                if (h < kHeavyCap) {
                                    if (Gr::lane() == 0) { work.heavy_items[h] = (uint32_t)(sq.item0 + u); work.heavy_ranks[h] = (uint32_t)total; }
                                    parked = true;
                                }
                            }
                            if (!HEAVY && !parked && total) {
                                const uint32_t* known = (EMIT && slot < work.lane_cap) ? work.lane_counts + (size_t)slot * 32 : nullptr;
                                item_sites = rank_walk<EMIT, false>(rc, 0u, (uint32_t)total, module, EMIT ? records + out_words : nullptr, known, lane_sites);
                                BSS_PHASE(5);
                            }
                        }
                    }
                }
            }
            if (!ranked) {
                BSS_PHASE(7);
                // Component 0: its chain from position 0 starts on a run start, i.e. it is the thinned set.
                const uint32_t* first = (cx.ovl & 1u) ? g_T : g_M;
                if (!EMIT) {
                    uint64_t mine = 0;
                    for (int w = Gr::lane(); w < sq.Wq; w += G)
                        for (uint32_t x = first[w]; x; x &= x - 1) mine += walk_subtree<false>(cx, (w << 5) + __ffs(x) - 1, module, nullptr);
                    item_sites = Gr::sum(mine);
                } else {
                    // sites below each word of component 0, then their exclusive prefix
                    uint64_t carry = 0;
                    for (int w0 = 0; w0 < sq.Wq; w0 += G) {
                        const int w    = w0 + (int)Gr::lane();
                        uint64_t  mine = 0;
                        uint32_t  x    = 0;
                        if (w < sq.Wq) {
                            x = first[w];
                            for (uint32_t y = x; y; y &= y - 1) mine += walk_subtree<false>(cx, (w << 5) + __ffs(y) - 1, module, nullptr);
                        }
                        uint64_t       total;
                        const uint64_t ex = Gr::exclusive(mine, total);
                        if (w < sq.Wq && mine) {
                            uint32_t* out = records + out_words + (carry + ex) * (uint64_t)(C + 1);
                            for (uint32_t y = x; y; y &= y - 1) {
                                const uint64_t wrote = walk_subtree<true>(cx, (w << 5) + __ffs(y) - 1, module, out);
                                out += wrote * (uint64_t)(C + 1);
                            }
                        }
                        carry += total;
                    }
                    item_sites = carry;
                }
                BSS_PHASE(6);
            }
        }
        Gr::sync();   // the group's scratch is reused by its next item
        if (MODE == kModeHeavyCount) {
            work.task_lane_counts[(size_t)slot * 32 + Gr::lane()] = lane_sites;
            if (Gr::lane() == 0) {
                work.task_sites[slot] = (uint32_t)item_sites;
                if (item_sites) {
                    const uint32_t old = atomicAdd(&work.counts[sq.item0 + u], (uint32_t)item_sites);
                    if (old + (uint32_t)item_sites < old) atomicOr((unsigned long long*)(work.totals + 2), 1ull);
                    atomicAdd((unsigned long long*)&work.chunk_sites[chunk_of_item], (unsigned long long)item_sites);
                    atomicAdd((unsigned long long*)&work.chunk_words[chunk_of_item], (unsigned long long)(item_sites * (uint64_t)(C + 1)));
                }
            }
        }
        if (MODE == kModeCount) {
            uint32_t where = 0;
            if (Gr::lane() == 0) {
                if (item_sites > 0xFFFFFFFFull) atomicOr((unsigned long long*)(work.totals + 2), 1ull);
                work.counts[sq.item0 + u] = (uint32_t)item_sites;
                my_sites += item_sites;
                my_words += item_sites * (uint64_t)(C + 1);
                if (item_sites) {
                    where             = (uint32_t)atomicAdd((unsigned long long*)(work.totals + 3), 1ull);
                    work.alive[where] = (uint32_t)(sq.item0 + u);
                }
            }
            if (G == 32 && ranked && item_sites) {   // sites of every lane's run of ranks, for the emit pass
                where = Gr::first(where);
                if (where < work.lane_cap) work.lane_counts[(size_t)where * 32 + Gr::lane()] = lane_sites;
            }
        }
    }

    if (MODE == kModeCount) {
        __shared__ unsigned long long cta_sites, cta_words;
        if (threadIdx.x == 0) { cta_sites = 0; cta_words = 0; }
        __syncthreads();
        if (my_sites) {
            atomicAdd(&cta_sites, (unsigned long long)my_sites);
            atomicAdd(&cta_words, (unsigned long long)my_words);
        }
        __syncthreads();
        // (added, not stored: the window pass and the heavy passes account for their items in the same totals, zeroed per search)
        if (threadIdx.x == 0 && cta_sites) {
            atomicAdd((unsigned long long*)&work.chunk_sites[chunk], cta_sites);
            atomicAdd((unsigned long long*)&work.chunk_words[chunk], cta_words);
        }
    }
}

// band[q] = max v - u over the set bits (u, v), u < v < n, of the pair mask of sequence q.
// One warp per mask row; grid = (sequences, row slices).
__global__ void __launch_bounds__(256) band_kernel(const DevQuery query, int32_t* __restrict__ band)
{
    const uint32_t qi = blockIdx.x;
    const int      n  = (int)(query.seq_begin[qi + 1] - query.seq_begin[qi]);
    const int      Wm = (n + 31) >> 5;
    const uint32_t* mask = query.masks + query.mask_begin[qi];
    const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int            best = -1;
    for (int u = blockIdx.y * 8 + warp; u < n; u += gridDim.y * 8) {
        for (int w = Wm - 1 - lane; w >= (u >> 5); w -= 32) {   // from the far end of the row
            uint32_t x = __ldg(mask + (size_t)u * Wm + w);
            if (w == (u >> 5)) x &= (u & 31) == 31 ? 0u : (0xFFFFFFFFu << ((u & 31) + 1));
            if (w == Wm - 1 && (n & 31)) x &= (1u << (n & 31)) - 1u;
            if (x) {
                best = max(best, (w << 5) + 31 - __clz(x) - u);
                break;   // words further left of this lane's are closer to the diagonal
            }
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) best = max(best, __shfl_xor_sync(0xFFFFFFFFu, best, d));
    if (lane == 0 && best >= 0) atomicMax(band + qi, best);
}

// Occurrence table of every sequence: row a (a < n_symbols) bit q = (seq[q] == alphabet[a]);
// with pair rows, row n_symbols + a * n_symbols + b bit q = (seq[q] == alphabet[a] && seq[q+1] == alphabet[b]).
// Rows are `stride` words long and zero beyond the sequence. With `kmers`, also the presence bits of
// the sequence's 4-, 5- and 6-mers (zeroed by the caller; gathered in shared memory, then OR-ed in).
// Grid (sequences, slices): a CTA handles kOccSliceWords words of every row of one sequence.
constexpr int kOccSliceWords = 16;

__global__ void __launch_bounds__(128) occ_kernel(const DevQuery query, const DevLibrary lib, uint32_t* __restrict__ occ,
                                                  uint32_t* __restrict__ kmers)
{
    __shared__ uint8_t  code_of[256];
    __shared__ uint32_t bits[kKmerWords];
    const uint32_t qi = blockIdx.x;
    const uint64_t sb = query.seq_begin[qi];
    const int      n  = (int)(query.seq_begin[qi + 1] - sb);
    const uint8_t* seq = query.seqs + sb;
    const int      WP = (int)query.occ_stride, ns = (int)lib.n_symbols, rows = (int)query.occ_rows;
    uint32_t*      mine = occ + (size_t)qi * rows * WP;
    const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int      w_lo = blockIdx.y * kOccSliceWords, w_hi = min(WP, w_lo + kOccSliceWords);
    for (int i = threadIdx.x; i < 256; i += 128) code_of[i] = 0xFF;
    for (int i = threadIdx.x; i < (int)kKmerWords; i += 128) bits[i] = 0;
    __syncthreads();
    if ((int)threadIdx.x < ns) code_of[lib.alphabet[threadIdx.x]] = (uint8_t)threadIdx.x;
    __syncthreads();
    for (int w = w_lo + warp; w < w_hi; w += 4) {
        const int q  = (w << 5) + lane;
        const int c0 = q < n ? (int)code_of[__ldg(seq + q)] : 0xFF, c1 = q + 1 < n ? (int)code_of[__ldg(seq + q + 1)] : 0xFF;
        for (int a = 0; a < ns; a++) {
            const uint32_t one = __ballot_sync(0xFFFFFFFFu, c0 == a);
            if (lane == 0) mine[a * WP + w] = one;
            if (rows > ns)
                for (int b = 0; b < ns; b++) {
                    const uint32_t two = __ballot_sync(0xFFFFFFFFu, c0 == a && c1 == b);
                    if (lane == 0) mine[(ns + a * ns + b) * WP + w] = two;
                }
        }
    }
    if (!kmers) return;
    for (int q = (w_lo << 5) + threadIdx.x; q < min(n, w_hi << 5) && q + (int)kKmerMin <= n; q += 128) {
        uint32_t code = 0;
        for (int j = 0; j < (int)kKmerMax && q + j < n; j++) {
            const uint32_t a = code_of[__ldg(seq + q + j)];
            if (a == 0xFFu) break;   // a byte outside the alphabet ends every k-mer that would contain it
            code |= a << (2 * j);
            if (j + 1 >= (int)kKmerMin) atomicOr(&bits[kmer_table_word((uint32_t)(j + 1) - kKmerMin) + (code >> 5)], 1u << (code & 31));
        }
    }
    __syncthreads();
    uint32_t* out = kmers + (size_t)qi * kKmerWords;
    for (int i = threadIdx.x; i < (int)kKmerWords; i += 128)
        if (bits[i]) atomicOr(&out[i], bits[i]);
}

// k-mer presence filter, one thread per (sequence, unit) item: a component that contains 4 to 6
// consecutive literal symbols cannot match a sequence in which that k-mer never occurs. Items
// ruled out this way get their count (0) here and one bit in `prefilter`; the count pass skips
// them before it touches their descriptor. Most search-bound items end here.
__global__ void __launch_bounds__(256) prefilter_kernel(const DevLibrary lib, const DevQuery query, uint32_t* __restrict__ counts,
                                                        uint32_t* __restrict__ prefilter, uint64_t n_items)
{
    for (uint64_t item = (uint64_t)blockIdx.x * 256 + threadIdx.x; item < ((n_items + 31) & ~31ull); item += (uint64_t)gridDim.x * 256) {
        bool alive = false;
        if (item < n_items) {
            const uint32_t  qi = (uint32_t)(item / lib.n_units), u = (uint32_t)(item - (uint64_t)qi * lib.n_units);
            const uint32_t* bits = query.kmers + (size_t)qi * kKmerWords;
            const uint2     pr = __ldg(reinterpret_cast<const uint2*>(lib.unit_probe) + u);   // four u16 probes
            alive = true;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t p = ((i < 2 ? pr.x : pr.y) >> (16 * (i & 1))) & 0xFFFFu;
                if (p != 0xFFFFu) {
                    const uint32_t code = p & 0xFFFu;
                    if (!((__ldg(bits + kmer_table_word(p >> 12) + (code >> 5)) >> (code & 31)) & 1u)) alive = false;
                }
            }
            if (!alive) counts[item] = 0;
        }
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, alive);
        if ((threadIdx.x & 31) == 0) prefilter[item >> 5] = word;
    }
}

cudaError_t launch_prefilter(const DevLibrary& lib, const DevQuery& q, const Work& work, int sm_count, cudaStream_t s)
{
    const uint64_t n_items = (uint64_t)q.n_seqs * lib.n_units;
    if (n_items == 0) return cudaSuccess;
    uint64_t grid = (n_items + 255) / 256;
    if (grid > (uint64_t)sm_count * 64) grid = (uint64_t)sm_count * 64;
    prefilter_kernel<<<(uint32_t)grid, 256, 0, s>>>(lib, q, work.counts, work.prefilter, n_items);
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_search(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records,
                          uint32_t n_alive, uint32_t heavy_grid, uint64_t capacity_words, cudaStream_t s)
{
    constexpr bool EMIT = (MODE & 1) != 0, HEAVY = MODE >= 2;
    const uint32_t groups = kThreads / plan.group;
    // (n_alive == UINT32_MAX: read on the device; the grid is then a fixed number of CTAs that stride over the list)
    uint32_t       grid   = HEAVY ? heavy_grid : EMIT ? (n_alive == 0xFFFFFFFFu ? heavy_grid * 4 : (n_alive + groups - 1) / groups) : plan.n_chunks;
    if (grid == 0) return cudaSuccess;
    if (grid > (1u << 20)) grid = 1u << 20;
    const uint32_t smem = EMIT ? plan.smem_bytes_emit : plan.smem_bytes;
#define BSS_LAUNCH(GG)                                                                                                    \
    do {                                                                                                                  \
        auto kernel = search_kernel<GG, MODE>;                                                                            \
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
        if (e != cudaSuccess) return e;                                                                                   \
        kernel<<<grid, kThreads, smem, s>>>(lib, q, plan, work, records, n_alive, capacity_words);                        \
        BSS_SYNC_CHECK("search_kernel", s);                                                                               \
    } while (0)
    if constexpr (HEAVY) {
        BSS_LAUNCH(32);
    } else {
        switch (plan.group) {
        case 4: BSS_LAUNCH(4); break;
        case 8: BSS_LAUNCH(8); break;
        case 16: BSS_LAUNCH(16); break;
        default: BSS_LAUNCH(32); break;
        }
    }
#undef BSS_LAUNCH
    return cudaGetLastError();
}

}   // namespace

Plan make_plan(const DevLibrary& lib, const DevQuery& q, int sm_count, const Tuning& tuning, uint32_t n_window_units, int known_band)
{
    const uint32_t n_seqs = q.n_seqs, max_len = q.max_len;
    Plan p{};
    p.W  = (max_len + 32) / 32;
    p.WS = (p.W + 31) / 32;
    // Lanes per item. Short sequences in large batches: narrow groups, several items per warp (most
    // items end in the match phase). Otherwise a whole warp per item, which also enables the
    // rank-space walk and the window walk.
    const uint64_t n_items = (uint64_t)n_seqs * lib.n_units;
    p.group = (p.W > 16 || n_items <= (1u << 17)) ? 32 : p.W <= 4 ? 4 : p.W <= 8 ? 8 : 16;
    // (tuning: developer knobs for experiments and for tests that force the rarely taken paths — narrow groups,
    //  tiny list capacity, heavy-item blocks — on small inputs; bss_library_set_tuning validates them)
    if (tuning.group > 0 && (tuning.group == 32 || (uint32_t)tuning.group >= p.W)) p.group = (uint32_t)tuning.group;   // (a narrow group holds a whole mask row)
    const uint32_t groups = kThreads / p.group;
    // Several waves of CTAs, so that the hardware scheduler evens out chunks of unequal cost;
    // a chunk holds at least one item per group.
    const uint32_t target = (uint32_t)sm_count * 24u;
    uint32_t       per_seq = n_seqs >= target ? 1u : (target + n_seqs - 1) / (n_seqs ? n_seqs : 1);
    uint32_t       upc     = (lib.n_units + per_seq - 1) / (per_seq ? per_seq : 1);
    upc                    = ((upc + groups - 1) / groups) * groups;
    if (upc < groups) upc = groups;
    // Every item takes the window pass (a library of window units on narrow-band sequences): the chunks only carry the
    // output offsets then — few large ones make the scan one short step (the emit pass sums the counts of a chunk's earlier units).
    const bool window_only = n_window_units == lib.n_units && lib.n_units > 0 && q.bmask != nullptr && known_band != INT32_MIN && known_band <= kWindowBandMax && p.group == 32 && tuning.window != 0;
    if (window_only && n_seqs <= 64) upc = std::max(upc, 256u);
    if (tuning.units_per_chunk > 0) upc = (uint32_t)tuning.units_per_chunk;
    if (upc > kMaxChunkUnits) upc = kMaxChunkUnits;
    p.units_per_chunk = upc;
    p.chunks_per_seq  = lib.n_units ? (lib.n_units + upc - 1) / upc : 0;
    p.n_chunks        = p.chunks_per_seq * n_seqs;
    // Rank-space walk (32-lane groups): room for the match lists of all but the densest few units
    // (lib.list_density = 98th percentile over the units of the expected matches per nucleotide,
    // all components together); denser items take the depth-first walk.
    p.list_cap = 0;
    if (p.group == 32) {
        const double want = 1.4 * lib.list_density * max_len + 16.0;
        p.list_cap        = want > 1024.0 ? 1024u : want < 64.0 ? 64u : (((uint32_t)want + 31u) & ~31u);
    }
    p.heavy_ranks = 1u << 15;
    if (tuning.heavy_ranks > 0) p.heavy_ranks = (uint32_t)tuning.heavy_ranks;
    p.heavy_block = p.heavy_ranks / 4 ? p.heavy_ranks / 4 : 1;
    if (tuning.list_cap >= 0) p.list_cap = p.group == 32 ? (uint32_t)tuning.list_cap : 0u;
    p.smem_bytes      = 4u * smem_words(lib.cmax, p.W, p.WS, groups, upc, p.list_cap, false);
    p.smem_bytes_emit = 4u * smem_words(lib.cmax, p.W, p.WS, groups, upc, p.list_cap, true);
    // Window walk: whole warps only; the query carries banded masks; the working set fits.
    // Window walk: whole warps only; the query carries banded masks; the band is narrow. The walk scans the words of a band
    // window, which pays while the window is a few words: measured with an all-allowed mask on 2000 nt (band 1999, the working
    // set still fits two CTAs per SM) it is 10x SLOWER than the rank-space passes, which step from match to match
    // (c4b_all 32.6 vs 3.4 ms, profiles/r02_variants_g1.jsonl) — so wide-band sequences stay with the ordinary passes.
    const int widest    = known_band == INT32_MIN ? kWindowBandMax : std::min(std::max(known_band, 0), kWindowBandMax);
    p.window_band       = kWindowBandMax;
    p.window_kw         = (uint32_t)band_words(widest);
    p.window_qcap       = std::max(64u, ((uint32_t)widest + 1u + 31u) & ~31u);
    p.smem_bytes_window = 4u * window_smem_words(lib.cmax, p.W, p.WS, kWinThreads, p.window_qcap, p.window_kw);
    // (a library with only a few window units — CaRNAval RINs, whose strands are mostly tied by non-row links — is not worth the four
    //  extra launches: c2 0.207 ms with the window pass for 18 of 400 units, 0.177 ms without)
    const bool enough = tuning.window > 0 || (uint64_t)n_window_units * 8 >= lib.n_units;
    p.window = (p.group == 32 && n_window_units > 0 && enough && q.bmask != nullptr && p.smem_bytes_window <= 110u * 1024u && tuning.window != 0 &&
                (known_band == INT32_MIN || known_band <= kWindowBandMax)) ? 1u : 0u;
    p.window_draw  = tuning.window_draw > 0 ? (uint32_t)tuning.window_draw : 1u;   // (measured on c4: 1 -> 0.146 ms, 4 -> 0.166, 32 -> 0.53: balance beats atomics)
    // Pattern table: worth it when the library shares patterns and the table of this query stays small (it is rebuilt by every search)
    const uint64_t table_bytes = 4ull * n_seqs * ((uint64_t)lib.n_slots * (p.W + p.WS) + (uint64_t)lib.n_tslots * p.W);
    p.dict         = (p.window && lib.n_slots > 0 && table_bytes <= (64ull << 20) && tuning.dict != 0) ? 1u : 0u;
    p.window_stage = 0;
    if (p.window && n_seqs == 1 && known_band >= 0 && known_band <= kWindowBandMax && tuning.window_stage > 0) {
        // occurrence table + banded mask of the one sequence in shared memory, next to the warps' scratch
        const uint32_t staged = 4u * (window_stage_words(q.occ_rows, q.occ_stride, max_len, (uint32_t)band_words(known_band)) +
                                      window_smem_words(lib.cmax, p.W, p.WS, kWinThreadsStaged, p.window_qcap, p.window_kw));
        if (staged <= 110u * 1024u) {   // two CTAs of 16 warps per SM
            p.window_stage      = 1;
            p.smem_bytes_window = staged;
        }
    }
    return p;
}

// band: [n_seqs]; band_host (may be null): pinned host copy of the bands, queued right behind the band kernel and marked by
// `copied` — the host can read the widest band while the banded masks (and whatever is queued next) are still being built.
cudaError_t launch_band(const DevQuery& q, int32_t* band, uint32_t* bmask, int32_t* band_host, cudaEvent_t copied, cudaStream_t s)
{
    if (q.n_seqs == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(band, 0xFF, sizeof(int32_t) * (size_t)q.n_seqs, s);   // -1: no pair allowed
    if (e != cudaSuccess) return e;
    // one warp per row, 8 rows per CTA and pass: slice long sequences over several CTAs
    uint32_t slices = q.n_seqs >= 1024 ? 1u : (q.max_len + 63) / 64;
    slices          = slices < 1 ? 1u : slices > 64 ? 64u : slices;
    dim3 grid(q.n_seqs, slices);
    band_kernel<<<grid, 256, 0, s>>>(q, band);
    BSS_SYNC_CHECK("band_kernel", s);
    if (band_host) {
        e = cudaMemcpyAsync(band_host, band, sizeof(int32_t) * (size_t)q.n_seqs, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaEventRecord(copied, s);
        if (e != cudaSuccess) return e;
    }
    if (bmask) bandmask_kernel<<<grid, 256, 0, s>>>(q, bmask);   // (reads the bands the kernel before it wrote)
    BSS_SYNC_CHECK("bandmask_kernel", s);
    return cudaGetLastError();
}

cudaError_t launch_occ(const DevLibrary& lib, const DevQuery& q, uint32_t* occ, uint32_t* kmers, cudaStream_t s)
{
    if (q.n_seqs == 0) return cudaSuccess;
    if (kmers) {
        cudaError_t e = cudaMemsetAsync(kmers, 0, sizeof(uint32_t) * (size_t)q.n_seqs * kKmerWords, s);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(q.n_seqs, (q.occ_stride + kOccSliceWords - 1) / kOccSliceWords);
    occ_kernel<<<grid, 128, 0, s>>>(q, lib, occ, kmers);
    BSS_SYNC_CHECK("occ_kernel", s);
    return cudaGetLastError();
}

#ifdef BSS_DEBUG_ASSERT
uint32_t g_dbg_grid = 0, g_dbg_smem_extra = 0;
#endif

template <bool EMIT>
cudaError_t launch_window(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records, uint64_t capacity_words,
                          int sm_count, cudaStream_t s)
{
    // persistent warps: the list lengths are read on the device
    const uint32_t most   = plan.window_stage ? 2u : (uint32_t)BSS_WIN_MIN_CTAS;
    const uint32_t per_sm = plan.smem_bytes_window ? std::min<uint32_t>(most, (200u * 1024u) / plan.smem_bytes_window) : most;
    uint32_t grid = (uint32_t)sm_count * (per_sm ? per_sm : 1u);
    uint32_t smem = plan.smem_bytes_window;
#ifdef BSS_DEBUG_ASSERT
    if (g_dbg_grid) grid = g_dbg_grid;
    smem += g_dbg_smem_extra;
#endif
#define BSS_LAUNCH_WIN(STAGE)                                                                                            \
    do {                                                                                                                 \
        auto kernel = window_kernel<EMIT, STAGE>;                                                                        \
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        if (e != cudaSuccess) return e;                                                                                  \
        kernel<<<grid, (STAGE) ? kWinThreadsStaged : kWinThreads, smem, s>>>(lib, q, plan, work, records, capacity_words);  \
        BSS_SYNC_CHECK(EMIT ? "window_kernel emit" : "window_kernel count", s);                                          \
    } while (0)
    if (plan.window_stage) BSS_LAUNCH_WIN(true); else BSS_LAUNCH_WIN(false);
#undef BSS_LAUNCH_WIN
    return cudaGetLastError();
}

cudaError_t launch_pattern(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, int sm_count, cudaStream_t s)
{
    const uint64_t tasks = (uint64_t)q.n_seqs * lib.n_slots;
    if (tasks == 0) return cudaSuccess;
    const uint32_t smem = 4u * (kThreads / 32) * 2u * (plan.W + plan.WS);
    cudaError_t    e    = cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t grid = std::min<uint64_t>((tasks + kThreads / 32 - 1) / (kThreads / 32), (uint64_t)sm_count * 16);
    pattern_kernel<<<(uint32_t)grid, kThreads, smem, s>>>(lib, q, plan, work);   // (also zeroes the search's totals: reset_is_folded)
    BSS_SYNC_CHECK("pattern_kernel", s);
    return cudaGetLastError();
}

cudaError_t launch_route(const DevLibrary& lib, const DevQuery& q, const Work& work, uint32_t dict, int window_band, int sm_count, cudaStream_t s)
{
    const uint64_t n_items = (uint64_t)q.n_seqs * lib.n_units;
    if (n_items == 0) return cudaSuccess;
    uint64_t grid = (n_items + 255) / 256;
    if (grid > (uint64_t)sm_count * 64) grid = (uint64_t)sm_count * 64;
    route_kernel<<<(uint32_t)grid, 256, 0, s>>>(lib, q, work, n_items, dict, window_band);
    BSS_SYNC_CHECK("route_kernel", s);
    return cudaGetLastError();
}

cudaError_t launch_count(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, bool ordinary, int sm_count, cudaStream_t s)
{
    cudaError_t e = cudaSuccess;
    if (plan.dict) e = launch_pattern(lib, q, plan, work, sm_count, s);
    if (plan.window && e == cudaSuccess) e = launch_route(lib, q, work, plan.dict, plan.window_band, sm_count, s);
    else if (work.prefilter) e = launch_prefilter(lib, q, work, sm_count, s);
    if (ordinary) {
        if (e == cudaSuccess) e = launch_search<kModeCount>(lib, q, plan, work, nullptr, 0, 0, 0, s);
        // the blocks of the items the count pass parked on the heavy list (returns at once when there are none)
        if (e == cudaSuccess && plan.group == 32 && plan.list_cap) e = launch_search<kModeHeavyCount>(lib, q, plan, work, nullptr, 0, (uint32_t)sm_count * 4u, 0, s);
    }
    if (e == cudaSuccess && plan.window && plan.n_chunks) e = launch_window<false>(lib, q, plan, work, nullptr, 0, sm_count, s);
    return e;
}

cudaError_t launch_reset(const Plan& plan, const Work& work, uint32_t n_seqs, cudaStream_t s)
{
    const uint32_t items = std::max(2 * (plan.n_chunks + 1), n_seqs + 1);
    reset_kernel<<<std::min<uint32_t>((items + 255) / 256, 64u), 256, 0, s>>>(work, plan.n_chunks, n_seqs);
    BSS_SYNC_CHECK("reset_kernel", s);
    return cudaGetLastError();
}

cudaError_t launch_scan(const DevQuery& q, const Plan& plan, const Work& work, uint64_t* seq_out, cudaStream_t s)
{
    scan_kernel<<<1, plan.n_chunks > 512 ? 1024 : 512, 0, s>>>(plan, work, q.n_seqs, seq_out);
    BSS_SYNC_CHECK("scan_kernel", s);
    return cudaGetLastError();
}

cudaError_t launch_emit(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records,
                        uint32_t n_alive, uint32_t n_heavy, uint64_t capacity_words, bool ordinary, int sm_count, cudaStream_t s)
{
    // n_alive == UINT32_MAX: sizes are read on the device (no host round trip after the count pass);
    // the heavy emit is then launched unconditionally (it returns at once without heavy items).
    const bool  blind = n_alive == 0xFFFFFFFFu;
    cudaError_t e     = cudaSuccess;
    if (ordinary) {
        e = launch_search<kModeEmit>(lib, q, plan, work, records, n_alive, (uint32_t)sm_count * 8u, capacity_words, s);   // (blind: 32 CTAs per SM stride over the list)
        if (e == cudaSuccess && (n_heavy || (blind && plan.group == 32 && plan.list_cap)))
            e = launch_search<kModeHeavyEmit>(lib, q, plan, work, records, 0, (uint32_t)sm_count * 4u, capacity_words, s);
    }
    if (e == cudaSuccess && plan.window) e = launch_window<true>(lib, q, plan, work, records, capacity_words, sm_count, s);
    return e;
}

cudaError_t launch_placements(const DevLibrary& lib, const DevQuery& q, const Plan& plan, uint64_t* scratch, size_t per_warp, uint32_t n_ctas,
                              unsigned long long* total, cudaStream_t s)
{
    if (q.n_seqs == 0 || lib.n_units == 0) return cudaSuccess;
    const uint32_t smem = 4u * (kThreads / 32) * (kBlobStage + 2 * lib.cmax * (plan.W + plan.WS));
    cudaError_t    e    = cudaFuncSetAttribute(placements_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    placements_kernel<<<n_ctas, kThreads, smem, s>>>(lib, q, plan, scratch, per_warp, total);
    return cudaGetLastError();
}

size_t placements_words_per_warp(uint32_t max_len) { return placements_scratch_words(max_len); }

}   // namespace bss


#ifdef BSS_PHASE_TIMING
extern "C" int bss_debug_phase_cycles(unsigned long long* out, int reset)
{
    cudaDeviceSynchronize();
    cudaError_t e = cudaMemcpyFromSymbol(out, bss::g_phase_cycles, sizeof(unsigned long long) * 16);
    if (e == cudaSuccess && reset) {
        unsigned long long zero[16] = {};
        e = cudaMemcpyToSymbol(bss::g_phase_cycles, zero, sizeof zero);
    }
    return (int)e;
}
#endif

#ifdef BSS_DEBUG_ASSERT
extern "C" void bss_debug_set_launch(unsigned grid, unsigned smem_extra)
{
    bss::g_dbg_grid       = grid;
    bss::g_dbg_smem_extra = smem_extra;
}

extern "C" int bss_debug_set_skip(int mask)
{
    return (int)cudaMemcpyToSymbol(bss::g_win_skip, &mask, sizeof mask);
}

extern "C" int bss_debug_assert_counts(unsigned int* out)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, bss::g_assert_fail, sizeof(unsigned int) * 32);
}
#endif
