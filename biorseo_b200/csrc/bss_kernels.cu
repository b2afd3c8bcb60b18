// Hand-written sm_100a kernels of the insertion-site search.
//
// What they replace (reference, CPU): one Pool job per module running std::regex over the
// query and materialising every placement before filtering it
// (cppsrc/Motif.cpp:432-494, cppsrc/MOIP.cpp:1069-1085, 1126-1143, 1178-1212).
//
// Shape of the computation. One CTA owns one chunk = (sequence, contiguous range of units).
// It first turns the query into per-symbol occurrence bit rows in shared memory
// (bit q of row a = "sequence byte q is alphabet symbol a"). A group of G lanes
// (G = 4..32, picked from the sequence length) then draws (sequence, unit) items from the
// chunk's counter and, for each:
//   1. match:  for every component, M_c = AND_j (occ[pattern_j] >> j), 32 positions per
//              lane and step, with a one-bit-per-word summary for fast "next match" queries;
//   2. thin:   components whose matches overlap on this sequence get T_c, the leftmost
//              non-overlapping matches of every run of overlapping matches. The reference's
//              candidates for a component in the suffix starting at t (sregex_iterator) are
//              then: a greedy walk over M_c from t until it lands on a bit of T_c, and T_c
//              itself from there on;
//   3. walk:   lanes split the candidates of component 0 by mask word and enumerate the
//              placements below each depth-first. Every base-pair check is tested as soon as
//              both of its ends are placed, and a level whose component is tied by a check to
//              an already placed position only scans the window the pair mask's band allows
//              (pruning is result-neutral: SURVEY.md appendix B);
//   4. count / emit: pass 1 counts sites per item, a scan turns counts into offsets,
//              pass 2 repeats the walk and writes [module, p_0..p_{C-1}] records at their
//              final, canonical (sequence, unit, depth-first) position. No sort, no atomics
//              on the output.
// Not a contraction: integer/bit work only, no tensor cores.
#include "bss_device.cuh"

namespace bss {
namespace {

__device__ __forceinline__ int unpack_anchor(uint32_t w) { return (int)(int8_t)(w & 0xFFu); }
__device__ __forceinline__ int unpack_off(uint32_t w) { return (int)(int16_t)((w >> 8) & 0xFFFFu); }

// First set bit at or after position t (t >= 0) of a bitset with a one-bit-per-word summary, or -1.
__device__ __forceinline__ int next_bit(const uint32_t* __restrict__ B, const uint32_t* __restrict__ S, int W, int WS, int t)
{
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

__device__ __forceinline__ bool test_bit(const uint32_t* __restrict__ B, int q) { return (B[q >> 5] >> (q & 31)) & 1u; }

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

__device__ __forceinline__ int placed(const ItemCtx& cx, int level) { return (int)(cx.stack[level * kThreads] & 0xFFFFu); }

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
        const int      a = min(u, v), b = max(u, v);
        if (a < 0 || b >= cx.n || a == b) return false;
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
// Returns false as soon as one component has no match at all (the unit cannot be placed).
template <int G>
__device__ __forceinline__ bool match_components(const uint32_t* blob, const uint32_t* occ, int WP, uint32_t* M, uint32_t* S,
                                                 int W, int WS, int n)
{
    using Gr       = Group<G>;
    const int  C   = (int)(blob[1] & 0xFFu);
    const int  Wq  = (n + 32) >> 5;   // ceil((n + 1) / 32): positions 0..n
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(blob);
    for (int c = 0; c < C; c++) {
        const uint32_t desc = blob[4 + 2 * c];
        const int      k    = (int)(desc & 0xFFu);
        const uint8_t* pat  = bytes + ((desc >> 8) & 0xFFFFu);
        const int      last = n - k;   // last position where the pattern still fits (k = 0: position n)
        bool           any  = false;
        for (int w0 = 0; w0 < W; w0 += G) {
            const int w = w0 + (int)Gr::lane();
            uint32_t  m = 0;
            if (w < Wq && last >= (w << 5)) {
                const int hi = last - (w << 5);
                m            = hi >= 31 ? 0xFFFFFFFFu : ((2u << hi) - 1u);
                for (int j = 0; j < k && m; j++) {
                    const uint32_t sym = pat[j];
                    if (sym == kWildcard) continue;
                    const uint32_t* row = occ + sym * WP + w + (j >> 5);
                    m &= __funnelshift_r(row[0], row[1], j & 31);
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
        for (int w0 = 0; w0 < W; w0 += G) {   // run starts: matches without a match in the k-1 positions before them
            const int w = w0 + (int)Gr::lane();
            uint32_t  m = 0, starts = 0;
            if (w < W) {
                m             = Mc[w];
                uint32_t near = 0;
                if (m)
                    for (int s = 1; s < k; s++) near |= shifted_up(Mc, w, s);
                starts = m & ~near;
                Tc[w]  = starts;
            }
            any |= Gr::ballot(m != starts) != 0;
        }
        if (!any) continue;
        ovl |= 1u << c;
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
    uint32_t* N;       // [cap] placements below each match (components c+1.. only)
    uint32_t* PS;      // [cap + C] per component a_c + 1 prefix sums of N over the thinned matches
    uint16_t* MWP;     // [C][W] matches before each mask word
    uint16_t* lbeg;    // [C + 1] first list entry of each component
};

__device__ __forceinline__ int list_len(const RankCtx& rc, int c) { return (int)rc.lbeg[c + 1] - (int)rc.lbeg[c]; }

// Number of matches of component c at positions < x.
__device__ __forceinline__ int match_rank(const RankCtx& rc, int c, int x)
{
    if (x <= 0) return 0;
    const int w = x >> 5;
    if (w >= rc.cx.W) return list_len(rc, c);
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
// UINT64_MAX when a count does not fit 31 bits (the caller falls back to the depth-first walk).
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
                        n += rc.N[rc.lbeg[c + 1] + match_rank(rc, c + 1, e)];
                        e = next_bit(cx.M + (c + 1) * cx.W, cx.SM + (c + 1) * cx.WS, cx.W, cx.WS, e + ke);
                    }
                    n += PSn[kd.jb] - PSn[kd.ja];
                }
                if (n >= 0x80000000ull) overflow = true;
                rc.N[b + j] = (uint32_t)n;
                if (!is_thinned(rc, c, p)) n = 0;   // only ever reached as a greedy match, never through a range
            }
            uint64_t inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t up = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += up;
            }
            if (j < a) PS[j + 1] = (uint32_t)(running + inc);
            running += __shfl_sync(0xFFFFFFFFu, inc, 31);
            if (running >= 0x80000000ull) overflow = true;
        }
        total = running;
        __syncwarp();
    }
    if (__any_sync(0xFFFFFFFFu, overflow)) return ~0ull;
    return total;
}

// Positions of the placement of depth-first rank r into this lane's stack column.
__device__ __forceinline__ void rank_unrank(const RankCtx& rc, uint32_t r)
{
    const ItemCtx& cx = rc.cx;
    int            p;
    {
        const uint32_t* PS = rc.PS;
        const int       j  = rank_search(PS, 0, list_len(rc, 0), r);
        r -= PS[j];
        p = rc.mpos[j];
        cx.stack[0] = (uint32_t)p;
    }
    for (int c = 1; c < cx.C; c++) {
        const Kids      kd = kids_of(rc, c, p);
        const int       b  = rc.lbeg[c];
        const uint32_t* PS = rc.PS + b + c;
        bool            done = false;
        if (kd.nex) {
            const int ke = max(1, (int)(cx.blob[4 + 2 * c] & 0xFFu));
            int       e  = kd.efirst;
            for (int x = 0; x < kd.nex; x++) {
                const uint32_t n = rc.N[b + match_rank(rc, c, e)];
                if (r < n) { p = e; done = true; break; }
                r -= n;
                e = next_bit(cx.M + c * cx.W, cx.SM + c * cx.WS, cx.W, cx.WS, e + ke);
            }
        }
        if (!done) {
            const int j = rank_search(PS, kd.ja, kd.jb, PS[kd.ja] + r);
            r -= PS[j] - PS[kd.ja];
            p = rc.mpos[b + j];
        }
        cx.stack[c * kThreads] = (uint32_t)p;
    }
}

__device__ __forceinline__ bool all_checks_pass(const ItemCtx& cx)
{
    const int       nk  = (int)((cx.blob[1] >> 8) & 0xFFu);
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16);
    for (int i = 0; i < nk; i++) {
        const uint32_t ea = chk[2 * i], eb = chk[2 * i + 1];
        const int      aa = unpack_anchor(ea), ba = unpack_anchor(eb);
        const int      u  = (aa < 0 ? 0 : placed(cx, aa)) + unpack_off(ea);
        const int      v  = (ba < 0 ? 0 : placed(cx, ba)) + unpack_off(eb);
        const int      a = min(u, v), b = max(u, v);
        if (a < 0 || b >= cx.n || a == b) return false;
        if (!((__ldg(cx.mask + (size_t)a * cx.mstride + (b >> 5)) >> (b & 31)) & 1u)) return false;
    }
    return true;
}

// Enumerates ranks [0, total) 32 at a time. Returns the number of sites; when EMIT, writes
// their records from `out` on, staged through `stage` (32 * (C + 1) words) so that every
// batch leaves as one contiguous block.
template <bool EMIT>
__device__ __forceinline__ uint64_t rank_walk(const RankCtx& rc, uint32_t total, uint32_t module, uint32_t* __restrict__ out,
                                              uint32_t* stage)
{
    const ItemCtx& cx   = rc.cx;
    const uint32_t lane = threadIdx.x & 31;
    const int      R    = cx.C + 1;
    uint64_t       found = 0;
    for (uint32_t f0 = 0; f0 < total; f0 += 32) {
        const uint32_t f  = f0 + lane;
        bool           ok = f < total;
        if (ok) {
            rank_unrank(rc, f);
            ok = all_checks_pass(cx);
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
        if (EMIT && bal) {
            if (ok) {
                uint32_t* rec = stage + __popc(bal & ((1u << lane) - 1u)) * R;
                rec[0]        = module;
                for (int c = 0; c < cx.C; c++) rec[1 + c] = (uint32_t)placed(cx, c);
            }
            __syncwarp();
            const uint32_t words = (uint32_t)__popc(bal) * R;
            uint32_t*      dst   = out + found * (uint64_t)R;
            for (uint32_t i = lane; i < words; i += 32) dst[i] = stage[i];
            __syncwarp();
        }
        found += __popc(bal);
    }
    return found;
}

struct Smem {
    uint32_t* occ;       // [n_symbols][WP]
    uint32_t* blob;      // [groups][kBlobStage]
    uint32_t* M;         // [groups][cmax][W]
    uint32_t* T;         // [groups][cmax][W]
    uint32_t* SM;        // [groups][cmax][WS]
    uint32_t* ST;        // [groups][cmax][WS]
    uint32_t* stack;     // [cmax][kThreads]
    uint32_t* rank;      // [groups][rank_words] rank-space scratch (32-lane groups only)
    uint64_t* item_off;  // [units_per_chunk] (emit only)
};

// Rank-space scratch of one group, in words: N, PS, staging, then the u16 arrays.
__host__ __device__ inline uint32_t rank_words(uint32_t cmax, uint32_t W, uint32_t cap)
{
    if (cap == 0) return 0;
    return cap + (cap + cmax + 1) + 32 * (cmax + 1) + (cap + 1) / 2 + (cmax * W + 1) / 2 + (cmax + 2) / 2;
}

__host__ __device__ inline uint32_t smem_words(uint32_t n_symbols, uint32_t cmax, uint32_t W, uint32_t WS, uint32_t groups,
                                               uint32_t units_per_chunk, uint32_t cap)
{
    uint32_t words = n_symbols * (W + kOccPad);
    words += groups * (kBlobStage + 2 * cmax * (W + WS) + rank_words(cmax, W, cap));
    words += cmax * kThreads;
    words = (words + 1) & ~1u;
    words += 2 * units_per_chunk;
    return words;
}

template <int G>
__device__ __forceinline__ Smem carve(uint32_t* base, const DevLibrary& lib, const Plan& plan)
{
    const uint32_t groups = kThreads / G, WP = plan.W + kOccPad;
    Smem           s;
    s.occ   = base;               base += lib.n_symbols * WP;
    s.blob  = base;               base += groups * kBlobStage;
    s.M     = base;               base += groups * lib.cmax * plan.W;
    s.T     = base;               base += groups * lib.cmax * plan.W;
    s.SM    = base;               base += groups * lib.cmax * plan.WS;
    s.ST    = base;               base += groups * lib.cmax * plan.WS;
    s.stack = base;               base += lib.cmax * kThreads;
    s.rank  = base;               base += groups * rank_words(lib.cmax, plan.W, plan.list_cap);
    uintptr_t p = reinterpret_cast<uintptr_t>(base);
    p           = (p + 7) & ~(uintptr_t)7;
    s.item_off  = reinterpret_cast<uint64_t*>(p);
    return s;
}

// Occurrence rows of the sequence: occ[a][w] bit b = (seq[32w + b] == alphabet[a]).
__device__ __forceinline__ void build_occurrence(uint32_t* occ, const DevLibrary& lib, const uint8_t* __restrict__ seq, int n, int W)
{
    const int WP = W + kOccPad;
    for (uint32_t i = threadIdx.x; i < lib.n_symbols * (uint32_t)WP; i += kThreads) occ[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int w = warp; w < ((n + 31) >> 5); w += kThreads / 32) {
        const int     q    = (w << 5) + lane;
        const int     byte = q < n ? (int)__ldg(seq + q) : -1;
        for (uint32_t a = 0; a < lib.n_symbols; a++) {
            const uint32_t bits = __ballot_sync(0xFFFFFFFFu, byte == (int)lib.alphabet[a]);
            if (lane == 0) occ[a * WP + w] = bits;
        }
    }
    __syncthreads();
}

template <int G, bool EMIT>
__global__ void __launch_bounds__(kThreads) search_kernel(const DevLibrary lib, const DevQuery query, const Plan plan,
                                                          const Work work, uint32_t* __restrict__ records)
{
    using Gr = Group<G>;
    extern __shared__ __align__(16) uint32_t smem_raw[];
    __shared__ uint32_t next_unit;
    const uint32_t chunk = blockIdx.x;
    if (EMIT && work.chunk_sites[chunk] == 0) return;   // nothing to write for this chunk

    const uint32_t qi = chunk / plan.chunks_per_seq;
    const uint32_t u0 = (chunk % plan.chunks_per_seq) * plan.units_per_chunk;
    const uint32_t u1 = min(lib.n_units, u0 + plan.units_per_chunk);
    const uint64_t sb = query.seq_begin[qi];
    const int      n  = (int)(query.seq_begin[qi + 1] - sb);
    const uint8_t* seq = query.seqs + sb;
    const uint32_t* pair_mask = query.masks + query.mask_begin[qi];
    const int      band = query.band[qi];
    const int      W = (int)plan.W, WS = (int)plan.WS, WP = W + kOccPad;
    const int      Wq = (n + 32) >> 5;

    Smem sm = carve<G>(smem_raw, lib, plan);
    if (threadIdx.x == 0) next_unit = u0;
    build_occurrence(sm.occ, lib, seq, n, W);

    const uint32_t gid    = threadIdx.x / G;
    uint32_t*      g_blob = sm.blob + gid * kBlobStage;
    uint32_t*      g_M    = sm.M + gid * lib.cmax * W;
    uint32_t*      g_T    = sm.T + gid * lib.cmax * W;
    uint32_t*      g_SM   = sm.SM + gid * lib.cmax * WS;
    uint32_t*      g_ST   = sm.ST + gid * lib.cmax * WS;
    const uint64_t item0  = (uint64_t)qi * lib.n_units;

    if (EMIT) {
        // Word offset of every item of the chunk: exclusive scan of count * record length.
        // One warp does it; chunks hold at most kMaxChunkUnits items.
        if (threadIdx.x < 32) {
            uint64_t carry = work.chunk_words[chunk];
            for (uint32_t b = u0; b < u1; b += 32) {
                const uint32_t u = b + threadIdx.x;
                uint64_t       v = 0;
                if (u < u1) {
                    const uint32_t c = work.counts[item0 + u];
                    if (c) v = (uint64_t)c * (1u + (__ldg(lib.blob + __ldg(lib.unit_off + u) + 1) & 0xFFu));
                }
                uint64_t total;
                const uint64_t ex = Group<32>::exclusive(v, total);
                if (u < u1) sm.item_off[u - u0] = carry + ex;
                carry += total;
            }
        }
        __syncthreads();
    }

    uint64_t my_sites = 0, my_words = 0;   // accumulated by lane 0 of each group (count pass)

    for (;;) {
        uint32_t u = 0;
        if (Gr::lane() == 0) u = atomicAdd(&next_unit, 1u);
        u = Gr::first(u);
        if (u >= u1) break;
        if (EMIT && work.counts[item0 + u] == 0) continue;

        // stage the unit descriptor
        const uint32_t  boff  = __ldg(lib.unit_off + u);
        const uint32_t  blen  = __ldg(lib.unit_off + u + 1) - boff;
        const uint32_t* gblob = lib.blob + boff;
        const uint32_t* blob  = gblob;
        if (blen <= kBlobStage) {
            for (uint32_t i = Gr::lane(); i < blen; i += G) g_blob[i] = __ldg(gblob + i);
            blob = g_blob;
            Gr::sync();
        }
        const uint32_t module = blob[0] + lib.module_base;
        const int      C      = (int)(blob[1] & 0xFFu);
        const int      delay  = (int)(blob[1] >> 16);
        const uint32_t gate   = blob[2];

        bool alive = true;
        if (!EMIT && gate != kNoGate) {   // an emitted item has already passed its gate
            const uint32_t* gb = lib.blob + __ldg(lib.unit_off + gate);
            alive              = gate_open<G>(gb, sm.occ, WP, g_M, g_SM, W, WS, n);
            Gr::sync();
        }
        // An empty first pattern at position 0 with delay 0 leaves "no room" (unsigned wrap in
        // the reference, cppsrc/Motif.cpp:461) and ends the enumeration before it starts.
        if (C > 1 && (blob[4] & 0xFFu) == 0 && delay == 0) alive = false;
        if (alive) alive = match_components<G>(blob, sm.occ, WP, g_M, g_SM, W, WS, n);

        uint64_t item_sites = 0;
        if (alive) {
            ItemCtx cx;
            cx.blob = blob; cx.M = g_M; cx.T = g_T; cx.SM = g_SM; cx.ST = g_ST; cx.mask = pair_mask;
            cx.stack = sm.stack + threadIdx.x;
            cx.ovl   = thin_components<G>(blob, g_M, g_SM, g_T, g_ST, W, WS);
            cx.W = W; cx.WS = WS; cx.n = n; cx.mstride = (n + 31) >> 5; cx.C = C; cx.delay = delay; cx.band = band;

            bool ranked = false;
            if constexpr (G == 32) {
                if (plan.list_cap) {
                    const uint32_t cap = plan.list_cap;
                    uint32_t*      rw  = sm.rank + gid * rank_words(lib.cmax, plan.W, cap);
                    RankCtx        rc;
                    rc.cx   = cx;
                    rc.N    = rw;                        rw += cap;
                    rc.PS   = rw;                        rw += cap + lib.cmax + 1;
                    uint32_t* stage = rw;                rw += 32 * (lib.cmax + 1);
                    rc.mpos = reinterpret_cast<uint16_t*>(rw);   rw += (cap + 1) / 2;
                    rc.MWP  = reinterpret_cast<uint16_t*>(rw);   rw += (lib.cmax * plan.W + 1) / 2;
                    rc.lbeg = reinterpret_cast<uint16_t*>(rw);
                    if (rank_build_lists(rc, (int)cap)) {
                        const uint64_t total = rank_count(rc);
                        if (total != ~0ull) {
                            ranked = true;
                            if (total) {
                                uint32_t* out = EMIT ? records + sm.item_off[u - u0] : nullptr;
                                item_sites    = rank_walk<EMIT>(rc, (uint32_t)total, module, out, stage);
                            }
                        }
                    }
                }
            }
            if (!ranked) {
            // Component 0: its chain from position 0 starts on a run start, i.e. it is the thinned set.
            const uint32_t* first = (cx.ovl & 1u) ? g_T : g_M;

            if (!EMIT) {
                uint64_t mine = 0;
                for (int w = Gr::lane(); w < Wq; w += G)
                    for (uint32_t x = first[w]; x; x &= x - 1) mine += walk_subtree<false>(cx, (w << 5) + __ffs(x) - 1, module, nullptr);
                item_sites = Gr::sum(mine);
            } else {
                // sites below each word of component 0, then their exclusive prefix
                uint64_t carry = 0;
                for (int w0 = 0; w0 < Wq; w0 += G) {
                    const int w    = w0 + (int)Gr::lane();
                    uint64_t  mine = 0;
                    uint32_t  x    = 0;
                    if (w < Wq) {
                        x = first[w];
                        for (uint32_t y = x; y; y &= y - 1) mine += walk_subtree<false>(cx, (w << 5) + __ffs(y) - 1, module, nullptr);
                    }
                    uint64_t       total;
                    const uint64_t ex = Gr::exclusive(mine, total);
                    if (w < Wq && mine) {
                        uint32_t* out = records + sm.item_off[u - u0] + (carry + ex) * (uint64_t)(C + 1);
                        for (uint32_t y = x; y; y &= y - 1) {
                            const uint64_t wrote = walk_subtree<true>(cx, (w << 5) + __ffs(y) - 1, module, out);
                            out += wrote * (uint64_t)(C + 1);
                        }
                    }
                    carry += total;
                }
                item_sites = carry;
            }
            }
        }
        Gr::sync();   // the group's scratch is reused by its next item
        if (!EMIT && Gr::lane() == 0) {
            if (item_sites > 0xFFFFFFFFull) atomicOr((unsigned long long*)(work.totals + 2), 1ull);
            work.counts[item0 + u] = (uint32_t)item_sites;
            my_sites += item_sites;
            my_words += item_sites * (uint64_t)(C + 1);
        }
    }

    if (!EMIT) {
        __shared__ unsigned long long cta_sites, cta_words;
        if (threadIdx.x == 0) { cta_sites = 0; cta_words = 0; }
        __syncthreads();
        if (my_sites) {
            atomicAdd(&cta_sites, (unsigned long long)my_sites);
            atomicAdd(&cta_words, (unsigned long long)my_words);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            work.chunk_sites[chunk] = cta_sites;
            work.chunk_words[chunk] = cta_words;
        }
    }
}

// Block-wide exclusive prefix of one u64 per thread (1024 threads); `total` = block sum.
__device__ __forceinline__ uint64_t block_exclusive(uint64_t v, uint64_t* warp_sums, uint64_t& total)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t       wt;
    const uint64_t in_warp = Group<32>::exclusive(v, wt);
    __syncthreads();   // warp_sums may still be read from the previous call
    if (lane == 31) warp_sums[warp] = wt;
    __syncthreads();
    if (warp == 0) {
        uint64_t       all;
        const uint64_t ex = Group<32>::exclusive(warp_sums[lane], all);
        warp_sums[lane]   = ex;
        if (lane == 0) warp_sums[32] = all;
    }
    __syncthreads();
    total = warp_sums[32];
    return warp_sums[warp] + in_warp;
}

// Exclusive prefix of the per-chunk word counts, totals, and per-sequence word offsets.
__global__ void __launch_bounds__(1024) scan_kernel(const Plan plan, const Work work, uint32_t n_seqs)
{
    __shared__ uint64_t warp_sums[33];
    uint64_t            carry_words = 0, carry_sites = 0;   // identical in every thread
    for (uint32_t base = 0; base < plan.n_chunks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t w = i < plan.n_chunks ? work.chunk_words[i] : 0;
        const uint64_t s = i < plan.n_chunks ? work.chunk_sites[i] : 0;
        uint64_t       tw, ts;
        const uint64_t ew = block_exclusive(w, warp_sums, tw);
        block_exclusive(s, warp_sums, ts);
        if (i < plan.n_chunks) {
            work.chunk_words[i] = carry_words + ew;
            if (i % plan.chunks_per_seq == 0) work.seq_word_begin[i / plan.chunks_per_seq] = carry_words + ew;
        }
        carry_words += tw;
        carry_sites += ts;
    }
    if (threadIdx.x == 0) {
        work.chunk_words[plan.n_chunks] = carry_words;
        work.seq_word_begin[n_seqs]     = carry_words;
        work.totals[0]                  = carry_words;
        work.totals[1]                  = carry_sites;
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

template <bool EMIT>
cudaError_t launch_search(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records,
                          cudaStream_t s)
{
    if (plan.n_chunks == 0) return cudaSuccess;
#define BSS_LAUNCH(GG)                                                                                                    \
    do {                                                                                                                  \
        auto kernel = search_kernel<GG, EMIT>;                                                                            \
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);  \
        if (e != cudaSuccess) return e;                                                                                   \
        kernel<<<plan.n_chunks, kThreads, plan.smem_bytes, s>>>(lib, q, plan, work, records);                            \
    } while (0)
    switch (plan.group) {
    case 4: BSS_LAUNCH(4); break;
    case 8: BSS_LAUNCH(8); break;
    case 16: BSS_LAUNCH(16); break;
    default: BSS_LAUNCH(32); break;
    }
#undef BSS_LAUNCH
    return cudaGetLastError();
}

}   // namespace

Plan make_plan(const DevLibrary& lib, uint32_t n_seqs, uint32_t max_len, int sm_count)
{
    Plan p{};
    p.W  = (max_len + 32) / 32;
    p.WS = (p.W + 31) / 32;
    p.group = p.W <= 4 ? 4 : p.W <= 8 ? 8 : p.W <= 16 ? 16 : 32;
    const uint32_t groups = kThreads / p.group;
    // Several waves of CTAs, so that the hardware scheduler evens out chunks of unequal cost;
    // a chunk holds at least a few items per group to amortise its occurrence rows.
    const uint32_t target = (uint32_t)sm_count * 24u;
    uint32_t       per_seq = n_seqs >= target ? 1u : (target + n_seqs - 1) / (n_seqs ? n_seqs : 1);
    uint32_t       upc     = (lib.n_units + per_seq - 1) / (per_seq ? per_seq : 1);
    upc                    = ((upc + groups - 1) / groups) * groups;
    if (upc < 4 * groups) upc = 4 * groups;
    if (upc > kMaxChunkUnits) upc = kMaxChunkUnits;
    p.units_per_chunk = upc;
    p.chunks_per_seq  = lib.n_units ? (lib.n_units + upc - 1) / upc : 0;
    p.n_chunks        = p.chunks_per_seq * n_seqs;
    p.list_cap        = p.group == 32 ? 384u : 0u;
    p.smem_bytes      = 4u * smem_words(lib.n_symbols, lib.cmax, p.W, p.WS, groups, upc, p.list_cap);
    return p;
}

cudaError_t launch_band(const DevQuery& q, int32_t* band, cudaStream_t s)
{
    if (q.n_seqs == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(band, 0xFF, sizeof(int32_t) * (size_t)q.n_seqs, s);   // -1: no pair allowed
    if (e != cudaSuccess) return e;
    // one warp per row, 8 rows per CTA and pass: slice long sequences over several CTAs
    uint32_t slices = q.n_seqs >= 1024 ? 1u : (q.max_len + 63) / 64;
    slices          = slices < 1 ? 1u : slices > 64 ? 64u : slices;
    dim3 grid(q.n_seqs, slices);
    band_kernel<<<grid, 256, 0, s>>>(q, band);
    return cudaGetLastError();
}

cudaError_t launch_count(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, cudaStream_t s)
{
    return launch_search<false>(lib, q, plan, work, nullptr, s);
}

cudaError_t launch_scan(const DevQuery& q, const Plan& plan, const Work& work, cudaStream_t s)
{
    scan_kernel<<<1, 1024, 0, s>>>(plan, work, q.n_seqs);
    return cudaGetLastError();
}

cudaError_t launch_emit(const DevLibrary& lib, const DevQuery& q, const Plan& plan, const Work& work, uint32_t* records,
                        cudaStream_t s)
{
    return launch_search<true>(lib, q, plan, work, records, s);
}

}   // namespace bss
