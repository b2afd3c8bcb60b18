// Hand-written sm_100a kernels of the insertion-site search.
//
// What they replace (reference, CPU): one Pool job per module running std::regex over the
// query and materialising every placement before filtering it
// (cppsrc/Motif.cpp:432-494, cppsrc/MOIP.cpp:1069-1085, 1126-1143, 1178-1212).
//
// Shape of the computation. One CTA owns one chunk = (sequence, contiguous range of units).
// It first turns the query into per-symbol occurrence bit rows in shared memory
// (bit q of row a = "sequence byte q is alphabet symbol a"). A group of G lanes
// (G = 4..32, picked from the sequence length) then owns one (sequence, unit) item:
//   1. match:  for every component, M_c = AND_j (occ[pattern_j] >> j), 32 positions per
//              lane and step, with a one-bit-per-word summary for fast "next match" queries;
//   2. walk:   lanes split the matches of component 0 by mask word and enumerate the
//              placements below each in depth-first order, following the reference's
//              "leftmost non-overlapping matches of the next pattern in the suffix after
//              this component" rule, testing every base-pair check as soon as both of its
//              ends are placed (pruning is result-neutral: SURVEY.md appendix B);
//   3. count / emit: pass 1 counts sites per item, a scan turns counts into offsets,
//              pass 2 repeats the walk and writes [module, p_0..p_{C-1}] records at their
//              final, canonical (sequence, unit, depth-first) position. No sort, no atomics
//              on the output.
// Not a contraction: integer/bit work only, no tensor cores.
#include "bss_device.cuh"

namespace bss {
namespace {

__device__ __forceinline__ int unpack_anchor(uint32_t w) { return (int)(int8_t)(w & 0xFFu); }
__device__ __forceinline__ int unpack_off(uint32_t w) { return (int)(int16_t)((w >> 8) & 0xFFFFu); }

// First set bit at or after position t in a component's match mask, or -1.
__device__ __forceinline__ int next_match(const uint32_t* __restrict__ M, const uint32_t* __restrict__ S, int W, int WS, int t)
{
    int w = t >> 5;
    if (w >= W) return -1;
    uint32_t x = M[w] & (0xFFFFFFFFu << (t & 31));
    if (x) return (w << 5) + __ffs(x) - 1;
    if (++w >= W) return -1;
    int      sw = w >> 5;
    uint32_t s  = S[sw] & (0xFFFFFFFFu << (w & 31));
    while (!s) {
        if (++sw >= WS) return -1;
        s = S[sw];
    }
    w = (sw << 5) + __ffs(s) - 1;
    return (w << 5) + __ffs(M[w]) - 1;
}

struct ItemCtx {
    const uint32_t* blob;     // unit descriptor (shared or global memory)
    const uint32_t* M;        // [C][W] match masks
    const uint32_t* S;        // [C][WS] summaries
    const uint32_t* mask;     // pair mask of the sequence, row stride mstride words
    uint32_t*       stack_p;  // this thread's column of the position stack, stride kThreads
    uint32_t*       stack_t;  // this thread's column of the cursor stack
    int             W, WS, n, mstride, C, delay;
};

// All checks that become decidable once component `level` is placed.
__device__ __forceinline__ bool checks_pass(const ItemCtx& cx, int level)
{
    const uint32_t range = cx.blob[5 + 2 * level];
    const int      cb = (int)(range & 0xFFu), ce = (int)((range >> 8) & 0xFFu);
    const uint32_t* chk = cx.blob + (cx.blob[3] >> 16);
    for (int i = cb; i < ce; i++) {
        const uint32_t ea = chk[2 * i], eb = chk[2 * i + 1];
        const int      aa = unpack_anchor(ea), ba = unpack_anchor(eb);
        const int      u  = (aa < 0 ? 0 : (int)cx.stack_p[aa * kThreads]) + unpack_off(ea);
        const int      v  = (ba < 0 ? 0 : (int)cx.stack_p[ba * kThreads]) + unpack_off(eb);
        const int      a = min(u, v), b = max(u, v);
        if (a < 0 || b >= cx.n || a == b) return false;
        if (!((__ldg(cx.mask + (size_t)a * cx.mstride + (b >> 5)) >> (b & 31)) & 1u)) return false;
    }
    return true;
}

// Enumerates every site whose component 0 sits at p0. Returns the number of sites; when
// EMIT, also writes their records starting at out (rec_words each).
template <bool EMIT>
__device__ __forceinline__ uint64_t walk_subtree(const ItemCtx& cx, int p0, uint32_t module, uint32_t* __restrict__ out)
{
    const int C = cx.C;
    cx.stack_p[0] = (uint32_t)p0;
    if (!checks_pass(cx, 0)) return 0;
    uint64_t found = 0;
    if (C == 1) {
        if (EMIT) { out[0] = module; out[1] = (uint32_t)p0; }
        return 1;
    }
    {
        const int k0 = (int)(cx.blob[4] & 0xFFu);
        const int t  = p0 + k0 - 1 + cx.delay;
        if (t < 0 || t >= cx.n) return 0;   // no room for the next component
        cx.stack_t[1 * kThreads] = (uint32_t)t;
    }
    int level = 1;
    while (level >= 1) {
        const int k = (int)(cx.blob[4 + 2 * level] & 0xFFu);
        const int q = next_match(cx.M + level * cx.W, cx.S + level * cx.WS, cx.W, cx.WS, (int)cx.stack_t[level * kThreads]);
        if (q < 0) { level--; continue; }
        cx.stack_t[level * kThreads] = (uint32_t)(q + (k > 0 ? k : 1));   // matches of one pattern never overlap
        cx.stack_p[level * kThreads] = (uint32_t)q;
        if (!checks_pass(cx, level)) continue;
        if (level == C - 1) {
            if (EMIT) {
                uint32_t* rec = out + found * (uint64_t)(C + 1);
                rec[0] = module;
                for (int c = 0; c < C; c++) rec[1 + c] = cx.stack_p[c * kThreads];
            }
            found++;
            continue;
        }
        const int t = q + k - 1 + cx.delay;
        if (t < 0 || t >= cx.n) { level--; continue; }   // later matches leave even less room
        level++;
        cx.stack_t[level * kThreads] = (uint32_t)t;
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
        const uint8_t* pat  = bytes + (desc >> 8);
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

// Greedy earliest placement: does the gate unit have any placement on this sequence?
template <int G>
__device__ __forceinline__ bool gate_open(const uint32_t* gblob, const uint32_t* occ, int WP, uint32_t* M, uint32_t* S, int W,
                                          int WS, int n)
{
    if (!match_components<G>(gblob, occ, WP, M, S, W, WS, n)) return false;
    const int C = (int)(gblob[1] & 0xFFu), delay = (int)(gblob[1] >> 16);
    int       t = 0;
    for (int c = 0; c < C; c++) {
        const int q = next_match(M + c * W, S + c * WS, W, WS, t);
        if (q < 0) return false;
        t = q + (int)(gblob[4 + 2 * c] & 0xFFu) - 1 + delay;
    }
    return true;
}

struct Smem {
    uint32_t* occ;       // [n_symbols][WP]
    uint32_t* blob;      // [groups][kBlobStage]
    uint32_t* M;         // [groups][cmax][W]
    uint32_t* S;         // [groups][cmax][WS]
    uint32_t* T;         // [groups][W]   thinned chain of component 0 / per-word site counts
    uint32_t* stack_p;   // [cmax][kThreads]
    uint32_t* stack_t;   // [cmax][kThreads]
    uint64_t* item_off;  // [units_per_chunk] (emit only)
};

__host__ __device__ inline uint32_t smem_words(uint32_t n_symbols, uint32_t cmax, uint32_t W, uint32_t WS, uint32_t groups,
                                               uint32_t units_per_chunk)
{
    uint32_t words = n_symbols * (W + kOccPad);
    words += groups * (kBlobStage + cmax * W + cmax * WS + W);
    words += 2 * cmax * kThreads;
    words = (words + 1) & ~1u;
    words += 2 * units_per_chunk;
    return words;
}

template <int G>
__device__ __forceinline__ Smem carve(uint32_t* base, const DevLibrary& lib, const Plan& plan)
{
    const uint32_t groups = kThreads / G, WP = plan.W + kOccPad;
    Smem           s;
    s.occ     = base;               base += lib.n_symbols * WP;
    s.blob    = base;               base += groups * kBlobStage;
    s.M       = base;               base += groups * lib.cmax * plan.W;
    s.S       = base;               base += groups * lib.cmax * plan.WS;
    s.T       = base;               base += groups * plan.W;
    s.stack_p = base;               base += lib.cmax * kThreads;
    s.stack_t = base;               base += lib.cmax * kThreads;
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
    const uint32_t chunk = blockIdx.x;
    if (EMIT && work.chunk_sites[chunk] == 0) return;   // nothing to write for this chunk

    const uint32_t qi = chunk / plan.chunks_per_seq;
    const uint32_t u0 = (chunk % plan.chunks_per_seq) * plan.units_per_chunk;
    const uint32_t u1 = min(lib.n_units, u0 + plan.units_per_chunk);
    const uint64_t sb = query.seq_begin[qi];
    const int      n  = (int)(query.seq_begin[qi + 1] - sb);
    const uint8_t* seq = query.seqs + sb;
    const uint32_t* pair_mask = query.masks + query.mask_begin[qi];
    const int      W = (int)plan.W, WS = (int)plan.WS, WP = W + kOccPad;
    const int      Wq = (n + 32) >> 5;

    Smem sm = carve<G>(smem_raw, lib, plan);
    build_occurrence(sm.occ, lib, seq, n, W);

    constexpr uint32_t groups = kThreads / G;
    const uint32_t     gid    = threadIdx.x / G;
    uint32_t*          g_blob = sm.blob + gid * kBlobStage;
    uint32_t*          g_M    = sm.M + gid * lib.cmax * W;
    uint32_t*          g_S    = sm.S + gid * lib.cmax * WS;
    uint32_t*          g_T    = sm.T + gid * W;
    const uint64_t     item0  = (uint64_t)qi * lib.n_units;

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

    for (uint32_t u = u0 + gid; u < u1; u += groups) {
        uint32_t count_known = 0;
        if (EMIT) {
            count_known = work.counts[item0 + u];
            if (count_known == 0) continue;
        }
        // stage the unit descriptor
        const uint32_t  boff  = __ldg(lib.unit_off + u);
        const uint32_t  blen  = __ldg(lib.unit_off + u + 1) - boff;
        const uint32_t* gblob = lib.blob + boff;
        const uint32_t* blob  = gblob;
        Gr::sync();
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
            alive              = gate_open<G>(gb, sm.occ, WP, g_M, g_S, W, WS, n);
            Gr::sync();
        }
        if (alive) alive = match_components<G>(blob, sm.occ, WP, g_M, g_S, W, WS, n);

        uint64_t item_sites = 0;
        if (alive) {
            // Component 0: its matches from position 0, leftmost and non-overlapping.
            const uint32_t* first = g_M;
            if (blob[3] & 1u) {
                for (int w = Gr::lane(); w < W; w += G) g_T[w] = 0;
                Gr::sync();
                if (Gr::lane() == 0) {
                    const int k0 = (int)(blob[4] & 0xFFu);
                    for (int q = next_match(g_M, g_S, W, WS, 0); q >= 0; q = next_match(g_M, g_S, W, WS, q + (k0 > 0 ? k0 : 1)))
                        g_T[q >> 5] |= 1u << (q & 31);
                }
                Gr::sync();
                first = g_T;
            }

            ItemCtx cx;
            cx.blob = blob; cx.M = g_M; cx.S = g_S; cx.mask = pair_mask;
            cx.stack_p = sm.stack_p + threadIdx.x; cx.stack_t = sm.stack_t + threadIdx.x;
            cx.W = W; cx.WS = WS; cx.n = n; cx.mstride = (n + 31) >> 5; cx.C = C; cx.delay = delay;

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
    // Enough chunks to give every SM several CTAs, but whole multiples of the groups in a CTA.
    const uint32_t target = (uint32_t)sm_count * 8u;
    uint32_t       per_seq = n_seqs >= target ? 1u : (target + n_seqs - 1) / (n_seqs ? n_seqs : 1);
    uint32_t       upc     = (lib.n_units + per_seq - 1) / (per_seq ? per_seq : 1);
    upc                    = ((upc + groups - 1) / groups) * groups;
    if (upc < groups) upc = groups;
    if (upc > kMaxChunkUnits) upc = kMaxChunkUnits;
    p.units_per_chunk = upc;
    p.chunks_per_seq  = lib.n_units ? (lib.n_units + upc - 1) / upc : 0;
    p.n_chunks        = p.chunks_per_seq * n_seqs;
    p.smem_bytes      = 4u * smem_words(lib.n_symbols, lib.cmax, p.W, p.WS, groups, upc);
    return p;
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
