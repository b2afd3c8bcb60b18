// Window pass of the insertion-site search (included by bss_kernels.cu, inside its anonymous namespace:
// it uses Group, match_components, thin_components, gate_open and load_seq defined there).
//
//   bandmask_kernel  once per query: the pair mask of every narrow-band sequence re-laid as rows of K words that
//                    start at the diagonal word (bss_device.cuh, DevQuery::bmask) — 40 KB instead of 500 KB at
//                    n = 2000, band 99: small enough to be staged in shared memory, and one multiply away;
//   route_kernel     one thread per (sequence, unit) item: k-mer presence filter, then the item goes on the window
//                    list (window-walk unit on a narrow-band sequence), stays with the ordinary passes (bit in
//                    `prefilter`), or is finished (count 0);
//   window_kernel    count: persistent warps draw items from the window list; per item match -> thin -> every lane
//                    walks the candidates of the first component in its two or so mask words (bss_window_walk.h);
//                    the per-lane site counts, the item's count, the alive list and the chunk totals are recorded.
//                    emit: the warps stride over the alive list and write the records where the scanned offsets
//                    say: canonical order, no sort, no output atomics.
// Not a contraction: integer/bit work only, no tensor cores.
#pragma once
// (bss_window_walk.h is included by bss_kernels.cu at file scope)

#ifdef BSS_DEBUG_ASSERT
__device__ int g_win_skip;   // debug build: 1 skip the walk, 2 skip thin + walk, 4 skip match too
#endif

// One item per warp at a time. 4 warps per CTA, 8 CTAs per SM; with staged tables 16 warps per CTA, 2 CTAs per SM,
// so that the 46 KB copy of the query's tables is shared by 16 warps and 32 warps stay resident.
#ifndef BSS_WIN_THREADS
#define BSS_WIN_THREADS 128
#endif
constexpr int kWinThreads = BSS_WIN_THREADS, kWinThreadsStaged = 512;

// 1-D bulk copy global -> shared through the TMA unit, completion on an mbarrier (sm_90+; UBLKCP in SASS).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes),
                 "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(phase)
        : "memory");
}

// Banded re-layout of the pair masks. One warp per row; lane j < K copies word (u >> 5) + j of row u.
__global__ void __launch_bounds__(256) bandmask_kernel(const DevQuery query, uint32_t* __restrict__ bmask)
{
    const uint32_t qi   = blockIdx.x;
    const int      band = query.band[qi];
    const int       K  = band_words(band);
    const uint64_t  sb = query.seq_begin[qi];
    const int       n  = (int)(query.seq_begin[qi + 1] - sb);
    const int       Wm = (n + 31) >> 5;
    const uint32_t* mask = query.masks + query.mask_begin[qi];
    uint32_t*       out  = bmask + (size_t)query.bm_stride * sb;
    const int       warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int u = blockIdx.y * 8 + warp; u < n; u += gridDim.y * 8) {
        for (int j = lane; j < K; j += 32) {   // (band < 0: K = 0, nothing to write)
            const int w = (u >> 5) + j;
            uint32_t  x = w < Wm ? __ldg(mask + (size_t)u * Wm + w) : 0u;
            if (j == 0) x &= (u & 31) == 31 ? 0u : (0xFFFFFFFFu << ((u & 31) + 1));   // strictly above the diagonal
            if (w == Wm - 1 && (n & 31)) x &= (1u << (n & 31)) - 1u;                    // inside the sequence
            out[(size_t)u * K + j] = x;
        }
    }
}

// One thread per item. See the header of this file.
__global__ void __launch_bounds__(256) route_kernel(const DevLibrary lib, const DevQuery query, const Work work, uint64_t n_items, uint32_t plan_dict,
                                                    int window_band)
{
    __shared__ uint32_t warp_win[8], warp_old[8], base_win;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    const uint64_t rounds = (n_items + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; r++) {
        const uint64_t item = r * stride + (uint64_t)blockIdx.x * 256 + threadIdx.x;
        bool alive = false, window = false;
        if (item < n_items) {
            const uint32_t qi = (uint32_t)(item / lib.n_units), u = (uint32_t)(item - (uint64_t)qi * lib.n_units);
            alive = true;
            if (lib.unit_probe && query.kmers) {
                const uint32_t* bits = query.kmers + (size_t)qi * kKmerWords;
                const uint2     pr = __ldg(reinterpret_cast<const uint2*>(lib.unit_probe) + u);   // four u16 probes
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t p = ((i < 2 ? pr.x : pr.y) >> (16 * (i & 1))) & 0xFFFFu;
                    if (p != 0xFFFFu) {
                        const uint32_t code = p & 0xFFFu;
                        if (!((__ldg(bits + kmer_table_word(p >> 12) + (code >> 5)) >> (code & 31)) & 1u)) alive = false;
                    }
                }
            }
            if (alive && plan_dict) {   // pattern table: a component whose pattern does not occur ends the item
                const uint2    sl = __ldg(reinterpret_cast<const uint2*>(lib.unit_slots) + u);
                const uint8_t* fl = work.pflag + (size_t)qi * lib.n_slots;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t slot = ((i < 2 ? sl.x : sl.y) >> (16 * (i & 1))) & 0xFFFFu;
                    if (slot != kNoSlot && !(fl[slot] & 1u)) alive = false;
                }
            }
            if (!alive) work.counts[item] = 0;
            window = alive && (__ldg(lib.unit_flags + u) & kUnitWindow) && __ldg(query.band + qi) <= window_band;
        }
        const uint32_t wbal = __ballot_sync(0xFFFFFFFFu, window), obal = __ballot_sync(0xFFFFFFFFu, alive && !window);
        if (lane == 0 && item < ((n_items + 31) & ~31ull)) work.prefilter[item >> 5] = obal;
        if (lane == 0) { warp_win[warp] = __popc(wbal); warp_old[warp] = __popc(obal); }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t w = 0, o = 0;
            for (int i = 0; i < 8; i++) { const uint32_t c = warp_win[i]; warp_win[i] = w; w += c; o += warp_old[i]; }
            base_win = w ? (uint32_t)atomicAdd((unsigned long long*)(work.totals + 6), (unsigned long long)w) : 0u;
            if (o) atomicAdd((unsigned long long*)(work.totals + 7), (unsigned long long)o);
        }
        __syncthreads();
        if (window) work.wlist[base_win + warp_win[warp] + __popc(wbal & ((1u << lane) - 1u))] = (uint32_t)item;
        __syncthreads();   // warp_win / base_win are reused by the next round
    }
}

// Per warp: descriptor, match sets M and T with the summaries of both, and the walk's scratch (bss_window_walk.h):
// decoded level parameters, saved cursors, one queue of `qcap` entries per level but the last, the stash of `kw` words per lane.
__host__ __device__ inline uint32_t window_smem_words(uint32_t cmax, uint32_t W, uint32_t WS, uint32_t threads, uint32_t qcap, uint32_t kw)
{
    return (threads / 32) * (kBlobStage + 2 * cmax * (W + WS) + (win::kLevelWords + 2 + 6) * cmax + (cmax > 1 ? cmax - 1 : 1) * qcap + 32 * kw);
}

// Occurrence table + banded mask of ONE sequence staged in shared memory (words), when the plan asks for it.
__host__ __device__ inline uint32_t window_stage_words(uint32_t occ_rows, uint32_t occ_stride, uint32_t n, uint32_t K)
{
    return ((occ_rows * occ_stride + 3u) & ~3u) + ((n * K + 3u) & ~3u);
}

// Match masks (as match_components) from an occurrence table that may live in shared memory: generic loads.
__device__ __forceinline__ bool match_components_any(const uint32_t* blob, const uint32_t* occ, int WP, uint32_t* M, uint32_t* S, int W, int WS, int n)
{
    using Gr       = Group<32>;
    const int  C   = (int)(blob[1] & 0xFFu);
    const int  Wq  = (n + 32) >> 5;
    const uint16_t* halves = reinterpret_cast<const uint16_t*>(blob);
    for (int c = 0; c < C; c++) {
        const uint32_t  desc  = blob[4 + 2 * c];
        const int       k     = (int)(desc & 0xFFu);
        const int       steps = (int)((blob[5 + 2 * c] >> 16) & 0xFFu);
        const uint16_t* prog  = halves + ((desc >> 8) & 0xFFFFu);
        const int       last  = n - k;
        bool            any   = false;
        for (int w0 = 0; w0 < W; w0 += 32) {
            const int w = w0 + (int)Gr::lane();
            uint32_t  m = 0;
            if (w < Wq && last >= (w << 5)) {
                const int hi = last - (w << 5);
                m            = hi >= 31 ? 0xFFFFFFFFu : ((2u << hi) - 1u);
#pragma unroll 4
                for (int i = 0; i < steps; i++) {
                    const uint32_t  st  = prog[i];
                    const uint32_t* row = occ + (st & 0x7Fu) * WP + w + (st >> 12);
                    m &= __funnelshift_r(row[0], row[1], (st >> 7) & 31u);
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

// Thinned match sets, as thin_components (bss_kernels.cu), with the two common outcomes on registers and two
// shared-memory words per lane: (1) no two matches of the component overlap on this sequence (nothing to thin:
// most items), (2) matches overlap in pairs only, so that the thinned set is just the run starts. Runs of three
// and more overlapping matches (a greedy walk per run), patterns longer than 32 bytes: the general routine.
// Does not write the summaries ST (the window walk does not read them).
__device__ __forceinline__ uint32_t thin_components_win(const uint32_t* blob, const uint32_t* M, const uint32_t* SM, uint32_t* T, uint32_t* ST, int W, int WS)
{
    using Gr     = Group<32>;
    const int C  = (int)(blob[1] & 0xFFu);
    uint32_t ovl = 0;
    bool     general = false;
    for (int c = 0; c < C && !general; c++) {
        const uint32_t desc = blob[4 + 2 * c];
        if (!(desc & kCompOverlap)) continue;
        const int k = (int)(desc & 0xFFu);
        if (k > 32) { general = true; break; }
        const uint32_t* Mc = M + c * W;
        uint32_t*       Tc = T + c * W;
        bool            any = false;
        for (int w0 = 0; w0 < W; w0 += 32) {   // followers: matches with another match in the k-1 positions before them
            const int      w = w0 + (int)Gr::lane();
            const uint32_t m = w < W ? Mc[w] : 0u, prev = (w > 0 && w < W) ? Mc[w - 1] : 0u;
            uint32_t       near = 0;
            for (int s = 1; s < k; s++) near |= __funnelshift_l(prev, m, s);
            near &= m;
            if (w < W) Tc[w] = near;
            any |= Gr::ballot(near != 0) != 0;
        }
        if (!any) continue;   // no two matches overlap on this sequence: the chain from any start is M itself
        Gr::sync();
        bool deep = false;    // a follower of a follower: a run of three matches or more
        for (int w0 = 0; w0 < W; w0 += 32) {
            const int      w = w0 + (int)Gr::lane();
            const uint32_t t = w < W ? Tc[w] : 0u, prev = (w > 0 && w < W) ? Tc[w - 1] : 0u;
            uint32_t       f2 = 0;
            for (int s = 1; s < k; s++) f2 |= __funnelshift_l(prev, t, s);
            deep |= Gr::ballot((f2 & t) != 0) != 0;
        }
        if (deep) { general = true; break; }
        ovl |= 1u << c;
        Gr::sync();
        for (int w = (int)Gr::lane(); w < W; w += 32) Tc[w] = Mc[w] & ~Tc[w];   // run starts
    }
    Gr::sync();
    if (general) return thin_components<32>(blob, M, SM, T, ST, W, WS);   // (recomputes every component)
    return ovl;
}

// Greedy earliest placement: does the gate unit have any placement on this sequence?
__device__ __forceinline__ bool gate_open_any(const uint32_t* gblob, const uint32_t* occ, int WP, uint32_t* M, uint32_t* S, int W, int WS, int n)
{
    if (!match_components_any(gblob, occ, WP, M, S, W, WS, n)) return false;
    const int C = (int)(gblob[1] & 0xFFu), delay = (int)(gblob[1] >> 16);
    int       t = 0;
    for (int c = 0; c < C; c++) {
        const int q = next_bit(M + c * W, S + c * WS, W, WS, t);
        if (q < 0) return false;
        t = q + (int)(gblob[4 + 2 * c] & 0xFFu) - 1 + delay;
    }
    return true;
}

// Pattern table of one search: one warp per (sequence, slot) computes the match set of the slot's pattern (with its word
// summary), and for patterns that can overlap themselves the thinned set, ONCE — every unit that carries the pattern
// reads the rows instead of matching and thinning again (the headline library holds 50 000 units over a few hundred
// distinct short strands). Same routines as the per-unit path, on the slot's one-component descriptor.
__global__ void __launch_bounds__(kThreads) pattern_kernel(const DevLibrary lib, const DevQuery query, const Plan plan, const Work work)
{
    extern __shared__ __align__(16) uint32_t smem_raw[];
    constexpr uint32_t warps = kThreads / 32;
    const int          W = (int)plan.W, WS = (int)plan.WS, WP = (int)query.occ_stride;
    const uint32_t     warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* g_M  = smem_raw + warp * (2 * (W + WS));
    uint32_t* g_SM = g_M + W;
    uint32_t* g_T  = g_SM + WS;
    uint32_t* g_ST = g_T + W;
    const uint64_t n_tasks = (uint64_t)query.n_seqs * lib.n_slots;
    // first kernel of the search: its first CTAs zero the totals, chunk totals and offsets (no reset kernel of its own)
    constexpr uint32_t kResetCtas = 16;
    if (blockIdx.x < kResetCtas) reset_work(work, plan.n_chunks, query.n_seqs, blockIdx.x * kThreads + threadIdx.x, min(gridDim.x, kResetCtas) * kThreads);
    for (uint64_t task = (uint64_t)blockIdx.x * warps + warp; task < n_tasks; task += (uint64_t)gridDim.x * warps) {
        const uint32_t  qi = (uint32_t)(task / lib.n_slots), slot = (uint32_t)(task - (uint64_t)qi * lib.n_slots);
        const SeqCtx    sq = load_seq(lib, query, qi);
        const uint32_t* blob = lib.pat_blob + __ldg(lib.pat_off + slot);
        uint32_t        flag = 0;
        if (match_components_any(blob, sq.occ, WP, g_M, g_SM, W, WS, sq.n)) {
            flag = 1u;
            uint32_t* out = work.ptab_m + task * (size_t)(W + WS);
            for (int i = (int)lane; i < W + WS; i += 32) out[i] = g_M[i];   // (M and its summary are contiguous)
            if ((blob[4] & kCompOverlap) && (thin_components_win(blob, g_M, g_SM, g_T, g_ST, W, WS) & 1u)) {
                flag |= 2u;
                uint32_t* tout = work.ptab_t + ((size_t)qi * lib.n_tslots + __ldg(lib.pat_tslot + slot)) * (size_t)W;
                for (int i = (int)lane; i < W; i += 32) tout[i] = g_T[i];
            }
        }
        if (lane == 0) work.pflag[task] = (uint8_t)flag;
        __syncwarp();
    }
}

template <bool EMIT, bool STAGE>
#ifndef BSS_WIN_MIN_CTAS
#define BSS_WIN_MIN_CTAS 8
#endif
__global__ void __launch_bounds__(STAGE ? kWinThreadsStaged : kWinThreads, STAGE ? 2 : BSS_WIN_MIN_CTAS)
    window_kernel(const DevLibrary lib, const DevQuery query, const Plan plan, const Work work, uint32_t* __restrict__ records, uint64_t capacity_words)
{
    constexpr int THREADS = STAGE ? kWinThreadsStaged : kWinThreads;
    using Gr = Group<32>;
    const uint32_t n_list = (uint32_t)work.totals[EMIT ? 5 : 6];
    if (EMIT && (n_list == 0 || work.totals[0] > capacity_words)) return;   // nothing to write / the records would not fit: nothing is written
    extern __shared__ __align__(16) uint32_t smem_raw[];
    constexpr uint32_t warps = THREADS / 32;
    const int          W = (int)plan.W, WS = (int)plan.WS;
    const uint32_t     warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    uint32_t* base  = smem_raw;
    // staged tables first (16-byte aligned for the bulk copies)
    const uint32_t* s_occ = nullptr;
    const uint32_t* s_bm  = nullptr;
    if (STAGE) {
        __shared__ __align__(8) uint64_t bar;
        const int      n0 = (int)(query.seq_begin[1] - query.seq_begin[0]), K0 = band_words(query.band[0]);
        const uint32_t occ_words = query.occ_rows * query.occ_stride, bm_words = (uint32_t)(n0 * K0);
        uint32_t*      d_occ = base;
        uint32_t*      d_bm  = base + ((occ_words + 3u) & ~3u);
        base += window_stage_words(query.occ_rows, query.occ_stride, (uint32_t)n0, (uint32_t)K0);
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            const uint32_t b0 = (occ_words * 4u) & ~15u, b1 = (bm_words * 4u) & ~15u;
            mbar_expect_tx(&bar, b0 + b1);
            if (b0) bulk_g2s(d_occ, query.occ, b0, &bar);
            if (b1) bulk_g2s(d_bm, query.bmask, b1, &bar);
        }
        // the tails that are not a multiple of 16 bytes
        for (uint32_t i = (occ_words & ~3u) + threadIdx.x; i < occ_words; i += THREADS) d_occ[i] = query.occ[i];
        for (uint32_t i = (bm_words & ~3u) + threadIdx.x; i < bm_words; i += THREADS) d_bm[i] = query.bmask[i];
        __syncthreads();   // the barrier is initialised before anyone waits on it
        mbar_wait(&bar, 0);
        __syncthreads();
        s_occ = d_occ;
        s_bm  = d_bm;
    }
    uint32_t* g_blob = base + warp * kBlobStage;                       base += warps * kBlobStage;
    uint32_t* g_M    = base + warp * lib.cmax * W;                     base += warps * lib.cmax * W;
    uint32_t* g_T    = base + warp * lib.cmax * W;                     base += warps * lib.cmax * W;
    uint32_t* g_SM   = base + warp * lib.cmax * WS;                    base += warps * lib.cmax * WS;
    uint32_t* g_ST   = base + warp * lib.cmax * WS;                    base += warps * lib.cmax * WS;
    const uint32_t qcap = plan.window_qcap, kw = plan.window_kw, qlevels = lib.cmax > 1 ? lib.cmax - 1 : 1;
    int*      g_lp    = reinterpret_cast<int*>(base) + warp * win::kLevelWords * lib.cmax;   base += warps * win::kLevelWords * lib.cmax;
    uint32_t* g_ctl   = base + warp * 2 * lib.cmax;                                          base += warps * 2 * lib.cmax;
    uint32_t* g_Q     = base + warp * qlevels * qcap;                                        base += warps * qlevels * qcap;
    uint32_t* g_stash = base + warp * 32 * kw;                                               base += warps * 32 * kw;
    base += (base - smem_raw) & 1u;   // (8-byte alignment of the row pointers)
    const uint32_t** g_rows = reinterpret_cast<const uint32_t**>(base) + warp * 3 * lib.cmax;      base += warps * 6 * lib.cmax;
#ifdef BSS_DEBUG_ASSERT
    if ((g_win_skip & 64) && blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        printf("[win dbg] EMIT %d dyn smem %u B, plan says %u B, used %u B, cmax %u W %d WS %d n_list %u\n", (int)EMIT, dyn, plan.smem_bytes_window,
               (uint32_t)(base - smem_raw) * 4u, lib.cmax, W, WS, n_list);
    }
#endif

    // Count pass: the warps draw blocks of plan.window_draw consecutive list entries from one cursor. Measured on the
    // headline workload, one entry per draw is best (0.146 ms; 4 per draw 0.166, 32 per draw 0.53 ms): the atomic's
    // latency hides behind the item in hand, while larger blocks leave warps without work at the end.
    unsigned int*  cursor = reinterpret_cast<unsigned int*>(work.totals + 8);
    const uint32_t draw   = plan.window_draw;
    uint32_t       block_end = 0;   // count pass: one past the last slot of the block in hand
    uint32_t            slot;
    if (EMIT) {
        slot = blockIdx.x * warps + warp;
    } else {
        slot = 0;
        if (lane == 0) slot = atomicAdd(cursor, draw);
        slot      = Gr::first(slot);
        block_end = slot + draw;
    }
    const bool one_seq = query.n_seqs == 1;
    SeqCtx     sq      = load_seq(lib, query, 0);
    if (!EMIT && n_list == 0) return;
    while (slot < n_list) {
        uint32_t next_slot = 0;
        if (!EMIT && lane == 0 && slot + 1 == block_end) next_slot = atomicAdd(cursor, draw);   // drawn ahead: its latency hides behind this item
        const uint32_t item = EMIT ? work.alive_w[slot] : work.wlist[slot];
        if (!BSS_WIN_CHECK((uint64_t)item < (uint64_t)query.n_seqs * lib.n_units, 21)) break;
        uint32_t qi = 0, u = item;
        if (!one_seq) {   // (one-sequence queries: no division, the sequence's context was loaded before the loop)
            qi = item / lib.n_units;
            u  = item - qi * lib.n_units;
            sq = load_seq(lib, query, qi);
        }
        const int n = sq.n;

        uint64_t out_words = 0;
        if (EMIT) {   // records of the chunk's earlier items
            const uint32_t chunk = qi * plan.chunks_per_seq + u / plan.units_per_chunk;
            const uint32_t v0 = (u / plan.units_per_chunk) * plan.units_per_chunk;
            uint64_t       before = 0;
            for (uint32_t v = v0 + lane; v < u; v += 32) before += (uint64_t)work.counts[sq.item0 + v] * __ldg(lib.unit_rlen + v);
            out_words = work.chunk_words[chunk] + Gr::sum(before);
        }

        // stage the unit descriptor
        const uint32_t  boff  = __ldg(lib.unit_off + u);
        const uint32_t  blen  = __ldg(lib.unit_off + u + 1) - boff;
        const uint32_t* gblob = lib.blob + boff;
        if (!BSS_WIN_CHECK(blen >= 6 && blen <= kBlobStage, 22)) break;
        // (window-walk units always fit the staging area: the library builder only flags those that do, so every
        //  descriptor read of the walk is a shared-memory load)
        // (streamed once per search: evict-first, so that the descriptors do not push the pattern-table rows and the banded mask out of L1)
        for (uint32_t i = lane; i < blen; i += 32) g_blob[i] = __ldcs(gblob + i);
        const uint32_t* blob = g_blob;
        __syncwarp();
        const uint32_t module = blob[0] + lib.module_base;
        const int      C      = (int)(blob[1] & 0xFFu);
        const uint32_t gate   = blob[2];
        if (!BSS_WIN_CHECK(C >= 1 && C <= (int)lib.cmax && n >= 0 && n <= (int)query.max_len, 23)) break;
        const uint32_t* occ   = STAGE ? s_occ : sq.occ;
        const int       WP    = (int)query.occ_stride;

        bool alive = true;
#ifdef BSS_DEBUG_ASSERT
        if (g_win_skip & 4) alive = false;
        if ((g_win_skip >> 8) && u != (uint32_t)(g_win_skip >> 8) - 1u) alive = false;   // one unit only
#endif
        // Lean units: every pattern has a row in the pattern table of this search (pattern_kernel): no matching, no thinning
        // here, the walk reads the table. (The route kernel has already dropped the items one of whose patterns does not occur.)
        const bool lean = plan.dict && (__ldg(lib.unit_flags + u) & kUnitLean);
        uint32_t   ovl  = 0;
        if (lean) {
            uint32_t flag = 0;
            if ((int)lane < C) {
                const uint32_t slot = __ldg(lib.unit_slots + 4 * (size_t)u + lane);
                const size_t   at   = (size_t)qi * lib.n_slots + slot;
                flag                = work.pflag[at];
                const uint32_t* m   = work.ptab_m + at * (size_t)(W + WS);
                g_rows[3 * lane]     = m;
                g_rows[3 * lane + 1] = (flag & 2u) ? work.ptab_t + ((size_t)qi * lib.n_tslots + __ldg(lib.pat_tslot + slot)) * (size_t)W : m;
                g_rows[3 * lane + 2] = m + W;
            }
            ovl = __ballot_sync(0xFFFFFFFFu, (flag & 2u) != 0);
        } else {
            if (!EMIT && gate != kNoGate) {   // items on the alive list have passed their gate
                alive = gate_open_any(lib.blob + __ldg(lib.unit_off + gate), occ, WP, g_M, g_SM, W, WS, n);
                __syncwarp();
            }
            if (alive) alive = match_components_any(blob, occ, WP, g_M, g_SM, W, WS, n);
        }

        uint64_t item_sites = 0;
#ifdef BSS_DEBUG_ASSERT
        if (g_win_skip & 2) alive = false;
#endif
        if (alive) {
            win::Walk cx;
            if (!lean) {
                ovl = thin_components_win(blob, g_M, g_SM, g_T, g_ST, W, WS);
                if ((int)lane < C) {
                    g_rows[3 * lane]     = g_M + lane * W;
                    g_rows[3 * lane + 1] = g_T + lane * W;
                    g_rows[3 * lane + 2] = g_SM + lane * WS;
                }
            }
            __syncwarp();
            cx.blob = blob; cx.rows = g_rows;
            cx.band = sq.band;
            cx.K    = min(band_words(sq.band), (int)kw);
            cx.bm   = STAGE ? s_bm : query.bmask + (size_t)query.bm_stride * query.seq_begin[qi];
            cx.lp = g_lp; cx.ctl = g_ctl; cx.Q = g_Q; cx.stash = g_stash; cx.qcap = (int)qcap;
            cx.ovl  = ovl;
            cx.W = W; cx.WS = WS; cx.n = n; cx.C = C; cx.delay = (int)(blob[1] >> 16);
#ifdef BSS_DEBUG_ASSERT
            if (g_win_skip & 1) { __syncwarp(); goto skip_walk; }
#endif
            // one warp walks the item breadth-first by level (bss_window_walk.h); the emit pass walks it again and writes
            if (EMIT) {
                const uint64_t wrote = win::walk<true>(cx, sq.Wq, module, records + out_words);
                (void)BSS_WIN_CHECK(wrote == work.counts[sq.item0 + u], 25);
            } else {
                item_sites = win::walk<false>(cx, sq.Wq, module, nullptr);
            }
        }
#ifdef BSS_DEBUG_ASSERT
    skip_walk:
        if (g_win_skip & 8) item_sites = 0;   // walk, but drop what it found
#endif
        __syncwarp();   // the warp's scratch is reused by its next item
        if (!EMIT) {
            if (lane == 0) {
                if (item_sites > 0xFFFFFFFFull) atomicOr((unsigned long long*)(work.totals + 2), 1ull);
                work.counts[sq.item0 + u] = (uint32_t)item_sites;
                if (item_sites) {
                    const uint32_t chunk = qi * plan.chunks_per_seq + u / plan.units_per_chunk;
                    const uint32_t where = (uint32_t)atomicAdd((unsigned long long*)(work.totals + 5), 1ull);
                    work.alive_w[where]  = item;
                    atomicAdd((unsigned long long*)&work.chunk_sites[chunk], (unsigned long long)item_sites);
                    atomicAdd((unsigned long long*)&work.chunk_words[chunk], (unsigned long long)(item_sites * (uint64_t)(C + 1)));
                }
            }
            if (slot + 1 == block_end) {   // (uniform)
                slot      = Gr::first(next_slot);
                block_end = slot + draw;
            } else {
                slot++;
            }
        } else {
            slot += gridDim.x * warps;
        }
    }
}

// ----------------------------------------------------------------------------------------
// placements_kernel (diagnostic, not part of a search): the exact number of placements the reference's
// enumerator materialises BEFORE its filters — the size of the vector find_next_ones_in returns
// (cppsrc/Motif.cpp:466-474), summed over every (sequence, unit) item whose gate is open. This is the
// reference's real work term; the search kernels never visit most of these placements (band windows,
// row filters, rank-space pruning). One warp per item; backward over the components:
//   N_c(q)  = placements of components c.. below a match of component c at q
//   A_c(t)  = sum of N_c over the thinned matches at positions >= t
// and the children of q are the reference's chain in the suffix starting at t = q + k_c - 1 + delay: a few
// greedy matches, then the thinned set (A). Per-warp scratch in global memory: 2 levels x (N, A).
// ----------------------------------------------------------------------------------------
__host__ __device__ inline size_t placements_scratch_words(uint32_t max_len) { return 4 * ((size_t)max_len + 2); }   // u64 words per warp

__global__ void __launch_bounds__(kThreads) placements_kernel(const DevLibrary lib, const DevQuery query, const Plan plan, uint64_t* __restrict__ scratch,
                                                              size_t per_warp, unsigned long long* __restrict__ total_out)
{
    using Gr = Group<32>;
    extern __shared__ __align__(16) uint32_t smem_raw[];
    constexpr uint32_t warps = kThreads / 32;
    const int          W = (int)plan.W, WS = (int)plan.WS, WP = (int)query.occ_stride;
    const uint32_t     warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* base   = smem_raw;
    uint32_t* g_blob = base + warp * kBlobStage;      base += warps * kBlobStage;
    uint32_t* g_M    = base + warp * lib.cmax * W;    base += warps * lib.cmax * W;
    uint32_t* g_T    = base + warp * lib.cmax * W;    base += warps * lib.cmax * W;
    uint32_t* g_SM   = base + warp * lib.cmax * WS;   base += warps * lib.cmax * WS;
    uint32_t* g_ST   = base + warp * lib.cmax * WS;
    uint64_t* mine   = scratch + (size_t)(blockIdx.x * warps + warp) * per_warp;
    const uint64_t n_items = (uint64_t)query.n_seqs * lib.n_units;
    unsigned long long sum = 0;
    for (uint64_t item = (uint64_t)blockIdx.x * warps + warp; item < n_items; item += (uint64_t)gridDim.x * warps) {
        const uint32_t qi = (uint32_t)(item / lib.n_units), u = (uint32_t)(item - (uint64_t)qi * lib.n_units);
        const SeqCtx   sq = load_seq(lib, query, qi);
        const int      n  = sq.n;
        const uint32_t  boff  = __ldg(lib.unit_off + u);
        const uint32_t  blen  = __ldg(lib.unit_off + u + 1) - boff;
        const uint32_t* blob  = lib.blob + boff;
        if (blen <= kBlobStage) {
            for (uint32_t i = lane; i < blen; i += 32) g_blob[i] = __ldg(blob + i);
            blob = g_blob;
            __syncwarp();
        }
        const int      C = (int)(blob[1] & 0xFFu), delay = (int)(blob[1] >> 16);
        const uint32_t gate = blob[2];
        bool alive = true;
        if (gate != kNoGate) {
            alive = gate_open<32>(lib.blob + __ldg(lib.unit_off + gate), sq.occ, WP, g_M, g_SM, W, WS, n);
            __syncwarp();
        }
        if (alive) alive = match_components<32>(blob, sq.occ, WP, g_M, g_SM, W, WS, n);
        if (alive) {
            const uint32_t ovl = thin_components<32>(blob, g_M, g_SM, g_T, g_ST, W, WS);
            const int      P   = n + 1;                    // positions 0 .. n (an empty pattern matches at n)
            const int      per = (P + 31) >> 5;            // positions per lane
            const int      p0 = min(P, (int)lane * per), p1 = min(P, p0 + per);
            uint64_t*      Nc = mine, *Ac = mine + (n + 2), *Nn = mine + 2 * (size_t)(n + 2), *An = mine + 3 * (size_t)(n + 2);
            for (int c = C - 1; c >= 0; c--) {
                const int       kc = (int)(blob[4 + 2 * c] & 0xFFu);
                const uint32_t* Mc = g_M + c * W;
                const bool      thinned = (ovl >> c) & 1u;
                const uint32_t* Sc = (thinned ? g_T : g_M) + c * W;
                // N_c at every match of the component
                for (int q = (int)lane; q < P; q += 32) {
                    if (!test_bit(Mc, q)) continue;
                    uint64_t v = 1;
                    if (c < C - 1) {
                        const int t = q + kc - 1 + delay;
                        v = 0;
                        if (t >= 0 && t < n) {   // (t < 0: the unsigned wrap of cppsrc/Motif.cpp:461 reads as "no room")
                            const int d = c + 1;
                            if ((ovl >> d) & 1u) {
                                const uint32_t *Md = g_M + d * W, *SMd = g_SM + d * WS, *Td = g_T + d * W;
                                const int       ke = max(1, (int)(blob[4 + 2 * d] & 0xFFu));
                                int             x  = next_bit(Md, SMd, W, WS, t);
                                while (x >= 0 && !test_bit(Td, x)) {
                                    v += Nn[x];
                                    x = next_bit(Md, SMd, W, WS, x + ke);
                                }
                                if (x >= 0) v += An[x];
                            } else {
                                v = An[t];
                            }
                        }
                    }
                    Nc[q] = v;
                }
                __syncwarp();
                // A_c = suffix sums of N_c over the chain set: block sums, a scan over the lanes, block sweeps
                uint64_t block = 0;
                for (int q = p0; q < p1; q++)
                    if (test_bit(Sc, q)) block += Nc[q];
                uint64_t       all;
                const uint64_t before = Gr::exclusive(block, all);   // sum of the blocks of lower lanes
                uint64_t       run    = all - before - block;        // sum of the blocks of higher lanes
                for (int q = p1 - 1; q >= p0; q--) {
                    if (test_bit(Sc, q)) run += Nc[q];
                    Ac[q] = run;
                }
                if (lane == 0) {
                    Ac[P] = 0;
                    // an empty pattern at position 0 with delay 0 leaves "no room" and ends the level (unsigned wrap)
                    if (c < C - 1 && kc == 0 && delay == 0) Ac[0] = 0;
                }
                __syncwarp();
                uint64_t* t0 = Nc; Nc = Nn; Nn = t0;
                t0 = Ac; Ac = An; An = t0;
            }
            if (lane == 0) sum += An[0];   // (after the last swap the first component's sums are in An)
        }
        __syncwarp();
    }
    if (lane == 0 && sum) atomicAdd(total_out, sum);
}
