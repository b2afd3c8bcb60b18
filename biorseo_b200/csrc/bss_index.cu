// Insertion-variable index: what the consumer of the site list — the variable and constraint
// construction of the MOIP constructor — needs from it, built on the device right behind a
// search (row "next" #1 of the path).
//
// What it replaces (reference, CPU, all host loops over insertion_sites_):
//   cppsrc/MOIP.cpp:302-328   C_{x,i,p} variables: index_of_first_components / index_of_Cxip_
//                             (one variable per component of every site, numbered in site order)
//   cppsrc/MOIP.cpp:507-566   "c3": per variable, the y_{u,v} of every allowed pair that touches
//                             a nucleotide of the component (jsonfolder / rinfolder: the whole
//                             component, the site's own links left out; descfolder: the interior
//                             nucleotides) — O(sites * C * k * n * links) on the host
//   cppsrc/MOIP.cpp:569-586   "c4": per nucleotide u, every variable whose component covers u
//                             — O(n * sites * C) on the host
//   cppsrc/MOIP.cpp:74-90, 467-487   y_{u,v} numbering (index_of_yuv_) and, per nucleotide, its
//                             allowed partners ("c1"), which also are the term lists of c3
// Here: the variables table (first, last, number of c3 terms), the per-nucleotide overlap lists
// as a CSR in variable order, and the pair adjacency as a CSR in partner order. Everything is
// integer work on streams of u32 — HBM/L2-bound, no tensor cores.
//
// The site records are not self-delimiting, so the builder reads the per-(sequence, unit) site
// counts the search left in the library's work buffers: it has to run before the next search.
#include "bss_device.cuh"

#include <algorithm>

namespace bss {
namespace {

constexpr int kScanThreads = 1024;

// Exclusive scan of n values by one CTA of kScanThreads threads: thread t owns the contiguous
// slice [t * per, (t + 1) * per). `value(i)` yields element i, `store(i, prefix)` receives the
// exclusive prefix of element i; the return value is the total (valid in every thread).
template <typename T, typename Value, typename Store>
__device__ T block_exclusive_scan(uint32_t n, Value value, Store store)
{
    __shared__ unsigned long long part[kScanThreads];
    const uint32_t per = (n + kScanThreads - 1) / kScanThreads, t = threadIdx.x;
    const uint32_t i0 = min(n, t * per), i1 = min(n, i0 + per);
    T              sum = 0;
    for (uint32_t i = i0; i < i1; i++) sum += value(i);
    part[t] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < kScanThreads; d <<= 1) {   // Hillis-Steele over the slice sums
        const unsigned long long x = t >= d ? part[t - d] : 0ull;
        __syncthreads();
        part[t] += x;
        __syncthreads();
    }
    T run = (T)(part[t] - sum);
    for (uint32_t i = i0; i < i1; i++) {
        const T v = value(i);
        store(i, run);
        run += v;
    }
    const T total = (T)part[kScanThreads - 1];
    __syncthreads();
    return total;
}

// Sites and variables before every unit of the sequence (from the search's per-item counts).
__global__ void __launch_bounds__(kScanThreads) idx_unit_scan_kernel(const uint32_t* __restrict__ counts, const uint8_t* __restrict__ unit_rlen,
                                                                      uint32_t n_units, uint64_t* __restrict__ unit_site_begin,
                                                                      uint64_t* __restrict__ unit_var_begin, uint64_t* __restrict__ totals)
{
    const uint64_t sites = block_exclusive_scan<uint64_t>(
        n_units, [&](uint32_t u) { return (uint64_t)counts[u]; }, [&](uint32_t u, uint64_t p) { unit_site_begin[u] = p; });
    const uint64_t vars = block_exclusive_scan<uint64_t>(
        n_units, [&](uint32_t u) { return (uint64_t)counts[u] * (uint32_t)(unit_rlen[u] - 1); }, [&](uint32_t u, uint64_t p) { unit_var_begin[u] = p; });
    if (threadIdx.x == 0) {
        unit_site_begin[n_units] = sites;
        unit_var_begin[n_units]  = vars;
        totals[0]                = sites;
        totals[1]                = vars;
    }
}

// Allowed pairs per row (u < v), from the pair mask: one warp per row, popcount of the words right of the diagonal.
// degree[u] starts as the row's count; the column kernel adds the partners below u.
__global__ void __launch_bounds__(256) idx_row_pairs_kernel(const uint32_t* __restrict__ mask, uint32_t n, uint32_t Wm, uint32_t* __restrict__ row_pairs,
                                                            uint32_t* __restrict__ degree)
{
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t u = blockIdx.x * 8 + (threadIdx.x >> 5); u < n; u += gridDim.x * 8) {
        uint32_t c = 0;
        for (uint32_t w = (u >> 5) + lane; w < Wm; w += 32) {
            uint32_t bits = mask[(uint64_t)u * Wm + w];
            if ((w << 5) <= u) bits &= (u & 31) == 31 ? 0u : ~0u << ((u & 31) + 1);      // keep v > u
            if (((w + 1) << 5) > n) bits &= (n & 31) ? (1u << (n & 31)) - 1u : ~0u;     // and v < n
            c += __popc(bits);
        }
        for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
        if (lane == 0) {
            row_pairs[u] = c;
            atomicAdd(degree + u, c);
        }
    }
}

// Partners below every nucleotide (column v above the diagonal): thread = column, CTA = 256 columns x 64 rows; the 32
// lanes of a warp read the same word of a row. One atomic per (row chunk, column) instead of one per pair.
__global__ void __launch_bounds__(256) idx_col_pairs_kernel(const uint32_t* __restrict__ mask, uint32_t n, uint32_t Wm, uint32_t* __restrict__ degree)
{
    const uint32_t v = blockIdx.x * 256 + threadIdx.x, u0 = blockIdx.y * 64;
    if (v >= n) return;
    const uint32_t u1 = min(min(u0 + 64, n), v);
    uint32_t       c  = 0;
    for (uint32_t u = u0; u < u1; u++) c += (mask[(uint64_t)u * Wm + (v >> 5)] >> (v & 31)) & 1u;
    if (c) atomicAdd(degree + v, c);
}

__global__ void __launch_bounds__(kScanThreads) idx_pair_scan_kernel(const uint32_t* __restrict__ row_pairs, const uint32_t* __restrict__ degree, uint32_t n,
                                                                      uint32_t* __restrict__ row_first_pair, uint32_t* __restrict__ pair_begin,
                                                                      uint64_t* __restrict__ totals)
{
    const uint32_t pairs = block_exclusive_scan<uint32_t>(
        n, [&](uint32_t u) { return row_pairs[u]; }, [&](uint32_t u, uint32_t p) { row_first_pair[u] = p; });
    const uint32_t ends = block_exclusive_scan<uint32_t>(
        n, [&](uint32_t u) { return degree[u]; }, [&](uint32_t u, uint32_t p) { pair_begin[u] = p; });
    if (threadIdx.x == 0) {
        row_first_pair[n] = pairs;
        pair_begin[n]     = ends;
        totals[2]         = ends;
    }
}

// Partners of every nucleotide in increasing order: column u above the diagonal, then row u.
// One warp per nucleotide, ballot compaction keeps the order.
__global__ void __launch_bounds__(256) idx_adjacency_kernel(const uint32_t* __restrict__ mask, uint32_t n, uint32_t Wm, const uint32_t* __restrict__ pair_begin,
                                                            uint32_t* __restrict__ partner)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    for (uint32_t u = blockIdx.x * 8 + warp; u < n; u += gridDim.x * 8) {
        uint32_t* out = partner + pair_begin[u];
        uint32_t  pos = 0;
        for (uint32_t v0 = 0; v0 < u; v0 += 128) {   // column u above the diagonal: four independent loads per lane, then four ballots
            bool on[4];
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                const uint32_t v = v0 + 32 * i + lane;
                on[i] = v < u && ((mask[(uint64_t)v * Wm + (u >> 5)] >> (u & 31)) & 1u);
            }
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                const uint32_t all = __ballot_sync(0xFFFFFFFFu, on[i]);
                if (on[i]) out[pos + __popc(all & lt)] = v0 + 32 * i + lane;
                pos += __popc(all);
            }
        }
        for (uint32_t w0 = u >> 5; w0 < Wm; w0 += 32) {   // row u: 32 words per coalesced load, handed round by shuffles
            const uint32_t word = w0 + lane < Wm ? mask[(uint64_t)u * Wm + w0 + lane] : 0u;
            const uint32_t cnt  = min(32u, Wm - w0);
            for (uint32_t i = 0; i < cnt; i++) {
                const uint32_t bits = __shfl_sync(0xFFFFFFFFu, word, i);
                const uint32_t v    = ((w0 + i) << 5) + lane;
                const bool     set  = v > u && v < n && ((bits >> lane) & 1u);
                const uint32_t all  = __ballot_sync(0xFFFFFFFFu, set);
                if (set) out[pos + __popc(all & lt)] = v;
                pos += __popc(all);
            }
        }
    }
}

__device__ __forceinline__ int end_position(uint32_t word, const uint32_t* p)
{
    const int anchor = (int)(int8_t)(word & 0xFFu), off = (int)(int16_t)((word >> 8) & 0xFFFFu);
    return anchor < 0 ? off : (int)p[anchor] + off;
}

// The variables of every site: one thread per site.
__global__ void __launch_bounds__(128) idx_vars_kernel(DevLibrary lib, const uint64_t* __restrict__ unit_site_begin, const uint64_t* __restrict__ unit_var_begin,
                                                       const uint32_t* __restrict__ records, uint32_t n_sites, uint32_t n_vars, const uint32_t* __restrict__ mask,
                                                       uint32_t n, uint32_t Wm, const uint32_t* __restrict__ pair_begin, int c3_rule,
                                                       uint32_t* __restrict__ var_begin, uint32_t* __restrict__ var_first, uint32_t* __restrict__ var_second,
                                                       uint32_t* __restrict__ var_c3, uint64_t* __restrict__ totals)
{
    for (uint32_t s = blockIdx.x * 128 + threadIdx.x; s <= n_sites; s += gridDim.x * 128) {
        if (s == n_sites) {
            var_begin[s] = n_vars;
            break;
        }
        // the unit that owns site s: last unit whose first site is <= s (units without sites share a boundary)
        uint32_t lo = 0, hi = lib.n_units;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (unit_site_begin[mid] <= s) lo = mid; else hi = mid;
        }
        const uint32_t  unit = lo;
        const uint32_t* d    = lib.blob + lib.unit_off[unit];
        const uint32_t  C = d[1] & 0xFFu, NK = (d[1] >> 8) & 0xFFu;
        const uint64_t  local = s - unit_site_begin[unit];
        const uint32_t* rec   = records + unit_site_begin[unit] + unit_var_begin[unit] + local * (C + 1);
        const uint32_t  v0    = (uint32_t)(unit_var_begin[unit] + local * C);
        if (rec[0] != d[0] + lib.module_base || (uint32_t)lib.unit_rlen[unit] != C + 1) {
            atomicOr((unsigned long long*)(totals + 4), 1ull);   // these records are not the last search's
            continue;
        }
        uint32_t p[16], excluded[16];
        for (uint32_t c = 0; c < C; c++) {
            p[c]        = rec[1 + c];
            excluded[c] = 0;
        }
        if (c3_rule == 0) {
            // the site's own links are left out of c3 (cppsrc/MOIP.cpp:524-537): every distinct allowed link
            // takes one term away from each component that holds one of its ends
            const uint32_t* chk = d + (d[3] >> 16);
            for (uint32_t i = 0; i < NK; i++) {
                const int a = end_position(chk[2 * i], p), b = end_position(chk[2 * i + 1], p);
                const int u = min(a, b), v = max(a, b);
                if (u < 0 || v >= (int)n || u == v) continue;
                if (!((mask[(uint64_t)u * Wm + (v >> 5)] >> (v & 31)) & 1u)) continue;
                bool seen = false;
                for (uint32_t j = 0; j < i && !seen; j++) {
                    const int a2 = end_position(chk[2 * j], p), b2 = end_position(chk[2 * j + 1], p);
                    seen = min(a2, b2) == u && max(a2, b2) == v;
                }
                if (seen) continue;
                for (uint32_t c = 0; c < C; c++) {
                    const uint32_t k = d[4 + 2 * c] & 0xFFu;
                    if ((uint32_t)u - p[c] < k) excluded[c]++;
                    if ((uint32_t)v - p[c] < k) excluded[c]++;
                }
            }
        }
        var_begin[s] = v0;
        for (uint32_t c = 0; c < C; c++) {
            const uint32_t k = d[4 + 2 * c] & 0xFFu;
            var_first[v0 + c]  = p[c];
            var_second[v0 + c] = p[c] + k - 1u;   // unsigned like Component::pos (an empty component ends before it starts)
            uint32_t a = c3_rule == 0 ? p[c] : p[c] + 1u, b = c3_rule == 0 ? p[c] + k : p[c] + k - 1u;
            if (c3_rule != 0 && k < 3) b = a;
            a = min(a, n);
            b = min(max(a, b), n);
            var_c3[v0 + c] = pair_begin[b] - pair_begin[a] - excluded[c];
        }
    }
}

// Overlap lists, pass 1: CTA b (one warp) owns the variables [b * per, (b + 1) * per) and counts, for
// every nucleotide, how many of them cover it (difference array in shared memory, then a running sum).
__global__ void __launch_bounds__(32) idx_overlap_hist_kernel(const uint32_t* __restrict__ var_first, const uint32_t* __restrict__ var_second, uint32_t n_vars,
                                                              uint32_t per, uint32_t n, uint32_t* __restrict__ hist)
{
    extern __shared__ int cover[];   // [n + 1]
    const uint32_t lane = threadIdx.x, b = blockIdx.x;
    for (uint32_t u = lane; u <= n; u += 32) cover[u] = 0;
    __syncwarp();
    const uint32_t v0 = b * per, v1 = min(n_vars, v0 + per);
    for (uint32_t v = v0 + lane; v < v1; v += 32) {
        const uint32_t f = var_first[v], k = var_second[v] - f + 1u;
        if (k && f < n) {
            atomicAdd(cover + f, 1);
            atomicAdd(cover + min(n, f + k), -1);
        }
    }
    __syncwarp();
    int carry = 0;
    for (uint32_t u0 = 0; u0 < n; u0 += 32) {
        const uint32_t u = u0 + lane;
        int            x = u < n ? cover[u] : 0;
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if ((int)lane >= d) x += y;
        }
        x += carry;
        if (u < n) hist[(uint64_t)b * n + u] = (uint32_t)x;
        carry = __shfl_sync(0xFFFFFFFFu, x, 31);
    }
}

// Pass 2: per nucleotide, exclusive scan over the slices' counts (in place) and the nucleotide's total. A CTA takes 8
// neighbouring nucleotides (one 32-byte sector per slice row): thread = (nucleotide, one of 32 runs of consecutive
// slices). Every thread sums its run (independent loads), the 32 run sums of a nucleotide are scanned in shared
// memory, and a second sweep over the run (now cached) writes the exclusive prefixes.
__global__ void __launch_bounds__(256) idx_overlap_colscan_kernel(uint32_t* __restrict__ hist, uint32_t n_blocks, uint32_t n, uint32_t* __restrict__ total)
{
    __shared__ uint32_t run_sum[32][8];
    const uint32_t j = threadIdx.x & 7, r = threadIdx.x >> 3, u = blockIdx.x * 8 + j;
    const uint32_t per = (n_blocks + 31) / 32, b0 = min(n_blocks, r * per), b1 = min(n_blocks, b0 + per);
    uint32_t       sum = 0;
    if (u < n) {
#pragma unroll 8
        for (uint32_t b = b0; b < b1; b++) sum += hist[(uint64_t)b * n + u];
    }
    run_sum[r][j] = sum;
    __syncthreads();
    if (threadIdx.x < 8) {   // exclusive scan of the 32 run sums of nucleotide j
        uint32_t acc = 0;
        for (uint32_t i = 0; i < 32; i++) {
            const uint32_t v = run_sum[i][threadIdx.x];
            run_sum[i][threadIdx.x] = acc;
            acc += v;
        }
        if (blockIdx.x * 8 + threadIdx.x < n) total[blockIdx.x * 8 + threadIdx.x] = acc;
    }
    __syncthreads();
    if (u < n) {
        uint32_t run = run_sum[r][j];
        for (uint32_t b = b0; b < b1; b += 4) {   // four loads in flight, then the four stores
            uint32_t t[4];
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) t[i] = b + i < b1 ? hist[(uint64_t)(b + i) * n + u] : 0u;
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                if (b + i < b1) hist[(uint64_t)(b + i) * n + u] = run;
                run += t[i];
            }
        }
    }
}

__global__ void __launch_bounds__(kScanThreads) idx_overlap_scan_kernel(const uint32_t* __restrict__ total, uint32_t n, uint32_t* __restrict__ overlap_begin,
                                                                         uint64_t* __restrict__ totals)
{
    const uint64_t entries = block_exclusive_scan<uint64_t>(
        n, [&](uint32_t u) { return (uint64_t)total[u]; }, [&](uint32_t u, uint64_t p) { overlap_begin[u] = (uint32_t)p; });
    if (threadIdx.x == 0) {
        overlap_begin[n] = (uint32_t)entries;
        totals[3]        = entries;
    }
}

// Pass 3: the same CTAs walk their variables in order and append each one to the list of every nucleotide
// it covers (lanes over the nucleotides of a component), so every list comes out in variable order.
__global__ void __launch_bounds__(32) idx_overlap_fill_kernel(const uint32_t* __restrict__ var_first, const uint32_t* __restrict__ var_second, uint32_t n_vars,
                                                              uint32_t per, uint32_t n, const uint32_t* __restrict__ hist,
                                                              const uint32_t* __restrict__ overlap_begin, uint32_t* __restrict__ overlap_vars)
{
    extern __shared__ int cover[];   // [n] next free slot of every nucleotide's list
    uint32_t*      cursor = reinterpret_cast<uint32_t*>(cover);
    const uint32_t lane = threadIdx.x, b = blockIdx.x;
    for (uint32_t u = lane; u < n; u += 32) cursor[u] = overlap_begin[u] + hist[(uint64_t)b * n + u];
    __syncwarp();
    const uint32_t v0 = b * per, v1 = min(n_vars, v0 + per);
    // this lane's variable of a batch: first nucleotide and length (0 past the end of the slice); the next batch is
    // loaded while the current one is walked
    auto fetch = [&](uint32_t mine, uint32_t& f, uint32_t& k) {
        f = 0;
        k = 0;
        if (mine < v1) {
            f = var_first[mine];
            k = var_second[mine] - f + 1u;
            if (f >= n) k = 0;
            k = min(k, n - min(f, n));
        }
    };
    uint32_t nf, nk;
    fetch(v0 + lane, nf, nk);
    for (uint32_t base = v0; base < v1; base += 32) {
        const uint32_t f = nf, k = nk;
        fetch(base + 32 + lane, nf, nk);
        const uint32_t cnt = min(32u, v1 - base), kmax = __reduce_max_sync(0xFFFFFFFFu, k);
        if (kmax == 0) continue;
        if (kmax <= 32) {
            // Groups of 4, 8, 16 or 32 lanes (the smallest that holds the longest component of the batch) take one
            // variable each per step, lanes over its nucleotides. Lanes of different groups that cover the same
            // nucleotide find each other with a match; their lane order is the variable order, so the rank among
            // them is the place in the list, and the last of them moves the cursor.
            const uint32_t lg = kmax <= 4 ? 2u : kmax <= 8 ? 3u : kmax <= 16 ? 4u : 5u;
            const uint32_t g = lane >> lg, l = lane & ((1u << lg) - 1u), ng = 32u >> lg, lt = (1u << lane) - 1u;
            for (uint32_t s0 = 0; s0 < cnt; s0 += ng) {
                const uint32_t src = s0 + g;   // < 32; variables past the end of the slice have k = 0
                const uint32_t fj = __shfl_sync(0xFFFFFFFFu, f, src), kj = __shfl_sync(0xFFFFFFFFu, k, src);
                const bool     on = l < kj;
                const uint32_t u  = on ? fj + l : 0xFFFFFFFFu - lane;
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, u);
                if (on) {
                    const uint32_t at = cursor[u] + __popc(peers & lt);
                    overlap_vars[at]  = base + src;
                    if ((peers >> lane) == 1u) cursor[u] = at + 1;
                }
                __syncwarp();   // the next step may cover the same nucleotides from other lanes
            }
        } else {
            for (uint32_t j = 0; j < cnt; j++) {
                const uint32_t fj = __shfl_sync(0xFFFFFFFFu, f, j), kj = __shfl_sync(0xFFFFFFFFu, k, j);
                for (uint32_t t = lane; t < kj; t += 32) overlap_vars[cursor[fj + t]++] = base + j;
                __syncwarp();
            }
        }
    }
}

}   // namespace

// One warp per CTA, each walking its slice of the variables in order: at most ONE wave of CTAs (as many as are
// resident at once: 32 per SM, fewer when the per-nucleotide cursors of a long sequence take the shared memory),
// at least 128 variables each, and a histogram (n_blocks x n words) of at most 128 MB.
uint32_t index_overlap_blocks(uint64_t n_vars, uint32_t n, int sm_count)
{
    const uint64_t smem     = 4ull * (n + 1) + 1024;   // dynamic + reserved per CTA
    const uint64_t resident = std::max<uint64_t>(1, std::min<uint64_t>(32, (227ull << 10) / smem));
    uint64_t       want     = (n_vars + 127) / 128;
    want = std::min<uint64_t>(want, (uint64_t)sm_count * resident);
    want = std::min<uint64_t>(want, (32ull << 20) / std::max<uint32_t>(n, 1u));
    return (uint32_t)std::max<uint64_t>(want, 1);
}

cudaError_t launch_index_scan(const DevLibrary& lib, const uint32_t* counts, uint64_t* unit_site_begin, uint64_t* unit_var_begin,
                              const uint32_t* mask, uint32_t n, uint32_t* row_pairs, uint32_t* degree, uint32_t* row_first_pair,
                              uint32_t* pair_begin, uint64_t* totals, int sm_count, cudaStream_t s)
{
    const uint32_t Wm = (n + 31) / 32;
    cudaError_t    e  = cudaMemsetAsync(row_pairs, 0, 4 * (size_t)(n + 1), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(degree, 0, 4 * (size_t)(n + 1), s);
    if (e != cudaSuccess) return e;
    idx_unit_scan_kernel<<<1, kScanThreads, 0, s>>>(counts, lib.unit_rlen, lib.n_units, unit_site_begin, unit_var_begin, totals);
    if (n) {
        idx_row_pairs_kernel<<<min((n + 7) / 8, (uint32_t)sm_count * 8), 256, 0, s>>>(mask, n, Wm, row_pairs, degree);
        idx_col_pairs_kernel<<<dim3((n + 255) / 256, (n + 63) / 64), 256, 0, s>>>(mask, n, Wm, degree);
    }
    idx_pair_scan_kernel<<<1, kScanThreads, 0, s>>>(row_pairs, degree, n, row_first_pair, pair_begin, totals);
    return cudaGetLastError();
}

cudaError_t launch_index_vars(const DevLibrary& lib, const uint64_t* unit_site_begin, const uint64_t* unit_var_begin, const uint32_t* records,
                              uint32_t n_sites, uint32_t n_vars, const uint32_t* mask, uint32_t n, const uint32_t* pair_begin, int c3_rule,
                              uint32_t* pair_partner, uint32_t* var_begin, uint32_t* var_first, uint32_t* var_second, uint32_t* var_c3,
                              uint32_t* hist, uint32_t n_blocks, uint32_t* overlap_total, uint32_t* overlap_begin, uint64_t* totals,
                              int sm_count, cudaStream_t s)
{
    const uint32_t Wm = (n + 31) / 32;
    if (n) idx_adjacency_kernel<<<min((n + 7) / 8, (uint32_t)sm_count * 8), 256, 0, s>>>(mask, n, Wm, pair_begin, pair_partner);
    idx_vars_kernel<<<min(n_sites / 128 + 1, (uint32_t)sm_count * 32), 128, 0, s>>>(lib, unit_site_begin, unit_var_begin, records, n_sites, n_vars, mask, n, Wm,
                                                                                   pair_begin, c3_rule, var_begin, var_first, var_second, var_c3, totals);
    const uint32_t per  = (uint32_t)(((uint64_t)n_vars + n_blocks - 1) / n_blocks);
    const size_t   smem = 4 * (size_t)(n + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(idx_overlap_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    idx_overlap_hist_kernel<<<n_blocks, 32, smem, s>>>(var_first, var_second, n_vars, per, n, hist);
    if (n) idx_overlap_colscan_kernel<<<(n + 7) / 8, 256, 0, s>>>(hist, n_blocks, n, overlap_total);
    idx_overlap_scan_kernel<<<1, kScanThreads, 0, s>>>(overlap_total, n, overlap_begin, totals);
    return cudaGetLastError();
}

cudaError_t launch_index_fill(const uint32_t* var_first, const uint32_t* var_second, uint32_t n_vars, uint32_t n, const uint32_t* hist,
                              uint32_t n_blocks, const uint32_t* overlap_begin, uint32_t* overlap_vars, cudaStream_t s)
{
    const uint32_t per  = (uint32_t)(((uint64_t)n_vars + n_blocks - 1) / n_blocks);
    const size_t   smem = 4 * (size_t)(n + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(idx_overlap_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    idx_overlap_fill_kernel<<<n_blocks, 32, smem, s>>>(var_first, var_second, n_vars, per, n, hist, overlap_begin, overlap_vars);
    return cudaGetLastError();
}

}   // namespace bss
