// Pair-mask producer: the packed allowed_basepair predicate from the base-pair probability
// matrix, on the device.
//
// What it replaces (reference, CPU): the y_{u,v} loop of the MOIP constructor,
// cppsrc/MOIP.cpp:77-90 (index_of_yuv_ from `pij(u,v) > theta`, u < n-6, v >= u+4), read back
// by MOIP::allowed_basepair (cppsrc/MOIP.cpp:896-910) for every link of every placement. Here
// the predicate becomes the bit matrix the search kernels consume (bit v of row u, u < v),
// built once per query from the float matrix: n^2 floats in, n^2/8 bytes out — a pure
// streaming kernel, bound by HBM.
#include "bss_device.cuh"

namespace bss {
namespace {

// Row-major matrix (col_stride == 1): one warp per row. The words left of the diagonal are stored
// as zeroes without reading anything; from the diagonal on, the warp takes 8 consecutive words per
// step: 8 independent, fully coalesced 128-byte loads in flight per lane, then 8 ballots; lanes
// 0..7 store one word each (one 32-byte sector).
__global__ void __launch_bounds__(256) pair_mask_rows_kernel(const float* __restrict__ pij, uint32_t n, uint64_t row_stride, float theta,
                                                             uint32_t* __restrict__ mask)
{
    const uint32_t Wm = (n + 31) >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t u = blockIdx.x * 8 + warp; u < n; u += gridDim.x * 8) {
        uint32_t*      out    = mask + (uint64_t)u * Wm;
        const bool     row_on = u + 6 < n;
        const uint32_t w_on   = row_on ? ((u + 4) >> 5) & ~7u : Wm;   // first group of 8 words that can hold a set bit
        for (uint32_t w = lane; w < min(w_on, Wm); w += 32) out[w] = 0u;
        const float* row = pij + (uint64_t)u * row_stride;
        for (uint32_t w0 = w_on; w0 < Wm; w0 += 8) {
            float p[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t v = ((w0 + k) << 5) + lane;
                p[k] = (v < n && v >= u + 4) ? __ldg(row + v) : theta;   // theta itself is "not allowed"
            }
            uint32_t mine = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t bits = __ballot_sync(0xFFFFFFFFu, p[k] > theta);
                if ((int)lane == k) mine = bits;
            }
            if (lane < 8 && w0 + lane < Wm) out[w0 + lane] = mine;
        }
    }
}

// Column-major matrix (row_stride == 1, Eigen's default): one warp per 32 x 32 tile, lane = row;
// every load reads 32 consecutive rows of one column (one line), each lane assembles the word of
// its own row.
__global__ void __launch_bounds__(256) pair_mask_cols_kernel(const float* __restrict__ pij, uint32_t n, uint64_t col_stride, float theta,
                                                             uint32_t* __restrict__ mask)
{
    const uint32_t Wm = (n + 31) >> 5, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tiles = Wm * Wm;
    for (uint32_t t = blockIdx.x * 8 + warp; t < tiles; t += gridDim.x * 8) {
        const uint32_t tu = t / Wm, w = t - tu * Wm, u = (tu << 5) + lane;
        uint32_t       bits = 0;
        if (((w + 1) << 5) > (tu << 5) + 4) {   // the tile reaches the allowed side of the diagonal: 32 loads in flight per lane
            float p[32];
#pragma unroll
            for (uint32_t b = 0; b < 32; b++) {
                const uint32_t v = (w << 5) + b;
                p[b] = (u + 6 < n && v < n && v >= u + 4) ? __ldg(pij + (uint64_t)v * col_stride + u) : theta;
            }
#pragma unroll
            for (uint32_t b = 0; b < 32; b++) bits |= (p[b] > theta ? 1u : 0u) << b;
        }
        if (u < n) mask[u * Wm + w] = bits;
    }
}

// Any other strides: one thread per word, 32 strided loads.
__global__ void __launch_bounds__(256) pair_mask_any_kernel(const float* __restrict__ pij, uint32_t n, uint64_t row_stride, uint64_t col_stride,
                                                            float theta, uint32_t* __restrict__ mask)
{
    const uint32_t Wm = (n + 31) >> 5, words = n * Wm;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < words; i += gridDim.x * 256) {
        const uint32_t u = i / Wm, w = i - u * Wm;
        uint32_t       bits = 0;
        if (u + 6 < n)
            for (uint32_t b = 0; b < 32; b++) {
                const uint32_t v = (w << 5) + b;
                if (v < n && v >= u + 4 && __ldg(pij + (uint64_t)u * row_stride + (uint64_t)v * col_stride) > theta) bits |= 1u << b;
            }
        mask[i] = bits;
    }
}

}   // namespace

cudaError_t launch_pair_mask(const float* pij, uint32_t n, uint64_t row_stride, uint64_t col_stride, float theta, uint32_t* mask,
                             int sm_count, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const uint32_t Wm = (n + 31) / 32;
    const uint64_t warps = col_stride == 1 ? (uint64_t)n : row_stride == 1 ? (uint64_t)Wm * Wm : ((uint64_t)n * Wm + 31) / 32;
    uint64_t       grid  = (warps + 7) / 8;
    const uint64_t cap   = (uint64_t)sm_count * 32;   // grid-stride beyond a few waves
    if (grid > cap) grid = cap;
    if (col_stride == 1)
        pair_mask_rows_kernel<<<(uint32_t)grid, 256, 0, s>>>(pij, n, row_stride, theta, mask);
    else if (row_stride == 1)
        pair_mask_cols_kernel<<<(uint32_t)grid, 256, 0, s>>>(pij, n, col_stride, theta, mask);
    else
        pair_mask_any_kernel<<<(uint32_t)grid, 256, 0, s>>>(pij, n, row_stride, col_stride, theta, mask);
    return cudaGetLastError();
}

// Cut of a padded all-gather (biorseo_b200/sharded.py: every rank's block is [header | records |
// padding], the header's last u64 is the number of record words): writes the blocks' records back
// to back in rank order = canonical (module, placement) order, and the per-rank counts, without the
// host having to know any size. One CTA per rank. A count that does not fit a block raises *overflow.
__global__ void __launch_bounds__(256) gather_cut_kernel(const uint32_t* __restrict__ blocks, uint32_t world, uint64_t block_words,
                                                         uint32_t header_words, uint32_t* __restrict__ out, uint64_t* __restrict__ counts,
                                                         uint32_t* __restrict__ overflow)
{
    const uint32_t r = blockIdx.x;
    __shared__ uint64_t before, mine;
    if (threadIdx.x == 0) {
        uint64_t sum = 0, c = 0;
        bool     bad = false;
        for (uint32_t p = 0; p < world; p++) {
            const uint64_t n = *reinterpret_cast<const uint64_t*>(blocks + p * block_words + header_words - 2);
            if (n > block_words - header_words) bad = true;
            if (p < r) sum += n;
            if (p == r) c = n;
        }
        before = sum;
        mine   = bad ? 0 : c;
        if (bad) *overflow = 1u;
        counts[r] = c;
    }
    __syncthreads();
    const uint32_t* src = blocks + r * block_words + header_words;
    uint32_t*       dst = out + before;
    for (uint64_t i = threadIdx.x; i < mine; i += 256) dst[i] = src[i];
}

cudaError_t launch_gather_cut(const uint32_t* blocks, uint32_t world, uint64_t block_words, uint32_t header_words, uint32_t* out,
                              uint64_t* counts, uint32_t* overflow, cudaStream_t s)
{
    if (world == 0) return cudaSuccess;
    gather_cut_kernel<<<world, 256, 0, s>>>(blocks, world, block_words, header_words, out, counts, overflow);
    return cudaGetLastError();
}

}   // namespace bss
