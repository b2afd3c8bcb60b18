// Integer-issue probe: the denominator of the second roofline of the search. The match and
// walk phases are funnel shifts and three-input logic on 32-bit words; this kernel issues nothing else
// (8 independent SHF + LOP3 chains per thread, all SMs, full occupancy) and reports the rate the GPU
// sustains at the clocks it runs at: thread-level integer operations per second.
#include "bss_device.cuh"

namespace bss {
namespace {

__global__ void __launch_bounds__(256) int_issue_kernel(uint32_t seed, uint32_t m, uint32_t iters, uint32_t* __restrict__ sink)
{
    uint32_t x[8];
    const uint32_t t = blockIdx.x * 256 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = seed * (t + 1) + 0x9E3779B9u * (i + 1);
    const uint32_t s = (seed & 15u) + 1u;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = (__funnelshift_r(x[i], x[(i + 1) & 7], s) & m) ^ x[i];   // one SHF + one LOP3
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i];
    if (acc == 0x12345678u) sink[0] = acc;   // keeps the chains alive
}

}   // namespace

// Thread-level operations issued by one launch: threads * iters * 4 rounds * 8 chains * 2 instructions.
cudaError_t launch_int_issue_probe(uint32_t iters, int sm_count, uint32_t* sink, uint64_t* thread_ops, cudaStream_t s)
{
    const uint32_t grid = (uint32_t)sm_count * 8;
    int_issue_kernel<<<grid, 256, 0, s>>>(0x2545F491u, 0x7F3DB5E9u, iters, sink);
    *thread_ops = (uint64_t)grid * 256 * iters * 4 * 8 * 2;
    return cudaGetLastError();
}

}   // namespace bss
