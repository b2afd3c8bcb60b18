// All-gather-v of the shards' site records in ONE kernel over NVLink peer memory (row (e) of the path:
// the one exchange step of the library-sharded search).
//
// Every rank (one process per GPU) owns one allocation that its peers map through CUDA IPC:
//   control  arrived[2][world] u64 (epoch << 40 | words)
//   inbox    2 x world x capacity u32: slot [parity][r] receives rank r's records
//   out      2 x world x capacity u32: the canonical concatenation, double-buffered by epoch parity
// and a local block [header | records] that bss_search_query_into fills (the header's last u64 is the
// number of record words). The kernel, launched by every rank right behind its search, needs no host
// and ONE trip over the links:
//   1. the CTAs of peer p copy this rank's records into peer p's inbox[parity][rank] (16-byte peer-to-peer
//      stores over NVLink / NVSwitch); the last of them fences and stores (epoch, words) into peer p's
//      arrived[parity][rank]. The rank's own records go straight from the block to its output.
//   2. every CTA waits until all ranks' (epoch, words) have arrived in the LOCAL control block — the
//      counts and the completion of the data in one flag — takes the displacements (prefix sums of the
//      counts) and the CTAs of peer p move inbox[parity][p] to out[parity] + displacement[p] (local).
// When the kernel ends, out[parity] is the rank-order concatenation = canonical order.
// Waiting is bounded: a peer that never shows up raises the timeout flag instead of hanging the GPU.
// Reuse of a parity two epochs later is safe: a rank enters epoch e + 2 only after finishing e + 1, for
// which it needed every peer's flag of e + 1, which a peer sends only after its kernel of epoch e ended.
#include "biorseo_b200.h"
#include "bss_device.cuh"

#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr int      kCtasPerPeer = 4;
constexpr uint64_t kSpinLimit   = 1ull << 22;   // tight at first, then x ~0.4 us of back-off: about two seconds

__device__ __forceinline__ uint64_t load_sys(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void store_sys(uint64_t* p, uint64_t v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct ExchangeArgs {
    const uint32_t* block;          // local [header | records]
    uint32_t        header_words;
    uint32_t        world, rank;
    uint64_t        capacity;       // record words per rank
    uint64_t        epoch;
    uint64_t*       local_ctrl;     // arrived[2][world]
    uint32_t*       local_inbox;    // [2][world][capacity]
    uint32_t*       local_out;      // [2][world * capacity]
    uint64_t*       peer_ctrl[16];
    uint32_t*       peer_inbox[16];
    uint32_t*       arrivals;       // [world] local: CTAs of a peer that finished their copy
    uint64_t*       result;         // local: [0..world) counts, [world] total words, [world + 1] flags (1 overflow, 2 timeout)
};

// dst[i] = src[i] for i in [i0, i1), 16 bytes at a time when both sides are 16-byte aligned at i0 (i0 is a multiple of 4)
__device__ __forceinline__ void copy_words(uint32_t* dst, const uint32_t* src, uint64_t i0, uint64_t i1)
{
    if ((((uintptr_t)(dst + i0) | (uintptr_t)(src + i0)) & 15) == 0) {
        const uint64_t v1 = i0 + ((i1 - i0) & ~3ull);
        for (uint64_t i = i0 + 4ull * threadIdx.x; i < v1; i += 4ull * 256) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
        for (uint64_t i = v1 + threadIdx.x; i < i1; i += 256) dst[i] = src[i];
    } else {
        for (uint64_t i = i0 + threadIdx.x; i < i1; i += 256) dst[i] = src[i];
    }
}

__global__ void __launch_bounds__(256) exchange_kernel(ExchangeArgs a)
{
    const uint32_t p = blockIdx.x / kCtasPerPeer, c = blockIdx.x % kCtasPerPeer, par = (uint32_t)(a.epoch & 1);
    const uint64_t tag  = a.epoch & 0xFFFFFFull;
    const uint64_t mine = *reinterpret_cast<const uint64_t*>(a.block + a.header_words - 2);
    const bool     fits = mine <= a.capacity;
    // 1. this rank's records into peer p's inbox (nothing when they do not fit: the count alone tells the peers)
    if (p != a.rank) {
        if (fits) {
            const uint64_t per = ((mine + kCtasPerPeer - 1) / kCtasPerPeer + 3) & ~3ull;
            const uint64_t i0 = min(mine, c * per), i1 = min(mine, i0 + per);
            copy_words(a.peer_inbox[p] + ((uint64_t)par * a.world + a.rank) * a.capacity, a.block + a.header_words, i0, i1);
        }
        __syncthreads();   // the CTA's stores happen before thread 0's fence
        if (threadIdx.x == 0) {
            __threadfence_system();   // release: this CTA's peer stores, before its arrival is counted
            if (atomicAdd(a.arrivals + p, 1u) == kCtasPerPeer - 1) {   // the last CTA of peer p publishes (epoch, words)
                a.arrivals[p] = 0;
                __threadfence_system();   // acquire side of the arrival counter, and release of the flag
                store_sys(a.peer_ctrl[p] + par * a.world + a.rank, (tag << 40) | (mine & 0xFFFFFFFFFFull));
            }
        }
    } else if (threadIdx.x == 0 && c == 0) {
        store_sys(a.local_ctrl + par * a.world + a.rank, (tag << 40) | (mine & 0xFFFFFFFFFFull));
    }
    // 2. all counts (= all data) have arrived: displacements, then inbox -> output
    __shared__ uint64_t s_disp, s_count;
    __shared__ uint32_t s_flags;
    if (threadIdx.x == 0) {
        uint64_t disp = 0, total = 0, count = 0;
        uint32_t flags = 0;
        for (uint32_t r = 0; r < a.world; r++) {
            const uint64_t* slot = a.local_ctrl + par * a.world + r;
            uint64_t        v    = load_sys(slot), spins = 0;
            while ((v >> 40) != tag && spins < kSpinLimit) {
                if (spins > 4096) __nanosleep(200);
                v = load_sys(slot);
                spins++;
            }
            if ((v >> 40) != tag) flags |= 2u;
            const uint64_t n = v & 0xFFFFFFFFFFull;
            if (n > a.capacity) flags |= 1u;
            if (r < p) disp += n;
            if (r == p) count = n;
            total += n;
            if (blockIdx.x == 0) a.result[r] = n;
        }
        s_disp  = disp;
        s_count = count;
        s_flags = flags;
        if (blockIdx.x == 0) {
            a.result[a.world]     = total;
            a.result[a.world + 1] = flags;
        }
    }
    __syncthreads();
    if (s_flags == 0) {
        const uint64_t  n   = s_count;
        const uint64_t  per = ((n + kCtasPerPeer - 1) / kCtasPerPeer + 3) & ~3ull;
        const uint64_t  i0 = min(n, c * per), i1 = min(n, i0 + per);
        const uint32_t* src = p == a.rank ? a.block + a.header_words : a.local_inbox + ((uint64_t)par * a.world + p) * a.capacity;
        uint32_t*       dst = a.local_out + (uint64_t)par * a.world * a.capacity + s_disp;
        if (p == a.rank) {
            copy_words(dst, src, i0, i1);
        } else {
            // peer-written memory: read it past L1
            for (uint64_t i = i0 + threadIdx.x; i < i1; i += 256) dst[i] = __ldcg(src + i);
        }
    }
}

}   // namespace

struct bss_exchange {
    int          device = 0;
    uint32_t     world = 0, rank = 0, header_words = 0;
    uint64_t     capacity = 0, epoch = 0;
    void*        shared = nullptr;   // control + inbox, mapped by the peers
    size_t       shared_bytes = 0, ctrl_bytes = 0;
    uint32_t*    outbuf = nullptr;   // local [2][world * capacity]
    uint32_t*    block = nullptr;    // local [header | records]
    uint32_t*    arrivals = nullptr;
    uint64_t*    result = nullptr;
    bool         opened = false;
    void*        peer_base[16] = {};
};

extern "C" {

int bss_exchange_create(int device, uint32_t world, uint32_t rank, uint64_t capacity_words, uint32_t header_words, bss_exchange** out)
{
    return bss::guarded([&]() -> int {
    if (!out) return bss::set_error(BSS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (world < 1 || world > 16 || rank >= world || header_words < 2 || (header_words & 1)) return bss::set_error(BSS_ERR_INVALID, "bad exchange geometry");
    if (capacity_words >= (1ull << 40)) return bss::set_error(BSS_ERR_LIMIT, "exchange capacity above 2^40 words");
    cudaError_t e = device >= 0 ? cudaSetDevice(device) : cudaGetDevice(&device);
    if (e != cudaSuccess) return bss::set_error(BSS_ERR_CUDA, cudaGetErrorString(e));
    bss_exchange* x = new bss_exchange();
    x->device = device; x->world = world; x->rank = rank; x->header_words = header_words;
    x->capacity     = (capacity_words + 3) & ~3ull;
    x->ctrl_bytes   = (2 * (size_t)world * 8 + 255) & ~(size_t)255;
    x->shared_bytes = x->ctrl_bytes + 2 * (size_t)world * x->capacity * 4;
    e = cudaMalloc(&x->shared, x->shared_bytes);
    if (e == cudaSuccess) e = cudaMemset(x->shared, 0, x->ctrl_bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->outbuf), 2 * (size_t)world * x->capacity * 4);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->block), ((size_t)header_words + x->capacity) * 4);
    if (e == cudaSuccess) e = cudaMemset(x->block, 0, (size_t)header_words * 4);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->arrivals), 4 * (size_t)world);
    if (e == cudaSuccess) e = cudaMemset(x->arrivals, 0, 4 * (size_t)world);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->result), 8 * (size_t)(world + 2));
    if (e == cudaSuccess) e = cudaMemset(x->result, 0, 8 * (size_t)(world + 2));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        bss_exchange_destroy(x);
        return bss::set_error(BSS_ERR_CUDA, std::string("bss_exchange_create: ") + cudaGetErrorString(e));
    }
    *out = x;
    return BSS_OK;
    });
}

int bss_exchange_handle(const bss_exchange* x, void* handle64)
{
    return bss::guarded([&]() -> int {
    if (!x || !handle64) return bss::set_error(BSS_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t        e = cudaSetDevice(x->device);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, x->shared);
    if (e != cudaSuccess) return bss::set_error(BSS_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    std::memcpy(handle64, &h, 64);
    return BSS_OK;
    });
}

int bss_exchange_open(bss_exchange* x, const void* handles)
{
    return bss::guarded([&]() -> int {
    if (!x || !handles) return bss::set_error(BSS_ERR_INVALID, "null argument");
    if (x->opened) return bss::set_error(BSS_ERR_INVALID, "exchange already opened");
    cudaError_t e = cudaSetDevice(x->device);
    for (uint32_t r = 0; e == cudaSuccess && r < x->world; r++) {
        if (r == x->rank) {
            x->peer_base[r] = x->shared;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + 64 * (size_t)r, 64);
        e = cudaIpcOpenMemHandle(&x->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
    }
    if (e != cudaSuccess) {
        for (uint32_t r = 0; r < x->world; r++)
            if (r != x->rank && x->peer_base[r]) { cudaIpcCloseMemHandle(x->peer_base[r]); x->peer_base[r] = nullptr; }
        cudaGetLastError();
        return bss::set_error(BSS_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    x->opened = true;
    return BSS_OK;
    });
}

uint32_t* bss_exchange_block(const bss_exchange* x) { return x ? x->block : nullptr; }
uint64_t  bss_exchange_capacity(const bss_exchange* x) { return x ? x->capacity : 0; }

int bss_exchange_run_async(bss_exchange* x, void* cuda_stream, const uint32_t** d_out, const uint64_t** d_result)
{
    return bss::guarded([&]() -> int {
    if (!x) return bss::set_error(BSS_ERR_INVALID, "null argument");
    if (!x->opened) return bss::set_error(BSS_ERR_INVALID, "exchange not opened");
    cudaError_t e = cudaSetDevice(x->device);
    if (e != cudaSuccess) return bss::set_error(BSS_ERR_CUDA, cudaGetErrorString(e));
    x->epoch++;
    ExchangeArgs a{};
    a.block = x->block; a.header_words = x->header_words; a.world = x->world; a.rank = x->rank; a.capacity = x->capacity; a.epoch = x->epoch;
    a.local_ctrl  = static_cast<uint64_t*>(x->shared);
    a.local_inbox = reinterpret_cast<uint32_t*>(static_cast<char*>(x->shared) + x->ctrl_bytes);
    a.local_out   = x->outbuf;
    for (uint32_t r = 0; r < x->world; r++) {
        a.peer_ctrl[r]  = static_cast<uint64_t*>(x->peer_base[r]);
        a.peer_inbox[r] = reinterpret_cast<uint32_t*>(static_cast<char*>(x->peer_base[r]) + x->ctrl_bytes);
    }
    a.arrivals = x->arrivals;
    a.result   = x->result;
    exchange_kernel<<<x->world * kCtasPerPeer, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return bss::set_error(BSS_ERR_CUDA, std::string("exchange launch: ") + cudaGetErrorString(e));
    if (d_out) *d_out = a.local_out + (x->epoch & 1) * (uint64_t)x->world * x->capacity;
    if (d_result) *d_result = x->result;
    return BSS_OK;
    });
}

void bss_exchange_destroy(bss_exchange* x)
{
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (uint32_t r = 0; r < x->world; r++)
        if (x->opened && r != x->rank && x->peer_base[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
    if (x->shared) cudaFree(x->shared);
    if (x->outbuf) cudaFree(x->outbuf);
    if (x->block) cudaFree(x->block);
    if (x->arrivals) cudaFree(x->arrivals);
    if (x->result) cudaFree(x->result);
    delete x;
}

}   // extern "C"
