// All-gather-v of the shards' site records in ONE kernel over NVLink peer memory (row (e) of the path:
// the one exchange step of the library-sharded search). No NCCL call, no host round trip.
//
// Every rank owns one allocation that its peers map (CUDA IPC between processes, plain peer access inside
// one process):
//   control  count[2][world] u64 (epoch << 40 | words), done[2][world] u64 (epoch)
//   out      2 x (world x capacity) u32: the canonical concatenation, double-buffered by epoch parity
// and a local block [header | records] that bss_search_query_into fills (the header's last u64 is the number
// of record words: always written, also when the records did not fit). The kernel, launched by every rank
// right behind its search:
//   A. counts first: CTA 0 stores (epoch, words) into every peer's count[parity][rank] — 8 bytes per peer;
//   B. every CTA waits until all ranks' counts of this epoch are in the LOCAL control block and takes the
//      displacements (prefix sums of the counts): the same numbers on every rank;
//   C. direct placement: the rank's records go straight into every peer's OUTPUT at the rank's displacement
//      (its own output included) — peer-to-peer stores over NVLink / NVSwitch, 16 bytes wide where the
//      displacement allows, tiles of 16 K words interleaved over the peers so that every link is busy at once.
//      No inbox, no second pass over the data: 2 x (world x capacity) words of memory in all;
//   D. the last CTA to finish fences and stores `done` into every peer's control block, then waits for the
//      peers' own `done` flags: when the kernel ends, out[parity] holds every rank's records.
// Waiting is bounded: a peer that never shows up raises the timeout flag instead of hanging the GPU.
// A parity is reused two epochs later: a rank writes into a peer's out[parity] of epoch e + 2 only after it has
// that peer's count of e + 2, which the peer sends from its kernel of epoch e + 2 — stream-ordered behind
// whatever read its out[parity] of epoch e.
#include "biorseo_b200.h"
#include "bss_device.cuh"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr uint32_t kTileWords = 16384;          // words per copy tile
constexpr uint64_t kSpinLimit = 1ull << 22;     // tight at first, then x ~0.4 us of back-off: about two seconds
constexpr int      kThreadsX  = 256;

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
    uint64_t*       local_ctrl;     // count[2][world], done[2][world]
    uint64_t*       peer_ctrl[16];
    uint32_t*       peer_out[16];   // out[2][world * capacity] of every rank (own included)
    uint32_t*       arrivals;       // local: CTAs that finished their tiles
    uint64_t*       result;         // local: [0..world) counts, [world] total words, [world + 1] flags (1 overflow, 2 timeout)
};

// Bounded wait until (*slot >> shift) == tag. `value` receives what was read last.
__device__ __forceinline__ bool wait_tag(const uint64_t* slot, uint64_t tag, int shift, uint64_t& value)
{
    uint64_t v = load_sys(slot), spins = 0;
    while ((v >> shift) != tag && spins < kSpinLimit) {
        if (spins > 4096) __nanosleep(200);
        v = load_sys(slot);
        spins++;
    }
    value = v;
    return (v >> shift) == tag;
}

// dst[0..n) = src[0..n): 16-byte stores where dst allows (the peer side), 16-byte loads when src allows too
__device__ __forceinline__ void copy_tile(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t n)
{
    const uint32_t head = min(n, (uint32_t)(((16u - ((uintptr_t)dst & 15u)) & 15u) / 4u));   // words until dst is 16-byte aligned
    for (uint32_t i = threadIdx.x; i < head; i += kThreadsX) dst[i] = src[i];
    const uint32_t  body = (n - head) & ~3u;
    uint4*          d4   = reinterpret_cast<uint4*>(dst + head);
    const uint32_t* s    = src + head;
    if (((uintptr_t)s & 15u) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(s);
        for (uint32_t i = threadIdx.x; i < body / 4; i += kThreadsX) d4[i] = s4[i];
    } else {
        for (uint32_t i = threadIdx.x; i < body / 4; i += kThreadsX) d4[i] = make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
    }
    for (uint32_t i = head + body + threadIdx.x; i < n; i += kThreadsX) dst[i] = src[i];
}

__global__ void __launch_bounds__(kThreadsX) exchange_kernel(ExchangeArgs a)
{
    const uint32_t par  = (uint32_t)(a.epoch & 1);
    const uint64_t tag  = a.epoch & 0xFFFFFFull;
    const uint64_t mine = *reinterpret_cast<const uint64_t*>(a.block + a.header_words - 2);
    // A. counts first
    if (blockIdx.x == 0 && threadIdx.x < a.world) store_sys(a.peer_ctrl[threadIdx.x] + par * a.world + a.rank, (tag << 40) | (mine & 0xFFFFFFFFFFull));
    // B. all counts of this epoch: displacement of this rank's records, total, flags
    __shared__ uint64_t s_disp;
    __shared__ uint32_t s_flags;
    if (threadIdx.x == 0) {
        uint64_t disp = 0, total = 0;
        uint32_t flags = 0;
        for (uint32_t r = 0; r < a.world; r++) {
            uint64_t v;
            if (!wait_tag(a.local_ctrl + par * a.world + r, tag, 40, v)) flags |= 2u;
            const uint64_t n = v & 0xFFFFFFFFFFull;
            if (n > a.capacity) flags |= 1u;
            if (r < a.rank) disp += n;
            total += n;
            if (blockIdx.x == 0) a.result[r] = n;
        }
        s_disp  = disp;
        s_flags = flags;
        if (blockIdx.x == 0) {
            a.result[a.world]     = total;
            a.result[a.world + 1] = flags;
        }
    }
    __syncthreads();
    // C. direct placement into every rank's output (nothing when a stream does not fit or a peer is missing)
    if (s_flags == 0) {
        const uint64_t  tiles_per_peer = (mine + kTileWords - 1) / kTileWords;
        const uint32_t* src = a.block + a.header_words;
        for (uint64_t t = blockIdx.x; t < tiles_per_peer * a.world; t += gridDim.x) {
            const uint32_t p  = (uint32_t)((t + a.rank) % a.world);   // (every rank starts on another peer)
            const uint64_t i0 = (t / a.world) * kTileWords;
            const uint32_t n  = (uint32_t)min((uint64_t)kTileWords, mine - i0);
            copy_tile(a.peer_out[p] + (uint64_t)par * a.world * a.capacity + s_disp + i0, src + i0, n);
        }
    }
    // D. completion
    __syncthreads();   // the CTA's stores happen before thread 0's fence
    if (threadIdx.x == 0) {
        __threadfence_system();   // release: this CTA's peer stores, before its arrival is counted
        if (atomicAdd(a.arrivals, 1u) == gridDim.x - 1) {   // the last CTA of this rank
            *a.arrivals = 0;
            __threadfence_system();
            for (uint32_t p = 0; p < a.world; p++) store_sys(a.peer_ctrl[p] + 2 * a.world + par * a.world + a.rank, tag);
            uint32_t flags = 0;
            for (uint32_t r = 0; r < a.world; r++) {
                uint64_t v;
                if (!wait_tag(a.local_ctrl + 2 * a.world + par * a.world + r, tag, 0, v)) flags |= 2u;
            }
            if (flags) a.result[a.world + 1] |= flags;
        }
    }
}

}   // namespace

struct bss_exchange {
    int          device = 0;
    uint32_t     world = 0, rank = 0, header_words = 0;
    uint64_t     capacity = 0, epoch = 0;
    void*        shared = nullptr;   // control + output, mapped by the peers
    size_t       shared_bytes = 0, ctrl_bytes = 0;
    uint32_t*    block = nullptr;    // local [header | records]
    uint32_t*    arrivals = nullptr;
    uint64_t*    result = nullptr;
    int          sm_count = 148;
    bool         opened = false, local = false;
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
    x->ctrl_bytes   = (4 * (size_t)world * 8 + 255) & ~(size_t)255;
    x->shared_bytes = x->ctrl_bytes + 2 * (size_t)world * x->capacity * 4;
    e = cudaDeviceGetAttribute(&x->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc(&x->shared, x->shared_bytes);
    if (e == cudaSuccess) e = cudaMemset(x->shared, 0, x->ctrl_bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->block), ((size_t)header_words + x->capacity) * 4);
    if (e == cudaSuccess) e = cudaMemset(x->block, 0, (size_t)header_words * 4);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&x->arrivals), 4);
    if (e == cudaSuccess) e = cudaMemset(x->arrivals, 0, 4);
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

int bss_exchange_open_local(bss_exchange* const* all, uint32_t world)
{
    return bss::guarded([&]() -> int {
    if (!all || world < 1 || world > 16) return bss::set_error(BSS_ERR_INVALID, "bad argument");
    for (uint32_t r = 0; r < world; r++)
        if (!all[r] || all[r]->world != world || all[r]->rank != r || all[r]->opened || all[r]->capacity != all[0]->capacity)
            return bss::set_error(BSS_ERR_INVALID, "the exchanges of one process must share one geometry, in rank order, unopened");
    for (uint32_t r = 0; r < world; r++) {
        cudaError_t e = cudaSetDevice(all[r]->device);
        for (uint32_t p = 0; e == cudaSuccess && p < world; p++) {
            if (all[p]->device == all[r]->device) continue;
            int can = 0;
            e = cudaDeviceCanAccessPeer(&can, all[r]->device, all[p]->device);
            if (e == cudaSuccess && !can) return bss::set_error(BSS_ERR_CUDA, "the GPUs cannot map each other's memory");
            if (e == cudaSuccess) {
                e = cudaDeviceEnablePeerAccess(all[p]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { e = cudaSuccess; cudaGetLastError(); }
            }
        }
        if (e != cudaSuccess) return bss::set_error(BSS_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    }
    for (uint32_t r = 0; r < world; r++) {
        for (uint32_t p = 0; p < world; p++) all[r]->peer_base[p] = all[p]->shared;
        all[r]->opened = true;
        all[r]->local  = true;
    }
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
    a.local_ctrl = static_cast<uint64_t*>(x->shared);
    for (uint32_t r = 0; r < x->world; r++) {
        a.peer_ctrl[r] = static_cast<uint64_t*>(x->peer_base[r]);
        a.peer_out[r]  = reinterpret_cast<uint32_t*>(static_cast<char*>(x->peer_base[r]) + x->ctrl_bytes);
    }
    a.arrivals = x->arrivals;
    a.result   = x->result;
    // CTAs: enough tiles in flight for the capacity, at most two per SM; small streams need one per peer
    const uint64_t tiles = ((x->capacity + kTileWords - 1) / kTileWords) * x->world;
    const uint32_t grid  = (uint32_t)std::max<uint64_t>(x->world, std::min<uint64_t>(tiles, 2ull * (uint64_t)x->sm_count));
    exchange_kernel<<<grid, kThreadsX, 0, static_cast<cudaStream_t>(cuda_stream)>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return bss::set_error(BSS_ERR_CUDA, std::string("exchange launch: ") + cudaGetErrorString(e));
    if (d_out) *d_out = a.peer_out[x->rank] + (x->epoch & 1) * (uint64_t)x->world * x->capacity;
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
        if (x->opened && !x->local && r != x->rank && x->peer_base[r]) cudaIpcCloseMemHandle(x->peer_base[r]);
    if (x->shared) cudaFree(x->shared);
    if (x->block) cudaFree(x->block);
    if (x->arrivals) cudaFree(x->arrivals);
    if (x->result) cudaFree(x->result);
    delete x;
}

}   // extern "C"
