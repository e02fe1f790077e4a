// The one exchange step of the path: padded detections, their counts and the matched-target
// statistics of the local images go into ONE [capacity, row_words] fp32-word buffer per rank
// (row_words = T*6 + 5 rounded up to a multiple of four: rows are whole 16-byte vectors).
//
//  * pack_shard_kernel      packs the local buffer (one launch instead of the ~10 element-wise torch kernels the
//    same packing costs with tensor ops); NCCL's all-gather then moves it (sharding.all_gather_packed).
//  * pack_exchange_kernel   packs AND exchanges in the same launch: every CTA reads its image's row ONCE (16-byte
//    loads) and stores it straight into slot `rank` of EVERY rank's gathered buffer through NVLink peer memory
//    (16-byte stores; the buffers are CUDA-IPC mappings of one arena per rank, sharding.PeerExchange) -- no NCCL
//    call, no host involvement, capturable into the step graph.
//
//    Protocol per ring slot k (one slot per captured step graph; its launches are serialised on one stream):
//        exchange_open_kernel(k)      FIRST node of the slot's graph: m = ++opened[k] (local) and ack[k][rank] = m on
//                                     every peer: "this rank has finished reading launch m - 1 of the slot" (the
//                                     contract of a slot: its gathered buffer stays valid until the slot's graph is
//                                     launched again).
//        pack_exchange_kernel(k)      LAST node: every CTA waits until ack[k][r] >= m for every rank r (bounded spin on
//                                     LOCAL memory: nobody still reads what is about to be overwritten; a peer that
//                                     lags spins the writer, a dead peer becomes an error word after 10 s, never a
//                                     hang), writes its row to every rank, and then `world` of its threads publish
//                                     rowflag[k][rank][row] = m on one rank each with a system-scope RELEASE store
//                                     (cumulative over the CTA's stores through the CTA barrier in front of it).
//                                     No ticket, no last-CTA tail, no second fence.
//        exchange_wait_kernel(k)      spins until rowflag[k][r][row] >= opened[k] for every rank and row: slot complete.
#include "common.cuh"

namespace ssd {

__host__ __device__ inline int shard_row_words(int max_total) { return (max_total * 6 + 5 + 3) & ~3; }

// {positives, hard negatives selected, ignored, detections} of image b
__device__ __forceinline__ int4 row_stats(const int32_t* __restrict__ counts, const int32_t* __restrict__ assign_stats,
                                          const int32_t* __restrict__ mining_stats, int b) {
    return make_int4(assign_stats ? assign_stats[b * 4 + 0] : 0, mining_stats ? mining_stats[b * 4 + 2] : 0,
                     assign_stats ? assign_stats[b * 4 + 1] : 0, counts[b]);
}

// word e of the packed row of image b (b >= batch: padding row, count = -1, everything else 0)
__device__ __forceinline__ float row_word(const float* __restrict__ src, int4 st, bool padding, int max_total, int e) {
    const int t6 = max_total * 6;
    if (padding) return e == t6 ? __int_as_float(-1) : 0.f;
    if (e < t6) return src[e];
    const int k = e - t6;
    return __int_as_float(k == 0 ? st.w : k == 1 ? st.x : k == 2 ? st.y : k == 3 ? st.z : k == 4 ? st.w : 0);
}

// one CTA per row of the buffer; rows >= batch are padding
__global__ void __launch_bounds__(256)
pack_shard_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts,
                  const int32_t* __restrict__ assign_stats, const int32_t* __restrict__ mining_stats, int batch,
                  int max_total, float* __restrict__ shard, int32_t* __restrict__ stats_out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    const int b = blockIdx.x;
    const int words = shard_row_words(max_total);
    const bool padding = b >= batch;
    const int4 st = padding ? make_int4(0, 0, 0, 0) : row_stats(counts, assign_stats, mining_stats, b);
    const float* src = dets + (size_t)b * max_total * 6;
    float* row = shard + (size_t)b * words;
    for (int e = threadIdx.x; e < words; e += blockDim.x) row[e] = row_word(src, st, padding, max_total, e);
    if (!padding && stats_out && threadIdx.x == 0) reinterpret_cast<int4*>(stats_out)[b] = st;
}

// ---- peer-memory exchange ----
// Arena of one rank (sharding.PeerExchange): [header | slots x world x capacity x row_words floats]
//   header (int64 words): [0] error, [8 + k] opened[k], [8 + S + k * W + r] ack[k][r],
//                         [8 + S + S * W + (k * W + r) * kMaxRows + row] rowflag[k][r][row]
constexpr int kMaxWorld = SSD_EXCHANGE_MAX_WORLD;
constexpr int kMaxSlots = SSD_EXCHANGE_MAX_SLOTS;
constexpr int kMaxRows = SSD_EXCHANGE_MAX_ROWS;                 // images per rank and slot
constexpr unsigned long long kSpinLimitNs = 10000000000ull;     // 10 s: a dead peer becomes an error, not a hang

__host__ __device__ inline size_t hdr_opened(int k) { return 8 + (size_t)k; }
__host__ __device__ inline size_t hdr_ack(int k, int r) { return 8 + kMaxSlots + (size_t)k * kMaxWorld + r; }
__host__ __device__ inline size_t hdr_rowflag(int k, int r, int row) {
    return 8 + kMaxSlots + (size_t)kMaxSlots * kMaxWorld + ((size_t)k * kMaxWorld + r) * kMaxRows + row;
}
constexpr size_t kHeaderWords = 8 + kMaxSlots + (size_t)kMaxSlots * kMaxWorld + (size_t)kMaxSlots * kMaxWorld * kMaxRows;
__host__ __device__ inline size_t header_bytes() { return round_up(kHeaderWords * 8, 256); }

struct PeerArenas {
    unsigned char* base[kMaxWorld];       // the arena of rank r as mapped into THIS process
};

__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// spin until *flag (LOCAL memory, written by a peer) has reached `need`; false on time-out
__device__ __forceinline__ bool wait_word(long long* hdr, const long long* flag, long long need) {
    if (ld_acquire_sys(flag) >= need) return true;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < need) {
        if (global_ns() - t0 > kSpinLimitNs) { hdr[0] = 1; return false; }
        __nanosleep(64);
    }
    return true;
}

// first node of a slot's step graph: a new launch of the slot begins, its previous contents are released
__global__ void exchange_open_kernel(PeerArenas peers, int world, int rank, int slot) {
    griddep_wait();
    griddep_launch_dependents();
    long long* hdr = reinterpret_cast<long long*>(peers.base[rank]);
    const long long m = hdr[hdr_opened(slot)] + 1;        // only this kernel writes it; launches of a slot are serialised
    __syncthreads();
    if (threadIdx.x == 0) hdr[hdr_opened(slot)] = m;
    if (threadIdx.x < world)
        st_release_sys(reinterpret_cast<long long*>(peers.base[threadIdx.x]) + hdr_ack(slot, rank), m);
}

__global__ void __launch_bounds__(256)
pack_exchange_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts,
                     const int32_t* __restrict__ assign_stats, const int32_t* __restrict__ mining_stats, int batch,
                     int max_total, int capacity, PeerArenas peers, int world, int rank, int slot,
                     int32_t* __restrict__ stats_out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    long long* hdr = reinterpret_cast<long long*>(peers.base[rank]);
    const long long m = hdr[hdr_opened(slot)];             // written by this graph's exchange_open_kernel
    const int words = shard_row_words(max_total);
    const int vecs = words >> 2;
    const int b = blockIdx.x;
    const bool padding = b >= batch;
    const int4 st = padding ? make_int4(0, 0, 0, 0) : row_stats(counts, assign_stats, mining_stats, b);
    // the row is read once, into registers (the loads go out before the flow-control poll) ...
    const float* src = dets + (size_t)b * max_total * 6;
    constexpr int kVecPerThread = 4;                       // 256 threads x 4 x 16 bytes = 16 KB per row (T <= 680)
    float4 v[kVecPerThread];
#pragma unroll
    for (int u = 0; u < kVecPerThread; ++u) {
        const int e = (threadIdx.x + u * 256) * 4;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < words) {
            // src rows start 8-byte aligned only (T*6 floats per image): 8-byte loads inside the detections,
            // word-wise across the boundary to the statistics
            if (!padding && e + 4 <= max_total * 6) {
                const float2 lo = *reinterpret_cast<const float2*>(src + e), hi = *reinterpret_cast<const float2*>(src + e + 2);
                v[u] = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else {
                v[u] = make_float4(row_word(src, st, padding, max_total, e), row_word(src, st, padding, max_total, e + 1),
                                   row_word(src, st, padding, max_total, e + 2), row_word(src, st, padding, max_total, e + 3));
            }
        }
    }
    // ... every rank has released the slot's previous contents (one thread per rank polls LOCAL memory) ...
    if (threadIdx.x < world) wait_word(hdr, hdr + hdr_ack(slot, threadIdx.x), m);
    __syncthreads();
    // ... and stored `world` times: the own arena first, the peers starting with a different one on every rank
    const size_t slot_floats = (size_t)world * capacity * words;
    const size_t row_off = (size_t)slot * slot_floats + ((size_t)rank * capacity + b) * words;
    for (int i = 0; i < world; ++i) {
        const int r = (rank + i) % world;
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(peers.base[r] + header_bytes()) + row_off);
#pragma unroll
        for (int u = 0; u < kVecPerThread; ++u) {
            const int q = threadIdx.x + u * 256;
            if (q < vecs) dst[q] = v[u];
        }
    }
    if (!padding && stats_out && threadIdx.x == 0) reinterpret_cast<int4*>(stats_out)[b] = st;
    // the CTA barrier orders every thread's stores before the release stores below, which are cumulative: rank r
    // sees the whole row once it has acquired rowflag[slot][rank][b] == m
    __syncthreads();
    if (threadIdx.x < world)
        st_release_sys(reinterpret_cast<long long*>(peers.base[threadIdx.x]) + hdr_rowflag(slot, rank, b), m);
}

__global__ void __launch_bounds__(256) exchange_wait_kernel(long long* hdr, int world, int capacity, int slot) {
    griddep_wait();
    griddep_launch_dependents();
    const long long m = hdr[hdr_opened(slot)];
    for (int i = threadIdx.x; i < world * capacity; i += blockDim.x)
        wait_word(hdr, hdr + hdr_rowflag(slot, i / capacity, i % capacity), m);
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_shard_row_words(int max_total) { return max_total < 0 ? 0 : shard_row_words(max_total); }

extern "C" int ssd_pack_shard(const float* dets, const int32_t* counts, const int32_t* assign_stats,
                              const int32_t* mining_stats, int batch, int max_total, int capacity, float* shard_out,
                              int32_t* stats_out, void* stream) {
    SSD_REQUIRE(batch >= 0 && max_total >= 0 && capacity >= batch, SSD_ERR_INVALID_ARGUMENT,
                "ssd_pack_shard: batch %d, max_total %d, capacity %d", batch, max_total, capacity);
    if (capacity == 0) return SSD_OK;
    SSD_REQUIRE(shard_out && (batch == 0 || (dets && counts)), SSD_ERR_INVALID_ARGUMENT, "ssd_pack_shard: null pointer");
    SSD_CUDA(launch_pdl(pack_shard_kernel, dim3(capacity), dim3(256), 0, (cudaStream_t)stream, dets, counts, assign_stats,
                        mining_stats, batch, max_total, shard_out, stats_out));
    count_launch();
    return SSD_OK;
}

// The kernels of `device` dereference memory that lives on `peer_device` (an IPC mapping opened in this process):
// peer access has to be enabled in that direction.  "Already enabled" is not an error.
extern "C" int ssd_exchange_enable_peer(int device, int peer_device) {
    if (device == peer_device) return SSD_OK;
    int prev = 0;
    SSD_CUDA(cudaGetDevice(&prev));
    int can = 0;
    SSD_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
    SSD_REQUIRE(can, SSD_ERR_UNSUPPORTED, "ssd_exchange_enable_peer: device %d cannot access device %d", device, peer_device);
    SSD_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
    (void)cudaSetDevice(prev);
    SSD_CUDA(e);
    return SSD_OK;
}

extern "C" size_t ssd_exchange_arena_bytes(int world, int slots, int capacity, int max_total) {
    if (world < 1 || world > kMaxWorld || slots < 1 || slots > kMaxSlots || capacity < 0 || max_total < 0) return 0;
    if (capacity > kMaxRows) return 0;
    return header_bytes() + round_up((size_t)slots * world * capacity * shard_row_words(max_total) * sizeof(float), 256);
}

extern "C" size_t ssd_exchange_slot_offset(int world, int slot, int capacity, int max_total) {
    return header_bytes() + (size_t)slot * world * capacity * shard_row_words(max_total) * sizeof(float);
}

static int fill_arenas(void* const* peer_arenas, int world, int rank, int slot, PeerArenas& pa, const char* who) {
    SSD_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && slot >= 0 && slot < kMaxSlots,
                SSD_ERR_INVALID_ARGUMENT, "%s: world %d, rank %d, slot %d", who, world, rank, slot);
    SSD_REQUIRE(peer_arenas != nullptr, SSD_ERR_INVALID_ARGUMENT, "%s: null arena table", who);
    memset(&pa, 0, sizeof(pa));
    for (int r = 0; r < world; ++r) {
        SSD_REQUIRE(peer_arenas[r] != nullptr && aligned(peer_arenas[r], 256), SSD_ERR_INVALID_ARGUMENT,
                    "%s: arena of rank %d is null or not 256-byte aligned", who, r);
        pa.base[r] = (unsigned char*)peer_arenas[r];
    }
    return SSD_OK;
}

extern "C" int ssd_pack_exchange(const float* dets, const int32_t* counts, const int32_t* assign_stats,
                                 const int32_t* mining_stats, int batch, int max_total, int capacity,
                                 void* const* peer_arenas, int world, int rank, int slot, int32_t* stats_out,
                                 void* stream) {
    SSD_REQUIRE(batch >= 0 && max_total >= 0 && capacity >= batch && capacity >= 1, SSD_ERR_INVALID_ARGUMENT,
                "ssd_pack_exchange: batch %d, max_total %d, capacity %d", batch, max_total, capacity);
    SSD_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && slot >= 0 && slot < kMaxSlots,
                SSD_ERR_INVALID_ARGUMENT, "ssd_pack_exchange: world %d, rank %d, slot %d", world, rank, slot);
    SSD_REQUIRE(peer_arenas && (batch == 0 || (dets && counts)), SSD_ERR_INVALID_ARGUMENT, "ssd_pack_exchange: null pointer");
    SSD_REQUIRE(capacity <= kMaxRows, SSD_ERR_UNSUPPORTED, "ssd_pack_exchange: more than %d images per rank", kMaxRows);
    SSD_REQUIRE(shard_row_words(max_total) <= 256 * 4 * 4, SSD_ERR_UNSUPPORTED, "ssd_pack_exchange: max_total %d too large", max_total);
    SSD_REQUIRE(batch == 0 || aligned(dets, 8), SSD_ERR_MISALIGNED, "ssd_pack_exchange: dets must be 8-byte aligned");
    SSD_REQUIRE(!stats_out || aligned(stats_out, 16), SSD_ERR_MISALIGNED, "ssd_pack_exchange: stats_out must be 16-byte aligned");
    PeerArenas pa;
    const int rc = fill_arenas(peer_arenas, world, rank, slot, pa, "ssd_pack_exchange");
    if (rc != SSD_OK) return rc;
    SSD_CUDA(launch_pdl(pack_exchange_kernel, dim3(capacity), dim3(256), 0, (cudaStream_t)stream, dets, counts,
                        assign_stats, mining_stats, batch, max_total, capacity, pa, world, rank, slot, stats_out));
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_exchange_open(void* const* peer_arenas, int world, int rank, int slot, void* stream) {
    PeerArenas pa;
    const int rc = fill_arenas(peer_arenas, world, rank, slot, pa, "ssd_exchange_open");
    if (rc != SSD_OK) return rc;
    SSD_CUDA(launch_pdl(exchange_open_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, pa, world, rank, slot));
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_exchange_wait(void* own_arena, int world, int capacity, int slot, void* stream) {
    SSD_REQUIRE(own_arena && world >= 1 && world <= kMaxWorld && slot >= 0 && slot < kMaxSlots && capacity >= 1 &&
                    capacity <= kMaxRows, SSD_ERR_INVALID_ARGUMENT, "ssd_exchange_wait: bad argument");
    SSD_CUDA(launch_pdl(exchange_wait_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, (long long*)own_arena, world,
                        capacity, slot));
    count_launch();
    return SSD_OK;
}

// ---- the arena itself: plain cudaMalloc memory exported / imported with the CUDA IPC calls (no framework in between:
//      explicit open and close, nothing to leak at exit) ----
extern "C" int ssd_exchange_arena_alloc(size_t bytes, void** arena_out, void* ipc_handle_out64) {
    SSD_REQUIRE(arena_out && ipc_handle_out64 && bytes > 0, SSD_ERR_INVALID_ARGUMENT, "ssd_exchange_arena_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the header promises 64-byte handles");
    void* p = nullptr;
    SSD_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "ssd_exchange_arena_alloc"); }
    memcpy(ipc_handle_out64, &h, sizeof(h));
    *arena_out = p;
    return SSD_OK;
}
extern "C" int ssd_exchange_arena_free(void* arena) {
    if (arena) SSD_CUDA(cudaFree(arena));
    return SSD_OK;
}
extern "C" int ssd_exchange_peer_open(const void* ipc_handle64, void** mapped_out) {
    SSD_REQUIRE(ipc_handle64 && mapped_out, SSD_ERR_INVALID_ARGUMENT, "ssd_exchange_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, sizeof(h));
    SSD_CUDA(cudaIpcOpenMemHandle(mapped_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SSD_OK;
}
extern "C" int ssd_exchange_peer_close(void* mapped) {
    if (mapped) SSD_CUDA(cudaIpcCloseMemHandle(mapped));
    return SSD_OK;
}

SSD_DEFINE_TRACE_SETTER(set_trace_exchange)
