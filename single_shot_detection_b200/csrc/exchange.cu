// The one exchange step of the path: padded detections, their counts and the matched-target
// statistics of the local images go into ONE [capacity, T*6 + 5] fp32-word buffer per rank.
//
//  * pack_shard_kernel      packs the local buffer (one launch instead of the ~10 element-wise torch kernels the
//    same packing costs with tensor ops); NCCL's all-gather then moves it (sharding.all_gather_packed).
//  * pack_exchange_kernel   packs AND exchanges in the same launch: every CTA writes its row straight into slot
//    `rank` of EVERY peer's gathered buffer through NVLink peer memory (the buffers are CUDA-IPC mappings of one
//    arena per rank, sharding.PeerExchange) -- no NCCL call, no host involvement, capturable into the step graph.
//    With NCCL the exchange cost 11.5 us of every 47 us step at N = 2 (host enqueue + a kernel of its own).
//    Protocol per ring slot k (one slot per captured step graph; its launches are serialised on one stream):
//        n = count[k]                                  launches of this slot so far (local)
//        wait until flag[k][r] >= n for every peer r   (every peer has finished ITS launch n: flow control, bounded
//                                                       spin on LOCAL memory, time-out -> error word, never a hang)
//        write the rows into slot k of every peer
//        __threadfence_system(); ticket;               the last CTA: fence, count[k] = n + 1, and
//        flag_on_peer[k][rank] = n + 1                 (system-scope store) on every peer
//    exchange_wait_kernel(k) spins until flag[k][r] >= count[k] for all r: the gathered slot is complete.
//    A gathered slot stays valid until its step graph is launched again (the contract of AnchorPipeline.stream's
//    slots).
#include "common.cuh"

namespace ssd {

__device__ __forceinline__ void pack_row(const float* __restrict__ dets, const int32_t* __restrict__ counts,
                                         const int32_t* __restrict__ assign_stats, const int32_t* __restrict__ mining_stats,
                                         int batch, int max_total, int b, float* __restrict__ row,
                                         int32_t* __restrict__ stats_out) {
    const int words = max_total * 6 + 5;
    int32_t* irow = reinterpret_cast<int32_t*>(row);
    if (b >= batch) {                                   // padding row: count = -1, everything else 0
        for (int e = threadIdx.x; e < words; e += blockDim.x) row[e] = e == max_total * 6 ? __int_as_float(-1) : 0.f;
        return;
    }
    const float* src = dets + (size_t)b * max_total * 6;
    for (int e = threadIdx.x; e < max_total * 6; e += blockDim.x) row[e] = src[e];
    if (threadIdx.x == 0) {
        // {positives, hard negatives selected, ignored, detections}
        const int32_t s0 = assign_stats ? assign_stats[b * 4 + 0] : 0;
        const int32_t s1 = mining_stats ? mining_stats[b * 4 + 2] : 0;
        const int32_t s2 = assign_stats ? assign_stats[b * 4 + 1] : 0;
        const int32_t s3 = counts[b];
        irow[max_total * 6] = s3;
        irow[max_total * 6 + 1] = s0; irow[max_total * 6 + 2] = s1; irow[max_total * 6 + 3] = s2; irow[max_total * 6 + 4] = s3;
        if (stats_out) { stats_out[b * 4] = s0; stats_out[b * 4 + 1] = s1; stats_out[b * 4 + 2] = s2; stats_out[b * 4 + 3] = s3; }
    }
}

// one CTA per row of the buffer; rows >= batch are padding
__global__ void __launch_bounds__(256)
pack_shard_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts,
                  const int32_t* __restrict__ assign_stats, const int32_t* __restrict__ mining_stats, int batch,
                  int max_total, float* __restrict__ shard, int32_t* __restrict__ stats_out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    pack_row(dets, counts, assign_stats, mining_stats, batch, max_total, blockIdx.x,
             shard + (size_t)blockIdx.x * (max_total * 6 + 5), stats_out);
}

// ---- peer-memory exchange ----
// Arena of one rank (sharding.PeerExchange): [header | slots x world x capacity x words floats]
//   header (int64 words): [0] error, [8 + k] count[k], [8 + kMaxSlots + k * kMaxWorld + r] flag[k][r],
//                         [8 + kMaxSlots + kMaxSlots * kMaxWorld + k] ticket[k]
constexpr int kMaxWorld = SSD_EXCHANGE_MAX_WORLD;
constexpr int kMaxSlots = SSD_EXCHANGE_MAX_SLOTS;
constexpr unsigned long long kSpinLimitNs = 4000000000ull;      // 4 s: a dead peer becomes an error, not a hang

__host__ __device__ inline size_t hdr_count(int k) { return 8 + (size_t)k; }
__host__ __device__ inline size_t hdr_flag(int k, int r) { return 8 + kMaxSlots + (size_t)k * kMaxWorld + r; }
__host__ __device__ inline size_t hdr_ticket(int k) { return 8 + kMaxSlots + (size_t)kMaxSlots * kMaxWorld + k; }
constexpr size_t kHeaderWords = 8 + kMaxSlots + (size_t)kMaxSlots * kMaxWorld + kMaxSlots;
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

// spin until flag[k][r] (LOCAL memory, written by rank r) has reached `need`; false on time-out
__device__ __forceinline__ bool wait_flag(long long* hdr, int k, int r, long long need) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(hdr + hdr_flag(k, r)) < need) {
        if (global_ns() - t0 > kSpinLimitNs) { hdr[0] = 1; return false; }
        __nanosleep(100);
    }
    return true;
}

__global__ void __launch_bounds__(256)
pack_exchange_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts,
                     const int32_t* __restrict__ assign_stats, const int32_t* __restrict__ mining_stats, int batch,
                     int max_total, int capacity, PeerArenas peers, int world, int rank, int slot,
                     int32_t* __restrict__ stats_out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    __shared__ long long s_n;
    __shared__ int s_last;
    long long* hdr = reinterpret_cast<long long*>(peers.base[rank]);
    if (threadIdx.x < world) {                             // one thread per peer: the polls run in parallel
        const long long n = hdr[hdr_count(slot)];          // only the last CTA of the previous launch wrote it
        wait_flag(hdr, slot, threadIdx.x, n);              // that peer has finished its launch n of this slot
        if (threadIdx.x == 0) s_n = n;
    }
    __syncthreads();
    const int words = max_total * 6 + 5;
    const int b = blockIdx.x;
    const size_t slot_floats = (size_t)world * capacity * words;
    // pack once into the own arena, then copy the finished row to the peers (coalesced 4-byte stores over NVLink)
    float* mine = reinterpret_cast<float*>(peers.base[rank] + header_bytes()) + (size_t)slot * slot_floats +
                  ((size_t)rank * capacity + b) * words;
    pack_row(dets, counts, assign_stats, mining_stats, batch, max_total, b, mine, stats_out);
    __syncthreads();
    for (int i = 1; i < world; ++i) {
        const int r = (rank + i) % world;                  // every rank starts with a different peer
        float* dst = reinterpret_cast<float*>(peers.base[r] + header_bytes()) + (size_t)slot * slot_floats +
                     ((size_t)rank * capacity + b) * words;
        for (int e = threadIdx.x; e < words; e += blockDim.x) dst[e] = mine[e];
    }
    // the CTA barrier orders every thread's stores before thread 0's system-scope fence, which is cumulative:
    // ONE fence per CTA (it waits for the NVLink write acknowledgements), not one per thread
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(reinterpret_cast<unsigned long long*>(hdr + hdr_ticket(slot)), 1ull) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) {
        __threadfence_system();
        hdr[hdr_ticket(slot)] = 0;
        hdr[hdr_count(slot)] = s_n + 1;
    }
    __syncthreads();
    if (threadIdx.x < world)
        st_release_sys(reinterpret_cast<long long*>(peers.base[threadIdx.x]) + hdr_flag(slot, rank), s_n + 1);
}

__global__ void exchange_wait_kernel(long long* hdr, int world, int slot) {
    griddep_wait();
    griddep_launch_dependents();
    if (threadIdx.x < world && blockIdx.x == 0) wait_flag(hdr, slot, threadIdx.x, hdr[hdr_count(slot)]);
}

}  // namespace ssd

using namespace ssd;

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
    return header_bytes() + round_up((size_t)slots * world * capacity * (max_total * 6 + 5) * sizeof(float), 256);
}

extern "C" size_t ssd_exchange_slot_offset(int world, int slot, int capacity, int max_total) {
    return header_bytes() + (size_t)slot * world * capacity * (max_total * 6 + 5) * sizeof(float);
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
    PeerArenas pa;
    memset(&pa, 0, sizeof(pa));
    for (int r = 0; r < world; ++r) {
        SSD_REQUIRE(peer_arenas[r] != nullptr && aligned(peer_arenas[r], 256), SSD_ERR_INVALID_ARGUMENT,
                    "ssd_pack_exchange: arena of rank %d is null or not 256-byte aligned", r);
        pa.base[r] = (unsigned char*)peer_arenas[r];
    }
    SSD_CUDA(launch_pdl(pack_exchange_kernel, dim3(capacity), dim3(256), 0, (cudaStream_t)stream, dets, counts,
                        assign_stats, mining_stats, batch, max_total, capacity, pa, world, rank, slot, stats_out));
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_exchange_wait(void* own_arena, int world, int slot, void* stream) {
    SSD_REQUIRE(own_arena && world >= 1 && world <= kMaxWorld && slot >= 0 && slot < kMaxSlots, SSD_ERR_INVALID_ARGUMENT,
                "ssd_exchange_wait: bad argument");
    SSD_CUDA(launch_pdl(exchange_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, (long long*)own_arena, world, slot));
    count_launch();
    return SSD_OK;
}

SSD_DEFINE_TRACE_SETTER(set_trace_exchange)
