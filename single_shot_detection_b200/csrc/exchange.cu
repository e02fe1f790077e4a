// Packing for the one exchange step of the path: padded detections, their counts and
// the matched-target statistics of the local images go into ONE [capacity, T*6 + 5] fp32-word buffer
// that a single all-gather moves.  One launch instead of the ~10 element-wise torch kernels (zeros,
// fills, slice copies) the same packing costs when written with tensor ops -- those ran serially at
// the very end of every step.
#include "common.cuh"

namespace ssd {

// one CTA per row of the buffer; rows >= batch are padding (count = -1, everything else 0)
__global__ void __launch_bounds__(256)
pack_shard_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts,
                  const int32_t* __restrict__ assign_stats, const int32_t* __restrict__ mining_stats, int batch,
                  int max_total, float* __restrict__ shard, int32_t* __restrict__ stats_out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    const int b = blockIdx.x;
    const int words = max_total * 6 + 5;
    float* row = shard + (size_t)b * words;
    int32_t* irow = reinterpret_cast<int32_t*>(row);
    if (b >= batch) {
        for (int e = threadIdx.x; e < words; e += blockDim.x) row[e] = 0.f;
        __syncthreads();
        if (threadIdx.x == 0) irow[max_total * 6] = -1;
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

SSD_DEFINE_TRACE_SETTER(set_trace_exchange)
