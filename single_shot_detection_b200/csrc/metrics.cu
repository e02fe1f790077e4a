// Detection metric: detection/metrics/mean_average_precision.py:10-116 (the step right after the
// post-processor in bf/eval.py:54-70).  The reference walks the score-sorted detections in a Python
// loop -- one box_utils.iou call per detection -- and keeps `matched` sets per (image, class).
// Here the detections stay on the device (map_append_kernel compacts the post-processor's padded
// [B, T, 6] output into [N, 7] rows step after step) and the loop becomes data-parallel launches:
//
//  0. map_keys_kernel    64-bit sort key per detection: class in the high word, the complement of
//     the order-preserving score key in the low word, so ONE ascending radix sort (CUB -- library
//     code, like cuBLAS for a plain GEMM) yields class-major / descending-score order.  Only
//     detections of a box's own class can select it, so the rank INSIDE the class decides the greedy
//     matching exactly as the global rank does.  The same launch counts the non-difficult boxes per class;
//  1. map_match_kernel   one thread per detection (in sorted order): IoU
//     (bf/utils/box_utils.py:83-101, clamped areas, separately rounded fp32 ops) against the boxes
//     of its class in its image, first maximum, `value > threshold` in fp32.  The greedy rule "the
//     first detection that reaches an unmatched box takes it" needs no sequential walk: a box is
//     taken by the LOWEST-ranked detection that selects it, so every selecting detection does one
//     atomicMin(first[box], rank);
//  2. map_flag_kernel    TP if first[box] == rank, FP otherwise; a difficult box counts as neither;
//  3. map_ap_kernel      one CTA per class over its detections (sorted by class, then score):
//     block scans give cumulative tp / fp, precision = tp / (tp + fp) with a trailing 0, a
//     right-to-left NaN-propagating max scan gives the envelope (torch.max semantics), then VOC
//     11-point or area AP exactly as the reference composes them.
#include <math.h>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace ssd {

constexpr int kMapThreads = 256;
constexpr int kMapFlagTp = 1, kMapFlagFp = 2;

// ---- accumulation: padded post-processor output -> [N, 7] rows (bf/eval.py:54-59 without the per-image cat) ----
// one CTA per image; its write offset is the sum of the counts before it (deterministic order)
__global__ void __launch_bounds__(kMapThreads)
map_append_kernel(const float* __restrict__ dets, const int32_t* __restrict__ counts, int batch, int max_total,
                  int image_base, float* __restrict__ rows, int64_t capacity, const int64_t* __restrict__ cursor_in,
                  int64_t* __restrict__ cursor_out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    __shared__ int red[kMapThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    int before = 0, all = 0;
    for (int i = tid; i < batch; i += kMapThreads) {
        const int c = min(max(counts[i], 0), max_total);
        all += c;
        if (i < b) before += c;
    }
    before = __reduce_add_sync(FULL, before);
    all = __reduce_add_sync(FULL, all);
    if (lane_id() == 0) red[warp_id()] = before;
    __syncthreads();
    before = 0;
    for (int w = 0; w < kMapThreads / 32; ++w) before += red[w];
    __syncthreads();
    if (lane_id() == 0) red[warp_id()] = all;
    __syncthreads();
    all = 0;
    for (int w = 0; w < kMapThreads / 32; ++w) all += red[w];
    const int64_t base = cursor_in[0];
    if (b == 0 && tid == 0) {
        cursor_out[0] = base + all;                          // the caller checks it against the capacity
        cursor_out[1] = capacity;
    }
    const int n = min(max(counts[b], 0), max_total);
    const float img = (float)(image_base + b);              // torch.full(..., index, dtype=float32), eval.py:57
    for (int e = tid; e < n * 7; e += kMapThreads) {
        const int r = e / 7, c = e - r * 7;
        const int64_t dst = base + before + r;
        if (dst < capacity) rows[dst * 7 + c] = c == 0 ? img : dets[((size_t)b * max_total + r) * 6 + (c - 1)];
    }
}

// ---- sort keys + non-difficult box counts ----
__global__ void __launch_bounds__(kMapThreads)
map_keys_kernel(const float* __restrict__ preds, int64_t n, const float* __restrict__ gt_rows, int gt_cols, int total_gt,
                int num_classes, int difficult_col, uint64_t* __restrict__ keys, uint32_t* __restrict__ index,
                int32_t* __restrict__ totals) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const float cf = preds[k * 7 + 5];
        long long c = (long long)cf;                                       // int(pred[5].item())
        if (!(cf == cf) || c < 0 || c >= num_classes) c = num_classes;      // no such class in the ground truth
        // descending score, NaN first (torch.argsort(descending=True))
        keys[k] = ((uint64_t)c << 32) | (uint32_t)~ordered_key(preds[k * 7 + 6]);
        index[k] = (uint32_t)k;
    } else if (k < n + total_gt) {
        const float* row = gt_rows + (size_t)(k - n) * gt_cols;
        const long long c = (long long)row[SSD_CLASS_COL];                  // .long()
        if (c >= 0 && c < num_classes && (difficult_col < 0 || row[difficult_col] == 0.f)) atomicAdd(totals + c, 1);
    }
}

__global__ void __launch_bounds__(kMapThreads)
map_match_kernel(const float* __restrict__ preds, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ order,
                 int64_t n, const float* __restrict__ gt_rows, int gt_cols, const int32_t* __restrict__ gt_offsets,
                 int num_images, int num_classes, float iou_thr, int difficult_col, int32_t* __restrict__ first,
                 int32_t* __restrict__ best_out, int32_t* __restrict__ seg_offsets) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    // segment boundaries of the class-major order: seg_offsets[c] = first position of a class >= c
    const int seg = (int)(keys[k] >> 32);
    const int seg_prev = k > 0 ? (int)(keys[k - 1] >> 32) : -1;
    for (int c = seg_prev + 1; c <= seg; ++c) seg_offsets[c] = (int32_t)k;
    if (k == n - 1)
        for (int c = seg + 1; c <= num_classes + 1; ++c) seg_offsets[c] = (int32_t)n;
    const float* p = preds + (size_t)order[k] * 7;
    const float imgf = p[0];
    const int img = (int)imgf;                                   // int(pred[0].item())
    int best = -1;
    float best_v = 0.f;
    if (seg < num_classes && img >= 0 && img < num_images) {
        const float4 a = make_float4(p[1], p[2], p[3], p[4]);
        const float area_a = fmul(fmaxf(fsub(a.z, a.x), 0.f), fmaxf(fsub(a.w, a.y), 0.f));
        for (int g = gt_offsets[img]; g < gt_offsets[img + 1]; ++g) {
            const float* row = gt_rows + (size_t)g * gt_cols;
            if ((long long)row[SSD_CLASS_COL] != (long long)seg) continue;
            const float4 b = make_float4(row[0], row[1], row[2], row[3]);
            const float iw = fmaxf(fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
            const float ih = fmaxf(fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
            const float inter = fmul(iw, ih);
            const float area_b = fmul(fmaxf(fsub(b.z, b.x), 0.f), fmaxf(fsub(b.w, b.y), 0.f));
            const float v = fdiv(inter, fsub(fadd(area_a, area_b), inter));
            // torch.max(dim=0): first maximum, NaN propagates and sticks
            if (best < 0 || (!(best_v != best_v) && !(v <= best_v))) { best_v = v; best = g; }
        }
    }
    const bool above = best >= 0 && best_v > iou_thr;
    best_out[k] = above ? best : -1;
    if (above) {
        const bool difficult = difficult_col >= 0 && gt_rows[(size_t)best * gt_cols + difficult_col] != 0.f;
        if (!difficult) atomicMin(first + best, (int32_t)k);
    }
}

__global__ void __launch_bounds__(kMapThreads)
map_flag_kernel(int64_t n, const float* __restrict__ gt_rows, int gt_cols, int difficult_col,
                const int32_t* __restrict__ first, const int32_t* __restrict__ best, uint8_t* __restrict__ flags) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint8_t f = kMapFlagFp;
    const int g = best[k];
    if (g >= 0) {
        const bool difficult = difficult_col >= 0 && gt_rows[(size_t)g * gt_cols + difficult_col] != 0.f;
        f = difficult ? 0 : (first[g] == (int32_t)k ? kMapFlagTp : kMapFlagFp);
    }
    flags[k] = f;
}

// torch.max on two scalars: NaN if either is NaN
__device__ __forceinline__ float nanmax(float a, float b) { return (a != a || b != b) ? NAN : fmaxf(a, b); }

struct MapScan {
    int warp_tp[kMapThreads / 32], warp_fp[kMapThreads / 32];
    float warp_env[kMapThreads / 32];
    float red[kMapThreads / 32];
    int cnt[11];
};

__global__ void __launch_bounds__(kMapThreads)
map_ap_kernel(const uint8_t* __restrict__ flags, const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ totals,
              int voc, float* __restrict__ scratch, int64_t scratch_half, float* __restrict__ ap_out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    __shared__ MapScan sh;
    const int s = blockIdx.x;
    const int lo = seg_offsets[s], n = seg_offsets[s + 1] - lo;
    const int lane = lane_id(), wid = warp_id(), tid = threadIdx.x;
    constexpr int nw = kMapThreads / 32;
    if (totals[s] <= 0) {              // not a key of total_positive (:26-33): the class does not enter the mean
        if (tid == 0) ap_out[s] = -1.f;
        return;
    }
    const float total = (float)totals[s];
    if (n == 0) {                      // tp = [0], fp = [1]: precision (0, 0), AP 0 either way
        if (tid == 0) ap_out[s] = 0.f;
        return;
    }
    float* prec = scratch + lo + s;                         // n + 1 entries
    float* rec = scratch + scratch_half + lo + s;           // n entries
    float thr[11];
#pragma unroll
    for (int t = 0; t < 11; ++t) thr[t] = (float)((double)t * 0.1);      // torch.arange(0, 1.1, .1), bit-identical (tests)
    int below[11];
#pragma unroll
    for (int t = 0; t < 11; ++t) below[t] = 0;
    if (tid < 11) sh.cnt[tid] = 0;

    // ---- pass 1, left to right: cumulative tp / fp, raw precision, recall ----
    int carry_tp = 0, carry_fp = 0;
    for (int base = 0; base < n; base += kMapThreads) {
        const int k = base + tid;
        const uint8_t f = k < n ? flags[lo + k] : 0;
        int tp = f == kMapFlagTp, fp = f == kMapFlagFp;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(FULL, tp, o), b = __shfl_up_sync(FULL, fp, o);
            if (lane >= o) { tp += a; fp += b; }
        }
        __syncthreads();                                    // the previous round's warp totals are consumed
        if (lane == 31) { sh.warp_tp[wid] = tp; sh.warp_fp[wid] = fp; }
        __syncthreads();
        int off_tp = carry_tp, off_fp = carry_fp, all_tp = 0, all_fp = 0;
#pragma unroll
        for (int w = 0; w < nw; ++w) {
            if (w < wid) { off_tp += sh.warp_tp[w]; off_fp += sh.warp_fp[w]; }
            all_tp += sh.warp_tp[w]; all_fp += sh.warp_fp[w];
        }
        tp += off_tp; fp += off_fp;
        carry_tp += all_tp; carry_fp += all_fp;
        if (k < n) {
            const float ftp = (float)tp, ffp = (float)fp;
            prec[k] = fdiv(ftp, fadd(ftp, ffp));
            const float r = fdiv(ftp, total);
            rec[k] = r;
#pragma unroll
            for (int t = 0; t < 11; ++t) below[t] += thr[t] > r;
        }
    }
    if (tid == 0) prec[n] = 0.f;
    if (voc) {
#pragma unroll
        for (int t = 0; t < 11; ++t) {
            const int c = __reduce_add_sync(FULL, below[t]);
            if (lane == 0 && c) atomicAdd(&sh.cnt[t], c);
        }
    }
    __syncthreads();

    // ---- pass 2, right to left: precision envelope (and the area sum) ----
    float carry_env = 0.f;                                  // the trailing 0
    float area = 0.f;
    const int rounds = (n + kMapThreads - 1) / kMapThreads;
    for (int r_ = rounds - 1; r_ >= 0; --r_) {
        const int k = r_ * kMapThreads + tid;
        float v = k < n ? prec[k] : -INFINITY;              // -inf is neutral for the max
        // suffix scan inside the warp (towards lower lanes)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_down_sync(FULL, v, o);
            if (lane + o < 32) v = nanmax(v, t);
        }
        __syncthreads();
        if (lane == 0) sh.warp_env[wid] = v;
        __syncthreads();
        float right = carry_env, all = carry_env;
#pragma unroll
        for (int w = nw - 1; w >= 0; --w) {
            if (w > wid) right = nanmax(right, sh.warp_env[w]);
            all = nanmax(all, sh.warp_env[w]);
        }
        v = nanmax(v, right);
        carry_env = all;
        if (k < n) {
            prec[k] = v;
            if (!voc) {
                const float prev = k > 0 ? rec[k - 1] : 0.f;
                area = fadd(area, fmul(fsub(rec[k], prev), v));         // (recall[k+1] - recall[k]) * precision[k]
            }
        }
    }
    __syncthreads();
    float result;
    if (voc) {
        // recall + trailing 1: index = number of entries the threshold exceeds (the 1 never is: thr <= 1)
        float acc = 0.f;
        if (tid == 0) {
            for (int t = 0; t < 11; ++t) {
                int idx = sh.cnt[t] + (thr[t] > 1.f ? 1 : 0);
                if (idx > n) idx = n;
                acc = fadd(acc, prec[idx]);
            }
            ap_out[s] = fdiv(acc, 11.f);
        }
        return;
    }
    // area: + (1 - recall[n-1]) * precision[n] (= 0 unless NaN), block sum in fp32
    if (tid == 0) area = fadd(area, fmul(fsub(1.f, rec[n - 1]), prec[n]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) area = fadd(area, __shfl_xor_sync(FULL, area, o));
    if (lane == 0) sh.red[wid] = area;
    __syncthreads();
    if (tid == 0) {
        result = 0.f;
        for (int w = 0; w < nw; ++w) result = fadd(result, sh.red[w]);
        ap_out[s] = result;
    }
}

// mean over the classes that have non-difficult ground truth, summed in double in class order (:115)
__global__ void map_mean_kernel(const float* __restrict__ ap, int num_classes, double* __restrict__ out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double sum = 0.0;
    int cnt = 0;
    for (int c = 0; c < num_classes; ++c) {
        const float v = ap[c];
        if (v == -1.f) continue;
        sum += (double)v;
        ++cnt;
    }
    out[0] = cnt ? sum / (double)cnt : NAN;
    out[1] = (double)cnt;
}

struct MapLayout {
    size_t keys_in, keys_out, index_in, index_out, first, best, totals, seg, scratch, cub, end;
    size_t cub_bytes;
};

static int map_end_bit(int num_classes) {
    int bits = 1;
    while ((1LL << bits) <= (long long)num_classes) ++bits;     // classes 0 .. num_classes (the overflow bucket)
    return 32 + bits;
}

static cudaError_t map_layout(int64_t count, int total_gt, int num_classes, MapLayout* L) {
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)count, 0,
                                                    map_end_bit(num_classes), (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = round_up(off + bytes, 256); return o; };
    L->keys_in = take(8 * (size_t)count);
    L->keys_out = take(8 * (size_t)count);
    L->index_in = take(4 * (size_t)count);
    L->index_out = take(4 * (size_t)count);
    L->first = take(4 * (size_t)total_gt);
    L->best = take(4 * (size_t)count);
    L->totals = take(4 * (size_t)(num_classes + 1));
    L->seg = take(4 * (size_t)(num_classes + 2));
    L->scratch = take(4 * 2 * (size_t)(count + num_classes + 1));
    L->cub = take(cub_bytes);
    L->cub_bytes = cub_bytes;
    L->end = off;
    return cudaSuccess;
}

}  // namespace ssd

using namespace ssd;

extern "C" size_t ssd_map_workspace_bytes(int64_t count, int total_gt, int num_classes) {
    if (count < 0 || total_gt < 0 || num_classes < 0) return 0;
    MapLayout L;
    if (map_layout(count, total_gt, num_classes, &L) != cudaSuccess) return 0;
    return L.end;
}

extern "C" int ssd_map_append(const float* dets, const int32_t* counts, int batch, int max_total, int image_base,
                              float* rows, int64_t capacity, const int64_t* cursor_in, int64_t* cursor_out,
                              void* stream) {
    SSD_REQUIRE(batch >= 0 && max_total >= 0 && capacity >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_map_append: negative size");
    SSD_REQUIRE(cursor_in && cursor_out && cursor_in != cursor_out, SSD_ERR_INVALID_ARGUMENT,
                "ssd_map_append: cursor_in / cursor_out must be two different device words");
    if (batch == 0) {
        SSD_CUDA(cudaMemcpyAsync(cursor_out, cursor_in, sizeof(int64_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return SSD_OK;
    }
    SSD_REQUIRE(dets && counts && rows, SSD_ERR_INVALID_ARGUMENT, "ssd_map_append: null pointer");
    SSD_CUDA(launch_pdl(map_append_kernel, dim3(batch), dim3(kMapThreads), 0, (cudaStream_t)stream, dets, counts, batch,
                        max_total, image_base, rows, capacity, cursor_in, cursor_out));
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_mean_average_precision(const float* preds, int64_t count, const float* gt_rows, int gt_cols,
                                          const int32_t* gt_offsets, int num_images, int total_gt, int num_classes,
                                          float iou_threshold, int use_difficult, int voc, float* ap_out,
                                          double* map_out, uint8_t* flags_out, uint32_t* order_out, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    SSD_REQUIRE(count >= 0 && num_images >= 0 && total_gt >= 0 && num_classes >= 0, SSD_ERR_INVALID_ARGUMENT,
                "ssd_mean_average_precision: negative size");
    SSD_REQUIRE(count < 0x7F000000LL, SSD_ERR_UNSUPPORTED, "ssd_mean_average_precision: too many detections");
    SSD_REQUIRE(map_out && flags_out && order_out && (ap_out || num_classes == 0), SSD_ERR_INVALID_ARGUMENT,
                "ssd_mean_average_precision: null output pointer");
    SSD_REQUIRE((preds || count == 0) && gt_offsets && (gt_rows || total_gt == 0) && workspace, SSD_ERR_INVALID_ARGUMENT,
                "ssd_mean_average_precision: null input pointer");
    SSD_REQUIRE(gt_cols >= 6 && (!use_difficult || gt_cols >= 7), SSD_ERR_INVALID_ARGUMENT,
                "ssd_mean_average_precision: ground-truth rows need >= 6 columns (7 with a difficult flag), got %d", gt_cols);
    MapLayout L;
    SSD_CUDA(map_layout(count, total_gt, num_classes, &L));
    SSD_REQUIRE(workspace_bytes >= L.end && aligned(workspace, 256), SSD_ERR_WORKSPACE,
                "ssd_mean_average_precision: workspace needs %zu bytes (256-byte aligned), got %zu", L.end, workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    uint64_t* keys_in = (uint64_t*)(ws + L.keys_in);
    uint64_t* keys_out = (uint64_t*)(ws + L.keys_out);
    uint32_t* index_in = (uint32_t*)(ws + L.index_in);
    int32_t* first = (int32_t*)(ws + L.first);
    int32_t* best = (int32_t*)(ws + L.best);
    int32_t* totals = (int32_t*)(ws + L.totals);
    int32_t* seg = (int32_t*)(ws + L.seg);
    const int dcol = use_difficult ? 6 : -1;
    // totals and seg are adjacent: zero both (seg stays all-zero when there is no detection)
    SSD_CUDA(cudaMemsetAsync(totals, 0, L.scratch - L.totals, st));
    if (total_gt) SSD_CUDA(cudaMemsetAsync(first, 0x7F, sizeof(int32_t) * (size_t)total_gt, st));      // > any rank
    const int64_t items = count + total_gt;
    if (items) {
        SSD_CUDA(launch_pdl(map_keys_kernel, dim3((unsigned)((items + kMapThreads - 1) / kMapThreads)), dim3(kMapThreads),
                            0, st, preds, count, gt_rows, gt_cols, total_gt, num_classes, dcol, keys_in, index_in, totals));
        count_launch();
    }
    if (count) {
        size_t cub_bytes = L.cub_bytes;
        SSD_CUDA(cub::DeviceRadixSort::SortPairs((void*)(ws + L.cub), cub_bytes, (const uint64_t*)keys_in, keys_out,
                                                 (const uint32_t*)index_in, order_out, (int64_t)count, 0,
                                                 map_end_bit(num_classes), st));
        const unsigned blocks = (unsigned)((count + kMapThreads - 1) / kMapThreads);
        SSD_CUDA(launch_pdl(map_match_kernel, dim3(blocks), dim3(kMapThreads), 0, st, preds, (const uint64_t*)keys_out,
                            (const uint32_t*)order_out, count, gt_rows, gt_cols, gt_offsets, num_images, num_classes,
                            iou_threshold, dcol, first, best, seg));
        count_launch();
        SSD_CUDA(launch_pdl(map_flag_kernel, dim3(blocks), dim3(kMapThreads), 0, st, count, gt_rows, gt_cols, dcol,
                            (const int32_t*)first, (const int32_t*)best, flags_out));
        count_launch();
    }
    if (num_classes) {
        SSD_CUDA(launch_pdl(map_ap_kernel, dim3(num_classes), dim3(kMapThreads), 0, st, (const uint8_t*)flags_out,
                            (const int32_t*)seg, (const int32_t*)totals, voc, (float*)(ws + L.scratch),
                            (int64_t)(count + num_classes + 1), ap_out));
        count_launch();
    }
    SSD_CUDA(launch_pdl(map_mean_kernel, dim3(1), dim3(32), 0, st, (const float*)ap_out, num_classes, map_out));
    count_launch();
    return SSD_OK;
}

SSD_DEFINE_TRACE_SETTER(set_trace_metrics)
