// box_utils.nms for LONG lists: bf/utils/box_utils.py:165-194 without a bound on the boxes that enter the NMS
// (max_per_class=None, or max_per_class > 512) -- the case the batched post-processor kernel (one CTA per
// (image, class), <= 512 boxes in shared memory) does not take.  No sample configuration reaches it; it exists so
// that the mirrored API accepts everything the reference accepts.  Not a hot path: plain global-memory kernels.
//
//   keys      (score descending, input index ascending) as one ascending 64-bit key per box
//   sort      bitonic network over global memory: tiles of 2048 keys are sorted / merged in shared memory, the
//             long strides take one launch each (n <= 2^20: a few dozen launches)
//   top-k     the first k sorted keys (bf/utils/box_utils.py:186-188; ties at the cut go to the lower index)
//   hard NMS  torchvision.ops.nms semantics (unclamped areas, inter / (a_i + a_j - inter) compared float-vs-double,
//             keep list in descending score order): 64 x 64 tiles of the suppression bit matrix, then a sweep that
//             resolves 64 rows at a time against the diagonal words and ORs the kept rows' words into the
//             suppressed set
//   soft NMS  the reference's Gaussian loop (box_utils.py:145-163) statement by statement, one CTA, one pick per
//             iteration; the subset is taken in ascending input order like the batched kernel does
#include <math.h>

#include "common.cuh"

namespace ssd {

constexpr int kSortThreads = 1024;
constexpr int kSortTile = 2 * kSortThreads;
constexpr int kLargeMaxBoxes = 1 << 20;

__global__ void large_keys_kernel(const float* __restrict__ scores, int n, int n2, unsigned long long* __restrict__ keys) {
    griddep_wait();
    griddep_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    // ascending key order = descending score, then ascending index; padding sorts last
    keys[i] = i < n ? ((unsigned long long)(~ordered_key(scores[i])) << 32) | (unsigned long long)(uint32_t)i : ~0ull;
}

__device__ __forceinline__ void compare_exchange(unsigned long long& a, unsigned long long& b, bool ascending) {
    if ((a > b) == ascending) { const unsigned long long t = a; a = b; b = t; }
}

// every (size, stride) step with stride < kSortTile happens inside a tile: `first_size` = 2 sorts the tiles from
// scratch, otherwise only the tail of the merge of `first_size` (strides kSortTile / 2 .. 1) is applied
__global__ void __launch_bounds__(kSortThreads)
bitonic_tile_kernel(unsigned long long* __restrict__ keys, int first_size, int last_size) {
    __shared__ unsigned long long s[kSortTile];
    griddep_wait();
    griddep_launch_dependents();
    const size_t base = (size_t)blockIdx.x * kSortTile;
    s[threadIdx.x] = keys[base + threadIdx.x];
    s[threadIdx.x + kSortThreads] = keys[base + threadIdx.x + kSortThreads];
    for (int size = first_size; size <= last_size; size <<= 1) {
        int stride = size >> 1;
        if (stride >= kSortTile) stride = kSortTile >> 1;
        for (; stride > 0; stride >>= 1) {
            __syncthreads();
            const int t = threadIdx.x;
            const int lo = 2 * t - (t & (stride - 1));
            const bool ascending = ((base + lo) & (size_t)size) == 0;
            compare_exchange(s[lo], s[lo + stride], ascending);
        }
    }
    __syncthreads();
    keys[base + threadIdx.x] = s[threadIdx.x];
    keys[base + threadIdx.x + kSortThreads] = s[threadIdx.x + kSortThreads];
}

__global__ void bitonic_global_kernel(unsigned long long* __restrict__ keys, int n2, int size, int stride) {
    griddep_wait();
    griddep_launch_dependents();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)(n2 >> 1)) return;
    const size_t lo = 2 * t - (t & (size_t)(stride - 1));
    const bool ascending = (lo & (size_t)size) == 0;
    unsigned long long a = keys[lo], b = keys[lo + stride];
    if ((a > b) == ascending) { keys[lo] = b; keys[lo + stride] = a; }
}

// subset rows in the order the sweep wants them: rank r -> input row of the r-th key
__global__ void large_gather_kernel(const unsigned long long* __restrict__ keys, const float4* __restrict__ boxes,
                                    const float* __restrict__ scores, int k, int clamp_area, float4* __restrict__ sbox,
                                    float* __restrict__ sarea, float* __restrict__ sscore, int* __restrict__ sidx) {
    griddep_wait();
    griddep_launch_dependents();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= k) return;
    const int i = (int)(uint32_t)(keys[r] & 0xFFFFFFFFull);
    const float4 b = boxes[i];
    sbox[r] = b;
    const float w = fsub(b.z, b.x), h = fsub(b.w, b.y);
    sarea[r] = clamp_area ? fmul(fmaxf(w, 0.f), fmaxf(h, 0.f)) : fmul(w, h);       // box_utils.area / torchvision
    sscore[r] = scores[i];
    sidx[r] = i;
}

// bit j of mask[i][cb] set <=> row i suppresses row cb * 64 + j (only j > i)
__global__ void __launch_bounds__(64)
large_mask_kernel(const float4* __restrict__ sbox, const float* __restrict__ sarea, int k, int words, double thr,
                  unsigned long long* __restrict__ mask) {
    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    griddep_wait();
    griddep_launch_dependents();
    const int cb = blockIdx.x, rb = blockIdx.y;
    if (cb < rb) return;
    const int j0 = cb * 64;
    const int cols = min(64, k - j0);
    if ((int)threadIdx.x < cols) { cbox[threadIdx.x] = sbox[j0 + threadIdx.x]; carea[threadIdx.x] = sarea[j0 + threadIdx.x]; }
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i >= k) return;
    const float4 bi = sbox[i];
    const float ai = sarea[i];
    unsigned long long bits = 0ull;
    for (int j = (cb == rb ? (int)threadIdx.x + 1 : 0); j < cols; ++j) {
        const float4 bj = cbox[j];
        const float iw = fmaxf(0.f, fsub(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
        const float ih = fmaxf(0.f, fsub(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
        const float inter = fmul(iw, ih);
        const float iou = fdiv(inter, fsub(fadd(ai, carea[j]), inter));
        if ((double)iou > thr) bits |= 1ull << j;                         // float-vs-double, as torchvision's CPU kernel
    }
    mask[(size_t)i * words + cb] = bits;
}

__global__ void __launch_bounds__(1024)
large_sweep_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ sidx, int k, int words,
                   unsigned long long* __restrict__ removed, long long* __restrict__ keep_out, int* __restrict__ count_out) {
    __shared__ unsigned long long s_diag[64];
    __shared__ unsigned long long s_kept;
    __shared__ int s_nkeep;
    griddep_wait();
    griddep_launch_dependents();
    for (int w = threadIdx.x; w < words; w += blockDim.x) removed[w] = 0ull;
    if (threadIdx.x == 0) s_nkeep = 0;
    __syncthreads();
    for (int c = 0; c < words; ++c) {
        const int r0 = c * 64;
        const int rows = min(64, k - r0);
        if ((int)threadIdx.x < 64) s_diag[threadIdx.x] = (int)threadIdx.x < rows ? mask[(size_t)(r0 + threadIdx.x) * words + c] : 0ull;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cur = removed[c];
            unsigned long long kept = 0ull;
            int nk = s_nkeep;
            for (int b = 0; b < rows; ++b) {
                if (!((cur >> b) & 1ull)) {
                    kept |= 1ull << b;
                    cur |= s_diag[b];
                    keep_out[nk++] = (long long)sidx[r0 + b];
                }
            }
            s_kept = kept;
            s_nkeep = nk;
        }
        __syncthreads();
        const unsigned long long kept = s_kept;
        for (int w = c + 1 + threadIdx.x; w < words; w += blockDim.x) {
            unsigned long long acc = 0ull;
            for (int b = 0; b < rows; ++b)
                if ((kept >> b) & 1ull) acc |= mask[(size_t)(r0 + b) * words + w];
            removed[w] |= acc;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) count_out[0] = s_nkeep;
}

// ---- soft-NMS over a long list (one CTA; every iteration picks one box) ----
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* s_part) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(FULL, v, o);
        v = t > v ? t : v;
    }
    if (lane_id() == 0) s_part[warp_id()] = v;
    __syncthreads();
    unsigned long long r = lane_id() < (int)(blockDim.x >> 5) ? s_part[lane_id()] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(FULL, r, o);
        r = t > r ? t : r;
    }
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(1024)
large_soft_kernel(const float4* __restrict__ sbox, const float* __restrict__ sarea, float* __restrict__ sc,
                  const int* __restrict__ sidx, int k, float thr, float sigma, long long* __restrict__ keep_out,
                  int* __restrict__ count_out) {
    __shared__ unsigned long long s_part[32];
    griddep_wait();
    griddep_launch_dependents();
    // while mask.nonzero().sum(): a mask whose only element is index 0 ends the loop (the SUM of the indices)
    unsigned long long any = 0ull;
    for (int t = threadIdx.x; t < k; t += blockDim.x) any |= (sc[t] > thr && t != 0) ? 1ull : 0ull;
    bool more = block_max_u64(any, s_part) != 0ull;
    int nkeep = 0;
    while (more && nkeep < k) {
        unsigned long long best = 0ull;                                       // (score, lower index first)
        for (int t = threadIdx.x; t < k; t += blockDim.x) {
            const unsigned long long key = ((unsigned long long)ordered_key(sc[t]) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)t);
            best = key > best ? key : best;
        }
        best = block_max_u64(best, s_part);
        const int bt = (int)(0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull));
        if (threadIdx.x == 0) { keep_out[nkeep] = (long long)sidx[bt]; sc[bt] = 0.f; }
        ++nkeep;
        __syncthreads();
        const float4 bi = sbox[bt];
        const float ai = sarea[bt];
        any = 0ull;
        for (int t = threadIdx.x; t < k; t += blockDim.x) {
            const float v = sc[t];
            if (v > thr) {
                any |= t != 0 ? 1ull : 0ull;
                const float4 bj = sbox[t];
                const float iw = fmaxf(fsub(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), 0.f);
                const float ih = fmaxf(fsub(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)), 0.f);
                const float inter = fmul(iw, ih);
                const float iou = fdiv(inter, fsub(fadd(ai, sarea[t]), inter));
                sc[t] = fmul(v, expf(-fdiv(fmul(iou, iou), sigma)));
            }
        }
        more = block_max_u64(any, s_part) != 0ull;
    }
    if (threadIdx.x == 0) count_out[0] = nkeep;
}

// keys of the soft subset: the k selected input rows in ASCENDING input order
__global__ void large_index_keys_kernel(unsigned long long* __restrict__ keys, int k, int n2) {
    griddep_wait();
    griddep_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n2) return;
    keys[i] = i < k ? (keys[i] & 0xFFFFFFFFull) : ~0ull;
}

struct LargeLayout {
    int n2, k, words;
    size_t keys, sbox, sarea, sscore, sidx, removed, mask, end;
};

static int large_layout(int n, int max_keep, LargeLayout& L) {
    SSD_REQUIRE(n >= 1 && n <= kLargeMaxBoxes, SSD_ERR_UNSUPPORTED, "ssd_nms_large: %d boxes outside 1..%d", n, kLargeMaxBoxes);
    int n2 = kSortTile;
    while (n2 < n) n2 <<= 1;
    L.n2 = n2;
    L.k = (max_keep > 0 && max_keep < n) ? max_keep : n;
    L.words = (L.k + 63) / 64;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += round_up(bytes, 256); return o; };
    L.keys = take((size_t)n2 * 8);
    L.sbox = take((size_t)L.k * 16);
    L.sarea = take((size_t)L.k * 4);
    L.sscore = take((size_t)L.k * 4);
    L.sidx = take((size_t)L.k * 4);
    L.removed = take((size_t)L.words * 8);
    L.mask = take((size_t)L.k * L.words * 8);
    L.end = off;
    return SSD_OK;
}

static int sort_keys(unsigned long long* keys, int n2, cudaStream_t st) {
    const int tiles = n2 / kSortTile;
    SSD_CUDA(launch_pdl(bitonic_tile_kernel, dim3(tiles), dim3(kSortThreads), 0, st, keys, 2, kSortTile));
    count_launch();
    for (int size = 2 * kSortTile; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride >= kSortTile; stride >>= 1) {
            SSD_CUDA(launch_pdl(bitonic_global_kernel, dim3((unsigned)((n2 / 2 + 255) / 256)), dim3(256), 0, st, keys, n2, size, stride));
            count_launch();
        }
        SSD_CUDA(launch_pdl(bitonic_tile_kernel, dim3(tiles), dim3(kSortThreads), 0, st, keys, size, size));
        count_launch();
    }
    return SSD_OK;
}

}  // namespace ssd

using namespace ssd;

extern "C" size_t ssd_nms_large_workspace_bytes(int num_boxes, int max_keep) {
    LargeLayout L;
    if (num_boxes < 1 || large_layout(num_boxes, max_keep, L) != SSD_OK) return 256;
    return L.end;
}

extern "C" int ssd_nms_large(const float* corner_boxes, const float* scores, int num_boxes, int max_keep,
                             double overlap_threshold, int soft, float soft_threshold, float soft_sigma,
                             int64_t* keep_out, int32_t* count_out, void* workspace, size_t workspace_bytes,
                             void* stream) {
    SSD_REQUIRE(num_boxes >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_nms_large: negative box count");
    SSD_REQUIRE(count_out != nullptr, SSD_ERR_INVALID_ARGUMENT, "ssd_nms_large: null count_out");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_boxes == 0) {
        SSD_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), st));
        return SSD_OK;
    }
    LargeLayout L;
    const int rc = large_layout(num_boxes, max_keep, L);
    if (rc != SSD_OK) return rc;
    SSD_REQUIRE(corner_boxes && scores && keep_out && workspace, SSD_ERR_INVALID_ARGUMENT, "ssd_nms_large: null pointer");
    SSD_REQUIRE(aligned(corner_boxes, 16), SSD_ERR_MISALIGNED, "ssd_nms_large: boxes must be 16-byte aligned");
    SSD_REQUIRE(aligned(workspace, 256), SSD_ERR_MISALIGNED, "ssd_nms_large: workspace must be 256-byte aligned");
    SSD_REQUIRE(workspace_bytes >= L.end, SSD_ERR_WORKSPACE, "ssd_nms_large: workspace %zu < %zu bytes", workspace_bytes, L.end);
    SSD_REQUIRE(!soft || soft_sigma > 0.f, SSD_ERR_INVALID_ARGUMENT, "ssd_nms_large: soft-NMS needs sigma > 0");
    unsigned char* ws = (unsigned char*)workspace;
    unsigned long long* keys = (unsigned long long*)(ws + L.keys);
    float4* sbox = (float4*)(ws + L.sbox);
    float* sarea = (float*)(ws + L.sarea);
    float* sscore = (float*)(ws + L.sscore);
    int* sidx = (int*)(ws + L.sidx);
    const unsigned kb = (unsigned)((L.n2 + 255) / 256);
    SSD_CUDA(launch_pdl(large_keys_kernel, dim3(kb), dim3(256), 0, st, scores, num_boxes, L.n2, keys));
    count_launch();
    int rc2 = sort_keys(keys, L.n2, st);
    if (rc2 != SSD_OK) return rc2;
    if (soft) {
        // the reference's loop runs over the top-k SUBSET in its own order: ascending input index here (as in the
        // batched kernel); the first k sorted keys are re-sorted by index
        SSD_CUDA(launch_pdl(large_index_keys_kernel, dim3(kb), dim3(256), 0, st, keys, L.k, L.n2));
        count_launch();
        rc2 = sort_keys(keys, L.n2, st);
        if (rc2 != SSD_OK) return rc2;
    }
    SSD_CUDA(launch_pdl(large_gather_kernel, dim3((unsigned)((L.k + 255) / 256)), dim3(256), 0, st,
                        (const unsigned long long*)keys, (const float4*)corner_boxes, scores, L.k, soft ? 1 : 0, sbox, sarea,
                        sscore, sidx));
    count_launch();
    if (soft) {
        SSD_CUDA(launch_pdl(large_soft_kernel, dim3(1), dim3(1024), 0, st, (const float4*)sbox, (const float*)sarea, sscore,
                            (const int*)sidx, L.k, soft_threshold, soft_sigma, (long long*)keep_out, count_out));
        count_launch();
        return SSD_OK;
    }
    unsigned long long* mask = (unsigned long long*)(ws + L.mask);
    SSD_CUDA(launch_pdl(large_mask_kernel, dim3(L.words, L.words), dim3(64), 0, st, (const float4*)sbox, (const float*)sarea,
                        L.k, L.words, overlap_threshold, mask));
    count_launch();
    SSD_CUDA(launch_pdl(large_sweep_kernel, dim3(1), dim3(1024), 0, st, (const unsigned long long*)mask, (const int*)sidx, L.k,
                        L.words, (unsigned long long*)(ws + L.removed), (long long*)keep_out, count_out));
    count_launch();
    return SSD_OK;
}

