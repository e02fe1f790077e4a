// Hard-negative mining (detection/sampler.py:12-25) and the naive sampler (sampler.py:9-10).
//
// Two launches, no memset and no global atomics:
//   1. mining_loss_kernel   -- streams logits[B, A, C] once (TMA bulk -> smem ring, class ids as a
//      side array of the same ring), computes the mining criterion
//          loss = -log_softmax(x)[0] = (max + log(sum exp(x - max))) - x0
//      per anchor and folds the class id into one sortable uint32 key per anchor:
//          0           anchor is ignored (class -1): never selected
//          0xFFFFFFFF  anchor is positive: always selected
//          otherwise   ordered_key(loss) of a negative (class 0) anchor
//      This is the HBM-bound kernel: 4*C + 8 bytes read, 4 bytes written per anchor.
//   2. mining_select_kernel -- one CTA per image: a first pass over the image's keys (35 KB for
//      SSD300, L2 resident: the previous launch just wrote them) builds the histogram of the negative
//      losses in shared memory (kLossBins monotone bins, resolution 1/128) and counts the positives;
//      k = min(max(n_pos*ratio, min_neg), n_neg) exactly as the reference derives it (int64 or fp32
//      arithmetic depending on the Python type of `ratio`); a suffix scan of the histogram finds the
//      bin that holds the k-th largest loss; a second pass over the keys selects everything above that bin and collects the handful of keys inside
//      it, which are ranked exactly (key desc, anchor asc).  Loss ties across the cut
//      (implementation-defined in the reference, whose argsort is unstable) go to the lower anchor.
//      A boundary bin too crowded to rank in shared memory (heavily tied losses) falls back to an
//      MSB-first radix select over its members.
#include "rowstream.cuh"

namespace ssd {

constexpr uint32_t kKeyIgnored = 0u;
constexpr uint32_t kKeyPositive = 0xFFFFFFFFu;
constexpr int kLossBins = 4096;
constexpr int kHistStride = kLossBins + 32;      // [kLossBins] = positives, rest padding
constexpr int kBoundaryCap = 2048;               // boundary-bin members ranked in shared memory

__device__ __forceinline__ uint32_t mining_key(float loss, long long cls) {
    if (cls == SSD_NEGATIVE_CLASS) {
        uint32_t k = ordered_key(loss);
        return k == 0u ? 1u : k;                 // keep clear of the two sentinels
    }
    return cls == SSD_IGNORE_CLASS ? kKeyIgnored : kKeyPositive;
}

// Monotone (non-decreasing in the key order, NaN on top) map of a loss to a histogram bin.
__device__ __forceinline__ int loss_bin(float v) {
    if (v != v) return kLossBins - 1;
    const float s = fminf(fmaxf(__fmul_rn(v, 128.f), 0.f), (float)(kLossBins - 3));
    return 1 + (int)s;
}
__device__ __forceinline__ int key_bin(uint32_t key) { return loss_bin(key_to_float(key)); }


template <int Q, int NREG, int CMIN>
__global__ void __launch_bounds__(kStreamThreads)
mining_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ cls, uint32_t* __restrict__ keys,
                   ScoreGrid g) {
    extern __shared__ __align__(128) unsigned char smem[];
    KernelTrace trace_(TR_MINING_LOSS);
    stream_init(smem);
    griddep_wait();
    if (warp_id() == kConsumerWarps) {
        producer_loop(smem, g, logits, reinterpret_cast<const unsigned long long*>(cls), policy_evict_first());
        return;
    }
    const RowLanes<Q> ln;
    const RowShape<Q, NREG, CMIN> shape(ln.sub, g.C);
    const int rows_per_warp = g.tile_rows / kConsumerWarps;
    const uint32_t sb = smem_u32(smem);
    const int wbase = warp_id() * rows_per_warp;
    TileCursor cur;
    cur.start(g);
    for (; cur.valid(g); cur.next(g)) {
        const unsigned r0 = cur.first_row32(g);
        const int rows = cur.rows(g);
        ring_wait_full(sb, cur);
        const float* tile_logits = ring_logits(smem, g, cur, r0);
        const unsigned long long* side = ring_side(smem, g, cur, r0);
        for (int step = 0; step < rows_per_warp; step += RowLanes<Q>::kRowsPerWarpStep) {
            if (wbase + step >= rows) break;            // the rest of a partial tile (warp-uniform)
            const int lr = wbase + step + ln.rl;
            float v[NREG];
            shape.load(v, tile_logits + lr * g.C, ln.sub);
            float m, sum;
            row_max_sum<Q, NREG>(v, m, sum);
            if (lr < rows && ln.sub == 0) {
                // v[0] of lane sub == 0 is column 0
                const float loss = __fsub_rn(__fadd_rn(m, fast_log(sum)), v[0]);
                const long long c = (long long)side[lr];
                keys[r0 + lr] = mining_key(loss, c);
            }
        }
        ring_release(sb, cur);
    }
    griddep_launch_dependents();       // late: see the note at launch_pdl
}

// Criterion given (stage-boundary parity: identical fp32 inputs to the selection), or classes read
// from the class column of a target tensor (fused pipeline): CLS_FLOAT selects float rows of
// `cls_stride` floats, else int64 contiguous.
template <bool CLS_FLOAT>
__global__ void mining_keys_from_loss_kernel(const float* __restrict__ loss, const void* __restrict__ cls, int cls_stride,
                                             uint32_t* __restrict__ keys, int64_t n) {
    KernelTrace trace_(TR_MINING_KEYS);
    griddep_wait();
    griddep_launch_dependents();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long c;
    if (CLS_FLOAT) c = (long long)reinterpret_cast<const float*>(cls)[i * cls_stride];
    else c = reinterpret_cast<const long long*>(cls)[i];
    keys[i] = mining_key(loss[i], c);
}

__global__ void positive_mask_kernel(const long long* __restrict__ cls, uint8_t* __restrict__ mask, int64_t n) {
    griddep_wait();
    griddep_launch_dependents();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const long long c = cls[i];
        mask[i] = (c != SSD_NEGATIVE_CLASS && c != SSD_IGNORE_CLASS) ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// per-image selection
// ---------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;

struct KeySource {
    const uint32_t* keys;
    const void* cls;
    int stride;
    const int32_t* match;        // matcher output per anchor (optional): spares the strided class read for unmatched anchors
    __device__ __forceinline__ uint32_t key(int a) const {
        const uint32_t k = keys[a];
        if (cls == nullptr) return k;
        long long c;
        if (match != nullptr) {
            // target_assigner.py:38-58: an unmatched anchor keeps class 0, an ignored one gets -1; only a MATCHED anchor
            // (a few percent) carries its box's class, which is read from the row
            const int m = match[a];
            c = m == SSD_NOT_MATCHED ? (long long)SSD_NEGATIVE_CLASS
                                     : (m == SSD_IGNORE ? (long long)SSD_IGNORE_CLASS
                                                        : (stride ? (long long)reinterpret_cast<const float*>(cls)[(size_t)a * stride]
                                                                  : reinterpret_cast<const long long*>(cls)[a]));
        } else
        c = stride ? (long long)reinterpret_cast<const float*>(cls)[(size_t)a * stride]
                                   : reinterpret_cast<const long long*>(cls)[a];
        if (c == SSD_NEGATIVE_CLASS) return k == 0u ? 1u : (k == kKeyPositive ? kKeyPositive - 1u : k);
        return c == SSD_IGNORE_CLASS ? kKeyIgnored : kKeyPositive;
    }
};

struct SelShared {
    int warp_tot[32];
    int part[32][4];
    int total[4];
    int scan[32];
    int n_neg, cut_bin, above, in_bin;
    int n_cand;
    unsigned long long cand[kBoundaryCap];
    alignas(16) int hist[kHistStride];           // [kLossBins] = positives
};

// block-wide sum of four per-thread counters; result valid for every thread
__device__ __forceinline__ void block_sum4(SelShared& sh, int (&c)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = __reduce_add_sync(FULL, c[j]);
    if (lane_id() == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sh.part[warp_id()][j] = c[j];
    }
    __syncthreads();
    if (warp_id() == 0) {
        const int nw = blockDim.x >> 5;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int x = lane_id() < nw ? sh.part[lane_id()][j] : 0;
            x = __reduce_add_sync(FULL, x);
            if (lane_id() == 0) sh.total[j] = x;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = sh.total[j];
    __syncthreads();          // sh.total may be overwritten by the next call
}

__global__ void __launch_bounds__(kSelThreads, 2)
mining_select_kernel(const uint32_t* __restrict__ keys, const void* __restrict__ cls, int cls_stride,
                     const int32_t* __restrict__ match, int A, double ratio,
                     int ratio_is_integer, double min_negatives, uint8_t* __restrict__ mask,
                     int32_t* __restrict__ stats, int stage_keys) {
    __shared__ SelShared sh;
    // stage_keys: the image's final keys (class folded in) are kept in shared memory between the two passes, so the
    // second pass and the boundary ranking never go back to L2 (each src.key() is up to three dependent round trips)
    extern __shared__ __align__(16) uint32_t skeys[];
    KernelTrace trace_(TR_MINING_SELECT);
    griddep_wait();
    griddep_launch_dependents();
    const int b = blockIdx.x;
    const uint32_t* gk = keys + (size_t)b * A;
    // cls == nullptr: the keys already carry the class (mining_loss_kernel).  Otherwise they are the RAW
    // criterion (score_pass1_kernel) and the class comes from int64 [B, A] (cls_stride == 0) or from the
    // class column of float target rows `cls_stride` floats apart.
    const KeySource src{gk, cls == nullptr ? nullptr : (cls_stride ? (const void*)(reinterpret_cast<const float*>(cls) + (size_t)b * A * cls_stride)
                                                                  : (const void*)(reinterpret_cast<const long long*>(cls) + (size_t)b * A)),
                        cls_stride, match == nullptr ? nullptr : match + (size_t)b * A};
    uint8_t* gm = mask + (size_t)b * A;
    const int tid = threadIdx.x;
    static_assert(kLossBins == 4 * kSelThreads, "four bins per thread");

    // ---- histogram of the image's negative losses (+ the positives count) in shared memory ----
    reinterpret_cast<int4*>(sh.hist)[tid] = make_int4(0, 0, 0, 0);
    if (tid < kHistStride - kLossBins) sh.hist[kLossBins + tid] = 0;
    __syncthreads();
    {
        int pos = 0;
        for (int a0 = tid; a0 < A; a0 += 8 * kSelThreads) {
            uint32_t kk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int a = a0 + u * kSelThreads;
                kk[u] = a < A ? src.key(a) : kKeyIgnored;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t key = kk[u];
                if (stage_keys && a0 + u * kSelThreads < A) skeys[a0 + u * kSelThreads] = key;
                if (key == kKeyPositive) ++pos;
                else if (key != kKeyIgnored) atomicAdd(&sh.hist[key_bin(key)], 1);
            }
        }
        pos = __reduce_add_sync(FULL, pos);
        if (lane_id() == 0 && pos) atomicAdd(&sh.hist[kLossBins], pos);
    }
    __syncthreads();
    // ---- suffix scan of the histogram: thread t owns bins 4t..4t+3, higher bins = larger losses ----
    const int4 h4 = reinterpret_cast<const int4*>(sh.hist)[tid];
    const int n_pos = sh.hist[kLossBins];
    const int own = h4.x + h4.y + h4.z + h4.w;
    int incl = own;                                   // inclusive suffix sum inside the warp (towards higher lanes)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(FULL, incl, o);
        if (lane_id() + o < 32) incl += t;
    }
    if (lane_id() == 0) sh.warp_tot[warp_id()] = incl;
    __syncthreads();
    int higher_warps = 0;                             // keys in bins of the warps above this one
    int n_neg = 0;
    {
        const int wt = sh.warp_tot[lane_id()];
        n_neg = __reduce_add_sync(FULL, wt);
        higher_warps = __reduce_add_sync(FULL, lane_id() > warp_id() ? wt : 0);
    }
    const int above_thread = higher_warps + incl - own;      // keys in bins above 4t+3

    // k = min(clamp(n_pos * ratio, min=min_neg), n_neg)          detection/sampler.py:20
    long long k;
    if (ratio_is_integer) {
        long long want = (long long)n_pos * (long long)ratio;
        const long long floor_ = (long long)min_negatives;
        if (want < floor_) want = floor_;
        k = want < (long long)n_neg ? want : (long long)n_neg;
    } else {
        float want = __fmul_rn((float)n_pos, (float)ratio);       // int64 tensor * Python float -> fp32
        const float floor_ = (float)min_negatives;
        if (want < floor_) want = floor_;
        const float cap = (float)n_neg;
        if (cap < want) want = cap;
        k = (long long)ceilf(want);                               // rank < want  <=>  rank < ceil(want)
        if (k > n_neg) k = n_neg;
    }
    if (k < 0) k = 0;

    // the bin holding the k-th largest negative: above(bin) < k <= above(bin) + hist[bin]
    const bool partial = k > 0 && k < n_neg;
    if (tid == 0) { sh.cut_bin = partial ? -1 : (k > 0 ? 0 : kLossBins); sh.above = 0; sh.in_bin = 0; sh.n_cand = 0; }
    __syncthreads();
    if (partial && above_thread < k && above_thread + own >= k) {
        int above = above_thread;
        const int hb[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int j = 3; j >= 0; --j) {
            if (above < k && above + hb[j] >= k) { sh.cut_bin = 4 * tid + j; sh.above = above; sh.in_bin = hb[j]; }
            above += hb[j];
        }
    }
    __syncthreads();
    const int cut_bin = sh.cut_bin;                   // select bins > cut_bin, `need` keys of cut_bin itself
    const int in_bin = sh.in_bin;
    const int need = partial ? (int)k - sh.above : 0;
    const bool rank_bin = partial && need < in_bin;   // otherwise the whole boundary bin is taken
    const bool in_smem = in_bin <= kBoundaryCap;

    auto key_at = [&](int a) -> uint32_t { return stage_keys ? skeys[a] : src.key(a); };
    // ---- one pass over the keys (four loads in flight per thread) ----
    for (int a0 = tid; a0 < A; a0 += 4 * kSelThreads) {
        uint32_t kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = a0 + u * kSelThreads;
            kk[u] = a < A ? key_at(a) : kKeyIgnored;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = a0 + u * kSelThreads;
            if (a >= A) break;
            const uint32_t key = kk[u];
            uint8_t sel;
            if (key == kKeyPositive) sel = 1;
            else if (key == kKeyIgnored) sel = 0;
            else {
                const int bin = key_bin(key);
                sel = bin > cut_bin || (bin == cut_bin && !rank_bin);
                if (rank_bin && bin == cut_bin && in_smem) {
                    const int slot = atomicAdd(&sh.n_cand, 1);
                    sh.cand[slot] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)a);
                }
            }
            gm[a] = sel;
        }
    }
    int num_ties = 0;
    if (rank_bin) {
        __syncthreads();
        if (in_smem) {
            // exact rank inside the boundary bin: larger key first, then lower anchor
            const int n = sh.n_cand;
            for (int i = tid; i < n; i += kSelThreads) {
                const unsigned long long me = sh.cand[i];
                int rank = 0;
                for (int j = 0; j < n; ++j) rank += sh.cand[j] > me;
                if (rank < need) gm[0xFFFFFFFFu - (uint32_t)(me & 0xFFFFFFFFull)] = 1;
            }
        } else {
            // crowded boundary bin (heavily tied losses): MSB-first radix select over its members
            uint32_t prefix = 0u;
            int rem = need;
            num_ties = in_bin;
            for (int shift = 30; shift >= 0; shift -= 2) {
                int d[4] = {0, 0, 0, 0};
                const uint32_t pre_hi = shift + 2 >= 32 ? 0u : prefix >> (shift + 2);
                for (int a = tid; a < A; a += kSelThreads) {
                    const uint32_t key = key_at(a);
                    if (key != kKeyPositive && key != kKeyIgnored && key_bin(key) == cut_bin) {
                        const uint32_t khi = shift + 2 >= 32 ? 0u : key >> (shift + 2);
                        if (khi == pre_hi) d[(key >> shift) & 3u]++;
                    }
                }
                block_sum4(sh, d);
                int digit = 3;
                for (; digit > 0; --digit) {
                    if (rem <= d[digit]) break;
                    rem -= d[digit];
                }
                prefix |= (uint32_t)digit << shift;
                num_ties = d[digit];
            }
            const uint32_t thr_key = prefix;          // members > thr_key, plus the first `rem` members == thr_key
            int base = 0;
            const int rounds = (A + kSelThreads - 1) / kSelThreads;
            for (int i = 0; i < rounds; ++i) {
                const int a = i * kSelThreads + tid;
                uint32_t key = kKeyIgnored;
                if (a < A) key = key_at(a);
                const bool member = a < A && key != kKeyPositive && key != kKeyIgnored && key_bin(key) == cut_bin;
                const bool tie = member && key == thr_key;
                const unsigned bal = __ballot_sync(FULL, tie);
                if (lane_id() == 0) sh.scan[warp_id()] = __popc(bal);
                __syncthreads();
                int before = 0, round_total = 0;
                const int nw = blockDim.x >> 5;
                for (int w = 0; w < nw; ++w) {
                    const int x = sh.scan[w];
                    if (w < warp_id()) before += x;
                    round_total += x;
                }
                const int rank = base + before + __popc(bal & ((1u << lane_id()) - 1u));
                if (member && (key > thr_key || (tie && rank < rem))) gm[a] = 1;
                base += round_total;
                __syncthreads();
            }
        }
    }
    if (stats != nullptr && tid == 0) {
        stats[b * 4 + 0] = n_pos;
        stats[b * 4 + 1] = n_neg;
        stats[b * 4 + 2] = (int)k;
        stats[b * 4 + 3] = rank_bin ? (in_smem ? in_bin : num_ties) : 0;
    }
}

// keys staged in shared memory while an image's keys fit next to the static arrays (A <= 40960), else re-read from L2
constexpr size_t kSelStageMax = 160 * 1024;
static size_t select_smem_bytes(int num_anchors) {
    const size_t need = round_up((size_t)num_anchors * sizeof(uint32_t), 16);
    return need <= kSelStageMax ? need : 0;
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_positive_mask(const int64_t* target_classes, int64_t count, uint8_t* mask_out, void* stream) {
    SSD_REQUIRE(count >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_positive_mask: negative count");
    if (count == 0) return SSD_OK;
    SSD_REQUIRE(target_classes && mask_out, SSD_ERR_INVALID_ARGUMENT, "ssd_positive_mask: null pointer");
    const int threads = 256;
    const int64_t blocks = (count + threads - 1) / threads;
    SSD_CUDA(launch_pdl(positive_mask_kernel, dim3((unsigned)blocks), dim3(threads), 0, (cudaStream_t)stream,
                        (const long long*)target_classes, mask_out, count));
    count_launch();
    return SSD_OK;
}

static size_t keys_bytes(int batch, int num_anchors) { return round_up((size_t)batch * num_anchors * sizeof(uint32_t), 256); }

// keys + histograms of the whole batch from the logits
static int launch_mining_keys(const float* logits, const int64_t* target_classes, int batch, int num_anchors,
                              int num_cols, uint32_t* keys, cudaStream_t st) {
    SSD_REQUIRE(num_cols >= 1 && num_cols <= kMaxScoreCols, SSD_ERR_UNSUPPORTED,
                "mining: num_cols %d outside 1..%d", num_cols, kMaxScoreCols);
    SSD_REQUIRE(aligned(logits, 16), SSD_ERR_MISALIGNED, "mining: logits not 16-byte aligned");
    SSD_REQUIRE(aligned(target_classes, 16), SSD_ERR_MISALIGNED, "mining: target_classes not 16-byte aligned");
    SSD_REQUIRE((long long)batch * num_anchors < 0x7FFFFFFFll, SSD_ERR_UNSUPPORTED,
                "mining: %d x %d rows exceed the 32-bit row index of the streaming kernel", batch, num_anchors);
    ScoreGrid g;
    plan_tiles(g, batch, num_anchors, num_cols, true);
    const int grid = stream_grid(g);
    const size_t smem = stream_smem_bytes(g);
#define SSD_LAUNCH_MINING(QQ, NN, CM)                                                                                   \
    do {                                                                                                             \
        auto kern = mining_loss_kernel<QQ, NN, CM>;                                                                    \
        SSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
        LaunchTimer lt_("mining_loss", st);                                                            \
        SSD_CUDA(launch_pdl(kern, dim3(grid), dim3(kStreamThreads), smem, st, logits,                               \
                            (const long long*)target_classes, keys, g));                                      \
    } while (0)
    SSD_DISPATCH_ROW_SHAPE(num_cols, SSD_LAUNCH_MINING);
#undef SSD_LAUNCH_MINING
    SSD_CUDA(cudaGetLastError());
    count_launch();
    return SSD_OK;
}

extern "C" size_t ssd_hard_negative_workspace_bytes(int batch, int num_anchors) {
    if (batch <= 0 || num_anchors <= 0) return 256;
    return keys_bytes(batch, num_anchors);
}

extern "C" int ssd_mining_keys(const float* logits, const int64_t* target_classes, int batch, int num_anchors,
                               int num_cols, uint32_t* keys_out, void* workspace, size_t workspace_bytes,
                               void* stream) {
    SSD_REQUIRE(batch >= 0 && num_anchors >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_mining_keys: negative shape");
    if (batch == 0 || num_anchors == 0) return SSD_OK;
    SSD_REQUIRE(logits && target_classes && keys_out && workspace, SSD_ERR_INVALID_ARGUMENT,
                "ssd_mining_keys: null pointer");
    SSD_REQUIRE(workspace_bytes >= ssd_hard_negative_workspace_bytes(batch, num_anchors), SSD_ERR_WORKSPACE,
                "ssd_mining_keys: workspace too small");
    return launch_mining_keys(logits, target_classes, batch, num_anchors, num_cols, keys_out, (cudaStream_t)stream);
}

extern "C" int ssd_hard_negative_mask(const float* logits, const int64_t* target_classes, const float* loss_override,
                                      int batch, int num_anchors, int num_cols, double ratio, int ratio_is_integer,
                                      double min_negatives, uint8_t* mask_out, int32_t* stats_out, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    SSD_REQUIRE(batch >= 0 && num_anchors >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_hard_negative_mask: negative shape");
    if (batch == 0 || num_anchors == 0) return SSD_OK;
    SSD_REQUIRE(target_classes && mask_out && workspace, SSD_ERR_INVALID_ARGUMENT,
                "ssd_hard_negative_mask: null pointer");
    SSD_REQUIRE(logits || loss_override, SSD_ERR_INVALID_ARGUMENT, "ssd_hard_negative_mask: no logits and no loss");
    SSD_REQUIRE(workspace_bytes >= ssd_hard_negative_workspace_bytes(batch, num_anchors), SSD_ERR_WORKSPACE,
                "ssd_hard_negative_mask: workspace too small");
    SSD_REQUIRE(aligned(workspace, 256), SSD_ERR_MISALIGNED, "ssd_hard_negative_mask: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* keys = (uint32_t*)workspace;
    const int64_t total_rows = (int64_t)batch * num_anchors;

    if (loss_override != nullptr) {
        const int threads = 256;
        LaunchTimer lt_("keys_from_loss", st);
        SSD_CUDA(launch_pdl(mining_keys_from_loss_kernel<false>, dim3((unsigned)((total_rows + threads - 1) / threads)),
                            dim3(threads), 0, st, loss_override, (const void*)target_classes, 1, keys, total_rows));
        count_launch();
    } else {
        const int rc = launch_mining_keys(logits, target_classes, batch, num_anchors, num_cols, keys, st);
        if (rc != SSD_OK) return rc;
    }

    LaunchTimer lt_("mining_select", st);
    const size_t sel_smem = select_smem_bytes(num_anchors);
    SSD_CUDA(cudaFuncSetAttribute(mining_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelStageMax));
    SSD_CUDA(launch_pdl(mining_select_kernel, dim3(batch), dim3(kSelThreads), sel_smem, st, (const uint32_t*)keys,
                        (const void*)nullptr, 0, (const int32_t*)nullptr, num_anchors, ratio, ratio_is_integer, min_negatives, mask_out, stats_out,
                        sel_smem ? 1 : 0));
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_hard_negative_mask_from_keys(const uint32_t* loss_keys, const void* classes, int class_stride,
                                                const int32_t* match, int batch, int num_anchors, double ratio, int ratio_is_integer,
                                                double min_negatives, uint8_t* mask_out, int32_t* stats_out,
                                                void* stream) {
    SSD_REQUIRE(batch >= 0 && num_anchors >= 0 && class_stride >= 0, SSD_ERR_INVALID_ARGUMENT,
                "ssd_hard_negative_mask_from_keys: negative shape");
    if (batch == 0 || num_anchors == 0) return SSD_OK;
    SSD_REQUIRE(loss_keys && classes && mask_out, SSD_ERR_INVALID_ARGUMENT, "ssd_hard_negative_mask_from_keys: null pointer");
    LaunchTimer lt_("mining_select", (cudaStream_t)stream);
    const size_t sel_smem = select_smem_bytes(num_anchors);
    SSD_CUDA(cudaFuncSetAttribute(mining_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSelStageMax));
    SSD_CUDA(launch_pdl(mining_select_kernel, dim3(batch), dim3(kSelThreads), sel_smem, (cudaStream_t)stream, loss_keys, classes,
                        class_stride, match, num_anchors, ratio, ratio_is_integer, min_negatives, mask_out, stats_out,
                        sel_smem ? 1 : 0));
    count_launch();
    return SSD_OK;
}

SSD_DEFINE_TRACE_SETTER(set_trace_mining)
