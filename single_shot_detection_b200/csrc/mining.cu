// Hard-negative mining (detection/sampler.py:12-25) and the naive sampler (sampler.py:9-10).
//
// Two launches:
//   1. mining_loss_kernel   -- streams logits[B*A, C] once (TMA bulk -> smem ring), computes the
//      mining criterion  loss = -log_softmax(x)[0] = -((x0 - max) - log(sum exp(x - max)))  per
//      anchor and folds the class id into one sortable uint32 key per anchor:
//          0           anchor is ignored (class -1): never selected
//          0xFFFFFFFF  anchor is positive: always selected
//          otherwise   ordered_key(loss) of a negative (class 0) anchor
//      This is the HBM-bound kernel: 4*C + 8 bytes read, 4 bytes written per anchor.
//   2. mining_select_kernel -- one CTA per image: counts positives / negatives, derives
//      k = min(max(n_pos*ratio, min_neg), n_neg) exactly as the reference does (int64 or fp32
//      arithmetic depending on the Python type of `ratio`), finds the k-th largest negative key
//      with an MSB-first radix select on register-resident keys, and writes the bool mask.
//      Loss ties across the cut (implementation-defined in the reference, whose argsort is
//      unstable) go to the lower anchor index.
#include "rowstream.cuh"

namespace ssd {

constexpr uint32_t kKeyIgnored = 0u;
constexpr uint32_t kKeyPositive = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t mining_key(float loss, long long cls) {
    if (cls == SSD_NEGATIVE_CLASS) {
        uint32_t k = ordered_key(loss);
        return k == 0u ? 1u : k;                 // keep clear of the two sentinels
    }
    return cls == SSD_IGNORE_CLASS ? kKeyIgnored : kKeyPositive;
}

template <int Q, int NREG>
__global__ void __launch_bounds__(kStreamThreads)
mining_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ cls, uint32_t* __restrict__ keys,
                   int64_t total_rows, int C, int tile_rows, int stage_floats, int num_tiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    RowStream<kStreamStages> rs;
    stream_setup(rs, smem, stage_floats);
    const RowLanes<Q> ln;
    const int warps = blockDim.x >> 5;
    const int rows_per_step = warps * RowLanes<Q>::kRowsPerWarpStep;
    const int64_t total_floats = total_rows * C;
    const uint64_t policy = policy_evict_first();

    // prologue: fill the ring
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStreamStages; ++s) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < num_tiles) {
                const int64_t r0 = (int64_t)t * tile_rows;
                const int rows = (int)min((int64_t)tile_rows, total_rows - r0);
                rs.issue(s, logits, r0, rows, C, total_floats, policy);
            }
        }
    }
    int k = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++k) {
        const int s = k % kStreamStages;
        const uint32_t parity = (k / kStreamStages) & 1;
        const int64_t r0 = (int64_t)t * tile_rows;
        const int rows = (int)min((int64_t)tile_rows, total_rows - r0);
        mbar_wait(&rs.full[s], parity);
        const float* tile = rs.buf[s] + RowStream<kStreamStages>::head_of(r0, C);

        for (int base = 0; base < rows; base += rows_per_step) {
            const int lr = base + warp_id() * RowLanes<Q>::kRowsPerWarpStep + ln.rl;
            const bool valid = lr < rows;
            float v[NREG];
            load_row_slice<Q, NREG>(v, tile + (size_t)lr * C, ln.sub, C, valid, -INFINITY);
            float m = v[0];
#pragma unroll
            for (int i = 1; i < NREG; ++i) m = fmaxf(m, v[i]);
            m = group_max<Q>(m);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < NREG; ++i) {
                const int col = ln.sub + i * Q;
                if (col < C) sum = __fadd_rn(sum, expf(__fsub_rn(v[i], m)));
            }
            sum = group_sum<Q>(sum);
            const float x0 = __shfl_sync(FULL, v[0], lane_id() - ln.sub);
            if (valid && ln.sub == 0) {
                const float loss = -__fsub_rn(__fsub_rn(x0, m), logf(sum));
                const int64_t row = r0 + lr;
                keys[row] = mining_key(loss, cls[row]);
            }
        }
        __syncthreads();     // every warp is done with stage s
        if (threadIdx.x == 0) {
            const int tn = t + kStreamStages * gridDim.x;
            if (tn < num_tiles) {
                const int64_t rn = (int64_t)tn * tile_rows;
                const int rows_n = (int)min((int64_t)tile_rows, total_rows - rn);
                rs.issue(s, logits, rn, rows_n, C, total_floats, policy);
            }
        }
    }
}

// stage-boundary variant: the criterion is given (identical fp32 inputs to the selection)
__global__ void mining_keys_from_loss_kernel(const float* __restrict__ loss, const long long* __restrict__ cls,
                                             uint32_t* __restrict__ keys, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = mining_key(loss[i], cls[i]);
}

__global__ void positive_mask_kernel(const long long* __restrict__ cls, uint8_t* __restrict__ mask, int64_t n) {
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

struct SelShared {
    int part[32][4];
    int total[4];
    uint32_t umin[32], umax[32];
    uint32_t gmin, gmax;
    int scan[32];
};

// block-wide sum of four per-thread counters; result valid in sh.total for every thread
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

// KPT > 0: keys live in registers (A <= KPT*1024).  KPT == 0: keys are re-read every pass from
// `src`, which is either a shared-memory copy (A*4 bytes fit) or the global array (L2).
template <int KPT, typename F>
__device__ __forceinline__ void for_each_key(const uint32_t* src, const uint32_t* rk, int A, F f) {
    if constexpr (KPT > 0) {
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int a = i * kSelThreads + threadIdx.x;
            if (a < A) f(a, rk[i]);
        }
    } else {
        for (int a = threadIdx.x; a < A; a += kSelThreads) f(a, src[a]);
    }
}

template <int KPT>
__global__ void __launch_bounds__(kSelThreads)
mining_select_kernel(const uint32_t* __restrict__ keys, int A, int stage_in_smem, double ratio,
                     int ratio_is_integer, double min_negatives, uint8_t* __restrict__ mask,
                     int32_t* __restrict__ stats) {
    __shared__ SelShared sh;
    extern __shared__ __align__(16) uint32_t skeys[];
    const int b = blockIdx.x;
    const uint32_t* gkeys = keys + (size_t)b * A;
    const uint32_t* gk = gkeys;
    uint8_t* gm = mask + (size_t)b * A;
    if (KPT == 0 && stage_in_smem) {
        for (int a = threadIdx.x; a < A; a += kSelThreads) skeys[a] = gkeys[a];
        __syncthreads();
        gk = skeys;
    }

    uint32_t rk[KPT > 0 ? KPT : 1];
    if constexpr (KPT > 0) {
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int a = i * kSelThreads + threadIdx.x;
            rk[i] = a < A ? gk[a] : kKeyIgnored;
        }
    }

    // counts + range of the negative keys
    int c[4] = {0, 0, 0, 0};
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    for_each_key<KPT>(gk, rk, A, [&](int, uint32_t key) {
        if (key == kKeyPositive) c[0]++;
        else if (key != kKeyIgnored) { c[1]++; lo = min(lo, key); hi = max(hi, key); }
    });
    lo = __reduce_min_sync(FULL, lo);
    hi = __reduce_max_sync(FULL, hi);
    if (lane_id() == 0) { sh.umin[warp_id()] = lo; sh.umax[warp_id()] = hi; }
    block_sum4(sh, c);                           // includes the barriers that publish umin/umax
    if (warp_id() == 0) {
        const int nw = blockDim.x >> 5;
        uint32_t l = lane_id() < nw ? sh.umin[lane_id()] : 0xFFFFFFFFu;
        uint32_t h = lane_id() < nw ? sh.umax[lane_id()] : 0u;
        l = __reduce_min_sync(FULL, l);
        h = __reduce_max_sync(FULL, h);
        if (lane_id() == 0) { sh.gmin = l; sh.gmax = h; }
    }
    __syncthreads();
    const int n_pos = c[0], n_neg = c[1];
    const uint32_t kmin = sh.gmin, kmax = sh.gmax;

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

    uint32_t thr_key = 0xFFFFFFFFu;      // select keys > thr_key, plus `need_ties` keys == thr_key
    int need_ties = 0, num_ties = 0;
    if (k >= n_neg) {                    // every negative is taken (also covers n_neg == 0)
        thr_key = 0u;
    } else if (k > 0) {
        // MSB-first radix select, 2 bits per pass, skipping the prefix all negative keys share
        const uint32_t diff = kmin ^ kmax;
        uint32_t prefix;
        int shift;
        if (diff == 0u) {
            prefix = kmax;
            shift = -2;
        } else {
            const int hb = 31 - __clz(diff);
            shift = hb & ~1;
            prefix = shift + 2 >= 32 ? 0u : (kmax >> (shift + 2)) << (shift + 2);
        }
        long long rem = k;
        num_ties = n_neg;
        for (; shift >= 0; shift -= 2) {
            int d[4] = {0, 0, 0, 0};
            const uint32_t pre_hi = shift + 2 >= 32 ? 0u : prefix >> (shift + 2);
            for_each_key<KPT>(gk, rk, A, [&](int, uint32_t key) {
                if (key != kKeyPositive && key != kKeyIgnored) {
                    const uint32_t khi = shift + 2 >= 32 ? 0u : key >> (shift + 2);
                    if (khi == pre_hi) d[(key >> shift) & 3u]++;
                }
            });
            block_sum4(sh, d);
            int digit = 3;
            for (; digit > 0; --digit) {
                if (rem <= d[digit]) break;
                rem -= d[digit];
            }
            prefix |= (uint32_t)digit << shift;
            num_ties = d[digit];
        }
        thr_key = prefix;
        need_ties = (int)rem;
    }

    // mask.  Ties at the cut: lowest anchor index first.
    const bool rank_ties = (k > 0 && k < n_neg && need_ties < num_ties);
    if (!rank_ties) {
        for_each_key<KPT>(gk, rk, A, [&](int a, uint32_t key) {
            const bool neg = key != kKeyPositive && key != kKeyIgnored;
            gm[a] = (key == kKeyPositive || (neg && k > 0 && key >= thr_key)) ? 1 : 0;
        });
    } else {
        // rare path: ordered prefix count of the tied keys, anchor order = (round, thread)
        int base = 0;
        const int rounds = (A + kSelThreads - 1) / kSelThreads;
        for (int i = 0; i < rounds; ++i) {
            const int a = i * kSelThreads + threadIdx.x;
            uint32_t key = kKeyIgnored;
            if (a < A) key = gkeys[a];
            const bool tie = a < A && key == thr_key;
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
            if (a < A) {
                const bool neg = key != kKeyPositive && key != kKeyIgnored;
                gm[a] = (key == kKeyPositive || (neg && key > thr_key) || (tie && rank < need_ties)) ? 1 : 0;
            }
            base += round_total;
            __syncthreads();
        }
    }
    if (stats != nullptr && threadIdx.x == 0) {
        stats[b * 4 + 0] = n_pos;
        stats[b * 4 + 1] = n_neg;
        stats[b * 4 + 2] = (int)k;
        stats[b * 4 + 3] = rank_ties ? num_ties : 0;
    }
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_positive_mask(const int64_t* target_classes, int64_t count, uint8_t* mask_out, void* stream) {
    SSD_REQUIRE(count >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_positive_mask: negative count");
    if (count == 0) return SSD_OK;
    SSD_REQUIRE(target_classes && mask_out, SSD_ERR_INVALID_ARGUMENT, "ssd_positive_mask: null pointer");
    const int threads = 256;
    const int64_t blocks = (count + threads - 1) / threads;
    positive_mask_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        (const long long*)target_classes, mask_out, count);
    SSD_CUDA(cudaGetLastError());
    count_launch();
    return SSD_OK;
}

static int launch_mining_keys(const float* logits, const int64_t* target_classes, int64_t total_rows, int num_cols,
                              uint32_t* keys, cudaStream_t st) {
    SSD_REQUIRE(num_cols >= 1 && num_cols <= kMaxScoreCols, SSD_ERR_UNSUPPORTED,
                "mining: num_cols %d outside 1..%d", num_cols, kMaxScoreCols);
    SSD_REQUIRE(aligned(logits, 16), SSD_ERR_MISALIGNED, "mining: logits not 16-byte aligned");
    const StreamShape shp = make_stream_shape(num_cols, 32);
    const int num_tiles = (int)((total_rows + shp.tile_rows - 1) / shp.tile_rows);
    int grid = 2 * sm_count();
    if (grid > num_tiles) grid = num_tiles;
#define SSD_LAUNCH_MINING(QQ, NN)                                                                                   \
    do {                                                                                                             \
        auto kern = mining_loss_kernel<QQ, NN>;                                                                      \
        SSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shp.smem_bytes));      \
        kern<<<grid, kStreamThreads, shp.smem_bytes, st>>>(logits, (const long long*)target_classes, keys,          \
                                                            total_rows, num_cols, shp.tile_rows, shp.stage_floats,   \
                                                            num_tiles);                                              \
    } while (0)
    SSD_DISPATCH_ROW_SHAPE(num_cols, SSD_LAUNCH_MINING);
#undef SSD_LAUNCH_MINING
    SSD_CUDA(cudaGetLastError());
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_mining_keys(const float* logits, const int64_t* target_classes, int batch, int num_anchors,
                               int num_cols, uint32_t* keys_out, void* stream) {
    SSD_REQUIRE(batch >= 0 && num_anchors >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_mining_keys: negative shape");
    if (batch == 0 || num_anchors == 0) return SSD_OK;
    SSD_REQUIRE(logits && target_classes && keys_out, SSD_ERR_INVALID_ARGUMENT, "ssd_mining_keys: null pointer");
    return launch_mining_keys(logits, target_classes, (int64_t)batch * num_anchors, num_cols, keys_out,
                              (cudaStream_t)stream);
}

extern "C" size_t ssd_hard_negative_workspace_bytes(int batch, int num_anchors) {
    if (batch <= 0 || num_anchors <= 0) return 256;
    return round_up((size_t)batch * num_anchors * sizeof(uint32_t), 256);
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
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* keys = (uint32_t*)workspace;
    const int64_t total_rows = (int64_t)batch * num_anchors;

    if (loss_override != nullptr) {
        const int threads = 256;
        mining_keys_from_loss_kernel<<<(unsigned)((total_rows + threads - 1) / threads), threads, 0, st>>>(
            loss_override, (const long long*)target_classes, keys, total_rows);
        SSD_CUDA(cudaGetLastError());
    count_launch();
    } else {
        const int rc = launch_mining_keys(logits, target_classes, total_rows, num_cols, keys, st);
        if (rc != SSD_OK) return rc;
    }

    if (num_anchors <= 12 * kSelThreads) {
        mining_select_kernel<12><<<batch, kSelThreads, 0, st>>>(keys, num_anchors, 0, ratio, ratio_is_integer,
                                                                 min_negatives, mask_out, stats_out);
    } else {
        const size_t key_bytes = (size_t)num_anchors * sizeof(uint32_t);
        const int in_smem = key_bytes <= 220 * 1024 ? 1 : 0;
        const size_t dyn = in_smem ? key_bytes : 0;
        auto kern = mining_select_kernel<0>;
        SSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        kern<<<batch, kSelThreads, dyn, st>>>(keys, num_anchors, in_smem, ratio, ratio_is_integer, min_negatives,
                                              mask_out, stats_out);
    }
    SSD_CUDA(cudaGetLastError());
    count_launch();
    return SSD_OK;
}
