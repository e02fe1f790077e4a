// Target assignment: detection/target_assigner.py:22-63 with detection/matcher.py:33-56 and
// bf/utils/box_utils.py:16-23,38-101 fused into one launch for the whole batch.
//
// Grid = (anchor blocks, images): 1024 anchors per CTA, four per thread (strided by the CTA width, so
// every access stays coalesced); every CTA stages its image's ground-truth boxes in shared memory and
// walks them:
//   * IoU in the reference's operation order, every fp32 op rounded separately (no FMA):
//       inter = clamp(min(x2)-max(x1),0) * clamp(min(y2)-max(y1),0)
//       iou   = inter / ((area_gt + area_anchor) - inter)                 box_utils.py:93-101
//   * per-anchor running max over GT (first maximum wins = lowest GT index, NaN sticks),
//   * per-GT argmax over anchors: the thread's best of its four, warp REDUX max on the ordered IoU
//     key, REDUX min on the anchor among its holders, a shared-memory atomicMax on a 64-bit
//     (key, ~anchor) word per warp that beats the CTA's best, then ONE global atomicMax per (CTA, GT)
//     that saw an overlap at all;
//   * thresholds in fp32 (matcher.py:49-50), target rows written as coalesced float2 stores
//     (target_assigner.py:38-58);
//   * the forced match of every GT to its best anchor (matcher.py:53-54) needs the argmax over ALL
//     anchors of the image: the last CTA of an image to finish (fence + ticket) reads the merged
//     table and rewrites those <= G rows, colliding GTs resolved towards the highest GT index
//     (sequential index_put semantics).  No second launch, no cluster, any grid size.
#include "common.cuh"

namespace ssd {

constexpr int kAssignThreads = 256;
// Anchors per thread: 1024 anchors per CTA keep the grid at ~1/4 of a wave of thread slots (SSD300 b32:
// 288 CTAs), so the launch ramp is short, the kernel leaves room for the post-processor's first pass that
// runs beside it in the step graph, GT staging / per-GT atomics / tickets are amortised over 4x the work,
// and the four independent IoU chains per thread hide the shared-memory and divide latencies.
constexpr int kAssignPerThread = 4;
constexpr int kAssignTile = kAssignThreads * kAssignPerThread;
constexpr int kMaxGtPerImage = 4096;

__device__ __forceinline__ float clamped_area(float x1, float y1, float x2, float y2) {
    // (x2 - x1).clamp_(0) * (y2 - y1).clamp_(0)                      box_utils.py:46
    const float w = fmaxf(fsub(x2, x1), 0.f);
    const float h = fmaxf(fsub(y2, y1), 0.f);
    return fmul(w, h);
}

__device__ __forceinline__ float iou_exact(float4 g, float garea, float4 a, float aarea) {
    const float ix1 = fmaxf(g.x, a.x), iy1 = fmaxf(g.y, a.y);
    const float ix2 = fminf(g.z, a.z), iy2 = fminf(g.w, a.w);
    const float iw = fmaxf(fsub(ix2, ix1), 0.f);
    const float ih = fmaxf(fsub(iy2, iy1), 0.f);
    const float inter = fmul(iw, ih);
    const float uni = fsub(fadd(garea, aarea), inter);
    // 0 / positive is +0 exactly; everything else takes the IEEE divide (0/0 -> NaN as torch)
    return (inter == 0.f && uni > 0.f) ? 0.f : fdiv(inter, uni);
}

__device__ __forceinline__ float4 corners_of(float4 c) {
    // box_utils.py:23 -- w/2 is exact
    const float hw = fmul(c.z, 0.5f), hh = fmul(c.w, 0.5f);
    return make_float4(fsub(c.x, hw), fsub(c.y, hh), fadd(c.x, hw), fadd(c.y, hh));
}

// Per-image scratch (zeroed by the caller's memset): [0] CTAs done, [1..3] positives / ignored /
// positives with a NaN box, summed over the CTAs.
constexpr int kAssignCounters = 4;

__device__ __forceinline__ void row_class_counts(float cls, float4 bx, int& pos, int& ign, int& nan) {
    const bool p = cls != (float)SSD_NEGATIVE_CLASS && cls != (float)SSD_IGNORE_CLASS;
    pos = p;
    ign = cls == (float)SSD_IGNORE_CLASS;
    nan = p && (bx.x != bx.x || bx.y != bx.y || bx.z != bx.z || bx.w != bx.w);
}

// coding.on: the box columns of the target rows are written ALREADY passed through the loss route's
// to_centroids(inplace) + encode_box(inplace) (multibox_loss.py:81-82) -- the only consumer of the target
// does exactly that to every row right away, so the two box passes over [B, A, 4] disappear.
struct BoxCoding {
    int on;
    float xy, wh, eps;
};
__device__ __forceinline__ float4 coded_box(float4 corners, float4 prior, const BoxCoding& c) {
    const Box e = encode_inplace(to_centroids_inplace(Box{corners.x, corners.y, corners.z, corners.w}), prior, c.xy, c.wh, c.eps);
    return make_float4(e.a, e.b, e.c, e.d);
}

__global__ void __launch_bounds__(kAssignThreads, 5)
assign_targets_kernel(const float4* __restrict__ anchors, const float* __restrict__ gt_rows, int gt_cols,
                      const int32_t* __restrict__ gt_offsets, int A, int max_gt, float matched_thr,
                      float unmatched_thr, int force_match, float* __restrict__ target, int32_t* __restrict__ match_out,
                      int32_t* __restrict__ stats, int* __restrict__ counters, unsigned long long* __restrict__ gbest,
                      BoxCoding coding) {
    KernelTrace trace_(TR_ASSIGN);
    griddep_wait();
    const int img = blockIdx.y;
    const int g0 = gt_offsets[img];
    const int G_raw = gt_offsets[img + 1] - g0;
    // shared memory and the per-GT table are sized for max_gt boxes: an image that (against the contract) has more is
    // cut there -- never an out-of-bounds access -- and shows as stats[3] > max_gt
    const int G = G_raw < max_gt ? G_raw : max_gt;
    const int a_begin = blockIdx.x * kAssignTile;
    const int n_local = min(kAssignTile, A - a_begin);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Gcap = G > 0 ? G : 1;
    float4* gbox = reinterpret_cast<float4*>(smem_raw);
    unsigned long long* best = reinterpret_cast<unsigned long long*>(gbox + Gcap);      // later: int anchor[G]
    float2* gcs = reinterpret_cast<float2*>(best + Gcap);
    float* garea = reinterpret_cast<float*>(gcs + Gcap);
    int* gslow = reinterpret_cast<int*>(garea + Gcap);            // box with a non-finite coordinate: no quick reject
    __shared__ int s_last;

    // the anchors do not depend on the ground truth: their loads go out before the staging barrier
    float4 ab[kAssignPerThread];
    float aarea[kAssignPerThread];
    bool valid[kAssignPerThread];
    bool aslow[kAssignPerThread];          // zero / NaN area or a non-finite coordinate: never quick-rejected
#pragma unroll
    for (int j = 0; j < kAssignPerThread; ++j) {
        const int la = threadIdx.x + j * kAssignThreads;
        valid[j] = la < n_local;
        ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        aarea[j] = 0.f;
        if (valid[j]) {
            ab[j] = corners_of(anchors[a_begin + la]);
            aarea[j] = clamped_area(ab[j].x, ab[j].y, ab[j].z, ab[j].w);
        }
        aslow[j] = !(aarea[j] > 0.f) || !(isfinite(ab[j].x) && isfinite(ab[j].y) && isfinite(ab[j].z) && isfinite(ab[j].w));
    }

    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float* row = gt_rows + (size_t)(g0 + g) * gt_cols;
        const float4 bx = make_float4(row[0], row[1], row[2], row[3]);
        gbox[g] = bx;
        garea[g] = clamped_area(bx.x, bx.y, bx.z, bx.w);
        gslow[g] = !(isfinite(bx.x) && isfinite(bx.y) && isfinite(bx.z) && isfinite(bx.w));
        gcs[g] = make_float2(row[SSD_CLASS_COL], row[SSD_SCORE_COL]);
        best[g] = 0ull;                   // nothing overlapping seen: stands for (IoU 0, anchor 0)
    }
    __syncthreads();
    trace_.mark(0);

    int n_pos = 0, n_ign = 0, n_nan = 0;
    // ---- phase 1: per-anchor best GT, per-GT best anchor of this CTA ----
    {
        // IoU 0 against every box so far: the state torch.max would be in after a row of zeros
        float best_iou[kAssignPerThread];
        int best_g[kAssignPerThread];
#pragma unroll
        for (int j = 0; j < kAssignPerThread; ++j) { best_iou[j] = 0.f; best_g[j] = 0; }
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const float4 gb = gbox[g];
            const float ga = garea[g];
            // Quick reject, four compares per pair: if the boxes do not overlap with positive extent on both
            // axes, one clamped side of the intersection is exactly 0 and (finite inputs) so is `inter` -- the
            // pair is disjoint, IoU +0: it can neither raise the anchor's running maximum nor beat the
            // (0, anchor 0) entry every GT starts with.  ~90 % of the pairs end here.  Non-finite boxes and
            // zero-area anchors (0/0 -> NaN) always take the full path.
            const bool gs = gslow[g] != 0;
            bool ov[kAssignPerThread];
            bool any_ov = false;
#pragma unroll
            for (int j = 0; j < kAssignPerThread; ++j) {
                ov[j] = valid[j] && (gs || aslow[j] ||
                                     (gb.z > ab[j].x && ab[j].z > gb.x && gb.w > ab[j].y && ab[j].w > gb.y));
                any_ov = any_ov || ov[j];
            }
            if (!__any_sync(FULL, any_ov)) continue;
            // this thread's best (key, anchor) for the box: j ascending = anchor ascending, strict > keeps the lower
            uint32_t kbest = 0u;
            uint32_t abest = 0xFFFFFFFFu;
#pragma unroll
            for (int j = 0; j < kAssignPerThread; ++j) {
                if (!ov[j]) continue;
                const float iw = fmaxf(fsub(fminf(gb.z, ab[j].z), fmaxf(gb.x, ab[j].x)), 0.f);
                const float ih = fmaxf(fsub(fminf(gb.w, ab[j].w), fmaxf(gb.y, ab[j].y)), 0.f);
                const float inter = fmul(iw, ih);
                // (0/0 -> NaN needs both areas to be zero / NaN; NaN inter is not == 0.)
                bool live = !(inter == 0.f);
                if (!(aarea[j] > 0.f)) live = live || !(ga > 0.f);
                if (!live) continue;
                const float uni = fsub(fadd(ga, aarea[j]), inter);
                const float v = (inter == 0.f && uni > 0.f) ? 0.f : fdiv(inter, uni);
                // torch.max(dim=0): first maximum wins, NaN propagates and sticks
                if (!(v <= best_iou[j]) && !(best_iou[j] != best_iou[j])) { best_iou[j] = v; best_g[j] = g; }
                uint32_t key = ordered_key(v);
                if (v != v) key = 0xFFFFFFFFu;
                if (key > kbest) { kbest = key; abest = (uint32_t)(a_begin + threadIdx.x + j * kAssignThreads); }
            }
            // per-GT argmax over anchors: largest key, lowest anchor among its holders
            const uint32_t wmax = __reduce_max_sync(FULL, kbest);
            if (wmax > 0x80000000u) {               // somebody overlaps (key(+0) == 0x80000000)
                const uint32_t amin = __reduce_min_sync(FULL, kbest == wmax ? abest : 0xFFFFFFFFu);
                if (lane_id() == 0) {
                    const unsigned long long w = ((unsigned long long)wmax << 32) | (unsigned long long)(0xFFFFFFFFu - amin);
                    if (w > best[g]) atomicMax(&best[g], w);
                }
            }
        }
        // ---- phase 3 (same registers): thresholds (matcher.py:49-50) and the target rows.  Each thread writes
        //      its four rows as three float2 stores per row (rows are 24 bytes: 8-byte aligned); the three
        //      store instructions of a warp cover the same six 128-byte lines. ----
        float2* out = reinterpret_cast<float2*>(target + ((size_t)img * A + a_begin) * SSD_TARGET_COLS);
        const float zero_wh = fmul(logf(fadd(0.f, coding.eps)), coding.wh);
#pragma unroll
        for (int j = 0; j < kAssignPerThread; ++j) {
            if (!valid[j]) continue;
            const int la = threadIdx.x + j * kAssignThreads;
            int m = SSD_NOT_MATCHED;
            if (G > 0) {
                m = best_g[j];
                if (best_iou[j] < unmatched_thr) m = SSD_NOT_MATCHED;              // matcher.py:49
                else if (best_iou[j] < matched_thr) m = SSD_IGNORE;                // matcher.py:50
            }
            float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);                  // target_assigner.py:38: rows start as zeros
            float2 cs;
            if (m >= 0) {
                bx = gbox[m];
                cs = gcs[m];
                int p_, i_, n_;
                row_class_counts(cs.x, bx, p_, i_, n_);
                n_pos += p_; n_ign += i_; n_nan += n_;
            } else if (m == SSD_IGNORE) {
                cs = make_float2((float)SSD_IGNORE_CLASS, (float)SSD_IGNORE_CLASS);
                n_ign += 1;
            } else {
                cs = make_float2((float)SSD_NEGATIVE_CLASS, 1.f);
            }
            if (coding.on) {
                const float4 pr = anchors[a_begin + la];
                if (m < 0 && pr.z > 0.f && pr.w > 0.f) {
                    // an all-zero box (97 % of the rows): w / p_w = +0 exactly, so both size terms are the constant
                    // log(0 + eps) * wh_scale -- the same operations the full path would execute, minus two divides
                    // and two logarithms per row
                    bx = make_float4(fmul(fdiv(fsub(0.f, pr.x), pr.z), coding.xy), fmul(fdiv(fsub(0.f, pr.y), pr.w), coding.xy),
                                     zero_wh, zero_wh);
                } else {
                    bx = coded_box(bx, pr, coding);
                }
            }
            out[la * 3 + 0] = make_float2(bx.x, bx.y);
            out[la * 3 + 1] = make_float2(bx.z, bx.w);
            out[la * 3 + 2] = cs;
            if (match_out != nullptr) match_out[(size_t)img * A + a_begin + la] = m;
        }
    }
    trace_.mark(1);
    // dependents may be scheduled from here on (not earlier: an early dependent only squats on the SMs)
    griddep_launch_dependents();
    __syncthreads();                       // every warp is done with best[] (phase 1 atomics)

    // ---- phase 2: publish this CTA's per-GT winners ----
    if (force_match) {
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            const unsigned long long w = best[g];
            if (w != 0ull) atomicMax(&gbest[(size_t)img * max_gt + g], w);
        }
    }
    int* cnt = counters + (size_t)img * kAssignCounters;
    n_pos = __reduce_add_sync(FULL, n_pos);
    n_ign = __reduce_add_sync(FULL, n_ign);
    n_nan = __reduce_add_sync(FULL, n_nan);
    if (lane_id() == 0) {
        if (n_pos) atomicAdd(cnt + 1, n_pos);
        if (n_ign) atomicAdd(cnt + 2, n_ign);
        if (n_nan) atomicAdd(cnt + 3, n_nan);
    }

    // ---- phase 4: the last CTA of the image applies the forced matches and closes the statistics ----
    // (the CTA barrier orders every thread's stores before thread 0's fence, which is cumulative)
    __syncthreads();
    trace_.mark(2);
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(cnt, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    trace_.mark(3);
    if (!s_last) return;
    __threadfence();
    // The tail is a chain of dependent global round trips on ONE CTA per image while the rest of the GPU
    // waits for the kernel to end, so independent loads are issued together: the other CTAs' totals (all
    // published before their tickets) go out with the per-GT winners, and the forced-match deltas are
    // added in shared memory instead of bouncing off the global counters again.
    __shared__ int s_delta[3];
    int base_pos = 0, base_ign = 0, base_nan = 0;
    if (threadIdx.x == 0) {
        base_pos = __ldcg(cnt + 1); base_ign = __ldcg(cnt + 2); base_nan = __ldcg(cnt + 3);
        s_delta[0] = 0; s_delta[1] = 0; s_delta[2] = 0;
    }
    if (force_match && G > 0) {
        int* ganchor = reinterpret_cast<int*>(best);                   // [G] (the 64-bit table is done with)
        for (int g0_ = 0; g0_ < G; g0_ += blockDim.x) {
            const int g = g0_ + threadIdx.x;
            unsigned long long w = 0ull;
            if (g < G) w = __ldcg(&gbest[(size_t)img * max_gt + g]);
            if (g < G) ganchor[g] = w == 0ull ? 0 : (int)(0xFFFFFFFFu - (uint32_t)(w & 0xFFFFFFFFull));
        }
        __syncthreads();
        int d_pos = 0, d_ign = 0, d_nan = 0;
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            const int a = ganchor[g];
            bool winner = true;                                        // the highest GT index wins a collision
            for (int h = g + 1; h < G; ++h) winner = winner && ganchor[h] != a;
            if (!winner) continue;
            float* row = target + ((size_t)img * A + a) * SSD_TARGET_COLS;
            const float2 o0 = __ldcg(reinterpret_cast<const float2*>(row));
            const float2 o1 = __ldcg(reinterpret_cast<const float2*>(row) + 1);
            const float2 o2 = __ldcg(reinterpret_cast<const float2*>(row) + 2);
            int m_old = 0;
            float4 prior = make_float4(0.f, 0.f, 1.f, 1.f);
            if (coding.on) {                                           // the row holds coded values: its corner box
                m_old = __ldcg(match_out + (size_t)img * A + a);       // comes from its previous match
                prior = anchors[a];
            }
            int p_, i_, n_;
            float4 old = make_float4(o0.x, o0.y, o1.x, o1.y);
            if (coding.on) old = m_old >= 0 ? gbox[m_old] : make_float4(0.f, 0.f, 0.f, 0.f);
            row_class_counts(o2.x, old, p_, i_, n_);
            d_pos -= p_; d_ign -= i_; d_nan -= n_;
            const float4 bx = gbox[g];
            const float2 cs = gcs[g];
            row_class_counts(cs.x, bx, p_, i_, n_);
            d_pos += p_; d_ign += i_; d_nan += n_;
            const float4 wr = coding.on ? coded_box(bx, prior, coding) : bx;
            reinterpret_cast<float2*>(row)[0] = make_float2(wr.x, wr.y);
            reinterpret_cast<float2*>(row)[1] = make_float2(wr.z, wr.w);
            reinterpret_cast<float2*>(row)[2] = cs;
            if (match_out != nullptr) match_out[(size_t)img * A + a] = g;
        }
        if (d_pos) atomicAdd(&s_delta[0], d_pos);
        if (d_ign) atomicAdd(&s_delta[1], d_ign);
        if (d_nan) atomicAdd(&s_delta[2], d_nan);
        __syncthreads();
    }
    if (stats != nullptr && threadIdx.x == 0) {
        stats[img * 4 + 0] = base_pos + s_delta[0];
        stats[img * 4 + 1] = base_ign + s_delta[1];
        stats[img * 4 + 2] = base_nan + s_delta[2];
        stats[img * 4 + 3] = G_raw;
    }
}

// ---------------------------------------------------------------------------------------------
// stand-alone pieces of the same path (API completeness: box_utils.iou, matcher.match_per_prediction)
// ---------------------------------------------------------------------------------------------
__global__ void pairwise_iou_kernel(const float4* __restrict__ a, int na, const float4* __restrict__ b, int nb,
                                    float* __restrict__ out) {
    griddep_wait();
    griddep_launch_dependents();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= nb) return;
    const float4 ga = a[i], gb = b[j];
    const float area_a = clamped_area(ga.x, ga.y, ga.z, ga.w);
    const float area_b = clamped_area(gb.x, gb.y, gb.z, gb.w);
    const float ix1 = fmaxf(ga.x, gb.x), iy1 = fmaxf(ga.y, gb.y);
    const float ix2 = fminf(ga.z, gb.z), iy2 = fminf(ga.w, gb.w);
    const float inter = fmul(fmaxf(fsub(ix2, ix1), 0.f), fmaxf(fsub(iy2, iy1), 0.f));
    out[(size_t)i * nb + j] = fdiv(inter, fsub(fadd(area_a, area_b), inter));
}

// bf/utils/box_utils.py:104-143 generalized IoU (https://arxiv.org/abs/1902.09630): iou - (enclosing - union) / enclosing.
// cartesian: out[na, nb]; otherwise element-wise over na == nb rows (grid.y == 1, i = j).
__global__ void generalized_iou_kernel(const float4* __restrict__ a, int na, const float4* __restrict__ b, int nb,
                                       int cartesian, float* __restrict__ out) {
    griddep_wait();
    griddep_launch_dependents();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    const int i = cartesian ? blockIdx.y : j;
    const float4 ga = a[i], gb = b[j];
    const float area_a = clamped_area(ga.x, ga.y, ga.z, ga.w);
    const float area_b = clamped_area(gb.x, gb.y, gb.z, gb.w);
    const float inter = clamped_area(fmaxf(ga.x, gb.x), fmaxf(ga.y, gb.y), fminf(ga.z, gb.z), fminf(ga.w, gb.w));
    const float uni = fsub(fadd(area_a, area_b), inter);
    const float enc = clamped_area(fminf(ga.x, gb.x), fminf(ga.y, gb.y), fmaxf(ga.z, gb.z), fmaxf(ga.w, gb.w));
    out[cartesian ? (size_t)i * nb + j : (size_t)j] = fsub(fdiv(inter, uni), fdiv(fsub(enc, uni), enc));
}

// one CTA: weights[G, A] -> box_idx[A].  Columns are walked by threads (coalesced along A).
__global__ void __launch_bounds__(1024)
match_per_prediction_kernel(const float* __restrict__ w, int G, int A, float matched_thr, float unmatched_thr,
                            int force_match, long long* __restrict__ box_idx, unsigned long long* __restrict__ best) {
    griddep_wait();
    griddep_launch_dependents();
    // best[G] lives in global scratch (single CTA, so plain atomics + __syncthreads order it)
    // generic weights (may be negative): start below every real key
    for (int g = threadIdx.x; g < G; g += blockDim.x) best[g] = 0ull;
    __syncthreads();
    for (int base = 0; base < A; base += blockDim.x) {
        const int a = base + threadIdx.x;
        const bool valid = a < A;
        float best_v = -INFINITY;
        int best_g = 0;
        for (int g = 0; g < G; ++g) {
            const float v = valid ? w[(size_t)g * A + a] : 0.f;
            if (valid && !(v <= best_v) && !(best_v != best_v)) { best_v = v; best_g = g; }
            uint32_t key = valid ? ordered_key(v) : 0u;
            if (valid && v != v) key = 0xFFFFFFFFu;
            const uint32_t wmax = __reduce_max_sync(FULL, key);
            const unsigned bal = __ballot_sync(FULL, key == wmax);
            if (wmax != 0u && lane_id() == __ffs(bal) - 1) {
                const unsigned long long word = ((unsigned long long)key << 32) |
                                                (unsigned long long)(0xFFFFFFFFu - (uint32_t)a);
                atomicMax(&best[g], word);
            }
        }
        if (valid) {
            long long m = best_g;
            if (best_v < unmatched_thr) m = SSD_NOT_MATCHED;
            else if (best_v < matched_thr) m = SSD_IGNORE;
            box_idx[a] = m;
        }
    }
    __syncthreads();
    if (force_match && threadIdx.x == 0) {
        // sequential, last writer wins -- G is small on this API path
        for (int g = 0; g < G; ++g) {
            const int a = (int)(0xFFFFFFFFu - (uint32_t)(best[g] & 0xFFFFFFFFull));
            box_idx[a] = g;
        }
    }
}

// per-GT winners of ssd_match_per_prediction: one small buffer per DEVICE, allocated on first use (never inside a
// stream capture) and kept; calls on one device must be serialised by the caller (API convenience path)
static unsigned long long* g_match_scratch[64] = {nullptr};

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_pairwise_iou(const float* a_corners, int num_a, const float* b_corners, int num_b, float* out,
                                void* stream) {
    SSD_REQUIRE(num_a >= 0 && num_b >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_pairwise_iou: negative size");
    if (num_a == 0 || num_b == 0) return SSD_OK;
    SSD_REQUIRE(a_corners && b_corners && out, SSD_ERR_INVALID_ARGUMENT, "ssd_pairwise_iou: null pointer");
    SSD_REQUIRE(aligned(a_corners, 16) && aligned(b_corners, 16), SSD_ERR_MISALIGNED,
                "ssd_pairwise_iou: boxes must be 16-byte aligned");
    SSD_REQUIRE(num_a <= 65535, SSD_ERR_UNSUPPORTED, "ssd_pairwise_iou: more than 65535 rows");
    dim3 grid((num_b + 255) / 256, num_a);
    SSD_CUDA(launch_pdl(pairwise_iou_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (const float4*)a_corners, num_a,
                        (const float4*)b_corners, num_b, out));
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_generalized_iou(const float* a_corners, int num_a, const float* b_corners, int num_b, int cartesian,
                                   float* out, void* stream) {
    SSD_REQUIRE(num_a >= 0 && num_b >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_generalized_iou: negative size");
    SSD_REQUIRE(cartesian || num_a == num_b, SSD_ERR_INVALID_ARGUMENT,
                "ssd_generalized_iou: element-wise form needs as many a rows as b rows (%d vs %d)", num_a, num_b);
    if (num_a == 0 || num_b == 0) return SSD_OK;
    SSD_REQUIRE(a_corners && b_corners && out, SSD_ERR_INVALID_ARGUMENT, "ssd_generalized_iou: null pointer");
    SSD_REQUIRE(aligned(a_corners, 16) && aligned(b_corners, 16), SSD_ERR_MISALIGNED,
                "ssd_generalized_iou: boxes must be 16-byte aligned");
    SSD_REQUIRE(!cartesian || num_a <= 65535, SSD_ERR_UNSUPPORTED, "ssd_generalized_iou: more than 65535 rows");
    dim3 grid((num_b + 255) / 256, cartesian ? num_a : 1);
    SSD_CUDA(launch_pdl(generalized_iou_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (const float4*)a_corners, num_a,
                        (const float4*)b_corners, num_b, cartesian, out));
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_match_per_prediction(const float* weights, int num_gt, int num_anchors, float matched_threshold,
                                        float unmatched_threshold, int force_match, int64_t* box_idx_out,
                                        void* stream) {
    SSD_REQUIRE(num_gt >= 1 && num_anchors >= 1, SSD_ERR_INVALID_ARGUMENT,
                "ssd_match_per_prediction: needs at least one box and one anchor");
    SSD_REQUIRE(weights && box_idx_out, SSD_ERR_INVALID_ARGUMENT, "ssd_match_per_prediction: null pointer");
    SSD_REQUIRE(num_gt <= kMaxGtPerImage, SSD_ERR_UNSUPPORTED, "ssd_match_per_prediction: more than %d boxes",
                kMaxGtPerImage);
    int dev = 0;
    SSD_CUDA(cudaGetDevice(&dev));
    SSD_REQUIRE(dev >= 0 && dev < 64, SSD_ERR_UNSUPPORTED, "ssd_match_per_prediction: device index %d", dev);
    if (g_match_scratch[dev] == nullptr) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        SSD_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cap));
        SSD_REQUIRE(cap == cudaStreamCaptureStatusNone, SSD_ERR_UNSUPPORTED,
                    "ssd_match_per_prediction: first call on a device allocates scratch and cannot be captured");
        SSD_CUDA(cudaMalloc(&g_match_scratch[dev], sizeof(unsigned long long) * kMaxGtPerImage));
    }
    SSD_CUDA(launch_pdl(match_per_prediction_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, weights, num_gt,
                        num_anchors, matched_threshold, unmatched_threshold, force_match, (long long*)box_idx_out,
                        g_match_scratch[dev]));
    count_launch();
    return SSD_OK;
}

extern "C" size_t ssd_assign_workspace_bytes(int batch, int max_gt) {
    if (batch <= 0) return 256;
    if (max_gt < 1) max_gt = 1;
    return round_up((size_t)batch * kAssignCounters * sizeof(int), 256) +
           round_up((size_t)batch * max_gt * sizeof(unsigned long long), 256);
}

static int assign_impl(const float* anchors, const float* gt_rows, int gt_cols, const int32_t* gt_offsets,
                       int max_gt, int batch, int num_anchors, float matched_threshold,
                       float unmatched_threshold, int force_match, float* target_out, int32_t* match_out,
                       int32_t* stats_out, void* workspace, size_t workspace_bytes, void* stream, BoxCoding coding) {
    SSD_REQUIRE(batch >= 0 && num_anchors >= 0 && max_gt >= 0, SSD_ERR_INVALID_ARGUMENT,
                "ssd_assign_targets: negative shape");
    if (batch == 0 || num_anchors == 0) return SSD_OK;
    SSD_REQUIRE(anchors && gt_offsets && target_out && workspace, SSD_ERR_INVALID_ARGUMENT,
                "ssd_assign_targets: null pointer");
    SSD_REQUIRE(gt_rows || max_gt == 0, SSD_ERR_INVALID_ARGUMENT, "ssd_assign_targets: gt_rows is null");
    SSD_REQUIRE(gt_cols >= 6, SSD_ERR_INVALID_ARGUMENT, "ssd_assign_targets: gt rows need >= 6 columns, got %d", gt_cols);
    SSD_REQUIRE(matched_threshold >= unmatched_threshold, SSD_ERR_INVALID_ARGUMENT,
                "ssd_assign_targets: matched_threshold < unmatched_threshold");
    SSD_REQUIRE(max_gt <= kMaxGtPerImage, SSD_ERR_UNSUPPORTED, "ssd_assign_targets: more than %d boxes per image",
                kMaxGtPerImage);
    SSD_REQUIRE(batch <= 65535, SSD_ERR_UNSUPPORTED, "ssd_assign_targets: more than 65535 images per call");
    SSD_REQUIRE(aligned(anchors, 16), SSD_ERR_MISALIGNED, "ssd_assign_targets: anchors not 16-byte aligned");
    SSD_REQUIRE(aligned(target_out, 8), SSD_ERR_MISALIGNED, "ssd_assign_targets: target not 8-byte aligned");
    SSD_REQUIRE(aligned(workspace, 256), SSD_ERR_MISALIGNED, "ssd_assign_targets: workspace must be 256-byte aligned");
    const size_t need = ssd_assign_workspace_bytes(batch, max_gt);
    SSD_REQUIRE(workspace_bytes >= need, SSD_ERR_WORKSPACE, "ssd_assign_targets: workspace %zu < %zu bytes",
                workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    SSD_CUDA(zero_async(workspace, need, st));
    int* counters = (int*)workspace;
    unsigned long long* gbest = (unsigned long long*)((unsigned char*)workspace +
                                                      round_up((size_t)batch * kAssignCounters * sizeof(int), 256));
    const int gcap = max_gt > 0 ? max_gt : 1;
    const size_t smem = (size_t)gcap * (sizeof(float4) + sizeof(unsigned long long) + sizeof(float2) + sizeof(float) + sizeof(int));
    SSD_REQUIRE(smem <= 200 * 1024, SSD_ERR_UNSUPPORTED, "ssd_assign_targets: %zu bytes of shared memory needed (boxes %d)",
                smem, max_gt);
    SSD_CUDA(cudaFuncSetAttribute(assign_targets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((num_anchors + kAssignTile - 1) / kAssignTile), (unsigned)batch);
    LaunchTimer lt_("assign", st);
    SSD_CUDA(launch_pdl(assign_targets_kernel, grid, dim3(kAssignThreads), smem, st, (const float4*)anchors, gt_rows,
                        gt_cols, gt_offsets, num_anchors, gcap, matched_threshold, unmatched_threshold, force_match,
                        target_out, match_out, stats_out, counters, gbest, coding));
    count_launch();
    return SSD_OK;
}

extern "C" int ssd_assign_targets(const float* anchors, const float* gt_rows, int gt_cols, const int32_t* gt_offsets,
                                  int max_gt, int batch, int num_anchors, float matched_threshold,
                                  float unmatched_threshold, int force_match, float* target_out, int32_t* match_out,
                                  int32_t* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
    BoxCoding coding;
    coding.on = 0; coding.xy = 1.f; coding.wh = 1.f; coding.eps = 0.f;
    return assign_impl(anchors, gt_rows, gt_cols, gt_offsets, max_gt, batch, num_anchors, matched_threshold,
                       unmatched_threshold, force_match, target_out, match_out, stats_out, workspace, workspace_bytes,
                       stream, coding);
}

extern "C" int ssd_assign_targets_encoded(const float* anchors, const float* gt_rows, int gt_cols,
                                          const int32_t* gt_offsets, int max_gt, int batch, int num_anchors,
                                          float matched_threshold, float unmatched_threshold, int force_match,
                                          float xy_scale, float wh_scale, float eps, float* target_out,
                                          int32_t* match_out, int32_t* stats_out, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    SSD_REQUIRE(match_out != nullptr || batch == 0 || num_anchors == 0, SSD_ERR_INVALID_ARGUMENT,
                "ssd_assign_targets_encoded: match_out is required");
    BoxCoding coding;
    coding.on = 1; coding.xy = xy_scale; coding.wh = wh_scale; coding.eps = eps;
    return assign_impl(anchors, gt_rows, gt_cols, gt_offsets, max_gt, batch, num_anchors, matched_threshold,
                       unmatched_threshold, force_match, target_out, match_out, stats_out, workspace, workspace_bytes,
                       stream, coding);
}

SSD_DEFINE_TRACE_SETTER(set_trace_assign)
