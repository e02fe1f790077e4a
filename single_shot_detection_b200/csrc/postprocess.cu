// Post-processor: detection/postprocessor.py:24-78 (score conversion, decode, per-class
// threshold -> top-k -> NMS, final top-k) with bf/utils/box_utils.py:165-194 (top-k + hard NMS,
// torchvision.ops.nms semantics) -- batched over images x classes, five launches, no host sync.
//
// The reference walks B x (C-1) Python iterations, each gathering a score COLUMN out of the
// class-fastest [B, A, C] layout.  Here the logits are only ever streamed row-major:
//
//  1. score_pass1   streams logits (TMA bulk -> smem ring).  Per row: max, sum exp (SOFTMAX) ->
//                   rowstat[b,a] = (max, sum).  Per element a monotone "gate" value g (log-prob
//                   for SOFTMAX, the raw value for SIGMOID / IDENTITY).  Each warp keeps the
//                   running column maxima of the rows it sees and writes one vector per ROW BLOCK
//                   (~32 rows): blockmax[b, blk, c].
//  2. class_gate    one warp per (image, class): the K-th largest block maximum is a guaranteed
//                   lower bound of the K-th largest score of that class (K different rows reach
//                   it), and for i.i.d. scores only ~1.2 K elements exceed it.  gate[b,c] =
//                   max(score-threshold gate, that bound - slack).
//  3. score_pass2   streams the logits again (L2-resident for the SSD300-sized configs) with
//                   rowstat, recomputes g bit-identically and appends (anchor, raw value) of the
//                   few survivors g > gate[b,c] to a per-(image,class) candidate list.
//  4. segment_nms   one CTA per (image, class): exact score for each candidate (same fp32 ops as
//                   the reference: exp(x-max)/sum, 1/(1+exp(-x))), exact `score > threshold`,
//                   sort by (score desc, anchor asc), keep K, decode + to_corners only those
//                   boxes, IoU bit-matrix by warp ballot, sequential sweep -> kept rows.
//  5. image_topk    one CTA per image: class-major concatenation, or -- when more than T rows
//                   survive -- radix-select the T best and sort them descending.
//
// Exactness: steps 1-3 only PRUNE; every element whose exact score could rank in the top K of
// its class passes the gate (the slack is orders of magnitude above the fp32 error of g).  The
// exact scores, the exact threshold compare, the ordering and the NMS overlap test are evaluated
// in step 4 with separately rounded fp32 ops in the reference's order; the NMS overlap compare
// is float-vs-double as in torchvision's CPU kernel.
#include <cooperative_groups.h>
#include <float.h>
#include <math.h>

#include "rowstream.cuh"

namespace cg = cooperative_groups;

namespace ssd {

// ---------------------------------------------------------------------------------------------
// plan: shapes, tiling and workspace layout (host)
// ---------------------------------------------------------------------------------------------
struct PostPlan {
    int B, A, C, Cf, first_fg, K, T, det_cap, converter, box_input;
    ScoreGrid g;             // tiling of the streaming kernels (pass 2's item split)
    ScoreGrid g1;            // the same tiling with pass 1's item split (its grid differs)
    int grid;                // CTAs of pass 2
    int grid1;               // CTAs of pass 1 (less shared memory per CTA: one more per SM)
    int nblk;                // row blocks per image
    int cand_cap;            // candidate slots per (image, class)
    // workspace offsets (bytes)
    size_t off_rowstat, off_blockmax, off_gate, off_cand_count, off_cand, off_kept_count, off_kept, off_status,
        off_anchor_tmp, off_score_hist, off_bhist, off_image_done;
    size_t zero_begin, zero_bytes;   // counters and histograms of the LATER launches: zeroed by pass 1 itself
    bool gate_hist;                  // gates from a histogram of the block maxima (SOFTMAX / SIGMOID)
    bool block_per_step;             // pass 1: a warp step of 32 rows is a block (score_pass1_kernel<.., BPS = true>)
    float bin_lo, bin_scale;
    float soft_thr;
    size_t total_bytes;
    // fused candidate selection (fused_select_kernel): one thread-block cluster per image, 0 = not used
    int fused_cluster, fused_rows_per_cta, fused_merge, fused_split, fused_chunk_floats;
    unsigned fused_off_slab, fused_off_trow, fused_off_hist, fused_off_gate, fused_off_stage;
    size_t fused_smem;
};
static void plan_fused(PostPlan& pl);

constexpr int kKeptCols = 6;          // x1,y1,x2,y2,score,anchor(bits)
constexpr int kScoreBins = 2048;      // per-image histogram of the kept scores (final top-k)
constexpr int kTopkBoundaryCap = 1024;
constexpr int kQueueCap = 96;         // pass-2 survivor queue: entries per warp
// Class gates from a histogram of the block maxima: kGateBins linear bins over [lo, lo + kGateBins / scale)
// in the gate domain (log-probability / logit), two 16-bit counters per word.  The range starts
// just below the score threshold -- nothing under it can become a detection -- see gate_range().
constexpr int kGateBins = 256;
constexpr int kGateWords = kGateBins / 2;
constexpr int kGateStride = kGateWords + 1;      // shared-memory histogram of pass 2: words per column
constexpr size_t kQueueBytes = (size_t)kConsumerWarps * kQueueCap * 3 * sizeof(uint32_t);

// Monotone (non-decreasing) map of a kept score to a histogram bin; probabilities spread over the
// whole range, anything else is clamped (the boundary bin is ranked exactly, so only speed depends
// on the spread).
__device__ __forceinline__ int score_bin(float v) {
    const float s = fminf(fmaxf(__fmul_rn(v, (float)(kScoreBins - 2)), 0.f), (float)(kScoreBins - 2));
    return 1 + (int)s;                 // 1 .. kScoreBins-1 (NaN never reaches here: it fails score > thr)
}
constexpr int kMaxPerClass = 512;
constexpr int kNmsThreadsMax = 256;     // segment_nms_kernel<NT>: NT = 32, 64, 128 or 256 threads per (image, class)
constexpr int kTopkThreads = 512;

// Bin range of the block-maximum histograms.  SOFTMAX: log-probabilities from just under
// log(threshold) (at least -16) up to 0; SIGMOID: logits from just under logit(threshold) in steps
// of 1/32.  Values outside clamp to the end bins; the first bin stands for "-inf".
static void gate_range(int converter, float score_thr, float& lo, float& scale) {
    if (converter == SSD_CONVERT_SOFTMAX) {
        lo = -16.f;
        if (score_thr > 0.f && score_thr < 1.f) lo = fmaxf(-16.f, logf(score_thr) - 0.05f);
        scale = (float)kGateBins / -lo;
    } else {
        lo = -16.f;
        if (score_thr > 0.f && score_thr < 1.f) lo = fmaxf(-16.f, logf(score_thr / (1.f - score_thr)) - 0.05f);
        scale = 32.f;
    }
}

static int make_plan(const ssd_postprocess_params* p, PostPlan& pl) {
    SSD_REQUIRE(p != nullptr, SSD_ERR_INVALID_ARGUMENT, "ssd_postprocess: null params");
    SSD_REQUIRE(p->batch >= 0 && p->num_anchors >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_postprocess: negative shape");
    SSD_REQUIRE(p->num_cols >= 1 && p->num_cols <= kMaxScoreCols, SSD_ERR_UNSUPPORTED,
                "ssd_postprocess: num_cols %d outside 1..%d", p->num_cols, kMaxScoreCols);
    SSD_REQUIRE(p->converter >= SSD_CONVERT_SOFTMAX && p->converter <= SSD_CONVERT_IDENTITY, SSD_ERR_INVALID_ARGUMENT,
                "Wrong value for score_converter: %d", p->converter);
    SSD_REQUIRE(p->first_fg_col >= 0 && p->first_fg_col < p->num_cols, SSD_ERR_INVALID_ARGUMENT,
                "ssd_postprocess: first_fg_col %d outside the %d score columns", p->first_fg_col, p->num_cols);
    SSD_REQUIRE(p->box_input == SSD_BOXES_ENCODED || p->box_input == SSD_BOXES_CORNERS, SSD_ERR_INVALID_ARGUMENT,
                "ssd_postprocess: bad box_input %d", p->box_input);
    SSD_REQUIRE(p->max_per_class >= 1 && p->max_per_class <= kMaxPerClass, SSD_ERR_UNSUPPORTED,
                "ssd_postprocess: max_per_class %d outside 1..%d", p->max_per_class, kMaxPerClass);
    pl.B = p->batch; pl.A = p->num_anchors; pl.C = p->num_cols; pl.first_fg = p->first_fg_col;
    pl.Cf = pl.C - pl.first_fg; pl.K = p->max_per_class; pl.T = p->max_total; pl.det_cap = p->det_capacity;
    pl.converter = p->converter; pl.box_input = p->box_input;
    pl.soft_thr = p->soft_threshold;
    SSD_REQUIRE(!p->soft_nms || p->soft_sigma > 0.f, SSD_ERR_INVALID_ARGUMENT, "ssd_postprocess: soft-NMS needs sigma > 0");
    const long long all_rows = (long long)pl.Cf * pl.K;
    const long long need = (pl.T > 0 && pl.T < all_rows) ? pl.T : all_rows;
    SSD_REQUIRE(pl.det_cap >= need, SSD_ERR_INVALID_ARGUMENT,
                "ssd_postprocess: det_capacity %d below the %lld rows an image can produce", pl.det_cap, need);

    const int A1 = pl.A > 0 ? pl.A : 1;
    ScoreGrid& g = pl.g;
    plan_tiles(g, pl.B, A1, pl.C, pl.converter == SSD_CONVERT_SOFTMAX);
    g.first_fg = pl.first_fg;
    // rows per block: ~32, smaller when the image has few anchors so that >= ~2K blocks exist
    int want_rows = A1 / (2 * pl.K);
    if (want_rows > 32) want_rows = 32;
    const int rows_per_warp_tile = g.tile_rows / kConsumerWarps;
    int gt = want_rows / rows_per_warp_tile;
    if (gt < 1) gt = 1;
    g.group_tiles = gt;
    g.groups_per_image = (g.tiles_per_image + gt - 1) / gt;
    g.num_items = pl.B * g.groups_per_image;
    // fewer rows per block than a warp sees in one tile: keep `split` row slots separate
    int split = 1;
    const int slots = 32 / lanes_per_row(pl.C, true);     // row slots per warp step of pass 1
    while (split < slots && rows_per_warp_tile * gt / split > (want_rows > 0 ? want_rows : 1)) split <<= 1;
    g.split = split;
    g.nblk = g.groups_per_image * kConsumerWarps * split;
    pl.nblk = g.nblk;
    pl.gate_hist = pl.converter != SSD_CONVERT_IDENTITY && pl.nblk < 65536 && pl.C <= 32;
    {   // tuning knob, read once
        static const int gate_knob = [] { const char* e = getenv("SSD_GATE"); return e ? (e[0] == 'h' ? 1 : 0) : -1; }();
        if (gate_knob >= 0) pl.gate_hist = gate_knob == 1 && pl.converter != SSD_CONVERT_IDENTITY && pl.C <= 64;
    }
    // a warp step of the row-per-lane shape is exactly one block: the block maxima come straight out of the step.
    // (NaN block maxima -- a column that is NaN in all 32 rows -- are read as -inf by the histogram gate only.)
    pl.block_per_step = lanes_per_row(pl.C, true) == 1 && pl.gate_hist && split == 1 && gt == 1 && rows_per_warp_tile == 32;
    {
        static const int bps_knob = [] { const char* e = getenv("SSD_BPS"); return e ? (e[0] != '0') : 1; }();     // tuning knob, read once
        if (!bps_knob) pl.block_per_step = false;
    }
    SSD_REQUIRE((long long)pl.B * A1 < 0x7FFFFFFFll && (long long)pl.B * pl.nblk * g.bm_stride < 0x7FFFFFFFll, SSD_ERR_UNSUPPORTED,
                "ssd_postprocess: %d x %d rows exceed the 32-bit row index of the streaming kernels", pl.B, A1);
    pl.g1 = g;
    pl.grid1 = stream_grid(pl.g1);
    pl.grid = stream_grid(g, kQueueBytes + round_up((size_t)pl.C * sizeof(float), 16) +
                                 (pl.gate_hist ? (size_t)pl.C * kGateStride * sizeof(uint32_t) : 0));
    int cap = 512;                        // power of two (the segment sort pads to one), >= 8K
    while (cap < 8 * pl.K && cap < 4096) cap <<= 1;
    pl.cand_cap = cap;
    plan_fused(pl);

    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += round_up(bytes, 256); return o; };
    const size_t BA = (size_t)pl.B * A1 + 2;
    // block counts stay below 65536 (16-bit histogram counters) for every supported A
    // Few columns: the gates come from histograms inside pass 2 (one launch less).  Many columns
    // (COCO's 81): a pass-2 CTA would spend longer on its image's gates than the separate
    // class_gate launch costs, measured in the step graph (profiles/), so that one stays.
    gate_range(pl.converter, p->score_threshold, pl.bin_lo, pl.bin_scale);
    pl.off_rowstat = take(BA * sizeof(float2));
    pl.off_blockmax = take((size_t)pl.B * pl.nblk * pl.g.bm_stride * sizeof(float));
    pl.off_gate = take((size_t)pl.B * pl.C * sizeof(float));
    pl.off_cand = take((size_t)pl.B * pl.Cf * pl.cand_cap * sizeof(uint2));
    pl.off_kept_count = take((size_t)pl.B * pl.Cf * sizeof(int));
    pl.off_kept = take((size_t)pl.B * pl.Cf * pl.K * kKeptCols * sizeof(float));
    pl.off_anchor_tmp = take((size_t)pl.B * (pl.det_cap > 0 ? pl.det_cap : 1) * sizeof(int));
    pl.zero_begin = off;
    pl.off_bhist = 0;
    pl.off_cand_count = take((size_t)pl.B * pl.Cf * sizeof(int));
    pl.off_status = take(4 * sizeof(int));
    pl.off_score_hist = take((size_t)pl.B * kScoreBins * sizeof(int));
    pl.off_image_done = take((size_t)pl.B * sizeof(int));
    pl.zero_bytes = off - pl.zero_begin;
    pl.total_bytes = off;
    return SSD_OK;
}

// ---------------------------------------------------------------------------------------------
// 1. score_pass1: row statistics + per-block column maxima of the gate value
//    SOFTMAX: gate value of an element = its log-probability  x - (max + log(sum));
//    SIGMOID / IDENTITY: the raw value (both converters are monotone per element).
// ---------------------------------------------------------------------------------------------
// Histogram of the block maxima of one (image, column): bin = floor((g - lo) * scale), clamped; the
// lower edge of a bin, lo + bin / scale, can exceed a value of the bin only by the rounding of these
// few operations -- far inside the slack gate_from_kth subtracts.
struct GateBins {
    float lo, scale;
};
__device__ __forceinline__ int gate_bin(float gval, GateBins gb) {
    const float s = fminf(fmaxf(__fmul_rn(__fsub_rn(gval, gb.lo), gb.scale), 0.f), (float)(kGateBins - 1));
    return (int)s;                      // NaN -> 0
}
__device__ __forceinline__ void bump_gate_bin(uint32_t* __restrict__ words, float gval, GateBins gb) {
    const int b = gate_bin(gval, gb);
    atomicAdd(words + (b >> 1), (b & 1) ? 0x10000u : 1u);
}
__device__ __forceinline__ float gate_bin_edge(int bin, GateBins gb) {
    return bin <= 0 ? -INFINITY : __fadd_rn(gb.lo, __fdiv_rn((float)bin, gb.scale));
}

// BPS ("block per step", row-per-lane shapes with one warp step per tile and block): the block maxima are the
// CREDUX of the step's gate values themselves -- no running per-lane maxima, no resets, no block-end branch.
template <int Q, int NREG, int CMIN, int CONV, bool BPS>
__global__ void __launch_bounds__(kStreamThreads)
score_pass1_kernel(const float* __restrict__ scores, ScoreGrid g, float2* __restrict__ rowstat,
                   float* __restrict__ blockmax, uint32_t* __restrict__ loss_keys, uint4* __restrict__ zero_ptr,
                   unsigned zero_n16) {
    extern __shared__ __align__(128) unsigned char smem[];
    KernelTrace trace_(TR_PASS1);
    stream_init(smem);
    griddep_wait();
    if (warp_id() == kConsumerWarps) {
        producer_loop(smem, g, scores, nullptr, policy_evict_last());      // pass 2 re-reads the same bytes
        return;
    }
    // counters / histograms of the launches that FOLLOW (candidate counts, score histograms, tickets): zeroed
    // here, so the call needs no memset or zeroing launch in front of it.  Nothing this kernel accumulates
    // into needs a zero start: the block maxima are plain stores.
    for (unsigned i = blockIdx.x * (kConsumerWarps * 32) + threadIdx.x; i < zero_n16; i += gridDim.x * (kConsumerWarps * 32))
        zero_ptr[i] = make_uint4(0u, 0u, 0u, 0u);
    const RowLanes<Q> ln;
    const RowShape<Q, NREG, CMIN> shape(ln.sub, g.C);
    const int rows_per_warp = g.tile_rows / kConsumerWarps;
    const uint32_t sb = smem_u32(smem);
    const int wbase = warp_id() * rows_per_warp;
    float cmax[NREG];
#pragma unroll
    for (int i = 0; i < NREG; ++i) cmax[i] = -INFINITY;

    TileCursor cur;
    cur.start(g);
    for (; cur.valid(g); cur.next(g)) {
        const unsigned r0 = cur.first_row32(g);
        const int rows = cur.rows(g);
        ring_wait_full(sb, cur);
        const float* logits = ring_logits(smem, g, cur, r0);
        if constexpr (BPS) {
            // Q == 1, rows_per_warp == 32: lane l owns row wbase + l of the tile, the warp's 32 rows are one block
            static_assert(Q == 1, "block-per-step needs the row-per-lane shape");
            float4* dst = reinterpret_cast<float4*>(
                blockmax + (size_t)((unsigned)cur.img * (unsigned)g.nblk + (unsigned)cur.tile * kConsumerWarps + warp_id()) * g.bm_stride);
            constexpr int NV = (NREG + 3) / 4;
            if (wbase < rows) {
                const int lr = wbase + lane_id();
                const bool valid = lr < rows;
                float v[NREG];
                shape.load(v, logits + lr * g.C, 0);
                float t = 0.f;
                if (CONV == SSD_CONVERT_SOFTMAX) {
                    float m, sum;
                    row_max_sum<Q, NREG>(v, m, sum);
                    if (valid) rowstat[r0 + lr] = make_float2(m, sum);
                    // rows past the end of a partial tile: t = +inf turns every gate value into -inf / NaN, both
                    // of which the warp maximum drops (at least one lane of the warp holds a real row)
                    t = valid ? __fadd_rn(m, fast_log(sum)) : INFINITY;
                    // the sampler's criterion -log_softmax(x)[0] = (max + log(sum)) - x0 falls out of the same row
                    // statistics (mining.cu, same operations): one streamed read of the logits serves both
                    if (loss_keys != nullptr && valid) {
                        const uint32_t lk = ordered_key(__fsub_rn(t, v[0]));
                        loss_keys[r0 + lr] = lk == 0u ? 1u : lk;
                    }
                }
                ring_release(sb, cur);         // the row is in registers: the stage can be refilled during the reductions
                float r[NV * 4];
#pragma unroll
                for (int i = 0; i < NV * 4; ++i) {
                    if (i < NREG) {
                        const float x = CONV == SSD_CONVERT_SOFTMAX ? __fsub_rn(v[i], t) : (valid ? v[i] : -INFINITY);
                        r[i] = warp_max(x);
                    } else {
                        r[i] = -INFINITY;
                    }
                }
                if (lane_id() == 0) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        if (4 * j < g.bm_stride) dst[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                }
            } else {
                ring_release(sb, cur);
                if (lane_id() < NV && 4 * lane_id() < g.bm_stride) dst[lane_id()] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
        } else {
        for (int step = 0; step < rows_per_warp; step += RowLanes<Q>::kRowsPerWarpStep) {
            if (wbase + step >= rows) break;            // the rest of a partial tile (warp-uniform)
            const int lr = wbase + step + ln.rl;
            const bool valid = lr < rows;
            float v[NREG];
            shape.load(v, logits + lr * g.C, ln.sub);
            if (CONV == SSD_CONVERT_SOFTMAX) {
                float m, sum;
                row_max_sum<Q, NREG>(v, m, sum);
                if (valid && ln.sub == 0) rowstat[r0 + lr] = make_float2(m, sum);
                const float t = valid ? __fadd_rn(m, fast_log(sum)) : INFINITY;      // see the block-per-step branch
                if (loss_keys != nullptr && valid && ln.sub == 0) {
                    const uint32_t lk = ordered_key(__fsub_rn(t, v[0]));
                    loss_keys[r0 + lr] = lk == 0u ? 1u : lk;
                }
#pragma unroll
                for (int i = 0; i < NREG; ++i) cmax[i] = fmaxf(cmax[i], __fsub_rn(v[i], t));
            } else {
#pragma unroll
                for (int i = 0; i < NREG; ++i) cmax[i] = fmaxf(cmax[i], valid ? v[i] : -INFINITY);
            }
        }
        ring_release(sb, cur);
        if (cur.last_of_item()) {
            const size_t blk0 = (size_t)cur.image(g) * g.nblk + ((size_t)cur.group(g) * kConsumerWarps + warp_id()) * g.split;
            if (Q == 1 && g.split == 1) {
                // row-per-lane shape: one CREDUX per column; the results are warp-uniform, lane 0 stores them
                float r[NREG];
#pragma unroll
                for (int i = 0; i < NREG; ++i) r[i] = warp_max(cmax[i]);
                if (lane_id() == 0) {
                    float* dst = blockmax + blk0 * g.bm_stride;
#pragma unroll
                    for (int i = 0; i < NREG; ++i)
                        if (i < g.C) dst[i] = r[i];
                }
            } else {
                // merge the row slots of the warp down to `split` blocks, then one vector per block
#pragma unroll
                for (int i = 0; i < NREG; ++i) {
                    float x = cmax[i];
#pragma unroll
                    for (int o = Q; o < 32; o <<= 1)
                        if (o >= Q * g.split) x = fmaxf(x, __shfl_xor_sync(FULL, x, o));
                    cmax[i] = x;
                }
                if (ln.rl < g.split) {
                    float* dst = blockmax + (blk0 + ln.rl) * g.bm_stride;
#pragma unroll
                    for (int i = 0; i < NREG; ++i) {
                        const int col = ln.sub + i * Q;
                        if (col < g.C) dst[col] = cmax[i];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NREG; ++i) cmax[i] = -INFINITY;
        }
        }   // !BPS
    }
    griddep_launch_dependents();       // late: see the note at launch_pdl
}

template <int Q, int NREG, int CMIN, int CONV>
static auto pass1_kernel_ptr(bool block_per_step) {
    if constexpr (Q == 1) {
        if (block_per_step) return &score_pass1_kernel<Q, NREG, CMIN, CONV, true>;
    }
    return &score_pass1_kernel<Q, NREG, CMIN, CONV, false>;
}

// ---------------------------------------------------------------------------------------------
// 2. class_gate: one CTA per image, one warp per score column at a time.  The K-th largest block
//    maximum of a column is a guaranteed lower bound of the K-th largest score of that class.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float next_below(float x) {
    // largest float < x for finite x
    if (x == 0.f) return -FLT_MIN * FLT_EPSILON;     // -denorm_min
    uint32_t u = __float_as_uint(x);
    u = (x > 0.f) ? u - 1 : u + 1;
    return __uint_as_float(u);
}

__device__ __forceinline__ float gate_from_kth(int converter, float score_thr, float kth) {
    if (converter == SSD_CONVERT_SOFTMAX) {
        // gate domain = log-probability; 1e-4 of slack is ~100x the fp32 error of the gate value
        const float gthr = score_thr > 0.f ? logf(score_thr) - 1e-4f : -INFINITY;
        return fmaxf(gthr, kth - 1e-4f);
    }
    if (converter == SSD_CONVERT_SIGMOID) {
        // gate domain = logit.  Slack of 1e-4 RELATIVE IN PROBABILITY, evaluated in double so that
        // saturated scores (p == 1.0f for many logits) all stay candidates.
        float gthr;
        if (score_thr <= 0.f) gthr = -INFINITY;
        else if (score_thr >= 1.f) gthr = INFINITY;
        else {
            const double t = (double)score_thr * (1.0 - 1e-4);
            gthr = (float)(log(t / (1.0 - t))) - 1e-5f;
        }
        float gk = -INFINITY;
        if (kth > -INFINITY) {
            const double pk = 1.0 / (1.0 + exp(-(double)kth));
            const double pt = pk * (1.0 - 1e-4);
            const double x = log(pt / (1.0 - pt));
            gk = (float)x - 1e-5f * (1.f + fabsf((float)x));
            if (gk > kth) gk = kth - 1e-4f;
        }
        return fmaxf(gthr, gk);
    }
    // probabilities given: exact.  Candidates need score > thr and, once K blocks reach kth, score >= kth.
    float g = score_thr;
    if (kth > score_thr) g = next_below(kth);
    return g;
}

constexpr int kGateCols = 8;                 // score columns per CTA, one warp each
constexpr int kGateThreads = kGateCols * 32;
constexpr int kGateKeysPerLane = 16;         // (merged) blocks per lane held in registers
constexpr int kGateBits = 20;                // resolved MSBs of the K-th largest key; the rest rounds DOWN

// Any lower bound of the K-th largest block maximum is a valid gate, so (a) adjacent blocks are
// merged on load until at most 32*kGateKeysPerLane remain (a coarser partition of the rows is still
// a partition) and (b) the bisection stops after kGateBits bits and leaves the low bits zero
// (a slightly smaller key).  Both only let a few more candidates through.
__global__ void __launch_bounds__(kGateThreads)
class_gate_kernel(const float* __restrict__ blockmax, int C, int bm_stride, int first_fg, int nblk, int K, int converter,
                  float score_thr, float* __restrict__ gate) {
    extern __shared__ __align__(16) uint32_t skey[];          // [merged blocks][kGateCols + 1]
    KernelTrace trace_(TR_GATE);
    griddep_wait();
    griddep_launch_dependents();
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * kGateCols;
    const int lane = lane_id();
    const int merge = (nblk + 32 * kGateKeysPerLane - 1) / (32 * kGateKeysPerLane);
    const int nm = (nblk + merge - 1) / merge;                 // merged blocks, <= 512
    const float* src = blockmax + (size_t)b * nblk * bm_stride;
    // coalesced load: 8 consecutive columns of one block row per 8 threads, four rows in flight
    if (merge == 1) {
        const int total = nm * kGateCols;
        for (int t0 = threadIdx.x; t0 < total; t0 += 4 * kGateThreads) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + u * kGateThreads;
                const int mb = t / kGateCols, cc = t % kGateCols;
                v[u] = (t < total && c0 + cc < C) ? src[(size_t)mb * bm_stride + c0 + cc] : -INFINITY;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + u * kGateThreads;
                if (t < total) skey[(t / kGateCols) * (kGateCols + 1) + t % kGateCols] = ordered_key(v[u]);
            }
        }
    } else {
        // merged blocks: four of them per thread at a time, the `merge` loads of each issued without a
        // data-dependent exit (the loads do not depend on each other; a serial walk with an early break on
        // NaN made this loop one L2 round trip per element -- 14 of the kernel's 17 us at SSD512)
        const int total = nm * kGateCols;
        for (int t0 = threadIdx.x; t0 < total; t0 += 4 * kGateThreads) {
            float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            bool nan[4] = {false, false, false, false};
            for (int i = 0; i < merge; ++i) {
                float x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + u * kGateThreads;
                    const int mb = t / kGateCols, cc = t % kGateCols;
                    const int row = mb * merge + i;
                    x[u] = (t < total && c0 + cc < C && row < nblk) ? src[(size_t)row * bm_stride + c0 + cc] : -INFINITY;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    nan[u] = nan[u] || x[u] != x[u];
                    v[u] = x[u] > v[u] ? x[u] : v[u];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + u * kGateThreads;
                if (t < total) skey[(t / kGateCols) * (kGateCols + 1) + t % kGateCols] = ordered_key(nan[u] ? NAN : v[u]);   // NaN sorts on top
            }
        }
    }
    __syncthreads();
    const int col = c0 + warp_id();
    if (col >= C) return;
    if (col < first_fg) {
        if (lane == 0) gate[b * C + col] = INFINITY;
        return;
    }
    float kth = -INFINITY;
    if (nm >= K) {
        uint32_t key[kGateKeysPerLane];
#pragma unroll
        for (int j = 0; j < kGateKeysPerLane; ++j) {
            const int i = lane + 32 * j;
            key[j] = i < nm ? skey[i * (kGateCols + 1) + warp_id()] : 0u;      // 0 < every real key
        }
        // largest T (on the resolved bits) with count(key >= T) >= K
        uint32_t T = 0u;
        for (int bit = 31; bit > 31 - kGateBits; --bit) {
            const uint32_t trial = T | (1u << bit);
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < kGateKeysPerLane; ++j) cnt += key[j] >= trial;
            cnt = __reduce_add_sync(FULL, cnt);
            if (cnt >= K) T = trial;
        }
        // T == 0 only if fewer than K keys have any resolved bit set; key 0x007FFFFF is -inf
        kth = T < 0x00800000u ? -INFINITY : key_to_float(T);
        if (kth != kth) kth = -INFINITY;
    }
    if (lane == 0) gate[b * C + col] = gate_from_kth(converter, score_thr, kth);
}

// ---------------------------------------------------------------------------------------------
// 3. score_pass2: emit candidates (gate value above the class gate)
// ---------------------------------------------------------------------------------------------
// Survivors are parked in a per-warp shared-memory queue and appended to the per-(image, class)
// candidate lists in batches: one returning global atomic per entry, 32 in flight per warp, instead
// of a round trip in the middle of every row step.

struct CandQueue {
    uint32_t* seg;      // [kQueueCap] segment = image * Cf + foreground column
    uint32_t* anchor;   // [kQueueCap]
    uint32_t* val;      // [kQueueCap] raw score bits
    int n;              // warp-uniform fill
    __device__ __forceinline__ void flush(int* __restrict__ cand_count, uint2* __restrict__ cand, int cand_cap) {
        __syncwarp();
        for (int e = lane_id(); e < n; e += 32) {
            const uint32_t sg = seg[e];
            const int slot = atomicAdd(&cand_count[sg], 1);
            if (slot < cand_cap) cand[(size_t)sg * cand_cap + slot] = make_uint2(anchor[e], val[e]);
        }
        __syncwarp();
        n = 0;
    }
};

// Gates of one image from the block-maximum histograms (consumer warps only; a warp per column,
// kGateBatch columns' loads in flight).  Lane l holds bins 8 l .. 8 l + 7 of a column (four words);
// the gate is the lower edge of the highest bin whose suffix count reaches K -- a lower bound of
// the K-th largest block maximum, hence of the K-th largest score of the class.
struct GateHist {
    const float* blockmax;      // nullptr: gates are read from the `gate` array (class_gate_kernel)
    int nblk;
    int K, converter;
    float score_thr;
    GateBins bins;
};
constexpr int kGateBatch = 4;
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory"); }
__device__ __forceinline__ void image_gates(const GateHist& gh, const ScoreGrid& g, int img, float* __restrict__ sgate,
                                            uint32_t* __restrict__ shist) {
    const int lane = lane_id();
    static_assert(kGateWords == 32 * 4, "four words per lane");
    // histogram of the image's block maxima per column, built in shared memory (pass 1 stores the maxima,
    // [nblk, C] per image, 23 KB for SSD300: no global atomics and no zeroed scratch on that side).
    // Column stride kGateStride = kGateWords + 1 words: neighbouring lanes hold neighbouring COLUMNS of one
    // block whose maxima fall into similar bins -- with a stride of 128 words they would all hit one bank.
    for (int i = threadIdx.x; i < g.C * kGateStride; i += kConsumerWarps * 32) shist[i] = 0u;
    consumer_barrier();
    {
        const float* bm = gh.blockmax + (size_t)img * gh.nblk * g.bm_stride;
        const int total = gh.nblk * g.bm_stride;
        // a block's maxima are bm_stride floats (C rounded up to a multiple of four: whole 16-byte vectors, the padding
        // is skipped): every thread's loads (seven LDG.128 for SSD300) are in flight together -- one L2 round trip
        const float4* bm4 = reinterpret_cast<const float4*>(bm);
        const int total4 = total >> 2;
        constexpr int kDeep = 8;
        for (int i0 = threadIdx.x; i0 < total4; i0 += kDeep * kConsumerWarps * 32) {
            float4 v[kDeep];
#pragma unroll
            for (int u = 0; u < kDeep; ++u) {
                const int i = i0 + u * kConsumerWarps * 32;
                v[u] = i < total4 ? __ldcg(bm4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kDeep; ++u) {
                const int i = i0 + u * kConsumerWarps * 32;
                if (i < total4) {
                    const int col = (4 * i) % g.bm_stride;            // a vector never straddles two blocks
                    const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < g.C) bump_gate_bin(shist + (size_t)(col + j) * kGateStride, e[j], gh.bins);
                }
            }
        }
    }
    consumer_barrier();
    const uint32_t* base = shist + 4 * lane;
    for (int col0 = g.first_fg + warp_id(); col0 < g.C; col0 += kConsumerWarps * kGateBatch) {
        uint4 h[kGateBatch];
#pragma unroll
        for (int u = 0; u < kGateBatch; ++u) {
            const int col = col0 + u * kConsumerWarps;
            const uint32_t* src = base + (size_t)col * kGateStride;
            h[u] = col < g.C ? make_uint4(src[0], src[1], src[2], src[3]) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kGateBatch; ++u) {
            const int col = col0 + u * kConsumerWarps;
            if (col >= g.C) break;
            const uint32_t w[4] = {h[u].x, h[u].y, h[u].z, h[u].w};
            int own = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) own += (int)(w[j] & 0xFFFFu) + (int)(w[j] >> 16);
            int incl = own;                               // suffix sum towards higher lanes (= higher bins)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_down_sync(FULL, incl, o);
                if (lane + o < 32) incl += t;
            }
            int above = incl - own;
            int bin = -1;
            if (above < gh.K && incl >= gh.K) {
#pragma unroll
                for (int j = 7; j >= 0; --j) {
                    const int c = (int)((w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu);
                    if (bin < 0 && above + c >= gh.K) bin = 8 * lane + j;
                    above += c;
                }
            }
            const unsigned found = __ballot_sync(FULL, bin >= 0);
            if (bin >= 0) sgate[col] = gate_from_kth(gh.converter, gh.score_thr, gate_bin_edge(bin, gh.bins));
            if (found == 0u && lane == 0) sgate[col] = gate_from_kth(gh.converter, gh.score_thr, -INFINITY);
        }
    }
    for (int c = threadIdx.x; c < g.first_fg; c += kConsumerWarps * 32) sgate[c] = INFINITY;
}

// The same gates once per image in a launch of their own (SSD_GATE_KERNEL, default for C <= 32): one CTA per image
// instead of every pass-2 CTA repeating its image's histogram (~9 CTAs per image at SSD300 b32, 1.6 M warp
// instructions per step) -- one more launch in the chain, a quarter less work in pass 2; with two steps in
// flight the saved issue slots are what counts.
__global__ void __launch_bounds__(kConsumerWarps * 32)
hist_gate_kernel(GateHist gh, ScoreGrid g, float* __restrict__ gate) {
    extern __shared__ __align__(16) unsigned char gsm[];
    KernelTrace trace_(TR_GATE);
    griddep_wait();
    float* sgate = reinterpret_cast<float*>(gsm);
    uint32_t* shist = reinterpret_cast<uint32_t*>(gsm + round_up((size_t)g.C * sizeof(float), 16));
    const int img = blockIdx.x;
    image_gates(gh, g, img, sgate, shist);
    consumer_barrier();
    griddep_launch_dependents();
    for (int c = threadIdx.x; c < g.C; c += kConsumerWarps * 32) gate[(size_t)img * g.C + c] = sgate[c];
}

template <int Q, int NREG, int CMIN, int CONV>
__global__ void __launch_bounds__(kStreamThreads, 3)
score_pass2_kernel(const float* __restrict__ scores, ScoreGrid g, const float2* __restrict__ rowstat,
                   const float* __restrict__ gate, GateHist gh, int* __restrict__ cand_count, uint2* __restrict__ cand,
                   int cand_cap) {
    extern __shared__ __align__(128) unsigned char smem[];
    KernelTrace trace_(TR_PASS2);
    stream_init(smem);
    griddep_wait();
    if (warp_id() == kConsumerWarps) {
        producer_loop(smem, g, scores,
                      CONV == SSD_CONVERT_SOFTMAX ? reinterpret_cast<const unsigned long long*>(rowstat) : nullptr,
                      policy_evict_first());
        return;
    }
    const RowLanes<Q> ln;
    const RowShape<Q, NREG, CMIN> shape(ln.sub, g.C);
    const int rows_per_warp = g.tile_rows / kConsumerWarps;
    const int Cf = g.C - g.first_fg;
    CandQueue q;
    {
        uint32_t* base = reinterpret_cast<uint32_t*>(smem + stream_smem_bytes_dev(g)) + warp_id() * kQueueCap * 3;
        q.seg = base; q.anchor = base + kQueueCap; q.val = base + 2 * kQueueCap; q.n = 0;
    }
    float* sgate = reinterpret_cast<float*>(smem + stream_smem_bytes_dev(g) + kQueueBytes);      // [C]
    uint32_t* shist = reinterpret_cast<uint32_t*>(smem + stream_smem_bytes_dev(g) + kQueueBytes +
                                                  round_up((size_t)g.C * sizeof(float), 16));     // [C, kGateWords]
    const unsigned lt_mask = (1u << lane_id()) - 1u;

    float gv[NREG];
    int cur_image = -1;
    const uint32_t sb = smem_u32(smem);
    const int wbase = warp_id() * rows_per_warp;
    TileCursor cur;
    cur.start(g);
    for (; cur.valid(g); cur.next(g)) {
        const unsigned r0 = cur.first_row32(g);
        const int rows = cur.rows(g);
        const int img = cur.image(g);
        if (img != cur_image) {                 // per-lane slice of this image's gates
            cur_image = img;
            const float* gsrc = gate + (size_t)img * g.C;
            if (gh.blockmax != nullptr) {
                consumer_barrier();             // everybody is done with the previous image's gates
                image_gates(gh, g, img, sgate, shist);
                consumer_barrier();
                gsrc = sgate;
            }
#pragma unroll
            for (int i = 0; i < NREG; ++i) {
                const int col = ln.sub + i * Q;
                gv[i] = (col < g.C) ? gsrc[col] : INFINITY;
            }
        }
        ring_wait_full(sb, cur);
        const float* logits = ring_logits(smem, g, cur, r0);
        const float2* side = reinterpret_cast<const float2*>(ring_side(smem, g, cur, r0));
        const int a0 = cur.tile * g.tile_rows;
        for (int step = 0; step < rows_per_warp; step += RowLanes<Q>::kRowsPerWarpStep) {
            if (wbase + step >= rows) break;            // the rest of a partial tile (warp-uniform)
            const int lr = wbase + step + ln.rl;
            const bool valid = lr < rows;
            float v[NREG];
            shape.load(v, logits + lr * g.C, ln.sub);
            float t = 0.f;
            if (CONV == SSD_CONVERT_SOFTMAX) {
                const float2 st = side[lr];
                t = __fadd_rn(st.x, fast_log(st.y));            // bit-identical to pass 1
            }
            // bit i of `hit` = sign of gate[i] - x[i], i.e. x[i] > gate[i] (one funnel shift per column collects the sign
            // bits, most significant = column 0).  A NaN or a signed zero can set a bit the comparison would not: the
            // gate only PRUNES, the exact score test of the NMS kernel drops such a candidate again.
            unsigned hit = 0u;
#pragma unroll
            for (int i = 0; i < NREG; ++i) {
                const float x = CONV == SSD_CONVERT_SOFTMAX ? __fsub_rn(v[i], t) : v[i];
                hit = __funnelshift_l(__float_as_uint(__fsub_rn(gv[i], x)), hit, 1);
            }
            if (!valid) hit = 0u;
            hit = __brev(hit) >> (32 - NREG);                    // bit i = column slot i
            // park the survivors: every round each lane with a pending hit adds one entry
            unsigned bal = __ballot_sync(FULL, hit != 0u);
            while (bal) {
                const int add = __popc(bal);
                if (q.n + add > kQueueCap) q.flush(cand_count, cand, cand_cap);
                if (hit) {
                    const int i = __ffs(hit) - 1;
                    hit &= hit - 1;
                    const int col = ln.sub + i * Q;
                    const int pos = q.n + __popc(bal & lt_mask);
                    q.seg[pos] = (uint32_t)(img * Cf + (col - g.first_fg));
                    q.anchor[pos] = (uint32_t)(a0 + lr);
                    q.val[pos] = __float_as_uint(logits[lr * g.C + col]);     // (a dynamic register pick costs NREG selects)
                }
                q.n += add;
                bal = __ballot_sync(FULL, hit != 0u);
            }
        }
        ring_release(sb, cur);
    }
    griddep_launch_dependents();       // late: see the note at launch_pdl
    q.flush(cand_count, cand, cand_cap);
}


// ---------------------------------------------------------------------------------------------
// 1-3 fused: candidate selection of an image that FITS THE SHARED MEMORY OF ONE THREAD-BLOCK CLUSTER
//     (SSD300-VOC: 733 KB per image, SSD-MobileNetV2-COCO: 735 KB -- the launch-bound configurations).
//     One cluster of cs <= 8 CTAs per image; CTA r keeps rows [r R, (r+1) R) of the image's logits in ITS shared
//     memory (TMA bulk copies, a few chunks with one mbarrier each), so the logits leave HBM once and are never
//     re-read through L2, and the three launches of the streaming path (pass 1 -> gates -> pass 2) with their two
//     kernel boundaries become three phases of one kernel separated by cluster barriers:
//       P1  row statistics (max, sum exp) -> rowstat / criterion keys as score_pass1_kernel writes them, the
//           row's log-normaliser kept in shared memory; every lane keeps the running column maxima of the rows
//           IT sees over `merge` steps (a "block" of the row partition) and bumps the CTA's shared-memory
//           histogram of block maxima per column -- no cross-lane reduction, no block-maximum array at all;
//       P2  cluster barrier; column c belongs to one warp of CTA c % cs, which sums the cs histograms of the
//           column through distributed shared memory, finds the bin where the suffix count reaches K (the same
//           lower bound of the K-th largest score as in the streaming path) and stores the gate into every CTA's
//           shared memory; cluster barrier;
//       P3  the slab is walked again out of shared memory, survivors go to the candidate lists.
//     The NMS launch that follows is unchanged.
// ---------------------------------------------------------------------------------------------
constexpr int kFusedWarps = 16;
constexpr int kFusedThreads = kFusedWarps * 32;
constexpr int kFusedChunks = 8;
constexpr size_t kFusedQueueBytes = (size_t)kFusedWarps * kQueueCap * 3 * sizeof(uint32_t);
constexpr int kFusedStageBlocks = 8;           // blocks of the row partition per warp (at most)
constexpr size_t kFusedSmemLimit = 226 * 1024;        // 227 KB opt-in maximum per CTA on sm_100

int fused_select_mode();      // abi.cu: -1 automatic, 0 off, 1/2/4/8 forced cluster size
static int fused_max_clusters(int C, bool softmax, int cs, size_t smem);   // co-resident clusters of the launch (cached)

static void plan_fused(PostPlan& pl) {
    pl.fused_cluster = 0;
    const int mode = fused_select_mode();
    if (mode == 0 || pl.converter == SSD_CONVERT_IDENTITY || pl.A < 1 || pl.B < 1) return;
    const int q = lanes_per_row(pl.C, true);
    const int rps = 32 / q;                                   // rows per warp step
    const size_t hist_bytes = (size_t)pl.C * kGateStride * sizeof(uint32_t);
    // (+ the staging rows of the block maxima: up to 32 / q blocks per warp, see the end of P1)
    const size_t tail = round_up(hist_bytes > kFusedQueueBytes ? hist_bytes : kFusedQueueBytes, 16) +
                        round_up((size_t)pl.C * 3 * sizeof(float), 16) +       // gates, per-class counts and bases
                        round_up((size_t)kFusedWarps * kFusedStageBlocks * (pl.C | 1) * sizeof(float), 16);
    auto smem_for = [&](int cs, int& rows_per_cta) {
        rows_per_cta = (pl.A + cs - 1) / cs;
        rows_per_cta = (rows_per_cta + rps - 1) / rps * rps;
        const size_t slab = round_up(((size_t)rows_per_cta * pl.C + 8 + (size_t)(slots_per_row(pl.C) - pl.C)) * sizeof(float), 128);
        return (size_t)128 + slab + round_up((size_t)rows_per_cta * sizeof(float), 16) + tail;
    };
    int cs = 0, rows = 0;
    for (int c = 1; c <= 8; c <<= 1) {
        int r;
        if (smem_for(c, r) <= kFusedSmemLimit) { cs = c; rows = r; break; }
    }
    if (cs == 0) return;                                      // the image does not fit a cluster: streaming path
    if (mode > 0) {
        int r;
        if (mode < cs || smem_for(mode, r) > kFusedSmemLimit) return;
        cs = mode; rows = r;
    } else {
        // few images: more (smaller) CTAs per image while the grid stays within one wave of SMs
        while (cs < 8 && (long long)pl.B * cs * 2 <= sm_count()) {
            int r;
            smem_for(cs * 2, r);
            if (r < rps * 2) break;
            cs *= 2; rows = r;
        }
    }
    // Blocks of the row partition (any partition works: the K-th largest block maximum bounds the K-th largest
    // score from below).  A lane keeps the column maxima of ALL the rows it sees (steps_per_warp of them); at the
    // end of P1 the 32 / Q row slots of a warp are merged down to `split` blocks of ~A / 2K rows (at most ~32), so
    // that >= ~2K block maxima exist per column.
    const int steps = rows / rps;
    const int steps_per_warp = (steps + kFusedWarps - 1) / kFusedWarps;
    int want_rows = pl.A / (2 * pl.K);
    if (want_rows > 32) want_rows = 32;
    if (want_rows < 1) want_rows = 1;
    int split = 1;
    while (split < rps && split < kFusedStageBlocks && (long long)steps_per_warp * (rps / split) > want_rows) split <<= 1;
    const int merge = steps_per_warp;
    const long long blocks = (long long)cs * kFusedWarps * split;
    if (blocks >= 65536) return;                              // 16-bit histogram counters
    int r2;
    pl.fused_smem = smem_for(cs, r2);
    // automatic mode: only when every image's cluster is resident at once -- a second wave of clusters costs more than
    // the two kernel boundaries of the streaming path save
    if (mode < 0 && pl.B > fused_max_clusters(pl.C, pl.converter == SSD_CONVERT_SOFTMAX, cs, pl.fused_smem)) return;
    pl.fused_cluster = cs; pl.fused_rows_per_cta = rows; pl.fused_merge = merge; pl.fused_split = split;
    const size_t slab = round_up(((size_t)rows * pl.C + 8 + (size_t)(slots_per_row(pl.C) - pl.C)) * sizeof(float), 128);
    pl.fused_off_slab = 128;
    pl.fused_off_trow = (unsigned)(128 + slab);
    pl.fused_off_hist = (unsigned)(pl.fused_off_trow + round_up((size_t)rows * sizeof(float), 16));
    pl.fused_off_gate = (unsigned)(pl.fused_off_hist + round_up(hist_bytes > kFusedQueueBytes ? hist_bytes : kFusedQueueBytes, 16));
    pl.fused_off_stage = (unsigned)(pl.fused_off_gate + round_up((size_t)pl.C * 3 * sizeof(float), 16));
    int chunk = 1024;                                         // floats per bulk copy: a power of two (chunk of a float = a shift)
    while ((size_t)chunk * kFusedChunks < (size_t)rows * pl.C + 4) chunk <<= 1;
    pl.fused_chunk_floats = chunk;
}

// histogram bump in the shared memory of CTA `rank` of the cluster: red.shared::cluster, nothing comes back (a generic
// atomicAdd on the mapped address was observed to cost a blocking round trip per bump)
__device__ __forceinline__ void bump_gate_bin_remote(const uint32_t* words, int rank, float gval, GateBins gb) {
    const int b = gate_bin(gval, gb);
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(words + (b >> 1))), "r"(rank));
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"((b & 1) ? 0x10000u : 1u) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct FusedArgs {
    int A, C, first_fg, K, converter, cand_cap;
    int rows_per_cta, merge, split, chunk_floats, chunk_shift, queue_cap;
    float score_thr;
    GateBins bins;
    long long total_floats;
    unsigned off_slab, off_trow, off_hist, off_gate, off_stage;
};

template <int Q, int NREG, int CMIN, int CONV>
__global__ void __launch_bounds__(kFusedThreads, 1)
fused_select_kernel(const float* __restrict__ scores, FusedArgs fa, float2* __restrict__ rowstat,
                    uint32_t* __restrict__ loss_keys, int* __restrict__ cand_count, uint2* __restrict__ cand,
                    int* __restrict__ score_hist, int* __restrict__ image_done, int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem[];
    KernelTrace trace_(TR_PASS1);
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int img = blockIdx.x / cs;
    const int Cf = fa.C - fa.first_fg;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                          // [kFusedChunks]
    float* slab = reinterpret_cast<float*>(smem + fa.off_slab);
    float* trow = reinterpret_cast<float*>(smem + fa.off_trow);
    uint32_t* shist = reinterpret_cast<uint32_t*>(smem + fa.off_hist);           // [C, kGateStride]; later: the queues
    float* sgate = reinterpret_cast<float*>(smem + fa.off_gate);                 // [C]
    int* s_cnt = reinterpret_cast<int*>(sgate + fa.C);                           // [C] candidates of this CTA per class
    int* s_base = s_cnt + fa.C;                                                  // [C] their first slot in the global list
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < kFusedChunks; ++c) mbar_init(bars + c, 1);
        mbar_fence_init();
    }
    griddep_wait();

    // ---- P0: this CTA's rows -> shared memory ----
    const int row0 = crank * fa.rows_per_cta;
    const int rows = max(0, min(fa.rows_per_cta, fa.A - row0));
    const long long f0 = ((long long)img * fa.A + row0) * fa.C;
    const long long f1 = f0 + (long long)rows * fa.C;
    const long long fal = f0 & ~3ll;                           // 16-byte aligned superset (rows are 4 C bytes)
    const int head = (int)(f0 - fal);
    if (threadIdx.x == 0) {
        long long fe = (f1 + 3) & ~3ll;
        const long long n4 = fa.total_floats & ~3ll;
        if (fe > n4) fe = n4 > fal ? n4 : fal;
        for (long long f = fe; f < f1; ++f) slab[f - fal] = __ldg(scores + f);     // <= 3 floats, array tail only
        const uint64_t policy = policy_evict_first();                             // read once
#pragma unroll 1
        for (int c = 0; c < kFusedChunks; ++c) {
            const long long lo = fal + (long long)c * fa.chunk_floats;
            long long hi = lo + fa.chunk_floats;
            if (hi > fe) hi = fe;
            if (hi > lo) {
                const uint32_t bytes = (uint32_t)((hi - lo) * 4);
                mbar_expect_tx(bars + c, bytes);
                bulk_g2s(slab + (lo - fal), scores + lo, bytes, bars + c, policy);
            } else {
                mbar_arrive(bars + c);
            }
        }
    }
    trace_.mark(0);
    // (while the copies fly) this CTA's histogram -- the OWNED columns receive the whole cluster's bumps -- and counters
    for (int i = threadIdx.x; i < fa.C * kGateStride; i += kFusedThreads) shist[i] = 0u;
    for (int i = threadIdx.x; i < fa.C; i += kFusedThreads) s_cnt[i] = 0;
    __syncthreads();
    cluster_arrive();                      // "my histogram is zeroed"; waited for right before the first remote bump
    // counters / histograms the LATER launches accumulate into, and this image's candidate counts (P3)
    if (crank == 0) {
        for (int i = threadIdx.x; i < Cf; i += kFusedThreads) cand_count[(size_t)img * Cf + i] = 0;
        int4* sh4 = reinterpret_cast<int4*>(score_hist + (size_t)img * kScoreBins);
        for (int i = threadIdx.x; i < kScoreBins / 4; i += kFusedThreads) sh4[i] = make_int4(0, 0, 0, 0);
        if (threadIdx.x == 0) image_done[img] = 0;
        if (blockIdx.x == 0 && threadIdx.x < 4) status[threadIdx.x] = 0;
    }

    const RowLanes<Q> ln;
    const RowShape<Q, NREG, CMIN> shape(ln.sub, fa.C);
    constexpr int RPS = RowLanes<Q>::kRowsPerWarpStep;
    const int steps = (rows + RPS - 1) / RPS;
    const float* base = slab + head;

    // ---- P1: row statistics, criterion keys, histogram of the block maxima ----
    {
        float cmax[NREG];
#pragma unroll
        for (int i = 0; i < NREG; ++i) cmax[i] = -INFINITY;
        int ready = 0;
        for (int s = warp_id(); s < steps; s += kFusedWarps) {
            const int last_row = min((s + 1) * RPS, rows);
            const int need = (head + last_row * fa.C - 1) >> fa.chunk_shift;      // chunk holding the step's last float
            while (ready <= need) { mbar_wait(bars + ready, 0); ++ready; }
            const int lr = s * RPS + ln.rl;
            const bool valid = lr < rows;
            float v[NREG];
            shape.load(v, base + (size_t)lr * fa.C, ln.sub);
            if (CONV == SSD_CONVERT_SOFTMAX) {
                float m, sum;
                row_max_sum<Q, NREG>(v, m, sum);
                const float t = valid ? __fadd_rn(m, fast_log(sum)) : INFINITY;  // see score_pass1_kernel
                if (valid && ln.sub == 0) {
                    const size_t gr = (size_t)img * fa.A + row0 + lr;
                    rowstat[gr] = make_float2(m, sum);
                    trow[lr] = t;
                    if (loss_keys != nullptr) {
                        const uint32_t lk = ordered_key(__fsub_rn(t, v[0]));
                        loss_keys[gr] = lk == 0u ? 1u : lk;
                    }
                }
#pragma unroll
                for (int i = 0; i < NREG; ++i) cmax[i] = fmaxf(cmax[i], __fsub_rn(v[i], t));
            } else {
#pragma unroll
                for (int i = 0; i < NREG; ++i) cmax[i] = fmaxf(cmax[i], valid ? v[i] : -INFINITY);
            }
        }
        // The warp's row slots are merged down to `split` blocks and ONE lane per (block, column) bumps the histogram
        // (lanes that all bumped their own maxima would pile onto the same few bins of a column: 32-way serialised
        // shared-memory atomics).  Warps without a step contribute nothing.
        // End of P1 for this warp.  Its row slots are merged down to `split` <= 8 blocks (shuffles), the block maxima
        // are transposed through a small staging array so that lane c holds COLUMN c of every block, and that lane
        // bumps the histogram of the column -- in the shared memory of the column's OWNER, CTA (c - first_fg) % cs
        // (red.shared::cluster, nothing comes back), so that P2 reads local memory only.  (Every lane bumping its own
        // 21 maxima cost ~900 instructions per warp -- as much as four row steps.)
        trace_.mark(6);
        cluster_wait();                    // every CTA of the cluster has zeroed its histogram
        if (warp_id() < steps) {
            const int cpad = fa.C | 1;
            float* stage = reinterpret_cast<float*>(smem + fa.off_stage) + (size_t)warp_id() * kFusedStageBlocks * cpad;
#pragma unroll
            for (int i = 0; i < NREG; ++i) {
                float x = cmax[i];
#pragma unroll
                for (int o = Q; o < 32; o <<= 1)
                    if (o >= Q * fa.split) x = fmaxf(x, __shfl_xor_sync(FULL, x, o));
                const int col = ln.sub + i * Q;
                if (ln.rl < fa.split && col < fa.C) stage[ln.rl * cpad + col] = x;
            }
            __syncwarp();
            trace_.mark(7);
            for (int col = fa.first_fg + lane_id(); col < fa.C; col += 32) {
                const uint32_t* words = shist + (size_t)col * kGateStride;
                const int owner = (col - fa.first_fg) % cs;
                for (int b = 0; b < fa.split; ++b) bump_gate_bin_remote(words, owner, stage[b * cpad + col], fa.bins);
            }
        }
    }
    trace_.mark(1);
    cluster.sync();
    trace_.mark(2);

    // ---- P2: gates of the OWNED columns (CTA (c - first_fg) % cs, warp ((c - first_fg) / cs) % kFusedWarps) from the
    //          local histogram, stored into every CTA's gate array ----
    {
        const int lane = lane_id();
        for (int j = crank + cs * warp_id(); j < Cf; j += cs * kFusedWarps) {
            const int col = fa.first_fg + j;
            int cnt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cnt[u] = 0;
            {
                const uint32_t* rh = shist + (size_t)col * kGateStride + 4 * lane;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t w = rh[u];
                    cnt[2 * u] = (int)(w & 0xFFFFu);
                    cnt[2 * u + 1] = (int)(w >> 16);
                }
            }
            int own = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) own += cnt[u];
            int incl = own;                               // suffix sum towards higher lanes (= higher bins)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_down_sync(FULL, incl, o);
                if (lane + o < 32) incl += t;
            }
            int above = incl - own;
            int bin = -1;
            if (above < fa.K && incl >= fa.K) {
#pragma unroll
                for (int u = 7; u >= 0; --u) {
                    if (bin < 0 && above + cnt[u] >= fa.K) bin = 8 * lane + u;
                    above += cnt[u];
                }
            }
            const float found = bin >= 0 ? gate_from_kth(fa.converter, fa.score_thr, gate_bin_edge(bin, fa.bins)) : -INFINITY;
            const unsigned who = __ballot_sync(FULL, bin >= 0);
            float gate = __shfl_sync(FULL, found, who ? __ffs(who) - 1 : 0);
            if (who == 0u) gate = gate_from_kth(fa.converter, fa.score_thr, -INFINITY);
            if (lane < cs) cluster.map_shared_rank(sgate, lane)[col] = gate;
        }
        for (int c = threadIdx.x; c < fa.first_fg; c += kFusedThreads) sgate[c] = INFINITY;
    }
    trace_.mark(3);
    cluster.sync();
    trace_.mark(4);

    // ---- P3: candidates (nobody touches this CTA's histogram any more: the queues take its place).  A survivor takes
    //          a CTA-local slot of its class (shared-memory atomic) and is parked in the warp's queue; then ONE global
    //          atomic per (CTA, class) reserves the range in the candidate list -- 2000 same-address returning atomics
    //          per image would serialise at L2 -- and the queues are written out. ----
    {
        CandQueue q;
        uint32_t* qbase = shist + (size_t)warp_id() * fa.queue_cap * 3;
        q.seg = qbase; q.anchor = qbase + fa.queue_cap; q.val = qbase + 2 * fa.queue_cap; q.n = 0;
        const unsigned lt_mask = (1u << lane_id()) - 1u;
        float gv[NREG];
#pragma unroll
        for (int i = 0; i < NREG; ++i) {
            const int col = ln.sub + i * Q;
            gv[i] = col < fa.C ? sgate[col] : INFINITY;
        }
        for (int s = warp_id(); s < steps; s += kFusedWarps) {
            const int lr = s * RPS + ln.rl;
            const bool valid = lr < rows;
            float v[NREG];
            shape.load(v, base + (size_t)lr * fa.C, ln.sub);
            const float t = (CONV == SSD_CONVERT_SOFTMAX && valid) ? trow[lr] : 0.f;
            unsigned hit = 0u;
#pragma unroll
            for (int i = 0; i < NREG; ++i) {
                const float x = CONV == SSD_CONVERT_SOFTMAX ? __fsub_rn(v[i], t) : v[i];
                hit |= (x > gv[i]) ? (1u << i) : 0u;            // -inf padding and gv = +inf never pass
            }
            if (!valid) hit = 0u;
            unsigned bal = __ballot_sync(FULL, hit != 0u);
            while (bal) {
                const int add = __popc(bal);
                const bool direct = q.n + add > fa.queue_cap;   // queue full (heavily tied scores): straight to the list
                if (hit) {
                    const int i = __ffs(hit) - 1;
                    hit &= hit - 1;
                    const int col = ln.sub + i * Q;
                    const uint32_t seg = (uint32_t)(img * Cf + (col - fa.first_fg));
                    const uint32_t raw = __float_as_uint(base[(size_t)lr * fa.C + col]);   // (a dynamic register pick costs NREG selects)
                    if (direct) {
                        const int slot = atomicAdd(&cand_count[seg], 1);
                        if (slot < fa.cand_cap) cand[(size_t)seg * fa.cand_cap + slot] = make_uint2((uint32_t)(row0 + lr), raw);
                    } else {
                        const int pos = q.n + __popc(bal & lt_mask);
                        const int local = atomicAdd(&s_cnt[col], 1);
                        q.seg[pos] = ((uint32_t)col << 20) | (uint32_t)local;          // class column, CTA-local slot
                        q.anchor[pos] = (uint32_t)(row0 + lr);
                        q.val[pos] = raw;
                    }
                }
                if (!direct) q.n += add;
                bal = __ballot_sync(FULL, hit != 0u);
            }
        }
        griddep_launch_dependents();       // late: see the note at launch_pdl
        __syncthreads();
        for (int c = fa.first_fg + threadIdx.x; c < fa.C; c += kFusedThreads) {
            const int n = s_cnt[c];
            s_base[c] = n ? atomicAdd(&cand_count[(size_t)img * Cf + (c - fa.first_fg)], n) : 0;
        }
        __syncthreads();
        for (int e = lane_id(); e < q.n; e += 32) {
            const uint32_t w = q.seg[e];
            const int col = (int)(w >> 20);
            const int slot = s_base[col] + (int)(w & 0xFFFFFu);
            if (slot < fa.cand_cap)
                cand[((size_t)img * Cf + (col - fa.first_fg)) * fa.cand_cap + slot] = make_uint2(q.anchor[e], q.val[e]);
        }
    }
    trace_.mark(5);
}


static int fused_max_clusters(int C, bool softmax, int cs, size_t smem) {
    struct Entry { int C, softmax, cs; size_t smem; int dev, n; };
    static Entry cache[32];
    static int used = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }   // no device: planning only
    for (int i = 0; i < used; ++i)
        if (cache[i].C == C && cache[i].softmax == (int)softmax && cache[i].cs == cs && cache[i].smem == smem && cache[i].dev == dev)
            return cache[i].n;
    int n = 0;
#define SSD_QUERY_FUSED(QQ, NN, CM)                                                                              \
    do {                                                                                                        \
        auto query = [&](auto kern) {                                                                           \
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { \
                (void)cudaGetLastError();                                                                       \
                return;                                                                                         \
            }                                                                                                   \
            cudaLaunchConfig_t cfg = {};                                                                        \
            cfg.gridDim = dim3((unsigned)(cs * 1024));                                                          \
            cfg.blockDim = dim3(kFusedThreads);                                                                 \
            cfg.dynamicSmemBytes = smem;                                                                        \
            cudaLaunchAttribute attr[1];                                                                        \
            attr[0].id = cudaLaunchAttributeClusterDimension;                                                   \
            attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1; \
            cfg.attrs = attr;                                                                                   \
            cfg.numAttrs = 1;                                                                                   \
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = 0; } \
        };                                                                                                      \
        if (softmax) query(fused_select_kernel<QQ, NN, CM, SSD_CONVERT_SOFTMAX>);                               \
        else query(fused_select_kernel<QQ, NN, CM, SSD_CONVERT_IDENTITY>);                                      \
    } while (0)
    SSD_DISPATCH_ROW_SHAPE_PASS1(C, SSD_QUERY_FUSED);
#undef SSD_QUERY_FUSED
    if (used < 32) cache[used++] = Entry{C, (int)softmax, cs, smem, dev, n};
    return n;
}

// ---------------------------------------------------------------------------------------------
// shared-memory bitonic sort of 64-bit keys, descending
// ---------------------------------------------------------------------------------------------
__device__ void bitonic_sort_desc(unsigned long long* keys, int n_pow2) {
    for (int size = 2; size <= n_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// 4. segment_nms: one CTA per (image, foreground class)
// ---------------------------------------------------------------------------------------------
struct NmsArgs {
    int A, C, Cf, first_fg, K, cand_cap, converter, box_input;
    float score_thr, xy_scale, wh_scale;
    double iou_thr;
    // exact division-free form of `(double)RN32(inter / uni) > iou_thr` (see nms_threshold_split)
    double iou_mid;
    int iou_even, exact_mul;
    // fp32 screen (see pair_screen): threshold * (1 -/+ 4e-7), valid for thresholds in [0, 1e30)
    float iou_lo, iou_hi;
    int screen;
    // soft-NMS (box_utils.py:145-163): Gaussian decay instead of the hard sweep
    int soft;
    float soft_thr, soft_sigma;
};

// torchvision compares the fp32 quotient with the double threshold.  Let T32 be the smallest float
// whose value exceeds the threshold and P32 its predecessor: the quotient rounds to >= T32 exactly
// when the real quotient lies above the midpoint M of P32 and T32 (or on it, when T32's mantissa is
// even -- round to nearest even).  M has 25 significant bits, so M * uni is exact in double.
static void nms_threshold_split(double thr, NmsArgs& a) {
    a.iou_thr = thr;
    a.exact_mul = 0; a.iou_mid = 0.0; a.iou_even = 0;
    if (!(thr > 1e-30 && thr < 1e30)) return;
    const float f0 = (float)thr;
    const float t32 = ((double)f0 > thr) ? f0 : nextafterf(f0, INFINITY);
    const float p32 = nextafterf(t32, -INFINITY);
    uint32_t bits;
    memcpy(&bits, &t32, sizeof(bits));
    a.iou_mid = ((double)p32 + (double)t32) * 0.5;
    a.iou_even = (bits & 1u) == 0u;
    a.exact_mul = 1;
}
static void nms_threshold_screen(double thr, NmsArgs& a) {
    a.screen = 0; a.iou_lo = 0.f; a.iou_hi = 0.f;
    if (!(thr >= 0.0 && thr < 1e30)) return;
    a.iou_lo = nextafterf((float)(thr * (1.0 - 4e-7)), -INFINITY);
    a.iou_hi = nextafterf((float)(thr * (1.0 + 4e-7)), INFINITY);
    if (a.iou_lo < 0.f) a.iou_lo = 0.f;
    a.screen = 1;
}

__device__ __forceinline__ float exact_score(int converter, float x, float2 st) {
    if (converter == SSD_CONVERT_SOFTMAX) return __fdiv_rn(expf(__fsub_rn(x, st.x)), st.y);   // exp(x-max)/sum
    if (converter == SSD_CONVERT_SIGMOID) return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x)));     // 1/(1+exp(-x))
    return x;
}

// Exact top-K of one score COLUMN, used when a candidate list overflowed (dense / heavily tied
// scores).  Four strided sweeps over the column: three 11/11/10-bit histogram levels find the K-th
// largest exact score key, the last sweep collects the winners (ties: lower anchor first).
// Slow (strided 4-byte reads, exp + divide per element and sweep) but exact for any input.
__device__ int exact_select_column(const NmsArgs& a, int img, int col, const float* __restrict__ scores,
                                   const float2* __restrict__ rowstat, unsigned long long* keys_out,
                                   uint32_t* hist, int* s_misc) {
    const float* colp = scores + (size_t)img * a.A * a.C + col;
    const float2* rs = rowstat + (size_t)img * a.A;
    auto key_of = [&](int an) -> uint32_t {
        float2 st = make_float2(0.f, 1.f);
        if (a.converter == SSD_CONVERT_SOFTMAX) st = rs[an];
        const float p = exact_score(a.converter, colp[(size_t)an * a.C], st);
        return p > a.score_thr ? ordered_key(p) : 0u;
    };
    uint32_t prefix = 0u;
    int rem = 0, k = 0;
    for (int level = 0; level < 3; ++level) {
        const int shift = level == 0 ? 21 : (level == 1 ? 10 : 0);
        const int bins = level == 2 ? 1024 : 2048;
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0u;
        if (threadIdx.x == 0) s_misc[0] = 0;
        __syncthreads();
        int valid = 0;
        for (int an = threadIdx.x; an < a.A; an += blockDim.x) {
            const uint32_t key = key_of(an);
            if (key == 0u) continue;
            valid++;
            const bool match = level == 0 || (level == 1 ? (key >> 21) == (prefix >> 21) : (key >> 10) == (prefix >> 10));
            if (match) atomicAdd(&hist[(key >> shift) & (uint32_t)(bins - 1)], 1u);
        }
        if (level == 0) {
            valid = __reduce_add_sync(FULL, valid);
            if (lane_id() == 0 && valid) atomicAdd(&s_misc[0], valid);
        }
        __syncthreads();
        if (level == 0) {
            k = min(s_misc[0], a.K);
            rem = k;
            if (k == 0) return 0;
        }
        if (threadIdx.x == 0) {
            int r = rem, b = bins - 1;
            for (; b > 0; --b) {
                if (r <= (int)hist[b]) break;
                r -= (int)hist[b];
            }
            s_misc[1] = b;
            s_misc[2] = r;
            s_misc[3] = (int)hist[b];
        }
        __syncthreads();
        prefix |= (uint32_t)s_misc[1] << shift;
        rem = s_misc[2];
        __syncthreads();
    }
    const uint32_t thr_key = prefix;
    const int need_ties = rem;               // 1 <= need_ties <= number of keys == thr_key
    // collection sweep, anchor order = (round, thread) so that ties go to the lower anchor
    if (threadIdx.x == 0) { s_misc[0] = 0; s_misc[1] = 0; }     // [0] = slots used, [1] = ties seen so far
    __syncthreads();
    int* wcount = s_misc + 4;                                    // [nwarps]
    const int nw = blockDim.x >> 5;
    for (int base = 0; base < a.A; base += blockDim.x) {
        const int an = base + threadIdx.x;
        const uint32_t key = an < a.A ? key_of(an) : 0u;
        const bool tie = key == thr_key && key != 0u;
        const unsigned bal = __ballot_sync(FULL, tie);
        if (lane_id() == 0) wcount[warp_id()] = __popc(bal);
        __syncthreads();
        int before = s_misc[1], total = 0;
        for (int w = 0; w < nw; ++w) {
            if (w < warp_id()) before += wcount[w];
            total += wcount[w];
        }
        const int rank = before + __popc(bal & ((1u << lane_id()) - 1u));
        if (key > thr_key || (tie && rank < need_ties)) {
            const int slot = atomicAdd(&s_misc[0], 1);
            keys_out[slot] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)an);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_misc[1] += total;
        __syncthreads();
    }
    return k;
}

// torchvision's overlap test `(double)fp32(inter / union) > threshold` for one pair of corner boxes
// (areas unclamped, intersection sides clamped at 0), in three tiers:
//   * disjoint boxes (inter == 0) never exceed a non-negative threshold (0/u, -0 or 0/0 -> NaN);
//   * fp32 screen: inter against union * threshold with a 4e-7 relative guard band on both sides
//     (the product is off by <= 2^-24 relative, the rounding midpoint by <= 2^-24 more);
//   * anything inside the band, non-positive unions, NaNs and negative thresholds take the exact
//     form: the division-free double comparison (see nms_threshold_split) or the literal division.
// Returns 0 = keep, 1 = suppress, 2 = undecided (exact form needed).
__device__ __forceinline__ int pair_screen(const NmsArgs& a, float4 bi, float ai, float4 bj, float aj) {
    const float iw = fmaxf(0.f, fsub(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
    const float ih = fmaxf(0.f, fsub(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
    const float inter = fmul(iw, ih);
    if (inter == 0.f) return 0;
    const float uni = fsub(fadd(ai, aj), inter);
    const bool ok = inter > 0.f && uni > 0.f;
    if (ok && inter > fmul(uni, a.iou_hi)) return 1;
    if (ok && inter < fmul(uni, a.iou_lo)) return 0;
    return 2;
}
__device__ __noinline__ bool pair_suppressed_exact(const NmsArgs& a, float4 bi, float ai, float4 bj, float aj) {
    const float iw = fmaxf(0.f, fsub(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
    const float ih = fmaxf(0.f, fsub(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
    const float inter = fmul(iw, ih);
    const float uni = fsub(fadd(ai, aj), inter);
    if (a.exact_mul && inter > 0.f && uni > 0.f) {
        const double lhs = (double)inter, rhs = a.iou_mid * (double)uni;
        return lhs > rhs || (lhs == rhs && a.iou_even);
    }
    return (double)fdiv(inter, uni) > a.iou_thr;           // float-vs-double compare
}
__device__ __forceinline__ bool pair_suppressed(const NmsArgs& a, float4 bi, float ai, float4 bj, float aj, bool active) {
    int r = 2;
    if (a.screen) r = active ? pair_screen(a, bi, ai, bj, aj) : 0;
    else if (!active) r = 0;
    if (__any_sync(FULL, r == 2)) {
        if (r == 2) r = pair_suppressed_exact(a, bi, ai, bj, aj) ? 1 : 0;
    }
    return r == 1;
}

constexpr int kRankSortMax = 256;     // candidate lists up to this size are sorted by ranking
// 64-bit sort slots at the head of the NMS kernel's shared memory; later reused as the per-warp pair
// lists of the overlap filter (NT / 32 warps x 1024 16-bit codes), hence at least 1024
__host__ __device__ inline int nms_key_slots(int cand_cap, int nt) {
    const int lo = kMaxPerClass > 256 * (nt / 32) ? kMaxPerClass : 256 * (nt / 32);
    return cand_cap > lo ? cand_cap : lo;
}

// ---------------------------------------------------------------------------------------------
// 5. final top-k of one image (postprocessor.py:68-74), run by the LAST segment CTA of the image
//    (or by image_topk_kernel when its shared memory does not fit next to the NMS arrays)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_det_row(float* dets, int* anchors, const float* kept, int K, int cls, int slot,
                                              int out_row) {
    const float* src = kept + ((size_t)cls * K + slot) * kKeptCols;
    float* o = dets + (size_t)out_row * 6;
    o[0] = src[0]; o[1] = src[1]; o[2] = src[2]; o[3] = src[3];
    o[4] = (float)(cls + 1);                                           // postprocessor.py:66
    o[5] = src[4];
    if (anchors != nullptr) anchors[out_row] = (int)__float_as_uint(src[5]);
}

template <int THREADS>
struct TopkShared {
    int wsum[THREADS / 32];
    int warp_tot[THREADS / 32];
    int cnt[4];
    int n_total, n_sel, n_cand, cut_bin, above, in_bin;
};

struct TopkArgs {
    int Cf, K, T, det_cap;
    const int* kept_count;
    const float* kept;
    const int* score_hist;
    float* dets;
    int* det_count;
    int* det_anchor;
};

__host__ __device__ inline size_t topk_smem_bytes(int Cf, int T) {
    int t2 = 32;
    while (t2 < T) t2 <<= 1;
    return round_up((size_t)(Cf + 1) * 4, 16) * 2 + (size_t)(T > 0 ? t2 : 0) * 8 + (size_t)kTopkBoundaryCap * 8 + 64;
}

// The T best of the image's kept rows, descending score, ties by class-major position.  The NMS
// CTAs have already histogrammed the kept scores, so the cut is found with one suffix scan; rows
// above the cut bin are taken, the cut bin is ranked exactly.  A cut bin too crowded to rank
// (scores tied en masse) falls back to a bit-wise bisection over the rows in global memory.
template <int THREADS>
__device__ void image_topk_body(unsigned char* smem, TopkShared<THREADS>& sh, const TopkArgs& ta, int img) {
    const int lane = lane_id();
    const int tid = threadIdx.x;
    constexpr int nwarps = THREADS / 32;
    constexpr int BPT = kScoreBins / THREADS;                  // histogram bins per thread
    static_assert(BPT % 4 == 0 && BPT * THREADS == kScoreBins, "whole int4 loads per thread");
    const int Cf = ta.Cf, K = ta.K, T = ta.T;
    int t2 = 32;
    while (t2 < T) t2 <<= 1;
    const size_t offs_bytes = round_up((size_t)(Cf + 1) * 4, 16);
    int* offs = reinterpret_cast<int*>(smem);                                            // [Cf + 1]
    int* tie_base = reinterpret_cast<int*>(smem + offs_bytes);                           // [Cf + 1] fallback only
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(smem + 2 * offs_bytes);   // [t2]
    unsigned long long* cand = sel + (T > 0 ? t2 : 0);                                   // [kTopkBoundaryCap] (later: sorted)
    const int* kc = ta.kept_count + (size_t)img * Cf;
    const float* kimg = ta.kept + (size_t)img * Cf * K * kKeptCols;
    float* dimg = ta.dets + (size_t)img * ta.det_cap * 6;
    int* aimg = ta.det_anchor ? ta.det_anchor + (size_t)img * ta.det_cap : nullptr;

    // the histogram loads go out first (higher bins = larger scores)
    int hb[BPT];
    if (T > 0) {
        const int4* h4 = reinterpret_cast<const int4*>(ta.score_hist + (size_t)img * kScoreBins) + (BPT / 4) * tid;
#pragma unroll
        for (int j = 0; j < BPT / 4; ++j) {
            const int4 q = __ldcg(h4 + j);
            hb[4 * j] = q.x; hb[4 * j + 1] = q.y; hb[4 * j + 2] = q.z; hb[4 * j + 3] = q.w;
        }
    }
    // exclusive scan of the per-class counts (warp 0, 32 classes per round)
    if (warp_id() == 0) {
        int base = 0;
        for (int c0 = 0; c0 < Cf; c0 += 32) {
            const int c = c0 + lane;
            const int v = c < Cf ? __ldcg(kc + c) : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (c < Cf) offs[c] = base + incl - v;
            base += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) {
            offs[Cf] = base;
            sh.n_total = base; sh.n_sel = 0; sh.n_cand = 0; sh.cut_bin = -1; sh.above = 0; sh.in_bin = 0;
            sh.cnt[0] = sh.cnt[1] = sh.cnt[2] = sh.cnt[3] = 0;
        }
    }
    __syncthreads();
    const int n = sh.n_total;

    if (T <= 0 || n <= T) {
        // class-major order, descending score inside a class            postprocessor.py:68-70
        for (int c = warp_id(); c < Cf; c += nwarps) {
            const int cnt = offs[c + 1] - offs[c];
            for (int t = lane; t < cnt; t += 32) write_det_row(dimg, aimg, kimg, K, c, t, offs[c] + t);
        }
        if (tid == 0) ta.det_count[img] = n;
        return;
    }

    // ---- cut bin: above(bin) < T <= above(bin) + hist[bin] ----
    {
        int own = 0;
#pragma unroll
        for (int j = 0; j < BPT; ++j) own += hb[j];
        int incl = own;                                   // suffix sum inside the warp (towards higher lanes)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_down_sync(FULL, incl, o);
            if (lane + o < 32) incl += t;
        }
        if (lane == 0) sh.warp_tot[warp_id()] = incl;
        __syncthreads();
        const int wt = lane < nwarps ? sh.warp_tot[lane] : 0;
        const int higher = __reduce_add_sync(FULL, lane > warp_id() ? wt : 0);
        int above = higher + incl - own;
        if (above < T && above + own >= T) {
#pragma unroll
            for (int j = BPT - 1; j >= 0; --j) {
                if (above < T && above + hb[j] >= T) { sh.cut_bin = BPT * tid + j; sh.above = above; sh.in_bin = hb[j]; }
                above += hb[j];
            }
        }
        __syncthreads();
    }
    const int cut_bin = sh.cut_bin, in_bin = sh.in_bin;
    const int need = T - sh.above;                        // rows wanted from the cut bin, 1..in_bin

    if (in_bin <= kTopkBoundaryCap) {
        // one pass over the kept rows: composite = (score key, ~class-major position).  Flat over the
        // Cf x K slots, four independent loads in flight per thread.
        const int slots = Cf * K;
        for (int s0 = tid; s0 < slots; s0 += 4 * THREADS) {
            float sc[4];
            int pos[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int s = s0 + u * THREADS;
                const int c = s < slots ? s / K : 0;
                const int t = s - c * K;
                const bool ok = s < slots && t < offs[c + 1] - offs[c];
                pos[u] = ok ? offs[c] + t : -1;
                sc[u] = ok ? kimg[(size_t)s * kKeptCols + 4] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (pos[u] < 0) continue;
                const int bin = score_bin(sc[u]);
                if (bin >= cut_bin) {
                    const unsigned long long comp = ((unsigned long long)ordered_key(sc[u]) << 32) |
                                                    (unsigned long long)(0xFFFFFFFFu - (uint32_t)pos[u]);
                    if (bin > cut_bin || need == in_bin) sel[atomicAdd(&sh.n_sel, 1)] = comp;
                    else cand[atomicAdd(&sh.n_cand, 1)] = comp;
                }
            }
        }
        __syncthreads();
        if (need < in_bin) {
            const int nc = sh.n_cand;
            for (int i = tid; i < nc; i += THREADS) {
                const unsigned long long me = cand[i];
                int rank = 0;
                for (int j = 0; j < nc; ++j) rank += cand[j] > me;
                if (rank < need) sel[atomicAdd(&sh.n_sel, 1)] = me;
            }
            __syncthreads();
        }
    } else {
        // crowded cut bin: S = the T-th largest score key by bit-wise bisection over the rows (read
        // from global memory every pass -- slow, but only for en-masse ties).  One barrier per bit;
        // the counters rotate over four slots (slot p+2 is cleared while p is in use).
        auto count_keys = [&](uint32_t trial, bool strict) {
            int c_ = 0;
            for (int c = warp_id(); c < Cf; c += nwarps) {
                const int cnt = offs[c + 1] - offs[c];
                for (int t = lane; t < cnt; t += 32) {
                    const uint32_t key = ordered_key(kimg[((size_t)c * K + t) * kKeptCols + 4]);
                    c_ += strict ? key > trial : key >= trial;
                }
            }
            return __reduce_add_sync(FULL, c_);
        };
        uint32_t S = 0u;
        for (int bit = 31, pass = 0; bit >= 0; --bit, ++pass) {
            const uint32_t trial = S | (1u << bit);
            const int c = count_keys(trial, false);
            if (lane == 0 && c) atomicAdd(&sh.cnt[pass & 3], c);
            if (tid == 0) sh.cnt[(pass + 2) & 3] = 0;
            __syncthreads();
            if (sh.cnt[pass & 3] >= T) S = trial;
        }
        const int gt = count_keys(S, true);
        __syncthreads();                                   // every thread is done reading cnt[]
        if (tid == 0) sh.cnt[0] = 0;
        __syncthreads();
        if (lane == 0 && gt) atomicAdd(&sh.cnt[0], gt);
        // ties per class, then their exclusive scan: ties are taken from the front (position order)
        for (int c = warp_id(); c < Cf; c += nwarps) {
            const int cnt = offs[c + 1] - offs[c];
            int ties = 0;
            for (int t = lane; t < cnt; t += 32) ties += ordered_key(kimg[((size_t)c * K + t) * kKeptCols + 4]) == S;
            ties = __reduce_add_sync(FULL, ties);
            if (lane == 0) tie_base[c] = ties;
        }
        __syncthreads();
        if (warp_id() == 0) {
            int base = 0;
            for (int c0 = 0; c0 < Cf; c0 += 32) {
                const int c = c0 + lane;
                const int v = c < Cf ? tie_base[c] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                if (c < Cf) tie_base[c] = base + incl - v;
                base += __shfl_sync(FULL, incl, 31);
            }
        }
        __syncthreads();
        const int need_ties = T - sh.cnt[0];               // >= 1
        for (int c = warp_id(); c < Cf; c += nwarps) {
            const int cnt = offs[c + 1] - offs[c];
            int running = tie_base[c];
            for (int t0 = 0; t0 < cnt; t0 += 32) {
                const int t = t0 + lane;
                const uint32_t key = t < cnt ? ordered_key(kimg[((size_t)c * K + t) * kKeptCols + 4]) : 0u;
                const bool tie = t < cnt && key == S;
                const unsigned bal = __ballot_sync(FULL, tie);
                const int trank = running + __popc(bal & ((1u << lane) - 1u));
                if (t < cnt && (key > S || (tie && trank < need_ties)))
                    sel[atomicAdd(&sh.n_sel, 1)] = ((unsigned long long)key << 32) |
                                                   (unsigned long long)(0xFFFFFFFFu - (uint32_t)(offs[c] + t));
                running += __popc(bal);
            }
        }
        __syncthreads();
    }

    // ---- order the T selected rows: descending score, ties by class-major position ----
    const unsigned long long* ordered;
    if (T <= kTopkBoundaryCap) {
        unsigned long long* sorted = cand;                 // the candidates are no longer needed
        for (int r = tid; r < T; r += THREADS) {
            const unsigned long long me = sel[r];
            int rank = 0;
            for (int j = 0; j < T; ++j) rank += sel[j] > me;
            sorted[rank] = me;
        }
        ordered = sorted;
        __syncthreads();
    } else {
        for (int i = T + tid; i < t2; i += THREADS) sel[i] = 0ull;
        bitonic_sort_desc(sel, t2);
        ordered = sel;
    }
    for (int r = tid; r < T; r += THREADS) {
        const int pos = (int)(0xFFFFFFFFu - (uint32_t)(ordered[r] & 0xFFFFFFFFull));
        int c_lo = 0, c_hi = Cf;               // largest c with offs[c] <= pos
        while (c_hi - c_lo > 1) {
            const int mid = (c_lo + c_hi) >> 1;
            if (offs[mid] <= pos) c_lo = mid; else c_hi = mid;
        }
        write_det_row(dimg, aimg, kimg, K, c_lo, pos - offs[c_lo], r);
    }
    if (tid == 0) ta.det_count[img] = T;
}

__global__ void __launch_bounds__(kTopkThreads)
image_topk_kernel(TopkArgs ta) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ TopkShared<kTopkThreads> sh;
    KernelTrace trace_(TR_TOPK);
    griddep_wait();
    griddep_launch_dependents();
    image_topk_body<kTopkThreads>(smem, sh, ta, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// 4. segment_nms: one CTA per (image, foreground class)
// ---------------------------------------------------------------------------------------------
// Overlap filter of the bit matrix: bit U of `bits` is set when the two boxes intersect with positive extent,
// bi.z > bj.x && bj.z > bi.x && bi.w > bj.y && bj.w > bi.y (NaN fails) -- four chained FSETP and ONE predicated OR
// (the compiler's form of `bits |= hit ? 1u << U : 0u` was MOV 0 + SEL + LOP3 per pair).
template <int U>
__device__ __forceinline__ void overlap_bit(uint32_t& bits, const float4& bi, const float4 bj) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.f32 p, %1, %2;\n\t"
        "setp.gt.and.f32 p, %3, %4, p;\n\t"
        "setp.gt.and.f32 p, %5, %6, p;\n\t"
        "setp.gt.and.f32 p, %7, %8, p;\n\t"
        "@p or.b32 %0, %0, %9;\n\t}"
        : "+r"(bits)
        : "f"(bi.z), "f"(bj.x), "f"(bj.z), "f"(bi.x), "f"(bi.w), "f"(bj.y), "f"(bj.w), "f"(bi.y), "n"(1u << U));
}

// Returns the number of kept rows of the segment (kept_count[seg] is written by the caller).
template <int NT>
__device__ int nms_segment(const KernelTrace& tr, const NmsArgs& a, unsigned char* smem, uint32_t* s_hist, int* s_misc, int& s_valid, int& s_nkeep,
                           const float* __restrict__ scores, const float2* __restrict__ rowstat,
                           const int* __restrict__ cand_count, const uint2* __restrict__ cand,
                           const float4* __restrict__ boxes, const float4* __restrict__ priors, float* __restrict__ kept,
                           int* __restrict__ status, int* __restrict__ score_hist) {
    const int seg = blockIdx.x;
    const int img = seg / a.Cf;
    const int lane = lane_id();
    tr.mark(0);
    // the first candidate of every thread is requested together with the count (one round trip less)
    const uint2 first_cand = cand[(size_t)seg * a.cand_cap + threadIdx.x];
    int n_raw = cand_count[seg];
    if (n_raw == 0) return 0;
    // carve: keys[key_slots] u64 | sorted[K] u64 | box[K] float4 | fbox[K] float4 | area[K] | mask[K * kwords] | keep[K]
    const int key_slots = nms_key_slots(a.cand_cap, NT);
    const int kwords = (a.K + 31) >> 5;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    unsigned long long* sorted = keys + key_slots;
    float4* sbox = reinterpret_cast<float4*>(sorted + ((a.K + 1) & ~1));          // 16-byte aligned
    float4* fbox = sbox + a.K;                 // the boxes as the overlap filter sees them (NaN if degenerate)
    float* sarea = reinterpret_cast<float*>(fbox + a.K);
    uint32_t* mask = reinterpret_cast<uint32_t*>(sarea + a.K);
    int* keep = reinterpret_cast<int*>(mask + (size_t)a.K * kwords);
    // 32-bit score words of the rank sort [kRankSortMax + 4]: the rank sort touches only the first kRankSortMax of the
    // >= 1024 key slots, the words live behind them
    static_assert(kRankSortMax <= 512, "hk sits at key slot 512");
    uint32_t* hk = reinterpret_cast<uint32_t*>(keys + 512);

    const bool overflow = n_raw > a.cand_cap;
    if (threadIdx.x == 0) s_valid = 0;
    if (threadIdx.x == 0 && status != nullptr) {
        if (n_raw > kRankSortMax) atomicAdd(&status[2], 1);
        atomicMax(&status[3], n_raw);
    }
    if (overflow) {
        // the candidate list is incomplete: redo this (image, class) exactly from the score column
        if (threadIdx.x == 0 && status != nullptr) atomicAdd(&status[1], 1);
        n_raw = exact_select_column(a, img, a.first_fg + (seg - img * a.Cf), scores, rowstat, keys, s_hist, s_misc);
        if (n_raw == 0) return 0;
        if (threadIdx.x == 0) s_valid = n_raw;
    }
    __syncthreads();

    // ---- exact scores: key = (ordered score, ~anchor), 0 = below the threshold ----
    const bool by_rank = n_raw <= kRankSortMax;
    int n2 = 32;
    while (n2 < n_raw) n2 <<= 1;
    const int fill = by_rank ? ((n_raw + 1) & ~1) : n2;
    if (!overflow) {
        int local_valid = 0;
        for (int t = threadIdx.x; t < fill; t += blockDim.x) {
            unsigned long long key = 0ull;
            if (t < n_raw) {
                const uint2 e = t == (int)threadIdx.x ? first_cand : cand[(size_t)seg * a.cand_cap + t];
                float2 st = make_float2(0.f, 1.f);
                if (a.converter == SSD_CONVERT_SOFTMAX) st = rowstat[(size_t)img * a.A + e.x];
                const float p = exact_score(a.converter, __uint_as_float(e.y), st);
                if (p > a.score_thr) {                                   // postprocessor.py:62 (fp32 compare)
                    key = ((unsigned long long)ordered_key(p) << 32) | (unsigned long long)(0xFFFFFFFFu - e.x);
                    local_valid++;
                }
            }
            keys[t] = key;
        }
        local_valid = __reduce_add_sync(FULL, local_valid);
        if (lane == 0 && local_valid) atomicAdd(&s_valid, local_valid);
    } else {
        for (int t = n_raw + threadIdx.x; t < fill; t += blockDim.x) keys[t] = 0ull;
    }
    __syncthreads();
    tr.mark(1);
    const int n = min(s_valid, a.K);             // box_utils.py:186-188 top-k
    if (n == 0) return 0;

    // ---- order: (score desc, anchor asc) ----
    if (by_rank) {
        // Keys are unique, so the rank of a key is the number of larger keys.  Fast path on the 32-bit SCORE
        // word alone: three instructions per key, four keys per LDS.128, two independent counters (the 64-bit
        // compare-and-add it replaces was 18 % of the kernel's instructions).  Two candidates
        // with the SAME score word (rare: equal fp32 scores) get the same rank; the collision is detected when
        // a key does not find itself in its slot, and the exact 64-bit ranking (ties: lower anchor first) runs.
        // The words are the top 31 bits of the score key (an invalid entry is 0): both operands below 2^31, so the sign
        // bit of mine - other says `other > mine` and one shift-and-add (LEA.HI) accumulates it -- two instructions per
        // key instead of compare + add + select.  Dropping the lowest key bit only makes a collision (adjacent fp32
        // scores) marginally more likely; collisions are detected below.
        const int fill4 = (n_raw + 3) & ~3;
        for (int t = threadIdx.x; t < fill4; t += blockDim.x) hk[t] = t < n_raw ? (uint32_t)(keys[t] >> 33) : 0u;
        __syncthreads();
        const uint4* h4 = reinterpret_cast<const uint4*>(hk);
        constexpr int kPerThread = (kRankSortMax + NT - 1) / NT;
        int myrank[kPerThread];
#pragma unroll
        for (int u = 0; u < kPerThread; ++u) {
            const int t = threadIdx.x + u * NT;
            myrank[u] = -1;
            if (t >= n_raw) continue;
            const unsigned long long me = keys[t];
            if (me == 0ull) continue;
            const uint32_t mine = (uint32_t)(me >> 33);
            uint32_t rank = 0u, rank_b = 0u;
#pragma unroll 4
            for (int j = 0; j < (fill4 >> 2); ++j) {
                const uint4 o = h4[j];
                rank += (mine - o.x) >> 31;
                rank_b += (mine - o.y) >> 31;
                rank += (mine - o.z) >> 31;
                rank_b += (mine - o.w) >> 31;
            }
            rank += rank_b;
            if ((int)rank < n) { sorted[rank] = me; myrank[u] = (int)rank; }
        }
        __syncthreads();
        bool bad = false;
#pragma unroll
        for (int u = 0; u < kPerThread; ++u)          // a key that is not in its own slot lost it to a tied score word
            if (myrank[u] >= 0) bad = bad || sorted[myrank[u]] != keys[threadIdx.x + u * NT];
        if (__syncthreads_or(bad)) {
            // exact ranking on the full 64-bit keys
            const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(keys);
            for (int t = threadIdx.x; t < n_raw; t += blockDim.x) {
                const unsigned long long me = keys[t];
                if (me == 0ull) continue;
                int rank = 0;
#pragma unroll 4
                for (int j = 0; j < (fill >> 1); ++j) {
                    const ulonglong2 o = k2[j];
                    rank += (o.x > me) + (o.y > me);
                }
                if (rank < n) sorted[rank] = me;
            }
            __syncthreads();
        }
    } else {
        bitonic_sort_desc(keys, n2);                 // starts and ends with a barrier
        for (int t = threadIdx.x; t < n; t += blockDim.x) sorted[t] = keys[t];
        __syncthreads();
    }
    const int words = (n + 31) >> 5;
    tr.mark(2);

    // ---- boxes of the n best: decode + to_corners only these; the diagonal mask words start at 0 ----
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const uint32_t anchor = 0xFFFFFFFFu - (uint32_t)(sorted[t] & 0xFFFFFFFFull);
        float4 bx = boxes[(size_t)img * a.A + anchor];
        if (a.box_input == SSD_BOXES_ENCODED) {
            const float4 p = priors[anchor];
            // box_coder.py:55-57 then box_utils.py:23
            const float cx = fadd(p.x, fdiv(fmul(p.z, bx.x), a.xy_scale));
            const float cy = fadd(p.y, fdiv(fmul(p.w, bx.y), a.xy_scale));
            const float w = fmul(p.z, expf(fdiv(bx.z, a.wh_scale)));
            const float h = fmul(p.w, expf(fdiv(bx.w, a.wh_scale)));
            const float hw = fmul(w, 0.5f), hh = fmul(h, 0.5f);
            bx = make_float4(fsub(cx, hw), fsub(cy, hh), fadd(cx, hw), fadd(cy, hh));
        }
        sbox[t] = bx;
        sarea[t] = a.soft ? fmul(fmaxf(fsub(bx.z, bx.x), 0.f), fmaxf(fsub(bx.w, bx.y), 0.f))   // box_utils.area: clamped
                          : fmul(fsub(bx.z, bx.x), fsub(bx.w, bx.y));                          // torchvision: unclamped
        // a box without positive extent intersects nothing (the clamped side is 0): NaN fails every compare
        const bool extent = bx.z > bx.x && bx.w > bx.y;
        fbox[t] = extent ? bx : make_float4(NAN, NAN, NAN, NAN);
        for (int w = 0; w < words; ++w) mask[(size_t)t * words + w] = 0u;
    }
    __syncthreads();
    tr.mark(3);

    if (a.soft) {
        // ---- soft-NMS, box_utils.py:145-163, statement by statement (warp 0; lanes = candidates):
        //        mask = scores > thr
        //        while mask.nonzero().sum():                 <- the SUM OF THE INDICES: a mask whose only
        //            idx = scores_copy.argmax()                 element is index 0 ends the loop
        //            scores_copy[idx] = 0; picked += idx
        //            mask = scores_copy > thr                <- tested by the NEXT loop head, before ...
        //            scores_copy[mask] *= exp(-(iou(idx, mask) ** 2) / sigma)      ... this decay
        //      Index 0 of the reference's subset is the lowest anchor (boolean-mask order); ties of
        //      the argmax go to the lowest anchor as well. ----
        float* sc = reinterpret_cast<float*>(mask);
        for (int t = threadIdx.x; t < n; t += blockDim.x) sc[t] = key_to_float((uint32_t)(sorted[t] >> 32));
        __syncthreads();
        if (warp_id() == 0) {
            unsigned long long lowest = ~0ull;
            for (int t = lane; t < n; t += 32) {
                const unsigned long long anchor = 0xFFFFFFFFull - (sorted[t] & 0xFFFFFFFFull);
                lowest = min(lowest, (anchor << 32) | (unsigned long long)t);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lowest = min(lowest, __shfl_xor_sync(FULL, lowest, o));
            const int p0 = (int)(lowest & 0xFFFFFFFFull);
            bool more = false;
            for (int t = lane; t < n; t += 32) more |= sc[t] > a.soft_thr && t != p0;
            more = __any_sync(FULL, more);
            int nkeep = 0;
            while (more && nkeep < n) {
                // argmax over (score, lower anchor first): key = (ordered score, ~anchor) like `sorted`
                unsigned long long best = 0ull;
                int bt = 0;
                for (int t = lane; t < n; t += 32) {
                    const unsigned long long k = ((unsigned long long)ordered_key(sc[t]) << 32) | (sorted[t] & 0xFFFFFFFFull);
                    if (k > best) { best = k; bt = t; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long ob = __shfl_xor_sync(FULL, best, o);
                    const int ot = __shfl_xor_sync(FULL, bt, o);
                    if (ob > best) { best = ob; bt = ot; }
                }
                if (lane == 0) { keep[nkeep] = bt; sc[bt] = 0.f; }
                ++nkeep;
                __syncwarp();
                const float4 bi = sbox[bt];
                const float ai = sarea[bt];
                more = false;
                for (int t = lane; t < n; t += 32) {
                    const float v = sc[t];
                    if (v > a.soft_thr) {
                        more |= t != p0;
                        const float4 bj = sbox[t];
                        const float iw = fmaxf(fsub(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)), 0.f);
                        const float ih = fmaxf(fsub(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)), 0.f);
                        const float inter = fmul(iw, ih);
                        const float iou = fdiv(inter, fsub(fadd(ai, sarea[t]), inter));
                        sc[t] = fmul(v, expf(-fdiv(fmul(iou, iou), a.soft_sigma)));
                    }
                }
                more = __any_sync(FULL, more);
                __syncwarp();
            }
            if (lane == 0) s_nkeep = nkeep;
        }
        __syncthreads();
    } else {
    // ---- IoU bit matrix: bit j of row i set <=> box i suppresses box j (j > i) ----
    // Two tiers per 32 x 32 block (row chunk c, column word w >= c), one warp per block:
    //  (1) overlap filter, lanes = rows: the 32 boxes of the word are broadcast one by one and every
    //      lane collects a bit per column whose box intersects its row's box with positive extent
    //      (four compares).  A pair that fails this has intersection 0 (or NaN) and can never exceed
    //      a non-negative threshold -- for random boxes that is ~90 % of the pairs;
    //  (2) the surviving pairs are compacted into a per-warp list (16-bit row/column codes in the
    //      shared memory the sort no longer needs) and only those take the IoU test, lanes = pairs,
    //      every lane busy; suppressions are recorded with a shared-memory atomicOr.
    const int nwarps = blockDim.x >> 5;
    unsigned short* plist = reinterpret_cast<unsigned short*>(keys) + (size_t)warp_id() * 1024;     // [1024] per warp
    const int nblocks = words * (words + 1) / 2;
    for (int blk = warp_id(); blk < nblocks; blk += nwarps) {
        // blk -> (c, w) with c <= w, enumerated word-major: w = 0: (0,0); w = 1: (0,1), (1,1); ...
        int w = 0;
        while ((w + 1) * (w + 2) / 2 <= blk) ++w;
        const int c = blk - w * (w + 1) / 2;
        const int i = (c << 5) + lane;                      // this lane's row
        const int j0 = w << 5;
        const int cols = min(32, n - j0);
        uint32_t bits = 0u;
        if (a.screen) {
            const float4 bi = i < n ? fbox[i] : make_float4(NAN, NAN, NAN, NAN);
            // whole groups of eight columns with compile-time bit positions (a runtime `1u << jj` costs a MOV + SHF per
            // pair on top of the four compares), then the <= 7 columns that are left
            int jb = 0;
            for (; jb + 8 <= cols; jb += 8) {
                uint32_t g8 = 0u;
                overlap_bit<0>(g8, bi, fbox[j0 + jb + 0]);      // same address for every lane: broadcast
                overlap_bit<1>(g8, bi, fbox[j0 + jb + 1]);
                overlap_bit<2>(g8, bi, fbox[j0 + jb + 2]);
                overlap_bit<3>(g8, bi, fbox[j0 + jb + 3]);
                overlap_bit<4>(g8, bi, fbox[j0 + jb + 4]);
                overlap_bit<5>(g8, bi, fbox[j0 + jb + 5]);
                overlap_bit<6>(g8, bi, fbox[j0 + jb + 6]);
                overlap_bit<7>(g8, bi, fbox[j0 + jb + 7]);
                bits |= g8 << jb;
            }
            for (int jj = jb; jj < cols; ++jj) {
                const float4 bj = fbox[j0 + jj];
                const bool hit = bi.z > bj.x && bj.z > bi.x && bi.w > bj.y && bj.w > bi.y;
                bits |= hit ? 1u << jj : 0u;
            }
        } else {
            bits = i < n ? (cols == 32 ? ~0u : (1u << cols) - 1u) : 0u;      // negative threshold: every pair counts
        }
        if (c == w) bits &= lane == 31 ? 0u : ~0u << (lane + 1);             // diagonal block: only j > i
        // compaction: exclusive prefix of the popcounts, then every lane writes its (row, column) codes
        const int mine = __popc(bits);
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        int pos = incl - mine;
        while (bits) {
            const int jj = __ffs(bits) - 1;
            bits &= bits - 1;
            plist[pos++] = (unsigned short)((lane << 5) | jj);
        }
        __syncwarp();
        for (int e0 = 0; e0 < total; e0 += 32) {
            const int e = e0 + lane;
            const bool active = e < total;
            const int code = active ? plist[e] : 0;
            const int pi = (c << 5) + (code >> 5), pj = j0 + (code & 31);
            const bool sup = pair_suppressed(a, sbox[pi], sarea[pi], sbox[pj], sarea[pj], active);
            if (sup) atomicOr(&mask[(size_t)pi * words + w], 1u << (code & 31));
        }
        __syncwarp();
    }
    __syncthreads();
    tr.mark(4);

    // ---- greedy sweep in score order (warp 0), 32 rows at a time, lanes = rows of the chunk.
    //      Lane w also owns word w of the suppressed set.  Only rows that suppress something inside
    //      the chunk take part in the sequential chain (a row with an empty diagonal word changes
    //      nothing; whether it is kept can be read off the final set, since later rows never touch
    //      its bit); the words right of the diagonal are OR-reduced over the kept rows (REDUX). ----
    if (warp_id() == 0) {
        uint32_t removed = 0u;
        int nkeep = 0;
        for (int c = 0; c < words; ++c) {
            const int r0 = c << 5;
            const int rows_c = min(32, n - r0);
            const uint32_t* row = mask + (size_t)(r0 + lane) * words;
            const uint32_t diag = lane < rows_c ? row[c] : 0u;
            uint32_t cur = __shfl_sync(FULL, removed, c);
            if (rows_c < 32) cur |= ~0u << rows_c;
            unsigned todo = __ballot_sync(FULL, diag != 0u);
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t d = __shfl_sync(FULL, diag, l);
                if (!((cur >> l) & 1u)) cur |= d;
            }
            const uint32_t kbits = ~cur;
            const bool kept_row = (kbits >> lane) & 1u;
            for (int w = c + 1; w < words; ++w) {
                const uint32_t r = __reduce_or_sync(FULL, kept_row ? row[w] : 0u);
                if (lane == w) removed |= r;
            }
            if (kept_row) keep[nkeep + __popc(kbits & ((1u << lane) - 1u))] = r0 + lane;
            nkeep += __popc(kbits);
        }
        if (lane == 0) s_nkeep = nkeep;
    }
    __syncthreads();
    }   // hard NMS
    tr.mark(5);
    const int nkeep = s_nkeep;
    float* out = kept + (size_t)seg * a.K * kKeptCols;
    for (int t = threadIdx.x; t < nkeep; t += blockDim.x) {
        const int i = keep[t];
        const unsigned long long key = sorted[i];
        const float4 bx = sbox[i];
        float* o = out + (size_t)t * kKeptCols;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
        const float score = key_to_float((uint32_t)(key >> 32));
        o[4] = score;
        o[5] = __uint_as_float(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
        atomicAdd(score_hist + (size_t)img * kScoreBins + score_bin(score), 1);
    }
    return nkeep;
}

template <int NT>
__global__ void __launch_bounds__(NT)
segment_nms_kernel(NmsArgs a, TopkArgs ta, const float* __restrict__ scores, const float2* __restrict__ rowstat,
                   const int* __restrict__ cand_count, const uint2* __restrict__ cand,
                   const float4* __restrict__ boxes, const float4* __restrict__ priors, int* __restrict__ kept_count,
                   float* __restrict__ kept, int* __restrict__ status, int* __restrict__ score_hist,
                   int* __restrict__ image_done) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t s_hist[2048];
    __shared__ int s_misc[4 + NT / 32];
    __shared__ int s_valid, s_nkeep, s_ticket;
    KernelTrace trace_(TR_NMS);
    griddep_wait();
    const int seg = blockIdx.x;
    const int img = seg / a.Cf;
    const int nkeep = nms_segment<NT>(trace_, a, smem, s_hist, s_misc, s_valid, s_nkeep, scores, rowstat, cand_count, cand, boxes,
                                  priors, kept, status, score_hist);
    griddep_launch_dependents();       // late: see the note at launch_pdl
    if (threadIdx.x == 0) kept_count[seg] = nkeep;
    trace_.mark(6);
    if (image_done == nullptr) return;                     // the final top-k has its own launch
    // The last segment of an image to finish runs the image's final top-k (fence + ticket: its
    // reads of the other segments' kept rows, counts and histogram bumps are ordered after them).
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(image_done + img, 1);
    __syncthreads();
    if (s_ticket != a.Cf - 1) return;
    __threadfence();
    TopkShared<NT>& sh = *reinterpret_cast<TopkShared<NT>*>(s_hist);
    image_topk_body<NT>(smem, sh, ta, img);
}

__global__ void widen_keep_kernel(const int* __restrict__ src, const int* __restrict__ count, long long* __restrict__ dst,
                                  int cap) {
    griddep_wait();
    griddep_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && i < count[0]) dst[i] = src[i];
}

}  // namespace ssd

using namespace ssd;

extern "C" size_t ssd_postprocess_workspace_bytes(const ssd_postprocess_params* p) {
    PostPlan pl;
    if (make_plan(p, pl) != SSD_OK) return 0;
    return pl.total_bytes;
}

// threads per (image, class) CTA of the NMS kernel: SSD_NMS_THREADS = 32 / 64 / 128 / 256, read once
namespace ssd { extern int g_nms_threads; }          // abi.cu (ssd_b200_set_nms_threads)
static int nms_threads() {
    if (g_nms_threads > 0) return g_nms_threads;
    static const int nt = [] {
        const int v = env_int("SSD_NMS_THREADS", 128);
        return (v == 32 || v == 64 || v == 256) ? v : 128;
    }();
    return nt;
}

constexpr int kStagePass1 = 1, kStageRest = 2;
static int run_postprocess(const PostPlan& pl, const ssd_postprocess_params* p, const float* scores,
                           const float* boxes, const float* priors, float* dets_out, int32_t* count_out,
                           int32_t* anchor_out, int32_t* status_out, unsigned char* ws, cudaStream_t st,
                           int stages = kStagePass1 | kStageRest, uint32_t* loss_keys = nullptr) {
    float2* rowstat = (float2*)(ws + pl.off_rowstat);
    float* blockmax = (float*)(ws + pl.off_blockmax);
    float* gate = (float*)(ws + pl.off_gate);
    int* cand_count = (int*)(ws + pl.off_cand_count);
    uint2* cand = (uint2*)(ws + pl.off_cand);
    int* kept_count = (int*)(ws + pl.off_kept_count);
    float* kept = (float*)(ws + pl.off_kept);
    int* status = (int*)(ws + pl.off_status);
    int* score_hist = (int*)(ws + pl.off_score_hist);
    int* image_done = (int*)(ws + pl.off_image_done);

    const ScoreGrid& g = pl.g;
    const int grid = pl.grid;
    const size_t stream_smem = stream_smem_bytes(g);
    const size_t pass2_smem = stream_smem + kQueueBytes + round_up((size_t)pl.C * sizeof(float), 16) +
                              (pl.gate_hist ? (size_t)pl.C * kGateStride * sizeof(uint32_t) : 0);
    GateBins gbins;
    gbins.lo = pl.bin_lo; gbins.scale = pl.bin_scale;

    if ((stages & kStagePass1) && pl.fused_cluster > 0) {
        // one cluster per image: row statistics, gates and candidates in one launch (fused_select_kernel)
        FusedArgs fa;
        fa.A = pl.A; fa.C = pl.C; fa.first_fg = pl.first_fg; fa.K = pl.K; fa.converter = pl.converter; fa.cand_cap = pl.cand_cap;
        fa.rows_per_cta = pl.fused_rows_per_cta; fa.merge = pl.fused_merge; fa.split = pl.fused_split; fa.chunk_floats = pl.fused_chunk_floats;
        fa.chunk_shift = 0;
        while ((1 << fa.chunk_shift) < fa.chunk_floats) ++fa.chunk_shift;
        {   // the queues take the place of the histogram: whatever that region holds
            const size_t hist_bytes = (size_t)pl.C * kGateStride * sizeof(uint32_t);
            const size_t region = hist_bytes > kFusedQueueBytes ? hist_bytes : kFusedQueueBytes;
            fa.queue_cap = (int)(region / (kFusedWarps * 3 * sizeof(uint32_t)));
        }
        fa.score_thr = p->score_threshold; fa.bins = gbins; fa.total_floats = (long long)pl.B * pl.A * pl.C;
        fa.off_slab = pl.fused_off_slab; fa.off_trow = pl.fused_off_trow; fa.off_hist = pl.fused_off_hist; fa.off_gate = pl.fused_off_gate; fa.off_stage = pl.fused_off_stage;
#define SSD_LAUNCH_FUSED(QQ, NN, CM)                                                                                     \
    do {                                                                                                               \
        auto launch = [&](auto kern) -> int {                                                                          \
            SSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.fused_smem));    \
            prefer_max_shared(kern);                                                                                   \
            cudaLaunchConfig_t cfg = {};                                                                               \
            cfg.gridDim = dim3((unsigned)(pl.B * pl.fused_cluster));                                                   \
            cfg.blockDim = dim3(kFusedThreads);                                                                        \
            cfg.dynamicSmemBytes = pl.fused_smem;                                                                      \
            cfg.stream = st;                                                                                           \
            cudaLaunchAttribute attr[2];                                                                               \
            attr[0].id = cudaLaunchAttributeClusterDimension;                                                          \
            attr[0].val.clusterDim.x = (unsigned)pl.fused_cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1; \
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                           \
            attr[1].val.programmaticStreamSerializationAllowed = 1;                                                    \
            cfg.attrs = attr;                                                                                          \
            cfg.numAttrs = 2;                                                                                          \
            LaunchTimer lt_("fused_select", st);                                                                       \
            SSD_CUDA(cudaLaunchKernelEx(&cfg, kern, scores, fa, rowstat, loss_keys, cand_count, cand, score_hist,     \
                                        image_done, status));                                                          \
            return SSD_OK;                                                                                             \
        };                                                                                                             \
        int rc;                                                                                                        \
        if (pl.converter == SSD_CONVERT_SOFTMAX) rc = launch(fused_select_kernel<QQ, NN, CM, SSD_CONVERT_SOFTMAX>);     \
        else rc = launch(fused_select_kernel<QQ, NN, CM, SSD_CONVERT_IDENTITY>);                                        \
        if (rc != SSD_OK) return rc;                                                                                   \
    } while (0)
        SSD_DISPATCH_ROW_SHAPE_PASS1(pl.C, SSD_LAUNCH_FUSED);
#undef SSD_LAUNCH_FUSED
        SSD_CUDA(cudaGetLastError());
        count_launch();
    } else if (stages & kStagePass1) {
    // (the counters, histograms and status words of the later launches are zeroed by pass 1 itself)
    uint4* zero_ptr = (uint4*)(ws + pl.zero_begin);
    const unsigned zero_n16 = (unsigned)(pl.zero_bytes / 16);

#define SSD_LAUNCH_PASS1(QQ, NN, CM)                                                                                     \
    do {                                                                                                               \
        auto launch = [&](auto kern) -> int {                                                                          \
            SSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem));    \
            LaunchTimer lt_("pass1", st);                                                            \
            SSD_CUDA(launch_pdl(kern, dim3(pl.grid1), dim3(kStreamThreads), stream_smem, st, scores, pl.g1, rowstat, blockmax, \
                                loss_keys, zero_ptr, zero_n16));                                                \
            return SSD_OK;                                                                                             \
        };                                                                                                             \
        int rc;                                                                                                        \
        if (pl.converter == SSD_CONVERT_SOFTMAX) rc = launch(pass1_kernel_ptr<QQ, NN, CM, SSD_CONVERT_SOFTMAX>(pl.block_per_step)); \
        else rc = launch(pass1_kernel_ptr<QQ, NN, CM, SSD_CONVERT_IDENTITY>(pl.block_per_step));                        \
        if (rc != SSD_OK) return rc;                                                                                   \
    } while (0)
    SSD_DISPATCH_ROW_SHAPE_PASS1(pl.C, SSD_LAUNCH_PASS1);
#undef SSD_LAUNCH_PASS1
    SSD_CUDA(cudaGetLastError());
    count_launch();
    }
    if (!(stages & kStageRest)) return SSD_OK;

    if (pl.fused_cluster == 0) {             // streaming path: gates + second pass
    if (!pl.gate_hist) {
        LaunchTimer lt_("gate", st);
        const int merge = (pl.nblk + 32 * kGateKeysPerLane - 1) / (32 * kGateKeysPerLane);
        const int nm = (pl.nblk + merge - 1) / merge;
        const size_t gsmem = (size_t)nm * (kGateCols + 1) * sizeof(uint32_t);
        dim3 ggrid((pl.C + kGateCols - 1) / kGateCols, pl.B);
        SSD_CUDA(launch_pdl(class_gate_kernel, ggrid, dim3(kGateThreads), gsmem, st, (const float*)blockmax, pl.C,
                            pl.g.bm_stride, pl.first_fg, pl.nblk, pl.K, pl.converter, p->score_threshold, gate));
        SSD_CUDA(cudaGetLastError());
        count_launch();
    }

    GateHist gh;
    gh.blockmax = pl.gate_hist ? (const float*)blockmax : nullptr; gh.nblk = pl.nblk; gh.K = pl.K; gh.converter = pl.converter; gh.score_thr = p->score_threshold;
    gh.bins = gbins;
    bool gate_in_pass2 = pl.gate_hist;
    static const bool gate_kernel_knob = [] { const char* e = getenv("SSD_GATE_KERNEL"); return !(e && e[0] == '0'); }();   // read once
    if (pl.gate_hist && gate_kernel_knob) gate_in_pass2 = false;
    if (pl.gate_hist && !gate_in_pass2) {
        LaunchTimer lt_("gate", st);
        const size_t gsmem = round_up((size_t)pl.C * sizeof(float), 16) + (size_t)pl.C * kGateStride * sizeof(uint32_t);
        SSD_CUDA(cudaFuncSetAttribute(hist_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
        SSD_CUDA(launch_pdl(hist_gate_kernel, dim3(pl.B), dim3(kConsumerWarps * 32), gsmem, st, gh, g, gate));
        count_launch();
        gh.blockmax = nullptr;                    // pass 2 reads the gate array
    }
#define SSD_LAUNCH_PASS2(QQ, NN, CM)                                                                                     \
    do {                                                                                                               \
        auto launch = [&](auto kern) -> int {                                                                          \
            SSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass2_smem));    \
            LaunchTimer lt_("pass2", st);                                                            \
            SSD_CUDA(launch_pdl(kern, dim3(grid), dim3(kStreamThreads), pass2_smem, st, scores, g,                   \
                                (const float2*)rowstat, (const float*)gate, gh, cand_count, cand, pl.cand_cap));     \
            return SSD_OK;                                                                                             \
        };                                                                                                             \
        int rc;                                                                                                        \
        if (pl.converter == SSD_CONVERT_SOFTMAX) rc = launch(score_pass2_kernel<QQ, NN, CM, SSD_CONVERT_SOFTMAX>);         \
        else rc = launch(score_pass2_kernel<QQ, NN, CM, SSD_CONVERT_IDENTITY>);                                            \
        if (rc != SSD_OK) return rc;                                                                                   \
    } while (0)
    SSD_DISPATCH_ROW_SHAPE(pl.C, SSD_LAUNCH_PASS2);
#undef SSD_LAUNCH_PASS2
    SSD_CUDA(cudaGetLastError());
    count_launch();
    }   // streaming path

    {
        NmsArgs a;
        a.A = pl.A; a.C = pl.C; a.Cf = pl.Cf; a.first_fg = pl.first_fg; a.K = pl.K; a.cand_cap = pl.cand_cap;
        a.converter = pl.converter; a.box_input = pl.box_input; a.score_thr = p->score_threshold;
        a.xy_scale = p->xy_scale; a.wh_scale = p->wh_scale;
        nms_threshold_split(p->overlap_threshold, a);
        nms_threshold_screen(p->overlap_threshold, a);
        a.soft = p->soft_nms != 0;
        a.soft_sigma = p->soft_sigma;
        a.soft_thr = pl.soft_thr;
        TopkArgs ta;
        ta.Cf = pl.Cf; ta.K = pl.K; ta.T = pl.T; ta.det_cap = pl.det_cap; ta.kept_count = kept_count; ta.kept = kept;
        ta.score_hist = score_hist; ta.dets = dets_out; ta.det_count = count_out; ta.det_anchor = anchor_out;
        const int kwords = (pl.K + 31) / 32;
        const int nt = nms_threads();
        const size_t key_slots = nms_key_slots(pl.cand_cap, nt);
        const size_t nms_smem = key_slots * 8 + (size_t)(pl.K + 1) * (8 + 16 + 16 + 4 + 4) + (size_t)pl.K * kwords * 4 + 64;
        // The final top-k can run in the last segment CTA of every image (fence + ticket) instead of
        // in its own launch.  Measured on B200 (tools/graph_timeline.py) the 128-thread tail is slower
        // than the launch it saves unless the batch is large, so it is opt-in: SSD_TOPK=fused.
        const size_t topk_smem = topk_smem_bytes(pl.Cf, pl.T);
        static const bool want_fused_topk = [] { const char* e = getenv("SSD_TOPK"); return e && e[0] == 'f'; }();
        const bool fused_topk = want_fused_topk && topk_smem <= nms_smem + 16 * 1024;
        const size_t smem = fused_topk && topk_smem > nms_smem ? topk_smem : nms_smem;
        auto launch_nms = [&](auto kern, int threads) -> int {
            SSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            LaunchTimer lt_("nms", st);
            SSD_CUDA(launch_pdl(kern, dim3(pl.B * pl.Cf), dim3(threads), smem, st, a, ta, scores,
                                (const float2*)rowstat, (const int*)cand_count, (const uint2*)cand, (const float4*)boxes,
                                (const float4*)priors, kept_count, kept, status, score_hist,
                                fused_topk ? image_done : (int*)nullptr));
            count_launch();
            return SSD_OK;
        };
        int rc_nms;
        if (nt == 32) rc_nms = launch_nms(segment_nms_kernel<32>, 32);
        else if (nt == 64) rc_nms = launch_nms(segment_nms_kernel<64>, 64);
        else if (nt == 256) rc_nms = launch_nms(segment_nms_kernel<256>, 256);
        else rc_nms = launch_nms(segment_nms_kernel<128>, 128);
        if (rc_nms != SSD_OK) return rc_nms;
        if (!fused_topk) {
            SSD_REQUIRE(topk_smem <= 224 * 1024, SSD_ERR_UNSUPPORTED,
                        "ssd_postprocess: final top-k needs %zu bytes of shared memory", topk_smem);
            SSD_CUDA(cudaFuncSetAttribute(image_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)topk_smem));
            LaunchTimer lt_("topk", st);
            SSD_CUDA(launch_pdl(image_topk_kernel, dim3(pl.B), dim3(kTopkThreads), topk_smem, st, ta));
            count_launch();
        }
    }
    if (status_out != nullptr)
        SSD_CUDA(cudaMemcpyAsync(status_out, status, 4 * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return SSD_OK;
}

extern "C" int ssd_postprocess(const ssd_postprocess_params* p, const float* scores, const float* boxes,
                               const float* priors, float* dets_out, int32_t* count_out, int32_t* anchor_out,
                               int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream) {
    PostPlan pl;
    const int rc = make_plan(p, pl);
    if (rc != SSD_OK) return rc;
    if (pl.B == 0) return SSD_OK;
    SSD_REQUIRE(count_out && dets_out, SSD_ERR_INVALID_ARGUMENT, "ssd_postprocess: null output pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (pl.A == 0) {
        SSD_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t) * pl.B, st));
        if (status_out) SSD_CUDA(cudaMemsetAsync(status_out, 0, sizeof(int32_t) * 4, st));
        return SSD_OK;
    }
    SSD_REQUIRE(scores && boxes && workspace, SSD_ERR_INVALID_ARGUMENT, "ssd_postprocess: null pointer");
    SSD_REQUIRE(priors || pl.box_input == SSD_BOXES_CORNERS, SSD_ERR_INVALID_ARGUMENT,
                "ssd_postprocess: priors are needed to decode locs");
    SSD_REQUIRE(aligned(scores, 16) && aligned(boxes, 16) && (!priors || aligned(priors, 16)), SSD_ERR_MISALIGNED,
                "ssd_postprocess: scores, boxes and priors must be 16-byte aligned");
    SSD_REQUIRE(aligned(workspace, 256), SSD_ERR_MISALIGNED, "ssd_postprocess: workspace must be 256-byte aligned");
    SSD_REQUIRE(workspace_bytes >= pl.total_bytes, SSD_ERR_WORKSPACE, "ssd_postprocess: workspace %zu < %zu bytes",
                workspace_bytes, pl.total_bytes);
    return run_postprocess(pl, p, scores, boxes, priors, dets_out, count_out, anchor_out, status_out,
                           (unsigned char*)workspace, st, p->resume_after_pass1 ? kStageRest : (kStagePass1 | kStageRest));
}

// Pass 1 alone (memset + row statistics + gate bookkeeping), optionally emitting the sampler's
// criterion; ssd_postprocess with resume_after_pass1 = 1 on the SAME workspace finishes the job.
extern "C" int ssd_postprocess_pass1(const ssd_postprocess_params* p, const float* scores, uint32_t* loss_keys_out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    PostPlan pl;
    const int rc = make_plan(p, pl);
    if (rc != SSD_OK) return rc;
    if (pl.B == 0 || pl.A == 0) return SSD_OK;
    SSD_REQUIRE(scores && workspace, SSD_ERR_INVALID_ARGUMENT, "ssd_postprocess_pass1: null pointer");
    SSD_REQUIRE(loss_keys_out == nullptr || pl.converter == SSD_CONVERT_SOFTMAX, SSD_ERR_INVALID_ARGUMENT,
                "ssd_postprocess_pass1: the mining criterion needs the SOFTMAX converter");
    SSD_REQUIRE(aligned(scores, 16), SSD_ERR_MISALIGNED, "ssd_postprocess_pass1: scores must be 16-byte aligned");
    SSD_REQUIRE(aligned(workspace, 256), SSD_ERR_MISALIGNED, "ssd_postprocess_pass1: workspace must be 256-byte aligned");
    SSD_REQUIRE(workspace_bytes >= pl.total_bytes, SSD_ERR_WORKSPACE, "ssd_postprocess_pass1: workspace %zu < %zu bytes",
                workspace_bytes, pl.total_bytes);
    return run_postprocess(pl, p, scores, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, (unsigned char*)workspace,
                           (cudaStream_t)stream, kStagePass1, loss_keys_out);
}

// ---- box_utils.nms for one box set: the same machinery with B = 1, one score column ----
static void nms_params(int num_boxes, int max_per_class, double thr, ssd_postprocess_params& p) {
    memset(&p, 0, sizeof(p));
    p.batch = 1; p.num_anchors = num_boxes; p.num_cols = 1; p.converter = SSD_CONVERT_IDENTITY; p.first_fg_col = 0;
    p.box_input = SSD_BOXES_CORNERS; p.xy_scale = 1.f; p.wh_scale = 1.f; p.score_threshold = -INFINITY;
    p.max_per_class = max_per_class; p.overlap_threshold = thr; p.max_total = 0; p.det_capacity = max_per_class;
}

extern "C" size_t ssd_nms_workspace_bytes(int num_boxes, int max_per_class) {
    ssd_postprocess_params p;
    nms_params(num_boxes, max_per_class, 0.5, p);
    PostPlan pl;
    if (make_plan(&p, pl) != SSD_OK) return 0;
    return pl.total_bytes + round_up((size_t)max_per_class * 6 * sizeof(float), 256);
}

extern "C" int ssd_nms(const float* corner_boxes, const float* scores, int num_boxes, int max_per_class,
                       double overlap_threshold, int64_t* keep_out, int32_t* count_out, void* workspace,
                       size_t workspace_bytes, void* stream) {
    SSD_REQUIRE(num_boxes >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_nms: negative box count");
    SSD_REQUIRE(count_out != nullptr, SSD_ERR_INVALID_ARGUMENT, "ssd_nms: null count_out");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_boxes == 0) {
        SSD_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), st));
        return SSD_OK;
    }
    ssd_postprocess_params p;
    nms_params(num_boxes, max_per_class, overlap_threshold, p);
    PostPlan pl;
    const int rc = make_plan(&p, pl);
    if (rc != SSD_OK) return rc;
    SSD_REQUIRE(corner_boxes && scores && keep_out && workspace, SSD_ERR_INVALID_ARGUMENT, "ssd_nms: null pointer");
    SSD_REQUIRE(aligned(corner_boxes, 16) && aligned(scores, 16), SSD_ERR_MISALIGNED,
                "ssd_nms: boxes and scores must be 16-byte aligned");
    SSD_REQUIRE(aligned(workspace, 256), SSD_ERR_MISALIGNED, "ssd_nms: workspace must be 256-byte aligned");
    SSD_REQUIRE(workspace_bytes >= ssd_nms_workspace_bytes(num_boxes, max_per_class), SSD_ERR_WORKSPACE,
                "ssd_nms: workspace too small");
    unsigned char* ws = (unsigned char*)workspace;
    float* dets = (float*)(ws + pl.total_bytes);
    int* anchors = (int*)(ws + pl.off_anchor_tmp);
    const int rc2 = run_postprocess(pl, &p, scores, corner_boxes, nullptr, dets, count_out, anchors, nullptr, ws, st);
    if (rc2 != SSD_OK) return rc2;
    SSD_CUDA(launch_pdl(widen_keep_kernel, dim3((max_per_class + 127) / 128), dim3(128), 0, st, (const int*)anchors,
                        (const int*)count_out, (long long*)keep_out, max_per_class));
    SSD_CUDA(cudaGetLastError());
    count_launch();
    return SSD_OK;
}

// ---- box_utils.nms(soft=True): the same call with the Gaussian sweep ----
extern "C" int ssd_soft_nms(const float* corner_boxes, const float* scores, int num_boxes, int max_per_class,
                            float score_threshold, float sigma, int64_t* keep_out, int32_t* count_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
    SSD_REQUIRE(num_boxes >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_soft_nms: negative box count");
    SSD_REQUIRE(count_out != nullptr, SSD_ERR_INVALID_ARGUMENT, "ssd_soft_nms: null count_out");
    SSD_REQUIRE(sigma > 0.f, SSD_ERR_INVALID_ARGUMENT, "ssd_soft_nms: sigma must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_boxes == 0) {
        SSD_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), st));
        return SSD_OK;
    }
    ssd_postprocess_params p;
    nms_params(num_boxes, max_per_class, 0.5, p);
    p.soft_nms = 1; p.soft_sigma = sigma; p.soft_threshold = score_threshold;
    PostPlan pl;
    const int rc = make_plan(&p, pl);
    if (rc != SSD_OK) return rc;
    SSD_REQUIRE(corner_boxes && scores && keep_out && workspace, SSD_ERR_INVALID_ARGUMENT, "ssd_soft_nms: null pointer");
    SSD_REQUIRE(aligned(corner_boxes, 16) && aligned(scores, 16), SSD_ERR_MISALIGNED,
                "ssd_soft_nms: boxes and scores must be 16-byte aligned");
    SSD_REQUIRE(aligned(workspace, 256), SSD_ERR_MISALIGNED, "ssd_soft_nms: workspace must be 256-byte aligned");
    SSD_REQUIRE(workspace_bytes >= ssd_nms_workspace_bytes(num_boxes, max_per_class), SSD_ERR_WORKSPACE,
                "ssd_soft_nms: workspace too small");
    unsigned char* ws = (unsigned char*)workspace;
    float* dets = (float*)(ws + pl.total_bytes);
    int* anchors = (int*)(ws + pl.off_anchor_tmp);
    const int rc2 = run_postprocess(pl, &p, scores, corner_boxes, nullptr, dets, count_out, anchors, nullptr, ws, st);
    if (rc2 != SSD_OK) return rc2;
    SSD_CUDA(launch_pdl(widen_keep_kernel, dim3((max_per_class + 127) / 128), dim3(128), 0, st, (const int*)anchors,
                        (const int*)count_out, (long long*)keep_out, max_per_class));
    count_launch();
    return SSD_OK;
}

SSD_DEFINE_TRACE_SETTER(set_trace_postprocess)
