// Row-tile streaming skeleton shared by the logit-streaming kernels (hard-negative mining loss,
// post-processor score passes).
//
// Layout recap: logits are [B, A, C] fp32 with the class index fastest (detection/detector.py:50-66
// emits [B, A*C]).  The CTA is warp specialised:
//   * one PRODUCER warp walks the CTA's tile list and, per tile, issues cp.async.bulk (TMA, 1-D)
//     copies -- the tile's logits plus an optional 8-byte-per-row side array (class ids or the
//     (max, sum) row statistics) -- into a STAGES-deep shared-memory ring; completion is tracked
//     by the stage's `full` mbarrier (expect_tx), reuse by its `empty` mbarrier;
//   * kConsumerWarps CONSUMER warps wait on `full`, reduce rows out of shared memory and arrive
//     on `empty`.  No block-wide barrier in the steady state.
// A row is owned by Q adjacent lanes (lane `sub` holds columns sub, sub+Q, ... in NREG registers),
// so one warp covers 32/Q rows per step and every value is read from shared memory exactly once.
// Copies are made of the 16-byte aligned superset of the wanted bytes (TMA needs 16-byte aligned
// addresses and sizes; rows are 4*C bytes, so tile starts are only 4-byte aligned in general).
#pragma once

#include <stdlib.h>

#include "common.cuh"

namespace ssd {

constexpr int kConsumerWarps = 8;
constexpr int kStreamThreads = (kConsumerWarps + 1) * 32;
constexpr int kStreamStages = 3;
constexpr int kMaxScoreCols = 1024;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// exp(d) for the STREAMED row sums, d = x - max already rounded as the reference rounds it: one
// FMUL + one MUFU.  Relative error <= 2^-22 per term, exp(0) == 1 exactly (rows whose scores tie in
// the reference keep tying); the exact per-candidate scores use expf / IEEE divide.
__device__ __forceinline__ float fast_exp(float d) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__fmul_rn(d, kLog2e)));
    return y;
}

// log for the STREAMED quantities (gate values, mining criterion): one MUFU + one FMUL, absolute
// error ~1e-7 for the sums seen here (>= 1).  Both score passes use the same function, so the gate
// value of an element is bit-identical in pass 1 and pass 2.
__device__ __forceinline__ float fast_log(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return __fmul_rn(y, kLn2);
}

// Tiling of a [B images x A rows x C cols] array.  Work item = `group_tiles` consecutive tiles of
// one image; every CTA takes a contiguous range of items.  Tiles never cross an image boundary.
struct ScoreGrid {
    int A, C, first_fg;
    int tile_rows, tiles_per_image, group_tiles, groups_per_image, num_items;
    int stage_bytes;          // bytes per ring stage (logit tile + side array, 16-byte multiples)
    int side_offset;          // byte offset of the side array inside a stage
    int nblk, split;          // post-processor block-max bookkeeping (unused by mining)
    int bm_stride;            // floats per block-maximum vector: C rounded up to a multiple of four (16-byte rows)
    int step_img, step_grp;   // gridDim.x decomposed as step_img * groups_per_image + step_grp (round-robin)
    int contiguous;           // item dealing, see TileCursor
    int items_q, items_r;     // contiguous dealing for THIS launch's grid: num_items = grid * items_q + items_r (host)
    unsigned gpi_magic;       // ceil(2^32 / groups_per_image) when item / groups_per_image == umulhi(item, magic), else 0
    int64_t total_floats;     // B*A*C
    int64_t total_rows;       // B*A
};

struct TileCursor {
    int item, item_end, img, grp, tile, tile_end;
    int stage;                 // ring stage of the current tile
    unsigned phase;            // parity of the stage's `full` barrier for the current tile
    // contiguous: CTA c owns q or q + 1 consecutive items (q = n / grid, split on the host: set_item_split) -- at most
    // two images per CTA unless an image has fewer items than a CTA's share; otherwise items are dealt round-robin.
    __device__ __forceinline__ void start(const ScoreGrid& g) {
        if (g.contiguous) {
            const int c = (int)blockIdx.x;
            item = c * g.items_q + min(c, g.items_r);
            item_end = item + g.items_q + (c < g.items_r ? 1 : 0);
        } else {
            item = blockIdx.x;
            item_end = g.num_items;
        }
        img = g.gpi_magic ? (int)__umulhi((unsigned)item, g.gpi_magic) : item / g.groups_per_image;
        grp = item - img * g.groups_per_image;
        stage = 0;
        phase = 0u;
        open(g);
    }
    __device__ __forceinline__ void open(const ScoreGrid& g) {
        tile = grp * g.group_tiles;
        tile_end = min(tile + g.group_tiles, g.tiles_per_image);
    }
    __device__ __forceinline__ bool valid(const ScoreGrid&) const { return item < item_end; }
    __device__ __forceinline__ int image(const ScoreGrid&) const { return img; }
    __device__ __forceinline__ int group(const ScoreGrid&) const { return grp; }
    __device__ __forceinline__ bool last_of_item() const { return tile + 1 == tile_end; }
    __device__ __forceinline__ void next(const ScoreGrid& g) {
        if (++stage == kStreamStages) { stage = 0; phase ^= 1u; }
        if (++tile == tile_end) {
            if (g.contiguous) {
                ++item;
                if (++grp == g.groups_per_image) { grp = 0; ++img; }
            } else {
                item += gridDim.x;
                img += g.step_img;
                grp += g.step_grp;
                if (grp >= g.groups_per_image) { grp -= g.groups_per_image; ++img; }
            }
            open(g);
        }
    }
    __device__ __forceinline__ int rows(const ScoreGrid& g) const { return min(g.tile_rows, g.A - tile * g.tile_rows); }
    __device__ __forceinline__ int64_t first_row(const ScoreGrid& g) const {
        return (int64_t)img * g.A + (int64_t)tile * g.tile_rows;
    }
    // the same in 32 bits for the consumers (plan_tiles refuses B * A >= 2^31)
    __device__ __forceinline__ unsigned first_row32(const ScoreGrid& g) const {
        return (unsigned)img * (unsigned)g.A + (unsigned)tile * (unsigned)g.tile_rows;
    }
};

// shared memory: [0,64) full barriers, [64,128) empty barriers, then the stages
__device__ __forceinline__ uint64_t* full_bar(unsigned char* smem, int s) { return reinterpret_cast<uint64_t*>(smem) + s; }
__device__ __forceinline__ uint64_t* empty_bar(unsigned char* smem, int s) { return reinterpret_cast<uint64_t*>(smem + 64) + s; }
__device__ __forceinline__ unsigned char* stage_ptr(unsigned char* smem, const ScoreGrid& g, int s) {
    return smem + 128 + (size_t)s * g.stage_bytes;
}
__device__ __forceinline__ int tile_head(int64_t first_row, int C) { return (int)(((unsigned)first_row * (unsigned)C) & 3u); }
__device__ __forceinline__ int side_head(int64_t first_row) { return (int)(first_row & 1); }

__device__ __forceinline__ void stream_init(unsigned char* smem) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStreamStages; ++s) {
            mbar_init(full_bar(smem, s), 1);
            mbar_init(empty_bar(smem, s), kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

// Producer side: called by ONE lane.  Copies the tile [first_row, first_row+rows) x C floats (and
// rows 8-byte side elements when `side` != nullptr) into stage `dst`.
__device__ __forceinline__ void produce_tile(unsigned char* dst, uint64_t* full, const ScoreGrid& g,
                                             const float* __restrict__ base, const unsigned long long* __restrict__ side,
                                             int64_t first_row, int rows, uint64_t policy) {
    const int64_t f0 = first_row * g.C;
    const int64_t f1 = f0 + (int64_t)rows * g.C;
    const int64_t fa = f0 & ~int64_t(3);
    int64_t fe = (f1 + 3) & ~int64_t(3);
    const int64_t n4 = g.total_floats & ~int64_t(3);
    if (fe > n4) fe = n4 > fa ? n4 : fa;
    const uint32_t bytes = (uint32_t)((fe - fa) * 4);
    float* tdst = reinterpret_cast<float*>(dst);
    for (int64_t f = fe; f < f1; ++f) tdst[f - fa] = __ldg(base + f);          // <= 3 floats, array tail only
    uint32_t sbytes = 0;
    int64_t sa = 0;
    unsigned long long* sdst = reinterpret_cast<unsigned long long*>(dst + g.side_offset);
    if (side != nullptr) {
        const int64_t r1 = first_row + rows;
        sa = first_row & ~int64_t(1);
        int64_t se = (r1 + 1) & ~int64_t(1);
        const int64_t n2 = g.total_rows & ~int64_t(1);
        if (se > n2) se = n2 > sa ? n2 : sa;
        sbytes = (uint32_t)((se - sa) * 8);
        for (int64_t r = se; r < r1; ++r) sdst[r - sa] = __ldg(side + r);      // <= 1 element, array tail only
    }
    if (bytes + sbytes) {
        mbar_expect_tx(full, bytes + sbytes);
        if (bytes) bulk_g2s(tdst, base + fa, bytes, full, policy);
        if (sbytes) bulk_g2s(sdst, side + sa, sbytes, full, policy);
    } else {
        mbar_arrive(full);
    }
}

// The producer warp's whole life: stream every tile of this CTA through the ring.
__device__ __forceinline__ void producer_loop(unsigned char* smem, const ScoreGrid& g, const float* __restrict__ base,
                                              const unsigned long long* __restrict__ side, uint64_t policy) {
    if (lane_id() != 0) return;
    TileCursor cur;
    cur.start(g);
    bool wrapped = false;                       // the ring has been filled once: wait for the consumers from here on
    for (; cur.valid(g); cur.next(g)) {
        const int s = cur.stage;
        if (wrapped) mbar_wait(empty_bar(smem, s), cur.phase ^ 1u);
        produce_tile(stage_ptr(smem, g, s), full_bar(smem, s), g, base, side, cur.first_row(g), cur.rows(g), policy);
        if (s == kStreamStages - 1) wrapped = true;
    }
}

// Consumer side of the ring in shared-space addresses: `sb` = smem_u32(smem), taken ONCE per thread (every
// cvta of a generic pointer costs S2R + MOV + LEA where it is used: ~25 instructions per tile in the old form).
__device__ __forceinline__ void ring_wait_full(uint32_t sb, const TileCursor& cur) {
    mbar_wait(sb + 8u * (unsigned)cur.stage, cur.phase);
}
__device__ __forceinline__ void ring_release(uint32_t sb, const TileCursor& cur) {
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(sb + 64u + 8u * (unsigned)cur.stage);
}
__device__ __forceinline__ const float* ring_logits(unsigned char* smem, const ScoreGrid& g, const TileCursor& cur, unsigned first_row) {
    return reinterpret_cast<const float*>(smem + 128 + (size_t)(cur.stage * g.stage_bytes)) + ((first_row * (unsigned)g.C) & 3u);
}
__device__ __forceinline__ const unsigned long long* ring_side(unsigned char* smem, const ScoreGrid& g, const TileCursor& cur,
                                                               unsigned first_row) {
    return reinterpret_cast<const unsigned long long*>(smem + 128 + (size_t)(cur.stage * g.stage_bytes + g.side_offset)) + (first_row & 1u);
}

// Consumer side of the ring: wait for the cursor's tile, hand out its pointers, release it.
struct StagedTile {
    const float* logits;                 // first float of the tile's first row
    const unsigned long long* side;      // first side element of the tile (8 bytes per row)
};
__device__ __forceinline__ StagedTile consumer_acquire(unsigned char* smem, const ScoreGrid& g, const TileCursor& cur,
                                                       int64_t first_row) {
    mbar_wait(full_bar(smem, cur.stage), cur.phase);
    unsigned char* st = stage_ptr(smem, g, cur.stage);
    StagedTile t;
    t.logits = reinterpret_cast<const float*>(st) + tile_head(first_row, g.C);
    t.side = reinterpret_cast<const unsigned long long*>(st + g.side_offset) + side_head(first_row);
    return t;
}
__device__ __forceinline__ void consumer_release(unsigned char* smem, const TileCursor& cur) {
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(empty_bar(smem, cur.stage));
}

// Lane geometry of the Q-lanes-per-row mapping.
template <int Q>
struct RowLanes {
    int sub;       // column phase of this lane
    int rl;        // row slot of this lane inside the warp step
    static constexpr int kRowsPerWarpStep = 32 / Q;
    __device__ __forceinline__ RowLanes() : sub(lane_id() % Q), rl(lane_id() / Q) {}
};

// Slots [0, IFULL) hold a valid column for every lane and every C of the dispatch bucket; the
// remaining slots may fall past the end of the row and are forced to -inf with a per-lane bit mask
// (exp(-inf - m) == 0 and max(-inf, x) == x, so nothing downstream needs a predicate).
template <int Q, int NREG, int CMIN>
struct RowShape {
    static constexpr int kFull = CMIN / Q < NREG ? CMIN / Q : NREG;      // slots valid for every lane
    static constexpr int kMasked = NREG - kFull;
    uint32_t keep[kMasked > 0 ? kMasked : 1];
    __device__ __forceinline__ RowShape(int sub, int C) {
#pragma unroll
        for (int j = 0; j < kMasked; ++j) keep[j] = (sub + (kFull + j) * Q < C) ? 0xFFFFFFFFu : 0u;
    }
    // Loads are unconditional: rows past the end of a partial tile and columns past the end of the
    // row read whatever is in the stage (which has slack for it); the caller discards those.
    __device__ __forceinline__ void load(float (&v)[NREG], const float* row, int sub) const {
#pragma unroll
        for (int i = 0; i < NREG; ++i) {
            float x = row[sub + i * Q];
            if (i >= kFull) {
                const uint32_t k = keep[i - kFull];
                x = __uint_as_float((__float_as_uint(x) & k) | (~k & 0xFF800000u));
            }
            v[i] = x;
        }
    }
};

// three-input maximum (FMNMX3 on sm_100a); NaN operands are dropped like fmaxf
__device__ __forceinline__ float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// warp-wide maximum in one instruction (CREDUX.MAX.F32 on sm_100a); NaN lanes are dropped
__device__ __forceinline__ float warp_max(float x) {
    float y;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(y) : "f"(x));
    return y;
}

// row max and sum of exp(x - max) over the Q lanes that own the row.  The order of the maxima is
// immaterial; the sum is a streamed quantity (see fast_exp) accumulated in three independent chains.
template <int Q, int NREG>
__device__ __forceinline__ void row_max_sum(const float (&v)[NREG], float& m, float& sum) {
    float ma = v[0], mb = v[0];
    int i = 1;
#pragma unroll
    for (; i + 3 < NREG; i += 4) {
        ma = max3(ma, v[i], v[i + 1]);
        mb = max3(mb, v[i + 2], v[i + 3]);
    }
#pragma unroll
    for (; i < NREG; ++i) ma = fmaxf(ma, v[i]);
    m = group_max<Q>(fmaxf(ma, mb));
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NREG; j += 3) {
        s0 = __fadd_rn(s0, fast_exp(__fsub_rn(v[j], m)));
        if (j + 1 < NREG) s1 = __fadd_rn(s1, fast_exp(__fsub_rn(v[j + 1], m)));
        if (j + 2 < NREG) s2 = __fadd_rn(s2, fast_exp(__fsub_rn(v[j + 2], m)));
    }
    sum = group_sum<Q>(__fadd_rn(__fadd_rn(s0, s1), s2));
}

// Dispatch table C -> (Q, NREG, CMIN): NREG*Q >= C, CMIN = smallest C of the bucket.
// Odd C <= 32 (VOC's 21 columns) take the row-per-lane shape Q == 1: lane l walks row l of the warp
// step at a stride of C words -- C odd, so the 32 lanes hit 32 different banks -- with no lane idle,
// no shuffle in the row reductions and ~3x fewer warp instructions per element than a split row.
// WIDE81: the shape for COCO's 81 columns.  Four lanes x 21 registers (no padding slot, fewer
// instructions per row) is faster for the kernels that only reduce ROWS (mining criterion, pass 2);
// pass 1 also keeps per-lane COLUMN maxima that are merged across the row slots of the warp, which
// is cheaper with eight lanes x 11 registers (measured, profiles/).
#define SSD_DISPATCH_ROW_SHAPE(C, CALL) SSD_DISPATCH_ROW_SHAPE_(C, CALL, 1)
#define SSD_DISPATCH_ROW_SHAPE_PASS1(C, CALL) SSD_DISPATCH_ROW_SHAPE_(C, CALL, 0)
#define SSD_DISPATCH_ROW_SHAPE_(C, CALL, WIDE81)                 \
    do {                                                         \
        if ((C) <= 8) { CALL(1, 8, 1); }                         \
        else if ((C) == 21) { CALL(1, 21, 21); }                 \
        else if (((C) & 1) && (C) <= 16) { CALL(1, 16, 9); }     \
        else if (((C) & 1) && (C) <= 24) { CALL(1, 24, 17); }    \
        else if (((C) & 1) && (C) <= 32) { CALL(1, 32, 25); }    \
        else if ((C) <= 16) { CALL(2, 8, 9); }                   \
        else if ((C) <= 24) { CALL(4, 6, 17); }                  \
        else if ((C) <= 32) { CALL(4, 8, 25); }                  \
        else if ((C) <= 64) { CALL(8, 8, 33); }                  \
        else if ((C) == 81 && (WIDE81)) { CALL(4, 21, 81); }     \
        else if ((C) <= 88) { CALL(8, 11, 65); }                 \
        else if ((C) <= 128) { CALL(8, 16, 89); }                \
        else if ((C) <= 256) { CALL(32, 8, 129); }               \
        else { CALL(32, 32, 257); }                              \
    } while (0)

inline int lanes_per_row(int C, bool pass1 = false) {
    int q = 0;
#define SSD_Q_(QQ, NN, CM) q = QQ
    if (pass1) SSD_DISPATCH_ROW_SHAPE_PASS1(C, SSD_Q_);
    else SSD_DISPATCH_ROW_SHAPE(C, SSD_Q_);
#undef SSD_Q_
    return q;
}
// register slots per row (>= C): the widest of the shapes any kernel uses for this C
inline int slots_per_row(int C) {
    int n = 0, n1 = 0;
#define SSD_N_(QQ, NN, CM) n = QQ * NN
    SSD_DISPATCH_ROW_SHAPE(C, SSD_N_);
#undef SSD_N_
#define SSD_N_(QQ, NN, CM) n1 = QQ * NN
    SSD_DISPATCH_ROW_SHAPE_PASS1(C, SSD_N_);
#undef SSD_N_
    return n > n1 ? n : n1;
}

// Host: tile geometry.  tile_rows is a multiple of the rows all consumer warps cover in one step.
inline int env_int(const char* name, int fallback) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : fallback;
}
inline void plan_tiles(ScoreGrid& g, int images, int A, int C, bool with_side, int target_tile_bytes = 24 * 1024) {
    static const int tile_bytes_knob = env_int("SSD_TILE_BYTES", 0);           // tuning knob, read once (tools/kernel_times.sh)
    if (tile_bytes_knob > 0) target_tile_bytes = tile_bytes_knob;
    const int q_min = lanes_per_row(C) < lanes_per_row(C, true) ? lanes_per_row(C) : lanes_per_row(C, true);
    const int quantum = kConsumerWarps * (32 / q_min);        // a whole number of warp steps for every kernel
    int rows = target_tile_bytes / (C * 4);
    rows = rows / quantum * quantum;
    if (rows > 4 * quantum) rows = 4 * quantum;
    if (rows < quantum) rows = quantum;
    g.A = A; g.C = C;
    g.tile_rows = rows;
    g.tiles_per_image = (A + rows - 1) / rows;
    g.group_tiles = 1;
    g.groups_per_image = g.tiles_per_image;
    g.num_items = images * g.groups_per_image;
    // head alignment slack (<= 12 bytes) + the columns an unconditional row load may overshoot
    const size_t tile_bytes = round_up((size_t)rows * C * 4 + 16 + (size_t)(slots_per_row(C) - C) * 4 + 16, 16);
    g.side_offset = (int)tile_bytes;
    g.stage_bytes = (int)(tile_bytes + (with_side ? round_up((size_t)rows * 8 + 32, 16) : 0));
    g.total_floats = (int64_t)images * A * C;
    g.total_rows = (int64_t)images * A;
    g.first_fg = 0; g.nblk = 0; g.split = 1;
    g.bm_stride = (C + 3) & ~3;
    g.step_img = 0; g.step_grp = 0;
    g.items_q = 0; g.items_r = 0; g.gpi_magic = 0u;
    static const int contiguous_knob = [] { const char* e = getenv("SSD_CURSOR"); return (e && e[0] == 'r') ? 0 : 1; }();   // read once
    g.contiguous = contiguous_knob;
}
// contiguous item ranges of a launch with `grid` CTAs (TileCursor::start) and the division-free image index
inline void set_item_split(ScoreGrid& g, int grid) {
    if (grid < 1) grid = 1;
    g.items_q = g.num_items / grid;
    g.items_r = g.num_items % grid;
    const unsigned d = (unsigned)g.groups_per_image;
    // umulhi(n, ceil(2^32 / d)) == n / d for n, d < 2^16 (the error term n * (M d - 2^32) / (d 2^32) stays below 1 / d)
    g.gpi_magic = (d >= 2u && d < 65536u && g.num_items < 65536) ? (unsigned)((0x100000000ull + d - 1) / d) : 0u;
}
inline size_t stream_smem_bytes(const ScoreGrid& g) { return 128 + (size_t)kStreamStages * g.stage_bytes; }
__device__ __forceinline__ size_t stream_smem_bytes_dev(const ScoreGrid& g) { return 128 + (size_t)kStreamStages * g.stage_bytes; }
// Grid size: a whole number of resident waves.
inline int stream_grid(ScoreGrid& g, size_t extra_smem = 0) {
    int per_sm = (int)((size_t)(224 * 1024) / (stream_smem_bytes(g) + extra_smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (stream_ctas_override() > 0) per_sm = stream_ctas_override();            // see abi.cu
    if (per_sm < 1) per_sm = 1;
    int grid = per_sm * sm_count();
    if (grid > g.num_items) grid = g.num_items;
    if (grid < 1) grid = 1;
    g.step_img = grid / g.groups_per_image;
    g.step_grp = grid % g.groups_per_image;
    set_item_split(g, grid);
    return grid;
}

}  // namespace ssd
