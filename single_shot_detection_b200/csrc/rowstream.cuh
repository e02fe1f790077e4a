// Row-tile streaming skeleton shared by the logit-streaming kernels (hard-negative mining loss,
// post-processor score passes).
//
// Layout recap: logits are [rows, C] fp32 with the class index fastest (detection/detector.py:50-66
// emits [B, A*C]).  A CTA takes tiles of `tile_rows` consecutive rows, fetches each tile with ONE
// cp.async.bulk (TMA, 1-D) into a STAGES-deep shared-memory ring guarded by mbarriers, and its
// warps reduce rows out of shared memory.  A row is owned by Q adjacent lanes (lane `sub` holds
// columns sub, sub+Q, ... in NREG registers), so one warp covers 32/Q rows per step and every
// value is read from shared memory exactly once.
#pragma once

#include "common.cuh"

namespace ssd {

constexpr int kStreamThreads = 256;
constexpr int kStreamStages = 3;

struct StreamShape {
    int tile_rows;       // rows per tile
    int stage_floats;    // floats per stage buffer (tile_rows*C + 8, rounded to 4)
    size_t smem_bytes;   // dynamic shared memory for the ring + barriers
};

// tile_rows is a multiple of `row_quantum` (so that tiles hold whole 32-row blocks etc.)
inline StreamShape make_stream_shape(int C, int row_quantum, int target_tile_bytes = 24 * 1024) {
    StreamShape s;
    int rows = target_tile_bytes / (C * 4);
    rows = rows / row_quantum * row_quantum;
    if (rows < row_quantum) rows = row_quantum;
    s.tile_rows = rows;
    s.stage_floats = (int)round_up((size_t)rows * C + 8, 4);
    s.smem_bytes = 128 + (size_t)kStreamStages * s.stage_floats * 4;
    return s;
}

// Carves dynamic shared memory and initialises the barriers.  Must be called by all threads.
template <int STAGES>
__device__ __forceinline__ void stream_setup(RowStream<STAGES>& rs, unsigned char* smem, int stage_floats) {
    rs.full = reinterpret_cast<uint64_t*>(smem);
    rs.stage_floats = stage_floats;
#pragma unroll
    for (int s = 0; s < STAGES; ++s) rs.buf[s] = reinterpret_cast<float*>(smem + 128) + (size_t)s * stage_floats;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&rs.full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
}

// Lane geometry of the Q-lanes-per-row mapping.
template <int Q>
struct RowLanes {
    int sub;       // column phase of this lane
    int rl;        // row slot of this lane inside the warp step
    static constexpr int kRowsPerWarpStep = 32 / Q;
    __device__ __forceinline__ RowLanes() : sub(lane_id() % Q), rl(lane_id() / Q) {}
};

// Load the NREG register slice of one row from a staged tile; out-of-range slots get `fill`.
template <int Q, int NREG>
__device__ __forceinline__ void load_row_slice(float (&v)[NREG], const float* __restrict__ row, int sub, int C,
                                               bool row_valid, float fill) {
#pragma unroll
    for (int i = 0; i < NREG; ++i) {
        const int col = sub + i * Q;
        v[i] = (row_valid && col < C) ? row[col] : fill;
    }
}

// Dispatch table C -> (Q, NREG).  NREG*Q >= C.
#define SSD_DISPATCH_ROW_SHAPE(C, CALL)                      \
    do {                                                     \
        if ((C) <= 8) { CALL(1, 8); }                        \
        else if ((C) <= 24) { CALL(4, 6); }                  \
        else if ((C) <= 32) { CALL(4, 8); }                  \
        else if ((C) <= 64) { CALL(8, 8); }                  \
        else if ((C) <= 88) { CALL(8, 11); }                 \
        else if ((C) <= 128) { CALL(8, 16); }                \
        else if ((C) <= 256) { CALL(32, 8); }                \
        else { CALL(32, 32); }                               \
    } while (0)

constexpr int kMaxScoreCols = 1024;

}  // namespace ssd
