// Anchor (prior box) tables on the device: detection/anchor_generators/ssd.py:106-151,
// retina_net.py:28-54 and the per-level concatenation of detection/detector.py:82-86.
// The reference fills [H, W, boxes, 4] per level on the CPU (linspace -> meshgrid -> four strided
// assignments), flattens and concatenates the levels every forward pass, and the pipeline then ships
// the table to the GPU.  Here ONE launch writes the whole [A, 4] table of every level in HBM, one
// thread per anchor (a float4 store).
//
// Bit parity with the CPU table: torch.linspace (ATen RangeFactoriesKernel, fp32) evaluates
//     step = (end - start) / (steps - 1)
//     x[i] = fma(step, i, start)                  for i <  steps / 2
//     x[i] = fma(-step, steps - 1 - i, end)       otherwise
// on x86 builds with FMA (the AVX2 / AVX512 kernels; probed against this image's torch 2.11 for every
// feature-map size of the sample configs).  start / end arrive as the fp32 casts of the doubles the
// Python code computes (`offset * step_w`, `(offset + cells - 1) * step_w`); the (w, h) shapes are a
// few scalars per level, computed by the host exactly as the reference does and passed by value.
#include "common.cuh"

namespace ssd {

struct AnchorLevels {
    SsdAnchorLevel level[SSD_MAX_ANCHOR_LEVELS];
    long long first[SSD_MAX_ANCHOR_LEVELS + 1];        // first anchor of each level
    int count;
};

__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
    if (steps <= 1) return start;
    const float step = fdiv(fsub(end, start), (float)(steps - 1));
    return i < steps / 2 ? __fmaf_rn(step, (float)i, start) : __fmaf_rn(-step, (float)(steps - 1 - i), end);
}

__global__ void __launch_bounds__(256)
generate_anchors_kernel(const __grid_constant__ AnchorLevels L, float4* __restrict__ out, long long total) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= total) return;
    int lv = 0;
    while (lv + 1 < L.count && a >= L.first[lv + 1]) ++lv;
    const SsdAnchorLevel& P = L.level[lv];
    const long long r = a - L.first[lv];
    const int box = (int)(r % P.num_boxes);
    const long long cell = r / P.num_boxes;
    const int cx = (int)(cell % P.cells_x), cy = (int)(cell / P.cells_x);
    out[a] = make_float4(linspace_at(P.x_start, P.x_end, P.cells_x, cx), linspace_at(P.y_start, P.y_end, P.cells_y, cy),
                         P.wh[2 * box], P.wh[2 * box + 1]);
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_generate_anchors(const SsdAnchorLevel* levels, int num_levels, float* anchors_out,
                                    int64_t num_anchors, void* stream) {
    SSD_REQUIRE(num_levels >= 0 && num_levels <= SSD_MAX_ANCHOR_LEVELS, SSD_ERR_INVALID_ARGUMENT,
                "ssd_generate_anchors: %d levels (at most %d)", num_levels, SSD_MAX_ANCHOR_LEVELS);
    SSD_REQUIRE(levels || num_levels == 0, SSD_ERR_INVALID_ARGUMENT, "ssd_generate_anchors: levels is null");
    AnchorLevels L;
    memset(&L, 0, sizeof(L));
    long long total = 0;
    for (int i = 0; i < num_levels; ++i) {
        const SsdAnchorLevel& p = levels[i];
        SSD_REQUIRE(p.cells_x >= 0 && p.cells_y >= 0 && p.num_boxes >= 1 && p.num_boxes <= SSD_MAX_BOXES_PER_CELL,
                    SSD_ERR_INVALID_ARGUMENT, "ssd_generate_anchors: level %d: %d x %d cells, %d boxes per cell", i,
                    p.cells_x, p.cells_y, p.num_boxes);
        L.level[i] = p;
        L.first[i] = total;
        total += (long long)p.cells_x * p.cells_y * p.num_boxes;
    }
    L.first[num_levels] = total;
    L.count = num_levels;
    SSD_REQUIRE(total == num_anchors, SSD_ERR_INVALID_ARGUMENT,
                "ssd_generate_anchors: the levels hold %lld anchors, the output %lld", total, (long long)num_anchors);
    if (total == 0) return SSD_OK;
    SSD_REQUIRE(anchors_out && aligned(anchors_out, 16), SSD_ERR_MISALIGNED,
                "ssd_generate_anchors: anchors_out must be a 16-byte aligned device pointer");
    SSD_CUDA(launch_pdl(generate_anchors_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
                        L, reinterpret_cast<float4*>(anchors_out), total));
    count_launch();
    return SSD_OK;
}

SSD_DEFINE_TRACE_SETTER(set_trace_anchors)
