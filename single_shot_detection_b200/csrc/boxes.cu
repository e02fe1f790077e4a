// Box format conversion and centre-size coding: bf/utils/box_utils.py:16-36,
// detection/box_coder.py:13-57.  Element-wise, one thread per box row, 16-byte (or 2x8-byte when a
// row is a 24-byte target row) vector accesses.  Every fp32 op is rounded separately and follows
// the operation order of the reference branch it replaces -- the in-place and out-of-place
// branches of the reference differ in rounding, so both orders exist.
#include "common.cuh"

namespace ssd {

template <bool VEC4>
__device__ __forceinline__ Box load_box(const float* p) {
    if (VEC4) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        return {v.x, v.y, v.z, v.w};
    }
    const float2 lo = *reinterpret_cast<const float2*>(p);
    const float2 hi = *reinterpret_cast<const float2*>(p + 2);
    return {lo.x, lo.y, hi.x, hi.y};
}
template <bool VEC4>
__device__ __forceinline__ void store_box(float* p, Box v) {
    if (VEC4) {
        *reinterpret_cast<float4*>(p) = make_float4(v.a, v.b, v.c, v.d);
    } else {
        *reinterpret_cast<float2*>(p) = make_float2(v.a, v.b);
        *reinterpret_cast<float2*>(p + 2) = make_float2(v.c, v.d);
    }
}

__device__ __forceinline__ Box to_corners(Box c) {                       // box_utils.py:23
    const float hw = fmul(c.c, 0.5f), hh = fmul(c.d, 0.5f);
    return {fsub(c.a, hw), fsub(c.b, hh), fadd(c.a, hw), fadd(c.b, hh)};
}
__device__ __forceinline__ Box to_centroids(Box m) {                     // box_utils.py:36
    return {fmul(fadd(m.c, m.a), 0.5f), fmul(fadd(m.d, m.b), 0.5f), fsub(m.c, m.a), fsub(m.d, m.b)};
}
__device__ __forceinline__ Box encode(Box b, float4 p, float xy, float wh, float eps) {   // box_coder.py:32-34
    return {fmul(fdiv(fsub(b.a, p.x), p.z), xy), fmul(fdiv(fsub(b.b, p.y), p.w), xy),
            fmul(logf(fdiv(fadd(b.c, eps), p.z)), wh), fmul(logf(fdiv(fadd(b.d, eps), p.w)), wh)};
}
__device__ __forceinline__ Box decode(Box l, float4 p, float xy, float wh) {              // box_coder.py:55-57
    return {fadd(p.x, fdiv(fmul(p.z, l.a), xy)), fadd(p.y, fdiv(fmul(p.w, l.b), xy)),
            fmul(p.z, expf(fdiv(l.c, wh))), fmul(p.w, expf(fdiv(l.d, wh)))};
}
__device__ __forceinline__ Box decode_inplace(Box l, float4 p, float xy, float wh) {      // box_coder.py:47-52
    return {fadd(fmul(fdiv(l.a, xy), p.z), p.x), fadd(fmul(fdiv(l.b, xy), p.w), p.y),
            fmul(expf(fdiv(l.c, wh)), p.z), fmul(expf(fdiv(l.d, wh)), p.w)};
}

template <int OP, bool VEC4>
__global__ void __launch_bounds__(256)
box_transform_kernel(const float* __restrict__ src, int64_t src_stride, float* dst, int64_t dst_stride,
                     const float4* __restrict__ priors, int64_t rows, int A, float xy, float wh, float eps) {
    KernelTrace trace_(TR_BOX0 + OP);
    griddep_wait();
    griddep_launch_dependents();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    Box b = load_box<VEC4>(src + r * src_stride);
    float4 p = make_float4(0.f, 0.f, 1.f, 1.f);
    if (OP >= SSD_BOX_ENCODE) p = priors[r % A];
    Box o;
    switch (OP) {
        case SSD_BOX_TO_CORNERS: o = to_corners(b); break;
        case SSD_BOX_TO_CENTROIDS: o = to_centroids(b); break;
        case SSD_BOX_TO_CENTROIDS_INPLACE: o = to_centroids_inplace(b); break;
        case SSD_BOX_ENCODE: o = encode(b, p, xy, wh, eps); break;
        case SSD_BOX_ENCODE_INPLACE: o = encode_inplace(b, p, xy, wh, eps); break;
        case SSD_BOX_DECODE: o = decode(b, p, xy, wh); break;
        case SSD_BOX_DECODE_INPLACE: o = decode_inplace(b, p, xy, wh); break;
        case SSD_BOX_CENTROIDS_ENCODE_INPLACE: o = encode_inplace(to_centroids_inplace(b), p, xy, wh, eps); break;
        default: o = to_corners(decode(b, p, xy, wh)); break;
    }
    store_box<VEC4>(dst + r * dst_stride, o);
}

template <int OP>
static int launch_box(bool vec4, const float* src, int64_t ss, float* dst, int64_t ds, const float* priors,
                      int64_t rows, int A, float xy, float wh, float eps, cudaStream_t st) {
    const unsigned blocks = (unsigned)((rows + 255) / 256);
    LaunchTimer lt_("box", st);
    if (vec4)
        SSD_CUDA(launch_plain(box_transform_kernel<OP, true>, dim3(blocks), dim3(256), 0, st, src, ss, dst, ds,
                            (const float4*)priors, rows, A, xy, wh, eps));
    else
        SSD_CUDA(launch_plain(box_transform_kernel<OP, false>, dim3(blocks), dim3(256), 0, st, src, ss, dst, ds,
                            (const float4*)priors, rows, A, xy, wh, eps));
    SSD_CUDA(cudaGetLastError());
    count_launch();
    return SSD_OK;
}

}  // namespace ssd

using namespace ssd;

extern "C" int ssd_box_transform(int op, const float* src, int64_t src_row_stride, float* dst, int64_t dst_row_stride,
                                 const float* priors, int64_t num_rows, int num_anchors, float xy_scale,
                                 float wh_scale, float eps, void* stream) {
    SSD_REQUIRE(op >= 0 && op <= SSD_BOX_DECODE_TO_CORNERS, SSD_ERR_INVALID_ARGUMENT, "ssd_box_transform: bad op %d", op);
    SSD_REQUIRE(num_rows >= 0, SSD_ERR_INVALID_ARGUMENT, "ssd_box_transform: negative row count");
    if (num_rows == 0) return SSD_OK;
    SSD_REQUIRE(src && dst, SSD_ERR_INVALID_ARGUMENT, "ssd_box_transform: null pointer");
    SSD_REQUIRE(src_row_stride >= 4 && dst_row_stride >= 4, SSD_ERR_INVALID_ARGUMENT,
                "ssd_box_transform: row stride below 4 floats");
    const bool needs_priors = op >= SSD_BOX_ENCODE;
    SSD_REQUIRE(!needs_priors || (priors && num_anchors > 0), SSD_ERR_INVALID_ARGUMENT,
                "ssd_box_transform: this op needs priors");
    SSD_REQUIRE(!needs_priors || aligned(priors, 16), SSD_ERR_MISALIGNED, "ssd_box_transform: priors not 16-byte aligned");
    const bool vec4 = aligned(src, 16) && aligned(dst, 16) && src_row_stride % 4 == 0 && dst_row_stride % 4 == 0;
    const bool vec2 = aligned(src, 8) && aligned(dst, 8) && src_row_stride % 2 == 0 && dst_row_stride % 2 == 0;
    SSD_REQUIRE(vec4 || vec2, SSD_ERR_MISALIGNED,
                "ssd_box_transform: rows must be at least 8-byte aligned (even float strides)");
    cudaStream_t st = (cudaStream_t)stream;
    if (num_anchors <= 0) num_anchors = 1;
#define SSD_BOX_CASE(OPV) \
    case OPV: return launch_box<OPV>(vec4, src, src_row_stride, dst, dst_row_stride, priors, num_rows, num_anchors, xy_scale, wh_scale, eps, st)
    switch (op) {
        SSD_BOX_CASE(SSD_BOX_TO_CORNERS);
        SSD_BOX_CASE(SSD_BOX_TO_CENTROIDS);
        SSD_BOX_CASE(SSD_BOX_TO_CENTROIDS_INPLACE);
        SSD_BOX_CASE(SSD_BOX_ENCODE);
        SSD_BOX_CASE(SSD_BOX_ENCODE_INPLACE);
        SSD_BOX_CASE(SSD_BOX_DECODE);
        SSD_BOX_CASE(SSD_BOX_DECODE_INPLACE);
        SSD_BOX_CASE(SSD_BOX_CENTROIDS_ENCODE_INPLACE);
        SSD_BOX_CASE(SSD_BOX_DECODE_TO_CORNERS);
    }
#undef SSD_BOX_CASE
    return SSD_ERR_INVALID_ARGUMENT;
}

SSD_DEFINE_TRACE_SETTER(set_trace_boxes)
