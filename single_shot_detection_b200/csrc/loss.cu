// Masked multibox losses, forward and gradient in one pass: detection/losses/multibox_loss.py:56-92
// with the classification losses the samples configure (torch CrossEntropyLoss(reduction='sum',
// ignore_index=-1); SigmoidFocalLoss, bf/modules/losses.py:34-54) and SmoothL1Loss(reduction='sum').
//
// The reference gathers scores[sampled_mask] / locs[positive_mask] into fresh tensors, runs the
// loss modules on them and lets autograd scatter the gradients back.  Only a few percent of the
// anchors are sampled, so here nothing is gathered:
//   1. loss_count_kernel   counts the positives (the divider of both terms, multibox_loss.py:88);
//   2. loss_rows_kernel    one warp per 32 anchors: the mask bytes and target rows are read
//      coalesced, a ballot picks the sampled rows and the warp walks those with its lanes over the
//      score columns (row max / sum by shuffles, accurate expf / logf); the gradient rows
//      d(loss)/d(logits) and d(loss)/d(locs) are written already scaled by weight / divider, zeros
//      for the anchors outside the masks; per-CTA partial sums go to the workspace in double;
//   3. loss_finish_kernel  adds the partials in a fixed order (deterministic) and writes
//      (loss, class_loss, loc_loss).
// The logits of unsampled anchors are never read: the pass is bound by the gradient WRITE
// (4 C + 16 bytes per anchor) when gradients are requested and by the mask / target read otherwise.
#include <math.h>

#include "common.cuh"

namespace ssd {

constexpr int kLossThreads = 256;
constexpr int kLossRowsPerCta = kLossThreads;          // one anchor per thread in the coalesced phases

struct LossArgs {
    int A, C, kind;                 // kind: SSD_LOSS_SOFTMAX_CE / SSD_LOSS_SIGMOID_FOCAL
    float gamma, alpha, class_weight, loc_weight;
    int64_t rows;                   // B * A
    // localisation term: SmoothL1 on coded boxes (priors == nullptr) or GeneralizedIoULoss on decoded corner
    // boxes against the CORNER target rows (multibox_loss.py:77-79, bf/modules/losses.py:109-114)
    const float4* priors;
    float xy_scale, wh_scale;
};

// derivative of min(a, b) / max(a, b) with respect to a, torch's convention at ties (half each)
__device__ __forceinline__ float dmin_da(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float dmax_da(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }

// 1 - generalized_iou(decode -> to_corners (locs), target corners) of one anchor and d/d locs.
//   decode (out of place, box_coder.py:55-57): c = p_xy + (p_wh * l_xy) / xy_scale, s = p_wh * exp(l_wh / wh_scale)
//   giou = I / U - (E - U) / E  (box_utils.py:104-143; clamped areas), loss = 1 - giou
__device__ __forceinline__ float giou_loss_term(float4 l, float4 p, float4 t, float xs, float ws, float (&g)[4]) {
    const float cx = p.x + (p.z * l.x) / xs, cy = p.y + (p.w * l.y) / xs;
    const float w = p.z * expf(l.z / ws), h = p.w * expf(l.w / ws);
    const float x1 = cx - w * 0.5f, y1 = cy - h * 0.5f, x2 = cx + w * 0.5f, y2 = cy + h * 0.5f;
    const float aw = fmaxf(x2 - x1, 0.f), ah = fmaxf(y2 - y1, 0.f);
    const float area_a = aw * ah;
    const float area_b = fmaxf(t.z - t.x, 0.f) * fmaxf(t.w - t.y, 0.f);
    const float ix1 = fmaxf(x1, t.x), iy1 = fmaxf(y1, t.y), ix2 = fminf(x2, t.z), iy2 = fminf(y2, t.w);
    const float iw = fmaxf(ix2 - ix1, 0.f), ih = fmaxf(iy2 - iy1, 0.f);
    const float I = iw * ih;
    const float U = area_a + area_b - I;
    const float ex1 = fminf(x1, t.x), ey1 = fminf(y1, t.y), ex2 = fmaxf(x2, t.z), ey2 = fmaxf(y2, t.w);
    const float ew = fmaxf(ex2 - ex1, 0.f), eh = fmaxf(ey2 - ey1, 0.f);
    const float E = ew * eh;
    const float giou = I / U - (E - U) / E;
    // d loss: giou = I/U - 1 + U/E
    const float GA = I / (U * U) - 1.f / E;                  // d loss / d area_a  (through U)
    const float GI = -1.f / U - GA;                          // d loss / d I       (direct and through U)
    const float GE = U / (E * E);                            // d loss / d E
    // clamp(min = 0) passes the gradient where its input is >= 0 (torch)
    const float m_aw = (x2 - x1) >= 0.f, m_ah = (y2 - y1) >= 0.f;
    const float m_iw = (ix2 - ix1) >= 0.f, m_ih = (iy2 - iy1) >= 0.f;
    const float m_ew = (ex2 - ex1) >= 0.f, m_eh = (ey2 - ey1) >= 0.f;
    const float gx2 = GA * m_aw * ah + GI * ih * m_iw * dmin_da(x2, t.z) + GE * eh * m_ew * dmax_da(x2, t.z);
    const float gx1 = -GA * m_aw * ah - GI * ih * m_iw * dmax_da(x1, t.x) - GE * eh * m_ew * dmin_da(x1, t.x);
    const float gy2 = GA * m_ah * aw + GI * iw * m_ih * dmin_da(y2, t.w) + GE * ew * m_eh * dmax_da(y2, t.w);
    const float gy1 = -GA * m_ah * aw - GI * iw * m_ih * dmax_da(y1, t.y) - GE * ew * m_eh * dmin_da(y1, t.y);
    g[0] = (gx1 + gx2) * p.z / xs;
    g[1] = (gy1 + gy2) * p.w / xs;
    g[2] = (gx2 - gx1) * 0.5f * w / ws;
    g[3] = (gy2 - gy1) * 0.5f * h / ws;
    return 1.f - giou;
}

__device__ __forceinline__ bool is_positive_class(float c) {
    return c != (float)SSD_NEGATIVE_CLASS && c != (float)SSD_IGNORE_CLASS;
}

// counts[0] = positives (the divider, multibox_loss.py:88), counts[1] = sampled anchors
__global__ void __launch_bounds__(kLossThreads)
loss_count_kernel(const float* __restrict__ target, const uint8_t* __restrict__ sampled, int64_t rows,
                  int* __restrict__ counts) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    int c = 0, m = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        c += is_positive_class(target[r * SSD_TARGET_COLS + SSD_CLASS_COL]);
        m += sampled[r] != 0;
    }
    c = __reduce_add_sync(FULL, c);
    m = __reduce_add_sync(FULL, m);
    if (lane_id() == 0 && c) atomicAdd(counts, c);
    if (lane_id() == 0 && m) atomicAdd(counts + 1, m);
}

// focal term of one (logit, soft target) pair and its derivative, bf/modules/losses.py:42-52
__device__ __forceinline__ void focal_term(float x, float t, float gamma, float alpha, float& loss, float& grad) {
    const float s = 1.f / (1.f + expf(-x));
    const float aw = t * alpha + (1.f - t) * (1.f - alpha);
    const float pb = s * t + (1.f - s) * (1.f - t);
    // binary_cross_entropy_with_logits: max(x, 0) - x t + log(1 + exp(-|x|))
    const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
    const float q = 1.f - pb;
    const float qg = powf(q, gamma);
    loss = aw * qg * ce;
    const float dpb = s * (1.f - s) * (2.f * t - 1.f);
    const float dqg = q > 0.f ? gamma * powf(q, gamma - 1.f) : 0.f;
    grad = aw * (qg * (s - t) - dqg * dpb * ce);
}

__global__ void __launch_bounds__(kLossThreads)
loss_rows_kernel(LossArgs a, const float* __restrict__ logits, const float* __restrict__ locs,
                 const float* __restrict__ target, const uint8_t* __restrict__ sampled, const int* __restrict__ n_pos,
                 float* __restrict__ grad_logits, float* __restrict__ grad_locs, double* __restrict__ partials) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    __shared__ double s_cls[kLossThreads / 32], s_loc[kLossThreads / 32];
    const int lane = lane_id();
    const int npos = n_pos[0];
    const float divider = (float)(npos > 1 ? npos : 1);                      // .clamp_(min=1).float()
    // SigmoidFocalLoss reduces with 'mean' over the sampled anchors in the reference (see the header
    // of this function's C-ABI entry in include/ssd_b200.h): one more division for that kind
    const float cmean = a.kind == SSD_LOSS_SIGMOID_FOCAL ? (float)n_pos[1] : 1.f;
    const float cscale = a.class_weight / divider / cmean, lscale = a.loc_weight / divider;
    double cls_sum = 0.0, loc_sum = 0.0;

    const int64_t r = (int64_t)blockIdx.x * kLossRowsPerCta + threadIdx.x;
    const bool in_range = r < a.rows;
    float cls = 0.f, tscore = 0.f;
    bool take = false;
    if (in_range) {
        const float2 cs = *reinterpret_cast<const float2*>(target + r * SSD_TARGET_COLS + SSD_CLASS_COL);
        cls = cs.x; tscore = cs.y;
        take = sampled[r] != 0;
    }
    // ---- localisation: SmoothL1 (beta = 1) over the positives, one thread per anchor ----
    if (in_range) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (is_positive_class(cls)) {
            const float4 p = *reinterpret_cast<const float4*>(locs + r * 4);
            const float2 t0 = *reinterpret_cast<const float2*>(target + r * SSD_TARGET_COLS);
            const float2 t1 = *reinterpret_cast<const float2*>(target + r * SSD_TARGET_COLS + 2);
            float gd[4];
            float l = 0.f;
            if (a.priors != nullptr) {
                l = giou_loss_term(p, a.priors[r % a.A], make_float4(t0.x, t0.y, t1.x, t1.y), a.xy_scale, a.wh_scale, gd);
            } else {
                const float d[4] = {p.x - t0.x, p.y - t0.y, p.z - t1.x, p.w - t1.y};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float ad = fabsf(d[k]);
                    l += ad < 1.f ? 0.5f * d[k] * d[k] : ad - 0.5f;
                    gd[k] = ad < 1.f ? d[k] : (d[k] > 0.f ? 1.f : -1.f);
                }
            }
            loc_sum += (double)l;
            g = make_float4(gd[0] * lscale, gd[1] * lscale, gd[2] * lscale, gd[3] * lscale);
        }
        if (grad_locs != nullptr) *reinterpret_cast<float4*>(grad_locs + r * 4) = g;
    }
    // ---- classification: the warp walks its sampled rows, lanes over the score columns ----
    const int64_t warp_row0 = (int64_t)blockIdx.x * kLossRowsPerCta + (threadIdx.x & ~31);
    unsigned todo = __ballot_sync(FULL, take);
    if (grad_logits != nullptr) {
        // zero rows of the anchors outside the mask: the warp's 32 rows are one contiguous span
        const unsigned skip = __ballot_sync(FULL, in_range && !take);
        const int64_t base = warp_row0 * a.C;
        for (int i = lane; i < 32 * a.C; i += 32) {
            const int rr = i / a.C;
            if ((skip >> rr) & 1u) grad_logits[base + i] = 0.f;
        }
    }
    while (todo) {
        const int rr = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t row = warp_row0 + rr;
        const float rcls = __shfl_sync(FULL, cls, rr);
        const float rscore = __shfl_sync(FULL, tscore, rr);
        const float* x = logits + row * a.C;
        float* gx = grad_logits != nullptr ? grad_logits + row * a.C : nullptr;
        if (a.kind == SSD_LOSS_SOFTMAX_CE) {
            // -log_softmax(x)[cls], ignore_index = -1 (a sampled ignored anchor contributes nothing)
            const int c_t = (int)rcls;
            float m = -INFINITY;
            for (int c = lane; c < a.C; c += 32) m = fmaxf(m, x[c]);
            m = group_max<32>(m);
            float s = 0.f;
            for (int c = lane; c < a.C; c += 32) s += expf(x[c] - m);
            s = group_sum<32>(s);
            const float lse = m + logf(s);
            const bool live = rcls != (float)SSD_IGNORE_CLASS && c_t >= 0 && c_t < a.C;
            if (live && lane == 0) cls_sum += (double)(lse - x[c_t]);
            if (gx != nullptr)
                for (int c = lane; c < a.C; c += 32)
                    gx[c] = live ? (expf(x[c] - lse) - (c == c_t ? 1.f : 0.f)) * cscale : 0.f;
        } else {
            // one-hot soft target at column cls - 1 scaled by the GT score (multibox_loss.py:64-67)
            const int c_t = is_positive_class(rcls) ? (int)rcls - 1 : -1;
            float l = 0.f;
            for (int c = lane; c < a.C; c += 32) {
                float li, gi;
                focal_term(x[c], c == c_t ? rscore : 0.f, a.gamma, a.alpha, li, gi);
                l += li;
                if (gx != nullptr) gx[c] = gi * cscale;
            }
            l = group_sum<32>(l);
            if (lane == 0) cls_sum += (double)l;
        }
    }
    // ---- per-CTA partials (fixed order inside the CTA: warp shuffles, then warp 0 adds the warps) ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cls_sum += __shfl_xor_sync(FULL, cls_sum, o);
        loc_sum += __shfl_xor_sync(FULL, loc_sum, o);
    }
    if (lane == 0) { s_cls[warp_id()] = cls_sum; s_loc[warp_id()] = loc_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, l = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) { c += s_cls[w]; l += s_loc[w]; }
        partials[2 * (size_t)blockIdx.x] = c;
        partials[2 * (size_t)blockIdx.x + 1] = l;
    }
}

__global__ void __launch_bounds__(1024)
loss_finish_kernel(const double* __restrict__ partials, int n, const int* __restrict__ n_pos, int kind,
                   float class_weight, float loc_weight, float* __restrict__ loss_out) {
    KernelTrace trace_(TR_MISC);
    griddep_wait();
    griddep_launch_dependents();
    __shared__ double s[2][32];
    double c = 0.0, l = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { c += partials[2 * (size_t)i]; l += partials[2 * (size_t)i + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(FULL, c, o);
        l += __shfl_xor_sync(FULL, l, o);
    }
    if (lane_id() == 0) { s[0][warp_id()] = c; s[1][warp_id()] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cs = 0.0, ls = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { cs += s[0][w]; ls += s[1][w]; }
        const int np = n_pos[0];
        const float divider = (float)(np > 1 ? np : 1);
        // the reference reduces in fp32, scales by the weight, then divides (multibox_loss.py:89-90)
        float class_sum = (float)cs;
        if (kind == SSD_LOSS_SIGMOID_FOCAL) class_sum = class_sum / (float)n_pos[1];      // 'mean'; 0/0 -> NaN as torch
        const float class_loss = class_sum * class_weight / divider;
        const float loc_loss = (float)ls * loc_weight / divider;
        loss_out[0] = class_loss + loc_loss;
        loss_out[1] = class_loss;
        loss_out[2] = loc_loss;
    }
}

}  // namespace ssd

using namespace ssd;

static int loss_ctas(int64_t rows) { return (int)((rows + kLossRowsPerCta - 1) / kLossRowsPerCta); }

extern "C" size_t ssd_multibox_loss_workspace_bytes(int batch, int num_anchors) {
    if (batch <= 0 || num_anchors <= 0) return 256;
    return 256 + round_up((size_t)loss_ctas((int64_t)batch * num_anchors) * 2 * sizeof(double), 256);
}

static int multibox_loss_impl(const float* logits, const float* locs, const float* target, const uint8_t* sampled_mask,
                              int batch, int num_anchors, int num_cols, int kind, float gamma, float alpha,
                              float class_weight, float loc_weight, float* grad_logits, float* grad_locs,
                              float* loss_out, void* workspace, size_t workspace_bytes, void* stream,
                              const float* priors, float xy_scale, float wh_scale) {
    SSD_REQUIRE(batch >= 0 && num_anchors >= 0 && num_cols >= 1, SSD_ERR_INVALID_ARGUMENT, "ssd_multibox_loss: bad shape");
    SSD_REQUIRE(kind == SSD_LOSS_SOFTMAX_CE || kind == SSD_LOSS_SIGMOID_FOCAL, SSD_ERR_INVALID_ARGUMENT,
                "ssd_multibox_loss: unknown classification loss %d", kind);
    SSD_REQUIRE(loss_out != nullptr, SSD_ERR_INVALID_ARGUMENT, "ssd_multibox_loss: null loss_out");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = (int64_t)batch * num_anchors;
    if (rows == 0) {
        SSD_CUDA(cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), st));
        return SSD_OK;
    }
    SSD_REQUIRE(logits && locs && target && sampled_mask && workspace, SSD_ERR_INVALID_ARGUMENT,
                "ssd_multibox_loss: null pointer");
    SSD_REQUIRE(aligned(locs, 16) && aligned(target, 8) && (!grad_locs || aligned(grad_locs, 16)), SSD_ERR_MISALIGNED,
                "ssd_multibox_loss: locs / grad_locs need 16-byte, target 8-byte alignment");
    SSD_REQUIRE(aligned(workspace, 256), SSD_ERR_MISALIGNED, "ssd_multibox_loss: workspace must be 256-byte aligned");
    SSD_REQUIRE(workspace_bytes >= ssd_multibox_loss_workspace_bytes(batch, num_anchors), SSD_ERR_WORKSPACE,
                "ssd_multibox_loss: workspace too small");
    int* n_pos = (int*)workspace;
    double* partials = (double*)((unsigned char*)workspace + 256);
    SSD_CUDA(cudaMemsetAsync(n_pos, 0, 2 * sizeof(int), st));
    {
        LaunchTimer lt_("loss_count", st);
        const int blocks = (int)((rows + 4 * kLossThreads - 1) / (4 * kLossThreads));
        SSD_CUDA(launch_pdl(loss_count_kernel, dim3(blocks), dim3(kLossThreads), 0, st, target, sampled_mask, rows, n_pos));
        count_launch();
    }
    LossArgs a;
    a.A = num_anchors; a.C = num_cols; a.kind = kind; a.gamma = gamma; a.alpha = alpha;
    a.class_weight = class_weight; a.loc_weight = loc_weight; a.rows = rows;
    a.priors = reinterpret_cast<const float4*>(priors); a.xy_scale = xy_scale; a.wh_scale = wh_scale;
    const int ctas = loss_ctas(rows);
    {
        LaunchTimer lt_("loss_rows", st);
        SSD_CUDA(launch_pdl(loss_rows_kernel, dim3(ctas), dim3(kLossThreads), 0, st, a, logits, locs, target, sampled_mask,
                            (const int*)n_pos, grad_logits, grad_locs, partials));
        count_launch();
    }
    {
        LaunchTimer lt_("loss_finish", st);
        SSD_CUDA(launch_pdl(loss_finish_kernel, dim3(1), dim3(1024), 0, st, (const double*)partials, ctas,
                            (const int*)n_pos, kind, class_weight, loc_weight, loss_out));
        count_launch();
    }
    return SSD_OK;
}

extern "C" int ssd_multibox_loss(const float* logits, const float* locs, const float* target, const uint8_t* sampled_mask,
                                 int batch, int num_anchors, int num_cols, int kind, float gamma, float alpha,
                                 float class_weight, float loc_weight, float* grad_logits, float* grad_locs,
                                 float* loss_out, void* workspace, size_t workspace_bytes, void* stream) {
    return multibox_loss_impl(logits, locs, target, sampled_mask, batch, num_anchors, num_cols, kind, gamma, alpha,
                              class_weight, loc_weight, grad_logits, grad_locs, loss_out, workspace, workspace_bytes,
                              stream, nullptr, 1.f, 1.f);
}

extern "C" int ssd_multibox_loss_giou(const float* logits, const float* locs, const float* target, const float* priors,
                                      const uint8_t* sampled_mask, int batch, int num_anchors, int num_cols, int kind,
                                      float gamma, float alpha, float class_weight, float loc_weight, float xy_scale,
                                      float wh_scale, float* grad_logits, float* grad_locs, float* loss_out,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    SSD_REQUIRE(priors != nullptr && aligned(priors, 16), SSD_ERR_INVALID_ARGUMENT,
                "ssd_multibox_loss_giou: priors must be a 16-byte aligned device pointer");
    return multibox_loss_impl(logits, locs, target, sampled_mask, batch, num_anchors, num_cols, kind, gamma, alpha,
                              class_weight, loc_weight, grad_logits, grad_locs, loss_out, workspace, workspace_bytes,
                              stream, priors, xy_scale, wh_scale);
}

SSD_DEFINE_TRACE_SETTER(set_trace_loss)
