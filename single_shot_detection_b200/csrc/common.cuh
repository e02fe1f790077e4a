// Shared device/host helpers for the sm_100a anchor-pipeline kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ssd_b200.h"

namespace ssd {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SSD_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            ::ssd::set_error(__VA_ARGS__);      \
            return (code);                      \
        }                                       \
    } while (0)

#define SSD_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return ::ssd::cuda_fail(e__, #call); \
    } while (0)

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
__host__ __device__ inline size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
int sm_count();
int stream_ctas_override();        // 0 = automatic (ssd_b200_set_stream_ctas_per_sm / SSD_CTAS_PER_SM)
cudaError_t zero_async(void* p, size_t bytes, cudaStream_t st);   // scratch zeroing as a kernel node (abi.cu)
void count_launch(int n = 1);      // bumps the library-wide kernel launch counter (ssd_b200_launch_count)
// diagnostics: per-launch CUDA-event timing when enabled through ssd_b200_timing_enable()
int timing_begin(const char* label, cudaStream_t st);
void timing_end(int slot, cudaStream_t st);
struct LaunchTimer {
    int slot;
    cudaStream_t st;
    LaunchTimer(const char* label, cudaStream_t s) : slot(timing_begin(label, s)), st(s) {}
    ~LaunchTimer() { timing_end(slot, st); }
};

// ---------------------------------------------------------------------------------------------
// exact fp32 arithmetic: every op separately rounded (the file is also compiled with -fmad=false)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// The loss route's box coding (shared by boxes.cu and the fused assignment): box_utils.to_centroids(inplace=True)
// followed by BoxCoder.encode_box(inplace=True), in the reference's in-place operation order.
struct Box { float a, b, c, d; };
__device__ __forceinline__ Box to_centroids_inplace(Box m) {             // bf/utils/box_utils.py:33-34
    const float w = fsub(m.c, m.a), h = fsub(m.d, m.b);
    return {fadd(m.a, fmul(w, 0.5f)), fadd(m.b, fmul(h, 0.5f)), w, h};
}
__device__ __forceinline__ Box encode_inplace(Box b, float4 p, float xy, float wh, float eps) {  // detection/box_coder.py:22-29
    return {fmul(fdiv(fsub(b.a, p.x), p.z), xy), fmul(fdiv(fsub(b.b, p.y), p.w), xy),
            fmul(logf(fadd(fdiv(b.c, p.z), eps)), wh), fmul(logf(fadd(fdiv(b.d, p.w), eps)), wh)};
}

// Monotone map float -> uint32 (a < b  <=>  key(a) < key(b) for non-NaN), NaN -> top.
// Keys of real values lie in [0x007FFFFF (-inf), 0xFF800000 (+inf)].
__device__ __forceinline__ uint32_t ordered_key(float v) {
    uint32_t u = __float_as_uint(v);
    if (v != v) return 0xFFFFFFFEu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------
// device-side timeline (diagnostics): when a trace buffer is installed (ssd_b200_trace_enable),
// every kernel records min(start) / max(end) of %globaltimer over its warps in its slot, so the
// real overlap of the launches inside a replayed CUDA graph can be read back.  One pointer per
// translation unit in constant memory (no relocatable device code); a null pointer costs one
// constant load and a branch.
// ---------------------------------------------------------------------------------------------
enum TraceSlot {
    TR_ASSIGN = 0, TR_MINING_LOSS, TR_MINING_KEYS, TR_MINING_SELECT, TR_PASS1, TR_GATE, TR_PASS2, TR_NMS, TR_TOPK,
    TR_MISC, TR_BOX0 = 10, kTraceSlots = 24
};
static __constant__ unsigned long long* tu_trace_buf = nullptr;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct KernelTrace {
    unsigned long long* p;
    __device__ __forceinline__ explicit KernelTrace(int slot) {
        p = tu_trace_buf;
        if (p != nullptr) {
            p += 2 * slot;
            if ((threadIdx.x & 31) == 0) atomicMin(p, global_ns());
        }
    }
    // phase marks of one CTA (the middle one of the grid): slot 16 + k holds {time, 0}
    __device__ __forceinline__ void mark(int k) const {
        if (p != nullptr && threadIdx.x == 0 && blockIdx.x == gridDim.x / 2 && k < 8) {
            unsigned long long* q = tu_trace_buf + 2 * (16 + k);
            q[0] = global_ns();
            q[1] = 1;
        }
    }
    __device__ __forceinline__ ~KernelTrace() {
        if (p != nullptr && (threadIdx.x & 31) == 0) atomicMax(p + 1, global_ns());
    }
};
#define SSD_DEFINE_TRACE_SETTER(name)                                                            \
    namespace ssd {                                                                              \
    cudaError_t name(unsigned long long* buf) { return cudaMemcpyToSymbol(tu_trace_buf, &buf, sizeof(buf)); } \
    }

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch: every kernel of this library is launched with the
// programmatic-stream-serialization attribute and starts with griddep_wait(), so the launch
// latency and prologue of kernel N+1 overlap the tail of kernel N (also inside captured graphs).
// griddep_wait() returns once the preceding grid in the stream has completed and flushed; since
// every kernel waits before it finishes, completion is transitive along the chain.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// WHERE a kernel triggers its dependents matters (tools/graph_timeline.py): the block scheduler places the
// CTAs of concurrently pending grids in arrival order, so a dependent that was launched at the START of
// its predecessor and cannot become fully resident yet (the NMS grid needs the shared memory pass 2 still
// holds) blocks every grid that arrives after it -- the box transforms and the sampler's selection sat
// behind it for ~14 us.  The long kernels therefore trigger when a CTA has finished its main loop: the
// dependent's launch latency still overlaps the predecessor's tail, but nothing queues behind a grid
// that cannot run.
// A dependent launched early holds its thread slots / shared memory while it sits in griddep_wait():
// small kernels with large grids that hang off a long-running predecessor (the box transforms after
// the target assignment: 1092 CTAs x 256 threads) were observed to keep the post-processor's first
// pass off the SMs for the whole assignment.  Those use launch_plain (full stream-order dependency).
// Every kernel of the library asks for the same shared-memory carve-out (the maximum: the streaming
// kernels need 3 x 72 KB per SM).  Kernels with different carve-outs cannot share an SM, and switching
// costs a reconfiguration: with mixed preferences the post-processor's first pass started ~5 us after
// the zeroing kernel in front of it had finished (tools/graph_timeline.py).
template <typename F>
inline void prefer_max_shared(F kern) {
    static_cast<void>(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_plain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    prefer_max_shared(kern);
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    prefer_max_shared(kern);
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

template <int WIDTH>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, 1-D) + mbarrier: global -> shared staging for the logit streams
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// the same on shared-space addresses (no generic -> shared conversion at the point of use)
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// L2 eviction policies for the bulk copies
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

}  // namespace ssd
