// Library-level pieces of the C ABI: error text, version, device check.
#include <stdarg.h>

#include "common.cuh"

namespace ssd {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return SSD_ERR_CUDA;
}

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// Resident CTAs per SM of the logit-streaming kernels: 0 = as many as shared memory allows (best for a step that
// runs alone), 1 = one per SM, which leaves shared memory for the NMS / selection CTAs of OTHER steps when
// several step graphs are in flight (bench.py: 34.6 -> 32.2 us per step with four in flight, 64 -> 68 us alone).
// Read when a launch is issued or captured.
int g_nms_threads = 0;             // ssd_b200_set_nms_threads: 0 = default (SSD_NMS_THREADS or 128)
static int g_stream_ctas = -1;
int stream_ctas_override() {
    if (g_stream_ctas < 0) {
        const char* e = getenv("SSD_CTAS_PER_SM");
        g_stream_ctas = (e && e[0]) ? atoi(e) : 0;
    }
    return g_stream_ctas;
}

// Fused candidate selection of the post-processor (postprocess.cu: fused_select_kernel): 0 = off (the streaming
// two-pass path; the default: measured on B200 the cluster kernel is not faster), -1 = automatic (images
// that fit the shared memory of one thread-block cluster, all clusters resident at once), 1/2/4/8 = this cluster size
// when it fits.  SSD_FUSED_SELECT presets it; read when a launch is planned.
static int g_fused_select = -2;
int fused_select_mode() {
    if (g_fused_select == -2) {
        const char* e = getenv("SSD_FUSED_SELECT");
        g_fused_select = (e && e[0]) ? atoi(e) : 0;
    }
    return g_fused_select;
}

__global__ void probe_kernel(int* out) { *out = 100; }

// Scratch zeroing as a KERNEL node: inside a captured step graph a memset node in front of a branch was
// observed (tools/graph_timeline.py) to queue behind the kernels of an unrelated branch -- the
// post-processor's first pass then started only after the target assignment had finished.  A kernel
// launched with programmatic stream serialization has no such coupling and overlaps its successor's
// prologue.
__global__ void __launch_bounds__(256) zero_kernel(uint4* __restrict__ p, size_t n16) {
    KernelTrace trace_(n16 > 4096 ? TR_MISC : TR_MINING_KEYS);     // (diagnostics: large / small zeroing)
    griddep_wait();
    griddep_launch_dependents();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(0u, 0u, 0u, 0u);
}
cudaError_t zero_async(void* p, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return cudaSuccess;
    if (!aligned(p, 16) || (bytes & 15)) return cudaMemsetAsync(p, 0, bytes, st);
    const size_t n16 = bytes / 16;
    size_t blocks = (n16 + 255) / 256;
    if (blocks > 592) blocks = 592;
    const cudaError_t e = launch_pdl(zero_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (uint4*)p, n16);
    if (e == cudaSuccess) count_launch();
    return e;
}

// ---- optional per-launch timing (diagnostics; off by default, not capturable into a graph) ----
constexpr int kTimingSlots = 8192;
static bool g_timing = false;
static int g_timing_used = 0;
static cudaEvent_t g_ev[kTimingSlots][2];
static const char* g_label[kTimingSlots];
static bool g_ev_ready = false;

int timing_begin(const char* label, cudaStream_t st) {
    if (!g_timing || g_timing_used >= kTimingSlots) return -1;
    if (!g_ev_ready) {
        for (int i = 0; i < kTimingSlots; ++i) { cudaEventCreate(&g_ev[i][0]); cudaEventCreate(&g_ev[i][1]); }
        g_ev_ready = true;
    }
    const int i = g_timing_used++;
    g_label[i] = label;
    cudaEventRecord(g_ev[i][0], st);
    return i;
}
void timing_end(int slot, cudaStream_t st) {
    if (slot >= 0) cudaEventRecord(g_ev[slot][1], st);
}

}  // namespace ssd

extern "C" int ssd_b200_abi_version(void) { return SSD_B200_ABI_VERSION; }

extern "C" unsigned long long ssd_b200_launch_count(void) { return __atomic_load_n(&ssd::g_launches, __ATOMIC_RELAXED); }

extern "C" const char* ssd_b200_last_error(void) { return ssd::g_error; }

extern "C" int ssd_b200_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        ssd::set_error("no CUDA device: %s", cudaGetErrorString(e));
        return SSD_ERR_NO_DEVICE;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        ssd::set_error("device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
        return SSD_ERR_NO_DEVICE;
    }
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, ssd::probe_kernel);
    if (e != cudaSuccess) {
        ssd::set_error("sm_100a kernel image not loadable: %s", cudaGetErrorString(e));
        return SSD_ERR_NO_DEVICE;
    }
    return SSD_OK;
}

namespace ssd {
cudaError_t set_trace_assign(unsigned long long*);
cudaError_t set_trace_boxes(unsigned long long*);
cudaError_t set_trace_mining(unsigned long long*);
cudaError_t set_trace_postprocess(unsigned long long*);
cudaError_t set_trace_loss(unsigned long long*);
cudaError_t set_trace_metrics(unsigned long long*);
cudaError_t set_trace_anchors(unsigned long long*);
cudaError_t set_trace_exchange(unsigned long long*);
cudaError_t set_trace_abi(unsigned long long*);
}  // namespace ssd

extern "C" int ssd_b200_trace_enable(unsigned long long* device_slots) {
    SSD_CUDA(ssd::set_trace_assign(device_slots));
    SSD_CUDA(ssd::set_trace_boxes(device_slots));
    SSD_CUDA(ssd::set_trace_mining(device_slots));
    SSD_CUDA(ssd::set_trace_postprocess(device_slots));
    SSD_CUDA(ssd::set_trace_loss(device_slots));
    SSD_CUDA(ssd::set_trace_metrics(device_slots));
    SSD_CUDA(ssd::set_trace_anchors(device_slots));
    SSD_CUDA(ssd::set_trace_exchange(device_slots));
    SSD_CUDA(ssd::set_trace_abi(device_slots));
    return SSD_OK;
}
extern "C" int ssd_b200_trace_slots(void) { return ssd::kTraceSlots; }

extern "C" void ssd_b200_timing_enable(int on) {
    ssd::g_timing = on != 0;
    ssd::g_timing_used = 0;
}

extern "C" size_t ssd_b200_timing_report(char* buf, size_t capacity) {
    using namespace ssd;
    cudaDeviceSynchronize();
    struct Acc { const char* label; double us; int n; };
    Acc acc[64];
    int na = 0;
    for (int i = 0; i < g_timing_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_ev[i][0], g_ev[i][1]) != cudaSuccess) continue;
        int j = 0;
        for (; j < na; ++j) if (acc[j].label == g_label[i] || strcmp(acc[j].label, g_label[i]) == 0) break;
        if (j == na) { if (na == 64) continue; acc[na++] = {g_label[i], 0.0, 0}; }
        acc[j].us += 1e3 * ms;
        acc[j].n += 1;
    }
    size_t off = 0;
    for (int j = 0; j < na && buf && off + 96 < capacity; ++j)
        off += (size_t)snprintf(buf + off, capacity - off, "%s%s:%.2f:%d", j ? "," : "", acc[j].label, acc[j].us / acc[j].n, acc[j].n);
    if (buf && capacity) buf[off < capacity ? off : capacity - 1] = 0;
    g_timing_used = 0;
    return off;
}

SSD_DEFINE_TRACE_SETTER(set_trace_abi)

extern "C" int ssd_b200_set_fused_select(int mode) {
    SSD_REQUIRE(mode == -1 || mode == 0 || mode == 1 || mode == 2 || mode == 4 || mode == 8, SSD_ERR_INVALID_ARGUMENT,
                "ssd_b200_set_fused_select: %d is not one of -1, 0, 1, 2, 4, 8", mode);
    ssd::g_fused_select = mode;
    return SSD_OK;
}

extern "C" int ssd_b200_set_nms_threads(int threads) {
    SSD_REQUIRE(threads == 0 || threads == 32 || threads == 64 || threads == 128 || threads == 256, SSD_ERR_INVALID_ARGUMENT,
                "ssd_b200_set_nms_threads: %d is not one of 0, 32, 64, 128, 256", threads);
    ssd::g_nms_threads = threads;
    return SSD_OK;
}

extern "C" int ssd_b200_set_stream_ctas_per_sm(int ctas) {
    SSD_REQUIRE(ctas >= 0 && ctas <= 8, SSD_ERR_INVALID_ARGUMENT, "ssd_b200_set_stream_ctas_per_sm: %d outside 0..8", ctas);
    ssd::g_stream_ctas = ctas;
    return SSD_OK;
}
