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

__global__ void probe_kernel(int* out) { *out = 100; }

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
