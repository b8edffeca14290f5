// ABI bookkeeping: version, thread-local error message, device check, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace hpcs {

std::atomic<uint64_t> g_launches{0};

char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

}  // namespace hpcs

extern "C" {

int hpcs_abi_version(void) { return HPCS_ABI_VERSION; }

const char* hpcs_last_error(void) { return hpcs::last_error_buf(); }

uint64_t hpcs_launch_count(void) { return hpcs::g_launches.load(); }

int hpcs_device_check(void) {
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return hpcs::fail(HPCS_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10)
        return hpcs::fail(HPCS_ERR_DEVICE, "device %d has compute capability %d.x; this library is sm_100a only", dev, major);
    return HPCS_OK;
}

}  // extern "C"
