// errors.cu -- thread-local error string, version, and the per-device caches of libaurdf (the only process-wide
// state of the library: immutable facts about a device and "this never-changing kernel attribute has been set").
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "common.cuh"

namespace aurdf {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return AURDF_ECUDA;
}

namespace {
constexpr int kMaxDev = 64, kSlots = 16;
std::atomic<int> g_sms[kMaxDev];
std::atomic<unsigned char> g_once[kSlots][kMaxDev];
}  // namespace

int current_device_sms(int *device) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    if (device) *device = dev;
    int sms = dev >= 0 && dev < kMaxDev ? g_sms[dev].load(std::memory_order_relaxed) : 0;
    if (sms <= 0) {
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = kNumSMs;
        if (dev >= 0 && dev < kMaxDev) g_sms[dev].store(sms, std::memory_order_relaxed);
    }
    return sms;
}

// true exactly when the caller should perform the (idempotent) one-time action for this slot and device now
bool once_per_device(int slot, int device) {
    if (slot < 0 || slot >= kSlots || device < 0 || device >= kMaxDev) return true;
    return g_once[slot][device].exchange(1, std::memory_order_relaxed) == 0;
}
}  // namespace aurdf

extern "C" int aurdf_version(void) { return AURDF_VERSION; }
extern "C" const char *aurdf_last_error_string(void) { return aurdf::g_err; }
