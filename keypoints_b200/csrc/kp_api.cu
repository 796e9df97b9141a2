// Error plumbing and device queries behind the C ABI (include/keypoints_b200.h).
#include "kp_common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void kp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int kp_sm_count() {
    static int cached[64] = {0};          // per device: one process may drive several GPUs
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cached[dev & 63]) return cached[dev & 63];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached[dev & 63] = n;
    return n;
}

extern "C" const char* kp_last_error(void) { return g_err; }
extern "C" int kp_version(void) { return 100; }

extern "C" int kp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0, n = 0, ma = 0, mi = 0;
    KP_CUDA(cudaGetDevice(&dev));
    KP_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    KP_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    KP_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    if (ma != 10) {
        kp_set_error("keypoints_b200 is built for sm_100a only; device reports sm_%d%d", ma, mi);
        return KP_ERR_UNSUPPORTED;
    }
    return KP_OK;
}
