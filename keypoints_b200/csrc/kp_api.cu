// Error plumbing and device queries behind the C ABI (include/keypoints_b200.h).
#include "kp_common.cuh"
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <unordered_map>

static thread_local char g_err[512] = "";

void kp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct KpMapKeyHash {
    size_t operator()(const KpMapKey& k) const {
        size_t h = (size_t)k.ptr;
        for (long long v : {k.d0, k.d1, k.d2, k.d3}) h = h * 1000003u ^ (size_t)v;
        return h;
    }
};
struct KpMap128 { unsigned char b[128]; };

struct kp_ctx {
    int device = 0;
    int sm_count = 0;
    int sm_limit = 0;                    // 0 = all SMs
    long long hits = 0, misses = 0;
    std::mutex mu;
    std::unordered_map<KpMapKey, KpMap128, KpMapKeyHash> maps;
};
static thread_local kp_ctx* g_ctx = nullptr;

bool kp_ctx_map_get(const KpMapKey& key, void* map128) {
    kp_ctx* c = g_ctx;
    if (!c) return false;
    std::lock_guard<std::mutex> lock(c->mu);
    auto it = c->maps.find(key);
    if (it == c->maps.end()) { ++c->misses; return false; }
    memcpy(map128, it->second.b, 128);
    ++c->hits;
    return true;
}
void kp_ctx_map_put(const KpMapKey& key, const void* map128) {
    kp_ctx* c = g_ctx;
    if (!c) return;
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->maps.size() >= 4096) c->maps.clear();       // a training step uses a few hundred; bound the table anyway
    KpMap128 m;
    memcpy(m.b, map128, 128);
    c->maps[key] = m;
}

extern "C" int kp_ctx_create(kp_ctx** ctx) {
    KP_CHECK_ARG(ctx, "kp_ctx_create: null argument");
    int dev = 0, n = 0, ma = 0;
    KP_CUDA(cudaGetDevice(&dev));
    KP_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    KP_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    if (ma != 10) {
        kp_set_error("kp_ctx_create: keypoints_b200 is built for sm_100a only; device %d reports cc major %d", dev, ma);
        return KP_ERR_UNSUPPORTED;
    }
    kp_ctx* c = new kp_ctx();
    c->device = dev;
    c->sm_count = n;
    *ctx = c;
    return KP_OK;
}
extern "C" int kp_ctx_destroy(kp_ctx* ctx) {
    if (g_ctx == ctx) g_ctx = nullptr;
    delete ctx;
    return KP_OK;
}
extern "C" int kp_ctx_set_current(kp_ctx* ctx) {
    g_ctx = ctx;
    return KP_OK;
}
extern "C" int kp_ctx_set_sm_limit(kp_ctx* ctx, int sms) {
    KP_CHECK_ARG(ctx && sms >= 0, "kp_ctx_set_sm_limit: bad arguments");
    ctx->sm_limit = (sms == 0 || sms >= ctx->sm_count) ? 0 : (sms < 2 ? 2 : sms);
    return KP_OK;
}
extern "C" int kp_ctx_info(kp_ctx* ctx, int* device, int* sm_count, int* sm_limit, int64_t* map_hits, int64_t* map_misses,
                           int64_t* maps_cached) {
    KP_CHECK_ARG(ctx, "kp_ctx_info: null context");
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (device) *device = ctx->device;
    if (sm_count) *sm_count = ctx->sm_count;
    if (sm_limit) *sm_limit = ctx->sm_limit;
    if (map_hits) *map_hits = ctx->hits;
    if (map_misses) *map_misses = ctx->misses;
    if (maps_cached) *maps_cached = (int64_t)ctx->maps.size();
    return KP_OK;
}

int kp_sm_count() {
    if (g_ctx && g_ctx->sm_limit > 0) return g_ctx->sm_limit;
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
