// Shared helpers for the keypoints_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/keypoints_b200.h"

typedef __nv_bfloat16 bf16;

void kp_set_error(const char* fmt, ...);

#define KP_CHECK_ARG(cond, ...)                          \
    do {                                                 \
        if (!(cond)) {                                   \
            kp_set_error(__VA_ARGS__);                   \
            return KP_ERR_ARG;                           \
        }                                                \
    } while (0)

#define KP_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            kp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return KP_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

#define KP_LAUNCH_CHECK() KP_CUDA(cudaGetLastError())

// device-side view (strides in elements)
template <typename T>
struct View {
    T* p;
    long long sn, sy, sx, sc;
    __device__ __forceinline__ T* at(int n, int y, int x, int c) const {
        return p + n * sn + y * sy + x * sx + c * sc;
    }
};
template <typename T>
static inline View<T> make_view(const kp_view* v) {
    View<T> r;
    r.p = (T*)v->ptr; r.sn = v->sn; r.sy = v->sy; r.sx = v->sx; r.sc = v->sc;
    return r;
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void from_f(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// V contiguous channels -> float registers (V = 1 or 8; V=8 requires 16-byte alignment)
template <typename T, int V> struct Vec;
template <typename T> struct Vec<T, 1> {
    static __device__ __forceinline__ void load(const T* p, float (&f)[1]) { f[0] = to_f(*p); }
    static __device__ __forceinline__ void store(T* p, const float (&f)[1]) { from_f(p, f[0]); }
};
template <> struct Vec<float, 8> {
    struct Raw { float4 a, b; };
    static __device__ __forceinline__ Raw load_raw(const float* p) {
        Raw r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r;
    }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&f)[8]) {
        f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w; f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
    }
    static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
};
template <> struct Vec<bf16, 8> {
    struct Raw { uint4 a; };
    static __device__ __forceinline__ Raw load_raw(const bf16* p) { Raw r; r.a = *reinterpret_cast<const uint4*>(p); return r; }
    static __device__ __forceinline__ void unpack(const Raw& r, float (&f)[8]) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.a);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    static __device__ __forceinline__ void load(const bf16* p, float (&f)[8]) {
        uint4 u = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    static __device__ __forceinline__ void store(bf16* p, const float (&f)[8]) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};

static inline bool view_vec8_ok(const kp_view* v, int C) {
    size_t es = v->dtype == KP_BF16 ? 2 : 4;
    return v->sc == 1 && (C % 8) == 0 && (v->sn % 8) == 0 && (v->sy % 8) == 0 && (v->sx % 8) == 0 &&
           (((uintptr_t)v->ptr) % 16) == 0 && es != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float apply_act(float z, int act) {
    if (act == KP_ACT_LEAKY) return z > 0.f ? z : 0.01f * z;
    if (act == KP_ACT_RELU) return z > 0.f ? z : 0.f;
    return z;
}
__device__ __forceinline__ float act_grad(float z, int act) {
    if (act == KP_ACT_LEAKY) return z > 0.f ? 1.f : 0.01f;
    if (act == KP_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    return 1.f;
}

int kp_sm_count();

// Experiment switches (A/B of kernel generations on the GPU box): compiled in only with `make EXPERIMENTS=1`
// (-DKP_EXPERIMENTS).  The product library never reads the environment: every switch takes its default.
#ifdef KP_EXPERIMENTS
#include <stdlib.h>
static inline const char* kp_env(const char* name) { return getenv(name); }
#else
static inline const char* kp_env(const char*) { return nullptr; }
#endif

// kp_ctx (kp_api.cu): per-device state a caller may bind to its thread - SM budget of the persistent kernels and a cache of
// encoded TMA tensor maps.  The tensor-core launchers ask for maps through kp_ctx_map_{get,put}; without a current
// context every call encodes its maps afresh (the behaviour before contexts existed).
struct KpMapKey {
    const void* ptr;
    long long d0, d1, d2, d3;     // dims (2-D: rows, cols, box_rows, -1; 4-D: N, PH*65536+PW, C, bx*256+by)
    bool operator==(const KpMapKey& o) const { return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3; }
};
bool kp_ctx_map_get(const KpMapKey& key, void* map128);        // true: the 128-byte CUtensorMap was copied out of the cache
void kp_ctx_map_put(const KpMapKey& key, const void* map128);

// per-device one-time setup guard (cudaFuncSetAttribute is per device; one process may drive several GPUs)
struct KpOncePerDevice {
    unsigned long long mask = 0;
    bool first() {
        int d = 0;
        cudaGetDevice(&d);
        const unsigned long long b = 1ull << (d & 63);
        if (mask & b) return false;
        mask |= b;
        return true;
    }
};

// mma.sync kernels for the first (Cin <= 3) convolution of a Unit in bf16 mode (kp_conv_thin_mma.cu)
bool kp_thin_mma_fprop_ok(const kp_view* in, const kp_view* out, int OH, int OW, int IH, int IW, int Cin, int Cout, int ks,
                          int off);
int kp_thin_mma_fprop(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out,
                      double* stats, int N, int OH, int OW, int Cin, int Cout);
bool kp_thin_mma_wgrad_ok(const kp_view* x, const kp_view* dy, int W, int Cin, int Cout, int ks);
int kp_thin_mma_wgrad(cudaStream_t st, const kp_view* x, const kp_view* dy, float* dw, int N, int H, int W, int Cin,
                      int Cout);
// mma.sync kernels for 3x3 convolutions with 16 / 32 input and output channels in bf16 mode (kp_conv_small_mma.cu)
bool kp_small_mma_conv_ok(const kp_view* in, const kp_view* out, int N, int OH, int OW, int Cin, int Cout, int ks);
int kp_small_mma_conv(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out,
                      double* stats, int N, int OH, int OW, int IH, int IW, int Cin, int Cout, int ks, int off);
bool kp_small_mma_wgrad_ok(const kp_view* x, const kp_view* dy, int N, int H, int W, int Cin, int Cout, int ks);
int kp_small_mma_wgrad(cudaStream_t st, const kp_view* x, const kp_view* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                       int ks);
// single-channel first layer (1 -> 8..64 channels, 3x3) in bf16 mode (kp_conv_small_mma.cu)
bool kp_c1_conv_ok(const kp_view* in, const kp_view* out, int OH, int OW, int IH, int IW, int Cin, int Cout, int ks, int off);
int kp_c1_fprop(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out, double* stats, int N,
                int OH, int OW, int Cout);
bool kp_c1_wgrad_ok(const kp_view* x, const kp_view* dy, int Cin, int Cout, int ks);
int kp_c1_wgrad(cudaStream_t st, const kp_view* x, const kp_view* dy, float* dw, int N, int H, int W, int Cout);
// streaming kernels for a 1x1 head with <= 4 outputs on a 64..256-channel bf16 input (kp_conv_thin_mma.cu)
bool kp_head_mma_fprop_ok(const kp_view* in, const kp_view* out, int N, int H, int W, int Cw, int Ct);
int kp_head_mma_fprop(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out, int N, int H,
                      int W, int Cw, int Ct);
bool kp_head1x1_ok(const kp_view* wide, const kp_view* thin, int Cw, int Ct);
int kp_head1x1_fprop(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out, int N, int H,
                     int W, int Cw, int Ct);
int kp_head1x1_dgrad(cudaStream_t st, const kp_view* dy, const float* wd, const kp_view* dx, int N, int H, int W, int Cw, int Ct);
int kp_head1x1_wgrad(cudaStream_t st, const kp_view* x, const kp_view* dy, float* dw, int N, int H, int W, int Cw, int Ct);
