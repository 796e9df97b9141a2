// BatchNorm(train) + activation + pool/upsample + replicate-pad, forward and backward.
// HBM-bound passes: channel-fastest thread mapping, 16-byte vector accesses (8 channels / thread)
// whenever the views allow it, per-thread register accumulation then one double atomic per
// channel per block for the reductions.
#include "kp_common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace {

template <int V> using IC = std::integral_constant<int, V>;

template <typename F>
int dispatch1(int dt, F&& f) {
    if (dt == KP_F32) return f(float{});
    if (dt == KP_BF16) return f(bf16{});
    kp_set_error("bad dtype %d", dt);
    return KP_ERR_ARG;
}

struct Launch2D {
    dim3 grid, block;
};
static Launch2D plan2d(long long pixels, int C, int V) {
    int ncg = (C + V - 1) / V;
    int bx = 1;
    while (bx < ncg && bx < 32) bx <<= 1;
    int by = 256 / bx;
    int gy = (ncg + bx - 1) / bx;
    long long gx = (pixels + by - 1) / by;
    long long cap = (long long)kp_sm_count() * 8 / gy;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    Launch2D l;
    l.grid = dim3((unsigned)gx, (unsigned)gy, 1);
    l.block = dim3(bx, by, 1);
    return l;
}

// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_k(const double* __restrict__ stats, int C, double count, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean,
                              float* running_var, long long* nbt, float* scale, float* shift,
                              float* save_mean, float* save_invstd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= C) return;
    double mean = stats[c] / count;
    double var = stats[C + c] / count - mean * mean;
    if (var < 0) var = 0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;
    if (save_mean) save_mean[c] = (float)mean;
    if (save_invstd) save_invstd[c] = invstd;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
        double unb = count > 1 ? var * count / (count - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

__global__ void bn_grad_finalize_k(const double* __restrict__ sums, int C, float* dgamma, float* dbeta) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (dbeta) dbeta[c] = (float)sums[c];
    if (dgamma) dgamma[c] = (float)sums[C + c];
}

// ------------------------------------------------------------------------------------------------
template <typename TI, int V>
__device__ __forceinline__ void load_act(const View<TI>& y, int n, int yy, int xx, int c0, const float (&sc)[V],
                                         const float (&sh)[V], int act, float (&v)[V]) {
    Vec<TI, V>::load(y.at(n, yy, xx, c0), v);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = apply_act(fmaf(v[i], sc[i], sh[i]), act);
}

template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256)
bn_act_fwd_k(View<TI> y, View<TO> out, const float* __restrict__ scale, const float* __restrict__ shift, int act,
             int post, int pad, int N, int H, int W, int C, int OH, int OW) {
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * V;
    if (c0 >= C) return;
    float sc[V], sh[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        sc[i] = scale ? scale[c0 + i] : 1.f;
        sh[i] = shift ? shift[c0 + i] : 0.f;
    }
    const int PH = OH + 2 * pad, PW = OW + 2 * pad;
    const long long P = (long long)N * PH * PW;
    const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.y + threadIdx.y; p < P; p += (long long)gridDim.x * blockDim.y) {
        int px, py, n;
        if (P < (1LL << 31)) {                         // 32-bit divisions
            const unsigned u = (unsigned)p, r = u / (unsigned)PW;
            px = (int)(u - r * (unsigned)PW);
            n = (int)(r / (unsigned)PH);
            py = (int)(r - (unsigned)n * (unsigned)PH);
        } else {
            px = (int)(p % PW);
            long long r = p / PW;
            py = (int)(r % PH);
            n = (int)(r / PH);
        }
        int oy = min(max(py - pad, 0), OH - 1), ox = min(max(px - pad, 0), OW - 1);
        float v[V];
        if (post == KP_POST_NONE) {
            load_act<TI, V>(y, n, oy, ox, c0, sc, sh, act, v);
        } else if (post == KP_POST_POOL) {
            float a[V];
            load_act<TI, V>(y, n, 2 * oy, 2 * ox, c0, sc, sh, act, v);
            load_act<TI, V>(y, n, 2 * oy, 2 * ox + 1, c0, sc, sh, act, a);
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = fmaxf(v[i], a[i]);
            load_act<TI, V>(y, n, 2 * oy + 1, 2 * ox, c0, sc, sh, act, a);
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = fmaxf(v[i], a[i]);
            load_act<TI, V>(y, n, 2 * oy + 1, 2 * ox + 1, c0, sc, sh, act, a);
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = fmaxf(v[i], a[i]);
        } else {   // bilinear x2, align_corners=True (vgg.py:26)
            float sy = ry * oy, sx = rx * ox;
            int y0 = (int)sy, x0 = (int)sx;
            float ly = sy - y0, lx = sx - x0;
            int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
            float a00[V], a01[V], a10[V], a11[V];
            load_act<TI, V>(y, n, y0, x0, c0, sc, sh, act, a00);
            load_act<TI, V>(y, n, y0, x1, c0, sc, sh, act, a01);
            load_act<TI, V>(y, n, y1, x0, c0, sc, sh, act, a10);
            load_act<TI, V>(y, n, y1, x1, c0, sc, sh, act, a11);
#pragma unroll
            for (int i = 0; i < V; ++i)
                v[i] = (1.f - ly) * ((1.f - lx) * a00[i] + lx * a01[i]) + ly * ((1.f - lx) * a10[i] + lx * a11[i]);
        }
        Vec<TO, V>::store(out.at(n, py, px, c0), v);
    }
}

// ------------------------------------------------------------------------------------------------
// gradient w.r.t. the (unpadded) pixel (oy,ox) of a replicate-padded buffer: acc += wgt * fold
template <typename TG, int V>
__device__ __forceinline__ void fold_read(const View<TG>& d, int pad, int n, int oy, int ox, int c0, int OH, int OW,
                                          float wgt, float (&acc)[V]) {
    float t[V];
    if (!pad) {
        Vec<TG, V>::load(d.at(n, oy, ox, c0), t);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = fmaf(wgt, t[i], acc[i]);
        return;
    }
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = oy + 1;
    if (oy == 0) ys[ny++] = 0;
    if (oy == OH - 1) ys[ny++] = OH + 1;
    xs[nx++] = ox + 1;
    if (ox == 0) xs[nx++] = 0;
    if (ox == OW - 1) xs[nx++] = OW + 1;
    for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) {
            Vec<TG, V>::load(d.at(n, ys[a], xs[b], c0), t);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = fmaf(wgt, t[i], acc[i]);
        }
}

// dz (gradient w.r.t. the BatchNorm output) and the raw conv output yv at pixel (n,yy,xx)
template <typename TG, typename TY, int V>
__device__ __forceinline__ void compute_dz(const View<TG>& dout, const View<TY>& y, const float (&sc)[V],
                                           const float (&sh)[V], int act, int post, int pad, int n, int yy, int xx,
                                           int c0, int H, int W, int OH, int OW, float ry, float rx, float (&g)[V],
                                           float (&yv)[V]) {
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = 0.f;
    Vec<TY, V>::load(y.at(n, yy, xx, c0), yv);
    if (post == KP_POST_NONE) {
        fold_read<TG, V>(dout, pad, n, yy, xx, c0, OH, OW, 1.f, g);
    } else if (post == KP_POST_POOL) {
        int wy = yy & ~1, wx = xx & ~1;
        if (wy + 1 < H && wx + 1 < W) {
            float best[V];
            int bi[V];
#pragma unroll
            for (int i = 0; i < V; ++i) { best[i] = -INFINITY; bi[i] = 0; }
            for (int q = 0; q < 4; ++q) {          // first maximum wins, row-major window order
                float a[V];
                load_act<TY, V>(y, n, wy + (q >> 1), wx + (q & 1), c0, sc, sh, act, a);
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (a[i] > best[i] || q == 0) { best[i] = a[i]; bi[i] = q; }
            }
            float t[V];
#pragma unroll
            for (int i = 0; i < V; ++i) t[i] = 0.f;
            fold_read<TG, V>(dout, pad, n, wy >> 1, wx >> 1, c0, OH, OW, 1.f, t);
            int mine = ((yy & 1) << 1) | (xx & 1);
#pragma unroll
            for (int i = 0; i < V; ++i) g[i] = (bi[i] == mine) ? t[i] : 0.f;
        }
    } else {   // bilinear x2 align_corners=True backward: gather every output row/col touching (yy,xx)
        int ylo = 0, yhi = OH - 1, xlo = 0, xhi = OW - 1;
        if (H > 1) {
            ylo = max(0, (int)floorf((yy - 1) / ry) - 1);
            yhi = min(OH - 1, (int)ceilf((yy + 1) / ry) + 1);
        }
        if (W > 1) {
            xlo = max(0, (int)floorf((xx - 1) / rx) - 1);
            xhi = min(OW - 1, (int)ceilf((xx + 1) / rx) + 1);
        }
        for (int Y = ylo; Y <= yhi; ++Y) {
            float sy = ry * Y;
            int y0 = (int)sy;
            float ly = sy - y0;
            int y1 = min(y0 + 1, H - 1);
            float wyv = (y0 == yy ? 1.f - ly : 0.f) + (y1 == yy ? ly : 0.f);
            if (wyv == 0.f) continue;
            for (int X = xlo; X <= xhi; ++X) {
                float sx = rx * X;
                int x0 = (int)sx;
                float lx = sx - x0;
                int x1 = min(x0 + 1, W - 1);
                float wxv = (x0 == xx ? 1.f - lx : 0.f) + (x1 == xx ? lx : 0.f);
                if (wxv == 0.f) continue;
                fold_read<TG, V>(dout, pad, n, Y, X, c0, OH, OW, wyv * wxv, g);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] *= act_grad(fmaf(yv[i], sc[i], sh[i]), act);
}

template <typename TG, typename TY, typename TD, int V>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_k(View<TG> dout, View<TY> y, View<TD> dy, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ invstd, double* sums, int act, int post, int pad, int N, int H, int W,
                    int C, int OH, int OW) {
    __shared__ float red[2][256 * V];
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * V;
    const bool active = c0 < C;
    float sc[V], sh[V], mu[V], is[V], s1[V], s2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        sc[i] = (active && scale) ? scale[c0 + i] : 1.f;
        sh[i] = (active && shift) ? shift[c0 + i] : 0.f;
        mu[i] = (active && mean) ? mean[c0 + i] : 0.f;
        is[i] = (active && invstd) ? invstd[c0 + i] : 1.f;
        s1[i] = 0.f; s2[i] = 0.f;
    }
    const long long P = (long long)N * H * W;
    const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    if (active) {
        for (long long p = (long long)blockIdx.x * blockDim.y + threadIdx.y; p < P;
             p += (long long)gridDim.x * blockDim.y) {
            int xx, yy, n;
            if (P < (1LL << 31)) {                     // 32-bit divisions
                const unsigned u = (unsigned)p, r = u / (unsigned)W;
                xx = (int)(u - r * (unsigned)W);
                n = (int)(r / (unsigned)H);
                yy = (int)(r - (unsigned)n * (unsigned)H);
            } else {
                xx = (int)(p % W);
                long long r = p / W;
                yy = (int)(r % H);
                n = (int)(r / H);
            }
            float g[V], yv[V];
            compute_dz<TG, TY, V>(dout, y, sc, sh, act, post, pad, n, yy, xx, c0, H, W, OH, OW, ry, rx, g, yv);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                s1[i] += g[i];
                s2[i] += g[i] * (yv[i] - mu[i]) * is[i];
            }
            if (dy.p) Vec<TD, V>::store(dy.at(n, yy, xx, c0), g);
        }
    }
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < V; ++i) { red[0][t * V + i] = s1[i]; red[1][t * V + i] = s2[i]; }
    __syncthreads();
    if (threadIdx.y == 0 && active) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float a = 0.f, b = 0.f;
            for (int j = 0; j < (int)blockDim.y; ++j) {
                a += red[0][(j * blockDim.x + threadIdx.x) * V + i];
                b += red[1][(j * blockDim.x + threadIdx.x) * V + i];
            }
            atomicAdd(&sums[c0 + i], (double)a);
            atomicAdd(&sums[C + c0 + i], (double)b);
        }
    }
}

// ================================================================================================
// Fast path: NHWC with unit channel stride, 8 channels (16 bytes) per thread, C/8 a power of two <= 256.
// One block walks whole image rows: 32-bit index math only, the channel group of a thread is fixed (so the
// per-channel affine lives in registers and the reductions accumulate in registers), consecutive threads
// touch consecutive 16-byte chunks of a row, and the inner loops are unrolled x4 with every load issued
// before the first use so each thread keeps 4-10 independent 16-byte requests in flight.
// ================================================================================================
constexpr int V8 = 8;
#ifndef KP_EW_UNR
#define KP_EW_UNR 4
#endif
#ifndef KP_EW_MINB
#define KP_EW_MINB 2
#endif
constexpr int UNR = KP_EW_UNR;

// Optional fused BatchNorm finalize: when stats != nullptr the forward kernels derive the per-channel affine from
// the batch statistics themselves (every block recomputes its 8 channels; block 0 publishes scale / shift / mean /
// invstd for the backward pass and updates the running statistics) — removes one dependent launch per layer.
struct BnFuse {
    const double* stats;
    double count;
    const float* gamma;
    const float* beta;
    float eps, momentum;
    float* rmean;
    float* rvar;
    long long* nbt;
    float* scale;
    float* shift;
    float* mean;
    float* invstd;
};

__device__ __forceinline__ void fused_affine(const BnFuse& f, int C, int c0, bool publish, float (&sc)[8], float (&sh)[8]) {
    // every thread of every block runs this: one double division, the rest double FMAs and fp32 (a double division or
    // square root per channel per thread costs ~15 us per launch on the fp64 pipe)
    const double rcount = 1.0 / f.count;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        const double mean = f.stats[c] * rcount;
        double var = f.stats[C + c] * rcount - mean * mean;
        if (var < 0) var = 0;
        const float invstd = 1.0f / sqrtf((float)var + f.eps);
        const float g = f.gamma ? f.gamma[c] : 1.f, b = f.beta ? f.beta[c] : 0.f;
        sc[i] = g * invstd;
        sh[i] = b - (float)mean * sc[i];
        if (publish) {
            f.scale[c] = sc[i];
            f.shift[c] = sh[i];
            if (f.mean) f.mean[c] = (float)mean;
            if (f.invstd) f.invstd[c] = invstd;
            if (f.rmean) f.rmean[c] = (1.f - f.momentum) * f.rmean[c] + f.momentum * (float)mean;
            if (f.rvar) {
                const double unb = f.count > 1 ? var * f.count / (f.count - 1) : var;
                f.rvar[c] = (1.f - f.momentum) * f.rvar[c] + f.momentum * (float)unb;
            }
            if (c == 0 && f.nbt) *f.nbt += 1;
        }
    }
}

#include "kp_bn_lean.cuh"
#include "kp_bn_pipe.cuh"

#define KP_ACT_SWITCH(actv_, CALL)                          \
    do {                                                    \
        if ((actv_) == KP_ACT_LEAKY) { CALL(KP_ACT_LEAKY); }  \
        else if ((actv_) == KP_ACT_RELU) { CALL(KP_ACT_RELU); } \
        else { CALL(KP_ACT_NONE); }                          \
    } while (0)

struct RowFold {   // rows / columns of a replicate-padded gradient that fold into one unpadded coordinate
    int n, v[3];
    __device__ __forceinline__ RowFold(int o, int O, int pad) {
        if (!pad) { n = 1; v[0] = o; return; }
        n = 0; v[n++] = o + 1;
        if (o == 0) v[n++] = 0;
        if (o == O - 1) v[n++] = O + 1;
    }
};

template <typename TG>
__device__ __forceinline__ void fold_acc(const View<TG>& d, int n, const RowFold& fy, const RowFold& fx, int c0, float wgt,
                                         float (&acc)[V8]) {
    float t[V8];
    for (int a = 0; a < fy.n; ++a)
        for (int b = 0; b < fx.n; ++b) {
            Vec<TG, V8>::load(d.at(n, fy.v[a], fx.v[b], c0), t);
#pragma unroll
            for (int i = 0; i < V8; ++i) acc[i] = fmaf(wgt, t[i], acc[i]);
        }
}

template <typename T>
__device__ __forceinline__ void raw_act(const typename Vec<T, V8>::Raw& r, const float (&sc)[V8], const float (&sh)[V8],
                                        int act, float (&v)[V8]) {
    Vec<T, V8>::unpack(r, v);
#pragma unroll
    for (int i = 0; i < V8; ++i) v[i] = apply_act(fmaf(v[i], sc[i], sh[i]), act);
}

template <typename TI, typename TO, int POST>
__global__ void __launch_bounds__(256, KP_EW_MINB)
bn_act_fwd_rows_k(View<TI> y, View<TO> out, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                  int pad, int N, int H, int W, int C, int OH, int OW, int cg_shift, const BnFuse fuse) {
    using RawI = typename Vec<TI, V8>::Raw;
    const int ncg = C >> 3;
    const int cg = threadIdx.x & (ncg - 1), c0 = cg * V8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float sc[V8], sh[V8];
    if (fuse.stats) fused_affine(fuse, C, c0, blockIdx.x == 0 && x0 == 0, sc, sh);
    else {
#pragma unroll
        for (int i = 0; i < V8; ++i) { sc[i] = scale ? scale[c0 + i] : 1.f; sh[i] = shift ? shift[c0 + i] : 0.f; }
    }
    const int PH = OH + 2 * pad, PW = OW + 2 * pad, rows = N * PH;
    const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / PH, py = row - n * PH;
        const int oy = min(max(py - pad, 0), OH - 1);
        int ya = oy, yb = oy;
        float ly = 0.f;
        if (POST == KP_POST_POOL) { ya = 2 * oy; yb = 2 * oy + 1; }
        if (POST == KP_POST_UP) {
            float sy = ry * oy;
            ya = (int)sy; ly = sy - ya; yb = min(ya + 1, H - 1);
        }
        for (int pxb = x0; pxb < PW; pxb += UNR * xstep) {
            RawI r[UNR][POST == KP_POST_NONE ? 1 : 4];
            float lx[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int px = pxb + j * xstep;
                if (px < PW) {
                    const int ox = min(max(px - pad, 0), OW - 1);
                    if (POST == KP_POST_NONE) {
                        r[j][0] = Vec<TI, V8>::load_raw(y.at(n, oy, ox, c0));
                    } else {
                        int xa, xb;
                        if (POST == KP_POST_POOL) { xa = 2 * ox; xb = 2 * ox + 1; lx[j] = 0.f; }
                        else { float sx = rx * ox; xa = (int)sx; lx[j] = sx - xa; xb = min(xa + 1, W - 1); }
                        r[j][0] = Vec<TI, V8>::load_raw(y.at(n, ya, xa, c0));
                        r[j][1] = Vec<TI, V8>::load_raw(y.at(n, ya, xb, c0));
                        r[j][2] = Vec<TI, V8>::load_raw(y.at(n, yb, xa, c0));
                        r[j][3] = Vec<TI, V8>::load_raw(y.at(n, yb, xb, c0));
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int px = pxb + j * xstep;
                if (px < PW) {
                    float v[V8];
                    raw_act<TI>(r[j][0], sc, sh, act, v);
                    if (POST != KP_POST_NONE) {
                        float a[V8], b[V8], c[V8];
                        raw_act<TI>(r[j][1], sc, sh, act, a);
                        raw_act<TI>(r[j][2], sc, sh, act, b);
                        raw_act<TI>(r[j][3], sc, sh, act, c);
                        if (POST == KP_POST_POOL) {
#pragma unroll
                            for (int i = 0; i < V8; ++i) v[i] = fmaxf(fmaxf(v[i], a[i]), fmaxf(b[i], c[i]));
                        } else {
                            const float l = lx[j];
#pragma unroll
                            for (int i = 0; i < V8; ++i)
                                v[i] = (1.f - ly) * ((1.f - l) * v[i] + l * a[i]) + ly * ((1.f - l) * b[i] + l * c[i]);
                        }
                    }
                    Vec<TO, V8>::store(out.at(n, py, px, c0), v);
                }
            }
        }
    }
}

// Backward pass 1: dz = gradient w.r.t. the BatchNorm output (replicate-pad fold + pool/upsample backward +
// activation backward), written to dy; per-channel sums of dz and dz*xhat accumulated.
// APPLY = false: pass 1 (dz written if dy given, sums accumulated).  APPLY = true: "recompute" pass 2 — the gather is
// repeated and dy = scale (dz - mean(dz) - xhat mean(dz xhat)) is written directly, so dz never round-trips through HBM
// (cheaper whenever d(out) is not larger than y: post = none / pool).
template <typename TG, typename TY, typename TD, int POST, bool APPLY>
__global__ void __launch_bounds__(256, KP_EW_MINB)
bn_act_bwd_rows_k(View<TG> dout, View<TY> y, View<TD> dy, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                  double* sums, double count, int act, int pad, int N, int H, int W, int C, int OH, int OW, int cg_shift) {
    using RawG = typename Vec<TG, V8>::Raw;
    using RawY = typename Vec<TY, V8>::Raw;
    __shared__ float red[2][256 * V8];
    const int ncg = C >> 3;
    const int cg = threadIdx.x & (ncg - 1), c0 = cg * V8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float sc[V8], sh[V8], mu[V8], is[V8], s1[V8], s2[V8];
#pragma unroll
    for (int i = 0; i < V8; ++i) {
        sc[i] = scale ? scale[c0 + i] : 1.f; sh[i] = shift ? shift[c0 + i] : 0.f;
        mu[i] = mean ? mean[c0 + i] : 0.f; is[i] = invstd ? invstd[c0 + i] : 1.f;
        if (APPLY) { const double rc = 1.0 / count; s1[i] = (float)(sums[c0 + i] * rc); s2[i] = (float)(sums[C + c0 + i] * rc); }
        else { s1[i] = 0.f; s2[i] = 0.f; }
    }
    const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    // consume one element: g = gradient w.r.t. the activation output, yv = raw conv output
    auto emit = [&](int n, int yy, int xx, float (&g)[V8], const float (&yv)[V8]) {
        if (APPLY) {
#pragma unroll
            for (int i = 0; i < V8; ++i) {
                const float dz = g[i] * act_grad(fmaf(yv[i], sc[i], sh[i]), act);
                g[i] = sc[i] * (dz - s1[i] - (yv[i] - mu[i]) * is[i] * s2[i]);
            }
            Vec<TD, V8>::store(dy.at(n, yy, xx, c0), g);
            return;
        }
#pragma unroll
        for (int i = 0; i < V8; ++i) {
            g[i] *= act_grad(fmaf(yv[i], sc[i], sh[i]), act);
            s1[i] += g[i];
            s2[i] = fmaf(g[i], (yv[i] - mu[i]) * is[i], s2[i]);
        }
        if (dy.p) Vec<TD, V8>::store(dy.at(n, yy, xx, c0), g);
    };
    if (POST == KP_POST_POOL) {
        const int HW2 = (H + 1) >> 1, WW2 = (W + 1) >> 1, rows = N * HW2;
        for (int row = blockIdx.x; row < rows; row += gridDim.x) {
            const int n = row / HW2, wy = (row - n * HW2) * 2;
            const RowFold fy(wy >> 1, OH, pad);
            for (int wi = x0; wi < WW2; wi += xstep) {
                const int wx = wi * 2;
                const bool full = (wy + 1 < H) && (wx + 1 < W);
                RawY ry4[4];
                float yv[4][V8], best[V8], t[V8];
                int bi[V8];
                bool ok[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int yy = wy + (q >> 1), xx = wx + (q & 1);
                    ok[q] = yy < H && xx < W;
                    if (ok[q]) ry4[q] = Vec<TY, V8>::load_raw(y.at(n, yy, xx, c0));
                }
#pragma unroll
                for (int i = 0; i < V8; ++i) { best[i] = -INFINITY; bi[i] = 0; t[i] = 0.f; }
                if (full) {
                    const RowFold fx(wx >> 1, OW, pad);
                    fold_acc<TG>(dout, n, fy, fx, c0, 1.f, t);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (ok[q]) Vec<TY, V8>::unpack(ry4[q], yv[q]);
                    else {
#pragma unroll
                        for (int i = 0; i < V8; ++i) yv[q][i] = 0.f;
                    }
                }
                if (full) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int i = 0; i < V8; ++i) {
                            float a = apply_act(fmaf(yv[q][i], sc[i], sh[i]), act);
                            if (a > best[i] || q == 0) { best[i] = a; bi[i] = q; }      // first maximum wins
                        }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (ok[q]) {
                        float g[V8];
#pragma unroll
                        for (int i = 0; i < V8; ++i) g[i] = (full && bi[i] == q) ? t[i] : 0.f;
                        emit(n, wy + (q >> 1), wx + (q & 1), g, yv[q]);
                    }
                }
            }
        }
    } else if (POST == KP_POST_NONE) {
        const int rows = N * H;
        for (int row = blockIdx.x; row < rows; row += gridDim.x) {
            const int n = row / H, yy = row - n * H;
            const RowFold fy(yy, OH, pad);
            for (int xb = x0; xb < W; xb += UNR * xstep) {
                RawG rg[UNR];
                RawY ryv[UNR];
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int xx = xb + j * xstep;
                    if (xx < W) {
                        rg[j] = Vec<TG, V8>::load_raw(dout.at(n, fy.v[0], xx + pad, c0));
                        ryv[j] = Vec<TY, V8>::load_raw(y.at(n, yy, xx, c0));
                    }
                }
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    const int xx = xb + j * xstep;
                    if (xx < W) {
                        float g[V8], yv[V8];
                        Vec<TG, V8>::unpack(rg[j], g);
                        Vec<TY, V8>::unpack(ryv[j], yv);
                        if (pad && (fy.n > 1 || xx == 0 || xx == W - 1)) {      // border pixel: add the folded copies
                            const RowFold fx(xx, OW, pad);
                            float t[V8];
                            for (int a = 0; a < fy.n; ++a)
                                for (int b = 0; b < fx.n; ++b) {
                                    if (a == 0 && b == 0) continue;
                                    Vec<TG, V8>::load(dout.at(n, fy.v[a], fx.v[b], c0), t);
#pragma unroll
                                    for (int i = 0; i < V8; ++i) g[i] += t[i];
                                }
                        }
                        emit(n, yy, xx, g, yv);
                    }
                }
            }
        }
    } else {   // bilinear x2 (align_corners=True) backward: gather the <= 4x4 outputs that read (yy,xx)
        const int rows = N * H;
        for (int row = blockIdx.x; row < rows; row += gridDim.x) {
            const int n = row / H, yy = row - n * H;
            int ylo = 0, yhi = OH - 1;
            if (H > 1) { ylo = max(0, 2 * yy - 2); yhi = min(OH - 1, 2 * yy + 3); }
            for (int xx = x0; xx < W; xx += xstep) {
                int xlo = 0, xhi = OW - 1;
                if (W > 1) { xlo = max(0, 2 * xx - 2); xhi = min(OW - 1, 2 * xx + 3); }
                const RawY ryv = Vec<TY, V8>::load_raw(y.at(n, yy, xx, c0));
                // horizontal taps first (at most 6 candidates, <= 4 non-zero), kept in registers
                float wxs[6];
                int nx = 0, xsel[6];
                for (int X = xlo; X <= xhi; ++X) {
                    float sx = rx * X;
                    int xa = (int)sx;
                    float lx = sx - xa;
                    int xb = min(xa + 1, W - 1);
                    float wv = (xa == xx ? 1.f - lx : 0.f) + (xb == xx ? lx : 0.f);
                    if (wv != 0.f && nx < 6) { wxs[nx] = wv; xsel[nx] = X; ++nx; }
                }
                float g[V8], yv[V8];
#pragma unroll
                for (int i = 0; i < V8; ++i) g[i] = 0.f;
                for (int Y = ylo; Y <= yhi; ++Y) {
                    float sy = ry * Y;
                    int ya = (int)sy;
                    float ly = sy - ya;
                    int yb = min(ya + 1, H - 1);
                    float wyv = (ya == yy ? 1.f - ly : 0.f) + (yb == yy ? ly : 0.f);
                    if (wyv == 0.f) continue;
                    const RowFold fy(Y, OH, pad);
                    for (int k = 0; k < nx; ++k) {
                        const RowFold fx(xsel[k], OW, pad);
                        fold_acc<TG>(dout, n, fy, fx, c0, wyv * wxs[k], g);
                    }
                }
                Vec<TY, V8>::unpack(ryv, yv);
                emit(n, yy, xx, g, yv);
            }
        }
    }
    if (APPLY) return;
#pragma unroll
    for (int i = 0; i < V8; ++i) { red[0][threadIdx.x * V8 + i] = s1[i]; red[1][threadIdx.x * V8 + i] = s2[i]; }
    __syncthreads();
    // one channel per thread: sum over the 256/ncg threads that share its channel group
    for (int ch = threadIdx.x; ch < C; ch += 256) {
        const int g8 = ch >> 3, i = ch & 7;
        float a = 0.f, b = 0.f;
        for (int t = g8; t < 256; t += ncg) { a += red[0][t * V8 + i]; b += red[1][t * V8 + i]; }
        atomicAdd(&sums[ch], (double)a);
        atomicAdd(&sums[C + ch], (double)b);
    }
}

// Backward pass 2 (BatchNorm only), in place: dy = scale * (dz - mean(dz) - xhat * mean(dz * xhat))
template <typename TY, typename TD>
__global__ void __launch_bounds__(256, KP_EW_MINB)
bn_bwd_apply_rows_k(View<TY> y, View<TD> dy, const float* __restrict__ scale, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const double* __restrict__ sums, double count, int N, int H, int W,
                    int C, int cg_shift, float* dgamma, float* dbeta) {
    using RawY = typename Vec<TY, V8>::Raw;
    using RawD = typename Vec<TD, V8>::Raw;
    const int ncg = C >> 3;
    const int cg = threadIdx.x & (ncg - 1), c0 = cg * V8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float a0[V8], a1[V8], a2[V8];      // dy = a0*dz + a1*y + a2
#pragma unroll
    for (int i = 0; i < V8; ++i) {
        const float sc = scale[c0 + i], mu = mean[c0 + i], is = invstd[c0 + i];
        const double rc = 1.0 / count;
        const float m1 = (float)(sums[c0 + i] * rc), m2 = (float)(sums[C + c0 + i] * rc);
        a0[i] = sc;
        a1[i] = -sc * is * m2;
        a2[i] = -sc * m1 + sc * is * m2 * mu;
        if (blockIdx.x == 0 && x0 == 0) {           // fused kp_bn_grad_finalize
            if (dbeta) dbeta[c0 + i] = (float)sums[c0 + i];
            if (dgamma) dgamma[c0 + i] = (float)sums[C + c0 + i];
        }
    }
    const int rows = N * H;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / H, yy = row - n * H;
        for (int xb = x0; xb < W; xb += UNR * xstep) {
            RawY ryv[UNR];
            RawD rd[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int xx = xb + j * xstep;
                if (xx < W) {
                    ryv[j] = Vec<TY, V8>::load_raw(y.at(n, yy, xx, c0));
                    rd[j] = Vec<TD, V8>::load_raw(dy.at(n, yy, xx, c0));
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int xx = xb + j * xstep;
                if (xx < W) {
                    float g[V8], yv[V8];
                    Vec<TD, V8>::unpack(rd[j], g);
                    Vec<TY, V8>::unpack(ryv[j], yv);
#pragma unroll
                    for (int i = 0; i < V8; ++i) g[i] = fmaf(a0[i], g[i], fmaf(a1[i], yv[i], a2[i]));
                    Vec<TD, V8>::store(dy.at(n, yy, xx, c0), g);
                }
            }
        }
    }
}

// generic (any strides / channel count) version of pass 2
template <typename TY, typename TD>
__global__ void bn_bwd_apply_generic_k(View<TY> y, View<TD> dy, const float* __restrict__ scale,
                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                       const double* __restrict__ sums, double count, int N, int H, int W, int C) {
    const long long total = (long long)N * H * W * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long p = i / C;
        int xx = (int)(p % W);
        long long r = p / W;
        int yy = (int)(r % H);
        int n = (int)(r / H);
        float dz = to_f(*dy.at(n, yy, xx, c)), yv = to_f(*y.at(n, yy, xx, c));
        float m1 = (float)(sums[c] / count), m2 = (float)(sums[C + c] / count);
        from_f(dy.at(n, yy, xx, c), scale[c] * (dz - m1 - (yv - mean[c]) * invstd[c] * m2));
    }
}

// ------------------------------------------------------------------------------------------------
// Bilinear x2 (align_corners=True) specialisations.  Forward: one thread produces a 2x2 output quad from the
// 3x3 input neighbourhood it needs, so BatchNorm + activation run 9 (not 16) times per 4 outputs.  Backward: one
// thread gathers the 4x4 outputs that read its input pixel with all 16 loads issued up front; only pixels whose
// gather touches the replicate-padded border take the slow folding path.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float actf(float z, float slope) { return fmaxf(z, slope * z); }

template <typename TI, typename TO>
__global__ void __launch_bounds__(256, KP_EW_MINB)
bn_act_fwd_up_quad_k(View<TI> y, View<TO> out, const float* __restrict__ scale, const float* __restrict__ shift,
                     float slope, int pad, int N, int H, int W, int C, int cg_shift, const BnFuse fuse) {
    // Outputs (2k, 2k+1) x (2j, 2j+1).  With src = r * dst, r = (H-1)/(2H-1): floor(r*2k) = k-1 (k >= 1) and
    // floor(r*(2k+1)) = k, so output row 2k blends input rows (k-1, k) and row 2k+1 blends (k, k+1): a fixed
    // 3x3 window, no index selects.  The fractions come from the same float formula PyTorch uses.
    using RawI = typename Vec<TI, V8>::Raw;
    const int OH = 2 * H, OW = 2 * W;
    const int ncg = C >> 3;
    const int cg = threadIdx.x & (ncg - 1), c0 = cg * V8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float sc[V8], sh[V8];
    if (fuse.stats) fused_affine(fuse, C, c0, blockIdx.x == 0 && x0 == 0, sc, sh);
    else {
#pragma unroll
        for (int i = 0; i < V8; ++i) { sc[i] = scale ? scale[c0 + i] : 1.f; sh[i] = shift ? shift[c0 + i] : 0.f; }
    }
    const float ry = (float)(H - 1) / (float)(OH - 1);
    const float rx = (float)(W - 1) / (float)(OW - 1);
    const int rows = N * H;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / H, k = row - n * H;
        const int yr[3] = {max(k - 1, 0), k, min(k + 1, H - 1)};
        const float ly0 = k > 0 ? fminf(fmaxf(ry * (2 * k) - (float)(k - 1), 0.f), 1.f) : 0.f;
        const float ly1 = fminf(fmaxf(ry * (2 * k + 1) - (float)k, 0.f), 1.f);
        for (int j = x0; j < W; j += xstep) {
            const int xr[3] = {max(j - 1, 0), j, min(j + 1, W - 1)};
            const float lx0 = j > 0 ? fminf(fmaxf(rx * (2 * j) - (float)(j - 1), 0.f), 1.f) : 0.f;
            const float lx1 = fminf(fmaxf(rx * (2 * j + 1) - (float)j, 0.f), 1.f);
            RawI r[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) r[a][b] = Vec<TI, V8>::load_raw(y.at(n, yr[a], xr[b], c0));
            float h[3][2][V8];          // horizontally interpolated rows: [input row][output column parity]
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float v0[V8], v1[V8], v2[V8];
                Vec<TI, V8>::unpack(r[a][0], v0);
                Vec<TI, V8>::unpack(r[a][1], v1);
                Vec<TI, V8>::unpack(r[a][2], v2);
#pragma unroll
                for (int i = 0; i < V8; ++i) {
                    const float p0 = actf(fmaf(v0[i], sc[i], sh[i]), slope);
                    const float p1 = actf(fmaf(v1[i], sc[i], sh[i]), slope);
                    const float p2 = actf(fmaf(v2[i], sc[i], sh[i]), slope);
                    h[a][0][i] = (1.f - lx0) * p0 + lx0 * p1;
                    h[a][1][i] = (1.f - lx1) * p1 + lx1 * p2;
                }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    float o[V8];
                    const float l = a == 0 ? ly0 : ly1;
#pragma unroll
                    for (int i = 0; i < V8; ++i) o[i] = (1.f - l) * h[a][b][i] + l * h[a + 1][b][i];
                    const int oy = 2 * k + a, ox = 2 * j + b;
                    Vec<TO, V8>::store(out.at(n, oy + pad, ox + pad, c0), o);
                    if (pad) {   // replicate border copies
                        const bool top = oy == 0, bot = oy == OH - 1, lef = ox == 0, rig = ox == OW - 1;
                        if (top) Vec<TO, V8>::store(out.at(n, 0, ox + 1, c0), o);
                        if (bot) Vec<TO, V8>::store(out.at(n, OH + 1, ox + 1, c0), o);
                        if (lef) Vec<TO, V8>::store(out.at(n, oy + 1, 0, c0), o);
                        if (rig) Vec<TO, V8>::store(out.at(n, oy + 1, OW + 1, c0), o);
                        if (top && lef) Vec<TO, V8>::store(out.at(n, 0, 0, c0), o);
                        if (top && rig) Vec<TO, V8>::store(out.at(n, 0, OW + 1, c0), o);
                        if (bot && lef) Vec<TO, V8>::store(out.at(n, OH + 1, 0, c0), o);
                        if (bot && rig) Vec<TO, V8>::store(out.at(n, OH + 1, OW + 1, c0), o);
                    }
                }
        }
    }
}

template <typename TG, typename TY, typename TD>
__global__ void __launch_bounds__(256, KP_EW_MINB)
bn_act_bwd_up_k(View<TG> dout, View<TY> y, View<TD> dy, const float* __restrict__ scale,
                const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                double* sums, float slope, int pad, int N, int H, int W, int C, int cg_shift) {
    using RawG = typename Vec<TG, V8>::Raw;
    __shared__ float red[2][256 * V8];
    const int OH = 2 * H, OW = 2 * W;
    const int ncg = C >> 3;
    const int cg = threadIdx.x & (ncg - 1), c0 = cg * V8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float sc[V8], sh[V8], mu[V8], is[V8], s1[V8], s2[V8];
#pragma unroll
    for (int i = 0; i < V8; ++i) {
        sc[i] = scale ? scale[c0 + i] : 1.f; sh[i] = shift ? shift[c0 + i] : 0.f;
        mu[i] = mean ? mean[c0 + i] : 0.f; is[i] = invstd ? invstd[c0 + i] : 1.f;
        s1[i] = 0.f; s2[i] = 0.f;
    }
    const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    const int rows = N * H;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / H, yy = row - n * H;
        // vertical taps: outputs 2yy-1 .. 2yy+2
        float wy[4];
        int Ys[4];
        bool ybord = false;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            int Y = 2 * yy - 1 + a;
            float wv = 0.f;
            if (Y >= 0 && Y < OH) {
                float sy = ry * Y;
                int t = (int)sy;
                float l = sy - t;
                int tb = min(t + 1, H - 1);
                wv = (t == yy ? 1.f - l : 0.f) + (tb == yy ? l : 0.f);
                if (wv != 0.f && (Y == 0 || Y == OH - 1)) ybord = true;
            }
            wy[a] = wv;
            Ys[a] = min(max(Y, 0), OH - 1);
        }
        for (int xx = x0; xx < W; xx += xstep) {
            float wx[4];
            int Xs[4];
            bool bord = ybord;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int X = 2 * xx - 1 + b;
                float wv = 0.f;
                if (X >= 0 && X < OW) {
                    float sx = rx * X;
                    int t = (int)sx;
                    float l = sx - t;
                    int tb = min(t + 1, W - 1);
                    wv = (t == xx ? 1.f - l : 0.f) + (tb == xx ? l : 0.f);
                    if (wv != 0.f && (X == 0 || X == OW - 1)) bord = true;
                }
                wx[b] = wv;
                Xs[b] = min(max(X, 0), OW - 1);
            }
            float g[V8], yv[V8];
#pragma unroll
            for (int i = 0; i < V8; ++i) g[i] = 0.f;
            const typename Vec<TY, V8>::Raw ryv = Vec<TY, V8>::load_raw(y.at(n, yy, xx, c0));
            if (!(pad && bord)) {
                RawG r[4][4];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) r[a][b] = Vec<TG, V8>::load_raw(dout.at(n, Ys[a] + pad, Xs[b] + pad, c0));
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        float t[V8];
                        Vec<TG, V8>::unpack(r[a][b], t);
                        const float wgt = wy[a] * wx[b];
#pragma unroll
                        for (int i = 0; i < V8; ++i) g[i] = fmaf(wgt, t[i], g[i]);
                    }
            } else {
                for (int a = 0; a < 4; ++a) {
                    if (wy[a] == 0.f) continue;
                    const RowFold fy(Ys[a], OH, pad);
                    for (int b = 0; b < 4; ++b) {
                        if (wx[b] == 0.f) continue;
                        const RowFold fx(Xs[b], OW, pad);
                        fold_acc<TG>(dout, n, fy, fx, c0, wy[a] * wx[b], g);
                    }
                }
            }
            Vec<TY, V8>::unpack(ryv, yv);
#pragma unroll
            for (int i = 0; i < V8; ++i) {
                g[i] *= (fmaf(yv[i], sc[i], sh[i]) > 0.f) ? 1.f : slope;
                s1[i] += g[i];
                s2[i] = fmaf(g[i], (yv[i] - mu[i]) * is[i], s2[i]);
            }
            if (dy.p) Vec<TD, V8>::store(dy.at(n, yy, xx, c0), g);
        }
    }
#pragma unroll
    for (int i = 0; i < V8; ++i) { red[0][threadIdx.x * V8 + i] = s1[i]; red[1][threadIdx.x * V8 + i] = s2[i]; }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C; ch += 256) {
        const int g8 = ch >> 3, i = ch & 7;
        float a = 0.f, b = 0.f;
        for (int t = g8; t < 256; t += ncg) { a += red[0][t * V8 + i]; b += red[1][t * V8 + i]; }
        atomicAdd(&sums[ch], (double)a);
        atomicAdd(&sums[C + ch], (double)b);
    }
}

static float act_slope(int act) { return act == KP_ACT_LEAKY ? 0.01f : (act == KP_ACT_RELU ? 0.f : 1.f); }


static bool fast_ok(int C, long long rows_px) {
    if (C % 8) return false;
    int ncg = C / 8;
    return ncg >= 1 && ncg <= 256 && (ncg & (ncg - 1)) == 0 && rows_px < (1LL << 31);
}
static bool lean_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = kp_env("KP_BN_LEAN"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
static bool pipe_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = kp_env("KP_BN_PIPE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
// dense bf16 NHWC rows the bulk-async kernels can stream: pixel stride == C, 512-item chunks
static bool pipe_up_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = kp_env("KP_BN_PIPE_UP"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
static bool pipe_view_ok(const kp_view* v, int C) { return v->dtype == KP_BF16 && v->sx == C && view_vec8_ok(v, C); }
static bool pipe_shape_ok(int C, int W) {
    if (C % 8) return false;
    const int ncg = C / 8;
    return (ncg & (ncg - 1)) == 0 && ncg <= 64 && ((long long)W * ncg) % pipe::IPC == 0;   // C <= 512 (PIPE_EXT)
}
static bool pipe_pool_shape_ok(int C, int OW) {
    if (C % 8) return false;
    const int ncg = C / 8;
    return (ncg & (ncg - 1)) == 0 && ncg <= 64 && ((long long)OW * ncg) % PIPE_POOL_ITEMS == 0;
}
template <typename K>
static int pipe_attr(K kernel, int smem) {
    KP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return KP_OK;
}
static int pipe_grid(long long units) {
    const long long cap = 2LL * kp_sm_count();
    return (int)(units < cap ? units : cap);
}
static int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }
// image order of the bulk-async BatchNorm passes (kp_bn_pipe.cuh Cursor): bit 0 forward passes, bit 1 backward pass 1,
// bit 2 backward pass 2 run last image first.  Default 5 (measured, KeyNet F batch 64: mask 0 4307-4333 pairs/s, 1 4363,
// 2 4345, 3 4384, 7 4340, 5 4415): the forward pass starts on the tail of y the conv just wrote, pass 1 sweeps forward and
// pass 2 comes back over the same (dout, y) tail; the convs that follow sweep forward again.  KP_BN_REV overrides the mask.
static int bn_rev(int role, int N) {
    static int mask = -1;
    if (mask < 0) { const char* e = kp_env("KP_BN_REV"); mask = e ? atoi(e) : 5; }
    return (mask & role) ? N : 0;
}

static int rows_grid(long long rows) {
    long long cap = (long long)kp_sm_count() * 8;
    return (int)(rows < cap ? (rows < 1 ? 1 : rows) : cap);
}

static void out_dims(int post, int H, int W, int* OH, int* OW) {
    if (post == KP_POST_POOL) { *OH = H / 2; *OW = W / 2; }
    else if (post == KP_POST_UP) { *OH = 2 * H; *OW = 2 * W; }
    else { *OH = H; *OW = W; }
}

}  // namespace

extern "C" int kp_bn_finalize(kp_stream stream, const double* stats, int C, double count, const float* gamma,
                              const float* beta, float eps, float momentum, float* running_mean,
                              float* running_var, int64_t* nbt, float* scale, float* shift, float* save_mean,
                              float* save_invstd) {
    KP_CHECK_ARG(stats && scale && shift && C > 0 && count > 0, "kp_bn_finalize: bad arguments");
    bn_finalize_k<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, C, count, gamma, beta, eps, momentum,
                                                                    running_mean, running_var, (long long*)nbt,
                                                                    scale, shift, save_mean, save_invstd);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

extern "C" int kp_bn_grad_finalize(kp_stream stream, const double* sums, int C, float* dgamma, float* dbeta) {
    KP_CHECK_ARG(sums && C > 0, "kp_bn_grad_finalize: bad arguments");
    bn_grad_finalize_k<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, C, dgamma, dbeta);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

// fuse == nullptr: plain normalise/activate.  Otherwise the fast kernels do the BatchNorm finalize themselves and
// *fused_done is set; if the fast path does not apply nothing is launched and *fused_done stays 0.
static int bn_act_fwd_impl(kp_stream stream, const kp_view* y, const kp_view* out, const float* scale,
                           const float* shift, int act, int post, int pad, int N, int H, int W, int C,
                           const BnFuse* fuse, int* fused_done) {
    const bool coresident = (post & KP_POST_CORESIDENT) != 0;
    post &= ~KP_POST_CORESIDENT;
    KP_CHECK_ARG(y && out && y->ptr && out->ptr && N > 0 && H > 0 && W > 0 && C > 0, "kp_bn_act_fwd: bad arguments");
    KP_CHECK_ARG(post != KP_POST_POOL || (H >= 2 && W >= 2), "kp_bn_act_fwd: pool needs H,W >= 2");
    int OH, OW;
    out_dims(post, H, W, &OH, &OW);
    const bool vec = view_vec8_ok(y, C) && view_vec8_ok(out, C);
    const long long P = (long long)N * (OH + 2 * pad) * (OW + 2 * pad);
    return dispatch1(y->dtype, [&](auto ti) -> int {
        return dispatch1(out->dtype, [&](auto to) -> int {
            using TI = decltype(ti);
            using TO = decltype(to);
            BnFuse nofuse;
            nofuse.stats = nullptr;
            const BnFuse fz = fuse ? *fuse : nofuse;
            if (pipe_enabled() && post == KP_POST_NONE && pipe_shape_ok(C, W) && pipe_view_ok(y, C) && pipe_view_ok(out, C)) {
                const int cpr = (int)((long long)W * (C / 8) / pipe::IPC), sh = ilog2(C / 8);
                const long long units = (long long)N * (H + 2 * pad) * cpr;
                constexpr int smem = pipe_smem_bytes<PIPE_FWD_STAGE, PIPE_FWD_STAGES>();
                cudaStream_t st = (cudaStream_t)stream;
#define KP_FWDP(ACTV)                                                                                                  \
    do {                                                                                                               \
        int rc_ = pipe_attr(bn_fwd_none_pipe_k<ACTV>, smem);                                                           \
        if (rc_) return rc_;                                                                                           \
        bn_fwd_none_pipe_k<ACTV><<<pipe_grid(units), pipe::THREADS, smem, st>>>(                                       \
            make_rows<const bf16>(y), make_rows<bf16>(out), scale, shift, pad, N, H, W, C, sh, cpr, fz, bn_rev(1, N));              \
    } while (0)
                // small-footprint variant (2-stage ring = 16 KB, one CTA per SM): fits next to a persistent tensor-core conv
                // CTA of the other stream (208 KB of the SM's 228), so the pass runs UNDER that conv instead of waiting for
                // its CTAs to exit.  Slower alone (a sixth of the bytes in flight), faster in the step: 14.49 -> 14.28-14.37
                // ms (gpurun_out/r2_b_slim*.log).  Chosen by the caller's KP_POST_CORESIDENT hint.
                if (coresident) {
                    constexpr int smem2 = pipe_smem_bytes<PIPE_FWD_STAGE, 2>();
                    const long long cap1 = kp_sm_count();
                    const int g1 = (int)(units < cap1 ? units : cap1);
#define KP_FWDS(ACTV)                                                                                                  \
    do {                                                                                                               \
        int rc_ = pipe_attr(bn_fwd_none_pipe_k<ACTV, 2>, smem2);                                                       \
        if (rc_) return rc_;                                                                                           \
        bn_fwd_none_pipe_k<ACTV, 2><<<g1, pipe::THREADS, smem2, st>>>(                                                 \
            make_rows<const bf16>(y), make_rows<bf16>(out), scale, shift, pad, N, H, W, C, sh, cpr, fz, bn_rev(1, N)); \
    } while (0)
                    KP_ACT_SWITCH(act, KP_FWDS);
#undef KP_FWDS
                    if (fuse && fused_done) *fused_done = 1;
                    KP_LAUNCH_CHECK();
                    return KP_OK;
                }
                KP_ACT_SWITCH(act, KP_FWDP);
#undef KP_FWDP
                if (fuse && fused_done) *fused_done = 1;
            } else if (pipe_enabled() && post == KP_POST_POOL && H % 2 == 0 && W % 2 == 0 && pipe_pool_shape_ok(C, OW) &&
                       pipe_view_ok(y, C) && pipe_view_ok(out, C)) {
                const int cpr = (int)((long long)OW * (C / 8) / PIPE_POOL_ITEMS), sh = ilog2(C / 8);
                const long long units = (long long)N * (OH + 2 * pad) * cpr;
                constexpr int smem = pipe_smem_bytes<PIPE_FPOOL_STAGE, PIPE_FPOOL_STAGES>();
                cudaStream_t st = (cudaStream_t)stream;
#define KP_FWDPP(ACTV)                                                                                                 \
    do {                                                                                                               \
        int rc_ = pipe_attr(bn_fwd_pool_pipe_k<ACTV>, smem);                                                           \
        if (rc_) return rc_;                                                                                           \
        bn_fwd_pool_pipe_k<ACTV><<<pipe_grid(units), pipe::THREADS, smem, st>>>(                                       \
            make_rows<const bf16>(y), make_rows<bf16>(out), scale, shift, pad, N, OH, OW, C, sh, cpr, fz, bn_rev(1, N));            \
    } while (0)
                KP_ACT_SWITCH(act, KP_FWDPP);
#undef KP_FWDPP
                if (fuse && fused_done) *fused_done = 1;
            } else if (pipe_enabled() && post == KP_POST_UP && H >= 2 && W >= 2 && pipe_pool_shape_ok(C, W) &&
                       pipe_view_ok(y, C) && pipe_view_ok(out, C) && pipe_up_enabled()) {
                const int cpr = (int)((long long)W * (C / 8) / PIPE_UP_ITEMS), sh = ilog2(C / 8);
                const long long strips = (long long)N * ((H + PIPE_UP_BAND - 1) / PIPE_UP_BAND) * cpr;
                constexpr int smem = pipe_smem_bytes<PIPE_FUP_STAGE, PIPE_FUP_STAGES>();
                cudaStream_t st = (cudaStream_t)stream;
#define KP_FWDUP(ACTV)                                                                                                 \
    do {                                                                                                               \
        int rc_ = pipe_attr(bn_fwd_up_pipe_k<ACTV>, smem);                                                             \
        if (rc_) return rc_;                                                                                           \
        bn_fwd_up_pipe_k<ACTV><<<pipe_grid(strips), pipe::THREADS, smem, st>>>(                                        \
            make_rows<const bf16>(y), make_rows<bf16>(out), scale, shift, pad, N, H, W, C, sh, cpr, fz);               \
    } while (0)
                KP_ACT_SWITCH(act, KP_FWDUP);
#undef KP_FWDUP
                if (fuse && fused_done) *fused_done = 1;
            } else if (vec && fast_ok(C, P * C) && std::is_same<TI, TO>::value && lean_enabled() &&
                (post == KP_POST_NONE || post == KP_POST_POOL)) {
                const int g = rows_grid((long long)N * (OH + 2 * pad)), sh = ilog2(C / 8);
                cudaStream_t st = (cudaStream_t)stream;
#define KP_FWDL(ACTV)                                                                                                   \
    do {                                                                                                                \
        if (post == KP_POST_NONE)                                                                                       \
            bn_fwd_lean_k<TI, KP_POST_NONE, ACTV><<<g, 256, 0, st>>>(make_rows<const TI>(y), make_rows<TI>(out), scale, \
                                                                    shift, pad, N, H, W, C, OH, OW, sh, fz);            \
        else                                                                                                            \
            bn_fwd_lean_k<TI, KP_POST_POOL, ACTV><<<g, 256, 0, st>>>(make_rows<const TI>(y), make_rows<TI>(out), scale, \
                                                                    shift, pad, N, H, W, C, OH, OW, sh, fz);            \
    } while (0)
                KP_ACT_SWITCH(act, KP_FWDL);
#undef KP_FWDL
                if (fuse && fused_done) *fused_done = 1;
            } else if (vec && fast_ok(C, P * C)) {
                const int g = rows_grid((long long)N * (OH + 2 * pad)), sh = ilog2(C / 8);
                cudaStream_t st = (cudaStream_t)stream;
#define KP_FWD(POSTV) bn_act_fwd_rows_k<TI, TO, POSTV><<<g, 256, 0, st>>>(make_view<TI>(y), make_view<TO>(out), scale, shift, act, pad, N, H, W, C, OH, OW, sh, fz)
                if (post == KP_POST_NONE) KP_FWD(KP_POST_NONE);
                else if (post == KP_POST_POOL) KP_FWD(KP_POST_POOL);
                else if (H < 2 || W < 2) KP_FWD(KP_POST_UP);
                else bn_act_fwd_up_quad_k<TI, TO><<<rows_grid((long long)N * H), 256, 0, st>>>(make_view<TI>(y), make_view<TO>(out), scale, shift, act_slope(act), pad, N, H, W, C, sh, fz);
#undef KP_FWD
                if (fuse && fused_done) *fused_done = 1;
            } else if (fuse) {
                return (int)KP_OK;                              // caller falls back to finalize + plain forward
            } else if (vec) {
                Launch2D l = plan2d(P, C, 8);
                bn_act_fwd_k<TI, TO, 8><<<l.grid, l.block, 0, (cudaStream_t)stream>>>(
                    make_view<TI>(y), make_view<TO>(out), scale, shift, act, post, pad, N, H, W, C, OH, OW);
            } else {
                Launch2D l = plan2d(P, C, 1);
                bn_act_fwd_k<TI, TO, 1><<<l.grid, l.block, 0, (cudaStream_t)stream>>>(
                    make_view<TI>(y), make_view<TO>(out), scale, shift, act, post, pad, N, H, W, C, OH, OW);
            }
            KP_LAUNCH_CHECK();
            return (int)KP_OK;
        });
    });
}

extern "C" int kp_bn_act_fwd(kp_stream stream, const kp_view* y, const kp_view* out, const float* scale,
                             const float* shift, int act, int post, int pad, int N, int H, int W, int C) {
    return bn_act_fwd_impl(stream, y, out, scale, shift, act, post, pad, N, H, W, C, nullptr, nullptr);
}

extern "C" int kp_bn_finalize_act_fwd(kp_stream stream, const double* stats, double count, const float* gamma,
                                      const float* beta, float eps, float momentum, float* running_mean,
                                      float* running_var, int64_t* nbt, float* scale, float* shift, float* save_mean,
                                      float* save_invstd, const kp_view* y, const kp_view* out, int act, int post,
                                      int pad, int N, int H, int W, int C) {
    KP_CHECK_ARG(stats && scale && shift && count > 0, "kp_bn_finalize_act_fwd: bad arguments");
    BnFuse f;
    f.stats = stats; f.count = count; f.gamma = gamma; f.beta = beta; f.eps = eps; f.momentum = momentum;
    f.rmean = running_mean; f.rvar = running_var; f.nbt = (long long*)nbt; f.scale = scale; f.shift = shift;
    f.mean = save_mean; f.invstd = save_invstd;
    int done = 0;
    int rc = bn_act_fwd_impl(stream, y, out, scale, shift, act, post, pad, N, H, W, C, &f, &done);
    if (rc || done) return rc;
    rc = kp_bn_finalize(stream, stats, C, count, gamma, beta, eps, momentum, running_mean, running_var, nbt, scale, shift,
                        save_mean, save_invstd);
    if (rc) return rc;
    return bn_act_fwd_impl(stream, y, out, scale, shift, act, post, pad, N, H, W, C, nullptr, nullptr);
}

extern "C" int kp_bn_act_bwd_reduce(kp_stream stream, const kp_view* dout, const kp_view* y, const kp_view* dy,
                                    const float* scale, const float* shift, const float* mean, const float* invstd,
                                    double* sums, int act, int post, int pad, int N, int H, int W, int C) {
    KP_CHECK_ARG(dout && y && dout->ptr && y->ptr && sums && N > 0 && H > 0 && W > 0 && C > 0,
                 "kp_bn_act_bwd_reduce: bad arguments");
    int OH, OW;
    out_dims(post, H, W, &OH, &OW);
    kp_view none = {nullptr, 0, 0, 0, 1, y->dtype, 0};
    const kp_view* dyv = (dy && dy->ptr) ? dy : &none;
    const bool vec = view_vec8_ok(y, C) && view_vec8_ok(dout, C) && (dyv->ptr == nullptr || view_vec8_ok(dyv, C));
    const long long P = (long long)N * H * W;
    return dispatch1(dout->dtype, [&](auto tg) -> int {
        return dispatch1(y->dtype, [&](auto ty) -> int {
            return dispatch1(dyv->dtype, [&](auto td) -> int {
                using TG = decltype(tg);
                using TY = decltype(ty);
                using TD = decltype(td);
                if (pipe_enabled() && post == KP_POST_NONE && pipe_shape_ok(C, W) && pipe_view_ok(dout, C) &&
                    pipe_view_ok(y, C) && (dyv->ptr == nullptr || pipe_view_ok(dyv, C))) {
                    const int cpr = (int)((long long)W * (C / 8) / pipe::IPC), sh = ilog2(C / 8);
                    const long long units = (long long)N * H * cpr;
                    constexpr int smem = pipe_smem_bytes<PIPE_BWD_STAGE, PIPE_BWD_STAGES>();
                    cudaStream_t st = (cudaStream_t)stream;
#define KP_BWDP(ACTV)                                                                                                  \
    do {                                                                                                               \
        if (dyv->ptr) {                                                                                                \
            int rc_ = pipe_attr(bn_bwd_none_pipe_k<ACTV, PASS1_WRITE>, smem);                                          \
            if (rc_) return rc_;                                                                                       \
            bn_bwd_none_pipe_k<ACTV, PASS1_WRITE><<<pipe_grid(units), pipe::THREADS, smem, st>>>(                      \
                make_rows<const bf16>(dout), make_rows<const bf16>(y), make_rows<bf16>(dyv), scale, shift, mean,       \
                invstd, sums, 1.0, pad, N, H, W, C, sh, cpr, nullptr, nullptr, bn_rev(2, N));                                       \
        } else {                                                                                                       \
            int rc_ = pipe_attr(bn_bwd_none_pipe_k<ACTV, PASS1_SUMS>, smem);                                           \
            if (rc_) return rc_;                                                                                       \
            bn_bwd_none_pipe_k<ACTV, PASS1_SUMS><<<pipe_grid(units), pipe::THREADS, smem, st>>>(                       \
                make_rows<const bf16>(dout), make_rows<const bf16>(y), make_rows<bf16>(dyv), scale, shift, mean,       \
                invstd, sums, 1.0, pad, N, H, W, C, sh, cpr, nullptr, nullptr, bn_rev(2, N));                                       \
        }                                                                                                              \
    } while (0)
                    KP_ACT_SWITCH(act, KP_BWDP);
#undef KP_BWDP
                } else if (pipe_enabled() && post == KP_POST_POOL && H % 2 == 0 && W % 2 == 0 && pipe_pool_shape_ok(C, OW) &&
                           pipe_view_ok(dout, C) && pipe_view_ok(y, C) && (dyv->ptr == nullptr || pipe_view_ok(dyv, C))) {
                    const int cpr = (int)((long long)OW * (C / 8) / PIPE_POOL_ITEMS), sh = ilog2(C / 8);
                    const long long units = (long long)N * OH * cpr;
                    constexpr int smem = pipe_smem_bytes<PIPE_BPOOL_STAGE, PIPE_BPOOL_STAGES>();
                    cudaStream_t st = (cudaStream_t)stream;
#define KP_BWDPP(ACTV)                                                                                                 \
    do {                                                                                                               \
        if (dyv->ptr) {                                                                                                \
            int rc_ = pipe_attr(bn_bwd_pool_pipe_k<ACTV, PASS1_WRITE>, smem);                                          \
            if (rc_) return rc_;                                                                                       \
            bn_bwd_pool_pipe_k<ACTV, PASS1_WRITE><<<pipe_grid(units), pipe::THREADS, smem, st>>>(                      \
                make_rows<const bf16>(dout), make_rows<const bf16>(y), make_rows<bf16>(dyv), scale, shift, mean,       \
                invstd, sums, 1.0, pad, N, OH, OW, C, sh, cpr, nullptr, nullptr, bn_rev(2, N));                                     \
        } else {                                                                                                       \
            int rc_ = pipe_attr(bn_bwd_pool_pipe_k<ACTV, PASS1_SUMS>, smem);                                           \
            if (rc_) return rc_;                                                                                       \
            bn_bwd_pool_pipe_k<ACTV, PASS1_SUMS><<<pipe_grid(units), pipe::THREADS, smem, st>>>(                       \
                make_rows<const bf16>(dout), make_rows<const bf16>(y), make_rows<bf16>(dyv), scale, shift, mean,       \
                invstd, sums, 1.0, pad, N, OH, OW, C, sh, cpr, nullptr, nullptr, bn_rev(2, N));                                     \
        }                                                                                                              \
    } while (0)
                    KP_ACT_SWITCH(act, KP_BWDPP);
#undef KP_BWDPP
                } else if (pipe_enabled() && pipe_up_enabled() && post == KP_POST_UP && pad == 1 && H >= 2 && W >= 2 &&
                           pipe_pool_shape_ok(C, W) && pipe_view_ok(dout, C) && pipe_view_ok(y, C) && dyv->ptr != nullptr &&
                           pipe_view_ok(dyv, C)) {
                    const int cpr = (int)((long long)W * (C / 8) / PIPE_UP_ITEMS), sh = ilog2(C / 8);
                    const long long strips = (long long)N * ((H + PIPE_UP_BAND - 1) / PIPE_UP_BAND) * cpr;
                    constexpr int smem = pipe_smem_bytes<PIPE_BUP_STAGE, PIPE_BUP_STAGES>();
                    cudaStream_t st = (cudaStream_t)stream;
#define KP_BWDUP(ACTV)                                                                                                 \
    do {                                                                                                               \
        int rc_ = pipe_attr(bn_bwd_up_pipe_k<ACTV>, smem);                                                             \
        if (rc_) return rc_;                                                                                           \
        bn_bwd_up_pipe_k<ACTV><<<pipe_grid(strips), pipe::THREADS, smem, st>>>(                                        \
            make_rows<const bf16>(dout), make_rows<const bf16>(y), make_rows<bf16>(dyv), scale, shift, mean, invstd,   \
            sums, N, H, W, C, sh, cpr);                                                                                \
    } while (0)
                    KP_ACT_SWITCH(act, KP_BWDUP);
#undef KP_BWDUP
                } else if (vec && fast_ok(C, P * C) && std::is_same<TG, TY>::value && std::is_same<TY, TD>::value &&
                    lean_enabled() && (post == KP_POST_NONE || post == KP_POST_POOL)) {
                    const long long rows = post == KP_POST_POOL ? (long long)N * ((H + 1) / 2) : (long long)N * H;
                    const long long cap = 2LL * kp_sm_count();
                    const int g = (int)(rows < cap ? rows : cap), sh = ilog2(C / 8);
                    cudaStream_t st = (cudaStream_t)stream;
#define KP_BWDL(ACTV)                                                                                                   \
    do {                                                                                                                \
        if (post == KP_POST_NONE)                                                                                       \
            bn_bwd_lean_k<TY, KP_POST_NONE, ACTV, false><<<g, 256, 0, st>>>(                                            \
                make_rows<const TY>(dout), make_rows<const TY>(y), make_rows<TY>(dyv), scale, shift, mean, invstd, sums, \
                1.0, pad, N, H, W, C, OH, OW, sh);                                                                      \
        else                                                                                                            \
            bn_bwd_lean_k<TY, KP_POST_POOL, ACTV, false><<<g, 256, 0, st>>>(                                            \
                make_rows<const TY>(dout), make_rows<const TY>(y), make_rows<TY>(dyv), scale, shift, mean, invstd, sums, \
                1.0, pad, N, H, W, C, OH, OW, sh);                                                                      \
    } while (0)
                    KP_ACT_SWITCH(act, KP_BWDL);
#undef KP_BWDL
                } else if (vec && fast_ok(C, P * C)) {
                    const long long rows = post == KP_POST_POOL ? (long long)N * ((H + 1) / 2) : (long long)N * H;
                    // one resident wave (2 blocks / SM at 128 registers): every block ends with 2C double atomics
                    const long long cap = 2LL * kp_sm_count();
                    const int g = (int)(rows < cap ? rows : cap), sh = ilog2(C / 8);
                    cudaStream_t st = (cudaStream_t)stream;
#define KP_BWD(POSTV) bn_act_bwd_rows_k<TG, TY, TD, POSTV, false><<<g, 256, 0, st>>>(make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dyv), scale, shift, mean, invstd, sums, 1.0, act, pad, N, H, W, C, OH, OW, sh)
                    if (post == KP_POST_NONE) KP_BWD(KP_POST_NONE);
                    else if (post == KP_POST_POOL) KP_BWD(KP_POST_POOL);
                    else if (H < 2 || W < 2) KP_BWD(KP_POST_UP);
                    else bn_act_bwd_up_k<TG, TY, TD><<<g, 256, 0, st>>>(make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dyv), scale, shift, mean, invstd, sums, act_slope(act), pad, N, H, W, C, sh);
#undef KP_BWD
                } else if (vec) {
                    Launch2D l = plan2d(P, C, 8);
                    bn_act_bwd_reduce_k<TG, TY, TD, 8><<<l.grid, l.block, 0, (cudaStream_t)stream>>>(
                        make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dyv), scale, shift, mean, invstd, sums,
                        act, post, pad, N, H, W, C, OH, OW);
                } else {
                    Launch2D l = plan2d(P, C, 1);
                    bn_act_bwd_reduce_k<TG, TY, TD, 1><<<l.grid, l.block, 0, (cudaStream_t)stream>>>(
                        make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dyv), scale, shift, mean, invstd, sums,
                        act, post, pad, N, H, W, C, OH, OW);
                }
                KP_LAUNCH_CHECK();
                return (int)KP_OK;
            });
        });
    });
}

extern "C" int kp_bn_act_bwd_apply_gather(kp_stream stream, const kp_view* dout, const kp_view* y, const kp_view* dy,
                                          const float* scale, const float* shift, const float* mean,
                                          const float* invstd, const double* sums, double count, int act, int post,
                                          int pad, int N, int H, int W, int C, float* dgamma, float* dbeta) {
    KP_CHECK_ARG(dout && y && dy && dout->ptr && y->ptr && dy->ptr && sums && scale && shift && mean && invstd &&
                     count > 0 && N > 0 && H > 0 && W > 0 && C > 0 && (post == KP_POST_NONE || post == KP_POST_POOL),
                 "kp_bn_act_bwd_apply_gather: bad arguments");
    int OH, OW;
    out_dims(post, H, W, &OH, &OW);
    const long long P = (long long)N * H * W;
    if (!(view_vec8_ok(y, C) && view_vec8_ok(dout, C) && view_vec8_ok(dy, C) && fast_ok(C, P * C))) {
        kp_set_error("kp_bn_act_bwd_apply_gather: only the vectorised NHWC path is implemented (C=%d)", C);
        return KP_ERR_UNSUPPORTED;
    }
    const long long rows = post == KP_POST_POOL ? (long long)N * ((H + 1) / 2) : (long long)N * H;
    const int g = rows_grid(rows), sh = ilog2(C / 8);
    cudaStream_t st = (cudaStream_t)stream;
    const bool views_ok = pipe_view_ok(dout, C) && pipe_view_ok(y, C) && pipe_view_ok(dy, C);
    if (pipe_enabled() && views_ok && post == KP_POST_NONE && pipe_shape_ok(C, W)) {
        const int cpr = (int)((long long)W * (C / 8) / pipe::IPC);
        const long long units = (long long)N * H * cpr;
        constexpr int smem = pipe_smem_bytes<PIPE_BWD_STAGE, PIPE_BWD_STAGES>();
#define KP_GATHN(ACTV)                                                                                                 \
    do {                                                                                                               \
        int rc_ = pipe_attr(bn_bwd_none_pipe_k<ACTV, PASS2_GATHER>, smem);                                             \
        if (rc_) return rc_;                                                                                           \
        bn_bwd_none_pipe_k<ACTV, PASS2_GATHER><<<pipe_grid(units), pipe::THREADS, smem, st>>>(                         \
            make_rows<const bf16>(dout), make_rows<const bf16>(y), make_rows<bf16>(dy), scale, shift, mean, invstd,    \
            const_cast<double*>(sums), count, pad, N, H, W, C, sh, cpr, dgamma, dbeta, bn_rev(4, N));                               \
    } while (0)
        KP_ACT_SWITCH(act, KP_GATHN);
#undef KP_GATHN
        KP_LAUNCH_CHECK();
        return KP_OK;
    }
    if (pipe_enabled() && views_ok && post == KP_POST_POOL && H % 2 == 0 && W % 2 == 0 && pipe_pool_shape_ok(C, OW)) {
        const int cpr = (int)((long long)OW * (C / 8) / PIPE_POOL_ITEMS);
        const long long units = (long long)N * OH * cpr;
        constexpr int smem = pipe_smem_bytes<PIPE_BPOOL_STAGE, PIPE_BPOOL_STAGES>();
#define KP_GATHP(ACTV)                                                                                                 \
    do {                                                                                                               \
        int rc_ = pipe_attr(bn_bwd_pool_pipe_k<ACTV, PASS2_GATHER>, smem);                                             \
        if (rc_) return rc_;                                                                                           \
        bn_bwd_pool_pipe_k<ACTV, PASS2_GATHER><<<pipe_grid(units), pipe::THREADS, smem, st>>>(                         \
            make_rows<const bf16>(dout), make_rows<const bf16>(y), make_rows<bf16>(dy), scale, shift, mean, invstd,    \
            const_cast<double*>(sums), count, pad, N, OH, OW, C, sh, cpr, dgamma, dbeta, bn_rev(4, N));                             \
    } while (0)
        KP_ACT_SWITCH(act, KP_GATHP);
#undef KP_GATHP
        KP_LAUNCH_CHECK();
        return KP_OK;
    }
    int rc = dispatch1(dout->dtype, [&](auto tg) -> int {
        return dispatch1(y->dtype, [&](auto ty) -> int {
            return dispatch1(dy->dtype, [&](auto td) -> int {
                using TG = decltype(tg);
                using TY = decltype(ty);
                using TD = decltype(td);
                if (post == KP_POST_NONE)
                    bn_act_bwd_rows_k<TG, TY, TD, KP_POST_NONE, true><<<g, 256, 0, st>>>(
                        make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dy), scale, shift, mean, invstd,
                        const_cast<double*>(sums), count, act, pad, N, H, W, C, OH, OW, sh);
                else
                    bn_act_bwd_rows_k<TG, TY, TD, KP_POST_POOL, true><<<g, 256, 0, st>>>(
                        make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dy), scale, shift, mean, invstd,
                        const_cast<double*>(sums), count, act, pad, N, H, W, C, OH, OW, sh);
                KP_LAUNCH_CHECK();
                return KP_OK;
            });
        });
    });
    if (rc) return rc;
    if (dgamma || dbeta) bn_grad_finalize_k<<<(C + 127) / 128, 128, 0, st>>>(sums, C, dgamma, dbeta);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

extern "C" int kp_bn_act_bwd_apply(kp_stream stream, const kp_view* y, const kp_view* dy, const float* scale,
                                   const float* mean, const float* invstd, const double* sums, double count, int N,
                                   int H, int W, int C, float* dgamma, float* dbeta) {
    KP_CHECK_ARG(y && dy && y->ptr && dy->ptr && sums && scale && mean && invstd && count > 0 && N > 0 && H > 0 &&
                     W > 0 && C > 0,
                 "kp_bn_act_bwd_apply: bad arguments");
    const bool vec = view_vec8_ok(y, C) && view_vec8_ok(dy, C);
    const long long P = (long long)N * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch1(y->dtype, [&](auto ty) -> int {
        return dispatch1(dy->dtype, [&](auto td) -> int {
            using TY = decltype(ty);
            using TD = decltype(td);
            if (pipe_enabled() && pipe_shape_ok(C, W) && pipe_view_ok(y, C) && pipe_view_ok(dy, C)) {
                const int cpr = (int)((long long)W * (C / 8) / pipe::IPC);
                const long long units = (long long)N * H * cpr;
                constexpr int smem = pipe_smem_bytes<PIPE_APPLY_STAGE, PIPE_APPLY_STAGES>();
                int rc_ = pipe_attr(bn_bwd_apply_pipe_k, smem);
                if (rc_) return rc_;
                bn_bwd_apply_pipe_k<<<pipe_grid(units), pipe::THREADS, smem, st>>>(
                    make_rows<const bf16>(y), make_rows<bf16>(dy), scale, mean, invstd, sums, count, N, H, W, C, ilog2(C / 8),
                    cpr, dgamma, dbeta);
            } else if (vec && fast_ok(C, P * C)) {
                bn_bwd_apply_rows_k<TY, TD><<<rows_grid((long long)N * H), 256, 0, st>>>(
                    make_view<TY>(y), make_view<TD>(dy), scale, mean, invstd, sums, count, N, H, W, C, ilog2(C / 8), dgamma,
                    dbeta);
            } else {
                if (dgamma || dbeta) bn_grad_finalize_k<<<(C + 127) / 128, 128, 0, st>>>(sums, C, dgamma, dbeta);
                long long total = P * C;
                long long g = (total + 255) / 256;
                if (g > (long long)kp_sm_count() * 16) g = (long long)kp_sm_count() * 16;
                bn_bwd_apply_generic_k<TY, TD><<<(int)g, 256, 0, st>>>(make_view<TY>(y), make_view<TD>(dy), scale, mean,
                                                                      invstd, sums, count, N, H, W, C);
            }
            KP_LAUNCH_CHECK();
            return KP_OK;
        });
    });
}
