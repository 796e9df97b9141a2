// BatchNorm(train) + activation + pool/upsample + replicate-pad, forward and backward.
// HBM-bound passes: channel-fastest thread mapping, 16-byte vector accesses (8 channels / thread)
// whenever the views allow it, per-thread register accumulation then one double atomic per
// channel per block for the reductions.
#include "kp_common.cuh"
#include <type_traits>

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
        int px = (int)(p % PW);
        long long r = p / PW;
        int py = (int)(r % PH);
        int n = (int)(r / PH);
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
            int xx = (int)(p % W);
            long long r = p / W;
            int yy = (int)(r % H);
            int n = (int)(r / H);
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

template <typename TG, typename TY, typename TD, int V>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_k(View<TG> dout, View<TY> y, View<TD> dy, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ mean,
                   const float* __restrict__ invstd, const double* __restrict__ sums, double count, int act,
                   int post, int pad, int N, int H, int W, int C, int OH, int OW) {
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * V;
    if (c0 >= C) return;
    float sc[V], sh[V], mu[V], is[V], m1[V], m2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        sc[i] = scale[c0 + i]; sh[i] = shift[c0 + i]; mu[i] = mean[c0 + i]; is[i] = invstd[c0 + i];
        m1[i] = (float)(sums[c0 + i] / count);
        m2[i] = (float)(sums[C + c0 + i] / count);
    }
    const long long P = (long long)N * H * W;
    const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.y + threadIdx.y; p < P; p += (long long)gridDim.x * blockDim.y) {
        int xx = (int)(p % W);
        long long r = p / W;
        int yy = (int)(r % H);
        int n = (int)(r / H);
        float g[V], yv[V];
        compute_dz<TG, TY, V>(dout, y, sc, sh, act, post, pad, n, yy, xx, c0, H, W, OH, OW, ry, rx, g, yv);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = sc[i] * (g[i] - m1[i] - (yv[i] - mu[i]) * is[i] * m2[i]);
        Vec<TD, V>::store(dy.at(n, yy, xx, c0), g);
    }
}


// ================================================================================================
// Fast path: NHWC with unit channel stride, 8 channels (16 bytes) per thread, C/8 a power of two <= 256.
// One block walks whole image rows: 32-bit index math only, the channel group of a thread is fixed (so the
// per-channel affine lives in registers and the reductions accumulate in registers), consecutive threads
// touch consecutive 16-byte chunks of the row.
// ================================================================================================
constexpr int V8 = 8;

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

template <typename TI, typename TO, int POST>
__global__ void __launch_bounds__(256)
bn_act_fwd_rows_k(View<TI> y, View<TO> out, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                  int pad, int N, int H, int W, int C, int OH, int OW, int cg_shift) {
    const int ncg = C >> 3;
    const int cg = threadIdx.x & (ncg - 1), c0 = cg * V8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float sc[V8], sh[V8];
#pragma unroll
    for (int i = 0; i < V8; ++i) { sc[i] = scale ? scale[c0 + i] : 1.f; sh[i] = shift ? shift[c0 + i] : 0.f; }
    const int PH = OH + 2 * pad, PW = OW + 2 * pad, rows = N * PH;
    const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / PH, py = row - n * PH;
        const int oy = min(max(py - pad, 0), OH - 1);
        int y0 = 0, y1 = 0;
        float ly = 0.f;
        if (POST == KP_POST_UP) {
            float sy = ry * oy;
            y0 = (int)sy; ly = sy - y0; y1 = min(y0 + 1, H - 1);
        }
        for (int px = x0; px < PW; px += xstep) {
            const int ox = min(max(px - pad, 0), OW - 1);
            float v[V8];
            if (POST == KP_POST_NONE) {
                load_act<TI, V8>(y, n, oy, ox, c0, sc, sh, act, v);
            } else if (POST == KP_POST_POOL) {
                float a[V8], b[V8], c[V8];
                load_act<TI, V8>(y, n, 2 * oy, 2 * ox, c0, sc, sh, act, v);
                load_act<TI, V8>(y, n, 2 * oy, 2 * ox + 1, c0, sc, sh, act, a);
                load_act<TI, V8>(y, n, 2 * oy + 1, 2 * ox, c0, sc, sh, act, b);
                load_act<TI, V8>(y, n, 2 * oy + 1, 2 * ox + 1, c0, sc, sh, act, c);
#pragma unroll
                for (int i = 0; i < V8; ++i) v[i] = fmaxf(fmaxf(v[i], a[i]), fmaxf(b[i], c[i]));
            } else {
                float sx = rx * ox;
                int xa = (int)sx;
                float lx = sx - xa;
                int xb = min(xa + 1, W - 1);
                float a00[V8], a01[V8], a10[V8], a11[V8];
                load_act<TI, V8>(y, n, y0, xa, c0, sc, sh, act, a00);
                load_act<TI, V8>(y, n, y0, xb, c0, sc, sh, act, a01);
                load_act<TI, V8>(y, n, y1, xa, c0, sc, sh, act, a10);
                load_act<TI, V8>(y, n, y1, xb, c0, sc, sh, act, a11);
#pragma unroll
                for (int i = 0; i < V8; ++i)
                    v[i] = (1.f - ly) * ((1.f - lx) * a00[i] + lx * a01[i]) + ly * ((1.f - lx) * a10[i] + lx * a11[i]);
            }
            Vec<TO, V8>::store(out.at(n, py, px, c0), v);
        }
    }
}

// MODE 0: reduce only (BatchNorm pass 1)   MODE 1: reduce + write dy = dz (no BatchNorm)   MODE 2: apply (pass 2)
template <typename TG, typename TY, typename TD, int POST, int MODE>
__global__ void __launch_bounds__(256)
bn_act_bwd_rows_k(View<TG> dout, View<TY> y, View<TD> dy, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                  double* sums, double count, int act, int pad, int N, int H, int W, int C, int OH, int OW, int cg_shift) {
    __shared__ float red[2][256 * V8];
    const int ncg = C >> 3;
    const int cg = threadIdx.x & (ncg - 1), c0 = cg * V8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float sc[V8], sh[V8], mu[V8], is[V8], s1[V8], s2[V8];
#pragma unroll
    for (int i = 0; i < V8; ++i) {
        sc[i] = scale ? scale[c0 + i] : 1.f; sh[i] = shift ? shift[c0 + i] : 0.f;
        mu[i] = mean ? mean[c0 + i] : 0.f; is[i] = invstd ? invstd[c0 + i] : 1.f;
        if (MODE == 2) { s1[i] = (float)(sums[c0 + i] / count); s2[i] = (float)(sums[C + c0 + i] / count); }
        else { s1[i] = 0.f; s2[i] = 0.f; }
    }
    const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    // consume one element: g = gradient w.r.t. the activation output, yv = raw conv output
    auto emit = [&](int n, int yy, int xx, float (&g)[V8], const float (&yv)[V8]) {
#pragma unroll
        for (int i = 0; i < V8; ++i) g[i] *= act_grad(fmaf(yv[i], sc[i], sh[i]), act);
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < V8; ++i) g[i] = sc[i] * (g[i] - s1[i] - (yv[i] - mu[i]) * is[i] * s2[i]);
            Vec<TD, V8>::store(dy.at(n, yy, xx, c0), g);
        } else {
#pragma unroll
            for (int i = 0; i < V8; ++i) { s1[i] += g[i]; s2[i] += g[i] * (yv[i] - mu[i]) * is[i]; }
            if (MODE == 1) Vec<TD, V8>::store(dy.at(n, yy, xx, c0), g);
        }
    };
    if (POST == KP_POST_POOL) {
        const int HW2 = (H + 1) >> 1, WW2 = (W + 1) >> 1, rows = N * HW2;
        for (int row = blockIdx.x; row < rows; row += gridDim.x) {
            const int n = row / HW2, wy = (row - n * HW2) * 2;
            const RowFold fy(wy >> 1, OH, pad);
            for (int wi = x0; wi < WW2; wi += xstep) {
                const int wx = wi * 2;
                const bool full = (wy + 1 < H) && (wx + 1 < W);
                float yv[4][V8], best[V8], t[V8];
                int bi[V8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int yy = wy + (q >> 1), xx = wx + (q & 1);
                    if (yy < H && xx < W) Vec<TY, V8>::load(y.at(n, yy, xx, c0), yv[q]);
                    else {
#pragma unroll
                        for (int i = 0; i < V8; ++i) yv[q][i] = 0.f;
                    }
                }
#pragma unroll
                for (int i = 0; i < V8; ++i) { best[i] = -INFINITY; bi[i] = 0; t[i] = 0.f; }
                if (full) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int i = 0; i < V8; ++i) {
                            float a = apply_act(fmaf(yv[q][i], sc[i], sh[i]), act);
                            if (a > best[i] || q == 0) { best[i] = a; bi[i] = q; }      // first maximum wins
                        }
                    const RowFold fx(wx >> 1, OW, pad);
                    fold_acc<TG>(dout, n, fy, fx, c0, 1.f, t);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int yy = wy + (q >> 1), xx = wx + (q & 1);
                    if (yy < H && xx < W) {
                        float g[V8];
#pragma unroll
                        for (int i = 0; i < V8; ++i) g[i] = (full && bi[i] == q) ? t[i] : 0.f;
                        emit(n, yy, xx, g, yv[q]);
                    }
                }
            }
        }
    } else {
        const int rows = N * H;
        for (int row = blockIdx.x; row < rows; row += gridDim.x) {
            const int n = row / H, yy = row - n * H;
            if (POST == KP_POST_NONE) {
                const RowFold fy(yy, OH, pad);
                for (int xx = x0; xx < W; xx += xstep) {
                    float g[V8], yv[V8];
#pragma unroll
                    for (int i = 0; i < V8; ++i) g[i] = 0.f;
                    const RowFold fx(xx, OW, pad);
                    fold_acc<TG>(dout, n, fy, fx, c0, 1.f, g);
                    Vec<TY, V8>::load(y.at(n, yy, xx, c0), yv);
                    emit(n, yy, xx, g, yv);
                }
            } else {   // bilinear x2 (align_corners=True) backward: gather the <= 4x4 outputs that read (yy,xx)
                int ylo = 0, yhi = OH - 1;
                if (H > 1) { ylo = max(0, 2 * yy - 2); yhi = min(OH - 1, 2 * yy + 3); }
                for (int xx = x0; xx < W; xx += xstep) {
                    int xlo = 0, xhi = OW - 1;
                    if (W > 1) { xlo = max(0, 2 * xx - 2); xhi = min(OW - 1, 2 * xx + 3); }
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
                        for (int X = xlo; X <= xhi; ++X) {
                            float sx = rx * X;
                            int xa = (int)sx;
                            float lx = sx - xa;
                            int xb = min(xa + 1, W - 1);
                            float wxv = (xa == xx ? 1.f - lx : 0.f) + (xb == xx ? lx : 0.f);
                            if (wxv == 0.f) continue;
                            const RowFold fx(X, OW, pad);
                            fold_acc<TG>(dout, n, fy, fx, c0, wyv * wxv, g);
                        }
                    }
                    Vec<TY, V8>::load(y.at(n, yy, xx, c0), yv);
                    emit(n, yy, xx, g, yv);
                }
            }
        }
    }
    if (MODE != 2) {
#pragma unroll
        for (int i = 0; i < V8; ++i) { red[0][threadIdx.x * V8 + i] = s1[i]; red[1][threadIdx.x * V8 + i] = s2[i]; }
        __syncthreads();
        // threads 0 .. C-1: one channel each, sum over the 256/ncg threads that share its channel group
        for (int ch = threadIdx.x; ch < C; ch += 256) {
            const int g8 = ch >> 3, i = ch & 7;
            float a = 0.f, b = 0.f;
            for (int t = g8; t < 256; t += ncg) { a += red[0][t * V8 + i]; b += red[1][t * V8 + i]; }
            atomicAdd(&sums[ch], (double)a);
            atomicAdd(&sums[C + ch], (double)b);
        }
    }
}

static bool fast_ok(int C, long long rows_px) {
    if (C % 8) return false;
    int ncg = C / 8;
    return ncg >= 1 && ncg <= 256 && (ncg & (ncg - 1)) == 0 && rows_px < (1LL << 31);
}
static int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }
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

extern "C" int kp_bn_act_fwd(kp_stream stream, const kp_view* y, const kp_view* out, const float* scale,
                             const float* shift, int act, int post, int pad, int N, int H, int W, int C) {
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
            if (vec && fast_ok(C, P * C)) {
                const int g = rows_grid((long long)N * (OH + 2 * pad)), sh = ilog2(C / 8);
                cudaStream_t st = (cudaStream_t)stream;
#define KP_FWD(POSTV) bn_act_fwd_rows_k<TI, TO, POSTV><<<g, 256, 0, st>>>(make_view<TI>(y), make_view<TO>(out), scale, shift, act, pad, N, H, W, C, OH, OW, sh)
                if (post == KP_POST_NONE) KP_FWD(KP_POST_NONE);
                else if (post == KP_POST_POOL) KP_FWD(KP_POST_POOL);
                else KP_FWD(KP_POST_UP);
#undef KP_FWD
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
                if (vec && fast_ok(C, P * C)) {
                    const long long rows = post == KP_POST_POOL ? (long long)N * ((H + 1) / 2) : (long long)N * H;
                    const int g = rows_grid(rows), sh = ilog2(C / 8);
                    cudaStream_t st = (cudaStream_t)stream;
#define KP_BWD(POSTV, MODEV) bn_act_bwd_rows_k<TG, TY, TD, POSTV, MODEV><<<g, 256, 0, st>>>(make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dyv), scale, shift, mean, invstd, sums, 1.0, act, pad, N, H, W, C, OH, OW, sh)
                    if (dyv->ptr) {
                        if (post == KP_POST_NONE) KP_BWD(KP_POST_NONE, 1);
                        else if (post == KP_POST_POOL) KP_BWD(KP_POST_POOL, 1);
                        else KP_BWD(KP_POST_UP, 1);
                    } else {
                        if (post == KP_POST_NONE) KP_BWD(KP_POST_NONE, 0);
                        else if (post == KP_POST_POOL) KP_BWD(KP_POST_POOL, 0);
                        else KP_BWD(KP_POST_UP, 0);
                    }
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

extern "C" int kp_bn_act_bwd_apply(kp_stream stream, const kp_view* dout, const kp_view* y, const kp_view* dy,
                                   const float* scale, const float* shift, const float* mean, const float* invstd,
                                   const double* sums, double count, int act, int post, int pad, int N, int H, int W,
                                   int C) {
    KP_CHECK_ARG(dout && y && dy && dout->ptr && y->ptr && dy->ptr && sums && scale && shift && mean && invstd &&
                     count > 0 && N > 0 && H > 0 && W > 0 && C > 0,
                 "kp_bn_act_bwd_apply: bad arguments");
    int OH, OW;
    out_dims(post, H, W, &OH, &OW);
    const bool vec = view_vec8_ok(y, C) && view_vec8_ok(dout, C) && view_vec8_ok(dy, C);
    const long long P = (long long)N * H * W;
    return dispatch1(dout->dtype, [&](auto tg) -> int {
        return dispatch1(y->dtype, [&](auto ty) -> int {
            return dispatch1(dy->dtype, [&](auto td) -> int {
                using TG = decltype(tg);
                using TY = decltype(ty);
                using TD = decltype(td);
                if (vec && fast_ok(C, P * C)) {
                    const long long rows = post == KP_POST_POOL ? (long long)N * ((H + 1) / 2) : (long long)N * H;
                    const int g = rows_grid(rows), sh = ilog2(C / 8);
                    cudaStream_t st = (cudaStream_t)stream;
#define KP_APP(POSTV) bn_act_bwd_rows_k<TG, TY, TD, POSTV, 2><<<g, 256, 0, st>>>(make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dy), scale, shift, mean, invstd, const_cast<double*>(sums), count, act, pad, N, H, W, C, OH, OW, sh)
                    if (post == KP_POST_NONE) KP_APP(KP_POST_NONE);
                    else if (post == KP_POST_POOL) KP_APP(KP_POST_POOL);
                    else KP_APP(KP_POST_UP);
#undef KP_APP
                } else if (vec) {
                    Launch2D l = plan2d(P, C, 8);
                    bn_act_bwd_apply_k<TG, TY, TD, 8><<<l.grid, l.block, 0, (cudaStream_t)stream>>>(
                        make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dy), scale, shift, mean, invstd, sums,
                        count, act, post, pad, N, H, W, C, OH, OW);
                } else {
                    Launch2D l = plan2d(P, C, 1);
                    bn_act_bwd_apply_k<TG, TY, TD, 1><<<l.grid, l.block, 0, (cudaStream_t)stream>>>(
                        make_view<TG>(dout), make_view<TY>(y), make_view<TD>(dy), scale, shift, mean, invstd, sums,
                        count, act, post, pad, N, H, W, C, OH, OW);
                }
                KP_LAUNCH_CHECK();
                return (int)KP_OK;
            });
        });
    });
}
