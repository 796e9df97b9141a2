// Keypoint bottleneck: marginal spatial soft-max -> K (y,x) keypoints, gaussian-like heat-map render,
// Transporter feature transport ('max' combine), masked L2 loss — forward and closed-form backward.
// All latency/HBM-bound: one block per (n,k) plane, warp-shuffle reductions, coalesced plane reads.
#include "kp_common.cuh"

namespace {

template <typename F>
int dispatch1(int dt, F&& f) {
    if (dt == KP_F32) return f(float{});
    if (dt == KP_BF16) return f(bf16{});
    kp_set_error("bad dtype %d", dt);
    return KP_ERR_ARG;
}

__device__ __forceinline__ float ruler(int i, int n) { return n > 1 ? (float)i / (float)(n - 1) : 0.f; }

// block-wide sum of one float; every thread gets the result.  blockDim.x multiple of 32, <= 1024.
__device__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    float t = (lane < nw) ? scratch[lane] : 0.f;
    t = warp_sum(t);
    return t;
}

// softmax of s[0:n] in place by one warp, returns expectation against ruler
__device__ float warp_softmax_expect(float* s, int n, int lane) {
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) mx = fmaxf(mx, s[i]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) sum += expf(s[i] - mx);
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    float e = 0.f;
    for (int i = lane; i < n; i += 32) {
        float p = expf(s[i] - lse);        // exp(log_softmax), functional.py:11-14,44
        s[i] = p;
        e += p * ruler(i, n);
    }
    return warp_sum(e);
}

__global__ void __launch_bounds__(256) ssm_fwd_k(const float* __restrict__ heat, int h, int w, float* k, float* ph,
                                                 float* pw) {
    extern __shared__ float sm[];
    float* a = sm;          // [h] row means
    float* b = sm + h;      // [w] column means
    const float* pl = heat + (long long)blockIdx.x * h * w;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = wid; i < h; i += nw) {
        float s = 0.f;
        for (int j = lane; j < w; j += 32) s += pl[(long long)i * w + j];
        s = warp_sum(s);
        if (lane == 0) a[i] = s / (float)w;
    }
    for (int j = threadIdx.x; j < w; j += blockDim.x) {
        float s = 0.f;
        for (int i = 0; i < h; ++i) s += pl[(long long)i * w + j];
        b[j] = s / (float)h;
    }
    __syncthreads();
    if (wid == 0) {
        float ky = warp_softmax_expect(a, h, lane);
        if (lane == 0) k[blockIdx.x * 2 + 0] = ky;
    } else if (wid == 1) {
        float kx = warp_softmax_expect(b, w, lane);
        if (lane == 0) k[blockIdx.x * 2 + 1] = kx;
    }
    __syncthreads();
    if (ph) for (int i = threadIdx.x; i < h; i += blockDim.x) ph[(long long)blockIdx.x * h + i] = a[i];
    if (pw) for (int j = threadIdx.x; j < w; j += blockDim.x) pw[(long long)blockIdx.x * w + j] = b[j];
}

// da_i = p_i (y_i - k_y) dk_y ; db_j likewise ; dheat[i,j] = da_i / w + db_j / h
__global__ void __launch_bounds__(256) ssm_bwd_k(const float* __restrict__ dk, const float* __restrict__ k,
                                                 const float* __restrict__ ph, const float* __restrict__ pw, int h,
                                                 int w, float* dheat) {
    extern __shared__ float sm[];
    float* da = sm;
    float* db = sm + h;
    const long long pl = blockIdx.x;
    const float ky = k[pl * 2], kx = k[pl * 2 + 1], gy = dk[pl * 2], gx = dk[pl * 2 + 1];
    for (int i = threadIdx.x; i < h; i += blockDim.x) da[i] = ph[pl * h + i] * (ruler(i, h) - ky) * gy / (float)w;
    for (int j = threadIdx.x; j < w; j += blockDim.x) db[j] = pw[pl * w + j] * (ruler(j, w) - kx) * gx / (float)h;
    __syncthreads();
    float* o = dheat + pl * h * w;
    for (int e = threadIdx.x; e < h * w; e += blockDim.x) o[e] = da[e / w] + db[e % w];
}

__device__ __forceinline__ float gauss(float yi, float xj, float ky, float kx, float two_s2, float eps) {
    float dy = yi - ky, dx = xj - kx;
    return expf(-sqrtf(dy * dy + dx * dx + eps) / two_s2);
}

__global__ void __launch_bounds__(256) gaussian_fwd_k(const float* __restrict__ k, int h, int w, float two_s2,
                                                      float eps, float* m) {
    const long long pl = blockIdx.x;
    const float ky = k[pl * 2], kx = k[pl * 2 + 1];
    float* o = m + pl * h * w;
    for (int e = threadIdx.x; e < h * w; e += blockDim.x) o[e] = gauss(ruler(e / w, h), ruler(e % w, w), ky, kx, two_s2, eps);
}

template <typename TG>
__global__ void __launch_bounds__(256)
gaussian_bwd_k(View<TG> dm, int pad, const float* __restrict__ k, const int* __restrict__ argmax, int K, int h, int w,
               float two_s2, float eps, float* dk) {
    __shared__ float scratch[32];
    const int pl = blockIdx.x, n = pl / K, kk = pl % K;
    const float ky = k[pl * 2], kx = k[pl * 2 + 1];
    float sy = 0.f, sx = 0.f;
    for (int e = threadIdx.x; e < h * w; e += blockDim.x) {
        int i = e / w, j = e % w;
        if (argmax && argmax[(long long)n * h * w + e] != kk) continue;
        float g = 0.f;
        {   // gradient at (i,j), replicate-pad border folded in
            const int c = argmax ? 0 : kk;
            if (!pad) {
                g = to_f(*dm.at(n, i, j, c));
            } else {
                int ys[3], xs[3], ny = 0, nx = 0;
                ys[ny++] = i + 1; if (i == 0) ys[ny++] = 0; if (i == h - 1) ys[ny++] = h + 1;
                xs[nx++] = j + 1; if (j == 0) xs[nx++] = 0; if (j == w - 1) xs[nx++] = w + 1;
                for (int a = 0; a < ny; ++a)
                    for (int b = 0; b < nx; ++b) g += to_f(*dm.at(n, ys[a], xs[b], c));
            }
        }
        float yi = ruler(i, h), xj = ruler(j, w);
        float dy = yi - ky, dx = xj - kx;
        float r = sqrtf(dy * dy + dx * dx + eps);
        float mv = expf(-r / two_s2);
        float t = g * mv / (two_s2 * r);
        sy += t * dy;
        sx += t * dx;
    }
    sy = block_sum(sy, scratch);
    sx = block_sum(sx, scratch);
    if (threadIdx.x == 0) { dk[pl * 2] = sy; dk[pl * 2 + 1] = sx; }
}

template <typename TP, typename TO>
__global__ void __launch_bounds__(256)
transport_fwd_k(View<TP> phi_s, View<TP> phi_t, const float* __restrict__ k_s, const float* __restrict__ k_t,
                View<TO> out, int pad, float* mask_s, float* mask_t, int* argmax_t, int N, int h, int w, int C, int K,
                float two_s2, float eps) {
    const int PH = h + 2 * pad, PW = w + 2 * pad;
    const long long total = (long long)N * PH * PW * C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C);
        long long p = idx / C;
        int px = (int)(p % PW);
        long long r = p / PW;
        int py = (int)(r % PH);
        int n = (int)(r / PH);
        int i = min(max(py - pad, 0), h - 1), j = min(max(px - pad, 0), w - 1);
        float yi = ruler(i, h), xj = ruler(j, w);
        float ms = -INFINITY, mt = -INFINITY;
        int am = 0;
        for (int q = 0; q < K; ++q) {
            const float* ks_ = k_s + ((long long)n * K + q) * 2;
            const float* kt_ = k_t + ((long long)n * K + q) * 2;
            float a = gauss(yi, xj, ks_[0], ks_[1], two_s2, eps);
            float b = gauss(yi, xj, kt_[0], kt_[1], two_s2, eps);
            ms = fmaxf(ms, a);
            if (b > mt) { mt = b; am = q; }          // first maximum wins, as torch.max(dim=1)
        }
        float ps = to_f(*phi_s.at(n, i, j, c)), pt = to_f(*phi_t.at(n, i, j, c));
        from_f(out.at(n, py, px, c), ps * (1.f - ms) * (1.f - mt) + pt * mt);
        if (c == 0 && py - pad == i && px - pad == j) {
            long long e = ((long long)n * h + i) * w + j;
            if (mask_s) mask_s[e] = ms;
            if (mask_t) mask_t[e] = mt;
            if (argmax_t) argmax_t[e] = am;
        }
    }
}

// Vectorised 'max' transport (the shapes the nets use: C % 8 == 0, C / 8 a power of two <= 32, 16-byte aligned views).
// G = C / 8 threads share one padded output pixel: the 2K rendered maps are evaluated ONCE per pixel (keypoints strided over
// the group, max / first-argmax folded with shuffles) and every thread moves one 128-bit vector of 8 channels.
template <typename TP, typename TO, int G>
__global__ void __launch_bounds__(256)
transport_fwd_vec_k(View<TP> phi_s, View<TP> phi_t, const float* __restrict__ k_s, const float* __restrict__ k_t,
                    View<TO> out, int pad, float* mask_s, float* mask_t, int* argmax_t, int N, int h, int w, int K,
                    float two_s2, float eps) {
    const unsigned PH = h + 2 * pad, PW = w + 2 * pad;
    const unsigned total = (unsigned)N * PH * PW;
    const unsigned gl = threadIdx.x % G, ppb = 256 / G;
    const unsigned step = gridDim.x * ppb;
    const unsigned trips = (total + step - 1) / step;
    unsigned p = blockIdx.x * ppb + threadIdx.x / G;
    for (unsigned it = 0; it < trips; ++it, p += step) {
        const bool valid = p < total;
        const unsigned pc = valid ? p : 0;
        const unsigned r = pc / PW, px = pc - r * PW;
        const unsigned n = r / PH, py = r - n * PH;
        const int i = min(max((int)py - pad, 0), h - 1), j = min(max((int)px - pad, 0), w - 1);
        const float yi = ruler(i, h), xj = ruler(j, w);
        float ms = -INFINITY, mt = -INFINITY;
        int am = K;
        const float2* ks2 = reinterpret_cast<const float2*>(k_s) + (size_t)n * K;
        const float2* kt2 = reinterpret_cast<const float2*>(k_t) + (size_t)n * K;
        for (int q = gl; q < K; q += G) {
            const float2 a2 = ks2[q], b2 = kt2[q];
            const float a = gauss(yi, xj, a2.x, a2.y, two_s2, eps);
            const float b = gauss(yi, xj, b2.x, b2.y, two_s2, eps);
            ms = fmaxf(ms, a);
            if (b > mt) { mt = b; am = q; }
        }
#pragma unroll
        for (int o = 1; o < G; o <<= 1) {
            ms = fmaxf(ms, __shfl_xor_sync(0xffffffffu, ms, o));
            const float omt = __shfl_xor_sync(0xffffffffu, mt, o);
            const int oam = __shfl_xor_sync(0xffffffffu, am, o);
            if (omt > mt || (omt == mt && oam < am)) { mt = omt; am = oam; }   // first maximum wins, as torch.max(dim=1)
        }
        if (!valid) continue;
        float ps[8], pt[8], o8[8];
        Vec<TP, 8>::load(phi_s.at(n, i, j, gl * 8), ps);
        Vec<TP, 8>::load(phi_t.at(n, i, j, gl * 8), pt);
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = ps[e] * (1.f - ms) * (1.f - mt) + pt[e] * mt;
        Vec<TO, 8>::store(out.at(n, py, px, gl * 8), o8);
        if (gl == 0 && (int)py - pad == i && (int)px - pad == j) {
            const size_t e = ((size_t)n * h + i) * w + j;
            if (mask_s) mask_s[e] = ms;
            if (mask_t) mask_t[e] = mt;
            if (argmax_t) argmax_t[e] = am;
        }
    }
}

template <typename TG, typename TP, typename TD, int G>
__global__ void __launch_bounds__(256)
transport_bwd_vec_k(View<TG> dout, int pad, View<TP> phi_s, View<TP> phi_t, const float* __restrict__ mask_s,
                    const float* __restrict__ mask_t, View<TD> dphi_t, float* dmask_t, int N, int h, int w) {
    const unsigned total = (unsigned)N * h * w;
    const unsigned gl = threadIdx.x % G, ppb = 256 / G;
    const unsigned step = gridDim.x * ppb;
    const unsigned trips = (total + step - 1) / step;
    unsigned p = blockIdx.x * ppb + threadIdx.x / G;
    for (unsigned it = 0; it < trips; ++it, p += step) {
        const bool valid = p < total;
        const unsigned pc = valid ? p : 0;
        const unsigned r = pc / (unsigned)w;
        const int j = (int)(pc - r * w);
        const int n = (int)(r / (unsigned)h), i = (int)(r - (unsigned)n * h);
        const float ms = mask_s[pc], mt = mask_t[pc];
        float g[8], ps[8], pt[8], d8[8];
        Vec<TG, 8>::load(dout.at(n, i + pad, j + pad, gl * 8), g);
        if (pad && (i == 0 || i == h - 1 || j == 0 || j == w - 1)) {      // fold the replicate-pad border into the edge pixel
            int ys[3], xs[3], ny = 0, nx = 0;
            ys[ny++] = i + 1; if (i == 0) ys[ny++] = 0; if (i == h - 1) ys[ny++] = h + 1;
            xs[nx++] = j + 1; if (j == 0) xs[nx++] = 0; if (j == w - 1) xs[nx++] = w + 1;
            for (int a = 0; a < ny; ++a)
                for (int b = 0; b < nx; ++b) {
                    if (a == 0 && b == 0) continue;
                    float t[8];
                    Vec<TG, 8>::load(dout.at(n, ys[a], xs[b], gl * 8), t);
#pragma unroll
                    for (int e = 0; e < 8; ++e) g[e] += t[e];
                }
        }
        Vec<TP, 8>::load(phi_s.at(n, i, j, gl * 8), ps);
        Vec<TP, 8>::load(phi_t.at(n, i, j, gl * 8), pt);
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            d8[e] = g[e] * mt;
            acc += g[e] * (pt[e] - ps[e] * (1.f - ms));
        }
#pragma unroll
        for (int o = 1; o < G; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (!valid) continue;
        Vec<TD, 8>::store(dphi_t.at(n, i, j, gl * 8), d8);
        if (gl == 0) dmask_t[pc] = acc;
    }
}

static inline int transport_group(int C) {
    const int g = C / 8;
    return (C % 8 == 0 && g >= 1 && g <= 32 && (g & (g - 1)) == 0) ? g : 0;
}

// one warp per pixel: dphi_t = dout * M_t ; dmask_t = sum_c dout (phi_t - phi_s (1 - M_s))
template <typename TG, typename TP, typename TD>
__global__ void __launch_bounds__(256)
transport_bwd_k(View<TG> dout, int pad, View<TP> phi_s, View<TP> phi_t, const float* __restrict__ mask_s,
                const float* __restrict__ mask_t, View<TD> dphi_t, float* dmask_t, int N, int h, int w, int C) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long P = (long long)N * h * w;
    for (long long p = warp; p < P; p += nwarps) {
        int j = (int)(p % w);
        long long r = p / w;
        int i = (int)(r % h);
        int n = (int)(r / h);
        float ms = mask_s[p], mt = mask_t[p];
        int ys[3], xs[3], ny = 0, nx = 0;
        if (pad) {
            ys[ny++] = i + 1; if (i == 0) ys[ny++] = 0; if (i == h - 1) ys[ny++] = h + 1;
            xs[nx++] = j + 1; if (j == 0) xs[nx++] = 0; if (j == w - 1) xs[nx++] = w + 1;
        } else { ys[ny++] = i; xs[nx++] = j; }
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) {
            float g = 0.f;
            for (int a = 0; a < ny; ++a)
                for (int b = 0; b < nx; ++b) g += to_f(*dout.at(n, ys[a], xs[b], c));
            float ps = to_f(*phi_s.at(n, i, j, c)), pt = to_f(*phi_t.at(n, i, j, c));
            from_f(dphi_t.at(n, i, j, c), g * mt);
            acc += g * (pt - ps * (1.f - ms));
        }
        acc = warp_sum(acc);
        if (lane == 0) dmask_t[p] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// The non-default combine modes of TransporterNet.forward (models/transporter.py:41-50):
//   KP_COMBINE_SUM : M = clamp(sum_k m_k, 0, 1) for source and target, then the same blend as 'max'
//   KP_COMBINE_LOOP: phi <- phi (1 - m_s,k)(1 - m_t,k) + phi_t m_t,k for k = 0..K-1.  The recurrence is affine in phi with
//                    per-pixel coefficients, phi_K = phi_s A + phi_t B (A_{k+1} = A_k c_k, B_{k+1} = B_k c_k + m_t,k,
//                    c_k = (1 - m_s,k)(1 - m_t,k)), so it costs one pass over the keypoints per pixel, not per channel.
// KP_COMBINE_MAX is accepted too (the fused trainer keeps using kp_transport_fwd / _bwd for it).
// ------------------------------------------------------------------------------------------------
template <int MODE, typename TP, typename TO>
__global__ void __launch_bounds__(256)
transport_mode_fwd_k(View<TP> phi_s, View<TP> phi_t, const float* __restrict__ k_s, const float* __restrict__ k_t,
                     View<TO> out, int pad, float* mask_s, float* mask_t, int* aux, float* coef, int N, int h, int w, int C,
                     int K, float two_s2, float eps) {
    const int PH = h + 2 * pad, PW = w + 2 * pad;
    const long long total = (long long)N * PH * PW * C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C);
        long long p = idx / C;
        int px = (int)(p % PW);
        long long r = p / PW;
        int py = (int)(r % PH);
        int n = (int)(r / PH);
        int i = min(max(py - pad, 0), h - 1), j = min(max(px - pad, 0), w - 1);
        float yi = ruler(i, h), xj = ruler(j, w);
        float ms = MODE == KP_COMBINE_MAX ? -INFINITY : 0.f, mt = ms, A = 1.f, B = 0.f;
        int ax = 0;
        for (int q = 0; q < K; ++q) {
            const float* ks_ = k_s + ((long long)n * K + q) * 2;
            const float* kt_ = k_t + ((long long)n * K + q) * 2;
            float a = gauss(yi, xj, ks_[0], ks_[1], two_s2, eps);
            float b = gauss(yi, xj, kt_[0], kt_[1], two_s2, eps);
            if (MODE == KP_COMBINE_MAX) { ms = fmaxf(ms, a); if (b > mt) { mt = b; ax = q; } }
            else if (MODE == KP_COMBINE_SUM) { ms += a; mt += b; }
            else { const float cq = (1.f - a) * (1.f - b); A *= cq; B = fmaf(B, cq, b); ms = a; mt = b; }
        }
        float ps = to_f(*phi_s.at(n, i, j, c)), pt = to_f(*phi_t.at(n, i, j, c));
        float o;
        if (MODE == KP_COMBINE_LOOP) {
            o = ps * A + pt * B;
        } else {
            if (MODE == KP_COMBINE_SUM) {
                ax = (mt >= 0.f && mt <= 1.f) ? 1 : 0;          // clamp passes the gradient inside [0, 1] (inclusive, as torch)
                ms = fminf(fmaxf(ms, 0.f), 1.f); mt = fminf(fmaxf(mt, 0.f), 1.f);
            }
            o = ps * (1.f - ms) * (1.f - mt) + pt * mt;
        }
        from_f(out.at(n, py, px, c), o);
        if (c == 0 && py - pad == i && px - pad == j) {
            long long e = ((long long)n * h + i) * w + j;
            if (mask_s) mask_s[e] = ms;
            if (mask_t) mask_t[e] = mt;
            if (aux) aux[e] = ax;
            if (coef) { coef[2 * e] = A; coef[2 * e + 1] = B; }
        }
    }
}

// one warp per pixel: dphi_t and the gradient w.r.t. every rendered target map m_t[n][k][i][j]
template <int MODE, typename TG, typename TP, typename TD>
__global__ void __launch_bounds__(256)
transport_mode_bwd_k(View<TG> dout, int pad, View<TP> phi_s, View<TP> phi_t, const float* __restrict__ k_s,
                     const float* __restrict__ k_t, const float* __restrict__ mask_s, const float* __restrict__ mask_t,
                     const int* __restrict__ aux, const float* __restrict__ coef, View<TD> dphi_t, float* dm_t, int N, int h,
                     int w, int C, int K, float two_s2, float eps) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long P = (long long)N * h * w, hw = (long long)h * w;
    for (long long p = warp; p < P; p += nwarps) {
        int j = (int)(p % w);
        long long r = p / w;
        int i = (int)(r % h);
        int n = (int)(r / h);
        const float ms = mask_s[p], mt = mask_t[p];
        const float A = MODE == KP_COMBINE_LOOP ? coef[2 * p] : 0.f, B = MODE == KP_COMBINE_LOOP ? coef[2 * p + 1] : 0.f;
        int ys[3], xs[3], ny = 0, nx = 0;
        if (pad) {
            ys[ny++] = i + 1; if (i == 0) ys[ny++] = 0; if (i == h - 1) ys[ny++] = h + 1;
            xs[nx++] = j + 1; if (j == 0) xs[nx++] = 0; if (j == w - 1) xs[nx++] = w + 1;
        } else { ys[ny++] = i; xs[nx++] = j; }
        float acc = 0.f, accA = 0.f;
        for (int c = lane; c < C; c += 32) {
            float g = 0.f;
            for (int a = 0; a < ny; ++a)
                for (int b = 0; b < nx; ++b) g += to_f(*dout.at(n, ys[a], xs[b], c));
            float ps = to_f(*phi_s.at(n, i, j, c)), pt = to_f(*phi_t.at(n, i, j, c));
            if (MODE == KP_COMBINE_LOOP) {
                from_f(dphi_t.at(n, i, j, c), g * B);
                accA += g * ps;                                   // d/dA
                acc += g * pt;                                    // d/dB
            } else {
                from_f(dphi_t.at(n, i, j, c), g * mt);
                acc += g * (pt - ps * (1.f - ms));
            }
        }
        acc = warp_sum(acc);
        accA = warp_sum(accA);
        float* dm = dm_t + (long long)n * K * hw + (long long)i * w + j;
        if (MODE == KP_COMBINE_MAX) {
            const int ax = aux[p];
            for (int q = lane; q < K; q += 32) dm[q * hw] = q == ax ? acc : 0.f;
        } else if (MODE == KP_COMBINE_SUM) {
            const float v = aux[p] ? acc : 0.f;
            for (int q = lane; q < K; q += 32) dm[q * hw] = v;
        } else if (lane == 0) {
            // reverse the recurrence: state before step q is recovered by dividing out c_q (c_q > 0: both maps are < 1)
            float yi = ruler(i, h), xj = ruler(j, w);
            float gA = accA, gB = acc, Aq = A, Bq = B;
            for (int q = K - 1; q >= 0; --q) {
                const float* ks_ = k_s + ((long long)n * K + q) * 2;
                const float* kt_ = k_t + ((long long)n * K + q) * 2;
                const float a = gauss(yi, xj, ks_[0], ks_[1], two_s2, eps);
                const float b = gauss(yi, xj, kt_[0], kt_[1], two_s2, eps);
                const float cq = (1.f - a) * (1.f - b);
                Bq = (Bq - b) / cq;                               // B_q, A_q: the state entering step q
                Aq = Aq / cq;
                const float gc = gA * Aq + gB * Bq;
                dm[q * hw] = gB - gc * (1.f - a);
                gA *= cq;
                gB *= cq;
            }
        }
    }
}

__global__ void __launch_bounds__(256) l2_loss_k(const float* __restrict__ xhat, const float* __restrict__ target,
                                                 const float* __restrict__ mask, long long numel, float gscale,
                                                 double* loss_sum, float* dxhat) {
    __shared__ float scratch[32];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel;
         i += (long long)gridDim.x * blockDim.x) {
        float d = xhat[i] - target[i];
        float mk = mask ? mask[i] : 1.f;
        s += d * d * mk;
        if (dxhat) dxhat[i] = 2.f * d * mk * gscale;
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, (double)s);
}

static int grid_for(long long work, int block) {
    long long g = (work + block - 1) / block;
    long long cap = (long long)kp_sm_count() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

extern "C" int kp_spatial_softmax_fwd(kp_stream stream, const float* heat, int planes, int h, int w, float* k,
                                      float* p_h, float* p_w) {
    KP_CHECK_ARG(heat && k && planes > 0 && h > 0 && w > 0 && (h + w) * 4 <= 48 * 1024,
                 "kp_spatial_softmax_fwd: bad arguments");
    ssm_fwd_k<<<planes, 256, (h + w) * sizeof(float), (cudaStream_t)stream>>>(heat, h, w, k, p_h, p_w);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

extern "C" int kp_spatial_softmax_bwd(kp_stream stream, const float* dk, const float* k, const float* p_h,
                                      const float* p_w, int planes, int h, int w, float* dheat) {
    KP_CHECK_ARG(dk && k && p_h && p_w && dheat && planes > 0 && h > 0 && w > 0 && (h + w) * 4 <= 48 * 1024,
                 "kp_spatial_softmax_bwd: bad arguments");
    ssm_bwd_k<<<planes, 256, (h + w) * sizeof(float), (cudaStream_t)stream>>>(dk, k, p_h, p_w, h, w, dheat);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

extern "C" int kp_gaussian_fwd(kp_stream stream, const float* k, int planes, int h, int w, float sigma, float eps,
                               float* m) {
    KP_CHECK_ARG(k && m && planes > 0 && h > 0 && w > 0 && sigma > 0, "kp_gaussian_fwd: bad arguments");
    gaussian_fwd_k<<<planes, 256, 0, (cudaStream_t)stream>>>(k, h, w, (float)(2.0 * (double)sigma * (double)sigma), eps, m);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

extern "C" int kp_gaussian_bwd(kp_stream stream, const kp_view* dm, int pad, const float* k, const int32_t* argmax,
                               int N, int K, int h, int w, float sigma, float eps, float* dk) {
    KP_CHECK_ARG(dm && dm->ptr && k && dk && N > 0 && K > 0 && h > 0 && w > 0 && sigma > 0, "kp_gaussian_bwd: bad arguments");
    const float two_s2 = (float)(2.0 * (double)sigma * (double)sigma);
    return dispatch1(dm->dtype, [&](auto tg) -> int {
        using TG = decltype(tg);
        gaussian_bwd_k<TG><<<N * K, 256, 0, (cudaStream_t)stream>>>(make_view<TG>(dm), pad, k, argmax, K, h, w, two_s2,
                                                                   eps, dk);
        KP_LAUNCH_CHECK();
        return KP_OK;
    });
}

extern "C" int kp_transport_fwd(kp_stream stream, const kp_view* phi_s, const kp_view* phi_t, const float* k_s,
                                const float* k_t, const kp_view* out, int pad, float* mask_s, float* mask_t,
                                int32_t* argmax_t, int N, int h, int w, int C, int K, float sigma, float eps) {
    KP_CHECK_ARG(phi_s && phi_t && out && phi_s->ptr && phi_t->ptr && out->ptr && k_s && k_t && N > 0 && h > 0 &&
                     w > 0 && C > 0 && K > 0 && phi_s->dtype == phi_t->dtype,
                 "kp_transport_fwd: bad arguments");
    const float two_s2 = (float)(2.0 * (double)sigma * (double)sigma);
    long long total = (long long)N * (h + 2 * pad) * (w + 2 * pad) * C;
    int grid = grid_for(total, 256);
    const int G = transport_group(C);
    const long long pix = (long long)N * (h + 2 * pad) * (w + 2 * pad);
    if (G && pix < (1LL << 31) && view_vec8_ok(phi_s, C) && view_vec8_ok(phi_t, C) && view_vec8_ok(out, C) &&
        ((uintptr_t)k_s % 8) == 0 && ((uintptr_t)k_t % 8) == 0) {
        const int vgrid = grid_for(pix * G, 256);
        return dispatch1(phi_s->dtype, [&](auto tp) -> int {
            return dispatch1(out->dtype, [&](auto to) -> int {
                using TP = decltype(tp);
                using TO = decltype(to);
#define KP_TF(GV)                                                                                                          \
    transport_fwd_vec_k<TP, TO, GV><<<vgrid, 256, 0, (cudaStream_t)stream>>>(                                              \
        make_view<TP>(phi_s), make_view<TP>(phi_t), k_s, k_t, make_view<TO>(out), pad, mask_s, mask_t, argmax_t, N, h, w, K, \
        two_s2, eps)
                switch (G) {
                    case 1: KP_TF(1); break;
                    case 2: KP_TF(2); break;
                    case 4: KP_TF(4); break;
                    case 8: KP_TF(8); break;
                    case 16: KP_TF(16); break;
                    default: KP_TF(32); break;
                }
#undef KP_TF
                KP_LAUNCH_CHECK();
                return KP_OK;
            });
        });
    }
    return dispatch1(phi_s->dtype, [&](auto tp) -> int {
        return dispatch1(out->dtype, [&](auto to) -> int {
            using TP = decltype(tp);
            using TO = decltype(to);
            transport_fwd_k<TP, TO><<<grid, 256, 0, (cudaStream_t)stream>>>(
                make_view<TP>(phi_s), make_view<TP>(phi_t), k_s, k_t, make_view<TO>(out), pad, mask_s, mask_t, argmax_t,
                N, h, w, C, K, two_s2, eps);
            KP_LAUNCH_CHECK();
            return KP_OK;
        });
    });
}

extern "C" int kp_transport_bwd(kp_stream stream, const kp_view* dout, int pad, const kp_view* phi_s,
                                const kp_view* phi_t, const float* mask_s, const float* mask_t, const kp_view* dphi_t,
                                float* dmask_t, int N, int h, int w, int C) {
    KP_CHECK_ARG(dout && phi_s && phi_t && dphi_t && dout->ptr && phi_s->ptr && phi_t->ptr && dphi_t->ptr && mask_s &&
                     mask_t && dmask_t && N > 0 && h > 0 && w > 0 && C > 0 && phi_s->dtype == phi_t->dtype,
                 "kp_transport_bwd: bad arguments");
    long long P = (long long)N * h * w;
    int grid = grid_for(P * 32, 256);
    const int G = transport_group(C);
    if (G && P < (1LL << 31) && view_vec8_ok(dout, C) && view_vec8_ok(phi_s, C) && view_vec8_ok(phi_t, C) &&
        view_vec8_ok(dphi_t, C)) {
        const int vgrid = grid_for(P * G, 256);
        return dispatch1(dout->dtype, [&](auto tg) -> int {
            return dispatch1(phi_s->dtype, [&](auto tp) -> int {
                return dispatch1(dphi_t->dtype, [&](auto td) -> int {
                    using TG = decltype(tg);
                    using TP = decltype(tp);
                    using TD = decltype(td);
#define KP_TB(GV)                                                                                                          \
    transport_bwd_vec_k<TG, TP, TD, GV><<<vgrid, 256, 0, (cudaStream_t)stream>>>(                                          \
        make_view<TG>(dout), pad, make_view<TP>(phi_s), make_view<TP>(phi_t), mask_s, mask_t, make_view<TD>(dphi_t), dmask_t, \
        N, h, w)
                    switch (G) {
                        case 1: KP_TB(1); break;
                        case 2: KP_TB(2); break;
                        case 4: KP_TB(4); break;
                        case 8: KP_TB(8); break;
                        case 16: KP_TB(16); break;
                        default: KP_TB(32); break;
                    }
#undef KP_TB
                    KP_LAUNCH_CHECK();
                    return KP_OK;
                });
            });
        });
    }
    return dispatch1(dout->dtype, [&](auto tg) -> int {
        return dispatch1(phi_s->dtype, [&](auto tp) -> int {
            return dispatch1(dphi_t->dtype, [&](auto td) -> int {
                using TG = decltype(tg);
                using TP = decltype(tp);
                using TD = decltype(td);
                transport_bwd_k<TG, TP, TD><<<grid, 256, 0, (cudaStream_t)stream>>>(
                    make_view<TG>(dout), pad, make_view<TP>(phi_s), make_view<TP>(phi_t), mask_s, mask_t,
                    make_view<TD>(dphi_t), dmask_t, N, h, w, C);
                KP_LAUNCH_CHECK();
                return KP_OK;
            });
        });
    });
}

extern "C" int kp_transport_mode_fwd(kp_stream stream, int mode, const kp_view* phi_s, const kp_view* phi_t,
                                     const float* k_s, const float* k_t, const kp_view* out, int pad, float* mask_s,
                                     float* mask_t, int32_t* aux, float* coef, int N, int h, int w, int C, int K,
                                     float sigma, float eps) {
    KP_CHECK_ARG(phi_s && phi_t && out && phi_s->ptr && phi_t->ptr && out->ptr && k_s && k_t && N > 0 && h > 0 &&
                     w > 0 && C > 0 && K > 0 && phi_s->dtype == phi_t->dtype && mode >= KP_COMBINE_MAX &&
                     mode <= KP_COMBINE_LOOP && (mode != KP_COMBINE_LOOP || coef) && (mode == KP_COMBINE_LOOP || aux),
                 "kp_transport_mode_fwd: bad arguments");
    const float two_s2 = (float)(2.0 * (double)sigma * (double)sigma);
    long long total = (long long)N * (h + 2 * pad) * (w + 2 * pad) * C;
    int grid = grid_for(total, 256);
    return dispatch1(phi_s->dtype, [&](auto tp) -> int {
        return dispatch1(out->dtype, [&](auto to) -> int {
            using TP = decltype(tp);
            using TO = decltype(to);
#define KP_TMF(M) transport_mode_fwd_k<M, TP, TO><<<grid, 256, 0, (cudaStream_t)stream>>>(make_view<TP>(phi_s), make_view<TP>(phi_t), k_s, k_t, make_view<TO>(out), pad, mask_s, mask_t, aux, coef, N, h, w, C, K, two_s2, eps)
            if (mode == KP_COMBINE_MAX) KP_TMF(KP_COMBINE_MAX);
            else if (mode == KP_COMBINE_SUM) KP_TMF(KP_COMBINE_SUM);
            else KP_TMF(KP_COMBINE_LOOP);
#undef KP_TMF
            KP_LAUNCH_CHECK();
            return KP_OK;
        });
    });
}

extern "C" int kp_transport_mode_bwd(kp_stream stream, int mode, const kp_view* dout, int pad, const kp_view* phi_s,
                                     const kp_view* phi_t, const float* k_s, const float* k_t, const float* mask_s,
                                     const float* mask_t, const int32_t* aux, const float* coef, const kp_view* dphi_t,
                                     float* dm_t, int N, int h, int w, int C, int K, float sigma, float eps) {
    KP_CHECK_ARG(dout && phi_s && phi_t && dphi_t && dout->ptr && phi_s->ptr && phi_t->ptr && dphi_t->ptr && mask_s &&
                     mask_t && dm_t && k_s && k_t && N > 0 && h > 0 && w > 0 && C > 0 && K > 0 &&
                     phi_s->dtype == phi_t->dtype && mode >= KP_COMBINE_MAX && mode <= KP_COMBINE_LOOP &&
                     (mode != KP_COMBINE_LOOP || coef) && (mode == KP_COMBINE_LOOP || aux),
                 "kp_transport_mode_bwd: bad arguments");
    const float two_s2 = (float)(2.0 * (double)sigma * (double)sigma);
    long long P = (long long)N * h * w;
    int grid = grid_for(P * 32, 256);
    return dispatch1(dout->dtype, [&](auto tg) -> int {
        return dispatch1(phi_s->dtype, [&](auto tp) -> int {
            return dispatch1(dphi_t->dtype, [&](auto td) -> int {
                using TG = decltype(tg);
                using TP = decltype(tp);
                using TD = decltype(td);
#define KP_TMB(M) transport_mode_bwd_k<M, TG, TP, TD><<<grid, 256, 0, (cudaStream_t)stream>>>(make_view<TG>(dout), pad, make_view<TP>(phi_s), make_view<TP>(phi_t), k_s, k_t, mask_s, mask_t, aux, coef, make_view<TD>(dphi_t), dm_t, N, h, w, C, K, two_s2, eps)
                if (mode == KP_COMBINE_MAX) KP_TMB(KP_COMBINE_MAX);
                else if (mode == KP_COMBINE_SUM) KP_TMB(KP_COMBINE_SUM);
                else KP_TMB(KP_COMBINE_LOOP);
#undef KP_TMB
                KP_LAUNCH_CHECK();
                return KP_OK;
            });
        });
    });
}

extern "C" int kp_l2_loss(kp_stream stream, const float* xhat, const float* target, const float* mask, int64_t numel,
                          float gscale, double* loss_sum, float* dxhat) {
    KP_CHECK_ARG(xhat && target && numel > 0, "kp_l2_loss: bad arguments");
    l2_loss_k<<<grid_for(numel, 256 * 4), 256, 0, (cudaStream_t)stream>>>(xhat, target, mask, numel, gscale, loss_sum,
                                                                         dxhat);
    KP_LAUNCH_CHECK();
    return KP_OK;
}
