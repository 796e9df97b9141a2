// Direct (implicit-GEMM) convolution on the CUDA cores in fp32: the parity-grade path, and the
// path for thin layers (Cin in {1,3}, Cout in {1,3,K}) where a tensor-core tile would be empty.
// 64x64 output tile per block, K chunk of 16, 4x4 register tile per thread.
#include "kp_common.cuh"
#include <stdlib.h>

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename F>
int dispatch1(int dt, F&& f) {
    if (dt == KP_F32) return f(float{});
    if (dt == KP_BF16) return f(bf16{});
    kp_set_error("bad dtype %d", dt);
    return KP_ERR_ARG;
}

// out[m][co] = bias[co] + sum_kk A[m][kk] * wk[kk][co],  kk = t*Cin + ci, A gathered with zero fill
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
conv_simt_k(View<TI> in, const float* __restrict__ wk, const float* __restrict__ bias, View<TO> out, double* stats,
            int N, int OH, int OW, int IH, int IW, int Cin, int Cout, int ks, int off) {
    __shared__ float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ float red[2][16][BN];
    const int tid = threadIdx.x;
    const long long M = (long long)N * OH * OW;
    const int KK = ks * ks * Cin;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-load role: pixel am, 4 consecutive kk starting at ak
    const int am = tid >> 2, ak = (tid & 3) * 4;
    long long mg = m0 + am;
    const bool mvalid = mg < M;
    int an = 0, aoy = 0, aox = 0;
    if (mvalid) {
        aox = (int)(mg % OW);
        long long r = mg / OW;
        aoy = (int)(r % OH);
        an = (int)(r / OH);
    }
    // B-load role: row bk, 4 consecutive co starting at bc
    const int bk = tid >> 4, bc = (tid & 15) * 4;

    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < KK; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int kk = k0 + ak + i;
            float v = 0.f;
            if (mvalid && kk < KK) {
                int t = kk / Cin, ci = kk - t * Cin;
                int iy = aoy + t / ks + off, ix = aox + t % ks + off;
                if (iy >= 0 && iy < IH && ix >= 0 && ix < IW) v = to_f(*in.at(an, iy, ix, ci));
            }
            As[ak + i][am] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int kk = k0 + bk, co = n0 + bc + i;
            Bs[bk][bc + i] = (kk < KK && co < Cout) ? wk[(long long)kk * Cout + co] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
            float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
        int ox = (int)(m % OW);
        long long r = m / OW;
        int oy = (int)(r % OH);
        int n = (int)(r / OH);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co >= Cout) continue;
            float v = acc[i][j] + (bias ? bias[co] : 0.f);
            from_f(out.at(n, oy, ox, co), v);
            s1[j] += v;
            s2[j] += v * v;
        }
    }
    if (stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { red[0][ty][tx * 4 + j] = s1[j]; red[1][ty][tx * 4 + j] = s2[j]; }
        __syncthreads();
        if (tid < BN && n0 + tid < Cout) {
            float a = 0.f, b = 0.f;
            for (int r = 0; r < 16; ++r) { a += red[0][r][tid]; b += red[1][r][tid]; }
            atomicAdd(&stats[n0 + tid], (double)a);
            atomicAdd(&stats[Cout + n0 + tid], (double)b);
        }
    }
}

// dw[co][ci][t] += sum_p X[p + t][ci] * dY[p][co]; rows r = t*Cin + ci, cols co, reduction over pixels
template <typename TX, typename TG>
__global__ void __launch_bounds__(256)
conv_wgrad_simt_k(View<TX> x, View<TG> dy, float* dw, int N, int H, int W, int Cin, int Cout, int ks,
                  long long chunk) {
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    const int tid = threadIdx.x;
    const long long P = (long long)N * H * W;
    const int KK = ks * ks * Cin;
    const int r0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const long long pbeg = (long long)blockIdx.z * chunk;
    const long long pend = min(P, pbeg + chunk);
    const int lp = tid >> 4, l4 = (tid & 15) * 4;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long p0 = pbeg; p0 < pend; p0 += BK) {
        long long p = p0 + lp;
        bool pv = p < pend;
        int n = 0, yy = 0, xx = 0;
        if (pv) {
            xx = (int)(p % W);
            long long r = p / W;
            yy = (int)(r % H);
            n = (int)(r / H);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int rr = r0 + l4 + i;
            float v = 0.f;
            if (pv && rr < KK) {
                int t = rr / Cin, ci = rr - t * Cin;
                v = to_f(*x.at(n, yy + t / ks, xx + t % ks, ci));
            }
            As[lp][l4 + i] = v;
            int co = n0 + l4 + i;
            Bs[lp][l4 + i] = (pv && co < Cout) ? to_f(*dy.at(n, yy, xx, co)) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int kk2 = ks * ks;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int rr = r0 + ty * 4 + i;
        if (rr >= KK) continue;
        int t = rr / Cin, ci = rr - t * Cin;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co >= Cout) continue;
            atomicAdd(&dw[((long long)co * Cin + ci) * kk2 + t], acc[i][j]);
        }
    }
}

// OIHW fp32 -> kernel layouts
__global__ void pack_weights_k(const float* __restrict__ w, int Cout, int Cin, int ks, int ci_pad, float* simt_f,
                               float* simt_d, bf16* tc_f, bf16* tc_d) {
    const int T = ks * ks;
    const long long total = (long long)T * Cout * ci_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int ci = (int)(i % ci_pad);
        long long r = i / ci_pad;
        int co = (int)(r % Cout);
        int t = (int)(r / Cout);
        float v = ci < Cin ? w[((long long)co * Cin + ci) * T + t] : 0.f;
        int tf = T - 1 - t;   // flipped tap
        if (ci < Cin) {
            if (simt_f) simt_f[((long long)t * Cin + ci) * Cout + co] = v;
            if (simt_d) simt_d[((long long)tf * Cout + co) * Cin + ci] = v;
        }
        if (tc_f) tc_f[((long long)t * Cout + co) * ci_pad + ci] = __float2bfloat16_rn(v);
        if (tc_d) tc_d[((long long)tf * ci_pad + ci) * Cout + co] = __float2bfloat16_rn(v);
    }
}

// ================================================================================================
// Thin-layer kernels (bf16 or fp32 activations, fp32 math on the CUDA cores).  The first conv of a stack
// (Cin <= 4) and the 1x1 heads with a handful of outputs are HBM-bound: a 128x128x64 tile machine wastes most
// of its lanes on them.  Here a warp walks pixels; each lane owns two channels of the WIDE side (64 channels per
// warp pass), the <= 36 values of the THIN side of a pixel are staged in shared memory and broadcast.
// ================================================================================================
constexpr int THIN_MAX = 36;
constexpr int TILE_PX = 256;

template <typename T>
__device__ __forceinline__ void load2(const T* p, float& a, float& b);
template <> __device__ __forceinline__ void load2<float>(const float* p, float& a, float& b) {
    float2 v = *reinterpret_cast<const float2*>(p); a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void load2<bf16>(const bf16* p, float& a, float& b) {
    float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); a = v.x; b = v.y;
}
__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(bf16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// stage the thin values of TILE_PX pixels: thin[p][kk], kk = t*Cthin + c, gathered at (y + t/ks + off, x + t%ks + off).
// The per-kk element offsets / tap coordinates are tabulated once per block (sm_koff, sm_kyx) so the gather loop has
// no integer divisions.
__device__ __forceinline__ void build_koff(long long* sm_koff, int* sm_kyx, long long sy, long long sx, long long sc,
                                           int Cthin, int ks, int off, int KK, int KKP) {
    for (int kk = threadIdx.x; kk < KKP; kk += blockDim.x) {
        int t = kk / Cthin, c = kk - t * Cthin;
        int dy = t / ks + off, dx = t % ks + off;
        sm_koff[kk] = kk < KK ? dy * sy + dx * sx + c * sc : 0;
        sm_kyx[kk] = kk < KK ? ((dy & 0xffff) << 16) | (dx & 0xffff) : 0x7fff7fff;
    }
}

template <typename TT>
__device__ __forceinline__ void stage_thin(const View<TT>& thin, float* sm_thin, int* sm_xy, const long long* sm_koff,
                                           const int* sm_kyx, long long p0, long long M, int OH, int OW, int IH, int IW,
                                           int KK, int KKP, bool inbounds) {
    const int tid = threadIdx.x;
    long long p = p0 + tid;
    int n = 0, y = 0, x = 0;
    const bool ok = p < M;
    if (ok) {
        x = (int)(p % OW);
        long long r = p / OW;
        y = (int)(r % OH);
        n = (int)(r / OH);
    }
    sm_xy[tid * 3 + 0] = ok ? n : -1;
    sm_xy[tid * 3 + 1] = y;
    sm_xy[tid * 3 + 2] = x;
    float* dst = sm_thin + tid * KKP;
    const TT* base = thin.at(n, y, x, 0);
    if (ok && inbounds) {
        for (int kk = 0; kk < KKP; ++kk) dst[kk] = kk < KK ? to_f(base[sm_koff[kk]]) : 0.f;
    } else {
        for (int kk = 0; kk < KKP; ++kk) {
            float v = 0.f;
            if (ok && kk < KK) {
                const int pk = sm_kyx[kk];
                const int iy = y + (int)(short)(pk >> 16), ix = x + (int)(short)(pk & 0xffff);
                if (iy >= 0 && iy < IH && ix >= 0 && ix < IW) v = to_f(base[sm_koff[kk]]);
            }
            dst[kk] = v;
        }
    }
}

// out[pix][wide] = bias[wide] + sum_kk thin[pix][kk] * wk[kk][wide]      (wide = 64 channels per blockIdx.y)
template <typename TT, typename TO, int KKP>
__global__ void __launch_bounds__(256)
thin_fprop_k(View<TT> thin, const float* __restrict__ wk, const float* __restrict__ bias, View<TO> out, double* stats,
             int N, int OH, int OW, int IH, int IW, int Cthin, int Cwide, int ks, int off) {
    __shared__ __align__(16) float sm_thin[TILE_PX * KKP];
    __shared__ int sm_xy[TILE_PX * 3];
    __shared__ float sm_red[8][2][64];
    __shared__ long long sm_koff[KKP];
    __shared__ int sm_kyx[KKP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int KK = ks * ks * Cthin;
    const int c0 = blockIdx.y * 64 + lane * 2;
    const long long M = (long long)N * OH * OW;
    float w0[KKP], w1[KKP];
#pragma unroll
    for (int kk = 0; kk < KKP; ++kk) {
        w0[kk] = kk < KK ? wk[(long long)kk * Cwide + c0] : 0.f;
        w1[kk] = kk < KK ? wk[(long long)kk * Cwide + c0 + 1] : 0.f;
    }
    const float b0 = bias ? bias[c0] : 0.f, b1 = bias ? bias[c0 + 1] : 0.f;
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
    const long long ntiles = (M + TILE_PX - 1) / TILE_PX;
    build_koff(sm_koff, sm_kyx, thin.sy, thin.sx, thin.sc, Cthin, ks, off, KK, KKP);
    const bool inb = off >= 0 && OH + off + ks - 1 <= IH && OW + off + ks - 1 <= IW;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        stage_thin<TT>(thin, sm_thin, sm_xy, sm_koff, sm_kyx, tile * TILE_PX, M, OH, OW, IH, IW, KK, KKP, inb);
        __syncthreads();
        for (int i = 0; i < 32; ++i) {
            const int pp = warp * 32 + i;
            const int n = sm_xy[pp * 3];
            if (n < 0) break;
            const float4* pv = reinterpret_cast<const float4*>(sm_thin + pp * KKP);
            float a0 = b0, a1 = b1;
#pragma unroll
            for (int k4 = 0; k4 < KKP / 4; ++k4) {
                float4 v = pv[k4];
                a0 = fmaf(v.x, w0[4 * k4], a0); a1 = fmaf(v.x, w1[4 * k4], a1);
                a0 = fmaf(v.y, w0[4 * k4 + 1], a0); a1 = fmaf(v.y, w1[4 * k4 + 1], a1);
                a0 = fmaf(v.z, w0[4 * k4 + 2], a0); a1 = fmaf(v.z, w1[4 * k4 + 2], a1);
                a0 = fmaf(v.w, w0[4 * k4 + 3], a0); a1 = fmaf(v.w, w1[4 * k4 + 3], a1);
            }
            store2(out.at(n, sm_xy[pp * 3 + 1], sm_xy[pp * 3 + 2], c0), a0, a1);
            s1a += a0; s1b += a1; s2a = fmaf(a0, a0, s2a); s2b = fmaf(a1, a1, s2b);
        }
    }
    if (stats) {
        sm_red[warp][0][lane * 2] = s1a; sm_red[warp][0][lane * 2 + 1] = s1b;
        sm_red[warp][1][lane * 2] = s2a; sm_red[warp][1][lane * 2 + 1] = s2b;
        __syncthreads();
        if (threadIdx.x < 128) {
            const int q = threadIdx.x >> 6, ch = threadIdx.x & 63;
            float a = 0.f;
            for (int w = 0; w < 8; ++w) a += sm_red[w][q][ch];
            atomicAdd(&stats[q * Cwide + blockIdx.y * 64 + ch], (double)a);
        }
    }
}

// G[kk][wide] = sum_pix thin[pix][kk] * wide[pix][wide]; written (atomically) at dw[wide*sa + (kk%cdiv)*sb + (kk/cdiv)*sc]
template <typename TT, typename TW, int KKP>
__global__ void __launch_bounds__(256)
thin_wgrad_k(View<TT> thin, View<TW> wide, float* dw, int N, int OH, int OW, int IH, int IW, int Cthin, int ks, int off,
             int cdiv, long long sa, long long sb, long long sc) {
    __shared__ __align__(16) float sm_thin[TILE_PX * KKP];
    __shared__ int sm_xy[TILE_PX * 3];
    __shared__ long long sm_koff[KKP];
    __shared__ int sm_kyx[KKP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int KK = ks * ks * Cthin;
    const int c0 = blockIdx.y * 64 + lane * 2;
    const long long M = (long long)N * OH * OW;
    build_koff(sm_koff, sm_kyx, thin.sy, thin.sx, thin.sc, Cthin, ks, off, KK, KKP);
    const bool inb = off >= 0 && OH + off + ks - 1 <= IH && OW + off + ks - 1 <= IW;
    float g0[KKP], g1[KKP];
#pragma unroll
    for (int kk = 0; kk < KKP; ++kk) { g0[kk] = 0.f; g1[kk] = 0.f; }
    const long long ntiles = (M + TILE_PX - 1) / TILE_PX;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        stage_thin<TT>(thin, sm_thin, sm_xy, sm_koff, sm_kyx, tile * TILE_PX, M, OH, OW, IH, IW, KK, KKP, inb);
        __syncthreads();
        for (int i = 0; i < 32; ++i) {
            const int pp = warp * 32 + i;
            const int n = sm_xy[pp * 3];
            if (n < 0) break;
            float d0, d1;
            load2<TW>(wide.at(n, sm_xy[pp * 3 + 1], sm_xy[pp * 3 + 2], c0), d0, d1);
            const float4* pv = reinterpret_cast<const float4*>(sm_thin + pp * KKP);
#pragma unroll
            for (int k4 = 0; k4 < KKP / 4; ++k4) {
                float4 v = pv[k4];
                g0[4 * k4] = fmaf(v.x, d0, g0[4 * k4]); g1[4 * k4] = fmaf(v.x, d1, g1[4 * k4]);
                g0[4 * k4 + 1] = fmaf(v.y, d0, g0[4 * k4 + 1]); g1[4 * k4 + 1] = fmaf(v.y, d1, g1[4 * k4 + 1]);
                g0[4 * k4 + 2] = fmaf(v.z, d0, g0[4 * k4 + 2]); g1[4 * k4 + 2] = fmaf(v.z, d1, g1[4 * k4 + 2]);
                g0[4 * k4 + 3] = fmaf(v.w, d0, g0[4 * k4 + 3]); g1[4 * k4 + 3] = fmaf(v.w, d1, g1[4 * k4 + 3]);
            }
        }
    }
    // reduce the 8 warps of the block through shared memory (reusing the staging buffer), one atomic per output
    __syncthreads();
    float* red = sm_thin;                      // [KKP][64] floats <= 36*64*4 = 9 KB < staging size
    for (int i = threadIdx.x; i < KKP * 64; i += 256) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KKP; ++kk) {
        atomicAdd(&red[kk * 64 + lane * 2], g0[kk]);
        atomicAdd(&red[kk * 64 + lane * 2 + 1], g1[kk]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < KK * 64; i += 256) {
        const int kk = i >> 6, wc = blockIdx.y * 64 + (i & 63);
        atomicAdd(&dw[wc * sa + (kk % cdiv) * sb + (kk / cdiv) * sc], red[i]);
    }
}

// 1x1 conv with <= 16 outputs: one thread per pixel, 8 input channels per 16-byte load
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
thin_out_fprop_k(View<TI> in, const float* __restrict__ wk /*[Cin][Cout]*/, const float* __restrict__ bias, View<TO> out,
                 int N, int OH, int OW, int Cin, int Cout) {
    extern __shared__ float sm_w[];            // [Cin][16]
    for (int i = threadIdx.x; i < Cin * 16; i += blockDim.x) {
        int ci = i >> 4, co = i & 15;
        sm_w[i] = co < Cout ? wk[(long long)ci * Cout + co] : 0.f;
    }
    __syncthreads();
    const long long M = (long long)N * OH * OW;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < M; p += (long long)gridDim.x * blockDim.x) {
        int x = (int)(p % OW);
        long long r = p / OW;
        int y = (int)(r % OH);
        int n = (int)(r / OH);
        float acc[16];
#pragma unroll
        for (int co = 0; co < 16; ++co) acc[co] = (bias && co < Cout) ? bias[co] : 0.f;
        const TI* src = in.at(n, y, x, 0);
        for (int c8 = 0; c8 < Cin; c8 += 8) {
            float v[8];
            Vec<TI, 8>::load(src + c8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4* wr = reinterpret_cast<const float4*>(sm_w + (c8 + j) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q * 4 < Cout) {
                        float4 wv = wr[q];
                        acc[q * 4] = fmaf(v[j], wv.x, acc[q * 4]);
                        acc[q * 4 + 1] = fmaf(v[j], wv.y, acc[q * 4 + 1]);
                        acc[q * 4 + 2] = fmaf(v[j], wv.z, acc[q * 4 + 2]);
                        acc[q * 4 + 3] = fmaf(v[j], wv.w, acc[q * 4 + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int co = 0; co < 16; ++co)
            if (co < Cout) from_f(out.at(n, y, x, co), acc[co]);
    }
}

static bool thin_fast_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = kp_env("KP_SIMT_FAST");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
static bool thin_mma_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = kp_env("KP_THIN_MMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
static int thin_grid(long long M) {
    long long tiles = (M + TILE_PX - 1) / TILE_PX;
    long long cap = (long long)kp_sm_count() * 4;
    return (int)(tiles < cap ? (tiles < 1 ? 1 : tiles) : cap);
}

// multi-tensor variant: one launch repacks every layer (blockIdx.y = layer), descriptors live in device memory
struct PackDesc {
    const float* w;
    float* simt_f;
    float* simt_d;
    bf16* tc_f;
    bf16* tc_d;
    long long cout, cin, ks, ci_pad;
};
// One block repacks 32 (co) x 32 (ci) x T tiles through shared memory so that the OIHW master weights are read in
// contiguous runs of 32*T floats and every destination layout is written in contiguous 32-element segments.
__global__ void __launch_bounds__(256)
pack_weights_multi_k(const PackDesc* __restrict__ table) {
    __shared__ float sm[32][32 * 9 + 1];
    const PackDesc d = table[blockIdx.y];
    const int T = (int)(d.ks * d.ks), Cout = (int)d.cout, Cin = (int)d.cin, ci_pad = (int)d.ci_pad;
    const int tiles_ci = (ci_pad + 31) / 32, tiles = ((Cout + 31) / 32) * tiles_ci;
    const int row = 32 * T;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int co0 = (tile / tiles_ci) * 32, ci0 = (tile % tiles_ci) * 32;
#pragma unroll 12
        for (int idx = threadIdx.x; idx < 32 * row; idx += 256) {
            const int r = idx / row, j = idx - r * row;
            const int co = co0 + r, ci = ci0 + j / T;
            sm[r][j] = (co < Cout && ci < Cin) ? d.w[((long long)co * Cin + ci0) * T + j] : 0.f;
        }
        __syncthreads();
        // bf16 layouts are written two elements (4 bytes) per thread; ci_pad and Cout of a tensor-core layer are multiples of 64
        if (d.tc_f) {
            for (int idx = threadIdx.x; idx < 512 * T; idx += 256) {    // ci fastest: [t][co][ci]
                const int c = (idx & 15) * 2, r = (idx >> 4) & 31, t = idx >> 9;
                const int co = co0 + r, ci = ci0 + c;
                if (co < Cout && ci < ci_pad)
                    *reinterpret_cast<__nv_bfloat162*>(d.tc_f + ((long long)t * Cout + co) * ci_pad + ci) =
                        __floats2bfloat162_rn(sm[r][c * T + t], sm[r][(c + 1) * T + t]);
            }
        }
        if (d.tc_d) {
            for (int idx = threadIdx.x; idx < 512 * T; idx += 256) {    // co fastest: [t][ci][co], taps flipped
                const int r = (idx & 15) * 2, c = (idx >> 4) & 31, t = idx >> 9;
                const int co = co0 + r, ci = ci0 + c, tf = T - 1 - t;
                if (co < Cout && ci < ci_pad)
                    *reinterpret_cast<__nv_bfloat162*>(d.tc_d + ((long long)tf * ci_pad + ci) * Cout + co) =
                        __floats2bfloat162_rn(sm[r][c * T + t], sm[r + 1][c * T + t]);
            }
        }
        if (d.simt_d) {
            for (int idx = threadIdx.x; idx < 1024 * T; idx += 256) {
                const int c = idx & 31, r = (idx >> 5) & 31, t = idx >> 10;
                const int co = co0 + r, ci = ci0 + c, tf = T - 1 - t;
                if (co < Cout && ci < Cin) d.simt_d[((long long)tf * Cout + co) * Cin + ci] = sm[r][c * T + t];
            }
        }
        if (d.simt_f) {
            for (int idx = threadIdx.x; idx < 1024 * T; idx += 256) {
                const int r = idx & 31, c = (idx >> 5) & 31, t = idx >> 10;
                const int co = co0 + r, ci = ci0 + c;
                if (co < Cout && ci < Cin) d.simt_f[((long long)t * Cin + ci) * Cout + co] = sm[r][c * T + t];
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int kp_pack_weights_multi(kp_stream stream, const void* table_dev, int n_layers) {
    KP_CHECK_ARG(table_dev && n_layers > 0 && n_layers <= 65535, "kp_pack_weights_multi: bad arguments");
    dim3 grid(256u, (unsigned)n_layers, 1);
    pack_weights_multi_k<<<grid, 256, 0, (cudaStream_t)stream>>>((const PackDesc*)table_dev);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

extern "C" int kp_conv_simt(kp_stream stream, const kp_view* in, const float* wk, const float* bias,
                            const kp_view* out, double* stats, int N, int OH, int OW, int IH, int IW, int Cin,
                            int Cout, int ks, int off) {
    KP_CHECK_ARG(in && out && in->ptr && out->ptr && wk && N > 0 && OH > 0 && OW > 0 && Cin > 0 && Cout > 0 &&
                     (ks == 1 || ks == 3),
                 "kp_conv_simt: bad arguments");
    long long M = (long long)N * OH * OW;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Cout + BN - 1) / BN), 1);
    const int KKt = ks * ks * Cin;
    const bool out2 = out->sc == 1 && (((uintptr_t)out->ptr) % 8) == 0 && out->sx % 2 == 0 && out->sy % 2 == 0 && out->sn % 2 == 0;
    if (thin_mma_enabled() && ks == 1 && !stats && OH == IH && OW == IW && kp_head_mma_fprop_ok(in, out, N, OH, OW, Cin, Cout))
        return kp_head_mma_fprop((cudaStream_t)stream, in, wk, bias, out, N, OH, OW, Cin, Cout);
    if (thin_mma_enabled() && ks == 1 && !stats && OH == IH && OW == IW && M < (1LL << 31)) {
        // 1x1 head: forward (wide -> <= 4) or its data gradient (<= 4 -> wide); wk is [Cin][Cout] in both cases
        if (kp_head1x1_ok(in, out, Cin, Cout))
            return kp_head1x1_fprop((cudaStream_t)stream, in, wk, bias, out, N, OH, OW, Cin, Cout);
        if (!bias && out->dtype == KP_BF16 && in->dtype == KP_BF16 && kp_head1x1_ok(out, in, Cout, Cin))
            return kp_head1x1_dgrad((cudaStream_t)stream, in, wk, out, N, OH, OW, Cout, Cin);
    }
    if (thin_mma_enabled() && kp_c1_conv_ok(in, out, OH, OW, IH, IW, Cin, Cout, ks, off))
        return kp_c1_fprop((cudaStream_t)stream, in, wk, bias, out, stats, N, OH, OW, Cout);
    if (thin_mma_enabled() && kp_small_mma_conv_ok(in, out, N, OH, OW, Cin, Cout, ks))
        return kp_small_mma_conv((cudaStream_t)stream, in, wk, bias, out, stats, N, OH, OW, IH, IW, Cin, Cout, ks, off);
    if (thin_mma_enabled() && out->dtype == KP_BF16 && kp_thin_mma_fprop_ok(in, out, OH, OW, IH, IW, Cin, Cout, ks, off))
        return kp_thin_mma_fprop((cudaStream_t)stream, in, wk, bias, out, stats, N, OH, OW, Cin, Cout);
    if (thin_fast_enabled() && KKt <= THIN_MAX && Cout % 64 == 0 && out2) {      // thin-K: first conv / dgrad of a thin head
        dim3 g2((unsigned)thin_grid(M), (unsigned)(Cout / 64), 1);
        return dispatch1(in->dtype, [&](auto ti) -> int {
            return dispatch1(out->dtype, [&](auto to) -> int {
                using TI = decltype(ti);
                using TO = decltype(to);
                cudaStream_t st = (cudaStream_t)stream;
#define KP_THIN(KKPV) thin_fprop_k<TI, TO, KKPV><<<g2, 256, 0, st>>>(make_view<TI>(in), wk, bias, make_view<TO>(out), stats, N, OH, OW, IH, IW, Cin, Cout, ks, off)
                if (KKt <= 4) KP_THIN(4);
                else if (KKt <= 12) KP_THIN(12);
                else if (KKt <= 28) KP_THIN(28);
                else KP_THIN(36);
#undef KP_THIN
                KP_LAUNCH_CHECK();
                return KP_OK;
            });
        });
    }
    if (thin_fast_enabled() && ks == 1 && Cout <= 16 && !stats && view_vec8_ok(in, Cin) && Cin * 16 * 4 <= 48 * 1024) {
        long long blocks = (M + 255) / 256;
        if (blocks > (long long)kp_sm_count() * 8) blocks = (long long)kp_sm_count() * 8;
        return dispatch1(in->dtype, [&](auto ti) -> int {
            return dispatch1(out->dtype, [&](auto to) -> int {
                using TI = decltype(ti);
                using TO = decltype(to);
                thin_out_fprop_k<TI, TO><<<(int)blocks, 256, Cin * 16 * sizeof(float), (cudaStream_t)stream>>>(
                    make_view<TI>(in), wk, bias, make_view<TO>(out), N, OH, OW, Cin, Cout);
                KP_LAUNCH_CHECK();
                return KP_OK;
            });
        });
    }
    return dispatch1(in->dtype, [&](auto ti) -> int {
        return dispatch1(out->dtype, [&](auto to) -> int {
            using TI = decltype(ti);
            using TO = decltype(to);
            conv_simt_k<TI, TO><<<grid, 256, 0, (cudaStream_t)stream>>>(make_view<TI>(in), wk, bias, make_view<TO>(out),
                                                                       stats, N, OH, OW, IH, IW, Cin, Cout, ks, off);
            KP_LAUNCH_CHECK();
            return KP_OK;
        });
    });
}

extern "C" int kp_conv_wgrad_simt(kp_stream stream, const kp_view* x, const kp_view* dy, float* dw_oihw, int N, int H,
                                  int W, int Cin, int Cout, int ks) {
    KP_CHECK_ARG(x && dy && x->ptr && dy->ptr && dw_oihw && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 &&
                     (ks == 1 || ks == 3),
                 "kp_conv_wgrad_simt: bad arguments");
    const int KK = ks * ks * Cin;
    const long long P = (long long)N * H * W;
    auto pair_ok = [](const kp_view* v) {
        return v->sc == 1 && (((uintptr_t)v->ptr) % 8) == 0 && v->sx % 2 == 0 && v->sy % 2 == 0 && v->sn % 2 == 0;
    };
    if (thin_mma_enabled() && ks == 1 && P < (1LL << 31) && dy->dtype == KP_BF16 && kp_head1x1_ok(x, dy, Cin, Cout))
        return kp_head1x1_wgrad((cudaStream_t)stream, x, dy, dw_oihw, N, H, W, Cin, Cout);
    if (thin_mma_enabled() && kp_c1_wgrad_ok(x, dy, Cin, Cout, ks))
        return kp_c1_wgrad((cudaStream_t)stream, x, dy, dw_oihw, N, H, W, Cout);
    if (thin_mma_enabled() && kp_small_mma_wgrad_ok(x, dy, N, H, W, Cin, Cout, ks))
        return kp_small_mma_wgrad((cudaStream_t)stream, x, dy, dw_oihw, N, H, W, Cin, Cout, ks);
    if (thin_mma_enabled() && dy->dtype == KP_BF16 && kp_thin_mma_wgrad_ok(x, dy, W, Cin, Cout, ks))
        return kp_thin_mma_wgrad((cudaStream_t)stream, x, dy, dw_oihw, N, H, W, Cin, Cout);
    if (thin_fast_enabled() && KK <= THIN_MAX && Cout % 64 == 0 && pair_ok(dy)) {
        // thin = x patches (kk = t*Cin + ci), wide = dy (co): dw[co][ci][t]
        dim3 g2((unsigned)thin_grid(P), (unsigned)(Cout / 64), 1);
        const int T = ks * ks;
        return dispatch1(x->dtype, [&](auto tx) -> int {
            return dispatch1(dy->dtype, [&](auto tg) -> int {
                using TX = decltype(tx);
                using TG = decltype(tg);
                cudaStream_t st = (cudaStream_t)stream;
#define KP_THINW(KKPV) thin_wgrad_k<TX, TG, KKPV><<<g2, 256, 0, st>>>(make_view<TX>(x), make_view<TG>(dy), dw_oihw, N, H, W, H + ks - 1, W + ks - 1, Cin, ks, 0, Cin, (long long)Cin * T, (long long)T, 1LL)
                if (KK <= 4) KP_THINW(4);
                else if (KK <= 12) KP_THINW(12);
                else if (KK <= 28) KP_THINW(28);
                else KP_THINW(36);
#undef KP_THINW
                KP_LAUNCH_CHECK();
                return KP_OK;
            });
        });
    }
    if (thin_fast_enabled() && ks == 1 && Cout <= 16 && Cin % 64 == 0 && pair_ok(x)) {
        // thin = dy (kk = co), wide = x (ci): dw[co][ci]
        dim3 g2((unsigned)thin_grid(P), (unsigned)(Cin / 64), 1);
        return dispatch1(dy->dtype, [&](auto tg) -> int {
            return dispatch1(x->dtype, [&](auto tx) -> int {
                using TX = decltype(tx);
                using TG = decltype(tg);
                cudaStream_t st = (cudaStream_t)stream;
#define KP_THINW(KKPV) thin_wgrad_k<TG, TX, KKPV><<<g2, 256, 0, st>>>(make_view<TG>(dy), make_view<TX>(x), dw_oihw, N, H, W, H, W, Cout, 1, 0, 1 << 30, 1LL, (long long)Cin, 0LL)
                if (Cout <= 4) KP_THINW(4);
                else if (Cout <= 12) KP_THINW(12);
                else KP_THINW(28);
#undef KP_THINW
                KP_LAUNCH_CHECK();
                return KP_OK;
            });
        });
    }
    int tiles = ((KK + BM - 1) / BM) * ((Cout + BN - 1) / BN);
    long long split = (4LL * kp_sm_count() + tiles - 1) / tiles;
    long long maxsplit = (P + 4 * BK - 1) / (4 * BK);
    if (split > maxsplit) split = maxsplit;
    if (split < 1) split = 1;
    if (split > 65535) split = 65535;
    long long chunk = (P + split - 1) / split;
    chunk = (chunk + BK - 1) / BK * BK;
    split = (P + chunk - 1) / chunk;
    dim3 grid((unsigned)((KK + BM - 1) / BM), (unsigned)((Cout + BN - 1) / BN), (unsigned)split);
    return dispatch1(x->dtype, [&](auto tx) -> int {
        return dispatch1(dy->dtype, [&](auto tg) -> int {
            using TX = decltype(tx);
            using TG = decltype(tg);
            conv_wgrad_simt_k<TX, TG><<<grid, 256, 0, (cudaStream_t)stream>>>(make_view<TX>(x), make_view<TG>(dy),
                                                                             dw_oihw, N, H, W, Cin, Cout, ks, chunk);
            KP_LAUNCH_CHECK();
            return KP_OK;
        });
    });
}

extern "C" int kp_pack_weights(kp_stream stream, const float* w_oihw, int Cout, int Cin, int ks, int ci_pad,
                               float* simt_f, float* simt_d, void* tc_f, void* tc_d) {
    KP_CHECK_ARG(w_oihw && Cout > 0 && Cin > 0 && (ks == 1 || ks == 3) && ci_pad >= Cin, "kp_pack_weights: bad arguments");
    long long total = (long long)ks * ks * Cout * ci_pad;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_k<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, ks, ci_pad, simt_f, simt_d, (bf16*)tc_f,
                                                            (bf16*)tc_d);
    KP_LAUNCH_CHECK();
    return KP_OK;
}
