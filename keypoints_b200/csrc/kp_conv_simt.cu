// Direct (implicit-GEMM) convolution on the CUDA cores in fp32: the parity-grade path, and the
// path for thin layers (Cin in {1,3}, Cout in {1,3,K}) where a tensor-core tile would be empty.
// 64x64 output tile per block, K chunk of 16, 4x4 register tile per thread.
#include "kp_common.cuh"

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

}  // namespace

extern "C" int kp_conv_simt(kp_stream stream, const kp_view* in, const float* wk, const float* bias,
                            const kp_view* out, double* stats, int N, int OH, int OW, int IH, int IW, int Cin,
                            int Cout, int ks, int off) {
    KP_CHECK_ARG(in && out && in->ptr && out->ptr && wk && N > 0 && OH > 0 && OW > 0 && Cin > 0 && Cout > 0 &&
                     (ks == 1 || ks == 3),
                 "kp_conv_simt: bad arguments");
    long long M = (long long)N * OH * OW;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Cout + BN - 1) / BN), 1);
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
