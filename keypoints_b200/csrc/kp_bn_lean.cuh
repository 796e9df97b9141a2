// Lean row kernels for the BatchNorm / activation passes on dense NHWC buffers of ONE dtype (bf16 in throughput
// mode, fp32 in parity mode).  These passes are nominally HBM-bound, but at 16 resident warps per SM they only reach
// the memory roofline if a warp spends few issue slots per byte: the activation is a template parameter (no per-element
// branches), addresses are a 64-bit row base plus a 32-bit in-row offset, and 8 channels are handled as 4 float2 pairs.
// Included by kp_elementwise.cu (inside its anonymous namespace, after BnFuse / fused_affine).
#pragma once

template <typename T>
struct Rows {          // dense NHWC view with unit channel stride
    T* p;
    long long sn, sy;
    int sx;
    __device__ __forceinline__ T* row(int n, int y) const { return p + n * sn + y * sy; }
};
template <typename T>
static inline Rows<T> make_rows(const kp_view* v) {
    Rows<T> r;
    r.p = (T*)v->ptr; r.sn = v->sn; r.sy = v->sy; r.sx = (int)v->sx;
    return r;
}

template <typename T> struct P8;
template <> struct P8<bf16> {
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw ld(const bf16* p) { return *reinterpret_cast<const uint4*>(p); }
    static __device__ __forceinline__ float2 up1(uint32_t u) {
        return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    }
    static __device__ __forceinline__ void up(const Raw& r, float2 (&v)[4]) {
        v[0] = up1(r.x); v[1] = up1(r.y); v[2] = up1(r.z); v[3] = up1(r.w);
    }
    static __device__ __forceinline__ uint32_t pk(float2 f) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f.x, f.y);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ void st(bf16* p, const float2 (&v)[4]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(pk(v[0]), pk(v[1]), pk(v[2]), pk(v[3]));
    }
};
template <> struct P8<float> {
    struct Raw { float4 a, b; };
    static __device__ __forceinline__ Raw ld(const float* p) {
        Raw r; r.a = *reinterpret_cast<const float4*>(p); r.b = *reinterpret_cast<const float4*>(p + 4); return r;
    }
    static __device__ __forceinline__ void up(const Raw& r, float2 (&v)[4]) {
        v[0] = make_float2(r.a.x, r.a.y); v[1] = make_float2(r.a.z, r.a.w);
        v[2] = make_float2(r.b.x, r.b.y); v[3] = make_float2(r.b.z, r.b.w);
    }
    static __device__ __forceinline__ void st(float* p, const float2 (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
    }
};

template <int ACT> __device__ __forceinline__ float actv(float z) {
    if (ACT == KP_ACT_LEAKY) return z > 0.f ? z : 0.01f * z;
    if (ACT == KP_ACT_RELU) return z > 0.f ? z : 0.f;
    return z;
}
template <int ACT> __device__ __forceinline__ float actg(float z, float g) {
    if (ACT == KP_ACT_LEAKY) return z > 0.f ? g : 0.01f * g;
    if (ACT == KP_ACT_RELU) return z > 0.f ? g : 0.f;
    return g;
}
// packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2): half the issue slots of the scalar forms
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

__device__ __forceinline__ void load_c8(const float* p, int c0, float dflt, float2 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = p ? make_float2(p[c0 + 2 * i], p[c0 + 2 * i + 1]) : make_float2(dflt, dflt);
}

// ------------------------------------------------------------------------------------------------
// forward: out = post(act(scale * y + shift)) written into the (optionally replicate-padded) next input
// ------------------------------------------------------------------------------------------------
template <typename T, int POST, int ACT>
__global__ void __launch_bounds__(256, (POST == KP_POST_NONE && sizeof(T) == 2) ? 3 : 2)
bn_fwd_lean_k(Rows<const T> y, Rows<T> out, const float* __restrict__ scale, const float* __restrict__ shift, int pad, int N,
              int H, int W, int C, int OH, int OW, int cg_shift, const BnFuse fuse) {
    typedef typename P8<T>::Raw Raw;
    const int ncg = C >> 3;
    const int c0 = (threadIdx.x & (ncg - 1)) * 8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float2 sc[4], sh[4];
    if (fuse.stats) {
        float a[8], b[8];
        fused_affine(fuse, C, c0, blockIdx.x == 0 && x0 == 0, a, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) { sc[i] = make_float2(a[2 * i], a[2 * i + 1]); sh[i] = make_float2(b[2 * i], b[2 * i + 1]); }
    } else {
        load_c8(scale, c0, 1.f, sc);
        load_c8(shift, c0, 0.f, sh);
    }
    const int PH = OH + 2 * pad, PW = OW + 2 * pad, rows = N * PH;
    constexpr int NL = POST == KP_POST_POOL ? 4 : 1;
    constexpr int U = POST == KP_POST_POOL ? 2 : 4;
    if (POST == KP_POST_NONE) {
        // slot j of a pass -> (row offset jr, pixel offset jx).  Rows shorter than the block's pixel span put several
        // rows in flight per pass (S slots per row, RB = U / S rows) instead of leaving the upper slots idle.
        const int S = (PW + xstep - 1) / xstep, RB = (sizeof(T) == 2 && S <= U) ? U / S : 1;   // fp32 (parity) mode keeps one row per pass
        int jr[U], jx[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { jr[j] = RB > 1 ? j / S : 0; jx[j] = (RB > 1 ? j - jr[j] * S : j) * xstep; }
        for (int row0 = blockIdx.x * RB; row0 < rows; row0 += gridDim.x * RB) {
            for (int pxb = x0; pxb < PW; pxb += U * xstep) {
                Raw r[U];
                T* op[U];
                bool ok[U];
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int row = row0 + jr[j], px = pxb + jx[j];
                    ok[j] = jr[j] < RB && row < rows && px < PW;
                    if (ok[j]) {
                        const int n = row / PH, py = row - n * PH;
                        const int oy = min(max(py - pad, 0), OH - 1), ox = min(max(px - pad, 0), OW - 1);
                        r[j] = P8<T>::ld(y.row(n, oy) + c0 + ox * y.sx);
                        op[j] = out.row(n, py) + c0 + px * out.sx;
                    }
                }
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    if (ok[j]) {
                        float2 v[4];
                        P8<T>::up(r[j], v);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 z = fma2(v[i], sc[i], sh[i]);
                            v[i] = make_float2(actv<ACT>(z.x), actv<ACT>(z.y));
                        }
                        P8<T>::st(op[j], v);
                    }
                }
            }
        }
        return;
    }
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / PH, py = row - n * PH;
        const int oy = min(max(py - pad, 0), OH - 1);
        const T* r0 = y.row(n, POST == KP_POST_POOL ? 2 * oy : oy) + c0;
        const T* r1 = r0 + y.sy;
        T* orow = out.row(n, py) + c0;
        for (int pxb = x0; pxb < PW; pxb += U * xstep) {
            Raw r[U][NL];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int px = pxb + j * xstep;
                if (px < PW) {
                    const int ox = min(max(px - pad, 0), OW - 1);
                    if (POST == KP_POST_POOL) {
                        const T* a = r0 + 2 * ox * y.sx;
                        const T* b = r1 + 2 * ox * y.sx;
                        r[j][0] = P8<T>::ld(a); r[j][1 % NL] = P8<T>::ld(a + y.sx);
                        r[j][2 % NL] = P8<T>::ld(b); r[j][3 % NL] = P8<T>::ld(b + y.sx);
                    } else {
                        r[j][0] = P8<T>::ld(r0 + ox * y.sx);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int px = pxb + j * xstep;
                if (px < PW) {
                    float2 v[4];
                    P8<T>::up(r[j][0], v);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 z = fma2(v[i], sc[i], sh[i]);
                        v[i] = make_float2(actv<ACT>(z.x), actv<ACT>(z.y));
                    }
                    if (POST == KP_POST_POOL) {
#pragma unroll
                        for (int q = 1; q < NL; ++q) {
                            float2 a[4];
                            P8<T>::up(r[j][q], a);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 z = fma2(a[i], sc[i], sh[i]);
                                v[i].x = fmaxf(v[i].x, actv<ACT>(z.x));
                                v[i].y = fmaxf(v[i].y, actv<ACT>(z.y));
                            }
                        }
                    }
                    P8<T>::st(orow + px * out.sx, v);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward pass 1 (APPLY = false): dz = act'(z) * gather(dout)  [replicate-pad fold, max-pool routing], written to dy
// when given; per-channel sums of dz and dz * xhat.  APPLY = true: the gather is repeated and
// dy = scale (dz - mean(dz) - xhat mean(dz xhat)) is written directly.
// ------------------------------------------------------------------------------------------------
// the copies of an edge pixel in the replicate-padded border, other than its own interior position
struct F8 { float2 v[4]; };
template <typename T>
__device__ __noinline__ F8 fold_extra(const Rows<const T>& d, int n, int oy, int ox, int OH, int OW, int c0) {
    F8 s;
#pragma unroll
    for (int i = 0; i < 4; ++i) s.v[i] = make_float2(0.f, 0.f);
    for (int a = 0; a < 3; ++a) {
        if ((a == 1 && oy != 0) || (a == 2 && oy != OH - 1)) continue;
        const int ry = a == 0 ? oy + 1 : (a == 1 ? 0 : OH + 1);
        for (int b = 0; b < 3; ++b) {
            if ((b == 1 && ox != 0) || (b == 2 && ox != OW - 1) || (a == 0 && b == 0)) continue;
            const int rx = b == 0 ? ox + 1 : (b == 1 ? 0 : OW + 1);
            float2 t[4];
            P8<T>::up(P8<T>::ld(d.row(n, ry) + (long long)rx * d.sx + c0), t);
#pragma unroll
            for (int i = 0; i < 4; ++i) { s.v[i].x += t[i].x; s.v[i].y += t[i].y; }
        }
    }
    return s;
}
template <typename T>
__device__ __forceinline__ void add_fold(const Rows<const T>& d, int n, int oy, int ox, int OH, int OW, int c0, float2 (&g)[4]) {
    const F8 e = fold_extra<T>(d, n, oy, ox, OH, OW, c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) { g[i].x += e.v[i].x; g[i].y += e.v[i].y; }
}

template <typename T, int POST, int ACT, bool APPLY>
__global__ void __launch_bounds__(256, 2)
bn_bwd_lean_k(Rows<const T> dout, Rows<const T> y, Rows<T> dy, const float* __restrict__ scale,
              const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
              double* sums, double count, int pad, int N, int H, int W, int C, int OH, int OW, int cg_shift) {
    typedef typename P8<T>::Raw Raw;
    __shared__ float red[2][256 * 8];
    const int ncg = C >> 3;
    const int c0 = (threadIdx.x & (ncg - 1)) * 8;
    const int x0 = threadIdx.x >> cg_shift, xstep = 256 >> cg_shift;
    float2 sc[4], sh[4], nmu[4], s1[4], s2[4];
    load_c8(scale, c0, 1.f, sc);
    load_c8(shift, c0, 0.f, sh);
    load_c8(mean, c0, 0.f, nmu);
#pragma unroll
    for (int i = 0; i < 4; ++i) nmu[i] = make_float2(-nmu[i].x, -nmu[i].y);
    if (APPLY) {      // dy = sc*dz + s2*(y - mu) + s1   with s1 = -sc*mean(dz), s2 = -sc*invstd*mean(dz*xhat)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + 2 * i;
            const double rc = 1.0 / count;
            const float m1x = (float)(sums[c] * rc), m1y = (float)(sums[c + 1] * rc);
            const float m2x = (float)(sums[C + c] * rc), m2y = (float)(sums[C + c + 1] * rc);
            s1[i] = make_float2(-sc[i].x * m1x, -sc[i].y * m1y);
            s2[i] = make_float2(-sc[i].x * invstd[c] * m2x, -sc[i].y * invstd[c + 1] * m2y);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) { s1[i] = make_float2(0.f, 0.f); s2[i] = make_float2(0.f, 0.f); }
    }
    // g = gradient w.r.t. the activation output at one position, yv = raw conv output there
    auto emit = [&](T* dst, float2 (&g)[4], const float2 (&yv)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 z = fma2(yv[i], sc[i], sh[i]);
            const float2 dz = make_float2(actg<ACT>(z.x, g[i].x), actg<ACT>(z.y, g[i].y));
            const float2 yc = add2(yv[i], nmu[i]);
            if (APPLY) {
                g[i] = fma2(sc[i], dz, fma2(s2[i], yc, s1[i]));
            } else {
                g[i] = dz;
                s1[i] = add2(s1[i], dz);
                s2[i] = fma2(dz, yc, s2[i]);
            }
        }
        if (APPLY || dst) P8<T>::st(dst, g);
    };
    if (POST == KP_POST_NONE) {
        constexpr int U = 4;
        const int rows = N * H;
        // slot mapping as in the forward kernel: short rows put RB rows in flight per pass
        const int S = (W + xstep - 1) / xstep, RB = (sizeof(T) == 2 && S <= U) ? U / S : 1;   // fp32 (parity) mode keeps one row per pass
        int jr[U], jx[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { jr[j] = RB > 1 ? j / S : 0; jx[j] = (RB > 1 ? j - jr[j] * S : j) * xstep; }
        for (int row0 = blockIdx.x * RB; row0 < rows; row0 += gridDim.x * RB) {
            for (int xb = x0; xb < W; xb += U * xstep) {
                Raw rg[U], ry[U];
                int sn[U], sy[U];
                bool ok[U];
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int row = row0 + jr[j], xx = xb + jx[j];
                    ok[j] = jr[j] < RB && row < rows && xx < W;
                    if (ok[j]) {
                        sn[j] = row / H; sy[j] = row - sn[j] * H;
                        rg[j] = P8<T>::ld(dout.row(sn[j], sy[j] + pad) + (xx + pad) * dout.sx + c0);
                        ry[j] = P8<T>::ld(y.row(sn[j], sy[j]) + xx * y.sx + c0);
                    }
                }
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    if (ok[j]) {
                        const int xx = xb + jx[j], yy = sy[j];
                        float2 g[4], yv[4];
                        P8<T>::up(rg[j], g);
                        P8<T>::up(ry[j], yv);
                        if (pad && (yy == 0 || yy == H - 1 || xx == 0 || xx == W - 1)) add_fold<T>(dout, sn[j], yy, xx, OH, OW, c0, g);
                        emit(dy.p ? dy.row(sn[j], yy) + xx * dy.sx + c0 : nullptr, g, yv);
                    }
                }
            }
        }
    } else {   // 2x2 max-pool: one window per item, first maximum (row-major) wins
        const int HW2 = (H + 1) >> 1, WW2 = (W + 1) >> 1, rows = N * HW2;
        for (int row = blockIdx.x; row < rows; row += gridDim.x) {
            const int n = row / HW2, wy = (row - n * HW2) * 2;
            const bool has_r1 = wy + 1 < H;
            const T* y0 = y.row(n, wy) + c0;
            const T* y1 = y0 + y.sy;
            T* o0 = dy.p ? dy.row(n, wy) + c0 : nullptr;
            T* o1 = o0 ? o0 + dy.sy : nullptr;
            const int oy = wy >> 1;
            const T* drow = dout.row(n, min(oy, OH - 1) + pad) + pad * dout.sx + c0;
            const bool rowb = pad && (oy == 0 || oy == OH - 1);
            for (int wi = x0; wi < WW2; wi += xstep) {
                const int wx = wi * 2;
                const bool has_c1 = wx + 1 < W;
                const bool full = has_r1 && has_c1;
                Raw ry[4], rg;
                ry[0] = P8<T>::ld(y0 + wx * y.sx);
                if (has_c1) ry[1] = P8<T>::ld(y0 + (wx + 1) * y.sx);
                if (has_r1) ry[2] = P8<T>::ld(y1 + wx * y.sx);
                if (full) { ry[3] = P8<T>::ld(y1 + (wx + 1) * y.sx); rg = P8<T>::ld(drow + wi * dout.sx); }
                float2 yv[4][4], t[4];
                int bi[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 8; ++i) bi[i] = -1;
                if (full) {
                    P8<T>::up(rg, t);
                    if (pad && (rowb || wi == 0 || wi == OW - 1)) add_fold<T>(dout, n, oy, wi, OH, OW, c0, t);
                    float best[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        P8<T>::up(ry[q], yv[q]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 z = fma2(yv[q][i], sc[i], sh[i]);
                            const float ax = actv<ACT>(z.x), ay = actv<ACT>(z.y);
                            if (q == 0 || ax > best[2 * i]) { best[2 * i] = ax; bi[2 * i] = q; }
                            if (q == 0 || ay > best[2 * i + 1]) { best[2 * i + 1] = ay; bi[2 * i + 1] = q; }
                        }
                    }
                } else {
                    P8<T>::up(ry[0], yv[0]);
                    if (has_c1) P8<T>::up(ry[1], yv[1]);
                    if (has_r1) P8<T>::up(ry[2], yv[2]);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool ok = (q == 0) || (q == 1 && has_c1) || (q == 2 && has_r1) || (q == 3 && full);
                    if (ok) {
                        float2 g[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            g[i] = make_float2(bi[2 * i] == q ? t[i].x : 0.f, bi[2 * i + 1] == q ? t[i].y : 0.f);
                        T* orow = (q >> 1) ? o1 : o0;
                        emit(orow ? orow + (wx + (q & 1)) * dy.sx : nullptr, g, yv[q]);
                    }
                }
            }
        }
    }
    if (APPLY) return;
    float2 is[4];
    load_c8(invstd, c0, 1.f, is);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        red[0][threadIdx.x * 8 + 2 * i] = s1[i].x; red[0][threadIdx.x * 8 + 2 * i + 1] = s1[i].y;
        red[1][threadIdx.x * 8 + 2 * i] = s2[i].x * is[i].x; red[1][threadIdx.x * 8 + 2 * i + 1] = s2[i].y * is[i].y;
    }
    __syncthreads();
    // one channel per thread: sum over the 256/ncg threads that share its channel group
    for (int ch = threadIdx.x; ch < C; ch += 256) {
        const int g8 = ch >> 3, i = ch & 7;
        float a = 0.f, b = 0.f;
        for (int t = g8; t < 256; t += ncg) { a += red[0][t * 8 + i]; b += red[1][t * 8 + i]; }
        atomicAdd(&sums[ch], (double)a);
        atomicAdd(&sums[C + ch], (double)b);
    }
}
