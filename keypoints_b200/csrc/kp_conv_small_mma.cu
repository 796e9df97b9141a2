// 3x3 convolutions with 16 or 32 input AND output channels (the VGG_PONG* nets of the reference, vgg.py:48-70) in bf16
// mode.  A 64-channel tcgen05 k-block / 128-row UMMA tile would be 50-94 % padding at these widths, and the fp32 SIMT
// kernels reach ~6 TFLOP/s, so these layers run on warp-level mma.sync.m16n8k16 instead: A fragments are gathered
// straight from the NHWC buffer with 4-byte loads (16 consecutive pixels of one image row per MMA tile, masked at the
// row end), the weights sit in shared memory pre-arranged in B-fragment order, fp32 accumulation.
//   conv  : out[pix][co] = b[co] + sum_{t,ci} in[pix + tap t][ci] w[t*CIN + ci][co]   (fprop, and dgrad with the flipped /
//           transposed weights the caller packs; CHECK = bounds-checked taps for dgrad's zero halo)
//   wgrad : dw[co][ci][t] += sum_pix in[pix + tap t][ci] dy[pix][co]; one warp per tap, the dy fragments are built from
//           4-byte loads with byte permutes (both operands have pixels as the slow axis; output columns relabelled)
#include "kp_common.cuh"
#include <stdlib.h>

namespace {

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t ld32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint32_t ld16(const bf16* p) { return (uint32_t)*reinterpret_cast<const unsigned short*>(p); }

constexpr int SM_WARPS = 8;

// tile = 16 consecutive x of one output row; tiles_per_row = ceil(OW / 16)
template <int CIN, int COUT, bool CHECK, int TAPS>
__global__ void __launch_bounds__(SM_WARPS * 32, 2)
small_mma_conv_k(View<bf16> in, const float* __restrict__ wk, const float* __restrict__ bias, View<bf16> out, double* stats,
                 int N, int OH, int OW, int IH, int IW, int off) {
    constexpr int KH = CIN / 16, KS = TAPS * KH, NT = COUT / 8;      // TAPS = 9 (3x3) or 1 (1x1)
    __shared__ uint2 wf[KS][NT][32];
    __shared__ float red[SM_WARPS][2][COUT];
    __shared__ float sbias[COUT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) {
        const int ln = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
        const int co = nt * 8 + (ln >> 2), k0 = ks * 16 + 2 * (ln & 3);
        uint2 v;
        v.x = pk2(wk[(k0)*COUT + co], wk[(k0 + 1) * COUT + co]);
        v.y = pk2(wk[(k0 + 8) * COUT + co], wk[(k0 + 9) * COUT + co]);
        wf[ks][nt][ln] = v;
    }
    if (threadIdx.x < COUT) sbias[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    float s1[NT][2], s2[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) { s1[j][0] = s1[j][1] = s2[j][0] = s2[j][1] = 0.f; }
    const unsigned tpr = (unsigned)(OW + 15) >> 4;
    const unsigned ntiles = (unsigned)N * (unsigned)OH * tpr;
    for (unsigned tile = blockIdx.x * SM_WARPS + warp; tile < ntiles; tile += gridDim.x * SM_WARPS) {
        const unsigned r = tile / tpr;
        const int x0 = (int)(tile - r * tpr) << 4;
        const int n = (int)(r / (unsigned)OH), y = (int)(r - (unsigned)n * (unsigned)OH);
        const int xa = x0 + g, xb = xa + 8;
        const bool va = xa < OW, vb = xb < OW;
        const bf16* pa = in.at(n, y + off, xa + off, 2 * q);
        const bf16* pb = pa + 8 * in.sx;
        float c[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float b0 = sbias[8 * j + 2 * q], b1 = sbias[8 * j + 2 * q + 1];
            c[j][0] = b0; c[j][1] = b1; c[j][2] = b0; c[j][3] = b1;
        }
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
            const int ty = TAPS == 9 ? t / 3 : 0, tx = TAPS == 9 ? t % 3 : 0;
            bool oa = va, ob = vb;
            if (CHECK) {
                const int iy = y + ty + off;
                const bool oky = iy >= 0 && iy < IH;
                oa = oa && oky && (xa + tx + off) >= 0 && (xa + tx + off) < IW;
                ob = ob && oky && (xb + tx + off) >= 0 && (xb + tx + off) < IW;
            }
            const long long toff = ty * in.sy + tx * in.sx;
#pragma unroll
            for (int h = 0; h < KH; ++h) {
                uint32_t a[4];
                a[0] = oa ? ld32(pa + toff + 16 * h) : 0u;
                a[1] = ob ? ld32(pb + toff + 16 * h) : 0u;
                a[2] = oa ? ld32(pa + toff + 16 * h + 8) : 0u;
                a[3] = ob ? ld32(pb + toff + 16 * h + 8) : 0u;
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const uint2 w = wf[t * KH + h][j][lane];
                    mma16816(c[j], a, w.x, w.y);
                }
            }
        }
        bf16* oa_p = out.at(n, y, xa, 2 * q);
        bf16* ob_p = oa_p + 8 * out.sx;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            if (va) {
                *reinterpret_cast<uint32_t*>(oa_p + 8 * j) = pk2(c[j][0], c[j][1]);
                s1[j][0] += c[j][0]; s1[j][1] += c[j][1];
                s2[j][0] = fmaf(c[j][0], c[j][0], s2[j][0]); s2[j][1] = fmaf(c[j][1], c[j][1], s2[j][1]);
            }
            if (vb) {
                *reinterpret_cast<uint32_t*>(ob_p + 8 * j) = pk2(c[j][2], c[j][3]);
                s1[j][0] += c[j][2]; s1[j][1] += c[j][3];
                s2[j][0] = fmaf(c[j][2], c[j][2], s2[j][0]); s2[j][1] = fmaf(c[j][3], c[j][3], s2[j][1]);
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a = s1[j][e], b = s2[j][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
                if (g == 0) { red[warp][0][8 * j + 2 * q + e] = a; red[warp][1][8 * j + 2 * q + e] = b; }
            }
        __syncthreads();
        if (threadIdx.x < 2 * COUT) {
            const int w = threadIdx.x / COUT, ch = threadIdx.x % COUT;
            float a = 0.f;
            for (int k = 0; k < SM_WARPS; ++k) a += red[k][w][ch];
            atomicAdd(&stats[w * COUT + ch], (double)a);
        }
    }
}

// block = 9 warps, warp t accumulates tap t: D[ci (M)][co (N)] over the block's pixel tiles
template <int CIN, int COUT, int TAPS>
__global__ void __launch_bounds__(288, 2)
small_mma_wgrad_k(View<bf16> x, View<bf16> dy, float* __restrict__ dw, int N, int H, int W) {
    constexpr int MT = CIN / 16, NJ = COUT / 16;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    // 3x3: warp = tap, every warp walks all tiles of the block.  1x1: one tap, the 9 warps split the tiles.
    const int t = TAPS == 9 ? wid : 0;
    const int ty = TAPS == 9 ? t / 3 : 0, tx = TAPS == 9 ? t % 3 : 0;
    float c[MT][2 * NJ][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 2 * NJ; ++j) { c[m][j][0] = c[m][j][1] = c[m][j][2] = c[m][j][3] = 0.f; }
    const unsigned tpr = (unsigned)(W + 15) >> 4;
    const unsigned ntiles = (unsigned)N * (unsigned)H * tpr;
    const unsigned first = TAPS == 9 ? blockIdx.x : blockIdx.x * 9 + wid, stride = TAPS == 9 ? gridDim.x : gridDim.x * 9;
    for (unsigned tile = first; tile < ntiles; tile += stride) {
        const unsigned r = tile / tpr;
        const int x0 = (int)(tile - r * tpr) << 4;
        const int n = (int)(r / (unsigned)H), y = (int)(r - (unsigned)n * (unsigned)H);
        // the 4 pixels this thread contracts over: x0 + {2q, 2q+1, 2q+8, 2q+9}
        const int xp = x0 + 2 * q;
        bool v[4];
        v[0] = xp < W; v[1] = xp + 1 < W; v[2] = xp + 8 < W; v[3] = xp + 9 < W;
        const bf16* pd = dy.at(n, y, xp, 2 * g);
        const bf16* px = x.at(n, y + ty, xp + tx, g);
        uint32_t u[4][NJ];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            const bf16* row = pd + ((pp & 1) + 8 * (pp >> 1)) * dy.sx;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) u[pp][jj] = v[pp] ? ld32(row + 16 * jj) : 0u;
        }
        uint32_t a[MT][4];
#pragma unroll
        for (int m = 0; m < MT; ++m) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                     // ci = 16 m + g + 8 hf
                const bf16* p0 = px + 16 * m + 8 * hf;
                const uint32_t v0 = v[0] ? ld16(p0) : 0u, v1 = v[1] ? ld16(p0 + x.sx) : 0u;
                const uint32_t v8 = v[2] ? ld16(p0 + 8 * x.sx) : 0u, v9 = v[3] ? ld16(p0 + 9 * x.sx) : 0u;
                a[m][hf] = v0 | (v1 << 16);
                a[m][2 + hf] = v8 | (v9 << 16);
            }
        }
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
            const uint32_t e0 = __byte_perm(u[0][jj], u[1][jj], 0x5410), o0 = __byte_perm(u[0][jj], u[1][jj], 0x7632);
            const uint32_t e1 = __byte_perm(u[2][jj], u[3][jj], 0x5410), o1 = __byte_perm(u[2][jj], u[3][jj], 0x7632);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                mma16816(c[m][2 * jj], a[m], e0, e1);
                mma16816(c[m][2 * jj + 1], a[m], o0, o1);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 2 * NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ci = 16 * m + g + 8 * (e >> 1);
                const int col = 2 * q + (e & 1);
                const int co = 16 * (j >> 1) + 2 * col + (j & 1);
                atomicAdd(&dw[((long long)co * CIN + ci) * TAPS + t], c[m][j][e]);
            }
}

// ------------------------------------------------------------------------------------------------
// Pipelined 3x3 wgrad: the register-gather kernel above has one 2 KB tile in flight per block and is bound by global-load
// latency (~1 us per tile).  Here every thread issues one 16-byte cp.async per tile into an SW_STAGES-deep shared-memory
// ring (x: 3 rows x 18 pixels, dy: 16 pixels; out-of-row pixels zero-filled), and the nine tap warps read both operands
// with ldmatrix.trans (pixels are the contraction axis and the slow axis of both NHWC buffers).  The 16-byte chunks of a
// pixel are XOR-swizzled with the pixel index so the eight rows of one ldmatrix phase hit distinct bank groups.
// ------------------------------------------------------------------------------------------------
constexpr int SW_STAGES = 4, SW_TPS = 4;        // ring depth, 16-pixel tiles per stage

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 16-byte chunk position of logical chunk ch of pixel p (CH chunks per pixel)
template <int CH>
__device__ __forceinline__ int swz(int p, int ch) { return CH == 4 ? ch ^ ((p >> 1) & 3) : ch ^ ((p >> 2) & 1); }

template <int CIN, int COUT>
__global__ void __launch_bounds__(288, 2)
small_mma_wgrad_pipe_k(View<bf16> x, View<bf16> dy, float* __restrict__ dw, int N, int H, int W, unsigned per_block) {
    constexpr int MT = CIN / 16, NJ = COUT / 16, XCH = CIN / 8, DCH = COUT / 8;
    constexpr int XROW = 18 * XCH * 16, XB = 3 * XROW, DB = 16 * DCH * 16, TB = XB + DB;   // bytes per tile
    constexpr int SB = SW_TPS * TB;                                                      // bytes per stage
    constexpr int NX = 3 * 18 * XCH, ND = 16 * DCH;                                      // 16-byte chunks per tile
    static_assert(NX + ND <= 288, "one chunk per thread per tile");
    extern __shared__ __align__(128) unsigned char sw_smem[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sw_smem);
    const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int ty = t / 3, tx = t % 3;
    const unsigned tpr = (unsigned)(W + 15) >> 4;
    const unsigned ntiles = (unsigned)N * (unsigned)H * tpr;
    // the block owns the contiguous tile range [first, first + mine): the copy cursor (n, y, x0) advances without divisions
    const unsigned first = blockIdx.x * per_block;
    const unsigned mine = first < ntiles ? min(per_block, ntiles - first) : 0u;
    const unsigned nstages = (mine + SW_TPS - 1) / SW_TPS;

    // this thread's chunk of every tile (tile-invariant part of the addresses)
    const int i = threadIdx.x;
    const bool is_x = i < NX, is_d = !is_x && i < NX + ND;
    int cp_p, cp_lim;
    long long cp_src;
    uint32_t cp_dst;
    if (is_x) {
        const int row = i / (18 * XCH), rem = i - row * (18 * XCH), p = rem / XCH, ch = rem - p * XCH;
        cp_p = p; cp_lim = W + 2;
        cp_src = row * x.sy + p * x.sx + ch * 8;
        cp_dst = row * XROW + (p * XCH + swz<XCH>(p, ch)) * 16;
    } else {
        const int k2 = is_d ? i - NX : 0, p = k2 / DCH, ch = k2 - p * DCH;
        cp_p = p; cp_lim = W;
        cp_src = p * dy.sx + ch * 8;
        cp_dst = XB + (p * DCH + swz<DCH>(p, ch)) * 16;
    }
    const bf16* cp_base = is_x ? x.p : dy.p;
    const long long s_n = is_x ? x.sn : dy.sn, s_y = is_x ? x.sy : dy.sy, s_x = is_x ? x.sx : dy.sx;
    int cn, cy, cx0;                                   // cursor of the next tile to copy
    {
        const unsigned r = first / tpr;
        cx0 = (int)(first - r * tpr) << 4;
        cn = (int)(r / (unsigned)H); cy = (int)(r - (unsigned)cn * (unsigned)H);
    }
    unsigned copied = 0;
    auto issue_stage = [&](unsigned stage) {
        const uint32_t st = sbase + (stage % SW_STAGES) * SB;
#pragma unroll
        for (int j = 0; j < SW_TPS; ++j) {
            if (copied < mine) {
                if (is_x || is_d) {
                    const uint32_t dst = st + j * TB + cp_dst;
                    if (cx0 + cp_p < cp_lim) cp_async16(dst, cp_base + cn * s_n + cy * s_y + cx0 * s_x + cp_src);
                    else *reinterpret_cast<uint4*>(sw_smem + (dst - sbase)) = make_uint4(0u, 0u, 0u, 0u);
                }
                ++copied;
                cx0 += 16;
                if (cx0 >= W) { cx0 = 0; if (++cy == H) { cy = 0; ++cn; } }
            }
        }
    };

    float c[MT][2 * NJ][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 2 * NJ; ++j) { c[m][j][0] = c[m][j][1] = c[m][j][2] = c[m][j][3] = 0.f; }
    // ldmatrix row addresses of this lane: matrix mi = lane / 8, row rr = lane % 8
    const int mi = lane >> 3, rr = lane & 7;
    const int pa = rr + 8 * (mi >> 1) + tx, ca = mi & 1;           // A (x): pixel, chunk parity
    const int pb = rr + 8 * (mi & 1), cb = mi >> 1;                // B (dy)
    uint32_t offa[MT], offb[NJ];
#pragma unroll
    for (int m = 0; m < MT; ++m) offa[m] = ty * XROW + (pa * XCH + swz<XCH>(pa, 2 * m + ca)) * 16;
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) offb[jj] = XB + (pb * DCH + swz<DCH>(pb, 2 * jj + cb)) * 16;

    for (unsigned k = 0; k < SW_STAGES - 1; ++k) {
        if (k < nstages) issue_stage(k);
        cp_async_commit();
    }
    for (unsigned k = 0; k < nstages; ++k) {
        cp_async_wait<SW_STAGES - 2>();
        __syncthreads();
        if (k + SW_STAGES - 1 < nstages) issue_stage(k + SW_STAGES - 1);
        cp_async_commit();
        const uint32_t st = sbase + (k % SW_STAGES) * SB;
        const unsigned left = mine - k * SW_TPS;
#pragma unroll
        for (int j = 0; j < SW_TPS; ++j) {
            if ((unsigned)j < left) {
                uint32_t a[MT][4];
#pragma unroll
                for (int m = 0; m < MT; ++m) ldsm4t(a[m], st + j * TB + offa[m]);
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    uint32_t b[4];
                    ldsm4t(b, st + j * TB + offb[jj]);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        mma16816(c[m][2 * jj], a[m], b[0], b[1]);
                        mma16816(c[m][2 * jj + 1], a[m], b[2], b[3]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // block partial [co][ci][9] staged in shared memory (the ring is free now), then added to dw with coalesced vector REDs
    float* part = reinterpret_cast<float*>(sw_smem);
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 2 * NJ; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ci = 16 * m + g + 8 * (e >> 1);
                const int co = 8 * j + 2 * q + (e & 1);
                part[(co * CIN + ci) * 9 + t] = c[m][j][e];
            }
    __syncthreads();
    if (mine == 0) return;
    constexpr int TOT = CIN * COUT * 9;
    if ((reinterpret_cast<uintptr_t>(dw) & 15) == 0) {
        for (int v = threadIdx.x; v < TOT / 4; v += 288) red_add_v4(dw + 4 * v, *reinterpret_cast<const float4*>(part + 4 * v));
    } else {
        for (int v = threadIdx.x; v < TOT; v += 288) atomicAdd(dw + v, part[v]);
    }
}

// ------------------------------------------------------------------------------------------------
// Pipelined 3x3 fprop / dgrad: small_mma_conv_k gathers the A fragments with 4-byte global loads and is bound by L1
// throughput (ncu: l1tex 94 %, 9 taps x 16 pixels re-read per tile).  Here every warp owns a contiguous range of tiles and a
// private SC_DEPTH-deep cp.async ring: the 3 x 18-pixel input window of a tile is copied once in 16-byte chunks (zero
// outside the image: the bounds check of the dgrad halo comes for free), A fragments come from ldmatrix, weights from the
// B-fragment table in shared memory as before.  No block barrier in the main loop.
// ------------------------------------------------------------------------------------------------
constexpr int SC_DEPTH = 3;

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

struct TileCursor {
    int n, y, x0;
    __device__ __forceinline__ void seek(unsigned tile, unsigned tpr, int OH) {
        const unsigned r = tile / tpr;
        x0 = (int)(tile - r * tpr) << 4;
        n = (int)(r / (unsigned)OH); y = (int)(r - (unsigned)n * (unsigned)OH);
    }
    __device__ __forceinline__ void next(int OH, int OW) {
        x0 += 16;
        if (x0 >= OW) { x0 = 0; if (++y == OH) { y = 0; ++n; } }
    }
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(SM_WARPS * 32, 2)
small_mma_conv_pipe_k(View<bf16> in, const float* __restrict__ wk, const float* __restrict__ bias, View<bf16> out, double* stats,
                      int N, int OH, int OW, int IH, int IW, int off, unsigned per_warp) {
    constexpr int KH = CIN / 16, KS = 9 * KH, NT = COUT / 8, XCH = CIN / 8;
    constexpr int XROW = 18 * XCH * 16, TB = 3 * XROW, NXC = 3 * 18 * XCH;
    __shared__ uint2 wf[KS][NT][32];
    __shared__ float red[SM_WARPS][2][COUT];
    __shared__ float sbias[COUT];
    extern __shared__ __align__(128) unsigned char sc_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) {
        const int ln = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
        const int co = nt * 8 + (ln >> 2), k0 = ks * 16 + 2 * (ln & 3);
        uint2 v;
        v.x = pk2(wk[(k0)*COUT + co], wk[(k0 + 1) * COUT + co]);
        v.y = pk2(wk[(k0 + 8) * COUT + co], wk[(k0 + 9) * COUT + co]);
        wf[ks][nt][ln] = v;
    }
    if (threadIdx.x < COUT) sbias[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
    __syncthreads();
    float s1[NT][2], s2[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) { s1[j][0] = s1[j][1] = s2[j][0] = s2[j][1] = 0.f; }
    const unsigned tpr = (unsigned)(OW + 15) >> 4;
    const unsigned ntiles = (unsigned)N * (unsigned)OH * tpr;
    const unsigned first = (blockIdx.x * SM_WARPS + warp) * per_warp;
    const unsigned mine = first < ntiles ? min(per_warp, ntiles - first) : 0u;
    unsigned char* ring = sc_smem + warp * (SC_DEPTH * TB);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    TileCursor cc, cu;
    cc.seek(first, tpr, OH);
    cu = cc;

    auto copy_tile = [&](int slot) {
        const uint32_t st = ring_s + slot * TB;
        for (int i = lane; i < NXC; i += 32) {
            const int row = i / (18 * XCH), rem = i - row * (18 * XCH), p = rem / XCH, ch = rem - p * XCH;
            const int iy = cc.y + row + off, ix = cc.x0 + p + off;
            const uint32_t dst = st + row * XROW + (p * XCH + swz<XCH>(p, ch)) * 16;
            if (iy >= 0 && iy < IH && ix >= 0 && ix < IW) cp_async16(dst, in.at(cc.n, iy, ix, ch * 8));
            else *reinterpret_cast<uint4*>(ring + (dst - ring_s)) = make_uint4(0u, 0u, 0u, 0u);
        }
        cc.next(OH, OW);
    };
    // ldmatrix lane addressing: matrix mi = lane / 8 (a0..a3), row rr = lane % 8
    const int mi = lane >> 3, rr = lane & 7;
    uint32_t offa[3][KH];
#pragma unroll
    for (int tx = 0; tx < 3; ++tx)
#pragma unroll
        for (int h = 0; h < KH; ++h) {
            const int p = rr + 8 * (mi & 1) + tx;
            offa[tx][h] = (p * XCH + swz<XCH>(p, 2 * h + (mi >> 1))) * 16;
        }

    for (unsigned k = 0; k < SC_DEPTH - 1; ++k) {
        if (k < mine) copy_tile(k);
        cp_async_commit();
    }
    for (unsigned k = 0; k < mine; ++k) {
        cp_async_wait<SC_DEPTH - 2>();
        __syncwarp();
        if (k + SC_DEPTH - 1 < mine) copy_tile((k + SC_DEPTH - 1) % SC_DEPTH);
        cp_async_commit();
        const uint32_t st = ring_s + (k % SC_DEPTH) * TB;
        float c[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const float b0 = sbias[8 * j + 2 * q], b1 = sbias[8 * j + 2 * q + 1];
            c[j][0] = b0; c[j][1] = b1; c[j][2] = b0; c[j][3] = b1;
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int h = 0; h < KH; ++h) {
                uint32_t a[4];
                ldsm4(a, st + (t / 3) * XROW + offa[t % 3][h]);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const uint2 w = wf[t * KH + h][j][lane];
                    mma16816(c[j], a, w.x, w.y);
                }
            }
        }
        const int xa = cu.x0 + g, xb = xa + 8;
        const bool va = xa < OW, vb = xb < OW;
        bf16* oa_p = out.at(cu.n, cu.y, xa, 2 * q);
        bf16* ob_p = oa_p + 8 * out.sx;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            if (va) {
                *reinterpret_cast<uint32_t*>(oa_p + 8 * j) = pk2(c[j][0], c[j][1]);
                s1[j][0] += c[j][0]; s1[j][1] += c[j][1];
                s2[j][0] = fmaf(c[j][0], c[j][0], s2[j][0]); s2[j][1] = fmaf(c[j][1], c[j][1], s2[j][1]);
            }
            if (vb) {
                *reinterpret_cast<uint32_t*>(ob_p + 8 * j) = pk2(c[j][2], c[j][3]);
                s1[j][0] += c[j][2]; s1[j][1] += c[j][3];
                s2[j][0] = fmaf(c[j][2], c[j][2], s2[j][0]); s2[j][1] = fmaf(c[j][3], c[j][3], s2[j][1]);
            }
        }
        cu.next(OH, OW);
    }
    cp_async_wait<0>();
    if (stats) {
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a = s1[j][e], b = s2[j][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
                if (g == 0) { red[warp][0][8 * j + 2 * q + e] = a; red[warp][1][8 * j + 2 * q + e] = b; }
            }
        __syncthreads();
        if (threadIdx.x < 2 * COUT) {
            const int w = threadIdx.x / COUT, ch = threadIdx.x % COUT;
            float a = 0.f;
            for (int k = 0; k < SM_WARPS; ++k) a += red[k][w][ch];
            atomicAdd(&stats[w * COUT + ch], (double)a);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Single-channel first layer (grey-scale nets: 1 -> Cout, 3x3).  9 MACs per output: a streaming kernel, one thread per
// (pixel, 8 output channels), the 3x3 window comes from L1 (the 1-channel padded image is contiguous in x).
// ------------------------------------------------------------------------------------------------
template <int G>        // G = Cout / 8 threads per pixel
__global__ void __launch_bounds__(256)
c1_fprop_k(View<bf16> in, const float* __restrict__ wk /*[9][Cout]*/, const float* __restrict__ bias, View<bf16> out,
           double* stats, int N, int OH, int OW) {
    constexpr int COUT = 8 * G, PPB = 256 / G;                   // pixels per block iteration
    __shared__ float red[2][256 * 8];
    const int gl = threadIdx.x % G, px = threadIdx.x / G;
    float w[9][8], b[8], s1[8], s2[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) w[t][i] = wk[t * COUT + gl * 8 + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { b[i] = bias ? bias[gl * 8 + i] : 0.f; s1[i] = 0.f; s2[i] = 0.f; }
    const int rows = N * OH;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / OH, y = row - n * OH;
        const bf16* r0 = in.at(n, y, 0, 0);
        bf16* orow = out.at(n, y, 0, gl * 8);
        for (int x = px; x < OW; x += PPB) {
            float v[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) v[t] = __bfloat162float(r0[(t / 3) * in.sy + (x + t % 3) * in.sx]);
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a = b[i];
#pragma unroll
                for (int t = 0; t < 9; ++t) a = fmaf(v[t], w[t][i], a);
                o[i] = a;
                s1[i] += a;
                s2[i] = fmaf(a, a, s2[i]);
            }
            uint4 u;
            u.x = pk2(o[0], o[1]); u.y = pk2(o[2], o[3]); u.z = pk2(o[4], o[5]); u.w = pk2(o[6], o[7]);
            *reinterpret_cast<uint4*>(orow + (long long)x * out.sx) = u;
        }
    }
    if (stats) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { red[0][threadIdx.x * 8 + i] = s1[i]; red[1][threadIdx.x * 8 + i] = s2[i]; }
        __syncthreads();
        for (int ch = threadIdx.x; ch < COUT; ch += 256) {
            const int g8 = ch >> 3, i = ch & 7;
            float a = 0.f, c = 0.f;
            for (int t = g8; t < 256; t += G) { a += red[0][t * 8 + i]; c += red[1][t * 8 + i]; }
            atomicAdd(&stats[ch], (double)a);
            atomicAdd(&stats[COUT + ch], (double)c);
        }
    }
}

// dw[co][0][t] += sum_pix dy[pix][co] * x[pix + tap t]
template <int G>
__global__ void __launch_bounds__(256)
c1_wgrad_k(View<bf16> x, View<bf16> dy, float* __restrict__ dw, int N, int H, int W) {
    constexpr int COUT = 8 * G, PPB = 256 / G;
    __shared__ float red[9 * 64];
    const int gl = threadIdx.x % G, px = threadIdx.x / G;
    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
    const int rows = N * H;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int n = row / H, y = row - n * H;
        const bf16* r0 = x.at(n, y, 0, 0);
        const bf16* drow = dy.at(n, y, 0, gl * 8);
        for (int xx = px; xx < W; xx += PPB) {
            const uint4 u = *reinterpret_cast<const uint4*>(drow + (long long)xx * dy.sx);
            const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
            float g[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { g[2 * i] = __uint_as_float(uw[i] << 16); g[2 * i + 1] = __uint_as_float(uw[i] & 0xffff0000u); }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float v = __bfloat162float(r0[(t / 3) * x.sy + (xx + t % 3) * x.sx]);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[t][i] = fmaf(v, g[i], acc[t][i]);
            }
        }
    }
    for (int i = threadIdx.x; i < 9 * COUT; i += 256) red[i] = 0.f;
    __syncthreads();
    // lanes l and l + G, l + 2G, ... of a warp hold the same channel group: combine them with shuffles first (128-way
    // contended shared-memory atomics were the whole run time of this kernel)
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = acc[t][i];
#pragma unroll
            for (int o = G; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane < G) atomicAdd(&red[t * COUT + gl * 8 + i], v);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * COUT; i += 256) {
        const int t = i / COUT, co = i - t * COUT;
        atomicAdd(&dw[co * 9 + t], red[i]);
    }
}

static bool c1_view_ok(const kp_view* v) { return v->dtype == KP_BF16 && v->sc == 1; }
static bool c1_wide_ok(const kp_view* v, int C) {
    return v->dtype == KP_BF16 && v->sc == 1 && C % 8 == 0 && C >= 8 && C <= 64 && (C & (C - 1)) == 0 &&
           (((uintptr_t)v->ptr) % 16) == 0 && v->sx % 8 == 0 && v->sy % 8 == 0 && v->sn % 8 == 0;
}

static bool small_view_ok(const kp_view* v, int C) {
    return v->dtype == KP_BF16 && v->sc == 1 && (((uintptr_t)v->ptr) % 4) == 0 && v->sx % 2 == 0 && v->sy % 2 == 0 &&
           v->sn % 2 == 0 && v->sx >= C;
}

}  // namespace

bool kp_small_mma_conv_ok(const kp_view* in, const kp_view* out, int N, int OH, int OW, int Cin, int Cout, int ks) {
    return (ks == 3 || ks == 1) && (Cin == 16 || Cin == 32) && (Cout == 16 || Cout == 32) && small_view_ok(in, Cin) && small_view_ok(out, Cout) &&
           (long long)N * OH * ((OW + 15) / 16) < (1LL << 31);
}

int kp_small_mma_conv(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out,
                      double* stats, int N, int OH, int OW, int IH, int IW, int Cin, int Cout, int ks, int off) {
    const long long tiles = (long long)N * OH * ((OW + 15) / 16);
    long long blocks = (tiles + SM_WARPS - 1) / SM_WARPS;
    const long long cap = (long long)kp_sm_count() * 4;
    if (blocks > cap) blocks = cap;
    const bool check = ks == 3 && !(off >= 0 && OH + off + 2 <= IH && OW + off + 2 <= IW);
    static int pipe = -1;
    if (pipe < 0) { const char* e = kp_env("KP_SMALL_CONV_PIPE"); pipe = e ? atoi(e) : 1; }
    if (pipe && ks == 3 && (((uintptr_t)in->ptr) % 16) == 0 && in->sx % 8 == 0 && in->sy % 8 == 0 && in->sn % 8 == 0) {
        long long pb = (tiles + SM_WARPS - 1) / SM_WARPS;
        const long long pcap = (long long)kp_sm_count() * 2;
        if (pb > pcap) pb = pcap;
        const unsigned per_warp = (unsigned)((tiles + pb * SM_WARPS - 1) / (pb * SM_WARPS));
#define KP_SMP(CI, CO)                                                                                                    \
    do {                                                                                                                  \
        constexpr int smem = SM_WARPS * SC_DEPTH * 3 * 18 * CI * 2;                                                       \
        static KpOncePerDevice attr;                                                                                           \
        if (attr.first()) {                                                                                                      \
            cudaFuncSetAttribute(small_mma_conv_pipe_k<CI, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);       \
        }                                                                                                                 \
        small_mma_conv_pipe_k<CI, CO><<<(unsigned)pb, SM_WARPS * 32, smem, st>>>(make_view<bf16>(in), wk, bias,           \
                                                                                make_view<bf16>(out), stats, N, OH, OW, IH, IW, \
                                                                                off, per_warp);                          \
    } while (0)
        if (Cin == 16 && Cout == 16) KP_SMP(16, 16);
        else if (Cin == 16 && Cout == 32) KP_SMP(16, 32);
        else if (Cin == 32 && Cout == 16) KP_SMP(32, 16);
        else KP_SMP(32, 32);
#undef KP_SMP
        KP_LAUNCH_CHECK();
        return KP_OK;
    }
    const dim3 grid((unsigned)blocks);
#define KP_SM(CI, CO)                                                                                                       \
    do {                                                                                                                    \
        if (ks == 1)                                                                                                        \
            small_mma_conv_k<CI, CO, false, 1><<<grid, SM_WARPS * 32, 0, st>>>(make_view<bf16>(in), wk, bias,               \
                                                                              make_view<bf16>(out), stats, N, OH, OW, IH, IW, off); \
        else if (check)                                                                                                     \
            small_mma_conv_k<CI, CO, true, 9><<<grid, SM_WARPS * 32, 0, st>>>(make_view<bf16>(in), wk, bias, make_view<bf16>(out), \
                                                                             stats, N, OH, OW, IH, IW, off);                \
        else                                                                                                                \
            small_mma_conv_k<CI, CO, false, 9><<<grid, SM_WARPS * 32, 0, st>>>(make_view<bf16>(in), wk, bias,               \
                                                                              make_view<bf16>(out), stats, N, OH, OW, IH, IW, off); \
    } while (0)
    if (Cin == 16 && Cout == 16) KP_SM(16, 16);
    else if (Cin == 16 && Cout == 32) KP_SM(16, 32);
    else if (Cin == 32 && Cout == 16) KP_SM(32, 16);
    else KP_SM(32, 32);
#undef KP_SM
    KP_LAUNCH_CHECK();
    return KP_OK;
}

bool kp_small_mma_wgrad_ok(const kp_view* x, const kp_view* dy, int N, int H, int W, int Cin, int Cout, int ks) {
    return (ks == 3 || ks == 1) && (Cin == 16 || Cin == 32) && (Cout == 16 || Cout == 32) && small_view_ok(x, Cin) && small_view_ok(dy, Cout) &&
           (long long)N * H * ((W + 15) / 16) < (1LL << 31);
}

int kp_small_mma_wgrad(cudaStream_t st, const kp_view* x, const kp_view* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                       int ks) {
    const long long tiles = (long long)N * H * ((W + 15) / 16);
    long long blocks = tiles;
    static int mult = -1;
    if (mult < 0) { const char* e = kp_env("KP_SMALL_WGRAD_MULT"); mult = e ? atoi(e) : 2; if (mult < 1) mult = 1; }
    const long long cap = (long long)kp_sm_count() * mult;
    if (blocks > cap) blocks = cap;
    const dim3 grid((unsigned)blocks);
    static int pipe = -1;
    if (pipe < 0) { const char* e = kp_env("KP_SMALL_WGRAD_PIPE"); pipe = e ? atoi(e) : 1; }
    if (pipe && ks == 3 && (((uintptr_t)x->ptr) % 16) == 0 && (((uintptr_t)dy->ptr) % 16) == 0 && x->sx % 8 == 0 && x->sy % 8 == 0 &&
        x->sn % 8 == 0 && dy->sx % 8 == 0 && dy->sy % 8 == 0 && dy->sn % 8 == 0) {
#define KP_SWP(CI, CO)                                                                                                   \
    do {                                                                                                                 \
        constexpr int smem = SW_STAGES * SW_TPS * (3 * 18 * CI * 2 + 16 * CO * 2);                                       \
        static KpOncePerDevice attr;                                                                                          \
        if (attr.first()) {                                                                                                     \
            cudaFuncSetAttribute(small_mma_wgrad_pipe_k<CI, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);     \
        }                                                                                                                \
        small_mma_wgrad_pipe_k<CI, CO><<<grid, 288, smem, st>>>(make_view<bf16>(x), make_view<bf16>(dy), dw, N, H, W,    \
                                                                (unsigned)((tiles + blocks - 1) / blocks));              \
    } while (0)
        if (Cin == 16 && Cout == 16) KP_SWP(16, 16);
        else if (Cin == 16 && Cout == 32) KP_SWP(16, 32);
        else if (Cin == 32 && Cout == 16) KP_SWP(32, 16);
        else KP_SWP(32, 32);
#undef KP_SWP
        KP_LAUNCH_CHECK();
        return KP_OK;
    }
#define KP_SW(CI, CO)                                                                                                    \
    do {                                                                                                                 \
        if (ks == 1) small_mma_wgrad_k<CI, CO, 1><<<grid, 288, 0, st>>>(make_view<bf16>(x), make_view<bf16>(dy), dw, N, H, W); \
        else small_mma_wgrad_k<CI, CO, 9><<<grid, 288, 0, st>>>(make_view<bf16>(x), make_view<bf16>(dy), dw, N, H, W);   \
    } while (0)
    if (Cin == 16 && Cout == 16) KP_SW(16, 16);
    else if (Cin == 16 && Cout == 32) KP_SW(16, 32);
    else if (Cin == 32 && Cout == 16) KP_SW(32, 16);
    else KP_SW(32, 32);
#undef KP_SW
    KP_LAUNCH_CHECK();
    return KP_OK;
}

bool kp_c1_conv_ok(const kp_view* in, const kp_view* out, int OH, int OW, int IH, int IW, int Cin, int Cout, int ks, int off) {
    return ks == 3 && Cin == 1 && off == 0 && IH >= OH + 2 && IW >= OW + 2 && c1_view_ok(in) && c1_wide_ok(out, Cout);
}

int kp_c1_fprop(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out, double* stats, int N,
                int OH, int OW, int Cout) {
    long long rows = (long long)N * OH;
    static int mult = -1;
    // 2 blocks per SM: every block ends with 2 * Cout double atomics on the same statistics words; at 8 blocks per SM their
    // serialisation in the L2 cost more than the extra row passes per block (33 -> 26 us at 84x84, 128 images)
    if (mult < 0) { const char* e = kp_env("KP_C1_FPROP_MULT"); mult = e ? atoi(e) : 2; if (mult < 1) mult = 1; }
    const long long cap = (long long)kp_sm_count() * mult;
    const int grid = (int)(rows < cap ? rows : cap);
#define KP_C1(GV) c1_fprop_k<GV><<<grid, 256, 0, st>>>(make_view<bf16>(in), wk, bias, make_view<bf16>(out), stats, N, OH, OW)
    if (Cout == 8) KP_C1(1); else if (Cout == 16) KP_C1(2); else if (Cout == 32) KP_C1(4); else KP_C1(8);
#undef KP_C1
    KP_LAUNCH_CHECK();
    return KP_OK;
}

bool kp_c1_wgrad_ok(const kp_view* x, const kp_view* dy, int Cin, int Cout, int ks) {
    return ks == 3 && Cin == 1 && c1_view_ok(x) && c1_wide_ok(dy, Cout);
}

int kp_c1_wgrad(cudaStream_t st, const kp_view* x, const kp_view* dy, float* dw, int N, int H, int W, int Cout) {
    long long rows = (long long)N * H;
    static int mult = -1;
    if (mult < 0) { const char* e = kp_env("KP_C1_WGRAD_MULT"); mult = e ? atoi(e) : 2; if (mult < 1) mult = 1; }
    const long long cap = (long long)kp_sm_count() * mult;
    const int grid = (int)(rows < cap ? rows : cap);
#define KP_C1W(GV) c1_wgrad_k<GV><<<grid, 256, 0, st>>>(make_view<bf16>(x), make_view<bf16>(dy), dw, N, H, W)
    if (Cout == 8) KP_C1W(1); else if (Cout == 16) KP_C1W(2); else if (Cout == 32) KP_C1W(4); else KP_C1W(8);
#undef KP_C1W
    KP_LAUNCH_CHECK();
    return KP_OK;
}
