// First convolution of a Unit (Cin <= 3, 3x3, 64*k outputs) in bf16 mode on the warp-level tensor cores.
// The layer has K = 9*Cin <= 27 (padded to 32): far too thin for a tcgen05 tile (a 64-channel k-block would be
// >95 % zeros) but a perfect fit for mma.sync.m16n8k16 — the patch matrix [pixels][32] is gathered straight from the
// replicate-padded NHWC input into A fragments (no im2col buffer, no shared-memory staging), so the kernels are bound
// by the 128-byte-per-pixel output (fprop) / dy (wgrad) stream in HBM instead of fp32 FMA issue.
//   fprop : Y[pix][co]   = sum_kk P[pix][kk] W[kk][co]      (+bias, bf16 store, fp32 BatchNorm statistics)
//   wgrad : dW[kk][co]   = sum_pix P[pix][kk] dY[pix][co]   (contraction over pixels; both operands have pixels as
//           the slow axis, so the B fragments are built with byte permutes from 4-byte loads and the output columns
//           are relabelled: n-tile 2j holds the even channels of group j, n-tile 2j+1 the odd ones)
// kk = tap*Cin + ci.  16 consecutive pixels of one image row form an MMA tile (requires OW % 16 == 0).
#include "kp_common.cuh"

namespace {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// element offset of patch entry kk relative to the (y, x) corner of a pixel's 3x3 window, or -1 for the zero padding
__device__ __forceinline__ int patch_off(int kk, int KK, int Cin, long long sy, long long sx, long long sc) {
    if (kk >= KK) return -1;
    const int t = kk / Cin, c = kk - t * Cin;
    return (int)((t / 3) * sy + (t % 3) * sx + c * sc);      // a few rows of one image: fits 32 bits
}

__device__ __forceinline__ uint32_t ldg_u16(const bf16* p) {
    return (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p));
}

constexpr int THIN_WARPS = 8;

// grid = (blocks over 16-pixel tiles, Cout / 64); block = 256 threads
__global__ void __launch_bounds__(THIN_WARPS * 32, 2)
thin_mma_fprop_k(View<bf16> in, const float* __restrict__ wk /*[KK][Cout]*/, const float* __restrict__ bias,
                 View<bf16> out, double* stats, int N, int OH, int OW, int Cin, int Cout) {
    __shared__ __align__(16) uint8_t stage[THIN_WARPS][16 * 128];
    __shared__ float red[THIN_WARPS][2][64];
    __shared__ float sbias[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int KK = 9 * Cin;
    const int cb = blockIdx.y * 64;
    // B fragments (weights, rounded to bf16 like every tensor-core layer): b[s][j] for k-step s, n-tile j
    uint32_t b0[2][8], b1[2][8];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = cb + 8 * j + g;
            const int k0 = 16 * s + 2 * q;
            auto wv = [&](int kk) { return kk < KK ? wk[(long long)kk * Cout + co] : 0.f; };
            b0[s][j] = pack_bf16(wv(k0), wv(k0 + 1));
            b1[s][j] = pack_bf16(wv(k0 + 8), wv(k0 + 9));
        }
    if (threadIdx.x < 64) sbias[threadIdx.x] = bias ? bias[cb + threadIdx.x] : 0.f;
    __syncthreads();
    // this thread's 8 patch entries: kk = 16 s + 8 h + 2 q + e
    int off[2][2][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int e = 0; e < 2; ++e) off[s][h][e] = patch_off(16 * s + 8 * h + 2 * q + e, KK, Cin, in.sy, in.sx, in.sc);
    float s1[8][2], s2[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j][0] = s1[j][1] = s2[j][0] = s2[j][1] = 0.f; }

    const int tiles_per_row = OW >> 4;
    const long long ntiles = (long long)N * OH * tiles_per_row;
    for (long long tile = (long long)blockIdx.x * THIN_WARPS + warp; tile < ntiles; tile += (long long)gridDim.x * THIN_WARPS) {
        const unsigned tu = (unsigned)tile, r = tu / (unsigned)tiles_per_row;      // 32-bit divisions (tiles < 2^31)
        const int xt = (int)(tu - r * (unsigned)tiles_per_row);
        const int n = (int)(r / (unsigned)OH), y = (int)(r - (unsigned)n * (unsigned)OH);
        const int x0 = xt << 4;
        const bf16* pa = in.at(n, y, x0 + g, 0);
        const bf16* pb = pa + 8 * in.sx;
        uint32_t a[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t lo_a = 0, hi_a = 0, lo_b = 0, hi_b = 0;
                if (off[s][h][0] >= 0) { lo_a = ldg_u16(pa + off[s][h][0]); lo_b = ldg_u16(pb + off[s][h][0]); }
                if (off[s][h][1] >= 0) { hi_a = ldg_u16(pa + off[s][h][1]); hi_b = ldg_u16(pb + off[s][h][1]); }
                a[s][2 * h] = lo_a | (hi_a << 16);          // row g
                a[s][2 * h + 1] = lo_b | (hi_b << 16);      // row g + 8
            }
        float c[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 bv = *reinterpret_cast<const float2*>(&sbias[8 * j + 2 * q]);
            c[j][0] = bv.x; c[j][1] = bv.y; c[j][2] = bv.x; c[j][3] = bv.y;
        }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int j = 0; j < 8; ++j) mma_bf16_16816(c[j], a[s], b0[s][j], b1[s][j]);
        // stage the 16 x 64 bf16 tile (16-byte chunks XOR-swizzled by row) and write whole 128-byte rows
        uint8_t* st = stage[warp];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s1[j][0] += c[j][0] + c[j][2]; s1[j][1] += c[j][1] + c[j][3];
            s2[j][0] = fmaf(c[j][0], c[j][0], fmaf(c[j][2], c[j][2], s2[j][0]));
            s2[j][1] = fmaf(c[j][1], c[j][1], fmaf(c[j][3], c[j][3], s2[j][1]));
            *reinterpret_cast<uint32_t*>(st + g * 128 + ((j ^ g) << 4) + 4 * q) = pack_bf16(c[j][0], c[j][1]);
            *reinterpret_cast<uint32_t*>(st + (g + 8) * 128 + ((j ^ g) << 4) + 4 * q) = pack_bf16(c[j][2], c[j][3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = lane + 32 * i, row = idx >> 3, ch = idx & 7;
            const uint4 v = *reinterpret_cast<const uint4*>(st + row * 128 + ((ch ^ (row & 7)) << 4));
            *reinterpret_cast<uint4*>(out.at(n, y, x0 + row, cb + ch * 8)) = v;
        }
    }
    if (stats) {
        // column (2q+e of n-tile j) totals over the 8 row groups g of the warp, then over the warps of the block
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float a = s1[j][e], b = s2[j][e];
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
                if (g == 0) { red[warp][0][8 * j + 2 * q + e] = a; red[warp][1][8 * j + 2 * q + e] = b; }
            }
        __syncthreads();
        if (threadIdx.x < 128) {
            const int w = threadIdx.x >> 6, ch = threadIdx.x & 63;
            float a = 0.f;
            for (int k = 0; k < THIN_WARPS; ++k) a += red[k][w][ch];
            atomicAdd(&stats[w * Cout + cb + ch], (double)a);
        }
    }
}

// dw[co][ci][t] += sum_pix P[pix][kk] dY[pix][co];  grid = (blocks over 16-pixel tiles, Cout / 32).
// 32 output channels per warp keep the accumulators at 32 registers, so 3 blocks (24 warps) per SM hide the gather latency
__global__ void __launch_bounds__(THIN_WARPS * 32, 3)
thin_mma_wgrad_k(View<bf16> in, View<bf16> dy, float* __restrict__ dw, int N, int OH, int OW, int Cin, int Cout) {
    __shared__ float red[32 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int KK = 9 * Cin;
    const int cb = blockIdx.y * 32;
    // A = P^T: rows kk in {g, g+8} + 16 mt, columns = pixels {2q, 2q+1, 2q+8, 2q+9}
    int off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) off[i] = patch_off(g + 8 * i, KK, Cin, in.sy, in.sx, in.sc);
    float c[2][4][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[m][j][0] = c[m][j][1] = c[m][j][2] = c[m][j][3] = 0.f; }

    const int tiles_per_row = OW >> 4;
    const long long ntiles = (long long)N * OH * tiles_per_row;
    for (long long tile = (long long)blockIdx.x * THIN_WARPS + warp; tile < ntiles; tile += (long long)gridDim.x * THIN_WARPS) {
        const unsigned tu = (unsigned)tile, r = tu / (unsigned)tiles_per_row;      // 32-bit divisions (tiles < 2^31)
        const int xt = (int)(tu - r * (unsigned)tiles_per_row);
        const int n = (int)(r / (unsigned)OH), y = (int)(r - (unsigned)n * (unsigned)OH);
        const int x0 = xt << 4;
        const bf16* px = in.at(n, y, x0 + 2 * q, 0);
        // gradient rows of the 4 pixels this thread contracts over: channel pairs (2g, 2g+1) of each 16-channel group
        const bf16* pd = dy.at(n, y, x0 + 2 * q, cb + 2 * g);
        uint32_t u[4][2];                                       // [pixel 2q, 2q+1, 2q+8, 2q+9][group]
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            const bf16* row = pd + ((pp & 1) + 8 * (pp >> 1)) * dy.sx;
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) u[pp][jj] = __ldg(reinterpret_cast<const uint32_t*>(row + 16 * jj));
        }
        uint32_t a[2][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                           // kk = g + 8 i  ->  m-tile i >> 1, row half i & 1
            uint32_t v0 = 0, v1 = 0, v8 = 0, v9 = 0;
            if (off[i] >= 0) {
                const bf16* p0 = px + off[i];
                v0 = ldg_u16(p0); v1 = ldg_u16(p0 + in.sx); v8 = ldg_u16(p0 + 8 * in.sx); v9 = ldg_u16(p0 + 9 * in.sx);
            }
            a[i >> 1][i & 1] = v0 | (v1 << 16);                 // k = 2q, 2q+1
            a[i >> 1][2 + (i & 1)] = v8 | (v9 << 16);           // k = 2q+8, 2q+9
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const uint32_t e0 = __byte_perm(u[0][jj], u[1][jj], 0x5410), o0 = __byte_perm(u[0][jj], u[1][jj], 0x7632);
            const uint32_t e1 = __byte_perm(u[2][jj], u[3][jj], 0x5410), o1 = __byte_perm(u[2][jj], u[3][jj], 0x7632);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                mma_bf16_16816(c[m][2 * jj], a[m], e0, e1);      // even channels 16 jj + 2 col
                mma_bf16_16816(c[m][2 * jj + 1], a[m], o0, o1);  // odd channels  16 jj + 2 col + 1
            }
        }
    }
    // block reduction in shared memory, then one atomic per weight
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kk = 16 * m + g + 8 * (e >> 1);
                const int col = 2 * q + (e & 1);
                const int co = 16 * (j >> 1) + 2 * col + (j & 1);
                atomicAdd(&red[kk * 32 + co], c[m][j][e]);
            }
    __syncthreads();
    for (int i = threadIdx.x; i < KK * 32; i += blockDim.x) {
        const int kk = i >> 5, co = cb + (i & 31);
        const int t = kk / Cin, ci = kk - t * Cin;
        atomicAdd(&dw[((long long)co * Cin + ci) * 9 + t], red[i]);
    }
}

static bool vec_ok(const kp_view* v, int align_elems) {
    return v->dtype == KP_BF16 && v->sc == 1 && (((uintptr_t)v->ptr) % (2 * align_elems)) == 0 && v->sx % align_elems == 0 &&
           v->sy % align_elems == 0 && v->sn % align_elems == 0;
}

static int thin_mma_grid(long long tiles) {
    long long blocks = (tiles + THIN_WARPS - 1) / THIN_WARPS;
    long long cap = (long long)kp_sm_count() * 4;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace

bool kp_thin_mma_fprop_ok(const kp_view* in, const kp_view* out, int OH, int OW, int IH, int IW, int Cin, int Cout, int ks,
                          int off) {
    return ks == 3 && off == 0 && Cin >= 1 && Cin <= 3 && Cout % 64 == 0 && OW % 16 == 0 && IH >= OH + 2 && IW >= OW + 2 &&
           in->dtype == KP_BF16 && vec_ok(out, 8);
}

int kp_thin_mma_fprop(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out,
                      double* stats, int N, int OH, int OW, int Cin, int Cout) {
    const long long tiles = (long long)N * OH * (OW / 16);
    dim3 grid((unsigned)thin_mma_grid(tiles), (unsigned)(Cout / 64), 1);
    thin_mma_fprop_k<<<grid, THIN_WARPS * 32, 0, st>>>(make_view<bf16>(in), wk, bias, make_view<bf16>(out), stats, N, OH, OW,
                                                       Cin, Cout);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

bool kp_thin_mma_wgrad_ok(const kp_view* x, const kp_view* dy, int W, int Cin, int Cout, int ks) {
    return ks == 3 && Cin >= 1 && Cin <= 3 && Cout % 64 == 0 && W % 16 == 0 && x->dtype == KP_BF16 && vec_ok(dy, 2);
}

int kp_thin_mma_wgrad(cudaStream_t st, const kp_view* x, const kp_view* dy, float* dw, int N, int H, int W, int Cin,
                      int Cout) {
    const long long tiles = (long long)N * H * (W / 16);
    long long blocks = (tiles + THIN_WARPS - 1) / THIN_WARPS;
    const long long cap = ((long long)kp_sm_count() * 3) / (Cout / 32);
    if (blocks > cap) blocks = cap > 0 ? cap : 1;
    dim3 grid((unsigned)blocks, (unsigned)(Cout / 32), 1);
    thin_mma_wgrad_k<<<grid, THIN_WARPS * 32, 0, st>>>(make_view<bf16>(x), make_view<bf16>(dy), dw, N, H, W, Cin, Cout);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

// ------------------------------------------------------------------------------------------------
// 1x1 output head with <= 4 outputs (decoder out_block 64 -> 3 at full resolution): three streaming kernels that move
// 128 bytes per pixel with 16-byte accesses (8 wide channels per thread).
//   fprop : y[pix][co]  = b[co] + sum_ci x[pix][ci] w[ci][co]
//   dgrad : dx[pix][ci] = sum_co dy[pix][co] w[co][ci]
//   wgrad : dw[co][ci] += sum_pix dy[pix][co] x[pix][ci]
// ------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
    const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(u[i] << 16); f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
}

struct PixIdx { int n, y, x; };
__device__ __forceinline__ PixIdx pix_of(long long p, int H, int W) {      // p < 2^31 (checked by kp_head1x1_ok): 32-bit divisions
    PixIdx r;
    const unsigned pu = (unsigned)p;
    const unsigned q = pu / (unsigned)W;
    r.x = (int)(pu - q * (unsigned)W);
    r.n = (int)(q / (unsigned)H);
    r.y = (int)(q - (unsigned)r.n * (unsigned)H);
    return r;
}

// groups of G = Cw / 8 lanes share a pixel; wk: fp32 [Cw][CT]
template <int CT, typename TO>
__global__ void __launch_bounds__(256)
head1x1_fprop_k(View<bf16> in, const float* __restrict__ wk, const float* __restrict__ bias, View<TO> out, int N, int H, int W,
                int Cw, int g_shift) {
    const int G = 1 << g_shift;
    const int gl = threadIdx.x & (G - 1);
    float w[8][CT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < CT; ++c) w[i][c] = wk[(gl * 8 + i) * CT + c];
    const long long P = (long long)N * H * W;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> g_shift;
    // every lane of a warp runs the same number of iterations (the shuffles below are warp-wide)
    const long long first = ((long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> g_shift;
    for (long long pb = first; pb < P; pb += stride) {
        const long long pr = pb + ((threadIdx.x & 31) >> g_shift);
        const bool valid = pr < P;
        const long long p = valid ? pr : P - 1;
        const PixIdx q = pix_of(p, H, W);
        float f[8], acc[CT];
        unpack8(*reinterpret_cast<const uint4*>(in.at(q.n, q.y, q.x, gl * 8)), f);
#pragma unroll
        for (int c = 0; c < CT; ++c) {
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) a = fmaf(f[i], w[i][c], a);
            for (int o = 1; o < G; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            acc[c] = a;
        }
        if (gl == 0 && valid) {
#pragma unroll
            for (int c = 0; c < CT; ++c) from_f(out.at(q.n, q.y, q.x, c), acc[c] + (bias ? bias[c] : 0.f));
        }
    }
}

// wd: fp32 [CT][Cw]
template <int CT>
__global__ void __launch_bounds__(256)
head1x1_dgrad_k(View<bf16> dy, const float* __restrict__ wd, View<bf16> dx, int N, int H, int W, int Cw, int g_shift) {
    const int G = 1 << g_shift;
    const int gl = threadIdx.x & (G - 1);
    float w[CT][8];
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) w[c][i] = wd[c * Cw + gl * 8 + i];
    const long long P = (long long)N * H * W;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> g_shift;
    for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> g_shift; p < P; p += stride) {
        const PixIdx q = pix_of(p, H, W);
        float g[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) g[c] = to_f(*dy.at(q.n, q.y, q.x, c));
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < CT; ++c) a = fmaf(g[c], w[c][i], a);
            o[i] = a;
        }
        uint4 u;
        u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
        *reinterpret_cast<uint4*>(dx.at(q.n, q.y, q.x, gl * 8)) = u;
    }
}

template <int CT>
__global__ void __launch_bounds__(256)
head1x1_wgrad_k(View<bf16> x, View<bf16> dy, float* __restrict__ dw, int N, int H, int W, int Cw, int g_shift) {
    __shared__ float red[CT * 512];
    const int G = 1 << g_shift;
    const int gl = threadIdx.x & (G - 1);
    float acc[CT][8];
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
    const long long P = (long long)N * H * W;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> g_shift;
    for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> g_shift; p < P; p += stride) {
        const PixIdx q = pix_of(p, H, W);
        float f[8], g[CT];
        unpack8(*reinterpret_cast<const uint4*>(x.at(q.n, q.y, q.x, gl * 8)), f);
#pragma unroll
        for (int c = 0; c < CT; ++c) g[c] = to_f(*dy.at(q.n, q.y, q.x, c));
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[c][i] = fmaf(g[c], f[i], acc[c][i]);
    }
    for (int i = threadIdx.x; i < CT * Cw; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    // lanes gl, gl + G, ... of a warp hold the same channels: combine them with shuffles before the shared-memory atomics
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = acc[c][i];
            for (int o = G; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) < G) atomicAdd(&red[c * Cw + gl * 8 + i], v);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < CT * Cw; i += blockDim.x) atomicAdd(&dw[i], red[i]);      // dw OIHW [co][ci] for a 1x1 conv
}

static int head_grid(long long P, int g_shift) {
    long long threads = P << g_shift;
    long long blocks = (threads + 255) / 256;
    const long long cap = (long long)kp_sm_count() * 8;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}
static int head_shift(int Cw) { int s = 0; while ((8 << s) < Cw) ++s; return s; }
static bool wide_ok(const kp_view* v, int Cw) {      // bf16, 16-byte aligned 8-channel chunks, 16 <= Cw <= 256 a power of two
    return v->dtype == KP_BF16 && v->sc == 1 && Cw >= 16 && Cw <= 256 && (Cw & (Cw - 1)) == 0 && (((uintptr_t)v->ptr) % 16) == 0 &&
           v->sx % 8 == 0 && v->sy % 8 == 0 && v->sn % 8 == 0;
}

// 1x1 conv from a wide bf16 activation to <= 32 outputs (decoder 64 -> 3 head at full resolution, 512 -> K keypoint
// heads) on mma.sync.m16n8k16: the lane-group kernels above spend ~440 instructions per 16 pixels on unpack / FMA / shuffle
// chains and are issue-bound at 1.3 TB/s; here 16 flat pixels are one M tile, A fragments are 4-byte loads straight from
// the NHWC rows, the (bf16-rounded, like every tensor-core layer) weights sit in shared memory in B-fragment order.
template <int NT, typename TO>
__global__ void __launch_bounds__(256)
head_mma_fprop_k(View<bf16> in, const float* __restrict__ wk /*[Cw][Ct]*/, const float* __restrict__ bias, View<TO> out, int N,
                 int H, int W, int Cw, int Ct) {
    extern __shared__ uint2 hm_wf[];                 // [Cw / 16][NT][32]
    const int KS = Cw >> 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    for (int i = threadIdx.x; i < KS * NT * 32; i += blockDim.x) {
        const int ln = i & 31, nt = (i >> 5) % NT, ks = (i >> 5) / NT;
        const int co = nt * 8 + (ln >> 2), k0 = ks * 16 + 2 * (ln & 3);
        uint2 v = make_uint2(0u, 0u);
        if (co < Ct) {
            v.x = pack_bf16(wk[(long long)k0 * Ct + co], wk[(long long)(k0 + 1) * Ct + co]);
            v.y = pack_bf16(wk[(long long)(k0 + 8) * Ct + co], wk[(long long)(k0 + 9) * Ct + co]);
        }
        hm_wf[i] = v;
    }
    float bv[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) { const int co = nt * 8 + 2 * q + e; bv[nt][e] = (bias && co < Ct) ? bias[co] : 0.f; }
    __syncthreads();
    const unsigned P = (unsigned)N * (unsigned)H * (unsigned)W, ntiles = (P + 15) >> 4;
    for (unsigned tile = blockIdx.x * 8 + warp; tile < ntiles; tile += gridDim.x * 8) {
        const unsigned p0 = tile * 16 + g, p1 = p0 + 8;
        const bool va = p0 < P, vb = p1 < P;
        const PixIdx qa = pix_of(va ? p0 : P - 1, H, W), qb = pix_of(vb ? p1 : P - 1, H, W);
        const bf16* pa = in.at(qa.n, qa.y, qa.x, 2 * q);
        const bf16* pb = in.at(qb.n, qb.y, qb.x, 2 * q);
        float c[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { c[nt][0] = bv[nt][0]; c[nt][1] = bv[nt][1]; c[nt][2] = bv[nt][0]; c[nt][3] = bv[nt][1]; }
#pragma unroll 4
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t a[4];
            a[0] = *reinterpret_cast<const uint32_t*>(pa + 16 * ks);
            a[1] = *reinterpret_cast<const uint32_t*>(pb + 16 * ks);
            a[2] = *reinterpret_cast<const uint32_t*>(pa + 16 * ks + 8);
            a[3] = *reinterpret_cast<const uint32_t*>(pb + 16 * ks + 8);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint2 w = hm_wf[(ks * NT + nt) * 32 + lane];
                mma_bf16_16816(c[nt], a, w.x, w.y);
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int co = nt * 8 + 2 * q + e;
                if (co < Ct) {
                    if (va) from_f(out.at(qa.n, qa.y, qa.x, co), c[nt][e]);
                    if (vb) from_f(out.at(qb.n, qb.y, qb.x, co), c[nt][2 + e]);
                }
            }
    }
}

}  // namespace

bool kp_head1x1_ok(const kp_view* wide, const kp_view* thin, int Cw, int Ct) {      // callers pass images of < 2^31 pixels
    return Ct >= 1 && Ct <= 4 && wide_ok(wide, Cw) && (thin->dtype == KP_BF16 || thin->dtype == KP_F32);
}

#define KP_HEAD_CT(CTV, CALL) \
    do { if ((CTV) == 1) { CALL(1); } else if ((CTV) == 2) { CALL(2); } else if ((CTV) == 3) { CALL(3); } else { CALL(4); } } while (0)

int kp_head1x1_fprop(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out, int N, int H,
                     int W, int Cw, int Ct) {
    const int sh = head_shift(Cw), grid = head_grid((long long)N * H * W, sh);
    if (out->dtype == KP_BF16) {
#define KP_C(CT) head1x1_fprop_k<CT, bf16><<<grid, 256, 0, st>>>(make_view<bf16>(in), wk, bias, make_view<bf16>(out), N, H, W, Cw, sh)
        KP_HEAD_CT(Ct, KP_C);
#undef KP_C
    } else {
#define KP_C(CT) head1x1_fprop_k<CT, float><<<grid, 256, 0, st>>>(make_view<bf16>(in), wk, bias, make_view<float>(out), N, H, W, Cw, sh)
        KP_HEAD_CT(Ct, KP_C);
#undef KP_C
    }
    KP_LAUNCH_CHECK();
    return KP_OK;
}

int kp_head1x1_dgrad(cudaStream_t st, const kp_view* dy, const float* wd, const kp_view* dx, int N, int H, int W, int Cw, int Ct) {
    const int sh = head_shift(Cw), grid = head_grid((long long)N * H * W, sh);
#define KP_C(CT) head1x1_dgrad_k<CT><<<grid, 256, 0, st>>>(make_view<bf16>(dy), wd, make_view<bf16>(dx), N, H, W, Cw, sh)
    KP_HEAD_CT(Ct, KP_C);
#undef KP_C
    KP_LAUNCH_CHECK();
    return KP_OK;
}

int kp_head1x1_wgrad(cudaStream_t st, const kp_view* x, const kp_view* dy, float* dw, int N, int H, int W, int Cw, int Ct) {
    const int sh = head_shift(Cw);
    long long blocks = ((((long long)N * H * W) << sh) + 255) / 256;
    const long long cap = (long long)kp_sm_count() * 4;
    if (blocks > cap) blocks = cap;
#define KP_C(CT) head1x1_wgrad_k<CT><<<(int)blocks, 256, 0, st>>>(make_view<bf16>(x), make_view<bf16>(dy), dw, N, H, W, Cw, sh)
    KP_HEAD_CT(Ct, KP_C);
#undef KP_C
    KP_LAUNCH_CHECK();
    return KP_OK;
}

bool kp_head_mma_fprop_ok(const kp_view* in, const kp_view* out, int N, int H, int W, int Cw, int Ct) {
    static int on = -1;
    if (on < 0) { const char* e = kp_env("KP_HEAD_MMA"); on = e ? atoi(e) : 1; }
    return on && in->dtype == KP_BF16 && in->sc == 1 && (((uintptr_t)in->ptr) % 4) == 0 && in->sx % 2 == 0 && in->sy % 2 == 0 &&
           in->sn % 2 == 0 && Cw % 16 == 0 && Cw >= 16 && Cw <= 512 && Ct >= 1 && Ct <= 32 &&
           (out->dtype == KP_BF16 || out->dtype == KP_F32) && (long long)N * H * W < (1LL << 31) - 16;
}

int kp_head_mma_fprop(cudaStream_t st, const kp_view* in, const float* wk, const float* bias, const kp_view* out, int N, int H,
                      int W, int Cw, int Ct) {
    const long long tiles = ((long long)N * H * W + 15) / 16;
    long long blocks = (tiles + 7) / 8;
    const long long cap = (long long)kp_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const int NT = Ct <= 8 ? 1 : (Ct <= 16 ? 2 : 4);
    const size_t smem = (size_t)(Cw / 16) * NT * 32 * sizeof(uint2);
#define KP_HM(NTV, TOV) head_mma_fprop_k<NTV, TOV><<<(int)blocks, 256, smem, st>>>(make_view<bf16>(in), wk, bias, make_view<TOV>(out), N, H, W, Cw, Ct)
    if (out->dtype == KP_BF16) { if (NT == 1) KP_HM(1, bf16); else if (NT == 2) KP_HM(2, bf16); else KP_HM(4, bf16); }
    else { if (NT == 1) KP_HM(1, float); else if (NT == 2) KP_HM(2, float); else KP_HM(4, float); }
#undef KP_HM
    KP_LAUNCH_CHECK();
    return KP_OK;
}
