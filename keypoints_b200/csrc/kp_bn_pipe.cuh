// Bulk-async pipelined row kernels for the bf16 BatchNorm / activation passes (throughput mode).
// A register-staged streaming kernel needs its in-flight bytes in registers, and these passes also keep 16-40
// per-channel constants / accumulators per thread, so they top out near 60-70 % of HBM bandwidth at 16-24 warps per SM.
// Here one producer thread per CTA streams contiguous row chunks (8 KB per operand) into a 4-6 stage shared-memory ring
// with cp.async.bulk + mbarrier transaction counts (the TMA engine; no tensor map is needed for 1-D spans), so ~100 KB per
// SM is in flight regardless of register pressure; 8 consumer warps read the stage with 16-byte LDS, do the math and write
// their results with coalesced 16-byte global stores.  Requirements (checked by the dispatcher): bf16, dense rows
// (pixel stride == C), C/8 a power of two <= 256 and W*C/8 a multiple of 512 (two items per thread per chunk).
// Included by kp_elementwise.cu after kp_bn_lean.cuh.
#pragma once

namespace pipe {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int CONSUMERS = 256;
constexpr int THREADS = CONSUMERS + 32;
constexpr int IPC = 512;                 // items (8 channels = 16 bytes) per chunk
constexpr int CHUNK_BYTES = IPC * 16;    // 8 KB
__device__ __forceinline__ uint4 lds16(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
}  // namespace pipe

// Position of a CTA in the (image, row, chunk) unit space, advanced by gridDim.x units per step without divisions.
struct Cursor {
    int n, yy, ch;                     // n = image the unit belongs to (reversed order when nrev > 0)
    int nf, nrev;
    int dn, dyy, dch, R, cpr;
    __device__ __forceinline__ void init(int u0, int stride, int rows_per_image, int chunks_per_row, int n_rev) {
        R = rows_per_image; cpr = chunks_per_row; nrev = n_rev;
        ch = u0 % cpr; int row = u0 / cpr; yy = row % R; nf = row / R;
        dch = stride % cpr; int drow = stride / cpr; dyy = drow % R; dn = drow / R;
        n = nrev > 0 ? nrev - 1 - nf : nf;
    }
    __device__ __forceinline__ void next() {
        ch += dch;
        int c = ch >= cpr ? 1 : 0;
        ch -= c ? cpr : 0;
        yy += dyy + c;
        c = yy >= R ? 1 : 0;
        yy -= c ? R : 0;
        nf += dn + c;
        n = nrev > 0 ? nrev - 1 - nf : nf;
    }
};

// n_rev = 0: images first to last; n_rev = N: last to first.  A pass that follows a producer sweeping the tensor front to
// back (a conv, or the previous BatchNorm pass) starts where the producer ended, i.e. on the ~100 MB still in the L2.
// Shared skeleton: STAGES-deep ring of STAGE_BYTES.  `issue(cur, stage, full)` (one producer thread) posts the expected
// byte count and the bulk copies of the unit at cursor `cur`; `body(cur, stage)` consumes one unit (256 consumer threads).
template <int STAGE_BYTES, int STAGES, class IssueFn, class BodyFn>
__device__ __forceinline__ void pipe_run(uint8_t* smem, int units, int rows_per_image, int cpr, int n_rev, IssueFn issue,
                                         BodyFn body) {
    using namespace pipe;
    const uint32_t bars = s32(smem + STAGES * STAGE_BYTES);      // full[STAGES], empty[STAGES]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), CONSUMERS / 32);
        }
        fence_init();
    }
    __syncthreads();
    const int mine = ((int)blockIdx.x < units) ? (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    Cursor cur;
    cur.init(blockIdx.x, gridDim.x, rows_per_image, cpr, n_rev);
    if (warp == CONSUMERS / 32) {
        if (lane == 0) {
            int s = 0, ph = 0;
            for (int it = 0; it < mine; ++it) {
                mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                issue(cur, s32(smem + s * STAGE_BYTES), bars + 8 * s);
                cur.next();
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        int s = 0, ph = 0;
        for (int it = 0; it < mine; ++it) {
            mbar_wait(bars + 8 * s, ph);
            body(cur, smem + s * STAGE_BYTES);
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (STAGES + s));
            cur.next();
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    }
}

template <int STAGE_BYTES, int STAGES>
constexpr int pipe_smem_bytes() { return STAGES * STAGE_BYTES + 16 * STAGES + 16; }
constexpr int PIPE_FWD_STAGE = pipe::CHUNK_BYTES, PIPE_FWD_STAGES = 6;
constexpr int PIPE_EXT = 1024;                                   // one pixel of <= 512 channels either side of a dout chunk
constexpr int PIPE_BWD_STAGE = 2 * pipe::CHUNK_BYTES + 2 * PIPE_EXT, PIPE_BWD_STAGES = 4;
constexpr int PIPE_APPLY_STAGE = 2 * pipe::CHUNK_BYTES, PIPE_APPLY_STAGES = 4;

// ------------------------------------------------------------------------------------------------
// forward, post = none: out[py][px] = act(scale * y[clamp(py - pad)][clamp(px - pad)] + shift)
// unit = (output row, chunk of the source row); edge pixels also write their replicate-border copies
// ------------------------------------------------------------------------------------------------
template <int ACT, int STAGES = PIPE_FWD_STAGES>
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_fwd_none_pipe_k(Rows<const bf16> y, Rows<bf16> out, const float* __restrict__ scale, const float* __restrict__ shift,
                   int pad, int N, int H, int W, int C, int cg_shift, int cpr, const BnFuse fuse, int rev) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int ncg = C >> 3;
    const int c0 = (tid & (ncg - 1)) * 8;
    float2 sc[4], sh[4];
    if (tid < pipe::CONSUMERS) {
        if (fuse.stats) {
            float a[8], b[8];
            fused_affine(fuse, C, c0, blockIdx.x == 0 && (tid >> cg_shift) == 0, a, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) { sc[i] = make_float2(a[2 * i], a[2 * i + 1]); sh[i] = make_float2(b[2 * i], b[2 * i + 1]); }
        } else {
            load_c8(scale, c0, 1.f, sc);
            load_c8(shift, c0, 0.f, sh);
        }
    }
    const int PH = H + 2 * pad;
    const int units = N * PH * cpr;
    auto issue = [&](const Cursor& cu, uint32_t stage, uint32_t full) {
        const int oy = min(max(cu.yy - pad, 0), H - 1);
        pipe::mbar_expect_tx(full, pipe::CHUNK_BYTES);
        pipe::bulk_g2s(stage, y.row(cu.n, oy) + cu.ch * (pipe::IPC * 8), pipe::CHUNK_BYTES, full);
    };
    auto body = [&](const Cursor& cu, const uint8_t* st) {
        const int ch = cu.ch;
        bf16* orow = out.row(cu.n, cu.yy) + pad * out.sx;
        uint4 raw[pipe::IPC / pipe::CONSUMERS];
#pragma unroll
        for (int j = 0; j < pipe::IPC / pipe::CONSUMERS; ++j) raw[j] = pipe::lds16(st + (tid + j * pipe::CONSUMERS) * 16);
#pragma unroll
        for (int j = 0; j < pipe::IPC / pipe::CONSUMERS; ++j) {
            const int item = tid + j * pipe::CONSUMERS;
            float2 v[4];
            P8<bf16>::up(raw[j], v);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 z = fma2(v[i], sc[i], sh[i]);
                v[i] = make_float2(actv<ACT>(z.x), actv<ACT>(z.y));
            }
            const int e = (ch * pipe::IPC + item) * 8;          // element offset inside the source row = x * C + c
            const int xx = (ch * pipe::IPC + item) >> cg_shift;
            const uint4 o = make_uint4(P8<bf16>::pk(v[0]), P8<bf16>::pk(v[1]), P8<bf16>::pk(v[2]), P8<bf16>::pk(v[3]));
            *reinterpret_cast<uint4*>(orow + e) = o;
            if (pad) {
                if (xx == 0) *reinterpret_cast<uint4*>(orow + e - C) = o;
                if (xx == W - 1) *reinterpret_cast<uint4*>(orow + e + C) = o;
            }
        }
    };
    pipe_run<PIPE_FWD_STAGE, STAGES>(smem, units, PH, cpr, rev, issue, body);
}

// ------------------------------------------------------------------------------------------------
// backward pass 1, post = none: dz = act'(z) * fold(dout) -> dy, per-channel sums of dz and dz * xhat
// ------------------------------------------------------------------------------------------------
// MODE: PASS1_WRITE = sums + dz written to dy; PASS1_SUMS = sums only (dz is recomputed by pass 2);
//       PASS2_GATHER = the gather is repeated and dy = scale (dz - mean(dz) - xhat mean(dz xhat)) is written, so dz never
//       round-trips through HBM (10 instead of 12 bytes per element over the two passes).
enum { PASS1_WRITE = 0, PASS1_SUMS = 1, PASS2_GATHER = 2 };

// per-channel constants of the two backward passes; in PASS2_GATHER s1 / s2 hold -sc*mean(dz) / -sc*invstd*mean(dz*xhat)
__device__ __forceinline__ void bwd_consts(int MODE, const float* scale, const float* shift, const float* mean,
                                           const float* invstd, const double* sums, double count, int C, int c0,
                                           float2 (&sc)[4], float2 (&sh)[4], float2 (&nmu)[4], float2 (&s1)[4], float2 (&s2)[4]) {
    load_c8(scale, c0, 1.f, sc);
    load_c8(shift, c0, 0.f, sh);
    load_c8(mean, c0, 0.f, nmu);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        nmu[i] = make_float2(-nmu[i].x, -nmu[i].y);
        if (MODE == PASS2_GATHER) {
            const int c = c0 + 2 * i;
            const double rc = 1.0 / count;
            const float m1x = (float)(sums[c] * rc), m1y = (float)(sums[c + 1] * rc);
            const float m2x = (float)(sums[C + c] * rc), m2y = (float)(sums[C + c + 1] * rc);
            s1[i] = make_float2(-sc[i].x * m1x, -sc[i].y * m1y);
            s2[i] = make_float2(-sc[i].x * invstd[c] * m2x, -sc[i].y * invstd[c + 1] * m2y);
        } else {
            s1[i] = make_float2(0.f, 0.f); s2[i] = make_float2(0.f, 0.f);
        }
    }
}

template <int ACT, int MODE>
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_bwd_none_pipe_k(Rows<const bf16> dout, Rows<const bf16> y, Rows<bf16> dy, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                   double* sums, double count, int pad, int N, int H, int W, int C, int cg_shift, int cpr, float* dgamma,
                   float* dbeta, int rev) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int ncg = C >> 3;
    const int c0 = (tid & (ncg - 1)) * 8;
    float2 sc[4], sh[4], nmu[4], s1[4], s2[4];
    bwd_consts(MODE, scale, shift, mean, invstd, sums, count, C, c0, sc, sh, nmu, s1, s2);
    if (MODE == PASS2_GATHER && blockIdx.x == 0 && tid < pipe::CONSUMERS && (tid >> cg_shift) == 0) {     // fused kp_bn_grad_finalize
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (dbeta) dbeta[c0 + i] = (float)sums[c0 + i];
            if (dgamma) dgamma[c0 + i] = (float)sums[C + c0 + i];
        }
    }
    const int units = N * H * cpr;
    // stage layout: [ext | dout chunk | ext | y chunk]; with a replicate-padded dout the chunk is loaded together with
    // the pixel before and after it, so the left / right border copies of the edge pixels come from shared memory
    const int ext = pad ? C * 2 : 0;                            // bytes of one pixel
    auto issue = [&](const Cursor& cu, uint32_t stage, uint32_t full) {
        const bf16* d = dout.row(cu.n, cu.yy + pad) + pad * dout.sx + cu.ch * (pipe::IPC * 8);
        pipe::mbar_expect_tx(full, 2 * pipe::CHUNK_BYTES + 2 * ext);
        pipe::bulk_g2s(stage + PIPE_EXT - ext, reinterpret_cast<const uint8_t*>(d) - ext, pipe::CHUNK_BYTES + 2 * ext, full);
        pipe::bulk_g2s(stage + 2 * PIPE_EXT + pipe::CHUNK_BYTES, y.row(cu.n, cu.yy) + cu.ch * (pipe::IPC * 8),
                       pipe::CHUNK_BYTES, full);
    };
    auto body = [&](const Cursor& cu, const uint8_t* st) {
        const int n = cu.n, yy = cu.yy, ch = cu.ch;
        bf16* orow = MODE == PASS1_SUMS ? nullptr : dy.row(n, yy);
        const bool rowb = pad && (yy == 0 || yy == H - 1);
        const uint8_t* sd = st + PIPE_EXT;
        const uint8_t* sy = st + 2 * PIPE_EXT + pipe::CHUNK_BYTES;
        uint4 rawg[pipe::IPC / pipe::CONSUMERS], rawy[pipe::IPC / pipe::CONSUMERS];
#pragma unroll
        for (int j = 0; j < pipe::IPC / pipe::CONSUMERS; ++j) {
            rawg[j] = pipe::lds16(sd + (tid + j * pipe::CONSUMERS) * 16);
            rawy[j] = pipe::lds16(sy + (tid + j * pipe::CONSUMERS) * 16);
        }
#pragma unroll
        for (int j = 0; j < pipe::IPC / pipe::CONSUMERS; ++j) {
            const int item = tid + j * pipe::CONSUMERS;
            float2 g[4], yv[4];
            P8<bf16>::up(rawg[j], g);
            P8<bf16>::up(rawy[j], yv);
            const int e = (ch * pipe::IPC + item) * 8;
            if (pad) {
                const int xx = (ch * pipe::IPC + item) >> cg_shift;
                if (rowb) {                                      // first / last image row: rare, gathered from global memory
                    add_fold<bf16>(dout, n, yy, xx, H, W, c0, g);
                } else if (xx == 0 || xx == W - 1) {
                    float2 t[4];
                    P8<bf16>::up(pipe::lds16(sd + item * 16 + (xx == 0 ? -ext : ext)), t);
#pragma unroll
                    for (int i = 0; i < 4; ++i) g[i] = add2(g[i], t[i]);
                    if (W == 1) {
                        P8<bf16>::up(pipe::lds16(sd + item * 16 + ext), t);
#pragma unroll
                        for (int i = 0; i < 4; ++i) g[i] = add2(g[i], t[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 z = fma2(yv[i], sc[i], sh[i]);
                const float2 dz = make_float2(actg<ACT>(z.x, g[i].x), actg<ACT>(z.y, g[i].y));
                const float2 yc = add2(yv[i], nmu[i]);
                if (MODE == PASS2_GATHER) {
                    g[i] = fma2(sc[i], dz, fma2(s2[i], yc, s1[i]));
                } else {
                    g[i] = dz;
                    s1[i] = add2(s1[i], dz);
                    s2[i] = fma2(dz, yc, s2[i]);
                }
            }
            if (MODE != PASS1_SUMS) P8<bf16>::st(orow + e, g);
        }
    };
    pipe_run<PIPE_BWD_STAGE, PIPE_BWD_STAGES>(smem, units, H, cpr, rev, issue, body);
    if (tid >= pipe::CONSUMERS || MODE == PASS2_GATHER) return;
    // every unit of this CTA has been consumed: the ring is free, reuse it for the block reduction
    pipe::consumer_sync();
    float* red = reinterpret_cast<float*>(smem);               // [2][256 * 8]
    float2 is[4];
    load_c8(invstd, c0, 1.f, is);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        red[tid * 8 + 2 * i] = s1[i].x; red[tid * 8 + 2 * i + 1] = s1[i].y;
        red[2048 + tid * 8 + 2 * i] = s2[i].x * is[i].x; red[2048 + tid * 8 + 2 * i + 1] = s2[i].y * is[i].y;
    }
    pipe::consumer_sync();
    for (int ch = tid; ch < C; ch += pipe::CONSUMERS) {
        const int g8 = ch >> 3, i = ch & 7;
        float a = 0.f, b = 0.f;
        for (int t = g8; t < pipe::CONSUMERS; t += ncg) { a += red[t * 8 + i]; b += red[2048 + t * 8 + i]; }
        atomicAdd(&sums[ch], (double)a);
        atomicAdd(&sums[C + ch], (double)b);
    }
}

// ------------------------------------------------------------------------------------------------
// backward pass 2 (in place): dy = scale * (dz - mean(dz) - xhat * mean(dz * xhat)) = a0*dz + a1*y + a2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_bwd_apply_pipe_k(Rows<const bf16> y, Rows<bf16> dy, const float* __restrict__ scale, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const double* __restrict__ sums, double count, int N, int H, int W,
                    int C, int cg_shift, int cpr, float* dgamma, float* dbeta) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int ncg = C >> 3;
    const int c0 = (tid & (ncg - 1)) * 8;
    float2 a0[4], a1[4], a2[4];
    if (tid < pipe::CONSUMERS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + i;
            const float sc = scale[c], mu = mean[c], is = invstd[c];
            const double rc = 1.0 / count;
            const float m1 = (float)(sums[c] * rc), m2 = (float)(sums[C + c] * rc);
            const float v0 = sc, v1 = -sc * is * m2, v2 = -sc * m1 + sc * is * m2 * mu;
            if (i & 1) { a0[i >> 1].y = v0; a1[i >> 1].y = v1; a2[i >> 1].y = v2; }
            else { a0[i >> 1].x = v0; a1[i >> 1].x = v1; a2[i >> 1].x = v2; }
            if (blockIdx.x == 0 && (tid >> cg_shift) == 0) {           // fused kp_bn_grad_finalize
                if (dbeta) dbeta[c] = (float)sums[c];
                if (dgamma) dgamma[c] = (float)sums[C + c];
            }
        }
    }
    const int units = N * H * cpr;
    auto issue = [&](const Cursor& cu, uint32_t stage, uint32_t full) {
        pipe::mbar_expect_tx(full, 2 * pipe::CHUNK_BYTES);
        pipe::bulk_g2s(stage, dy.row(cu.n, cu.yy) + cu.ch * (pipe::IPC * 8), pipe::CHUNK_BYTES, full);
        pipe::bulk_g2s(stage + pipe::CHUNK_BYTES, y.row(cu.n, cu.yy) + cu.ch * (pipe::IPC * 8), pipe::CHUNK_BYTES, full);
    };
    auto body = [&](const Cursor& cu, const uint8_t* st) {
        bf16* orow = dy.row(cu.n, cu.yy) + cu.ch * (pipe::IPC * 8);
        uint4 rawg[pipe::IPC / pipe::CONSUMERS], rawy[pipe::IPC / pipe::CONSUMERS];
#pragma unroll
        for (int j = 0; j < pipe::IPC / pipe::CONSUMERS; ++j) {
            rawg[j] = pipe::lds16(st + (tid + j * pipe::CONSUMERS) * 16);
            rawy[j] = pipe::lds16(st + pipe::CHUNK_BYTES + (tid + j * pipe::CONSUMERS) * 16);
        }
#pragma unroll
        for (int j = 0; j < pipe::IPC / pipe::CONSUMERS; ++j) {
            const int item = tid + j * pipe::CONSUMERS;
            float2 g[4], yv[4];
            P8<bf16>::up(rawg[j], g);
            P8<bf16>::up(rawy[j], yv);
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] = fma2(a0[i], g[i], fma2(a1[i], yv[i], a2[i]));
            P8<bf16>::st(orow + item * 8, g);
        }
    };
    pipe_run<PIPE_APPLY_STAGE, PIPE_APPLY_STAGES>(smem, units, H, cpr, 0, issue, body);
}

// ------------------------------------------------------------------------------------------------
// 2x2 max-pool variants.  A unit is one output row x 256 output items (pixel, channel group): the two source rows of
// y arrive as 2 x 8 KB spans (the 2x2 windows of a run of output pixels are contiguous in each source row).
// ------------------------------------------------------------------------------------------------
constexpr int PIPE_POOL_ITEMS = 256;
constexpr int PIPE_POOL_YBYTES = 2 * PIPE_POOL_ITEMS * 16;                  // one source row span
constexpr int PIPE_FPOOL_STAGE = 2 * PIPE_POOL_YBYTES, PIPE_FPOOL_STAGES = 5;
constexpr int PIPE_BPOOL_STAGE = 2 * PIPE_POOL_YBYTES + PIPE_POOL_ITEMS * 16 + 2 * PIPE_EXT, PIPE_BPOOL_STAGES = 4;

// forward: out[py][px] = max over the window of act(scale * y + shift); H, W even; OH = H/2, OW = W/2
template <int ACT>
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_fwd_pool_pipe_k(Rows<const bf16> y, Rows<bf16> out, const float* __restrict__ scale, const float* __restrict__ shift,
                   int pad, int N, int OH, int OW, int C, int cg_shift, int cpr, const BnFuse fuse, int rev) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int ncg = C >> 3;
    const int cg = tid & (ncg - 1), c0 = cg * 8;
    float2 sc[4], sh[4];
    if (tid < pipe::CONSUMERS) {
        if (fuse.stats) {
            float a[8], b[8];
            fused_affine(fuse, C, c0, blockIdx.x == 0 && (tid >> cg_shift) == 0, a, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) { sc[i] = make_float2(a[2 * i], a[2 * i + 1]); sh[i] = make_float2(b[2 * i], b[2 * i + 1]); }
        } else {
            load_c8(scale, c0, 1.f, sc);
            load_c8(shift, c0, 0.f, sh);
        }
    }
    const int PH = OH + 2 * pad;
    const int units = N * PH * cpr;
    auto issue = [&](const Cursor& cu, uint32_t stage, uint32_t full) {
        const int oy = min(max(cu.yy - pad, 0), OH - 1);
        const bf16* r0 = y.row(cu.n, 2 * oy) + cu.ch * (2 * PIPE_POOL_ITEMS * 8);
        pipe::mbar_expect_tx(full, 2 * PIPE_POOL_YBYTES);
        pipe::bulk_g2s(stage, r0, PIPE_POOL_YBYTES, full);
        pipe::bulk_g2s(stage + PIPE_POOL_YBYTES, r0 + y.sy, PIPE_POOL_YBYTES, full);
    };
    const int pl = tid >> cg_shift;                              // output pixel of this thread inside the chunk
    const int yoff = ((2 * pl) << cg_shift | cg) * 16;           // its window's left pixel inside a source row span
    auto body = [&](const Cursor& cu, const uint8_t* st) {
        const uint4 r00 = pipe::lds16(st + yoff), r01 = pipe::lds16(st + yoff + ncg * 16);
        const uint4 r10 = pipe::lds16(st + PIPE_POOL_YBYTES + yoff), r11 = pipe::lds16(st + PIPE_POOL_YBYTES + yoff + ncg * 16);
        // the activation is monotone: max over the window of act(z) = act(max z)
        float2 v[4], a[4];
        auto z8 = [&](const uint4& r, float2 (&o)[4]) {
            P8<bf16>::up(r, o);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = fma2(o[i], sc[i], sh[i]);
        };
        z8(r00, v);
        z8(r01, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float2(fmaxf(v[i].x, a[i].x), fmaxf(v[i].y, a[i].y));
        z8(r10, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float2(fmaxf(v[i].x, a[i].x), fmaxf(v[i].y, a[i].y));
        z8(r11, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = make_float2(fmaxf(v[i].x, a[i].x), fmaxf(v[i].y, a[i].y));
            v[i] = make_float2(actv<ACT>(v[i].x), actv<ACT>(v[i].y));
        }
        const int ox = cu.ch * (PIPE_POOL_ITEMS >> cg_shift) + pl;
        bf16* dst = out.row(cu.n, cu.yy) + (ox + pad) * C + c0;
        const uint4 o = make_uint4(P8<bf16>::pk(v[0]), P8<bf16>::pk(v[1]), P8<bf16>::pk(v[2]), P8<bf16>::pk(v[3]));
        *reinterpret_cast<uint4*>(dst) = o;
        if (pad) {
            if (ox == 0) *reinterpret_cast<uint4*>(dst - C) = o;
            if (ox == OW - 1) *reinterpret_cast<uint4*>(dst + C) = o;
        }
    };
    pipe_run<PIPE_FPOOL_STAGE, PIPE_FPOOL_STAGES>(smem, units, PH, cpr, rev, issue, body);
}

// backward pass 1: the gradient of a pooled pixel goes to the first maximum of its window (row-major), zeros elsewhere
template <int ACT, int MODE>
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_bwd_pool_pipe_k(Rows<const bf16> dout, Rows<const bf16> y, Rows<bf16> dy, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                   double* sums, double count, int pad, int N, int OH, int OW, int C, int cg_shift, int cpr, float* dgamma,
                   float* dbeta, int rev) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int ncg = C >> 3;
    const int cg = tid & (ncg - 1), c0 = cg * 8;
    float2 sc[4], sh[4], nmu[4], s1[4], s2[4];
    bwd_consts(MODE, scale, shift, mean, invstd, sums, count, C, c0, sc, sh, nmu, s1, s2);
    if (MODE == PASS2_GATHER && blockIdx.x == 0 && tid < pipe::CONSUMERS && (tid >> cg_shift) == 0) {     // fused kp_bn_grad_finalize
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (dbeta) dbeta[c0 + i] = (float)sums[c0 + i];
            if (dgamma) dgamma[c0 + i] = (float)sums[C + c0 + i];
        }
    }
    const int units = N * OH * cpr;
    const int ext = pad ? C * 2 : 0;
    constexpr int DOFF = 2 * PIPE_POOL_YBYTES + PIPE_EXT;       // dout chunk inside the stage
    auto issue = [&](const Cursor& cu, uint32_t stage, uint32_t full) {
        const bf16* r0 = y.row(cu.n, 2 * cu.yy) + cu.ch * (2 * PIPE_POOL_ITEMS * 8);
        const bf16* d = dout.row(cu.n, cu.yy + pad) + pad * dout.sx + cu.ch * (PIPE_POOL_ITEMS * 8);
        pipe::mbar_expect_tx(full, 2 * PIPE_POOL_YBYTES + PIPE_POOL_ITEMS * 16 + 2 * ext);
        pipe::bulk_g2s(stage, r0, PIPE_POOL_YBYTES, full);
        pipe::bulk_g2s(stage + PIPE_POOL_YBYTES, r0 + y.sy, PIPE_POOL_YBYTES, full);
        pipe::bulk_g2s(stage + DOFF - ext, reinterpret_cast<const uint8_t*>(d) - ext, PIPE_POOL_ITEMS * 16 + 2 * ext, full);
    };
    const int pl = tid >> cg_shift;
    const int yoff = ((2 * pl) << cg_shift | cg) * 16;
    auto body = [&](const Cursor& cu, const uint8_t* st) {
        uint4 ry[4];
        ry[0] = pipe::lds16(st + yoff); ry[1] = pipe::lds16(st + yoff + ncg * 16);
        ry[2] = pipe::lds16(st + PIPE_POOL_YBYTES + yoff); ry[3] = pipe::lds16(st + PIPE_POOL_YBYTES + yoff + ncg * 16);
        float2 t[4];
        P8<bf16>::up(pipe::lds16(st + DOFF + tid * 16), t);
        const int oy = cu.yy;
        const int ox = cu.ch * (PIPE_POOL_ITEMS >> cg_shift) + pl;
        if (pad) {
            if (oy == 0 || oy == OH - 1) {
                add_fold<bf16>(dout, cu.n, oy, ox, OH, OW, c0, t);
            } else if (ox == 0 || ox == OW - 1) {
                float2 e[4];
                P8<bf16>::up(pipe::lds16(st + DOFF + tid * 16 + (ox == 0 ? -ext : ext)), e);
#pragma unroll
                for (int i = 0; i < 4; ++i) t[i] = add2(t[i], e[i]);
            }
        }
        // arg-max of the window per channel.  The activation is monotone, so the first maximum of z is the first maximum of
        // act(z) wherever the routed gradient is non-zero (ties of a ReLU at 0 route a zero gradient either way).
        float2 zb[4], yb[4];
        int bi[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 yq[4];
            P8<bf16>::up(ry[q], yq);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 z = fma2(yq[i], sc[i], sh[i]);
                if (q == 0 || z.x > zb[i].x) { zb[i].x = z.x; yb[i].x = yq[i].x; bi[2 * i] = q; }
                if (q == 0 || z.y > zb[i].y) { zb[i].y = z.y; yb[i].y = yq[i].y; bi[2 * i + 1] = q; }
            }
        }
        float2 dzw[4];                                              // gradient at the winner; every other window pixel gets 0
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dzw[i] = make_float2(actg<ACT>(zb[i].x, t[i].x), actg<ACT>(zb[i].y, t[i].y));
            if (MODE == PASS2_GATHER) {
                dzw[i] = make_float2(sc[i].x * dzw[i].x, sc[i].y * dzw[i].y);
            } else {
                s1[i] = add2(s1[i], dzw[i]);
                s2[i] = fma2(dzw[i], add2(yb[i], nmu[i]), s2[i]);
            }
        }
        if (MODE != PASS1_SUMS) {
            bf16* o0 = dy.row(cu.n, 2 * oy) + (2 * ox) * C + c0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float2 g[4];
                if (MODE == PASS2_GATHER) {                          // dy = sc*dz + s2*(y - mu) + s1
                    float2 yq[4];
                    P8<bf16>::up(ry[q], yq);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        g[i] = fma2(s2[i], add2(yq[i], nmu[i]), s1[i]);
                        g[i].x += bi[2 * i] == q ? dzw[i].x : 0.f;
                        g[i].y += bi[2 * i + 1] == q ? dzw[i].y : 0.f;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        g[i] = make_float2(bi[2 * i] == q ? dzw[i].x : 0.f, bi[2 * i + 1] == q ? dzw[i].y : 0.f);
                }
                P8<bf16>::st(o0 + (q >> 1) * dy.sy + (q & 1) * C, g);
            }
        }
    };
    pipe_run<PIPE_BPOOL_STAGE, PIPE_BPOOL_STAGES>(smem, units, OH, cpr, rev, issue, body);
    if (tid >= pipe::CONSUMERS || MODE == PASS2_GATHER) return;
    pipe::consumer_sync();
    float* red = reinterpret_cast<float*>(smem);
    float2 is[4];
    load_c8(invstd, c0, 1.f, is);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        red[tid * 8 + 2 * i] = s1[i].x; red[tid * 8 + 2 * i + 1] = s1[i].y;
        red[2048 + tid * 8 + 2 * i] = s2[i].x * is[i].x; red[2048 + tid * 8 + 2 * i + 1] = s2[i].y * is[i].y;
    }
    pipe::consumer_sync();
    for (int ch = tid; ch < C; ch += pipe::CONSUMERS) {
        const int g8 = ch >> 3, i = ch & 7;
        float a = 0.f, b = 0.f;
        for (int t = g8; t < pipe::CONSUMERS; t += ncg) { a += red[t * 8 + i]; b += red[2048 + t * 8 + i]; }
        atomicAdd(&sums[ch], (double)a);
        atomicAdd(&sums[C + ch], (double)b);
    }
}

// ------------------------------------------------------------------------------------------------
// Bilinear x2 (align_corners=True) variants.  Every output row blends two input rows, so a CTA walks DOWN a strip
// (image, band of input rows, 256-item column chunk): input rows stream through the ring once, each consumer thread owns
// one (pixel, channel group) column of the strip and carries the horizontally interpolated rows it still needs in
// registers - no re-reads, no shared-memory round trip for intermediates, no block barriers.
// ------------------------------------------------------------------------------------------------
constexpr int PIPE_UP_ITEMS = 256;
constexpr int PIPE_UP_BAND = 16;                                            // input rows per strip
constexpr int PIPE_FUP_STAGE = PIPE_UP_ITEMS * 16 + 2 * PIPE_EXT, PIPE_FUP_STAGES = 8;

struct Strip {
    int n, k0, k1, ch;
    __device__ __forceinline__ void set(int t, int cpr, int nbands, int H) {
        ch = t % cpr;
        const int q = t / cpr;
        const int band = q % nbands;
        n = q / nbands;
        k0 = band * PIPE_UP_BAND;
        k1 = min(k0 + PIPE_UP_BAND, H);
    }
};

// forward: out(2k+a, 2j+b) from the activated 3x3 neighbourhood of input pixel (k, j); fractions as PyTorch computes them
template <int ACT>
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_fwd_up_pipe_k(Rows<const bf16> y, Rows<bf16> out, const float* __restrict__ scale, const float* __restrict__ shift,
                 int pad, int N, int H, int W, int C, int cg_shift, int cpr, const BnFuse fuse) {
    using namespace pipe;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int STAGES = PIPE_FUP_STAGES, STAGE_BYTES = PIPE_FUP_STAGE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = s32(smem + STAGES * STAGE_BYTES);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (STAGES + s), CONSUMERS / 32); }
        fence_init();
    }
    __syncthreads();
    const int nbands = (H + PIPE_UP_BAND - 1) / PIPE_UP_BAND;
    const int nstrips = N * nbands * cpr;
    const int ext = C * 2;
    if (warp == CONSUMERS / 32) {
        if (lane == 0) {
            int s = 0, ph = 0;
            for (int t = blockIdx.x; t < nstrips; t += gridDim.x) {
                Strip sp;
                sp.set(t, cpr, nbands, H);
                const int rs = max(sp.k0 - 1, 0), re = min(sp.k1, H - 1);
                for (int r = rs; r <= re; ++r) {
                    mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                    const uint32_t full = bars + 8 * s, stage = s32(smem + s * STAGE_BYTES);
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(y.row(sp.n, r) + sp.ch * (PIPE_UP_ITEMS * 8));
                    const int lh = sp.ch == 0 ? 0 : ext;          // no left halo at the start of a row (it may lie before the buffer)
                    mbar_expect_tx(full, PIPE_UP_ITEMS * 16 + ext + lh);
                    bulk_g2s(stage + PIPE_EXT - lh, src - lh, PIPE_UP_ITEMS * 16 + ext + lh, full);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }
    const int ncg = C >> 3;
    const int cg = tid & (ncg - 1), c0 = cg * 8;
    const int pl = tid >> cg_shift;
    float2 sc[4], sh[4];
    if (fuse.stats) {
        float a[8], b[8];
        fused_affine(fuse, C, c0, blockIdx.x == 0 && pl == 0, a, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) { sc[i] = make_float2(a[2 * i], a[2 * i + 1]); sh[i] = make_float2(b[2 * i], b[2 * i + 1]); }
    } else {
        load_c8(scale, c0, 1.f, sc);
        load_c8(shift, c0, 0.f, sh);
    }
    const int OH = 2 * H, OW = 2 * W;
    const float ry = (float)(H - 1) / (float)(OH - 1), rx = (float)(W - 1) / (float)(OW - 1);
    auto act8 = [&](const uint4& r, float2 (&o)[4]) {
        P8<bf16>::up(r, o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 z = fma2(o[i], sc[i], sh[i]);
            o[i] = make_float2(actv<ACT>(z.x), actv<ACT>(z.y));
        }
    };
    auto lerp8 = [&](const float2 (&a)[4], const float2 (&b)[4], float l, float2 (&o)[4]) {
        const float2 l2 = make_float2(l, l);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = fma2(l2, sub2(b[i], a[i]), a[i]);
    };
    int s = 0, ph = 0;
    for (int t = blockIdx.x; t < nstrips; t += gridDim.x) {
        Strip sp;
        sp.set(t, cpr, nbands, H);
        const int j = sp.ch * (PIPE_UP_ITEMS >> cg_shift) + pl;              // input pixel column of this thread
        const int offL = j == 0 ? 0 : -ext, offR = j == W - 1 ? 0 : ext;
        const float lx0 = j > 0 ? fminf(fmaxf(rx * (float)(2 * j) - (float)(j - 1), 0.f), 1.f) : 0.f;
        const float lx1 = fminf(fmaxf(rx * (float)(2 * j + 1) - (float)j, 0.f), 1.f);
        bf16* obase = out.row(sp.n, 0) + (2 * j + pad) * C + c0;
        // h[parity][8 channels] of the two previous input rows
        float2 hA0[4], hA1[4], hB0[4], hB1[4];
        auto emit = [&](int k, const float2 (&a0)[4], const float2 (&a1)[4], const float2 (&b0)[4], const float2 (&b1)[4],
                        const float2 (&c0v)[4], const float2 (&c1v)[4]) {
            const float ly0 = k > 0 ? fminf(fmaxf(ry * (float)(2 * k) - (float)(k - 1), 0.f), 1.f) : 0.f;
            const float ly1 = fminf(fmaxf(ry * (float)(2 * k + 1) - (float)k, 0.f), 1.f);
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                float2 o0[4], o1[4];
                if (a == 0) { lerp8(a0, b0, ly0, o0); lerp8(a1, b1, ly0, o1); }
                else { lerp8(b0, c0v, ly1, o0); lerp8(b1, c1v, ly1, o1); }
                const uint4 u0 = make_uint4(P8<bf16>::pk(o0[0]), P8<bf16>::pk(o0[1]), P8<bf16>::pk(o0[2]), P8<bf16>::pk(o0[3]));
                const uint4 u1 = make_uint4(P8<bf16>::pk(o1[0]), P8<bf16>::pk(o1[1]), P8<bf16>::pk(o1[2]), P8<bf16>::pk(o1[3]));
                const int oy = 2 * k + a;
                bf16* p = obase + (long long)(oy + pad) * out.sy;
                auto put = [&](bf16* q) {
                    *reinterpret_cast<uint4*>(q) = u0;
                    *reinterpret_cast<uint4*>(q + C) = u1;
                    if (pad) {
                        if (j == 0) *reinterpret_cast<uint4*>(q - C) = u0;
                        if (j == W - 1) *reinterpret_cast<uint4*>(q + 2 * C) = u1;
                    }
                };
                put(p);
                if (pad) {
                    if (oy == 0) put(p - out.sy);
                    if (oy == OH - 1) put(p + out.sy);
                }
            }
        };
        const int rs = max(sp.k0 - 1, 0), re = min(sp.k1, H - 1);
        for (int r = rs; r <= re; ++r) {
            mbar_wait(bars + 8 * s, ph);
            const uint8_t* st = smem + s * STAGE_BYTES + PIPE_EXT + tid * 16;
            const uint4 rl = lds16(st + offL), rc = lds16(st), rr = lds16(st + offR);
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (STAGES + s));
            if (++s == STAGES) { s = 0; ph ^= 1; }
            float2 al[4], ac[4], ar[4], hC0[4], hC1[4];
            act8(rl, al); act8(rc, ac); act8(rr, ar);
            lerp8(al, ac, lx0, hC0);
            lerp8(ac, ar, lx1, hC1);
            if (r == rs) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { hA0[i] = hC0[i]; hA1[i] = hC1[i]; hB0[i] = hC0[i]; hB1[i] = hC1[i]; }
                if (!(H == 1)) continue;
            }
            // rows (r-2 | r-1 | r) are in (hA | hB | hC): emit the output pair of input row r-1 if this strip owns it
            if (r - 1 >= sp.k0 && r - 1 < sp.k1 && r >= 1) emit(r - 1, hA0, hA1, hB0, hB1, hC0, hC1);
#pragma unroll
            for (int i = 0; i < 4; ++i) { hA0[i] = hB0[i]; hA1[i] = hB1[i]; hB0[i] = hC0[i]; hB1[i] = hC1[i]; }
            if (r == H - 1 && r >= sp.k0 && r < sp.k1) emit(r, hA0, hA1, hB0, hB1, hB0, hB1);     // last row: (H-2 | H-1 | H-1)
        }
    }
}

// backward pass 1 of the x2 bilinear layers (replicate-padded dout only).  Input pixel (k, j) receives from the 4x4 output
// pixels (2k-1..2k+2, 2j-1..2j+2), i.e. PADDED rows 2k..2k+3 and PADDED columns 2j..2j+3 of dout; the border rows / columns
// of the padding fold into the first / last output row / column, which only changes the tap weight (never the address).
// A strip streams padded dout rows in pairs (2k+2, 2k+3) together with y row k; the horizontally gathered rows 2k, 2k+1
// are carried in registers from the previous step.
constexpr int PIPE_BUP_DROW = 2 * PIPE_UP_ITEMS * 16 + 2 * PIPE_EXT;        // one padded dout row span (+2 pixels)
constexpr int PIPE_BUP_STAGE = 2 * PIPE_BUP_DROW + PIPE_UP_ITEMS * 16, PIPE_BUP_STAGES = 4;

__device__ __forceinline__ float up_weight(int i, int O, float r, int n_in, int n_out) {
    if (O < 0 || O >= n_out) return 0.f;
    const float s = r * (float)O;
    const int t = (int)s;
    const float l = s - (float)t;
    const int tb = min(t + 1, n_in - 1);
    return (t == i ? 1.f - l : 0.f) + (tb == i ? l : 0.f);
}

template <int ACT>
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_bwd_up_pipe_k(Rows<const bf16> dout, Rows<const bf16> y, Rows<bf16> dy, const float* __restrict__ scale,
                 const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                 double* sums, int N, int H, int W, int C, int cg_shift, int cpr) {
    using namespace pipe;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int STAGES = PIPE_BUP_STAGES, STAGE_BYTES = PIPE_BUP_STAGE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = s32(smem + STAGES * STAGE_BYTES);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (STAGES + s), CONSUMERS / 32); }
        fence_init();
    }
    __syncthreads();
    const int nbands = (H + PIPE_UP_BAND - 1) / PIPE_UP_BAND;
    const int nstrips = N * nbands * cpr;
    const int ext = C * 2;                                        // bytes of one pixel
    const int drow_bytes = 2 * PIPE_UP_ITEMS * 16 + 2 * ext;      // 2*np + 2 padded pixels
    if (warp == CONSUMERS / 32) {
        if (lane == 0) {
            int s = 0, ph = 0;
            for (int t = blockIdx.x; t < nstrips; t += gridDim.x) {
                Strip sp;
                sp.set(t, cpr, nbands, H);
                // step k0-1 preloads padded rows 2*k0, 2*k0+1 (no y row); step k loads padded rows 2k+2, 2k+3 and y row k
                for (int k = sp.k0 - 1; k < sp.k1; ++k) {
                    mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                    const uint32_t full = bars + 8 * s, stage = s32(smem + s * STAGE_BYTES);
                    const bool with_y = k >= sp.k0;
                    mbar_expect_tx(full, 2 * drow_bytes + (with_y ? PIPE_UP_ITEMS * 16 : 0));
                    const bf16* d0 = dout.row(sp.n, 2 * k + 2) + sp.ch * (2 * PIPE_UP_ITEMS * 8);
                    bulk_g2s(stage, d0, drow_bytes, full);
                    bulk_g2s(stage + PIPE_BUP_DROW, d0 + dout.sy, drow_bytes, full);
                    if (with_y) bulk_g2s(stage + 2 * PIPE_BUP_DROW, y.row(sp.n, k) + sp.ch * (PIPE_UP_ITEMS * 8), PIPE_UP_ITEMS * 16, full);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }
    const int ncg = C >> 3;
    const int cg = tid & (ncg - 1), c0 = cg * 8;
    const int pl = tid >> cg_shift;
    float2 sc[4], sh[4], nmu[4], s1[4], s2[4];
    bwd_consts(PASS1_WRITE, scale, shift, mean, invstd, sums, 1.0, C, c0, sc, sh, nmu, s1, s2);
    const int OH = 2 * H, OW = 2 * W;
    const float ry = (float)(H - 1) / (float)(OH - 1), rx = (float)(W - 1) / (float)(OW - 1);
    const uint32_t doff = ((2 * pl) << cg_shift | cg) * 16;       // padded column 2j of this thread inside a dout row span
    // sum_b wx[b] * dout[row][padded col 2j + b]
    auto hgather = [&](const uint8_t* rowp, const float (&wx)[4], float2 (&o)[4]) {
        uint4 r[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) r[b] = lds16(rowp + doff + b * ext);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float2 v[4];
            P8<bf16>::up(r[b], v);
            const float2 w2 = make_float2(wx[b], wx[b]);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = b == 0 ? make_float2(w2.x * v[i].x, w2.y * v[i].y) : fma2(w2, v[i], o[i]);
        }
    };
    int s = 0, ph = 0;
    for (int t = blockIdx.x; t < nstrips; t += gridDim.x) {
        Strip sp;
        sp.set(t, cpr, nbands, H);
        const int j = sp.ch * (PIPE_UP_ITEMS >> cg_shift) + pl;
        float wx[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) wx[b] = up_weight(j, 2 * j - 1 + b, rx, W, OW);
        if (j == 0) wx[0] = wx[1];                               // left padding column folds into output column 0
        if (j == W - 1) wx[3] = wx[2];                           // right padding column folds into output column OW-1
        float2 hA[4], hB[4];
        for (int k = sp.k0 - 1; k < sp.k1; ++k) {
            mbar_wait(bars + 8 * s, ph);
            const uint8_t* st = smem + s * STAGE_BYTES;
            float2 hC[4], hD[4];
            hgather(st, wx, hC);
            hgather(st + PIPE_BUP_DROW, wx, hD);
            uint4 ry4 = make_uint4(0, 0, 0, 0);
            if (k >= sp.k0) ry4 = lds16(st + 2 * PIPE_BUP_DROW + tid * 16);
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (STAGES + s));
            if (++s == STAGES) { s = 0; ph ^= 1; }
            if (k >= sp.k0) {
                float wy[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) wy[a] = up_weight(k, 2 * k - 1 + a, ry, H, OH);
                if (k == 0) wy[0] = wy[1];
                if (k == H - 1) wy[3] = wy[2];
                float2 g[4], yv[4];
                P8<bf16>::up(ry4, yv);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 acc = make_float2(wy[0] * hA[i].x, wy[0] * hA[i].y);
                    acc = fma2(make_float2(wy[1], wy[1]), hB[i], acc);
                    acc = fma2(make_float2(wy[2], wy[2]), hC[i], acc);
                    acc = fma2(make_float2(wy[3], wy[3]), hD[i], acc);
                    const float2 z = fma2(yv[i], sc[i], sh[i]);
                    const float2 dz = make_float2(actg<ACT>(z.x, acc.x), actg<ACT>(z.y, acc.y));
                    g[i] = dz;
                    s1[i] = add2(s1[i], dz);
                    s2[i] = fma2(dz, add2(yv[i], nmu[i]), s2[i]);
                }
                P8<bf16>::st(dy.row(sp.n, k) + j * C + c0, g);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { hA[i] = hC[i]; hB[i] = hD[i]; }
        }
    }
    consumer_sync();
    float* red = reinterpret_cast<float*>(smem);
    float2 is[4];
    load_c8(invstd, c0, 1.f, is);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        red[tid * 8 + 2 * i] = s1[i].x; red[tid * 8 + 2 * i + 1] = s1[i].y;
        red[2048 + tid * 8 + 2 * i] = s2[i].x * is[i].x; red[2048 + tid * 8 + 2 * i + 1] = s2[i].y * is[i].y;
    }
    consumer_sync();
    for (int ch = tid; ch < C; ch += CONSUMERS) {
        const int g8 = ch >> 3, i = ch & 7;
        float a = 0.f, b = 0.f;
        for (int t = g8; t < CONSUMERS; t += ncg) { a += red[t * 8 + i]; b += red[2048 + t * 8 + i]; }
        atomicAdd(&sums[ch], (double)a);
        atomicAdd(&sums[C + ch], (double)b);
    }
}
