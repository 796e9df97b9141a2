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
    int n, yy, ch;
    int dn, dyy, dch, R, cpr;
    __device__ __forceinline__ void init(int u0, int stride, int rows_per_image, int chunks_per_row) {
        R = rows_per_image; cpr = chunks_per_row;
        ch = u0 % cpr; int row = u0 / cpr; yy = row % R; n = row / R;
        dch = stride % cpr; int drow = stride / cpr; dyy = drow % R; dn = drow / R;
    }
    __device__ __forceinline__ void next() {
        ch += dch;
        int c = ch >= cpr ? 1 : 0;
        ch -= c ? cpr : 0;
        yy += dyy + c;
        c = yy >= R ? 1 : 0;
        yy -= c ? R : 0;
        n += dn + c;
    }
};

// Shared skeleton: STAGES-deep ring of STAGE_BYTES.  `issue(cur, stage, full)` (one producer thread) posts the expected
// byte count and the bulk copies of the unit at cursor `cur`; `body(cur, stage)` consumes one unit (256 consumer threads).
template <int STAGE_BYTES, int STAGES, class IssueFn, class BodyFn>
__device__ __forceinline__ void pipe_run(uint8_t* smem, int units, int rows_per_image, int cpr, IssueFn issue, BodyFn body) {
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
    cur.init(blockIdx.x, gridDim.x, rows_per_image, cpr);
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
template <int ACT>
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_fwd_none_pipe_k(Rows<const bf16> y, Rows<bf16> out, const float* __restrict__ scale, const float* __restrict__ shift,
                   int pad, int N, int H, int W, int C, int cg_shift, int cpr, const BnFuse fuse) {
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
    pipe_run<PIPE_FWD_STAGE, PIPE_FWD_STAGES>(smem, units, PH, cpr, issue, body);
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
            const float m1x = (float)(sums[c] / count), m1y = (float)(sums[c + 1] / count);
            const float m2x = (float)(sums[C + c] / count), m2y = (float)(sums[C + c + 1] / count);
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
                   double* sums, double count, int pad, int N, int H, int W, int C, int cg_shift, int cpr) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int ncg = C >> 3;
    const int c0 = (tid & (ncg - 1)) * 8;
    float2 sc[4], sh[4], nmu[4], s1[4], s2[4];
    bwd_consts(MODE, scale, shift, mean, invstd, sums, count, C, c0, sc, sh, nmu, s1, s2);
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
    pipe_run<PIPE_BWD_STAGE, PIPE_BWD_STAGES>(smem, units, H, cpr, issue, body);
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
            const float m1 = (float)(sums[c] / count), m2 = (float)(sums[C + c] / count);
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
    pipe_run<PIPE_APPLY_STAGE, PIPE_APPLY_STAGES>(smem, units, H, cpr, issue, body);
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
                   int pad, int N, int OH, int OW, int C, int cg_shift, int cpr, const BnFuse fuse) {
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
        float2 v[4], a[4];
        auto act8 = [&](const uint4& r, float2 (&o)[4]) {
            P8<bf16>::up(r, o);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 z = fma2(o[i], sc[i], sh[i]);
                o[i] = make_float2(actv<ACT>(z.x), actv<ACT>(z.y));
            }
        };
        act8(r00, v);
        act8(r01, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float2(fmaxf(v[i].x, a[i].x), fmaxf(v[i].y, a[i].y));
        act8(r10, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float2(fmaxf(v[i].x, a[i].x), fmaxf(v[i].y, a[i].y));
        act8(r11, a);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = make_float2(fmaxf(v[i].x, a[i].x), fmaxf(v[i].y, a[i].y));
        const int ox = cu.ch * (PIPE_POOL_ITEMS >> cg_shift) + pl;
        bf16* dst = out.row(cu.n, cu.yy) + (ox + pad) * C + c0;
        const uint4 o = make_uint4(P8<bf16>::pk(v[0]), P8<bf16>::pk(v[1]), P8<bf16>::pk(v[2]), P8<bf16>::pk(v[3]));
        *reinterpret_cast<uint4*>(dst) = o;
        if (pad) {
            if (ox == 0) *reinterpret_cast<uint4*>(dst - C) = o;
            if (ox == OW - 1) *reinterpret_cast<uint4*>(dst + C) = o;
        }
    };
    pipe_run<PIPE_FPOOL_STAGE, PIPE_FPOOL_STAGES>(smem, units, PH, cpr, issue, body);
}

// backward pass 1: the gradient of a pooled pixel goes to the first maximum of its window (row-major), zeros elsewhere
template <int ACT, int MODE>
__global__ void __launch_bounds__(pipe::THREADS, 2)
bn_bwd_pool_pipe_k(Rows<const bf16> dout, Rows<const bf16> y, Rows<bf16> dy, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                   double* sums, double count, int pad, int N, int OH, int OW, int C, int cg_shift, int cpr) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int ncg = C >> 3;
    const int cg = tid & (ncg - 1), c0 = cg * 8;
    float2 sc[4], sh[4], nmu[4], s1[4], s2[4];
    bwd_consts(MODE, scale, shift, mean, invstd, sums, count, C, c0, sc, sh, nmu, s1, s2);
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
        float2 best[4];
        int bi[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 yq[4];
            P8<bf16>::up(ry[q], yq);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 z = fma2(yq[i], sc[i], sh[i]);
                const float ax = actv<ACT>(z.x), ay = actv<ACT>(z.y);
                if (q == 0 || ax > best[i].x) { best[i].x = ax; bi[2 * i] = q; }
                if (q == 0 || ay > best[i].y) { best[i].y = ay; bi[2 * i + 1] = q; }
            }
        }
        bf16* o0 = MODE == PASS1_SUMS ? nullptr : dy.row(cu.n, 2 * oy) + (2 * ox) * C + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float2 g[4], yq[4];
            P8<bf16>::up(ry[q], yq);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 z = fma2(yq[i], sc[i], sh[i]);
                const float2 dz = make_float2(bi[2 * i] == q ? actg<ACT>(z.x, t[i].x) : 0.f,
                                              bi[2 * i + 1] == q ? actg<ACT>(z.y, t[i].y) : 0.f);
                const float2 yc = add2(yq[i], nmu[i]);
                if (MODE == PASS2_GATHER) {
                    g[i] = fma2(sc[i], dz, fma2(s2[i], yc, s1[i]));
                } else {
                    g[i] = dz;
                    s1[i] = add2(s1[i], dz);
                    s2[i] = fma2(dz, yc, s2[i]);
                }
            }
            if (MODE != PASS1_SUMS) P8<bf16>::st(o0 + (q >> 1) * dy.sy + (q & 1) * C, g);
        }
    };
    pipe_run<PIPE_BPOOL_STAGE, PIPE_BPOOL_STAGES>(smem, units, OH, cpr, issue, body);
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
