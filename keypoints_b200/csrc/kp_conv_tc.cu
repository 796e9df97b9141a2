// Convolutions on the Blackwell tensor cores: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM),
// operands staged by TMA with the 128-byte swizzle, mbarrier producer/consumer ring, warp-specialised
// (warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue).
//
// "Flat pixel" formulation.  Every activation lives in a replicate-padded NHWC buffer
// [N][PH][PW][C] (PH = H+2, PW = W+2) which is also a 2-D matrix [Q = N*PH*PW pixels][C].  Because the
// input and output share the row pitch PW, a 3x3 tap is a CONSTANT shift of the flat pixel index:
//     out[q][co] = sum_t sum_ci in[q + shift_t][ci] * w[t][co][ci]
// so the implicit-GEMM A tile of a tap is one plain 2-D TMA box (64 channels x 128 pixels) at row
// q0 + shift_t, for any image width, and TMA's out-of-range zero fill covers both ends of the buffer.
//   fprop : in = replicate-padded X, shift_t = ty*PW + tx, out = Y aligned top-left (rows with
//           x >= W or y >= H are scratch and are masked out of the BatchNorm statistics).
//   dgrad : in = dY stored interior-aligned with a ZERO border, shift_t' = (t'y-1)*PW + (t'x-1),
//           weights flipped/transposed; every row of out = d(padded X) is valid.
//   wgrad : dw[t][co][ci] = sum_q dY[q][co] * X[q + shift_t][ci]; both operands are "MN-major"
//           (pixels = GEMM-K are the strided axis) which UMMA reads directly through MN-major
//           shared-memory descriptors; split over pixel ranges, fp32 vector reductions into a
//           staging gradient.
#include "kp_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"((uint64_t)tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(src) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 16 TMEM lanes x 32 columns in the mma C-fragment layout (probed on B200, scripts/exp_tmem_ld_shapes.cu): with g = lane / 4,
// q = lane % 4, r[4x + 0..1] = (row g, columns 8x + 2q, 8x + 2q + 1), r[4x + 2..3] = (row g + 8, same columns), x = 0..3.
// No wait inside: the caller issues its loads back to back and waits once.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// shared-memory matrix descriptor (sm_100 format, version 1), 128-byte swizzle
//   K-major : rows of 128 B (64 bf16 of K), 8-row groups every SBO = 1024 B; LBO unused (1)
//   MN-major: rows of 128 B (64 bf16 of M/N) per k, 8-k groups every SBO = 1024 B, next 64 M/N every LBO
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}
// instruction descriptor: D=f32, A=B=bf16, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int STAGE_A_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int SLAB_BYTES = 128 * 128;       // epilogue staging slab: 128 rows x 64 bf16, SWIZZLE_128B, x2 buffers
constexpr int PF_TILES = 3;                // L2 prefetch distance of the persistent producers, in rounds of tiles

struct ConvTcParams {
    long long Q;
    int Cin, Cout, taps;
    int shift[9];
    const float* bias;
    bf16* out;
    double* stats;
    int PH, PW, VH, VW;
};

struct WgradTcParams {
    long long Q;            // flat pixels
    long long kchunk;       // pixels per split (multiple of 64)
    int splits;
    int shiftA[9], shiftB[9];
    int Mtot, Ntot;         // staging is [taps][Mtot][Ntot] fp32
    float* stg;
    // image mode: the contraction runs over the VALID pixels only.  A k-block is a bx x by patch (bx * by = 64) of one image,
    // fetched with 4-D tensor maps (channel, x, y, image) of the padded buffers; tap t reads operand A at (+ax, +ay) and
    // operand B at (+bx_o, +by_o) from the patch origin.  The flat mode also multiplies the zero / pad border pixels
    // (3 % of the MMA work at 128x128, 27 % at 16x16).
    int img, bx, by, bpr, bpi;
    int ax[9], ay[9], bxo[9], byo[9];
};

// ================================================================================================
// Persistent variants: one CTA per SM walks a static round-robin list of output tiles.  The accumulator is double
// buffered in TMEM (2 x BN columns), so the epilogue of tile i (TMEM -> registers -> bias / BatchNorm statistics ->
// bf16 stores) overlaps the TMA + tcgen05.mma main loop of tile i+1, barriers / TMEM / tensor-map prefetch are paid
// once per CTA, and the BatchNorm statistics are accumulated per CTA and flushed with one double atomic per channel.
// ================================================================================================
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(320, 1)
conv_tc_persist_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, const ConvTcParams p) {
    constexpr int STAGE_B_BYTES = BN * 128;
    constexpr int STAGE_BYTES = STAGE_A_BYTES + STAGE_B_BYTES;
    constexpr int EPI_FLOATS = 8 * 32 * 17 + 4 * 2 * BN;   // per-warp 32x16 transpose tiles, per-lane-group running column sums
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t slab = base + STAGES * STAGE_BYTES;                 // 2 x SLAB_BYTES, 1024-byte aligned
    uint8_t* slab_gen = gen_base + STAGES * STAGE_BYTES;
    float* epi = reinterpret_cast<float*>(gen_base + STAGES * STAGE_BYTES + 2 * SLAB_BYTES);
    const uint32_t bars = base + STAGES * STAGE_BYTES + 2 * SLAB_BYTES + EPI_FLOATS * 4;   // full[S], empty[S], tfull[2], tempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + STAGES * STAGE_BYTES + 2 * SLAB_BYTES + EPI_FLOATS * 4 + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kchunks = p.Cin / 64;
    const int num_kb = p.taps * kchunks;
    const int num_m = (int)((p.Q + 127) / 128);
    const int total = num_m * (p.Cout / BN);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bars + 8 * (2 * STAGES + b), 1);        // tmem full  (tcgen05.commit)
            mbar_init(bars + 8 * (2 * STAGES + 2 + b), 1);    // tmem empty (one elected epilogue thread)
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * BN);
    if (threadIdx.x >= 64) {
        float* wsum0 = epi + 8 * 32 * 17;
        for (int i = threadIdx.x - 64; i < 4 * 2 * BN; i += 256) wsum0[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int n_t = t / num_m, m_t = t - n_t * num_m;
                const long long q0 = (long long)m_t * 128;
                const int n0 = n_t * BN;
                {   // pull the not-yet-touched activation rows of a tile several rounds ahead into L2 (DRAM latency
                    // is otherwise exposed when the K loop of a tile is short)
                    const int tp = t + PF_TILES * (int)gridDim.x;
                    if (tp < total) {
                        const int mp = tp - (tp / num_m) * num_m;
                        for (int kc = 0; kc < kchunks; ++kc)
                            tma_prefetch_l2_2d(&tmA, kc * 64, (int)((long long)mp * 128 + p.shift[p.taps - 1]));
                    }
                }
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                    const uint32_t full = bars + 8 * s;
                    mbar_expect_tx(full, STAGE_BYTES);
                    const int tap = kb / kchunks, kc = kb - tap * kchunks;
                    const uint32_t sa = base + s * STAGE_BYTES;
                    tma_load_2d(sa, &tmA, kc * 64, (int)(q0 + p.shift[tap]), full);
                    tma_load_2d(sa + STAGE_A_BYTES, &tmB, kc * 64, tap * p.Cout + n0, full);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(BN, 0, 0);
            int it = 0, lt = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
                const int buf = lt & 1;
                mbar_wait(bars + 8 * (2 * STAGES + 2 + buf), ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * STAGE_BYTES;
                    const uint64_t ad = make_desc(sa, 16, 1024), bd = make_desc(sa + STAGE_A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(dcol, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(bars + 8 * (STAGES + s));
                }
                umma_commit(bars + 8 * (2 * STAGES + buf));
            }
        }
    } else {
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int et = threadIdx.x - 64;                       // 0..255: 8 epilogue warps, 2 per TMEM lane group
        const int half = (warp - 2) >> 2;                      // which interleaved set of 32-column chunks
        float* tile = epi + (warp - 2) * (32 * 17);
        float* wsum_all = epi + 8 * 32 * 17;                   // [4 lane groups][2*BN] running column sums of this CTA
        float* wsum = wsum_all + wq * (2 * BN);
        int lt = 0, cur_n = -1, slabs = 0;
        auto flush = [&](int n_tile) {                         // all 128 epilogue threads
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int ch = et; ch < BN; ch += 256) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    s1 += wsum_all[w * 2 * BN + ch];
                    s2 += wsum_all[w * 2 * BN + BN + ch];
                    wsum_all[w * 2 * BN + ch] = 0.f;
                    wsum_all[w * 2 * BN + BN + ch] = 0.f;
                }
                atomicAdd(&p.stats[n_tile * BN + ch], (double)s1);
                atomicAdd(&p.stats[p.Cout + n_tile * BN + ch], (double)s2);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        };
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
            const int n_t = t / num_m, m_t = t - n_t * num_m;
            const int n0 = n_t * BN;
            const int buf = lt & 1;
            if (p.stats && cur_n >= 0 && cur_n != n_t) flush(cur_n);
            cur_n = n_t;
            const long long q_tile = (long long)m_t * 128;
            const long long q = q_tile + row;
            bool valid = q < p.Q;
            if (valid && p.stats) {
                const unsigned qu = (unsigned)q, row_i = qu / (unsigned)p.PW;       // Q < 2^31 (checked by the launcher): 32-bit divisions
                int x = (int)(qu - row_i * (unsigned)p.PW);
                int y = (int)(row_i % (unsigned)p.PH);
                valid = x < p.VW && y < p.VH;
            }
            mbar_wait(bars + 8 * (2 * STAGES + buf), (lt >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = half; c < BN / 32; c += 2) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + buf * BN + c * 32, r);
                float v[32];
                if (p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c * 32);
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 bv = bp[g];
                        v[4 * g] = __uint_as_float(r[4 * g]) + bv.x; v[4 * g + 1] = __uint_as_float(r[4 * g + 1]) + bv.y;
                        v[4 * g + 2] = __uint_as_float(r[4 * g + 2]) + bv.z; v[4 * g + 3] = __uint_as_float(r[4 * g + 3]) + bv.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[n0 + c * 32 + j] : 0.f);
                }
                {   // stage this warp's 32 rows x 32 columns into the 128-row x 64-column slab (128-byte swizzle: the
                    // 16-byte chunk j of row r lives at chunk j ^ (r & 7)), then ONE TMA store writes the slab
                    const int sb = slabs & 1;
                    if (et == 0) bulk_wait_read<1>();          // the store that last read this buffer is done with it
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    uint8_t* drow = slab_gen + sb * SLAB_BYTES + row * 128;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 u;
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                        *reinterpret_cast<uint4*>(drow + (((half * 4 + g) ^ (row & 7)) << 4)) = u;
                    }
                    fence_async_smem();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (et == 0) tma_store_2d(&tmO, slab + sb * SLAB_BYTES, n0 + (c >> 1) * 64, (int)q_tile);
                    ++slabs;
                }
                if (p.stats) {
                    // column sums over the warp's 32 rows, 16 columns at a time through a 32x17 shared tile: lane l sums
                    // column (l & 15) over rows [16 (l >> 4), +16), the two row halves are combined with one shuffle
#pragma unroll
                    for (int hc = 0; hc < 2; ++hc) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) tile[lane * 17 + j] = valid ? v[hc * 16 + j] : 0.f;
                        __syncwarp();
                        float s1 = 0.f, s2 = 0.f;
                        const int col = lane & 15, r0 = (lane >> 4) * 16;
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            float tv = tile[(r0 + rr) * 17 + col];
                            s1 += tv;
                            s2 = fmaf(tv, tv, s2);
                        }
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                        if (lane < 16) {
                            wsum[c * 32 + hc * 16 + col] += s1;
                            wsum[BN + c * 32 + hc * 16 + col] += s2;
                        }
                        __syncwarp();
                    }
                }
            }
            // every epilogue thread has drained its TMEM rows: release the accumulator buffer to the MMA warp
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) mbar_arrive(bars + 8 * (2 * STAGES + 2 + buf));
        }
        if (p.stats && cur_n >= 0) flush(cur_n);
        if (et == 0) bulk_wait_all();                          // every TMA store has completed before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster on one TPC compute a 256-pixel x BN tile with one
// M=256 MMA stream issued by the leader.  Each CTA stages its own 128 pixel rows of A and only HALF of the weight tile
// (BN/2 rows), so a pipeline stage is 16 KB + BN/2*128 B instead of 16 KB + BN*128 B: the ring holds 1.5x more k-blocks
// in flight for the same shared memory, which is what the TMA-latency-bound main loop needs.
// Barriers: the TMA loads of both CTAs complete on the LEADER's full barrier; tcgen05.commit multicasts the
// stage-free and accumulator-ready signals to both CTAs; both epilogues arrive on the leader's accumulator-free barrier.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address -> leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar & PEER_MASK)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar & PEER_MASK)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m256(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
conv_tc_pair_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, const ConvTcParams p) {
    constexpr int STAGE_B_BYTES = (BN / 2) * 128;              // this CTA's half of the weight tile
    constexpr int STAGE_BYTES = STAGE_A_BYTES + STAGE_B_BYTES;
    constexpr int EPI_FLOATS = 8 * 32 * 17 + 4 * 2 * BN;   // per-warp 32x16 transpose tiles, per-lane-group running column sums
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t slab = base + STAGES * STAGE_BYTES;                 // 2 x SLAB_BYTES, 1024-byte aligned
    uint8_t* slab_gen = gen_base + STAGES * STAGE_BYTES;
    float* epi = reinterpret_cast<float*>(gen_base + STAGES * STAGE_BYTES + 2 * SLAB_BYTES);
    const uint32_t bars = base + STAGES * STAGE_BYTES + 2 * SLAB_BYTES + EPI_FLOATS * 4;   // full[S], empty[S], tfull[2], tempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + STAGES * STAGE_BYTES + 2 * SLAB_BYTES + EPI_FLOATS * 4 + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int kchunks = p.Cin / 64;
    const int num_kb = p.taps * kchunks;
    const int num_m = (int)((p.Q + 255) / 256);                 // pair tiles of 256 pixels
    const int total = num_m * (p.Cout / BN);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bars + 8 * (2 * STAGES + b), 1);        // tmem full  (tcgen05.commit)
            mbar_init(bars + 8 * (2 * STAGES + 2 + b), 2);    // tmem empty: one arrive from each CTA's epilogue
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= 64) {
        float* wsum0 = epi + 8 * 32 * 17;
        for (int i = threadIdx.x - 64; i < 4 * 2 * BN; i += 256) wsum0[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                         // both CTAs' barriers are initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nclusters = gridDim.x >> 1, cid = blockIdx.x >> 1;

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int t = cid; t < total; t += nclusters) {
                const int n_t = t / num_m, m_t = t - n_t * num_m;
                const long long q0 = (long long)m_t * 256 + rank * 128;
                const int n0 = n_t * BN + rank * (BN / 2);
                {
                    const int tp = t + PF_TILES * nclusters;
                    if (tp < total) {
                        const int mp = tp - (tp / num_m) * num_m;
                        for (int kc = 0; kc < kchunks; ++kc)
                            tma_prefetch_l2_2d(&tmA, kc * 64, (int)((long long)mp * 256 + rank * 128 + p.shift[p.taps - 1]));
                    }
                }
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                    const uint32_t full = bars + 8 * s;
                    if (leader) mbar_expect_tx(full, 2 * STAGE_BYTES);      // bytes of both CTAs land on the leader's barrier
                    const int tap = kb / kchunks, kc = kb - tap * kchunks;
                    const uint32_t sa = base + s * STAGE_BYTES;
                    tma_load_2d_pair(sa, &tmA, kc * 64, (int)(q0 + p.shift[tap]), full);
                    tma_load_2d_pair(sa + STAGE_A_BYTES, &tmB, kc * 64, tap * p.Cout + n0, full);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            constexpr uint32_t idesc = make_idesc_m256(BN);
            int it = 0, lt = 0;
            for (int t = cid; t < total; t += nclusters, ++lt) {
                const int buf = lt & 1;
                mbar_wait(bars + 8 * (2 * STAGES + 2 + buf), ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * STAGE_BYTES;
                    const uint64_t ad = make_desc(sa, 16, 1024), bd = make_desc(sa + STAGE_A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_pair(dcol, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit_pair(bars + 8 * (STAGES + s));
                }
                umma_commit_pair(bars + 8 * (2 * STAGES + buf));
            }
        }
    } else {
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int et = threadIdx.x - 64;                       // 0..255: 8 epilogue warps, 2 per TMEM lane group
        const int half = (warp - 2) >> 2;                      // which interleaved set of 32-column chunks
        float* tile = epi + (warp - 2) * (32 * 17);
        float* wsum_all = epi + 8 * 32 * 17;                   // [4 lane groups][2*BN] running column sums of this CTA
        float* wsum = wsum_all + wq * (2 * BN);
        int lt = 0, cur_n = -1, slabs = 0;
        auto flush = [&](int n_tile) {                         // all 128 epilogue threads
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int ch = et; ch < BN; ch += 256) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    s1 += wsum_all[w * 2 * BN + ch];
                    s2 += wsum_all[w * 2 * BN + BN + ch];
                    wsum_all[w * 2 * BN + ch] = 0.f;
                    wsum_all[w * 2 * BN + BN + ch] = 0.f;
                }
                atomicAdd(&p.stats[n_tile * BN + ch], (double)s1);
                atomicAdd(&p.stats[p.Cout + n_tile * BN + ch], (double)s2);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        };
        for (int t = cid; t < total; t += nclusters, ++lt) {
            const int n_t = t / num_m, m_t = t - n_t * num_m;
            const int n0 = n_t * BN;
            const int buf = lt & 1;
            if (p.stats && cur_n >= 0 && cur_n != n_t) flush(cur_n);
            cur_n = n_t;
            const long long q_tile = (long long)m_t * 256 + rank * 128;
            const long long q = q_tile + row;
            bool valid = q < p.Q;
            if (valid && p.stats) {
                const unsigned qu = (unsigned)q, row_i = qu / (unsigned)p.PW;       // Q < 2^31 (checked by the launcher): 32-bit divisions
                int x = (int)(qu - row_i * (unsigned)p.PW);
                int y = (int)(row_i % (unsigned)p.PH);
                valid = x < p.VW && y < p.VH;
            }
            mbar_wait(bars + 8 * (2 * STAGES + buf), (lt >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = half; c < BN / 32; c += 2) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + buf * BN + c * 32, r);
                float v[32];
                if (p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c * 32);
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 bv = bp[g];
                        v[4 * g] = __uint_as_float(r[4 * g]) + bv.x; v[4 * g + 1] = __uint_as_float(r[4 * g + 1]) + bv.y;
                        v[4 * g + 2] = __uint_as_float(r[4 * g + 2]) + bv.z; v[4 * g + 3] = __uint_as_float(r[4 * g + 3]) + bv.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[n0 + c * 32 + j] : 0.f);
                }
                {   // stage this warp's 32 rows x 32 columns into the 128-row x 64-column slab (128-byte swizzle: the
                    // 16-byte chunk j of row r lives at chunk j ^ (r & 7)), then ONE TMA store writes the slab
                    const int sb = slabs & 1;
                    if (et == 0) bulk_wait_read<1>();          // the store that last read this buffer is done with it
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    uint8_t* drow = slab_gen + sb * SLAB_BYTES + row * 128;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 u;
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                        *reinterpret_cast<uint4*>(drow + (((half * 4 + g) ^ (row & 7)) << 4)) = u;
                    }
                    fence_async_smem();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (et == 0) tma_store_2d(&tmO, slab + sb * SLAB_BYTES, n0 + (c >> 1) * 64, (int)q_tile);
                    ++slabs;
                }
                if (p.stats) {
                    // column sums over the warp's 32 rows, 16 columns at a time through a 32x17 shared tile: lane l sums
                    // column (l & 15) over rows [16 (l >> 4), +16), the two row halves are combined with one shuffle
#pragma unroll
                    for (int hc = 0; hc < 2; ++hc) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) tile[lane * 17 + j] = valid ? v[hc * 16 + j] : 0.f;
                        __syncwarp();
                        float s1 = 0.f, s2 = 0.f;
                        const int col = lane & 15, r0 = (lane >> 4) * 16;
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            float tv = tile[(r0 + rr) * 17 + col];
                            s1 += tv;
                            s2 = fmaf(tv, tv, s2);
                        }
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                        if (lane < 16) {
                            wsum[c * 32 + hc * 16 + col] += s1;
                            wsum[BN + c * 32 + hc * 16 + col] += s2;
                        }
                        __syncwarp();
                    }
                }
            }
            // every epilogue thread has drained its TMEM rows: release the accumulator buffer to the MMA warp
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) mbar_arrive_leader(bars + 8 * (2 * STAGES + 2 + buf));
        }
        if (p.stats && cur_n >= 0) flush(cur_n);
        if (et == 0) bulk_wait_all();                          // every TMA store has completed before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                         // no remote signal may target a CTA that has exited
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 variant with A-halo reuse.  The three horizontal taps of one kernel row read activation rows that differ by ONE
// flat pixel, so a single TMA box of 128 + 2 (padded to 136) pixel rows serves all three: the MMA of tap tx simply starts tx
// rows (tx * 128 B) into the tile.  That is legal because the 128-byte swizzle is a function of the absolute shared-memory
// address bits (chunk ^= (addr >> 7) & 7) for TMA and UMMA alike - measured on B200 with scripts/exp_umma_shift.cu: every
// row shift 0..7 is exact with the descriptor's base-offset field left at 0 (and wrong if it is set).  Activation traffic
// from L2 into shared memory drops 3x (the weight tiles are unchanged), which is what bounds the 64- and 128-channel
// layers at 128x128.  Two rings: AS slots of 18 KB for the activation tiles (one per kernel row and 64-channel chunk),
// BS slots for the per-tap weight half-tiles.
// ------------------------------------------------------------------------------------------------
constexpr int A3_ROWS = 136;
constexpr int A3_BYTES = A3_ROWS * 128;        // 17408 bytes landed by TMA
constexpr int A3_SLOT = 18432;                 // slot pitch (multiple of 1024)

// Sum over the 32 lanes of a warp of 32 per-lane values, transposed: lane l ends with the total of x[l] (returned).
// Five exchange steps of 16, 8, 4, 2, 1 values: 31 shuffles for 32 columns.
__device__ __forceinline__ float warp_transpose_sum32(float* x, int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
            const float send = up ? x[j] : x[j + off];
            const float keep = up ? x[j + off] : x[j];
            x[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return x[0];
}

// RB ("resident B"): the layer's whole weight tile set (9 taps x Cin/64 chunks, this CTA's half of the BN rows) fits the B
// ring exactly (BS == 9 * Cin / 64, Cout == BN): it is loaded once, on the first tile, and never released - the 64 <-> 128
// channel layers at 128x128 otherwise re-stage 72 KB of weights for every 54 KB of activations.
// BN <= 128: the BatchNorm statistics are accumulated per thread in registers across the tiles of a CTA (one row, BN/2
// columns per thread) and reduced over rows once per CTA with warp shuffles; the BN = 256 tiles keep the per-tile
// shared-memory transposes (256 accumulators per thread do not fit the register file).
template <int BN, int AS, int BS, bool RB, bool FE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((BN <= 128 && !FE) ? 384 : 320, 1)
conv_tc_pair3_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, const ConvTcParams p) {
    constexpr bool RS = BN <= 128 && !FE;                      // register-resident statistics, one row per thread
    constexpr int EW0 = RS ? 4 : 2;                            // first epilogue warp (RS: warpgroups 1 and 2)
    constexpr int NCH = BN / 64;                               // 32-column chunks per epilogue thread
    constexpr int B_BYTES = (BN / 2) * 128;                    // this CTA's half of the weight tile
    constexpr int RING_BYTES = AS * A3_SLOT + BS * B_BYTES;
    constexpr int NBAR = 2 * AS + 2 * BS + 4;                  // a_full, a_empty, b_full, b_empty, tfull[2], tempty[2]
    constexpr int AFULL = 0, AEMPTY = AS, BFULL = 2 * AS, BEMPTY = 2 * AS + BS, TFULL = 2 * AS + 2 * BS, TEMPTY = TFULL + 2;
    constexpr int TILE_FLOATS = (RS || FE) ? 0 : 8 * 32 * 17;  // per-warp 32x16 transpose tiles (BN = 256 without FE only)
    constexpr int EPI_FLOATS = TILE_FLOATS + 4 * 2 * BN;       // + per-lane-group running column sums
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t slab = base + RING_BYTES;                 // 2 x SLAB_BYTES, 1024-byte aligned
    uint8_t* slab_gen = gen_base + RING_BYTES;
    float* epi = reinterpret_cast<float*>(gen_base + RING_BYTES + 2 * SLAB_BYTES);
    const uint32_t bars = base + RING_BYTES + 2 * SLAB_BYTES + EPI_FLOATS * 4;   // full[S], empty[S], tfull[2], tempty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + RING_BYTES + 2 * SLAB_BYTES + EPI_FLOATS * 4 + 8 * NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int kchunks = p.Cin / 64;
    const int num_m = (int)((p.Q + 255) / 256);                 // pair tiles of 256 pixels
    const int total = num_m * (p.Cout / BN);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmO);
        for (int i = 0; i < 2 * AS + 2 * BS + 2; ++i) mbar_init(bars + 8 * i, 1);
        for (int b = 0; b < 2; ++b) mbar_init(bars + 8 * (TEMPTY + b), 2);   // tmem empty: one arrive from each CTA's epilogue
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x >= EW0 * 32) {
        float* wsum0 = epi + TILE_FLOATS;
        for (int i = threadIdx.x - EW0 * 32; i < 4 * 2 * BN; i += 256) wsum0[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                         // both CTAs' barriers are initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nclusters = gridDim.x >> 1, cid = blockIdx.x >> 1;

    // BN <= 128: three warpgroups, 0 = TMA producer (warp 0) + MMA issuer (warp 1) + two idle warps, 1 and 2 = epilogue.
    // The epilogue's per-thread BatchNorm accumulators need more than the 168 registers a 384-thread CTA gets evenly, so
    // warpgroup 0 hands most of its share to the other two (setmaxnreg).  BN = 256: 320 threads, warps 2-9 = epilogue
    // (113 registers: leaves room for a co-resident BatchNorm CTA of the other stream).
    if (warp < EW0) {
    if constexpr (RS) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
    if (warp == 0) {
        if (elect_one()) {
            int ia = 0, ib = 0;
            for (int t = cid; t < total; t += nclusters) {
                const int n_t = t / num_m, m_t = t - n_t * num_m;
                const long long q0 = (long long)m_t * 256 + rank * 128;
                const int n0 = n_t * BN + rank * (BN / 2);
                const bool load_b = !RB || t == cid;
                for (int ty = 0; ty < 3; ++ty)
                    for (int kc = 0; kc < kchunks; ++kc) {
                        const int sa_i = ia % AS, pa = (ia / AS) & 1;
                        ++ia;
                        mbar_wait(bars + 8 * (AEMPTY + sa_i), pa ^ 1);
                        const uint32_t afull = bars + 8 * (AFULL + sa_i);
                        if (leader) mbar_expect_tx(afull, 2 * A3_BYTES);
                        tma_load_2d_pair(base + sa_i * A3_SLOT, &tmA, kc * 64, (int)(q0 + p.shift[3 * ty]), afull);
                        if (load_b) {
                            for (int tx = 0; tx < 3; ++tx) {
                                const int sb_i = ib % BS, pb = (ib / BS) & 1;
                                ++ib;
                                mbar_wait(bars + 8 * (BEMPTY + sb_i), pb ^ 1);
                                const uint32_t bfull = bars + 8 * (BFULL + sb_i);
                                if (leader) mbar_expect_tx(bfull, 2 * B_BYTES);
                                tma_load_2d_pair(base + AS * A3_SLOT + sb_i * B_BYTES, &tmB, kc * 64, (3 * ty + tx) * p.Cout + n0, bfull);
                            }
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            constexpr uint32_t idesc = make_idesc_m256(BN);
            int ia = 0, ib = 0, lt = 0;
            for (int t = cid; t < total; t += nclusters, ++lt) {
                const int buf = lt & 1;
                mbar_wait(bars + 8 * (TEMPTY + buf), ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * BN;
                uint32_t first = 1;
                if (RB) ib = 0;
                for (int ty = 0; ty < 3; ++ty)
                    for (int kc = 0; kc < kchunks; ++kc) {
                        const int sa_i = ia % AS, pa = (ia / AS) & 1;
                        ++ia;
                        mbar_wait(bars + 8 * (AFULL + sa_i), pa);
                        const uint32_t sa = base + sa_i * A3_SLOT;
                        for (int tx = 0; tx < 3; ++tx) {
                            const int sb_i = ib % BS, pb = (ib / BS) & 1;
                            ++ib;
                            if (!RB || lt == 0) mbar_wait(bars + 8 * (BFULL + sb_i), pb);
                            tc_fence_after();
                            const uint64_t ad = make_desc(sa + tx * 128, 16, 1024);
                            const uint64_t bd = make_desc(base + AS * A3_SLOT + sb_i * B_BYTES, 16, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16_pair(dcol, ad + 2 * k, bd + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
                            first = 0;
                            if (!RB) umma_commit_pair(bars + 8 * (BEMPTY + sb_i));
                        }
                        umma_commit_pair(bars + 8 * (AEMPTY + sa_i));
                    }
                umma_commit_pair(bars + 8 * (TFULL + buf));
            }
        }
    }
    } else {
        if constexpr (RS) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;" ::: "memory");
        if constexpr (FE) {
            // Fragment epilogue: the accumulator is read in the mma C-fragment layout (tcgen05.ld.16x256b), so a thread holds
            // 4 rows x 2 adjacent columns of every 8-column group instead of one row x 32 columns: its BatchNorm column sums
            // need BN / 4 accumulators (64 at BN = 256) and stay in registers across all tiles of the CTA for every tile width.
            constexpr int NCF = BN / 64;                        // 32-column chunks per thread
            const int wq = warp & 3, g = lane >> 2, q4 = lane & 3;
            const int et = threadIdx.x - EW0 * 32;
            const int half = (warp - EW0) >> 2;
            float* wsum_all = epi + TILE_FLOATS;
            float* wsum = wsum_all + wq * (2 * BN);
            float a1[NCF * 8], a2[NCF * 8];
#pragma unroll
            for (int j = 0; j < NCF * 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
            int lt = 0, cur_n = -1, slabs = 0;
            auto flush = [&](int n_tile) {                      // all 256 epilogue threads
#pragma unroll
                for (int j = 0; j < NCF * 8; ++j) {
                    float s1 = a1[j], s2 = a2[j];
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
                    if (g == 0) {
                        const int col = (half + 2 * (j >> 3)) * 32 + 8 * ((j & 7) >> 1) + 2 * q4 + (j & 1);
                        wsum[col] = s1;
                        wsum[BN + col] = s2;
                    }
                    a1[j] = 0.f; a2[j] = 0.f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int ch = et; ch < BN; ch += 256) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        s1 += wsum_all[w * 2 * BN + ch];
                        s2 += wsum_all[w * 2 * BN + BN + ch];
                        wsum_all[w * 2 * BN + ch] = 0.f;
                        wsum_all[w * 2 * BN + BN + ch] = 0.f;
                    }
                    atomicAdd(&p.stats[n_tile * BN + ch], (double)s1);
                    atomicAdd(&p.stats[p.Cout + n_tile * BN + ch], (double)s2);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            };
            const bool bias_v = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 7) == 0);
            for (int t = cid; t < total; t += nclusters, ++lt) {
                const int n_t = t / num_m, m_t = t - n_t * num_m;
                const int n0 = n_t * BN;
                const int buf = lt & 1;
                if (p.stats && cur_n >= 0 && cur_n != n_t) flush(cur_n);
                cur_n = n_t;
                const long long q_tile = (long long)m_t * 256 + rank * 128;
                float vm[4] = {0.f, 0.f, 0.f, 0.f};             // rows wq * 32 + g + 8 k of the tile: 1 = counts in the statistics
                bool allv = false;                              // every row of this WARP counts: unmasked sums
                if (p.stats) {
                    const long long q0r = q_tile + wq * 32 + g;
                    const unsigned qu = (unsigned)q0r, row_i = qu / (unsigned)p.PW;     // one division pair, then +8 pixels per step
                    int x = (int)(qu - row_i * (unsigned)p.PW), y = (int)(row_i % (unsigned)p.PH);
#pragma unroll
                    for (int kq = 0; kq < 4; ++kq) {
                        vm[kq] = (q0r + 8 * kq < p.Q && x < p.VW && y < p.VH) ? 1.f : 0.f;
                        x += 8;
                        if (x >= p.PW) { x -= p.PW; if (++y >= p.PH) y = 0; }
                    }
                    allv = __all_sync(0xffffffffu, vm[0] + vm[1] + vm[2] + vm[3] == 4.f);
                }
                mbar_wait(bars + 8 * (TFULL + buf), (lt >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int i = 0; i < NCF; ++i) {
                    const int c = half + 2 * i;
                    const int sb = slabs & 1;
                    uint32_t r0[16], r1[16];
                    const uint32_t tcol = tmem_base + buf * BN + c * 32;
                    tmem_ld_16x256b_x4(tcol + ((uint32_t)(wq * 32) << 16), r0);
                    tmem_ld_16x256b_x4(tcol + ((uint32_t)(wq * 32 + 16) << 16), r1);
                    float2 bv[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const float* bp = p.bias + n0 + c * 32 + 8 * x + 2 * q4;
                        bv[x] = bias_v ? *reinterpret_cast<const float2*>(bp) : (p.bias ? make_float2(bp[0], bp[1]) : make_float2(0.f, 0.f));
                    }
                    if (et == 0) bulk_wait_read<1>();           // the store that last read this slab buffer is done with it
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    tmem_ld_wait();
                    uint8_t* dbase = slab_gen + sb * SLAB_BYTES + (wq * 32 + g) * 128 + q4 * 4;
#pragma unroll
                    for (int sbk = 0; sbk < 2; ++sbk) {
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const uint32_t* rr = sbk ? r1 : r0;
                            const float v00 = __uint_as_float(rr[4 * x]) + bv[x].x, v01 = __uint_as_float(rr[4 * x + 1]) + bv[x].y;
                            const float v10 = __uint_as_float(rr[4 * x + 2]) + bv[x].x, v11 = __uint_as_float(rr[4 * x + 3]) + bv[x].y;
                            // slab row R = wq * 32 + 16 sbk + 8 e + g (R & 7 == g), 16-byte chunk (c & 1) * 4 + x, swizzled by the row
                            const int ch16 = ((((c & 1) << 2) + x) ^ g) << 4;
                            __nv_bfloat162 h0 = __floats2bfloat162_rn(v00, v01), h1 = __floats2bfloat162_rn(v10, v11);
                            *reinterpret_cast<__nv_bfloat162*>(dbase + (sbk * 16) * 128 + ch16) = h0;
                            *reinterpret_cast<__nv_bfloat162*>(dbase + (sbk * 16 + 8) * 128 + ch16) = h1;
                            if (allv) {
                                a1[i * 8 + 2 * x] += v00 + v10;
                                a1[i * 8 + 2 * x + 1] += v01 + v11;
                                a2[i * 8 + 2 * x] = fmaf(v00, v00, fmaf(v10, v10, a2[i * 8 + 2 * x]));
                                a2[i * 8 + 2 * x + 1] = fmaf(v01, v01, fmaf(v11, v11, a2[i * 8 + 2 * x + 1]));
                            } else if (p.stats) {
                                const float m0 = vm[2 * sbk], m1 = vm[2 * sbk + 1];
                                const float w00 = m0 * v00, w01 = m0 * v01, w10 = m1 * v10, w11 = m1 * v11;
                                a1[i * 8 + 2 * x] += w00 + w10;
                                a1[i * 8 + 2 * x + 1] += w01 + w11;
                                a2[i * 8 + 2 * x] = fmaf(w00, v00, fmaf(w10, v10, a2[i * 8 + 2 * x]));
                                a2[i * 8 + 2 * x + 1] = fmaf(w01, v01, fmaf(w11, v11, a2[i * 8 + 2 * x + 1]));
                            }
                        }
                    }
                    fence_async_smem();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (et == 0) tma_store_2d(&tmO, slab + sb * SLAB_BYTES, n0 + (c >> 1) * 64, (int)q_tile);
                    ++slabs;
                }
                // every epilogue thread has drained its TMEM rows: release the accumulator buffer to the MMA warp
                tc_fence_before();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (et == 0) mbar_arrive_leader(bars + 8 * (TEMPTY + buf));
            }
            if (p.stats && cur_n >= 0) flush(cur_n);
            if (et == 0) bulk_wait_all();                       // every TMA store has completed before the CTA exits
        } else {
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int et = threadIdx.x - EW0 * 32;                 // 0..255: 8 epilogue warps, 2 per TMEM lane group
        const int half = (warp - EW0) >> 2;                    // which interleaved set of 32-column chunks
        float* tile = epi + (warp - EW0) * (32 * 17);          // BN = 256 only
        float* wsum_all = epi + TILE_FLOATS;                   // [4 lane groups][2*BN] running column sums of this CTA
        float* wsum = wsum_all + wq * (2 * BN);
        float a1[RS ? NCH * 32 : 1], a2[RS ? NCH * 32 : 1];    // per-thread running sums (one row, this warp's columns)
        if constexpr (RS) {
#pragma unroll
            for (int j = 0; j < NCH * 32; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
        }
        int lt = 0, cur_n = -1, slabs = 0;
        auto flush = [&](int n_tile) {                         // all 256 epilogue threads
            if constexpr (RS) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    const float s1 = warp_transpose_sum32(a1 + i * 32, lane);     // in place: the accumulators restart at zero
                    const float s2 = warp_transpose_sum32(a2 + i * 32, lane);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { a1[i * 32 + j] = 0.f; a2[i * 32 + j] = 0.f; }
                    wsum[(half + 2 * i) * 32 + lane] = s1;
                    wsum[BN + (half + 2 * i) * 32 + lane] = s2;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int ch = et; ch < BN; ch += 256) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    s1 += wsum_all[w * 2 * BN + ch];
                    s2 += wsum_all[w * 2 * BN + BN + ch];
                    wsum_all[w * 2 * BN + ch] = 0.f;
                    wsum_all[w * 2 * BN + BN + ch] = 0.f;
                }
                atomicAdd(&p.stats[n_tile * BN + ch], (double)s1);
                atomicAdd(&p.stats[p.Cout + n_tile * BN + ch], (double)s2);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        };
        for (int t = cid; t < total; t += nclusters, ++lt) {
            const int n_t = t / num_m, m_t = t - n_t * num_m;
            const int n0 = n_t * BN;
            const int buf = lt & 1;
            if (p.stats && cur_n >= 0 && cur_n != n_t) flush(cur_n);
            cur_n = n_t;
            const long long q_tile = (long long)m_t * 256 + rank * 128;
            const long long q = q_tile + row;
            bool valid = q < p.Q;
            if (valid && p.stats) {
                const unsigned qu = (unsigned)q, row_i = qu / (unsigned)p.PW;       // Q < 2^31 (checked by the launcher): 32-bit divisions
                int x = (int)(qu - row_i * (unsigned)p.PW);
                int y = (int)(row_i % (unsigned)p.PH);
                valid = x < p.VW && y < p.VH;
            }
            mbar_wait(bars + 8 * (TFULL + buf), (lt >> 1) & 1);
            tc_fence_after();
            auto chunk_rs = [&](const int c, const int ci) {   // BN <= 128: two 16-column halves keep the live registers low
                const int sb = slabs & 1;
                if (et == 0) bulk_wait_read<1>();              // the store that last read this slab buffer is done with it
                asm volatile("bar.sync 1, 256;" ::: "memory");
                uint8_t* drow = slab_gen + sb * SLAB_BYTES + row * 128;
                const bool bias_v = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                    uint32_t r[16];
                    tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + buf * BN + c * 32 + hc * 16, r);
                    float v[16];
                    if (bias_v) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + hc * 16);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 bv = bp[g];
                            v[4 * g] = __uint_as_float(r[4 * g]) + bv.x; v[4 * g + 1] = __uint_as_float(r[4 * g + 1]) + bv.y;
                            v[4 * g + 2] = __uint_as_float(r[4 * g + 2]) + bv.z; v[4 * g + 3] = __uint_as_float(r[4 * g + 3]) + bv.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[n0 + c * 32 + hc * 16 + j] : 0.f);
                    }
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint4 u;
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                        *reinterpret_cast<uint4*>(drow + (((half * 4 + hc * 2 + g) ^ (row & 7)) << 4)) = u;
                    }
                    if (p.stats && valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            a1[ci * 32 + hc * 16 + j] += v[j];
                            a2[ci * 32 + hc * 16 + j] = fmaf(v[j], v[j], a2[ci * 32 + hc * 16 + j]);
                        }
                    }
                }
                fence_async_smem();
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (et == 0) tma_store_2d(&tmO, slab + sb * SLAB_BYTES, n0 + (c >> 1) * 64, (int)q_tile);
                ++slabs;
            };
            auto chunk = [&](const int c, const int ci) {      // c: 32-column chunk of the tile, ci: its index among this thread's chunks
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + buf * BN + c * 32, r);
                float v[32];
                if (p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c * 32);
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 bv = bp[g];
                        v[4 * g] = __uint_as_float(r[4 * g]) + bv.x; v[4 * g + 1] = __uint_as_float(r[4 * g + 1]) + bv.y;
                        v[4 * g + 2] = __uint_as_float(r[4 * g + 2]) + bv.z; v[4 * g + 3] = __uint_as_float(r[4 * g + 3]) + bv.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + (p.bias ? p.bias[n0 + c * 32 + j] : 0.f);
                }
                {   // stage this warp's 32 rows x 32 columns into the 128-row x 64-column slab (128-byte swizzle: the
                    // 16-byte chunk j of row r lives at chunk j ^ (r & 7)), then ONE TMA store writes the slab
                    const int sb = slabs & 1;
                    if (et == 0) bulk_wait_read<1>();          // the store that last read this buffer is done with it
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    uint8_t* drow = slab_gen + sb * SLAB_BYTES + row * 128;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 u;
                        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                        *reinterpret_cast<uint4*>(drow + (((half * 4 + g) ^ (row & 7)) << 4)) = u;
                    }
                    fence_async_smem();
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (et == 0) tma_store_2d(&tmO, slab + sb * SLAB_BYTES, n0 + (c >> 1) * 64, (int)q_tile);
                    ++slabs;
                }
                if (p.stats) {
                    // column sums over the warp's 32 rows, 16 columns at a time through a 32x17 shared tile: lane l sums
                    // column (l & 15) over rows [16 (l >> 4), +16), the two row halves are combined with one shuffle
#pragma unroll
                    for (int hc = 0; hc < 2; ++hc) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) tile[lane * 17 + j] = valid ? v[hc * 16 + j] : 0.f;
                        __syncwarp();
                        float s1 = 0.f, s2 = 0.f;
                        const int col = lane & 15, r0 = (lane >> 4) * 16;
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            float tv = tile[(r0 + rr) * 17 + col];
                            s1 += tv;
                            s2 = fmaf(tv, tv, s2);
                        }
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                        if (lane < 16) {
                            wsum[c * 32 + hc * 16 + col] += s1;
                            wsum[BN + c * 32 + hc * 16 + col] += s2;
                        }
                        __syncwarp();
                    }
                }
            };
            if constexpr (RS) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) chunk_rs(half + 2 * i, i);
            } else {
#pragma unroll 1
                for (int c = half; c < BN / 32; c += 2) chunk(c, 0);
            }
            // every epilogue thread has drained its TMEM rows: release the accumulator buffer to the MMA warp
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (et == 0) mbar_arrive_leader(bars + 8 * (TEMPTY + buf));
        }
        if (p.stats && cur_n >= 0) flush(cur_n);
        if (et == 0) bulk_wait_all();                          // every TMA store has completed before the CTA exits
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                         // no remote signal may target a CTA that has exited
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
wgrad_tc_persist_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradTcParams p,
                   int taps) {
    constexpr int BOX_BYTES = 64 * 128;
    constexpr int A_BYTES = 2 * BOX_BYTES;
    constexpr int B_BYTES = (BN / 64) * BOX_BYTES;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = p.Mtot / 128, num_n = p.Ntot / BN;
    // work item = (split, tap, n tile, m tile); m fastest so the CTAs of one wave share the same pixel range
    const int total = num_m * num_n * taps * p.splits;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bars + 8 * (2 * STAGES + b), 1);
            mbar_init(bars + 8 * (2 * STAGES + 2 + b), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int t, int& m0, int& n0, int& tap, long long& qbeg, int& nkb) {
        int mt = t % num_m; t /= num_m;
        int nt = t % num_n; t /= num_n;
        tap = t % taps;
        int split = t / taps;
        m0 = mt * 128; n0 = nt * BN;
        qbeg = (long long)split * p.kchunk;
        long long qend = qbeg + p.kchunk;
        if (qend > p.Q) qend = p.Q;
        nkb = qend > qbeg ? (int)((qend - qbeg + 63) / 64) : 0;
    };

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                int m0, n0, tap, nkb;
                long long qbeg;
                decode(t, m0, n0, tap, qbeg, nkb);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                    const uint32_t full = bars + 8 * s;
                    mbar_expect_tx(full, STAGE_BYTES);
                    const long long q = qbeg + (long long)kb * 64;
                    const uint32_t sa = base + s * STAGE_BYTES;
                    if (p.img) {
                        const int P = (int)(q >> 6), n = P / p.bpi, rem = P - n * p.bpi, yb = rem / p.bpr, xb = rem - yb * p.bpr;
                        const int x0 = xb * p.bx, y0 = yb * p.by;
#pragma unroll
                        for (int b = 0; b < 2; ++b)
                            tma_load_4d(sa + b * BOX_BYTES, &tmA, m0 + b * 64, x0 + p.ax[tap], y0 + p.ay[tap], n, full);
#pragma unroll
                        for (int b = 0; b < BN / 64; ++b)
                            tma_load_4d(sa + A_BYTES + b * BOX_BYTES, &tmB, n0 + b * 64, x0 + p.bxo[tap], y0 + p.byo[tap], n, full);
                        continue;
                    }
#pragma unroll
                    for (int b = 0; b < 2; ++b)
                        tma_load_2d(sa + b * BOX_BYTES, &tmA, m0 + b * 64, (int)(q + p.shiftA[tap]), full);
#pragma unroll
                    for (int b = 0; b < BN / 64; ++b)
                        tma_load_2d(sa + A_BYTES + b * BOX_BYTES, &tmB, n0 + b * 64, (int)(q + p.shiftB[tap]), full);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(BN, 1, 1);
            int it = 0, lt = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
                int m0, n0, tap, nkb;
                long long qbeg;
                decode(t, m0, n0, tap, qbeg, nkb);
                const int buf = lt & 1;
                mbar_wait(bars + 8 * (2 * STAGES + 2 + buf), ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * STAGE_BYTES;
                    const uint64_t ad = make_desc(sa, BOX_BYTES, 1024), bd = make_desc(sa + A_BYTES, BOX_BYTES, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(dcol, ad + 128 * k, bd + 128 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(bars + 8 * (STAGES + s));
                }
                umma_commit(bars + 8 * (2 * STAGES + buf));
            }
        }
    } else {
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int et = threadIdx.x - 64;
        int lt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
            int m0, n0, tap, nkb;
            long long qbeg;
            decode(t, m0, n0, tap, qbeg, nkb);
            const int buf = lt & 1;
            mbar_wait(bars + 8 * (2 * STAGES + buf), (lt >> 1) & 1);
            tc_fence_after();
            if (nkb > 0) {
                float* dst_row = p.stg + ((long long)tap * p.Mtot + m0 + row) * p.Ntot + n0;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + buf * BN + c * 32, r);
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        red_add_v4(dst_row + c * 32 + g * 4, __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                   __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                }
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) mbar_arrive(bars + 8 * (2 * STAGES + 2 + buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

// Row-stacked weight gradient for the layers whose narrow side has 64 channels (64 <-> 128 at 128x128 / 256x256).  With one
// tap per work item every staged byte feeds exactly one 128 x 64 x 16 MMA: 24 KB written by TMA + 24 KB read by the tensor
// core per 128 clocks of MMA = 375 B/clk against the SM's ~128 B/clk of shared-memory bandwidth, so those launches ran at
// 0.26 of the tensor peak (620 TFLOP/s).  Here a work item is a kernel ROW: the 128-channel operand (unshifted: dy for
// 64 -> 128, x for 128 -> 64, where the shift moves to dy with the opposite sign) is staged ONCE and multiplied against the
// three horizontally shifted tiles of the 64-channel operand laid side by side as one N = 192 operand (three MN-major
// 64-channel atoms, LBO = one box) -> 80 KB of shared-memory traffic per 384 clocks = 208 B/clk.
template <int STAGES>
__global__ void __launch_bounds__(192, 1)
wgrad_tc_rows_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradTcParams p) {
    constexpr int BOX_BYTES = 64 * 128;
    constexpr int A_BYTES = 2 * BOX_BYTES;
    constexpr int B_BYTES = 3 * BOX_BYTES;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int NACC = 256;                                  // TMEM column pitch of one accumulator buffer (192 used)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = p.Mtot / 128, num_n = p.Ntot / 64;
    // work item = (split, kernel row, n tile, m tile)
    const int total = num_m * num_n * 3 * p.splits;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bars + 8 * (2 * STAGES + b), 1);
            mbar_init(bars + 8 * (2 * STAGES + 2 + b), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * NACC);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int t, int& m0, int& n0, int& ty, long long& qbeg, int& nkb) {
        int mt = t % num_m; t /= num_m;
        int nt = t % num_n; t /= num_n;
        ty = t % 3;
        int split = t / 3;
        m0 = mt * 128; n0 = nt * 64;
        qbeg = (long long)split * p.kchunk;
        long long qend = qbeg + p.kchunk;
        if (qend > p.Q) qend = p.Q;
        nkb = qend > qbeg ? (int)((qend - qbeg + 63) / 64) : 0;
    };

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                int m0, n0, ty, nkb;
                long long qbeg;
                decode(t, m0, n0, ty, qbeg, nkb);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                    const uint32_t full = bars + 8 * s;
                    mbar_expect_tx(full, STAGE_BYTES);
                    const long long q = qbeg + (long long)kb * 64;
                    const uint32_t sa = base + s * STAGE_BYTES;
#pragma unroll
                    for (int b = 0; b < 2; ++b) tma_load_2d(sa + b * BOX_BYTES, &tmA, m0 + b * 64, (int)q, full);
#pragma unroll
                    for (int tx = 0; tx < 3; ++tx)
                        tma_load_2d(sa + A_BYTES + tx * BOX_BYTES, &tmB, n0, (int)(q + p.shiftB[3 * ty + tx]), full);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(192, 1, 1);
            int it = 0, lt = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
                int m0, n0, ty, nkb;
                long long qbeg;
                decode(t, m0, n0, ty, qbeg, nkb);
                const int buf = lt & 1;
                mbar_wait(bars + 8 * (2 * STAGES + 2 + buf), ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * NACC;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * STAGE_BYTES;
                    const uint64_t ad = make_desc(sa, BOX_BYTES, 1024), bd = make_desc(sa + A_BYTES, BOX_BYTES, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(dcol, ad + 128 * k, bd + 128 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(bars + 8 * (STAGES + s));
                }
                umma_commit(bars + 8 * (2 * STAGES + buf));
            }
        }
    } else {
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int et = threadIdx.x - 64;
        int lt = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++lt) {
            int m0, n0, ty, nkb;
            long long qbeg;
            decode(t, m0, n0, ty, qbeg, nkb);
            const int buf = lt & 1;
            mbar_wait(bars + 8 * (2 * STAGES + buf), (lt >> 1) & 1);
            tc_fence_after();
            if (nkb > 0) {
#pragma unroll 1
                for (int c = 0; c < 6; ++c) {                   // 192 columns = 3 taps x 64 channels, 32 at a time
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + buf * NACC + c * 32, r);
                    float* dst = p.stg + ((long long)(3 * ty + (c >> 1)) * p.Mtot + m0 + row) * p.Ntot + n0 + (c & 1) * 32;
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        red_add_v4(dst + g * 4, __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                   __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                }
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) mbar_arrive(bars + 8 * (2 * STAGES + 2 + buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, 2 * NACC);
    }
}

// CTA-pair weight gradient: D[256 (M channels)][BN (N channels)], each CTA stages its 128 M-channels of one operand and
// BN/2 N-channels of the other (MN-major boxes of 64 channels x 64 pixels); one M=256 MMA stream issued by the leader.
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
wgrad_tc_pair_k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradTcParams p,
                   int taps) {
    constexpr int BOX_BYTES = 64 * 128;
    constexpr int A_BYTES = 2 * BOX_BYTES;
    constexpr int B_BYTES = (BN / 128) * BOX_BYTES;           // this CTA's half of the N operand
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + STAGES * STAGE_BYTES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen_base + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int num_m = p.Mtot / 256, num_n = p.Ntot / BN;
    // work item = (split, tap, n tile, m tile); m fastest so the CTAs of one wave share the same pixel range
    const int total = num_m * num_n * taps * p.splits;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (STAGES + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bars + 8 * (2 * STAGES + b), 1);
            mbar_init(bars + 8 * (2 * STAGES + 2 + b), 2);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nclusters = gridDim.x >> 1, cid = blockIdx.x >> 1;

    auto decode = [&](int t, int& m0, int& n0, int& tap, long long& qbeg, int& nkb) {
        int mt = t % num_m; t /= num_m;
        int nt = t % num_n; t /= num_n;
        tap = t % taps;
        int split = t / taps;
        m0 = mt * 256 + (int)rank * 128; n0 = nt * BN + (int)rank * (BN / 2);
        qbeg = (long long)split * p.kchunk;
        long long qend = qbeg + p.kchunk;
        if (qend > p.Q) qend = p.Q;
        nkb = qend > qbeg ? (int)((qend - qbeg + 63) / 64) : 0;
    };

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int t = cid; t < total; t += nclusters) {
                int m0, n0, tap, nkb;
                long long qbeg;
                decode(t, m0, n0, tap, qbeg, nkb);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * (STAGES + s), ph ^ 1);
                    const uint32_t full = bars + 8 * s;
                    if (leader) mbar_expect_tx(full, 2 * STAGE_BYTES);
                    const long long q = qbeg + (long long)kb * 64;
                    const uint32_t sa = base + s * STAGE_BYTES;
                    if (p.img) {
                        const int P = (int)(q >> 6), n = P / p.bpi, rem = P - n * p.bpi, yb = rem / p.bpr, xb = rem - yb * p.bpr;
                        const int x0 = xb * p.bx, y0 = yb * p.by;
#pragma unroll
                        for (int b = 0; b < 2; ++b)
                            tma_load_4d_pair(sa + b * BOX_BYTES, &tmA, m0 + b * 64, x0 + p.ax[tap], y0 + p.ay[tap], n, full);
#pragma unroll
                        for (int b = 0; b < BN / 128; ++b)
                            tma_load_4d_pair(sa + A_BYTES + b * BOX_BYTES, &tmB, n0 + b * 64, x0 + p.bxo[tap], y0 + p.byo[tap], n, full);
                        continue;
                    }
#pragma unroll
                    for (int b = 0; b < 2; ++b)
                        tma_load_2d_pair(sa + b * BOX_BYTES, &tmA, m0 + b * 64, (int)(q + p.shiftA[tap]), full);
#pragma unroll
                    for (int b = 0; b < BN / 128; ++b)
                        tma_load_2d_pair(sa + A_BYTES + b * BOX_BYTES, &tmB, n0 + b * 64, (int)(q + p.shiftB[tap]), full);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            constexpr uint32_t idesc = make_idesc_m256(BN) | (1u << 15) | (1u << 16);
            int it = 0, lt = 0;
            for (int t = cid; t < total; t += nclusters, ++lt) {
                int m0, n0, tap, nkb;
                long long qbeg;
                decode(t, m0, n0, tap, qbeg, nkb);
                const int buf = lt & 1;
                mbar_wait(bars + 8 * (2 * STAGES + 2 + buf), ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t dcol = tmem_base + buf * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(bars + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t sa = base + s * STAGE_BYTES;
                    const uint64_t ad = make_desc(sa, BOX_BYTES, 1024), bd = make_desc(sa + A_BYTES, BOX_BYTES, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_pair(dcol, ad + 128 * k, bd + 128 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit_pair(bars + 8 * (STAGES + s));
                }
                umma_commit_pair(bars + 8 * (2 * STAGES + buf));
            }
        }
    } else {
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int et = threadIdx.x - 64;
        int lt = 0;
        for (int t = cid; t < total; t += nclusters, ++lt) {
            int m0, n0, tap, nkb;
            long long qbeg;
            decode(t, m0, n0, tap, qbeg, nkb);
            const int buf = lt & 1;
            mbar_wait(bars + 8 * (2 * STAGES + buf), (lt >> 1) & 1);
            tc_fence_after();
            if (nkb > 0) {
                float* dst_row = p.stg + ((long long)tap * p.Mtot + m0 + row) * p.Ntot + (n0 - (int)rank * (BN / 2));
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + buf * BN + c * 32, r);
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        red_add_v4(dst_row + c * 32 + g * 4, __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                   __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                }
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) mbar_arrive_leader(bars + 8 * (2 * STAGES + 2 + buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// staging [t][a][b] -> dw_oihw[co][ci][t] (+=), a/b = (co,ci) or (ci,co)
__global__ void wgrad_finalize_k(const float* __restrict__ stg, float* __restrict__ dw, int taps, int Cout, int Cin,
                                 int CinStg, int m_is_cout) {
    const long long total = (long long)Cout * Cin * taps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int t = (int)(i % taps);
        long long r = i / taps;
        int ci = (int)(r % Cin);
        int co = (int)(r / Cin);
        float v = m_is_cout ? stg[((long long)t * Cout + co) * CinStg + ci] : stg[((long long)t * CinStg + ci) * Cout + co];
        dw[i] += v;
    }
}

struct FoldDesc {
    const float* stg;
    float* dw;
    long long taps, cout, cin, cin_stg, m_is_cout;
};
// every layer's staging gradient -> OIHW bucket in one launch (blockIdx.y = layer)
__global__ void wgrad_finalize_multi_k(const FoldDesc* __restrict__ table) {
    const FoldDesc d = table[blockIdx.y];
    const int taps = (int)d.taps, Cout = (int)d.cout, Cin = (int)d.cin, CinStg = (int)d.cin_stg;
    const unsigned total = (unsigned)(Cout * Cin * taps);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned r = i / (unsigned)taps, t = i - r * (unsigned)taps;
        const unsigned co = r / (unsigned)Cin, ci = r - co * (unsigned)Cin;
        const float v = d.m_is_cout ? d.stg[((long long)t * Cout + co) * CinStg + ci] : d.stg[((long long)t * CinStg + ci) * Cout + co];
        d.dw[i] += v;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)ptr;
    return fn;
}

// 2-D bf16 matrix [rows][cols] (cols contiguous), box = 64 cols x box_rows, 128-byte swizzle
static int make_map(CUtensorMap* tm, const void* ptr, long long rows, long long cols, int box_rows) {
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    const KpMapKey key{ptr, rows, cols, box_rows, -1};
    if (kp_ctx_map_get(key, tm)) return KP_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) { kp_set_error("cuTensorMapEncodeTiled entry point not available"); return KP_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { kp_set_error("cuTensorMapEncodeTiled failed: %d (rows=%lld cols=%lld box=%d)", (int)r, rows, cols, box_rows); return KP_ERR_CUDA; }
    kp_ctx_map_put(key, tm);
    return KP_OK;
}

// 4-D bf16 tensor [N][PH][PW][C] (C contiguous), box = 64 channels x bx x by x 1 image, 128-byte swizzle
static int make_map4(CUtensorMap* tm, const void* ptr, int N, int PH, int PW, int C, int bx, int by) {
    const KpMapKey key{ptr, N, (long long)PH * 65536 + PW, C, (long long)bx * 256 + by};
    if (kp_ctx_map_get(key, tm)) return KP_OK;
    EncodeTiledFn enc = get_encode();
    if (!enc) { kp_set_error("cuTensorMapEncodeTiled entry point not available"); return KP_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)PW, (cuuint64_t)PH, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)PW * C * 2, (cuuint64_t)PH * PW * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { kp_set_error("cuTensorMapEncodeTiled(4d) failed: %d (N=%d PH=%d PW=%d C=%d box=%dx%d)", (int)r, N, PH, PW, C, bx, by); return KP_ERR_CUDA; }
    kp_ctx_map_put(key, tm);
    return KP_OK;
}

template <int BN, int STAGES>
static int launch_conv_persist(cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const ConvTcParams& p) {
    constexpr int smem = STAGES * (STAGE_A_BYTES + BN * 128) + 2 * SLAB_BYTES + (8 * 32 * 17 + 4 * 2 * BN) * 4 +
                         8 * (2 * STAGES + 4) + 16 + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static KpOncePerDevice attr_done;
    if (attr_done.first()) {
        KP_CUDA(cudaFuncSetAttribute(conv_tc_persist_k<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    long long total = ((p.Q + 127) / 128) * (p.Cout / BN);
    int grid = (int)(total < kp_sm_count() ? total : kp_sm_count());
    conv_tc_persist_k<BN, STAGES><<<grid, 320, smem, st>>>(a, b, o, p);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

template <int BN, int STAGES>
static int launch_conv_pair(cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const ConvTcParams& p) {
    constexpr int smem = STAGES * (STAGE_A_BYTES + (BN / 2) * 128) + 2 * SLAB_BYTES + (8 * 32 * 17 + 4 * 2 * BN) * 4 +
                         8 * (2 * STAGES + 4) + 16 + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static KpOncePerDevice attr_done;
    if (attr_done.first()) {
        KP_CUDA(cudaFuncSetAttribute(conv_tc_pair_k<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    long long total = ((p.Q + 255) / 256) * (p.Cout / BN);
    int clusters = kp_sm_count() / 2;
    if (clusters > total) clusters = (int)total;
    if (clusters < 1) clusters = 1;
    conv_tc_pair_k<BN, STAGES><<<2 * clusters, 320, smem, st>>>(a, b, o, p);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

template <int BN, int AS, int BS, bool RB, bool FE>
static int launch_conv_pair3(cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const ConvTcParams& p) {
    constexpr int smem = AS * A3_SLOT + BS * (BN / 2) * 128 + 2 * SLAB_BYTES + (((BN <= 128 || FE) ? 0 : 8 * 32 * 17) + 4 * 2 * BN) * 4 +
                         8 * (2 * AS + 2 * BS + 4) + 16 + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static KpOncePerDevice attr_done;
    if (attr_done.first()) {
        KP_CUDA(cudaFuncSetAttribute(conv_tc_pair3_k<BN, AS, BS, RB, FE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    long long total = ((p.Q + 255) / 256) * (p.Cout / BN);
    int clusters = kp_sm_count() / 2;
    if (clusters > total) clusters = (int)total;
    if (clusters < 1) clusters = 1;
    conv_tc_pair3_k<BN, AS, BS, RB, FE><<<2 * clusters, (BN <= 128 && !FE) ? 384 : 320, smem, st>>>(a, b, o, p);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

template <int BN, int STAGES>
static int launch_wgrad_pair(cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const WgradTcParams& p, int taps) {
    constexpr int smem = STAGES * (2 * 8192 + (BN / 128) * 8192) + 8 * (2 * STAGES + 4) + 16 + 1024;
    static KpOncePerDevice attr_done;
    if (attr_done.first()) {
        KP_CUDA(cudaFuncSetAttribute(wgrad_tc_pair_k<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    long long total = (long long)(p.Mtot / 256) * (p.Ntot / BN) * taps * p.splits;
    int clusters = kp_sm_count() / 2;
    if (clusters > total) clusters = (int)total;
    if (clusters < 1) clusters = 1;
    wgrad_tc_pair_k<BN, STAGES><<<2 * clusters, 192, smem, st>>>(a, b, p, taps);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

template <int BN, int STAGES>
static int launch_wgrad_persist(cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const WgradTcParams& p, int taps) {
    constexpr int smem = STAGES * (2 * 8192 + (BN / 64) * 8192) + 8 * (2 * STAGES + 4) + 16 + 1024;
    static KpOncePerDevice attr_done;
    if (attr_done.first()) {
        KP_CUDA(cudaFuncSetAttribute(wgrad_tc_persist_k<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    long long total = (long long)(p.Mtot / 128) * (p.Ntot / BN) * taps * p.splits;
    int grid = (int)(total < kp_sm_count() ? total : kp_sm_count());
    wgrad_tc_persist_k<BN, STAGES><<<grid, 192, smem, st>>>(a, b, p, taps);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

template <int STAGES>
static int launch_wgrad_rows(cudaStream_t st, const CUtensorMap& a, const CUtensorMap& b, const WgradTcParams& p) {
    constexpr int smem = STAGES * (5 * 8192) + 8 * (2 * STAGES + 4) + 16 + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static KpOncePerDevice attr_done;
    if (attr_done.first()) {
        KP_CUDA(cudaFuncSetAttribute(wgrad_tc_rows_k<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    }
    long long total = (long long)(p.Mtot / 128) * (p.Ntot / 64) * 3 * p.splits;
    int grid = (int)(total < kp_sm_count() ? total : kp_sm_count());
    wgrad_tc_rows_k<STAGES><<<grid, 192, smem, st>>>(a, b, p);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

}  // namespace

// in: bf16 [Q][Cin]; wt: bf16 [taps][Cout][Cin]; out: bf16 [Q][Cout]; shifts: host int[taps];
// stats masked to rows with (q % PW) < VW && (q / PW % PH) < VH.
extern "C" int kp_conv_tc(kp_stream stream, const void* in_bf16, int64_t Q, int Cin, const void* wt_bf16, int taps,
                          const int32_t* shifts, const float* bias, void* out_bf16, int Cout, double* stats, int PH,
                          int PW, int VH, int VW) {
    KP_CHECK_ARG(in_bf16 && wt_bf16 && out_bf16 && shifts && Q > 0 && Q < (1LL << 31) - 4096 && taps >= 1 && taps <= 9 &&
                     Cin % 64 == 0 && Cout % 64 == 0 && Cin > 0 && Cout > 0,
                 "kp_conv_tc: unsupported shape Q=%lld Cin=%d Cout=%d taps=%d", (long long)Q, Cin, Cout, taps);
    KP_CHECK_ARG(!stats || (PH > 0 && PW > 0), "kp_conv_tc: stats need PH/PW");
    int BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0 ? 128 : 64);
    if (BN == 256) {
        // wave quantisation: with few pixel tiles (16x16 layers) 256-wide tiles leave the last round of the persistent
        // CTA pairs mostly idle; 128-wide tiles fill the rounds better
        const long long mt = (Q + 255) / 256, workers = kp_sm_count() / 2;
        auto eff = [&](long long tiles) { return (double)tiles / (double)(((tiles + workers - 1) / workers) * workers); };
        static int q_on = -1;
        if (q_on < 0) { const char* e = kp_env("KP_TC_QUANT"); q_on = (e && e[0] == '0') ? 0 : 1; }
        if (q_on && eff(mt * (Cout / 256)) < 0.8 && eff(mt * (Cout / 128)) > eff(mt * (Cout / 256)) + 0.08) BN = 128;
    }
    CUtensorMap ta, tb;
    int rc = make_map(&ta, in_bf16, Q, Cin, 128);
    if (rc) return rc;
    rc = make_map(&tb, wt_bf16, (long long)taps * Cout, Cin, BN);
    if (rc) return rc;
    ConvTcParams p;
    p.Q = Q; p.Cin = Cin; p.Cout = Cout; p.taps = taps;
    for (int i = 0; i < 9; ++i) p.shift[i] = i < taps ? shifts[i] : 0;
    p.bias = bias; p.out = (bf16*)out_bf16; p.stats = stats; p.PH = PH; p.PW = PW; p.VH = VH; p.VW = VW;
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap to;                                            // epilogue TMA store: 64 channels x 128 pixels per slab
    rc = make_map(&to, out_bf16, Q, Cout, 128);
    if (rc) return rc;
    static int pair_on = -1;
    if (pair_on < 0) { const char* e = kp_env("KP_TC_PAIR"); pair_on = (e && e[0] == '0') ? 0 : 1; }
    if (pair_on && Q >= 256) {
        CUtensorMap tbh;                                       // each CTA of the pair loads half of the weight tile
        rc = make_map(&tbh, wt_bf16, (long long)taps * Cout, Cin, BN / 2);
        if (rc) return rc;
        static int halo_on = -1;
        if (halo_on < 0) { const char* e = kp_env("KP_TC_HALO"); halo_on = (e && e[0] == '0') ? 0 : 1; }
        bool regular = taps == 9;                              // shift[3 ty + tx] = shift[3 ty] + tx
        for (int t = 0; regular && t < 9; ++t) regular = shifts[t] == shifts[3 * (t / 3)] + (t % 3);
        if (halo_on && regular) {
            CUtensorMap ta3;                                   // 128 + 2 pixel rows (padded to 136) per kernel row
            rc = make_map(&ta3, in_bf16, Q, Cin, A3_ROWS);
            if (rc) return rc;
            // Fragment epilogue (tcgen05.ld.16x256b, 4-byte slab stores, statistics in BN / 4 registers per thread) everywhere
            // except the one launch type it loses on: fprop of the 64 -> 128 layers, whose 9-k-block tiles leave the epilogue
            // 2304 clocks and are 10 % faster with one row (one validity test) per thread.  Per launch the fragment form is
            // 1-5 % faster on the dgrads and 2-4 % on the other fprops (gpurun_out/r3_calls_fe*.txt, r3_calls_fex*.txt).
            // Experiments build: KP_TC_FE = 0 never, 1 dgrad only, 2 always, 3 (default) as described.
            static int fe_mode = -1;
            if (fe_mode < 0) { const char* e = kp_env("KP_TC_FE"); fe_mode = e ? atoi(e) : 3; }
            const bool short_stats = stats != nullptr && BN == 128 && Cin == 64;
            const bool fe_on = fe_mode == 2 || (fe_mode == 3 && !short_stats) || (fe_mode == 1 && stats == nullptr);
            // small weight sets stay resident in shared memory (loaded once per CTA): 64 -> 128 and 128 -> 64 (and their
            // dgrads), 72 KB per CTA; the activation ring takes the rest
            static int rb_on = -1;
            if (rb_on < 0) { const char* e = kp_env("KP_TC_RESB"); rb_on = (e && e[0] == '0') ? 0 : 1; }
            const bool rb = rb_on && Cout == BN;
            if (fe_on) {
                if (BN == 256) return launch_conv_pair3<256, 3, 6, false, true>(st, ta3, tbh, to, p);
                if (rb && BN == 128 && Cin == 64) return launch_conv_pair3<128, 5, 9, true, true>(st, ta3, tbh, to, p);
                if (rb && BN == 64 && Cin == 128) return launch_conv_pair3<64, 5, 18, true, true>(st, ta3, tbh, to, p);
                if (BN == 128) return launch_conv_pair3<128, 4, 8, false, true>(st, ta3, tbh, to, p);
                return launch_conv_pair3<64, 5, 12, false, true>(st, ta3, tbh, to, p);
            }
            if (rb && BN == 128 && Cin == 64) return launch_conv_pair3<128, 5, 9, true, false>(st, ta3, tbh, to, p);
#ifdef KP_EXPERIMENTS
            if (BN == 256) return launch_conv_pair3<256, 3, 6, false, false>(st, ta3, tbh, to, p);
            if (rb && BN == 64 && Cin == 128) return launch_conv_pair3<64, 5, 18, true, false>(st, ta3, tbh, to, p);
            if (BN == 128) return launch_conv_pair3<128, 4, 8, false, false>(st, ta3, tbh, to, p);
            return launch_conv_pair3<64, 5, 12, false, false>(st, ta3, tbh, to, p);
#else
            if (BN == 128) return launch_conv_pair3<128, 4, 8, false, false>(st, ta3, tbh, to, p);   // KP_TC_RESB off only
            return launch_conv_pair3<64, 5, 12, false, true>(st, ta3, tbh, to, p);
#endif
        }
        if (BN == 256) return launch_conv_pair<256, 5>(st, ta, tbh, to, p);
        if (BN == 128) return launch_conv_pair<128, 7>(st, ta, tbh, to, p);
        return launch_conv_pair<64, 8>(st, ta, tbh, to, p);
    }
    if (BN == 256) return launch_conv_persist<256, 3>(st, ta, tb, to, p);
    if (BN == 128) return launch_conv_persist<128, 5>(st, ta, tb, to, p);
    return launch_conv_persist<64, 6>(st, ta, tb, to, p);
}

// Split the pixel range so that (tiles * splits) work items fill `workers` persistent CTAs (or clusters) in whole
// rounds: maximise total / (ceil(total / workers) * workers), prefer fewer splits (longer K per item) on ties.
static void choose_split(int tiles, int workers, long long blocks64, long long* kchunk, int* splits) {
    long long max_sp = blocks64 / 8;
    if (max_sp < 1) max_sp = 1;
    long long lim = (4LL * workers + tiles - 1) / tiles;
    if (lim < 1) lim = 1;
    if (max_sp > lim) max_sp = lim;
    double best = -1.0;
    long long best_sp = 1;
    for (long long sp = 1; sp <= max_sp; ++sp) {
        long long per = (blocks64 + sp - 1) / sp;
        long long real = (blocks64 + per - 1) / per;
        long long total = real * tiles;
        long long rounds = (total + workers - 1) / workers;
        double eff = (double)total / (double)(rounds * workers);
        if (total < workers) eff *= 0.5;                     // not even one full round: keep splitting
        if (eff > best + 0.02) { best = eff; best_sp = real; }
    }
    long long per = (blocks64 + best_sp - 1) / best_sp;
    *splits = (int)((blocks64 + per - 1) / per);
    *kchunk = per * 64;
}

// dw_oihw[co][ci][t] += sum_q dy[q][co] * x[q + shifts[t]][ci];  x: bf16 [Q][CinP] (CinP >= Cin, extra
// channels ignored), dy: bf16 [Q][Cout] with zero rows wherever the product must not count.
// stg: fp32 workspace of taps*Cout*CinP floats (zeroed here).
// img != nullptr: image mode {N, H, W} (contraction over the valid pixels through 4-D tensor maps)
static int wgrad_tc_impl(kp_stream stream, const void* x_bf16, const void* dy_bf16, int64_t Q, int Cin, int CinP,
                         int Cout, int taps, const int32_t* shifts, float* stg, float* dw_oihw, const int* img) {
    KP_CHECK_ARG(x_bf16 && dy_bf16 && stg && (shifts || img) && (img || (Q > 0 && Q < (1LL << 31) - 4096)) && taps >= 1 &&
                     taps <= 9 && CinP % 64 == 0 && Cout % 64 == 0 && Cin <= CinP && (Cout % 128 == 0 || CinP % 128 == 0),
                 "kp_conv_wgrad_tc: unsupported shape Q=%lld Cin=%d/%d Cout=%d taps=%d", (long long)Q, Cin, CinP, Cout, taps);
    cudaStream_t st = (cudaStream_t)stream;
    const bool m_is_cout = (Cout % 128 == 0);
    const int Mtot = m_is_cout ? Cout : CinP, Ntot = m_is_cout ? CinP : Cout;
    const int BN = (Ntot % 256 == 0) ? 256 : (Ntot % 128 == 0 ? 128 : 64);
    CUtensorMap ta, tb;
    int rc;
    WgradTcParams p;
    p.Mtot = Mtot; p.Ntot = Ntot; p.stg = stg; p.img = 0;
    p.bx = p.by = p.bpr = p.bpi = 0;
    for (int i = 0; i < 9; ++i) { p.ax[i] = p.ay[i] = p.bxo[i] = p.byo[i] = 0; }
    if (img) {
        const int N = img[0], H = img[1], W = img[2], PH = H + 2, PW = W + 2;
        p.img = 1;
        p.bx = W < 64 ? W : 64; p.by = 64 / p.bx;
        p.bpr = W / p.bx; p.bpi = p.bpr * (H / p.by);
        Q = (int64_t)N * p.bpi * 64;                       // = N * H * W valid pixels
        const int ks = taps == 9 ? 3 : 1;
        for (int t = 0; t < taps; ++t) {
            const int xo = ks == 3 ? t % 3 : 1, yo = ks == 3 ? t / 3 : 1;     // x window of output pixel (y, x) starts at (y, x)
            p.ax[t] = m_is_cout ? 1 : xo; p.ay[t] = m_is_cout ? 1 : yo;      // dy is interior-aligned: (+1, +1)
            p.bxo[t] = m_is_cout ? xo : 1; p.byo[t] = m_is_cout ? yo : 1;
        }
        rc = make_map4(&ta, m_is_cout ? dy_bf16 : x_bf16, N, PH, PW, Mtot, p.bx, p.by);
        if (rc) return rc;
        rc = make_map4(&tb, m_is_cout ? x_bf16 : dy_bf16, N, PH, PW, Ntot, p.bx, p.by);
        if (rc) return rc;
    } else {
        rc = make_map(&ta, m_is_cout ? dy_bf16 : x_bf16, Q, Mtot, 64);
        if (rc) return rc;
        rc = make_map(&tb, m_is_cout ? x_bf16 : dy_bf16, Q, Ntot, 64);
        if (rc) return rc;
    }
    p.Q = Q;
    for (int i = 0; i < 9; ++i) {
        int s = (i < taps && shifts) ? shifts[i] : 0;
        p.shiftA[i] = m_is_cout ? 0 : s;
        p.shiftB[i] = m_is_cout ? s : 0;
    }
    const int tiles = (Mtot / 128) * (Ntot / BN) * taps;
    long long blocks64 = (Q + 63) / 64;
    choose_split(tiles, kp_sm_count(), blocks64, &p.kchunk, &p.splits);
    static int rows_on = -1;
    if (rows_on < 0) { const char* e = kp_env("KP_WGRAD_ROWS"); rows_on = (e && e[0] == '0') ? 0 : 1; }
    const bool rows = rows_on && !img && BN == 64 && taps == 9;
    if (rows) {
        // row-stacked form: the 128-channel operand is staged unshifted, the 64-channel one carries the (relative) shift
        for (int i = 0; i < 9; ++i) { p.shiftB[i] = p.shiftB[i] - p.shiftA[i]; p.shiftA[i] = 0; }
        choose_split((Mtot / 128) * (Ntot / 64) * 3, kp_sm_count(), blocks64, &p.kchunk, &p.splits);
    }
    // dw_oihw == NULL: deferred mode — the caller zeroes its staging arena once per step and folds every layer with one
    // kp_wgrad_finalize_multi launch (saves a memset and a fold launch per layer)
    if (dw_oihw) KP_CUDA(cudaMemsetAsync(stg, 0, sizeof(float) * (size_t)taps * Mtot * Ntot, st));
    static int wpair_on = -1;
    if (wpair_on < 0) { const char* e = kp_env("KP_TC_PAIR"); wpair_on = (e && e[0] == '0') ? 0 : 1; }
    if (wpair_on && Mtot % 256 == 0 && BN >= 128) {
        // recompute the split for 256-row tiles
        const int tiles2 = (Mtot / 256) * (Ntot / BN) * taps;
        choose_split(tiles2, kp_sm_count() / 2, blocks64, &p.kchunk, &p.splits);
        if (BN == 256) rc = launch_wgrad_pair<256, 6>(st, ta, tb, p, taps);
        else rc = launch_wgrad_pair<128, 8>(st, ta, tb, p, taps);
    } else if (rows) rc = launch_wgrad_rows<5>(st, ta, tb, p);
    else if (BN == 256) rc = launch_wgrad_persist<256, 4>(st, ta, tb, p, taps);
    else if (BN == 128) rc = launch_wgrad_persist<128, 6>(st, ta, tb, p, taps);
    else rc = launch_wgrad_persist<64, 8>(st, ta, tb, p, taps);
    if (rc) return rc;
    if (!dw_oihw) return KP_OK;
    long long total = (long long)Cout * Cin * taps;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kp_sm_count() * 8) blocks = kp_sm_count() * 8;
    wgrad_finalize_k<<<blocks, 256, 0, st>>>(stg, dw_oihw, taps, Cout, Cin, CinP, m_is_cout ? 1 : 0);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

extern "C" int kp_conv_wgrad_tc(kp_stream stream, const void* x_bf16, const void* dy_bf16, int64_t Q, int Cin, int CinP,
                                int Cout, int taps, const int32_t* shifts, float* stg, float* dw_oihw) {
    return wgrad_tc_impl(stream, x_bf16, dy_bf16, Q, Cin, CinP, Cout, taps, shifts, stg, dw_oihw, nullptr);
}

// Image-aware form: x = replicate-padded input [N][H+2][W+2][CinP], dy = interior-aligned gradient [N][H+2][W+2][Cout];
// only the N*H*W valid pixels are contracted.  Needs W a power of two >= 16 (W <= 64: 64 % W == 0 and H % (64 / W) == 0).
extern "C" int kp_conv_wgrad_tc_img(kp_stream stream, const void* x_bf16, const void* dy_bf16, int N, int H, int W, int Cin,
                                    int CinP, int Cout, int ks, float* stg, float* dw_oihw) {
    KP_CHECK_ARG(N > 0 && H > 0 && W >= 16 && (W & (W - 1)) == 0 && (ks == 1 || ks == 3) &&
                     (W >= 64 || H % (64 / W) == 0) && (long long)N * H * W < (1LL << 31),
                 "kp_conv_wgrad_tc_img: unsupported image %dx%dx%d k%d", N, H, W, ks);
    const int img[3] = {N, H, W};
    return wgrad_tc_impl(stream, x_bf16, dy_bf16, 0, Cin, CinP, Cout, ks * ks, nullptr, stg, dw_oihw, img);
}

// table_dev: n_layers rows of 7 int64 {stg, dw_oihw, taps, Cout, Cin, CinP, m_is_cout (= Cout % 128 == 0)} in device memory
extern "C" int kp_wgrad_finalize_multi(kp_stream stream, const void* table_dev, int n_layers) {
    KP_CHECK_ARG(table_dev && n_layers > 0 && n_layers <= 65535, "kp_wgrad_finalize_multi: bad arguments");
    dim3 grid(64u, (unsigned)n_layers, 1);
    wgrad_finalize_multi_k<<<grid, 256, 0, (cudaStream_t)stream>>>((const FoldDesc*)table_dev);
    KP_LAUNCH_CHECK();
    return KP_OK;
}
