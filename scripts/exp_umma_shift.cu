// Experiment: can a SWIZZLE_128B K-major UMMA operand start at a row that is NOT a multiple of 8 (1024 B)?  If yes, one
// TMA-loaded activation tile of 128+2 rows serves the three horizontal taps of a 3x3 convolution (A-halo reuse).
// Result on B200 (gpurun_out/exp_umma_shift.log, profiles/r1_umma_row_shift.md): EXACT for every shift 0..7 with the
// descriptor's base-offset field = 0 (the swizzle is a function of the absolute smem address); WRONG when the field is set
// to the phase ((addr >> 7) & 7).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/exp_umma_shift scripts/exp_umma_shift.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef __nv_bfloat16 bf16;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

constexpr int ROWS = 144, BOXR = 136, N = 64;

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int use_base) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 18432, bars = base + 18432 + 8192;
    uint32_t* slot = reinterpret_cast<uint32_t*>(raw + (bars + 64 - smem_u32(raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bars, BOXR * 128 + N * 128);
        tma_load_2d(sA, &tmA, 0, 0, bars);
        tma_load_2d(sB, &tmB, 0, 0, bars);
        mbar_wait(bars, 0);
    }
    __syncthreads();
    for (int s = 0; s < 8; ++s) {
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t ad = make_desc(sA + s * 128, 16, 1024, use_base ? s : 0), bd = make_desc(sB, 16, 1024, 0);
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t acc = kk ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 2 * kk), "l"(bd + 2 * kk), "r"(make_idesc(N)), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 8) : "memory");
        }
        mbar_wait(bars + 8, s & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c * 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) out[((size_t)s * 128 + warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeTiledFn enc, CUtensorMap* tm, void* ptr, int rows, int box_rows) {
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}


// ---- MN-major variant (the wgrad operands): A = [K = 72 pixel rows][M = 128 channels] as two 64-channel boxes, B = [K = 64][N = 64]
// D_s[m][n] = sum_k A[k + s][m] * B[k][n]; the shift is s rows (128 B) into each A box
__global__ void __launch_bounds__(128) kmn(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    constexpr uint32_t ABOX = 72 * 128;                     // 9216 bytes per 64-channel box
    const uint32_t sA = base, sB = base + 2 * ABOX, bars = sB + 8192;
    uint32_t* slot = reinterpret_cast<uint32_t*>(raw + (bars + 64 - smem_u32(raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bars, 2 * ABOX + 8192);
        tma_load_2d(sA, &tmA, 0, 0, bars);
        tma_load_2d(sA + ABOX, &tmA, 64, 0, bars);
        tma_load_2d(sB, &tmB, 0, 0, bars);
        mbar_wait(bars, 0);
    }
    __syncthreads();
    const uint32_t idesc = make_idesc(N) | (1u << 15) | (1u << 16);      // both operands MN-major
    for (int s = 0; s < 8; ++s) {
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t ad = make_desc(sA + s * 128, ABOX, 1024, 0), bd = make_desc(sB, 8192, 1024, 0);
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t acc = kk ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 128 * kk), "l"(bd + 128 * kk), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bars + 8) : "memory");
        }
        mbar_wait(bars + 8, s & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c * 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) out[((size_t)s * 128 + warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(r[j]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

static void make_map_w(EncodeTiledFn enc, CUtensorMap* tm, void* ptr, int rows, int cols, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

static void run_mn(EncodeTiledFn enc) {
    const int KR = 80;                                      // pixel rows of A in global memory
    std::vector<bf16> hA(KR * 128), hB(64 * 64);
    std::vector<float> fA(KR * 128), fB(64 * 64);
    srand(2);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (float)(rand() % 17 - 8); hA[i] = __float2bfloat16(v); fA[i] = v; }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (float)(rand() % 9 - 4); hB[i] = __float2bfloat16(v); fB[i] = v; }
    bf16 *dA, *dB;
    float* dO;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dO, 8 * 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tA, tB;
    make_map_w(enc, &tA, dA, KR, 128, 72);
    make_map_w(enc, &tB, dB, 64, 64, 64);
    const int smem = 2 * 9216 + 8192 + 256 + 1024;
    CK(cudaFuncSetAttribute(kmn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaMemset(dO, 0, 8 * 128 * N * 4));
    kmn<<<1, 128, smem>>>(tA, tB, dO);
    CK(cudaDeviceSynchronize());
    std::vector<float> hO(8 * 128 * N);
    CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
    for (int s = 0; s < 8; ++s) {
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int kk = 0; kk < 64; ++kk) ref += fA[(kk + s) * 128 + m] * fB[kk * 64 + n];
                double e = fabs(ref - hO[((size_t)s * 128 + m) * N + n]);
                if (e > maxerr) maxerr = e;
            }
        printf("MN-major operand, shift=%d rows  max|err|=%g %s\n", s, maxerr, maxerr == 0 ? "EXACT" : "WRONG");
    }
}

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    std::vector<bf16> hA(ROWS * 64), hB(N * 64);
    std::vector<float> fA(ROWS * 64), fB(N * 64);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { float v = (float)(rand() % 17 - 8); hA[i] = __float2bfloat16(v); fA[i] = v; }
    for (size_t i = 0; i < hB.size(); ++i) { float v = (float)(rand() % 9 - 4); hB[i] = __float2bfloat16(v); fB[i] = v; }
    bf16 *dA, *dB;
    float* dO;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dO, 8 * 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tA, tB;
    make_map(enc, &tA, dA, ROWS, BOXR);
    make_map(enc, &tB, dB, N, N);
    const int smem = 18432 + 8192 + 256 + 1024;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int use_base = 0; use_base < 2; ++use_base) {
        CK(cudaMemset(dO, 0, 8 * 128 * N * 4));
        k<<<1, 128, smem>>>(tA, tB, dO, use_base);
        CK(cudaDeviceSynchronize());
        std::vector<float> hO(8 * 128 * N);
        CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
        for (int s = 0; s < 8; ++s) {
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int kk = 0; kk < 64; ++kk) ref += fA[(m + s) * 64 + kk] * fB[n * 64 + kk];
                    double e = fabs(ref - hO[((size_t)s * 128 + m) * N + n]);
                    if (e > maxerr) maxerr = e;
                }
            printf("base_offset_field=%s shift=%d rows  max|err|=%g %s\n", use_base ? "phase" : "zero ", s, maxerr, maxerr == 0 ? "EXACT" : "WRONG");
        }
    }
    run_mn(enc);
    return 0;
}
