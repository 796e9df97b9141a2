// Probe (B200): register layout of tcgen05.ld.16x256b against the known 32x32b layout (lane = row, register j = column j).
// One CTA, 4 warps; TMEM[row][col] = row * 256 + col is written with tcgen05.st.32x32b.x32, then read back with 16x256b.x1 / .x2
// at lane offsets 0 and 16 of each warp's quarter.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/_bin/exp_tmem_ld_shapes scripts/exp_tmem_ld_shapes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    const uint32_t row = warp * 32 + lane;
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = row * 256 + j;
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int half = 0; half < 2; ++half) {
        uint32_t r[8];
        const uint32_t ta = tb + ((uint32_t)(warp * 32 + half * 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) out[((warp * 2 + half) * 32 + lane) * 8 + j] = r[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64) : "memory");
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 4 * 2 * 32 * 8 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    static uint32_t h[4 * 2 * 32 * 8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int w = 0; w < 4; ++w)
        for (int half = 0; half < 2; ++half)
            for (int l = 0; l < 32; ++l) {
                const uint32_t* r = h + ((w * 2 + half) * 32 + l) * 8;
                if (w == 1 && (l < 6 || l > 29)) {
                    printf("warp %d half %d lane %2d:", w, half, l);
                    for (int j = 0; j < 8; ++j) printf(" (r%u,c%u)", r[j] / 256, r[j] % 256);
                    printf("\n");
                }
                // hypothesis: r[4x + 0..1] = row base + l/4, cols 8x + 2(l%4) + {0,1}; r[4x + 2..3] = row + 8
                for (int x = 0; x < 2; ++x)
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t row = w * 32 + half * 16 + l / 4 + (e >> 1) * 8, col = 8 * x + 2 * (l % 4) + (e & 1);
                        if (r[4 * x + e] != row * 256 + col) ++bad;
                    }
            }
    printf("hypothesis (mma C-fragment layout per 8 columns) mismatches: %d\n", bad);
    return 0;
}
