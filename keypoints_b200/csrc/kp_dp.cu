// Data-parallel optimiser step over NVLink peer memory: gradient reduce-scatter + Adam + parameter all-gather
// fused into ONE kernel.  Every rank owns the slice [rank*chunk, (rank+1)*chunk) of the flat parameter bucket: it reads
// that slice of every replica's gradient buffer (peer loads, or one multimem.ld_reduce through the NVSwitch), applies
// Adam to its slice of the moments, and writes the updated parameters into every replica (peer stores, or one
// multimem.st).  Per step and GPU this moves ~ (N-1)/N * 4 bytes/parameter in each direction instead of the
// 2 * (N-1)/N * 4 of a ring all-reduce, needs no staging buffer, and the Adam arithmetic runs once per parameter
// instead of once per parameter per replica.  Two tiny barrier kernels (flags in peer memory) bracket it.
// The reference has no distributed code; this replaces an NCCL all-reduce + a replicated Adam.
#include "kp_common.cuh"

namespace {

struct Peers {
    float* g[KP_DP_MAX_WORLD];
    float* p[KP_DP_MAX_WORLD];
    int* flag[KP_DP_MAX_WORLD];
    float* mc_g;
    float* mc_p;
};

// Barrier over the replicas: rank r writes the new epoch into slot [r] of every peer's flag array (release, system
// scope) and waits until all slots of its own array have reached it (acquire).  Epochs only grow, so a fast rank
// signalling the next barrier early is harmless.
__global__ void dp_barrier_k(Peers peers, int rank, int world, int* __restrict__ epoch_ptr) {
    __shared__ int e_s;
    if (threadIdx.x == 0) e_s = *epoch_ptr + 1;
    __syncthreads();
    const int e = e_s;
    const int j = threadIdx.x;
    if (j < world) {
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peers.flag[j] + rank), "r"(e) : "memory");
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(peers.flag[rank] + j) : "memory");
        } while (v < e);
    }
    __syncthreads();
    if (threadIdx.x == 0) *epoch_ptr = e;
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ void mc_st(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float adam1(float g, float& m, float& v, float p, float lr_bc1, float b1, float b2, float omb1,
                                       float omb2, float bc2_sqrt, float eps) {
    m = b1 * m + omb1 * g;
    v = b2 * v + omb2 * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    return p - lr_bc1 * (m / denom);
}

template <bool MC>
__global__ void __launch_bounds__(256) dp_adam_k(Peers peers, int rank, int world, long long lo4, long long hi4,
                                                 float* __restrict__ m, float* __restrict__ v, float lr, float b1, float b2,
                                                 float eps, float gscale, const int* __restrict__ step_dev) {
    const double t = (double)*step_dev;
    const float bc1 = (float)(1.0 - pow((double)b1, t));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
    const float lr_bc1 = lr / bc1, omb1 = 1.f - b1, omb2 = 1.f - b2;
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    const float4* own_p = reinterpret_cast<const float4*>(peers.p[rank]);
    for (long long i = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += (long long)gridDim.x * blockDim.x) {
        float4 g;
        if (MC) {
            g = mc_ld_reduce(peers.mc_g + 4 * i);
        } else {
            float4 part[KP_DP_MAX_WORLD];
#pragma unroll
            for (int j = 0; j < KP_DP_MAX_WORLD; ++j)
                if (j < world) part[j] = reinterpret_cast<const float4*>(peers.g[j])[i];       // all peer loads in flight at once
            g = part[0];
#pragma unroll
            for (int j = 1; j < KP_DP_MAX_WORLD; ++j)
                if (j < world) { g.x += part[j].x; g.y += part[j].y; g.z += part[j].z; g.w += part[j].w; }
        }
        float4 mm = m4[i], vv = v4[i], pp = own_p[i];
        pp.x = adam1(g.x * gscale, mm.x, vv.x, pp.x, lr_bc1, b1, b2, omb1, omb2, bc2_sqrt, eps);
        pp.y = adam1(g.y * gscale, mm.y, vv.y, pp.y, lr_bc1, b1, b2, omb1, omb2, bc2_sqrt, eps);
        pp.z = adam1(g.z * gscale, mm.z, vv.z, pp.z, lr_bc1, b1, b2, omb1, omb2, bc2_sqrt, eps);
        pp.w = adam1(g.w * gscale, mm.w, vv.w, pp.w, lr_bc1, b1, b2, omb1, omb2, bc2_sqrt, eps);
        m4[i] = mm;
        v4[i] = vv;
        if (MC) {
            mc_st(peers.mc_p + 4 * i, pp);
        } else {
#pragma unroll
            for (int j = 0; j < KP_DP_MAX_WORLD; ++j)
                if (j < world) reinterpret_cast<float4*>(peers.p[j])[i] = pp;
        }
    }
    __threadfence_system();          // this thread's peer stores are performed before the kernel (and the barrier after it) ends
}

__global__ void dp_tick_k(int* step_dev) { *step_dev += 1; }

}  // namespace

extern "C" int kp_dp_adam_step(kp_stream stream, const kp_dp_peers* peers, int rank, int world, int64_t n, float* m, float* v,
                               double lr, double beta1, double beta2, double eps, float grad_scale, int32_t* step_dev,
                               int32_t* epoch_dev, int use_multicast) {
    KP_CHECK_ARG(peers && m && v && step_dev && epoch_dev && n > 0 && (n % 4) == 0 && world >= 1 && world <= KP_DP_MAX_WORLD &&
                     rank >= 0 && rank < world,
                 "kp_dp_adam_step: bad arguments");
    KP_CHECK_ARG(!use_multicast || (peers->mc_g && peers->mc_p), "kp_dp_adam_step: multicast pointers missing");
    Peers pr;
    for (int j = 0; j < KP_DP_MAX_WORLD; ++j) {
        pr.g[j] = j < world ? peers->g[j] : nullptr;
        pr.p[j] = j < world ? peers->p[j] : nullptr;
        pr.flag[j] = j < world ? peers->flag[j] : nullptr;
        KP_CHECK_ARG(j >= world || (pr.g[j] && pr.p[j] && pr.flag[j]), "kp_dp_adam_step: null peer pointer");
    }
    pr.mc_g = peers->mc_g;
    pr.mc_p = peers->mc_p;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n4 = n / 4;
    const long long chunk4 = (n4 + world - 1) / world;
    const long long lo4 = (long long)rank * chunk4;
    const long long hi4 = lo4 + chunk4 < n4 ? lo4 + chunk4 : n4;
    dp_tick_k<<<1, 1, 0, st>>>(step_dev);
    dp_barrier_k<<<1, 32, 0, st>>>(pr, rank, world, epoch_dev);            // every replica's gradients are complete
    if (hi4 > lo4) {
        long long blocks = (hi4 - lo4 + 255) / 256;
        const long long cap = (long long)kp_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        if (use_multicast)
            dp_adam_k<true><<<(int)blocks, 256, 0, st>>>(pr, rank, world, lo4, hi4, m, v, (float)lr, (float)beta1, (float)beta2,
                                                         (float)eps, grad_scale, step_dev);
        else
            dp_adam_k<false><<<(int)blocks, 256, 0, st>>>(pr, rank, world, lo4, hi4, m, v, (float)lr, (float)beta1, (float)beta2,
                                                          (float)eps, grad_scale, step_dev);
    }
    dp_barrier_k<<<1, 32, 0, st>>>(pr, rank, world, epoch_dev);            // every slice has been read and every replica written
    KP_LAUNCH_CHECK();
    return KP_OK;
}
