// Fused Adam over the flat fp32 parameter bucket (torch.optim.Adam defaults, transporter.py:47).
#include "kp_common.cuh"

namespace {
__global__ void adam_tick_k(int* step_dev) { *step_dev += 1; }

__global__ void loss_ring_push_k(const double* __restrict__ loss_sum, double scale, const int* __restrict__ step_dev,
                                 double* __restrict__ ring, int slots) {
    const int step = step_dev ? *step_dev : 0;
    double* slot = ring + 2 * (step % slots);
    slot[0] = (double)step;
    slot[1] = *loss_sum * scale;
}

__global__ void __launch_bounds__(256) adam_k(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, long long n, float lr, float b1, float b2,
                                              float eps, float bc1, float bc2_sqrt, float gscale, float omb1, float omb2,
                                              const int* __restrict__ step_dev) {
    if (step_dev) {   // CUDA-graph friendly: the step count lives on the device
        double t = (double)*step_dev;
        bc1 = (float)(1.0 - pow((double)b1, t));
        bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        float mi = b1 * m[i] + omb1 * gi;
        float vi = b2 * v[i] + omb2 * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}
}  // namespace

extern "C" int kp_adam_step(kp_stream stream, float* p, const float* g, float* m, float* v, int64_t n, double lr,
                            double beta1, double beta2, double eps, int step, float grad_scale, int32_t* step_dev) {
    KP_CHECK_ARG(p && g && m && v && n > 0 && (step > 0 || step_dev), "kp_adam_step: bad arguments");
    if (step_dev) adam_tick_k<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    if (step <= 0) step = 1;
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)kp_sm_count() * 8) blocks = (long long)kp_sm_count() * 8;
    adam_k<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)lr, (float)beta1, (float)beta2, (float)eps, (float)bc1,
                                                         (float)sqrt(bc2), grad_scale, (float)(1.0 - beta1),
                                                         (float)(1.0 - beta2), step_dev);
    KP_LAUNCH_CHECK();
    return KP_OK;
}

extern "C" int kp_loss_ring_push(kp_stream stream, const double* loss_sum, double scale, const int32_t* step_dev, double* ring,
                                 int slots) {
    KP_CHECK_ARG(loss_sum && ring && slots > 0, "kp_loss_ring_push: bad arguments");
    loss_ring_push_k<<<1, 1, 0, (cudaStream_t)stream>>>(loss_sum, scale, step_dev, ring, slots);
    KP_LAUNCH_CHECK();
    return KP_OK;
}
