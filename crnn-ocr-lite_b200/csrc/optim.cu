// optim.cu -- optimiser step of the reference (train.py:187-192; Keras 2.2.2 Adam / SGD-Nesterov with clipnorm=5,
// SURVEY A.5) as two launches over the flat parameter / gradient arenas: global sum of squares, then a fused
// clip-by-global-norm + update.  `gscale` folds the 1/world_size of the data-parallel gradient average.
#include "common.cuh"
#include "kernels.h"

namespace {
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out)
{ pdl_enter();
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = (double)g[i]; s = fma(v, v, s);
    }
    __shared__ double red[32];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        atomicAdd(out, t);
    }
}
__device__ __forceinline__ float clipped(float g, float gscale, float norm, float clipnorm)
{
    g *= gscale;
    if (clipnorm > 0.f && norm >= clipnorm) g = g * clipnorm / norm;   // Keras clip_norm: g * c / n
    return g;
}
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, const double* __restrict__ sumsq, float clipnorm, float lr_t, float b1, float b2, float eps, float gscale)
{ pdl_enter();
    const float norm = (float)(sqrt(*sumsq) * (double)gscale);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = clipped(g[i], gscale, norm, clipnorm);
        float mi = b1 * m[i] + (1.f - b1) * gi;
        float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        w[i] = w[i] - lr_t * mi / (sqrtf(vi) + eps);
    }
}
__global__ void sgd_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ vel, long long n,
                           const double* __restrict__ sumsq, float clipnorm, float lr_i, float mom, float gscale)
{ pdl_enter();
    const float norm = (float)(sqrt(*sumsq) * (double)gscale);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = clipped(g[i], gscale, norm, clipnorm);
        float ve = mom * vel[i] - lr_i * gi;
        vel[i] = ve;
        w[i] = w[i] + mom * ve - lr_i * gi;
    }
}
}  // namespace

int launch_sumsq(const float* g, long long n, double* out, cudaStream_t st) {
    (void)crnn_launch(sumsq_kernel, 148 * 4, 256, 0, st, g, n, out); LAUNCH_CHECK(); return CRNN_OK;
}
int launch_adam(float* w, const float* g, float* m, float* v, long long n, const double* sumsq, float clipnorm,
                float lr_t, float b1, float b2, float eps, float gscale, cudaStream_t st) {
    (void)crnn_launch(adam_kernel, 148 * 4, 256, 0, st, w, g, m, v, n, sumsq, clipnorm, lr_t, b1, b2, eps, gscale); LAUNCH_CHECK(); return CRNN_OK;
}
int launch_sgd_nesterov(float* w, const float* g, float* vel, long long n, const double* sumsq, float clipnorm,
                        float lr_i, float momentum, float gscale, cudaStream_t st) {
    (void)crnn_launch(sgd_kernel, 148 * 4, 256, 0, st, w, g, vel, n, sumsq, clipnorm, lr_i, momentum, gscale); LAUNCH_CHECK(); return CRNN_OK;
}
