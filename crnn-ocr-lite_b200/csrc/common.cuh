// common.cuh -- shared helpers for the sm_100a kernels of the CRNN-OCR hot path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#ifndef CRNN_OK   // same values as include/crnn_b200.h
#define CRNN_OK 0
#define CRNN_ERR_INVALID (-1)
#define CRNN_ERR_CUDA (-2)
#define CRNN_ERR_NOMEM (-3)
#define CRNN_ERR_UNKNOWN_NAME (-4)
#define CRNN_ERR_INFEASIBLE (-5)
#endif

void crnn_set_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            crnn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return CRNN_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

extern long long g_crnn_launches;   // kernels launched by this library (engine.cu)
#define LAUNCH_CHECK() do { ++g_crnn_launches; CUDA_TRY(cudaGetLastError()); } while (0)
// kernel family of the launch being issued (per-kernel roofline of bench.py: one kernel serves several stages of the step); a launcher of
// a tracked kernel sets it, the engine's profiling scope reads it back when the launcher returns
enum { CRNN_FAM_OTHER = 0, CRNN_FAM_XW_TC, CRNN_FAM_XTY_TC, CRNN_FAM_RNN_MMA, CRNN_FAM_DWROWS, CRNN_FAM_COUNT };
extern int g_crnn_family;

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
#define NEG_INF (__int_as_float(0xff800000))

__device__ __forceinline__ float lse2(float a, float b) {
    // TF ctc_loss_util.h LogSumExp, fp32, accurate (non fast-math) expf/log1pf
    if (a == NEG_INF && b == NEG_INF) return NEG_INF;
    return (a > b) ? a + log1pf(expf(b - a)) : b + log1pf(expf(a - b));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.f), 6.f); }
__device__ __forceinline__ float hard_sigmoid(float v) { return fminf(fmaxf(__fadd_rn(__fmul_rn(0.2f, v), 0.5f), 0.f), 1.f); }

// Stateless dropout RNG: keep mask for element `idx` of dropout layer `layer` at optimiser step `step`.
// (murmur3-style 64-bit finaliser; the same function is restated in tests for dropout-on parity.)
__host__ __device__ __forceinline__ uint32_t crnn_hash(uint64_t seed, uint32_t layer, uint64_t idx) {
    uint64_t x = seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(layer + 1)) ^ (idx * 0xD6E8FEB86659FD93ull);
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (uint32_t)(x >> 32);
}
// Stateless dropout masks.  Element i belongs to the 4-element group i >> 2; one 2 x 32-bit mix of (seed, layer, group) yields four
// 16-bit uniform fields, element i keeps its value iff field[i & 3] >= rate * 65536 (effective rate quantised to 1/65536).  One mix per
// float4 instead of one 64-bit hash per element: the elementwise kernels were instruction-bound on the old per-element hash
// (ncu r1e: 382 warp instructions per float4 in act_pool_fwd).
__host__ __device__ __forceinline__ uint32_t crnn_mix32(uint32_t x) {          // "lowbias32" finaliser
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ void crnn_dropout_bits(uint64_t seed, uint32_t layer, uint64_t group, uint32_t& lo, uint32_t& hi) {
    const uint32_t k = (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u) ^ ((layer + 1u) * 0x85EBCA6Bu);
    const uint32_t g = (uint32_t)group ^ ((uint32_t)(group >> 32) * 0xC2B2AE35u);
    lo = crnn_mix32(g ^ k);
    hi = crnn_mix32((g + 0x9E3779B9u) ^ (k * 0x27D4EB2Fu + 0x165667B1u));
}
__host__ __device__ __forceinline__ uint32_t crnn_dropout_thr16(float rate) { return (uint32_t)((double)rate * 65536.0); }
// masks (0 or inv_keep) of elements 4*group .. 4*group+3
__host__ __device__ __forceinline__ void crnn_dropout_mask4(uint64_t seed, uint32_t layer, uint64_t group, float rate, float inv_keep, float (&m)[4]) {
    uint32_t lo, hi; crnn_dropout_bits(seed, layer, group, lo, hi);
    const uint32_t thr = crnn_dropout_thr16(rate);
    m[0] = (lo & 0xffffu) >= thr ? inv_keep : 0.f; m[1] = (lo >> 16) >= thr ? inv_keep : 0.f;
    m[2] = (hi & 0xffffu) >= thr ? inv_keep : 0.f; m[3] = (hi >> 16) >= thr ? inv_keep : 0.f;
}
// returns 0 or 1/(1-rate) for element idx (same mask as crnn_dropout_mask4 gives that element)
__host__ __device__ __forceinline__ float crnn_dropout_mask(uint64_t seed, uint32_t layer, uint64_t idx, float rate, float inv_keep) {
    uint32_t lo, hi; crnn_dropout_bits(seed, layer, idx >> 2, lo, hi);
    const uint32_t w = (idx & 2) ? hi : lo;
    const uint32_t f = (idx & 1) ? (w >> 16) : (w & 0xffffu);
    return f >= crnn_dropout_thr16(rate) ? inv_keep : 0.f;
}
#endif
