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

// Programmatic dependent launch (PDL), an A/B switch that stays OFF (CRNN_PDL=0 is the default): every kernel of this library starts with
// pdl_enter() -- wait for the full completion (memory flushed) of the kernel before it in the stream, then allow the kernel after it to be
// scheduled -- and crnn_launch() can add the programmatic-stream-serialization attribute, so the CTAs of the next kernel are placed on SMs
// as the last wave of the current one drains (captured into the step's CUDA graph as programmatic edges).  Everything a kernel does to
// global memory happens after its wait, so the data dependencies are exactly those of plain stream order; without the attribute the two
// instructions are no-ops (4.982 ms/step with them, 4.984 before).
// Measured on B200 (bench.py, step graph, 30 steps): CRNN_PDL=0 4.98 ms | 1 (every launch) 5.21 | 2 (not after the recurrence) 5.27 |
// 3 (only while no side branch is open, i.e. most of the forward pass + optimiser) 5.08.  Two effects, both against it here: (a) the
// early-placed CTAs of the next main-stream kernel take the SM slots that free up in the tail of the current one -- exactly the slots in
// which the side branch (weight-gradient GEMMs) makes its progress -- and during the recurrence (128 of 148 SMs) they park on the 20 idle
// SMs; (b) kernels whose grids are cut to fill the CTA slots exactly (dwconv_fused.cu: 296 strips = 148 SMs x 2) get their CTAs placed
// greedily on whichever SMs drain first, up to the occupancy limit, instead of evenly -- the launch gap PDL saves (~1-2 us) is smaller
// than the imbalance it creates.
extern int g_crnn_pdl;               // CRNN_PDL: 0 off (default), 1 every launch, 2 not for the launch that follows a kernel that leaves SMs idle, 3 = 2 + not while a side branch is open
extern int g_crnn_side_open;         // engine: work is pending on the side stream (between side_after and side_join)
extern cudaStream_t g_crnn_pdl_sparse; extern int g_crnn_pdl_sparse_set;
// the kernel just launched on `st` does not fill the GPU (recurrence: 128 of 148 SMs for ~280 us; one CTA per image): a dependent kernel placed
// early would sit on the idle SMs for its whole duration and keep the side branch's kernels off them
static inline void crnn_pdl_mark_sparse(cudaStream_t st) { g_crnn_pdl_sparse = st; g_crnn_pdl_sparse_set = 1; }
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }
template <typename... KA, typename... A>
static inline cudaError_t crnn_launch(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    bool on = g_crnn_pdl != 0;
    if (g_crnn_pdl_sparse_set && g_crnn_pdl_sparse == st) { if (g_crnn_pdl >= 2) on = false; g_crnn_pdl_sparse_set = 0; }
    if (g_crnn_pdl >= 3 && g_crnn_side_open) on = false;
    cfg.attrs = at; cfg.numAttrs = on ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}
#endif

// BatchNorm finalize folded into the kernel that accumulates the batch statistics (training): the CTA that finishes last (ticket counter,
// zeroed together with the statistics) turns [sum | sumsq] into scale / shift / mean / invstd and updates the moving statistics -- what
// bn_finalize_kernel does as a launch of its own (14 per step on the critical path).  The engine offers the job through g_crnn_bn_fin right
// before it calls the launcher of the producing kernel; a launcher that supports the tail takes it (sets the pointer back to nullptr), any
// other path leaves it and the engine launches bn_finalize_kernel as before.
struct BnFin {
    const double* stats; unsigned int* ticket; double M; int C;
    const float* gamma; const float* beta; float* mm; float* mv; float eps, momentum;
    float* scale; float* shift; float* save_mean; float* save_invstd;
};
extern const BnFin* g_crnn_bn_fin;
static inline BnFin crnn_take_bn_fin() { BnFin f = {}; if (g_crnn_bn_fin) { f = *g_crnn_bn_fin; g_crnn_bn_fin = nullptr; } return f; }

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
// one channel of the BatchNorm finalize (training: batch statistics + Keras 2.2.2 moving-average update, SURVEY A.4: fused-op Bessel correction
// times n/(n-(1+eps)); inference: moving statistics); shared by bn_finalize_kernel and the last-CTA tail so both give the same bits
__device__ __forceinline__ void bn_finalize_channel(int c, double sum, double sumsq, double M, int training, const float* gamma, const float* beta,
                                                    float* mm, float* mv, float eps, float momentum, float* scale, float* shift,
                                                    float* save_mean, float* save_invstd) {
    double mean, var;
    if (training) {
        mean = sum / M;
        var = sumsq / M - mean * mean;
        if (var < 0.0) var = 0.0;
        double var_mov = var * (M / (M - 1.0)) * (M / (M - (1.0 + (double)eps)));
        float om = 1.0f - momentum;
        mm[c] = mm[c] - (mm[c] - (float)mean) * om;
        mv[c] = mv[c] - (mv[c] - (float)var_mov) * om;
    } else {
        mean = mm[c]; var = mv[c];
    }
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    if (save_mean) { save_mean[c] = (float)mean; save_invstd[c] = invstd; }
}
// Called by ALL threads of EVERY CTA at the very end of a statistics-producing kernel (after the CTA's atomics on f.stats).
__device__ __forceinline__ void bn_finalize_tail(const BnFin& f) {
    if (!f.ticket) return;                                   // uniform for the grid
    __shared__ int s_last;
    const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x;
    const int nthr = blockDim.x * blockDim.y * blockDim.z;
    __threadfence();                                         // this thread's statistics atomics are performed before the ticket is taken
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(f.ticket, 1u) == gridDim.x * gridDim.y * gridDim.z - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = tid; c < f.C; c += nthr)
        bn_finalize_channel(c, __ldcg(f.stats + c), __ldcg(f.stats + f.C + c), f.M, 1, f.gamma, f.beta, f.mm, f.mv, f.eps, f.momentum,
                            f.scale, f.shift, f.save_mean, f.save_invstd);
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
