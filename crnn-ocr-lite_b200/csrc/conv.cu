// conv.cu -- HBM-bound stages of the depthwise-separable conv stack (reference utils.py:43-56):
// DepthwiseConv2D 3x3 'same' (fwd, bwd-data, bwd-weight), BatchNormalization statistics / finalize / backward,
// ReLU6 + MaxPooling + Dropout (fwd / bwd), and the small element-wise glue of the recurrent head.
// All tensors NHWC fp32; channel counts are 1 or multiples of 4 (float4 path).
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
#ifndef DWCONV_MIN_CTAS
#define DWCONV_MIN_CTAS 1   // (a cap of 3 CTAs = 80 registers spills 56-112 B of the window and ran 15-25 % slower)
#endif
__device__ __forceinline__ void fma4(float4& a, const float4 x, const float4 k) {
    a.x = fmaf(x.x, k.x, a.x); a.y = fmaf(x.y, k.y, a.y); a.z = fmaf(x.z, k.z, a.z); a.w = fmaf(x.w, k.w, a.w);
}

// ------------------------------------------------------------------ depthwise 3x3 forward / backward-data
// FLIP=false: y[h,w] = sum_{i,j} x[h+i-1, w+j-1] * k[i][j]        (forward, cross-correlation like Keras)
// FLIP=true : dx[h,w] = sum_{i,j} dy[h-i+1, w-j+1] * k[i][j]      (backward wrt input)
template <bool FLIP>
__global__ void dwconv3x3_vec4(const float* __restrict__ x, const float* __restrict__ k, float* __restrict__ y,
                               int B, int H, int W, int C4, long long total)
{ pdl_enter();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int c4 = (int)(idx % C4); long long r = idx / C4;
        int w = (int)(r % W); r /= W;
        int h = (int)(r % H); int b = (int)(r / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int hh = FLIP ? h - i + 1 : h + i - 1;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                int ww = FLIP ? w - j + 1 : w + j - 1;
                if (ww < 0 || ww >= W) continue;
                float4 xv = ldg4(x + (((size_t)b * H + hh) * W + ww) * (size_t)(C4 * 4) + c4 * 4);
                float4 kv = ldg4(k + (size_t)(i * 3 + j) * (C4 * 4) + c4 * 4);
                fma4(acc, xv, kv);
            }
        }
        *reinterpret_cast<float4*>(y + (size_t)idx * 4) = acc;
    }
}
// 4 consecutive w positions per thread: 18 input + 9 weight float4 loads for 4 outputs (sliding window in registers)
template <bool FLIP>
__global__ void __launch_bounds__(256) dwconv3x3_vec4_w4(const float* __restrict__ x, const float* __restrict__ k, float* __restrict__ y,
                                                         int B, int H, int W, int C4, int WG, long long total)
{ pdl_enter();
    const int C = C4 * 4;
    const int total32 = (int)total;                 // < 2^31 (checked by the launcher)
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total32; idx += gridDim.x * blockDim.x) {
        const int c4 = idx % C4; int r = idx / C4;
        const int wg = r % WG; r /= WG;
        const int h = r % H; const int b = r / H;
        const int w0 = wg * 4;
        float4 kv[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) kv[q] = ldg4(k + (size_t)q * C + c4 * 4);
        float4 acc[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int hh = FLIP ? h - i + 1 : h + i - 1;
            if (hh < 0 || hh >= H) continue;
            const float* row = x + (((size_t)b * H + hh) * W) * C + c4 * 4;
            float4 xv[6];
#pragma unroll
            for (int t = 0; t < 6; ++t) {
                const int col = w0 - 1 + t;
                xv[t] = (col >= 0 && col < W) ? ldg4(row + (size_t)col * C) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int j = 0; j < 3; ++j) fma4(acc[o], xv[FLIP ? (o - j + 2) : (o + j)], kv[i * 3 + j]);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (w0 + o < W) *reinterpret_cast<float4*>(y + ((((size_t)b * H + h) * W + w0 + o) * C4 + c4) * 4) = acc[o];
    }
}
// Channel-block variant of the kernel above (blockDim = (CQ channel quads, PY lanes)): a thread keeps the nine taps of its 4 channels
// in registers and walks (image row, 4-column group) work items, so the per-item weight loads and two of the three integer divisions
// disappear; with STATS it also accumulates the BatchNorm statistics of its outputs (fp32 per item, fp64 across items, smem across the
// PY lanes, one fp64 atomic pair per channel and CTA) -- the separate colstats pass over the depthwise output is gone.
// (tried: prefetch.global.L1 of the next item's 3x6 window -- 15 % slower, the extra 18 instructions cost more than the hidden latency)
template <bool FLIP, bool STATS>
__global__ void __launch_bounds__(256, DWCONV_MIN_CTAS) dwconv3x3_cb_kernel(const float* __restrict__ x, const float* __restrict__ k, float* __restrict__ y,
                                                           int H, int W, int C4, int WG, int ngroups, double* __restrict__ stats, int rev)
{ pdl_enter();
    extern __shared__ double dsm[];   // STATS: [PY][8][CQ]
    const int CQ = blockDim.x, PY = blockDim.y;
    const int c4 = blockIdx.x * CQ + threadIdx.x;
    const int C = C4 * 4;
    double S[4] = {0.0, 0.0, 0.0, 0.0}, Q[4] = {0.0, 0.0, 0.0, 0.0};
    if (c4 < C4) {
        float4 kv[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) kv[q] = ldg4(k + (size_t)q * C + c4 * 4);
        const int gstride = gridDim.y * PY;
        for (int g0 = blockIdx.y * PY + threadIdx.y; g0 < ngroups; g0 += gstride) {
            const int g = rev ? ngroups - 1 - g0 : g0;
            const int rowi = g / WG, wg = g - rowi * WG;         // rowi = b*H + h
            const int h = rowi % H;
            const int w0 = wg * 4;
            float4 acc[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int dh = FLIP ? 1 - i : i - 1;
                if (h + dh < 0 || h + dh >= H) continue;
                const float* row = x + ((size_t)(rowi + dh) * W) * C + c4 * 4;
                float4 xv[6];
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    const int col = w0 - 1 + t;
                    xv[t] = (col >= 0 && col < W) ? ldg4(row + (size_t)col * C) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int o = 0; o < 4; ++o)
#pragma unroll
                    for (int j = 0; j < 3; ++j) fma4(acc[o], xv[FLIP ? (o - j + 2) : (o + j)], kv[i * 3 + j]);
            }
            float* dst = y + ((size_t)rowi * W + w0) * C + c4 * 4;
            float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (w0 + o < W) {
                    *reinterpret_cast<float4*>(dst + (size_t)o * C) = acc[o];
                    if (STATS) {
                        s[0] += acc[o].x; s[1] += acc[o].y; s[2] += acc[o].z; s[3] += acc[o].w;
                        q[0] = fmaf(acc[o].x, acc[o].x, q[0]); q[1] = fmaf(acc[o].y, acc[o].y, q[1]);
                        q[2] = fmaf(acc[o].z, acc[o].z, q[2]); q[3] = fmaf(acc[o].w, acc[o].w, q[3]);
                    }
                }
            if (STATS) {
#pragma unroll
                for (int e = 0; e < 4; ++e) { S[e] += (double)s[e]; Q[e] += (double)q[e]; }
            }
        }
    }
    if (!STATS) return;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        dsm[(threadIdx.y * 8 + e) * CQ + threadIdx.x] = S[e];
        dsm[(threadIdx.y * 8 + 4 + e) * CQ + threadIdx.x] = Q[e];
    }
    __syncthreads();
    if (threadIdx.y == 0 && c4 < C4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            double t1 = 0.0, t2 = 0.0;
            for (int i = 0; i < PY; ++i) { t1 += dsm[(i * 8 + e) * CQ + threadIdx.x]; t2 += dsm[(i * 8 + 4 + e) * CQ + threadIdx.x]; }
            atomicAdd(stats + c4 * 4 + e, t1); atomicAdd(stats + C + c4 * 4 + e, t2);
        }
    }
}
template <bool FLIP>
__global__ void dwconv3x3_c1(const float* __restrict__ x, const float* __restrict__ k, float* __restrict__ y,
                             int B, int H, int W, long long total)
{ pdl_enter();
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int w = (int)(idx % W); long long r = idx / W;
        int h = (int)(r % H); int b = (int)(r / H);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int hh = FLIP ? h - i + 1 : h + i - 1;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                int ww = FLIP ? w - j + 1 : w + j - 1;
                if (ww < 0 || ww >= W) continue;
                acc = fmaf(__ldg(x + ((size_t)b * H + hh) * W + ww), __ldg(k + i * 3 + j), acc);
            }
        }
        y[idx] = acc;
    }
}

// ------------------------------------------------------------------ depthwise 3x3 backward-weight
// dk[i][j][c] = sum_{b,h,w} x[b,h+i-1,w+j-1,c] * dy[b,h,w,c].  blockDim = (CQ channel-quads | CT channels, PY pixel lanes)
__global__ void __launch_bounds__(256, DWCONV_MIN_CTAS) dwconv3x3_bwd_weight_vec4(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dk,
                                          int B, int H, int W, int C4, int WG, long long ngroups)
{ pdl_enter();
    extern __shared__ float red[];   // [PY][36][CQ]
    const int CQ = blockDim.x, PY = blockDim.y, C = C4 * 4;
    const int c4 = blockIdx.x * CQ + threadIdx.x;
    float4 acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < C4) {
        // one iteration = 4 consecutive pixels of a row: 18 x loads + 4 dy loads feed 36 float4 FMAs
        const int ng32 = (int)ngroups, pstride = gridDim.y * PY;
        for (int p = blockIdx.y * PY + threadIdx.y; p < ng32; p += pstride) {
            const int wg = p % WG; const int r = p / WG;
            const int h = r % H; const int b = r / H;
            const int w0 = wg * 4;
            float4 g[4];
#pragma unroll
            for (int o = 0; o < 4; ++o)
                g[o] = (w0 + o < W) ? ldg4(dy + ((((size_t)b * H + h) * W + w0 + o) * C4 + c4) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int hh = h + i - 1;
                if (hh < 0 || hh >= H) continue;
                const float* row = x + (((size_t)b * H + hh) * W) * C + c4 * 4;
                float4 xv[6];
#pragma unroll
                for (int t = 0; t < 6; ++t) {
                    const int col = w0 - 1 + t;
                    xv[t] = (col >= 0 && col < W) ? ldg4(row + (size_t)col * C) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int o = 0; o < 4; ++o) fma4(acc[i * 3 + j], xv[o + j], g[o]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        red[((threadIdx.y * 36) + q * 4 + 0) * CQ + threadIdx.x] = acc[q].x; red[((threadIdx.y * 36) + q * 4 + 1) * CQ + threadIdx.x] = acc[q].y;
        red[((threadIdx.y * 36) + q * 4 + 2) * CQ + threadIdx.x] = acc[q].z; red[((threadIdx.y * 36) + q * 4 + 3) * CQ + threadIdx.x] = acc[q].w;
    }
    __syncthreads();
    if (c4 < C4)
        for (int e = threadIdx.y; e < 36; e += PY) {
            float sum = 0.f;
            for (int yy = 0; yy < PY; ++yy) sum += red[(yy * 36 + e) * CQ + threadIdx.x];
            atomicAdd(dk + (size_t)(e >> 2) * C + c4 * 4 + (e & 3), sum);
        }
}
__global__ void dwconv3x3_bwd_weight(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dk,
                                     int B, int H, int W, int C, long long npix)
{ pdl_enter();
    extern __shared__ float red[];   // [PY][9][CT]
    const int CT = blockDim.x, PY = blockDim.y;
    const int c = blockIdx.x * CT + threadIdx.x;
    float acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = 0.f;
    if (c < C) {
        for (long long p = blockIdx.y * (long long)PY + threadIdx.y; p < npix; p += (long long)gridDim.y * PY) {
            int w = (int)(p % W); long long r = p / W;
            int h = (int)(r % H); int b = (int)(r / H);
            float g = __ldg(dy + (size_t)p * C + c);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                int hh = h + i - 1;
                if (hh < 0 || hh >= H) continue;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int ww = w + j - 1;
                    if (ww < 0 || ww >= W) continue;
                    acc[i * 3 + j] = fmaf(__ldg(x + (((size_t)b * H + hh) * W + ww) * C + c), g, acc[i * 3 + j]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) red[(threadIdx.y * 9 + q) * CT + threadIdx.x] = acc[q];
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            float s = 0.f;
            for (int y = 0; y < PY; ++y) s += red[(y * 9 + q) * CT + threadIdx.x];
            atomicAdd(dk + (size_t)q * C + c, s);
        }
    }
}

// ------------------------------------------------------------------ per-channel statistics
// blockDim = (CT, PY); stats[c] += sum y, stats[C+c] += sum y^2 (double)
__global__ void colstats_kernel(const float* __restrict__ y, long long M, int C, double* __restrict__ stats, const BnFin fin)
{ pdl_enter();
    extern __shared__ double dred[];  // [PY][2][CT]
    const int CT = blockDim.x, PY = blockDim.y;
    const int c = blockIdx.x * CT + threadIdx.x;
    double s = 0.0, q = 0.0;
    if (c < C)
        for (long long m = blockIdx.y * (long long)PY + threadIdx.y; m < M; m += (long long)gridDim.y * PY) {
            double v = (double)__ldg(y + (size_t)m * C + c);
            s += v; q = fma(v, v, q);
        }
    dred[(threadIdx.y * 2 + 0) * CT + threadIdx.x] = s;
    dred[(threadIdx.y * 2 + 1) * CT + threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double ts = 0.0, tq = 0.0;
        for (int i = 0; i < PY; ++i) { ts += dred[(i * 2) * CT + threadIdx.x]; tq += dred[(i * 2 + 1) * CT + threadIdx.x]; }
        atomicAdd(stats + c, ts); atomicAdd(stats + C + c, tq);
    }
    bn_finalize_tail(fin);
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ mm, float* __restrict__ mv,
                                   float eps, float momentum, int training, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd)
{ pdl_enter();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    bn_finalize_channel(c, training ? stats[c] : 0.0, training ? stats[C + c] : 0.0, M, training, gamma, beta, mm, mv, eps, momentum, scale, shift, save_mean, save_invstd);
}

// ------------------------------------------------------------------ BN + ReLU6 + MaxPool + Dropout forward
// blockDim = (CQ channel quads, PY pixel lanes): a thread keeps its 4 channels' scale/shift in registers and walks output pixels
// (one integer division per pixel; two pixels' window loads in flight).  ncu r1e on the 1-D version: 382 warp instructions per float4
// (per-element 64-bit dropout hash + 4 div/mod) and 48 % issue-active at 3.3 TB/s, i.e. instruction-bound, not HBM-bound.
template <int PH, int PW>
__global__ void __launch_bounds__(256) act_pool_fwd_kernel(const float* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                                    float* __restrict__ a, int W, int C4, int Wo, int npix,
                                    float rate, float inv_keep, uint64_t seed, uint32_t layer, const uint64_t* __restrict__ seed_ptr, int rev)
{ pdl_enter();
    const int CQ = blockDim.x, PY = blockDim.y;
    const int c4 = blockIdx.x * CQ + threadIdx.x;
    if (c4 >= C4) return;
    if (seed_ptr) seed = *seed_ptr;                 // graph replay: the step's seed lives in device memory
    const int C = C4 * 4;
    const float4 sc = ldg4(scale + c4 * 4), sh = ldg4(shift + c4 * 4);
    const int pstride = gridDim.y * PY;
#pragma unroll 2
    for (int p0 = blockIdx.y * PY + threadIdx.y; p0 < npix; p0 += pstride) {
        const int p = rev ? npix - 1 - p0 : p0;                          // serpentine traversal (see engine.cu): start where the producer ended
        const int row = p / Wo, wo = p - row * Wo;                       // row = b*Ho + ho; input row = row*PH because H = Ho*PH
        const float* src = y + ((size_t)(row * PH) * W + wo * PW) * C + c4 * 4;
        float4 v[PH * PW];
#pragma unroll
        for (int i = 0; i < PH; ++i)
#pragma unroll
            for (int j = 0; j < PW; ++j) v[i * PW + j] = ldg4(src + ((size_t)i * W + j) * C);
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int n = 0; n < PH * PW; ++n) {
            m.x = fmaxf(m.x, relu6f(fmaf(v[n].x, sc.x, sh.x))); m.y = fmaxf(m.y, relu6f(fmaf(v[n].y, sc.y, sh.y)));
            m.z = fmaxf(m.z, relu6f(fmaf(v[n].z, sc.z, sh.z))); m.w = fmaxf(m.w, relu6f(fmaf(v[n].w, sc.w, sh.w)));
        }
        const size_t oidx = (size_t)p * C4 + c4;
        if (rate > 0.f) {
            float dm[4]; crnn_dropout_mask4(seed, layer, (uint64_t)oidx, rate, inv_keep, dm);
            m.x *= dm[0]; m.y *= dm[1]; m.z *= dm[2]; m.w *= dm[3];
        }
        *reinterpret_cast<float4*>(a + oidx * 4) = m;
    }
}

// backward of the above, two passes over (da, y) with blockDim = (CQ channel-quads, PY pixel lanes):
//   APPLY=false: only the reductions sum(dz), sum(dz*xhat) per channel (registers -> smem -> one double atomic per channel per CTA)
//   APPLY=true : recomputes dz and writes the BatchNorm-backward result dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat))
// so the intermediate dz never touches HBM.  dz = unpool(da * dropmask) * 1[0 <= z <= 6], z = y*scale+shift.
template <bool APPLY, int PH, int PW>
__global__ void __launch_bounds__(256) act_pool_bwd_kernel(const float* __restrict__ da, const float* __restrict__ y, const float* __restrict__ scale,
                                    const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, float* __restrict__ dy, double* __restrict__ red,
                                    int B, int H, int W, int C4, double invM,
                                    float rate, float inv_keep, uint64_t seed, uint32_t layer, long long npix_ll, const uint64_t* __restrict__ seed_ptr, int rev)
{ pdl_enter();
    extern __shared__ float sred[];   // [PY][8][CQ]
    if (seed_ptr) seed = *seed_ptr;
    constexpr int ph = PH, pw = PW, NW = PH * PW;
    const int CQ = blockDim.x, PY = blockDim.y;
    const int Wo = W / pw, C = C4 * 4;
    const int npix = (int)npix_ll;                  // < 2^31 (checked by the launcher): 32-bit index arithmetic
    const int c4 = blockIdx.x * CQ + threadIdx.x;
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    if (c4 < C4) {
        float sc[4], sh[4], mu[4], is[4], gs[4] = {0, 0, 0, 0}, m1[4] = {0, 0, 0, 0}, m2[4] = {0, 0, 0, 0};
        { float4 t = ldg4(scale + c4 * 4); sc[0] = t.x; sc[1] = t.y; sc[2] = t.z; sc[3] = t.w; }
        { float4 t = ldg4(shift + c4 * 4); sh[0] = t.x; sh[1] = t.y; sh[2] = t.z; sh[3] = t.w; }
        { float4 t = ldg4(mean + c4 * 4); mu[0] = t.x; mu[1] = t.y; mu[2] = t.z; mu[3] = t.w; }
        { float4 t = ldg4(invstd + c4 * 4); is[0] = t.x; is[1] = t.y; is[2] = t.z; is[3] = t.w; }
        if (APPLY) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                gs[q] = gamma[c4 * 4 + q] * is[q];
                m1[q] = (float)(red[c4 * 4 + q] * invM); m2[q] = (float)(red[C + c4 * 4 + q] * invM);
            }
        }
        const int pstride = gridDim.y * PY;
#pragma unroll 2
        for (int p0 = blockIdx.y * PY + threadIdx.y; p0 < npix; p0 += pstride) {
            const int p = rev ? npix - 1 - p0 : p0;
            const int row = p / Wo, wo = p - row * Wo;       // row = b*Ho + ho; input row = row*ph because H = Ho*ph
            const size_t oidx = (size_t)p * C4 + c4;
            float g[4];
            { float4 t = ldg4(da + oidx * 4); g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w; }
            if (rate > 0.f) {
                float dm[4]; crnn_dropout_mask4(seed, layer, (uint64_t)oidx, rate, inv_keep, dm);
#pragma unroll
                for (int q = 0; q < 4; ++q) g[q] *= dm[q];
            }
            const size_t base = ((size_t)(row * ph) * W + wo * pw) * C + c4 * 4;
            float yv[NW][4];
#pragma unroll
            for (int i = 0; i < ph; ++i)
#pragma unroll
                for (int j = 0; j < pw; ++j) {
                    const float4 t = ldg4(y + base + ((size_t)i * W + j) * C);
                    yv[i * pw + j][0] = t.x; yv[i * pw + j][1] = t.y; yv[i * pw + j][2] = t.z; yv[i * pw + j][3] = t.w;
                }
            // arg-max of relu6(z) over the window (first max wins, like TF / torch max-pool grad); only that element gets gradient
            float dq[4], xq[4]; int arg[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float best = -INFINITY, zb = 0.f, yb = 0.f; int ab = 0;
#pragma unroll
                for (int n = 0; n < NW; ++n) {
                    const float z = fmaf(yv[n][q], sc[q], sh[q]);
                    const float aq = relu6f(z);
                    if (aq > best) { best = aq; zb = z; yb = yv[n][q]; ab = n; }
                }
                arg[q] = ab;
                dq[q] = (zb >= 0.f && zb <= 6.f) ? g[q] : 0.f;
                xq[q] = (yb - mu[q]) * is[q];
            }
            if (!APPLY) {
#pragma unroll
                for (int q = 0; q < 4; ++q) { s1[q] += dq[q]; s2[q] = fmaf(dq[q], xq[q], s2[q]); }
            } else {
#pragma unroll
                for (int i = 0; i < ph; ++i)
#pragma unroll
                    for (int j = 0; j < pw; ++j) {
                        const int n = i * pw + j;
                        float o[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float xh = (yv[n][q] - mu[q]) * is[q];
                            const float d = (arg[q] == n) ? dq[q] : 0.f;
                            o[q] = gs[q] * (d - m1[q] - xh * m2[q]);
                        }
                        *reinterpret_cast<float4*>(dy + base + ((size_t)i * W + j) * C) = make_float4(o[0], o[1], o[2], o[3]);
                    }
            }
        }
    }
    if (APPLY) return;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sred[(threadIdx.y * 8 + q) * CQ + threadIdx.x] = s1[q];
        sred[(threadIdx.y * 8 + 4 + q) * CQ + threadIdx.x] = s2[q];
    }
    __syncthreads();
    if (threadIdx.y == 0 && c4 < C4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double t1 = 0.0, t2 = 0.0;
            for (int i = 0; i < PY; ++i) { t1 += sred[(i * 8 + q) * CQ + threadIdx.x]; t2 += sred[(i * 8 + 4 + q) * CQ + threadIdx.x]; }
            atomicAdd(red + c4 * 4 + q, t1); atomicAdd(red + C + c4 * 4 + q, t2);
        }
    }
}

// dgamma += sum(dz*xhat), dbeta += sum(dz) from the reduction buffer
__global__ void bn_param_grads_kernel(const double* __restrict__ red, float* __restrict__ dgamma, float* __restrict__ dbeta, int C)
{ pdl_enter();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { dbeta[c] += (float)red[c]; dgamma[c] += (float)red[C + c]; }
}
__global__ void bn_param_grads_all_kernel(BnGradTable t)       // blockIdx.y = layer
{ pdl_enter();
    const int l = blockIdx.y, C = t.C[l];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) { t.dbeta[l][c] += (float)t.red[l][c]; t.dgamma[l][c] += (float)t.red[l][C + c]; }
}

// ReLU6 + BatchNorm backward of the BN that follows the depthwise conv (no pool / dropout): y, da, dy are [M][C].
//   APPLY=false: reductions only;  APPLY=true: dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)), dy may alias da.
template <bool APPLY>
__global__ void relu6_bwd_scalar_kernel(const float* da /* may alias dy */, const float* __restrict__ y, const float* __restrict__ scale,
                                 const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                                 const float* __restrict__ gamma, float* dy, double* __restrict__ red, long long M, int C, double invM)
{ pdl_enter();
    extern __shared__ float sred[];   // [PY][2][CT]
    const int CT = blockDim.x, PY = blockDim.y;
    const int c = blockIdx.x * CT + threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    if (c < C) {
        const float sc = scale[c], sh = shift[c], mu = mean[c], is = invstd[c];
        float gs = 0.f, m1 = 0.f, m2 = 0.f;
        if (APPLY) { gs = gamma[c] * is; m1 = (float)(red[c] * invM); m2 = (float)(red[C + c] * invM); }
        for (long long m = blockIdx.y * (long long)PY + threadIdx.y; m < M; m += (long long)gridDim.y * PY) {
            const float yv = __ldg(y + (size_t)m * C + c);
            const float z = fmaf(yv, sc, sh);
            const float d = (z >= 0.f && z <= 6.f) ? da[(size_t)m * C + c] : 0.f;
            const float xh = (yv - mu) * is;
            if (APPLY) dy[(size_t)m * C + c] = gs * (d - m1 - xh * m2);
            else { s1 += d; s2 = fmaf(d, xh, s2); }
        }
    }
    if (APPLY) return;
    sred[(threadIdx.y * 2 + 0) * CT + threadIdx.x] = s1;
    sred[(threadIdx.y * 2 + 1) * CT + threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double t1 = 0.0, t2 = 0.0;
        for (int i = 0; i < PY; ++i) { t1 += sred[(i * 2) * CT + threadIdx.x]; t2 += sred[(i * 2 + 1) * CT + threadIdx.x]; }
        atomicAdd(red + c, t1); atomicAdd(red + C + c, t2);
    }
}

// float4 version (C % 4 == 0): blockDim = (CQ channel quads, PY row lanes), two rows per iteration for memory-level parallelism
template <bool APPLY>
__global__ void __launch_bounds__(256) relu6_bwd_kernel(const float* da /* may alias dy */, const float* __restrict__ y, const float* __restrict__ scale,
                                 const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                                 const float* __restrict__ gamma, float* dy, double* __restrict__ red, long long M, int C4, double invM, int rev)
{ pdl_enter();
    extern __shared__ float sred[];   // [PY][8][CQ]
    const int CQ = blockDim.x, PY = blockDim.y, C = C4 * 4;
    const int c4 = blockIdx.x * CQ + threadIdx.x;
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    if (c4 < C4) {
        float sc[4], sh[4], mu[4], is[4], gs[4] = {0, 0, 0, 0}, m1[4] = {0, 0, 0, 0}, m2[4] = {0, 0, 0, 0};
        { float4 t = ldg4(scale + c4 * 4); sc[0] = t.x; sc[1] = t.y; sc[2] = t.z; sc[3] = t.w; }
        { float4 t = ldg4(shift + c4 * 4); sh[0] = t.x; sh[1] = t.y; sh[2] = t.z; sh[3] = t.w; }
        { float4 t = ldg4(mean + c4 * 4); mu[0] = t.x; mu[1] = t.y; mu[2] = t.z; mu[3] = t.w; }
        { float4 t = ldg4(invstd + c4 * 4); is[0] = t.x; is[1] = t.y; is[2] = t.z; is[3] = t.w; }
        if (APPLY) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { gs[q] = gamma[c4 * 4 + q] * is[q]; m1[q] = (float)(red[c4 * 4 + q] * invM); m2[q] = (float)(red[C + c4 * 4 + q] * invM); }
        }
        const long long stride = (long long)gridDim.y * PY;
        for (long long m = blockIdx.y * (long long)PY + threadIdx.y; m < M; m += 2 * stride) {
            const long long m2r = m + stride;
            const bool two = m2r < M;
            const long long ma = rev ? M - 1 - m : m, mb = two ? (rev ? M - 1 - m2r : m2r) : ma;
            const size_t o0 = ((size_t)ma * C4 + c4) * 4, o1 = ((size_t)mb * C4 + c4) * 4;
            const float4 y0 = ldg4(y + o0), y1 = ldg4(y + o1);
            const float4 d0 = *reinterpret_cast<const float4*>(da + o0), d1 = *reinterpret_cast<const float4*>(da + o1);
            const float yv[2][4] = {{y0.x, y0.y, y0.z, y0.w}, {y1.x, y1.y, y1.z, y1.w}};
            const float dv[2][4] = {{d0.x, d0.y, d0.z, d0.w}, {d1.x, d1.y, d1.z, d1.w}};
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (rr == 1 && !two) break;
                float o[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float z = fmaf(yv[rr][q], sc[q], sh[q]);
                    const float d = (z >= 0.f && z <= 6.f) ? dv[rr][q] : 0.f;
                    const float xh = (yv[rr][q] - mu[q]) * is[q];
                    if (APPLY) o[q] = gs[q] * (d - m1[q] - xh * m2[q]);
                    else { s1[q] += d; s2[q] = fmaf(d, xh, s2[q]); }
                }
                if (APPLY) *reinterpret_cast<float4*>(dy + (rr ? o1 : o0)) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
    if (APPLY) return;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sred[(threadIdx.y * 8 + q) * CQ + threadIdx.x] = s1[q];
        sred[(threadIdx.y * 8 + 4 + q) * CQ + threadIdx.x] = s2[q];
    }
    __syncthreads();
    if (threadIdx.y == 0 && c4 < C4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double t1 = 0.0, t2 = 0.0;
            for (int i = 0; i < PY; ++i) { t1 += sred[(i * 8 + q) * CQ + threadIdx.x]; t2 += sred[(i * 8 + 4 + q) * CQ + threadIdx.x]; }
            atomicAdd(red + c4 * 4 + q, t1); atomicAdd(red + C + c4 * 4 + q, t2);
        }
    }
}

// Reduction pass of the BatchNorm(+ReLU6[+Dropout]) backward on a non-pooled [M][C] tensor: red[c] += sum_m dz, red[C+c] += sum_m dz*xhat with
// dz = da * dropmask * 1[0 <= z <= 6].  Dedicated kernel (instead of the APPLY=false instances of the two kernels above): 4 rows = 8 float4
// loads in flight per thread and ~64 registers, because this pass is pure HBM latency (ncu r1h: 2.8-4.0 TB/s, long-scoreboard 15-18 per issue).
template <bool DROP>
__global__ void __launch_bounds__(256, 3) bn_relu6_reduce_kernel(const float* __restrict__ da, const float* __restrict__ y, const float* __restrict__ scale,
                                 const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                                 double* __restrict__ red, int M, int C4, int rev,
                                 float rate, float inv_keep, uint64_t seed, uint32_t layer, const uint64_t* __restrict__ seed_ptr)
{ pdl_enter();
    extern __shared__ float sred[];   // [PY][8][CQ]
    const int CQ = blockDim.x, PY = blockDim.y, C = C4 * 4;
    const int c4 = blockIdx.x * CQ + threadIdx.x;
    if (DROP && seed_ptr) seed = *seed_ptr;
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    if (c4 < C4) {
        const float4 sc = ldg4(scale + c4 * 4), sh = ldg4(shift + c4 * 4);
        float4 xa = ldg4(invstd + c4 * 4), xb = ldg4(mean + c4 * 4);            // xhat = y*xa + xb
        xb.x = -xb.x * xa.x; xb.y = -xb.y * xa.y; xb.z = -xb.z * xa.z; xb.w = -xb.w * xa.w;
        const int stride = gridDim.y * PY;
        for (int m0 = blockIdx.y * PY + threadIdx.y; m0 < M; m0 += 4 * stride) {
            float4 yv[4], dv[4]; size_t o[4]; bool ok[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int mm = m0 + r * stride;
                ok[r] = mm < M;
                const int m = ok[r] ? (rev ? M - 1 - mm : mm) : (rev ? M - 1 - m0 : m0);
                o[r] = (size_t)m * C4 + c4;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) { yv[r] = ldg4(y + o[r] * 4); dv[r] = ldg4(da + o[r] * 4); }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (!ok[r]) continue;
                float d[4] = {dv[r].x, dv[r].y, dv[r].z, dv[r].w};
                const float yy[4] = {yv[r].x, yv[r].y, yv[r].z, yv[r].w};
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
                const float xav[4] = {xa.x, xa.y, xa.z, xa.w}, xbv[4] = {xb.x, xb.y, xb.z, xb.w};
                if (DROP) {
                    float dm[4]; crnn_dropout_mask4(seed, layer, (uint64_t)o[r], rate, inv_keep, dm);
#pragma unroll
                    for (int q = 0; q < 4; ++q) d[q] *= dm[q];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float z = fmaf(yy[q], scv[q], shv[q]);
                    const float dz = (z >= 0.f && z <= 6.f) ? d[q] : 0.f;
                    s1[q] += dz; s2[q] = fmaf(dz, fmaf(yy[q], xav[q], xbv[q]), s2[q]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        sred[(threadIdx.y * 8 + q) * CQ + threadIdx.x] = s1[q];
        sred[(threadIdx.y * 8 + 4 + q) * CQ + threadIdx.x] = s2[q];
    }
    __syncthreads();
    if (threadIdx.y == 0 && c4 < C4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double t1 = 0.0, t2 = 0.0;
            for (int i = 0; i < PY; ++i) { t1 += sred[(i * 8 + q) * CQ + threadIdx.x]; t2 += sred[(i * 8 + 4 + q) * CQ + threadIdx.x]; }
            atomicAdd(red + c4 * 4 + q, t1); atomicAdd(red + C + c4 * 4 + q, t2);
        }
    }
}

// dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat))  (in place);  block 0 also emits dgamma/dbeta
__global__ void bn_bwd_apply_kernel(float* __restrict__ dz, const float* __restrict__ y, const double* __restrict__ red,
                                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, long long M, int C, long long total)
{ pdl_enter();
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += blockDim.x) { dbeta[c] += (float)red[c]; dgamma[c] += (float)red[C + c]; }
    const double invM = 1.0 / (double)M;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C);
        float m1 = (float)(red[c] * invM), m2 = (float)(red[C + c] * invM);
        float is = invstd[c];
        float xh = (y[idx] - mean[c]) * is;
        dz[idx] = gamma[c] * is * (dz[idx] - m1 - xh * m2);
    }
}

// ------------------------------------------------------------------ misc element-wise
__global__ void colsum_kernel(const float* __restrict__ y, long long M, int C, int ldy, float* __restrict__ out)
{ pdl_enter();
    extern __shared__ float sred[];   // [PY][CT]
    const int CT = blockDim.x, PY = blockDim.y;
    const int c = blockIdx.x * CT + threadIdx.x;
    float s = 0.f;
    if (c < C)
        for (long long m = blockIdx.y * (long long)PY + threadIdx.y; m < M; m += (long long)gridDim.y * PY) s += __ldg(y + (size_t)m * ldy + c);
    sred[threadIdx.y * CT + threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
        for (int i = 0; i < PY; ++i) t += sred[i * CT + threadIdx.x];
        atomicAdd(out + c, t);
    }
}
__global__ void relu_dropout_bwd_kernel(float* __restrict__ g, const float* __restrict__ act, long long n, float inv_keep)
{ pdl_enter();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        g[i] = act[i] > 0.f ? g[i] * inv_keep : 0.f;
}
__global__ void dropout_kernel(const float* in, float* x, long long n, float rate, float inv_keep, uint64_t seed, uint32_t layer,
                               const uint64_t* __restrict__ seed_ptr)
{ pdl_enter();
    if (seed_ptr) seed = *seed_ptr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = in[i] * crnn_dropout_mask(seed, layer, (uint64_t)i, rate, inv_keep);
}
// Fixed-order sum of the S split-K copies of a GEMM output, fused with bias, ReLU and (training) the inverted-dropout mask of the layer.
__global__ void sum_partials_kernel(const float* __restrict__ part, int S, long long stride, int N4, long long total4, const float* __restrict__ bias, int relu,
                                    float* __restrict__ out, int ldo, float rate, float inv_keep, uint64_t seed, uint32_t layer, const uint64_t* __restrict__ seed_ptr)
{ pdl_enter();
    if (seed_ptr) seed = *seed_ptr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / N4; const int n4 = (int)(i - m * N4);
        float4 v = ldg4(part + i * 4);
        for (int s2 = 1; s2 < S; ++s2) { const float4 w = ldg4(part + (size_t)s2 * stride + i * 4); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        if (bias) { const float4 b = ldg4(bias + n4 * 4); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (rate > 0.f) { float mk[4]; crnn_dropout_mask4(seed, layer, (uint64_t)i, rate, inv_keep, mk); v.x *= mk[0]; v.y *= mk[1]; v.z *= mk[2]; v.w *= mk[3]; }
        *reinterpret_cast<float4*>(out + (size_t)m * ldo + n4 * 4) = v;
    }
}
// ------------------------------------------------------------------ block 1 (Cin = 1): pointwise conv = outer product
// blockDim = 256 = 16 pixel lanes x 16 channel quads (Cout = 64); a warp covers 2 pixels x 64 channels = 2 x 256 B contiguous.
// BN statistics of the output in closed form: sum_m f_m w_c = w_c * S1, sum_m (f_m w_c)^2 = w_c^2 * S2 (S1, S2 accumulated in double per CTA).
__global__ void __launch_bounds__(256) pw1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                                      const float* __restrict__ w, float* __restrict__ out, int M, int C4, double* __restrict__ stats, int rev, const BnFin fin)
{ pdl_enter();
    __shared__ float s1s[8], s2s[8];
    const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4, PL = blockDim.x / C4;
    const float sc = __ldg(scale), sh = __ldg(shift);
    const float4 wq = ldg4(w + c4 * 4);
    float s1 = 0.f, s2 = 0.f;
    for (int m0 = blockIdx.x * PL + pl; m0 < M; m0 += gridDim.x * PL) {
        const int m = rev ? M - 1 - m0 : m0;
        const float f = relu6f(fmaf(__ldg(x + m), sc, sh));
        *reinterpret_cast<float4*>(out + ((size_t)m * C4 + c4) * 4) = make_float4(f * wq.x, f * wq.y, f * wq.z, f * wq.w);
        if (c4 == 0) { s1 += f; s2 = fmaf(f, f, s2); }
    }
    if (!stats) return;
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { s1s[threadIdx.x >> 5] = s1; s2s[threadIdx.x >> 5] = s2; }
    __syncthreads();
    const int C = C4 * 4;
    if (threadIdx.x < C) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += (double)s1s[i]; b += (double)s2s[i]; }
        const double wc = (double)__ldg(w + threadIdx.x);
        atomicAdd(stats + threadIdx.x, wc * a);
        atomicAdd(stats + C + threadIdx.x, wc * wc * b);
    }
    bn_finalize_tail(fin);
}
// backward in one pass over dY: dX[m] = sum_c dY[m][c] w[c] (16-lane shuffle reduce), dW[c] += sum_m f(x[m]) dY[m][c]
__global__ void __launch_bounds__(256) pw1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                                      const float* __restrict__ dY, const float* __restrict__ w, float* __restrict__ dX,
                                                      float* __restrict__ dW, int M, int rev)
{ pdl_enter();
    __shared__ float4 sacc[256];
    const int c4 = threadIdx.x & 15, pl = threadIdx.x >> 4;
    const float sc = __ldg(scale), sh = __ldg(shift);
    const float4 wq = ldg4(w + c4 * 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int mstep = gridDim.x * 16;
    // every lane of a warp runs the same number of iterations (M is padded to the stride by the bound check inside)
    for (int m0 = blockIdx.x * 16; m0 < M; m0 += mstep) {
        const int mf = m0 + pl, m = rev ? M - 1 - mf : mf;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f); float f = 0.f;
        if (mf < M) { g = ldg4(dY + ((size_t)m * 16 + c4) * 4); f = relu6f(fmaf(__ldg(x + m), sc, sh)); }
        float d = g.x * wq.x + g.y * wq.y + g.z * wq.z + g.w * wq.w;
#pragma unroll
        for (int o = 8; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (c4 == 0 && mf < M) dX[m] = d;
        acc.x = fmaf(f, g.x, acc.x); acc.y = fmaf(f, g.y, acc.y); acc.z = fmaf(f, g.z, acc.z); acc.w = fmaf(f, g.w, acc.w);
    }
    sacc[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < 16) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < 16; ++i) { const float4 v = sacc[i * 16 + threadIdx.x]; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
        atomicAdd(dW + threadIdx.x * 4 + 0, t.x); atomicAdd(dW + threadIdx.x * 4 + 1, t.y);
        atomicAdd(dW + threadIdx.x * 4 + 2, t.z); atomicAdd(dW + threadIdx.x * 4 + 3, t.w);
    }
}
__global__ void set_u64_kernel(uint64_t* p, uint64_t v) { pdl_enter(); *p = v; }
__global__ void sum_dirs_kernel(const float* __restrict__ hs, float* __restrict__ out, long long rows, int U)
{ pdl_enter();
    long long n = rows * U;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / U; int u = (int)(i % U);
        out[i] = hs[(r * 2) * U + u] + hs[(r * 2 + 1) * U + u];
    }
}
__global__ void dup_dirs_kernel(const float* __restrict__ g, float* __restrict__ out, long long rows, int U)
{ pdl_enter();
    long long n = rows * U;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / U; int u = (int)(i % U);
        float v = g[i];
        out[(r * 2) * U + u] = v; out[(r * 2 + 1) * U + u] = v;
    }
}
// softmax over the last axis, one warp per row (same formula as Keras softmax: exp(z-max)/sum)
__global__ void softmax_rows_kernel(const float* __restrict__ z, float* __restrict__ p, long long rows, int V)
{ pdl_enter();
    const int lane = threadIdx.x & 31;
    long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* zr = z + (size_t)row * V; float* pr = p + (size_t)row * V;
    float mx = -INFINITY;
    for (int k = lane; k < V; k += 32) mx = fmaxf(mx, zr[k]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int k = lane; k < V; k += 32) s += expf(zr[k] - mx);
    s = warp_sum(s);
    for (int k = lane; k < V; k += 32) pr[k] = expf(zr[k] - mx) / s;
}

inline bool too_big(long long n) { if (n >= (1ll << 31)) { crnn_set_error("tensor too large for 32-bit indexing (%lld elements)", n); return true; } return false; }
inline int grid1d(long long total, int threads, int max_blocks = 148 * 16) {
    long long b = (total + threads - 1) / threads;
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return (int)b;
}
// 2-D (channel, pixel-lane) block: CT channels x PY lanes = 256 threads, grid.y sized for ~4 CTAs/SM
inline void chan_block(int C, long long M, dim3& grid, dim3& block) {
    int CT = C >= 64 ? 64 : (C >= 32 ? 32 : (C >= 16 ? 16 : (C >= 8 ? 8 : (C >= 4 ? 4 : (C >= 2 ? 2 : 1)))));
    int PY = 256 / CT;
    int gx = (C + CT - 1) / CT;
    long long gy = (148 * 4 + gx - 1) / gx;
    long long maxy = (M + PY - 1) / PY;
    if (gy > maxy) gy = maxy;
    if (gy < 1) gy = 1;
    grid = dim3(gx, (unsigned)gy); block = dim3(CT, PY);
}

}  // namespace

int launch_dwconv_fwd(const float* x, const float* k, float* y, int B, int H, int W, int C, cudaStream_t st, double* stats, int rev) {
    if (too_big((long long)B * H * W * C)) return CRNN_ERR_INVALID;
    { const int rc = launch_dwconv_rows(x, k, y, B, H, W, C, 0, stats, rev, st); if (rc <= 0) return rc; }   // row-marching kernel (dwconv_rows.cu) when the shape allows
    if (C % 4 == 0) {
        const int WG = (W + 3) / 4; const long long ngroups = (long long)B * H * WG;
        dim3 grid, block; chan_block(C / 4, ngroups, grid, block);
        if (stats) (void)crnn_launch(dwconv3x3_cb_kernel<false, true>, grid, block, sizeof(double) * 8 * 256, st, x, k, y, H, W, C / 4, WG, (int)ngroups, stats, rev);
        else (void)crnn_launch(dwconv3x3_cb_kernel<false, false>, grid, block, 0, st, x, k, y, H, W, C / 4, WG, (int)ngroups, nullptr, rev);
    }
    else if (C == 1 && !stats) { long long total = (long long)B * H * W; (void)crnn_launch(dwconv3x3_c1<false>, grid1d(total, 256), 256, 0, st, x, k, y, B, H, W, total); }
    else { crnn_set_error("dwconv: C must be 1 or a multiple of 4 (fused statistics need C %% 4 == 0)"); return CRNN_ERR_INVALID; }
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_dwconv_bwd_data(const float* dy, const float* k, float* dx, int B, int H, int W, int C, int accumulate, cudaStream_t st, int rev,
                           const DwRowsRed* red, double* red_buf, int* red_done) {
    if (accumulate) { crnn_set_error("dwconv_bwd_data: accumulate unsupported"); return CRNN_ERR_INVALID; }
    if (too_big((long long)B * H * W * C)) return CRNN_ERR_INVALID;
    if (red_done) *red_done = 0;
    if (red && red_buf) {
        const int rc = launch_dwconv_rows(dy, k, dx, B, H, W, C, 1, red_buf, rev, st, red);
        if (rc < 0) return rc;
        if (rc == 0) { if (red_done) *red_done = 1; return rc; }
    }
    { const int rc = launch_dwconv_rows(dy, k, dx, B, H, W, C, 1, nullptr, rev, st); if (rc <= 0) return rc; }
    if (C % 4 == 0) {
        const int WG = (W + 3) / 4; const long long ngroups = (long long)B * H * WG;
        dim3 grid, block; chan_block(C / 4, ngroups, grid, block);
        (void)crnn_launch(dwconv3x3_cb_kernel<true, false>, grid, block, 0, st, dy, k, dx, H, W, C / 4, WG, (int)ngroups, nullptr, rev);
    }
    else if (C == 1) { long long total = (long long)B * H * W; (void)crnn_launch(dwconv3x3_c1<true>, grid1d(total, 256), 256, 0, st, dy, k, dx, B, H, W, total); }
    else { crnn_set_error("dwconv: C must be 1 or a multiple of 4"); return CRNN_ERR_INVALID; }
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_dwconv_bwd_weight(const float* x, const float* dy, float* dk, int B, int H, int W, int C, cudaStream_t st) {
    { const int rc = launch_dwconv_rows_bwd_weight(x, dy, dk, B, H, W, C, st); if (rc <= 0) return rc; }
    dim3 grid, block; long long npix = (long long)B * H * W;
    if (C % 4 == 0) {
        const int WG = (W + 3) / 4; const long long ngroups = (long long)B * H * WG;
        chan_block(C / 4, ngroups, grid, block);
        (void)crnn_launch(dwconv3x3_bwd_weight_vec4, grid, block, sizeof(float) * 36 * 256, st, x, dy, dk, B, H, W, C / 4, WG, ngroups);
    } else {
        chan_block(C, npix, grid, block);
        (void)crnn_launch(dwconv3x3_bwd_weight, grid, block, sizeof(float) * block.y * 9 * block.x, st, x, dy, dk, B, H, W, C, npix);
    }
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_colstats(const float* y, long long M, int C, double* stats, cudaStream_t st) {
    dim3 grid, block; chan_block(C, M, grid, block);
    const BnFin fin = crnn_take_bn_fin();
    (void)crnn_launch(colstats_kernel, grid, block, sizeof(double) * 2 * 256, st, y, M, C, stats, fin);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_bn_finalize(const double* stats, long long M, int C, const float* gamma, const float* beta, float* mm, float* mv,
                       float eps, float momentum, int training, float* scale, float* shift, float* save_mean, float* save_invstd, cudaStream_t st) {
    (void)crnn_launch(bn_finalize_kernel, ceil_div(C, 128), 128, 0, st, stats, (double)M, C, gamma, beta, mm, mv, eps, momentum, training, scale, shift, save_mean, save_invstd);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_act_pool_fwd(const float* y, const float* scale, const float* shift, float* a, int B, int H, int W, int C, int ph, int pw,
                        float rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr, int rev) {
    if (C % 4 || H % ph || W % pw || ph * pw > 4) { crnn_set_error("act_pool: unsupported shape"); return CRNN_ERR_INVALID; }
    const long long npix = (long long)B * (H / ph) * (W / pw);
    if (too_big((long long)B * H * W * C)) return CRNN_ERR_INVALID;
    const float ik = rate > 0.f ? 1.f / (1.f - rate) : 1.f;
    dim3 grid, block; chan_block(C / 4, npix, grid, block);
    grid.y = (unsigned)std::min<long long>((npix + block.y - 1) / block.y, (long long)grid.y * 2);     // ~8 CTAs per SM: short dependent chains, many loads in flight
#define APF(PH_, PW_) (void)crnn_launch(act_pool_fwd_kernel<PH_, PW_>, grid, block, 0, st, y, scale, shift, a, W, C / 4, W / pw, (int)npix, rate, ik, seed, layer, seed_ptr, rev)
    if (ph == 1 && pw == 1) APF(1, 1);
    else if (ph == 2 && pw == 2) APF(2, 2);
    else if (ph == 1 && pw == 2) APF(1, 2);
    else { crnn_set_error("act_pool: unsupported pool %dx%d", ph, pw); return CRNN_ERR_INVALID; }
#undef APF
    LAUNCH_CHECK(); return CRNN_OK;
}
static int launch_bn_relu6_reduce(const float* da, const float* y, const float* scale, const float* shift, const float* mean, const float* invstd,
                                  double* red, long long M, int C, int rev, float rate, uint64_t seed, uint32_t layer, const uint64_t* seed_ptr, cudaStream_t st) {
    const int C4 = C / 4;
    const int CQ = C4 >= 64 ? 64 : (C4 >= 32 ? 32 : (C4 >= 16 ? 16 : (C4 >= 8 ? 8 : (C4 >= 4 ? 4 : (C4 >= 2 ? 2 : 1)))));
    const int PY = 256 / CQ, gx = (C4 + CQ - 1) / CQ;
    long long gy = (148 * 3 + gx - 1) / gx;                         // 3 resident CTAs per SM (launch bounds), one wave
    const long long maxy = (M + 4LL * PY - 1) / (4LL * PY);
    if (gy > maxy) gy = maxy;
    if (gy < 1) gy = 1;
    const dim3 grid(gx, (unsigned)gy), block(CQ, PY);
    const size_t sm = sizeof(float) * 8 * 256;
    if (rate > 0.f) (void)crnn_launch(bn_relu6_reduce_kernel<true>, grid, block, sm, st, da, y, scale, shift, mean, invstd, red, (int)M, C4, rev, rate, 1.f / (1.f - rate), seed, layer, seed_ptr);
    else (void)crnn_launch(bn_relu6_reduce_kernel<false>, grid, block, sm, st, da, y, scale, shift, mean, invstd, red, (int)M, C4, rev, 0.f, 1.f, 0, 0, nullptr);
    LAUNCH_CHECK(); return CRNN_OK;
}
// two launches: reductions, then apply (+ a tiny launch for dgamma/dbeta); `red` (double[2C]) must be pre-zeroed
int launch_act_pool_bn_bwd(const float* da, const float* y, const float* scale, const float* shift, const float* mean, const float* invstd,
                           const float* gamma, float* dy, double* red, float* dgamma, float* dbeta,
                           int B, int H, int W, int C, int ph, int pw, float rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr, int rev,
                           int emit_param_grads, int reduce_done) {
    if (C % 4 || H % ph || W % pw || ph * pw > 4) { crnn_set_error("act_pool: unsupported shape"); return CRNN_ERR_INVALID; }
    const long long npix = (long long)B * (H / ph) * (W / pw);
    if (too_big((long long)B * H * W * C)) return CRNN_ERR_INVALID;
    const double invM = 1.0 / ((double)B * H * W);
    const float ik = rate > 0.f ? 1.f / (1.f - rate) : 1.f;
    dim3 grid, block; chan_block(C / 4, npix, grid, block);
#define APB(A_, PH_, PW_, SM_) (void)crnn_launch(act_pool_bwd_kernel<A_, PH_, PW_>, grid, block, SM_, st, da, y, scale, shift, mean, invstd, gamma, dy, red, B, H, W, C / 4, invM, rate, ik, seed, layer, npix, seed_ptr, (A_) ? !rev : rev)
    const size_t sm = sizeof(float) * 8 * 256;
    if (reduce_done && (ph != 1 || pw != 1)) { crnn_set_error("act_pool_bn_bwd: a fused reduction only exists for the non-pooled blocks"); return CRNN_ERR_INVALID; }
    if (ph == 1 && pw == 1) {
        if (!reduce_done) {
            const int rc = launch_bn_relu6_reduce(da, y, scale, shift, mean, invstd, red, (long long)B * H * W, C, rev, rate, seed, layer, seed_ptr, st);
            if (rc != CRNN_OK) return rc;
        }
        APB(true, 1, 1, 0);
    }
    else if (ph == 2 && pw == 2) { APB(false, 2, 2, sm); LAUNCH_CHECK(); APB(true, 2, 2, 0); }
    else if (ph == 1 && pw == 2) { APB(false, 1, 2, sm); LAUNCH_CHECK(); APB(true, 1, 2, 0); }
    else { crnn_set_error("act_pool: unsupported pool %dx%d", ph, pw); return CRNN_ERR_INVALID; }
#undef APB
    LAUNCH_CHECK();
    if (!emit_param_grads) return CRNN_OK;
    (void)crnn_launch(bn_param_grads_kernel, ceil_div(C, 128), 128, 0, st, red, dgamma, dbeta, C);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_relu6_bn_bwd(const float* da, const float* y, const float* scale, const float* shift, const float* mean, const float* invstd,
                        const float* gamma, float* dy, double* red, float* dgamma, float* dbeta, long long M, int C, cudaStream_t st, int rev, int reduce_done, int emit_param_grads) {
    dim3 grid, block;
    const double invM = 1.0 / (double)M;
    if (reduce_done && (C % 4)) { crnn_set_error("relu6_bn_bwd: fused reduction needs C %% 4 == 0"); return CRNN_ERR_INVALID; }
    if (C % 4 == 0) {
        chan_block(C / 4, (M + 1) / 2, grid, block);
        if (too_big(M * C)) return CRNN_ERR_INVALID;
        if (!reduce_done) { const int rc = launch_bn_relu6_reduce(da, y, scale, shift, mean, invstd, red, M, C, rev, 0.f, 0, 0, nullptr, st); if (rc != CRNN_OK) return rc; }
        (void)crnn_launch(relu6_bwd_kernel<true>, grid, block, 0, st, da, y, scale, shift, mean, invstd, gamma, dy, red, M, C / 4, invM, !rev);
    } else {
        chan_block(C, M, grid, block);
        (void)crnn_launch(relu6_bwd_scalar_kernel<false>, grid, block, sizeof(float) * 2 * 256, st, da, y, scale, shift, mean, invstd, gamma, dy, red, M, C, invM);
        LAUNCH_CHECK();
        (void)crnn_launch(relu6_bwd_scalar_kernel<true>, grid, block, 0, st, da, y, scale, shift, mean, invstd, gamma, dy, red, M, C, invM);
    }
    LAUNCH_CHECK();
    if (!emit_param_grads) return CRNN_OK;
    (void)crnn_launch(bn_param_grads_kernel, ceil_div(C, 128), 128, 0, st, red, dgamma, dbeta, C);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_bn_param_grads(const double* red, float* dgamma, float* dbeta, int C, cudaStream_t st) {
    (void)crnn_launch(bn_param_grads_kernel, ceil_div(C, 128), 128, 0, st, red, dgamma, dbeta, C);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_bn_param_grads_all(const BnGradTable& t, cudaStream_t st) {
    if (t.n <= 0) return CRNN_OK;
    (void)crnn_launch(bn_param_grads_all_kernel, dim3(2, t.n), 256, 0, st, t);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_bn_bwd_apply(float* dz, const float* y, const double* red, const float* gamma, const float* mean, const float* invstd,
                        float* dgamma, float* dbeta, long long M, int C, cudaStream_t st) {
    long long total = M * C;
    (void)crnn_launch(bn_bwd_apply_kernel, grid1d(total, 256), 256, 0, st, dz, y, red, gamma, mean, invstd, dgamma, dbeta, M, C, total);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_colsum(const float* y, long long M, int C, int ldy, float* out, cudaStream_t st) {
    dim3 grid, block; chan_block(C, M, grid, block);
    (void)crnn_launch(colsum_kernel, grid, block, sizeof(float) * 256, st, y, M, C, ldy, out);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_relu_dropout_bwd(float* g, const float* act, long long n, float rate, uint64_t, uint32_t, cudaStream_t st) {
    (void)crnn_launch(relu_dropout_bwd_kernel, grid1d(n, 256), 256, 0, st, g, act, n, rate > 0.f ? 1.f / (1.f - rate) : 1.f);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_dropout_fwd(float* x, long long n, float rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr) {
    if (rate <= 0.f) return CRNN_OK;
    (void)crnn_launch(dropout_kernel, grid1d(n, 256), 256, 0, st, x, x, n, rate, 1.f / (1.f - rate), seed, layer, seed_ptr);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_sum_partials(const float* part, int S, long long stride, long long M, int N, const float* bias, int relu, float* out, int ldo,
                        float rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr) {
    if ((N % 4) || (ldo % 4) || (stride % 4) || S < 1) { crnn_set_error("sum_partials: N, ldo and stride must be multiples of 4"); return CRNN_ERR_INVALID; }
    const long long total4 = M * (N / 4);
    (void)crnn_launch(sum_partials_kernel, grid1d(total4, 256), 256, 0, st, part, S, stride, N / 4, total4, bias, relu, out, ldo, rate, rate > 0.f ? 1.f / (1.f - rate) : 1.f, seed, layer, seed_ptr);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_dropout_copy(const float* in, float* out, long long n, float rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr) {
    (void)crnn_launch(dropout_kernel, grid1d(n, 256), 256, 0, st, in, out, n, rate, 1.f / (1.f - rate), seed, layer, seed_ptr);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_pw1_fwd(const float* x, const float* scale, const float* shift, const float* w, float* out, long long M, int Cout, double* stats, cudaStream_t st, int rev) {
    if (Cout != 64 || too_big(M * Cout)) { crnn_set_error("pw1_fwd: Cout must be 64"); return CRNN_ERR_INVALID; }
    const int grid = (int)std::min<long long>((M + 15) / 16, 148 * 8);
    BnFin fin = {}; if (stats) fin = crnn_take_bn_fin();
    (void)crnn_launch(pw1_fwd_kernel, grid, 256, 0, st, x, scale, shift, w, out, (int)M, Cout / 4, stats, rev, fin);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_pw1_bwd(const float* x, const float* scale, const float* shift, const float* dY, const float* w, float* dX, float* dW, long long M, int Cout, cudaStream_t st, int rev) {
    if (Cout != 64 || too_big(M * Cout)) { crnn_set_error("pw1_bwd: Cout must be 64"); return CRNN_ERR_INVALID; }
    const int grid = (int)std::min<long long>((M + 15) / 16, 148 * 8);
    (void)crnn_launch(pw1_bwd_kernel, grid, 256, 0, st, x, scale, shift, dY, w, dX, dW, (int)M, rev);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_set_u64(uint64_t* p, uint64_t v, cudaStream_t st) {
    (void)crnn_launch(set_u64_kernel, 1, 1, 0, st, p, v);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_sum_dirs(const float* hs, float* out, long long rows, int U, cudaStream_t st) {
    (void)crnn_launch(sum_dirs_kernel, grid1d(rows * U, 256), 256, 0, st, hs, out, rows, U); LAUNCH_CHECK(); return CRNN_OK;
}
int launch_dup_dirs(const float* g, float* out, long long rows, int U, cudaStream_t st) {
    (void)crnn_launch(dup_dirs_kernel, grid1d(rows * U, 256), 256, 0, st, g, out, rows, U); LAUNCH_CHECK(); return CRNN_OK;
}
int launch_softmax_rows(const float* z, float* p, long long rows, int V, cudaStream_t st) {
    (void)crnn_launch(softmax_rows_kernel, ceil_div(rows, 8), 256, 0, st, z, p, rows, V); LAUNCH_CHECK(); return CRNN_OK;
}
