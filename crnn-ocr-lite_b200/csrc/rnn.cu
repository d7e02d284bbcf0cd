// rnn.cu -- Bidirectional GRU / LSTM recurrence over the width axis (reference utils.py:77-82; Keras 2.2.2 GRUCell
// reset_after=False / LSTMCell, recurrent_activation hard_sigmoid), forward and BPTT.
//
// v1 layout: the input projections x@W+b of BOTH directions are done by the shared GEMM beforehand; this kernel runs
// only the strictly sequential part.  One CTA = RB batch rows of one direction (gridDim.y = 2 directions run
// concurrently), one thread per hidden unit; h lives in shared memory, the recurrent matrix U (256 x G*256 fp32,
// 0.75-1 MB, does not fit one SM) is streamed from L2 every step with fully coalesced rows.
//   xp    (B,T,2,G*U)   x@W+b, gate order [z,r,h] (GRU) / [i,f,c,o] (LSTM)
//   hs    (B,T,2,U)     outputs in time order (backward direction already re-reversed)
//   gates (B,T,2,GS*U)  saved activations for BPTT: GRU [z,r,hh], LSTM [i,f,g,o,c]
#include "common.cuh"
#include "kernels.h"

namespace {
constexpr int UNITS = 256;

template <int RB>
__global__ void __launch_bounds__(UNITS)
gru_fwd_kernel(const float* __restrict__ xp, const float* __restrict__ U0, const float* __restrict__ U1,
               float* __restrict__ hs, float* __restrict__ gates, int B, int T)
{
    constexpr int U = UNITS, G = 3;
    __shared__ float h[RB][U];
    __shared__ float rh[RB][U];
    const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * RB;
    const float* Um = dir ? U1 : U0;
#pragma unroll
    for (int r = 0; r < RB; ++r) h[r][j] = 0.f;
    __syncthreads();
    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        float az[RB], ar[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) { az[r] = 0.f; ar[r] = 0.f; }
#pragma unroll 8
        for (int k = 0; k < U; ++k) {
            float uz = __ldg(Um + (size_t)k * (G * U) + j), ur = __ldg(Um + (size_t)k * (G * U) + U + j);
#pragma unroll
            for (int r = 0; r < RB; ++r) { float hv = h[r][k]; az[r] = fmaf(hv, uz, az[r]); ar[r] = fmaf(hv, ur, ar[r]); }
        }
        float z[RB], rr[RB], xh[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            int b = b0 + r; if (b >= B) b = B - 1;
            const float* x = xp + (((size_t)b * T + t) * 2 + dir) * (G * U);
            z[r] = hard_sigmoid(x[j] + az[r]);
            rr[r] = hard_sigmoid(x[U + j] + ar[r]);
            xh[r] = x[2 * U + j];
            rh[r][j] = rr[r] * h[r][j];
        }
        __syncthreads();
        float ah[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) ah[r] = 0.f;
#pragma unroll 8
        for (int k = 0; k < U; ++k) {
            float uh = __ldg(Um + (size_t)k * (G * U) + 2 * U + j);
#pragma unroll
            for (int r = 0; r < RB; ++r) ah[r] = fmaf(rh[r][k], uh, ah[r]);
        }
        float hn[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            float hh = tanhf(xh[r] + ah[r]);
            hn[r] = z[r] * h[r][j] + (1.f - z[r]) * hh;
            int b = b0 + r;
            if (b < B) {
                size_t o = ((size_t)b * T + t) * 2 + dir;
                hs[o * U + j] = hn[r];
                if (gates) { float* g = gates + o * (3 * U); g[j] = z[r]; g[U + j] = rr[r]; g[2 * U + j] = hh; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RB; ++r) h[r][j] = hn[r];
        __syncthreads();
    }
}

template <int RB>
__global__ void __launch_bounds__(UNITS)
lstm_fwd_kernel(const float* __restrict__ xp, const float* __restrict__ U0, const float* __restrict__ U1,
                float* __restrict__ hs, float* __restrict__ gates, int B, int T)
{
    constexpr int U = UNITS, G = 4;
    __shared__ float h[RB][U];
    const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * RB;
    const float* Um = dir ? U1 : U0;
    float c[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) { h[r][j] = 0.f; c[r] = 0.f; }
    __syncthreads();
    for (int s = 0; s < T; ++s) {
        const int t = dir ? T - 1 - s : s;
        float a[RB][4];
#pragma unroll
        for (int r = 0; r < RB; ++r) { a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.f; }
#pragma unroll 4
        for (int k = 0; k < U; ++k) {
            const float* ur = Um + (size_t)k * (G * U) + j;
            float u0 = __ldg(ur), u1 = __ldg(ur + U), u2 = __ldg(ur + 2 * U), u3 = __ldg(ur + 3 * U);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                float hv = h[r][k];
                a[r][0] = fmaf(hv, u0, a[r][0]); a[r][1] = fmaf(hv, u1, a[r][1]);
                a[r][2] = fmaf(hv, u2, a[r][2]); a[r][3] = fmaf(hv, u3, a[r][3]);
            }
        }
        float hn[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            int b = b0 + r; int bb = b < B ? b : B - 1;
            const float* x = xp + (((size_t)bb * T + t) * 2 + dir) * (G * U);
            float ig = hard_sigmoid(x[j] + a[r][0]);
            float fg = hard_sigmoid(x[U + j] + a[r][1]);
            float gg = tanhf(x[2 * U + j] + a[r][2]);
            float og = hard_sigmoid(x[3 * U + j] + a[r][3]);
            c[r] = fg * c[r] + ig * gg;
            hn[r] = og * tanhf(c[r]);
            if (b < B) {
                size_t o = ((size_t)b * T + t) * 2 + dir;
                hs[o * U + j] = hn[r];
                if (gates) { float* g = gates + o * (5 * U); g[j] = ig; g[U + j] = fg; g[2 * U + j] = gg; g[3 * U + j] = og; g[4 * U + j] = c[r]; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RB; ++r) h[r][j] = hn[r];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------- BPTT
// UT = transposed recurrent kernels (2, G*U, U): UT[dir][g*U + j'][k] = U_dir[k][g*U + j'] (coalesced mat-vec over j').
template <int RB>
__global__ void __launch_bounds__(UNITS)
gru_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ hs, const float* __restrict__ gates,
               const float* __restrict__ UT, float* __restrict__ dxp, float* __restrict__ hprev_out, float* __restrict__ rh_out,
               int B, int T)
{
    constexpr int U = UNITS, G = 3;
    __shared__ float sa[RB][3][U];   // da_z, da_r, da_h
    const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * RB;
    const float* Ut = UT + (size_t)dir * (G * U) * U;
    float dh[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) dh[r] = 0.f;
    for (int s = T - 1; s >= 0; --s) {
        const int t = dir ? T - 1 - s : s;
        const int tp = dir ? t + 1 : t - 1;
        float z[RB], rr[RB], hp[RB], dz[RB], dhn[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            int b = b0 + r; if (b >= B) b = B - 1;
            size_t o = ((size_t)b * T + t) * 2 + dir;
            const float* g = gates + o * (3 * U);
            z[r] = g[j]; rr[r] = g[U + j]; float hh = g[2 * U + j];
            hp[r] = (s > 0) ? hs[(((size_t)b * T + tp) * 2 + dir) * U + j] : 0.f;
            float dht = dout[o * U + j] + dh[r];
            float dhh = dht * (1.f - z[r]);
            dz[r] = dht * (hp[r] - hh);
            dhn[r] = dht * z[r];
            sa[r][2][j] = dhh * (1.f - hh * hh);
        }
        __syncthreads();
        float drh[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) drh[r] = 0.f;
#pragma unroll 8
        for (int k = 0; k < U; ++k) {
            float u = __ldg(Ut + (size_t)(2 * U + k) * U + j);
#pragma unroll
            for (int r = 0; r < RB; ++r) drh[r] = fmaf(sa[r][2][k], u, drh[r]);
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            float dr = drh[r] * hp[r];
            dhn[r] = fmaf(drh[r], rr[r], dhn[r]);
            sa[r][0][j] = (z[r] > 0.f && z[r] < 1.f) ? 0.2f * dz[r] : 0.f;
            sa[r][1][j] = (rr[r] > 0.f && rr[r] < 1.f) ? 0.2f * dr : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < U; ++k) {
            float uz = __ldg(Ut + (size_t)k * U + j), ur = __ldg(Ut + (size_t)(U + k) * U + j);
#pragma unroll
            for (int r = 0; r < RB; ++r) dhn[r] = fmaf(sa[r][0][k], uz, fmaf(sa[r][1][k], ur, dhn[r]));
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            int b = b0 + r;
            if (b < B) {
                size_t o = ((size_t)b * T + t) * 2 + dir;
                float* d = dxp + o * (G * U);
                d[j] = sa[r][0][j]; d[U + j] = sa[r][1][j]; d[2 * U + j] = sa[r][2][j];
                hprev_out[o * U + j] = hp[r];
                rh_out[o * U + j] = rr[r] * hp[r];
            }
            dh[r] = dhn[r];
        }
        __syncthreads();
    }
}

template <int RB>
__global__ void __launch_bounds__(UNITS)
lstm_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ hs, const float* __restrict__ gates,
                const float* __restrict__ UT, float* __restrict__ dxp, float* __restrict__ hprev_out, int B, int T)
{
    constexpr int U = UNITS, G = 4;
    __shared__ float sa[RB][4][U];
    const int j = threadIdx.x, dir = blockIdx.y, b0 = blockIdx.x * RB;
    const float* Ut = UT + (size_t)dir * (G * U) * U;
    float dh[RB], dc[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) { dh[r] = 0.f; dc[r] = 0.f; }
    for (int s = T - 1; s >= 0; --s) {
        const int t = dir ? T - 1 - s : s;
        const int tp = dir ? t + 1 : t - 1;
        float hp[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            int b = b0 + r; if (b >= B) b = B - 1;
            size_t o = ((size_t)b * T + t) * 2 + dir;
            const float* g = gates + o * (5 * U);
            float ig = g[j], fg = g[U + j], gg = g[2 * U + j], og = g[3 * U + j], cc = g[4 * U + j];
            float cp = 0.f; hp[r] = 0.f;
            if (s > 0) {
                size_t op = ((size_t)b * T + tp) * 2 + dir;
                cp = gates[op * (5 * U) + 4 * U + j]; hp[r] = hs[op * U + j];
            }
            float dht = dout[o * U + j] + dh[r];
            float tc = tanhf(cc);
            float dog = dht * tc;
            float dct = dc[r] + dht * og * (1.f - tc * tc);
            sa[r][0][j] = (ig > 0.f && ig < 1.f) ? 0.2f * dct * gg : 0.f;
            sa[r][1][j] = (fg > 0.f && fg < 1.f) ? 0.2f * dct * cp : 0.f;
            sa[r][2][j] = dct * ig * (1.f - gg * gg);
            sa[r][3][j] = (og > 0.f && og < 1.f) ? 0.2f * dog : 0.f;
            dc[r] = dct * fg;
        }
        __syncthreads();
        float dhn[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) dhn[r] = 0.f;
#pragma unroll 4
        for (int k = 0; k < U; ++k) {
            float u0 = __ldg(Ut + (size_t)k * U + j), u1 = __ldg(Ut + (size_t)(U + k) * U + j);
            float u2 = __ldg(Ut + (size_t)(2 * U + k) * U + j), u3 = __ldg(Ut + (size_t)(3 * U + k) * U + j);
#pragma unroll
            for (int r = 0; r < RB; ++r)
                dhn[r] = fmaf(sa[r][0][k], u0, fmaf(sa[r][1][k], u1, fmaf(sa[r][2][k], u2, fmaf(sa[r][3][k], u3, dhn[r]))));
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            int b = b0 + r;
            if (b < B) {
                size_t o = ((size_t)b * T + t) * 2 + dir;
                float* d = dxp + o * (G * U);
                d[j] = sa[r][0][j]; d[U + j] = sa[r][1][j]; d[2 * U + j] = sa[r][2][j]; d[3 * U + j] = sa[r][3][j];
                hprev_out[o * U + j] = hp[r];
            }
            dh[r] = dhn[r];
        }
        __syncthreads();
    }
}
}  // namespace

int launch_rnn_fwd(int cell, const float* xp, const float* U0, const float* U1, float* hs, float* gates, int B, int T, int U, cudaStream_t st)
{
    if (U != UNITS) { crnn_set_error("rnn: n_units must be %d", UNITS); return CRNN_ERR_INVALID; }
    if (B <= 0) return CRNN_OK;
    const int RB = B <= 74 ? 1 : (B <= 148 ? 2 : 4);
    dim3 grid(ceil_div(B, RB), 2);
#define RUN(K, R) K<R><<<grid, UNITS, 0, st>>>(xp, U0, U1, hs, gates, B, T)
    if (cell == 0) { if (RB == 1) RUN(gru_fwd_kernel, 1); else if (RB == 2) RUN(gru_fwd_kernel, 2); else RUN(gru_fwd_kernel, 4); }
    else           { if (RB == 1) RUN(lstm_fwd_kernel, 1); else if (RB == 2) RUN(lstm_fwd_kernel, 2); else RUN(lstm_fwd_kernel, 4); }
#undef RUN
    LAUNCH_CHECK(); return CRNN_OK;
}

int launch_rnn_bwd(int cell, const float* dout, const float* hs, const float* gates, const float* UT,
                   float* dxp, float* hprev, float* rh, int B, int T, int U, cudaStream_t st)
{
    if (U != UNITS) { crnn_set_error("rnn: n_units must be %d", UNITS); return CRNN_ERR_INVALID; }
    if (B <= 0) return CRNN_OK;
    const int RB = B <= 74 ? 1 : (B <= 148 ? 2 : 4);
    dim3 grid(ceil_div(B, RB), 2);
    if (cell == 0) {
        if (RB == 1) gru_bwd_kernel<1><<<grid, UNITS, 0, st>>>(dout, hs, gates, UT, dxp, hprev, rh, B, T);
        else if (RB == 2) gru_bwd_kernel<2><<<grid, UNITS, 0, st>>>(dout, hs, gates, UT, dxp, hprev, rh, B, T);
        else gru_bwd_kernel<4><<<grid, UNITS, 0, st>>>(dout, hs, gates, UT, dxp, hprev, rh, B, T);
    } else {
        if (RB == 1) lstm_bwd_kernel<1><<<grid, UNITS, 0, st>>>(dout, hs, gates, UT, dxp, hprev, B, T);
        else if (RB == 2) lstm_bwd_kernel<2><<<grid, UNITS, 0, st>>>(dout, hs, gates, UT, dxp, hprev, B, T);
        else lstm_bwd_kernel<4><<<grid, UNITS, 0, st>>>(dout, hs, gates, UT, dxp, hprev, B, T);
    }
    LAUNCH_CHECK(); return CRNN_OK;
}
