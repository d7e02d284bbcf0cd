// dwconv_fused.cu -- (1) backward: ONE kernel for everything between the pointwise dX GEMM and the block below in the backward pass of a
// depthwise-separable block (utils.py:43-52: DepthwiseConv2D 3x3 -> BatchNormalization -> ReLU6), sm_100a, NHWC fp32:
//     dz  = BN backward( dA * 1[0 <= bn(z) <= 6] )          (was relu6_bn_bwd apply: read dA, z; write dz)
//     dx  = depthwise3x3^T(dz)                               (was dwconv3x3_rows<FLIP>: read dz; write dx)
//     dk += sum x (*) dz                                     (was dwconv3x3_rows_bwd_weight on the side stream: read x, dz)
//     red(block below) += [dx*mask*gate, . * xhat]           (was the RED instance of the backward-data kernel: read y of the block below)
// Eight tensor passes over [B,H,W,C] become four or five: dA, z and the block input are read once, dx is written once, dz never
// reaches HBM.  When the block below is not pooled its output (this block's input x) is a pure function of its raw pointwise
// output y -- x = dropout(relu6(bn(y))) -- and y is needed for the fused reduction anyway, so x is RECOMPUTED from y (same fma / mask
// as act_pool_fwd_kernel, bit-identical) and the `block{i-1}` tensor is not read at all.
//
// Layout: a CTA owns (strip of RS rows of the batch's column of B * (H + 1) virtual rows -- see the kernel --, FQ channel quads); 192 threads = FQ quads x FP position slots, a slot = (one of G = 2
// rows, one three-column segment).  FQ is the largest of 8 / 16 / 32 whose 192 / FQ slots still hold two rows of W / 3 segments (W = 36 -> 8,
// 18 -> 16, 9 -> 32), so a warp is uniform in its row and strips stay long (a first version with 8-row iterations at W = 9 and short
// per-image strips ran 2 iterations per CTA at 55 % lane utilisation: ncu r2q, 3.1 TB/s).  dz lives in a shared-memory ring
// of 6 rows (+ zero halo columns): iteration k PRODUCES rows P_k = hs-1+2k .. (each thread its own 3 columns, from registers loaded
// one iteration earlier), one __syncthreads, then CONSUMES output rows O_k = hs-2+2k .. : the 3x5 dz neighbourhood of the
// thread's 3 columns is read once from shared memory (15 LDS.128) and used twice -- 27 fma4 into dx with the taps, 27 fma4 into the
// nine dk accumulators with x.  O_k only touches P_{k-1} and P_k, P_{k+1} goes to the third ring slot, so one barrier per iteration
// is enough.  All global loads of iteration k+1 are issued before the arithmetic of iteration k (9-12 x 16 B per thread in flight).
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& a, const float4 x, const float4 k) {
    a.x = fmaf(x.x, k.x, a.x); a.y = fmaf(x.y, k.y, a.y); a.z = fmaf(x.z, k.z, a.z); a.w = fmaf(x.w, k.w, a.w);
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

constexpr int FT = 192;          // threads per CTA = FQ channel quads x FP position slots
constexpr int G = 2, NR = 3 * G; // rows per iteration, rows of the dz ring
constexpr int NCONST = 9 + 7 + 4;   // taps | BN1: scale shift mean invstd gamma*invstd mean(dz) mean(dz*xhat) | BN2 below: scale shift xa xb

struct FusedArgs {
    const float* dA; const float* z; const float* x; const float* k; float* dx; float* dk;
    const float* scale; const float* shift; const float* mean; const float* invstd; const float* gamma; const double* red1; double invM;
    const float* pscale; const float* pshift; const float* pmean; const float* pinvstd; double* pred;
    float rate, inv_keep; uint64_t seed; uint32_t layer; const uint64_t* seed_ptr;
    int H, W, C4, NS, RS, nstrips, niter, V, rev;      // V = B * (H + 1) virtual rows
};

// RED: a.x is the RAW pointwise output y of the block below (x is recomputed from it) and the BN2-backward reduction of that block is
// accumulated into a.pred;  !RED: a.x is the block input itself.
template <int FQ, bool RED>
__global__ void __launch_bounds__(FT, 2) dwconv3x3_bwd_fused_kernel(const FusedArgs a)
{ pdl_enter();
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int FP = FT / FQ;
    const int NS = a.NS, W = a.W, H = a.H, C4 = a.C4, C = C4 * 4;
    const int WP = W + 2;
    float4* ring = reinterpret_cast<float4*>(smraw);            // [NR][WP][FQ]
    float4* cs = ring + (size_t)NR * WP * FQ;                   // [NCONST][FQ]
    const int tq = threadIdx.x, tp = threadIdx.y;
    const int g = tp / NS, seg = tp - g * NS, w0 = seg * 3;
    const int c4 = blockIdx.x * FQ + tq;
    const bool cok = c4 < C4, act = cok && g < G;
    // The batch is one column of V = B * (H + 1) VIRTUAL rows: image b owns rows b(H+1) .. b(H+1)+H-1, row b(H+1)+H is an all-zero separator
    // (the bottom padding of image b and the top padding of image b+1 at once).  A strip is any RS consecutive virtual rows, so the launcher
    // can cut the batch into exactly as many equal strips as there are CTA slots (whole-image strips: 256 CTAs on 296 slots).  A row cursor
    // {v, r = v mod (H+1), prow = v - image} is advanced by G rows per iteration WITHOUT divisions: decoding v / (H+1) per row made the
    // kernels 20 % slower than whole-image strips (issue-bound), the cursor makes them 2-9 % faster.
    const int HP1 = H + 1, V = a.V;
    const int strip = a.rev ? a.nstrips - 1 - (int)blockIdx.y : (int)blockIdx.y;
    const int hs = strip * a.RS, he = min(V, hs + a.RS);
    struct Cur { int v, r, prow; };
    auto cur_at = [&](int v) { Cur c; const int bb = (v + HP1) / HP1 - 1; c.v = v; c.r = v - bb * HP1; c.prow = v - bb; return c; };
    auto cur_adv = [&](Cur& c) { c.v += G; c.r += G; c.prow += G; if (c.r >= HP1) { c.r -= HP1; c.prow -= 1; } };
    auto cur_ok = [&](const Cur& c) { return c.v >= 0 && c.v < V && c.r != H; };

    if (cok) for (int task = tp; task < (RED ? 11 : 10); task += FP) {
        if (task < 9) cs[task * FQ + tq] = ldg4(a.k + (size_t)task * C + c4 * 4);
        else if (task == 9) {
            const float4 is = ldg4(a.invstd + c4 * 4), ga = ldg4(a.gamma + c4 * 4);
            cs[9 * FQ + tq] = ldg4(a.scale + c4 * 4); cs[10 * FQ + tq] = ldg4(a.shift + c4 * 4);
            cs[11 * FQ + tq] = ldg4(a.mean + c4 * 4); cs[12 * FQ + tq] = is;
            cs[13 * FQ + tq] = make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w);
            const double* r1 = a.red1 + c4 * 4; const double* r2 = a.red1 + C + c4 * 4;
            cs[14 * FQ + tq] = make_float4((float)(r1[0] * a.invM), (float)(r1[1] * a.invM), (float)(r1[2] * a.invM), (float)(r1[3] * a.invM));
            cs[15 * FQ + tq] = make_float4((float)(r2[0] * a.invM), (float)(r2[1] * a.invM), (float)(r2[2] * a.invM), (float)(r2[3] * a.invM));
        } else {
            const float4 xa = ldg4(a.pinvstd + c4 * 4), mu = ldg4(a.pmean + c4 * 4);
            cs[16 * FQ + tq] = ldg4(a.pscale + c4 * 4); cs[17 * FQ + tq] = ldg4(a.pshift + c4 * 4);
            cs[18 * FQ + tq] = xa;                                                        // xhat = y*xa + xb
            cs[19 * FQ + tq] = make_float4(-mu.x * xa.x, -mu.y * xa.y, -mu.z * xa.z, -mu.w * xa.w);
        }
    }
    for (int r = tp; r < NR; r += FP) { ring[((size_t)r * WP) * FQ + tq] = zero4(); ring[((size_t)r * WP + W + 1) * FQ + tq] = zero4(); }
    const uint64_t rseed = (RED && a.seed_ptr) ? *a.seed_ptr : a.seed;
    __syncthreads();

    const size_t rstride = (size_t)W * C;
    const size_t col0 = (size_t)w0 * C + (size_t)c4 * 4;      // (pixel row 0, w0, c4): add pixel row * rstride
    float4 pa[3], pz[3], px[3];
    bool p_ok = false, x_ok = false; int x_row = 0;           // validity / pixel row of the rows whose loads are in flight
    Cur cp = cur_at(hs - 1 + g), cx = cur_at(hs - 2 + g);     // next produce row, next output row
    auto load_p = [&]() {
        p_ok = cur_ok(cp) && act;
        const size_t o = col0 + (size_t)(p_ok ? cp.prow : 0) * rstride;
#pragma unroll
        for (int t = 0; t < 3; ++t) { pa[t] = p_ok ? ldg4(a.dA + o + (size_t)t * C) : zero4(); pz[t] = p_ok ? ldg4(a.z + o + (size_t)t * C) : zero4(); }
        cur_adv(cp);
    };
    auto load_x = [&]() {
        x_ok = cx.v >= hs && cx.v < he && cur_ok(cx) && act;
        x_row = cx.prow;
        const size_t o = col0 + (size_t)(x_ok ? cx.prow : 0) * rstride;
#pragma unroll
        for (int t = 0; t < 3; ++t) px[t] = x_ok ? ldg4(a.x + o + (size_t)t * C) : zero4();
        cur_adv(cx);
    };
    float4 acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = zero4();
    float s[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};

    load_p();
    load_x();
    int pslot = g, cslot = g + NR - 2;       // ring rows of P_k's row (2k+g) and of the first row O_k needs (2k+g-2), mod NR
    for (int k = 0; k < a.niter; ++k) {
        // ---- produce dz row hs-1+kG+g (columns w0..w0+2) into the ring
        if (act) {
            float4* dst = ring + ((size_t)pslot * WP + w0 + 1) * FQ + tq;
            if (p_ok) {
                const float4 sc = cs[9 * FQ + tq], sh = cs[10 * FQ + tq], mu = cs[11 * FQ + tq], is = cs[12 * FQ + tq];
                const float4 gs = cs[13 * FQ + tq], m1 = cs[14 * FQ + tq], m2 = cs[15 * FQ + tq];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    float4 o;
#define DZ1(f) { const float zz = fmaf(pz[t].f, sc.f, sh.f); const float d = (zz >= 0.f && zz <= 6.f) ? pa[t].f : 0.f; \
                 const float xh = (pz[t].f - mu.f) * is.f; o.f = gs.f * (d - m1.f - xh * m2.f); }
                    DZ1(x) DZ1(y) DZ1(z) DZ1(w)
#undef DZ1
                    dst[t * FQ] = o;
                }
            } else {
#pragma unroll
                for (int t = 0; t < 3; ++t) dst[t * FQ] = zero4();
            }
        }
        if (k + 1 < a.niter) load_p();
        float4 xc[3] = {px[0], px[1], px[2]};
        const bool c_ok = x_ok; const int r = x_row;          // this iteration's output row (pixel-row index), before the next load overwrites them
        if (k + 1 < a.niter) load_x(); else x_ok = false;
        __syncthreads();
        // ---- consume: virtual output row hs-2+kG+g needs dz rows -1..+1 around it = ring rows (kG+g-2 .. kG+g) mod NR
        if (c_ok) {
            float dm[3][4];
            float4 yv[3];
            if (RED) {
                const float4 psc = cs[16 * FQ + tq], psh = cs[17 * FQ + tq];
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    yv[o] = xc[o];
                    if (a.rate > 0.f) crnn_dropout_mask4(rseed, a.layer, (uint64_t)(((size_t)r * W + w0 + o) * C4 + c4), a.rate, a.inv_keep, dm[o]);
                    else { dm[o][0] = dm[o][1] = dm[o][2] = dm[o][3] = 1.f; }
                    xc[o].x = relu6f(fmaf(yv[o].x, psc.x, psh.x)); xc[o].y = relu6f(fmaf(yv[o].y, psc.y, psh.y));
                    xc[o].z = relu6f(fmaf(yv[o].z, psc.z, psh.z)); xc[o].w = relu6f(fmaf(yv[o].w, psc.w, psh.w));
                    if (a.rate > 0.f) { xc[o].x *= dm[o][0]; xc[o].y *= dm[o][1]; xc[o].z *= dm[o][2]; xc[o].w *= dm[o][3]; }
                }
            }
            float4 dxv[3] = {zero4(), zero4(), zero4()};
#pragma unroll
            for (int ar = 0; ar < 3; ++ar) {
                const int rs = cslot + ar;
                const float4* src = ring + ((size_t)(rs >= NR ? rs - NR : rs) * WP + w0) * FQ + tq;
                float4 D[5];
#pragma unroll
                for (int c = 0; c < 5; ++c) D[c] = src[c * FQ];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 kt = cs[(8 - (ar * 3 + c)) * FQ + tq];
#pragma unroll
                    for (int o = 0; o < 3; ++o) { fma4(dxv[o], D[c + o], kt); fma4(acc[8 - (ar * 3 + c)], xc[o], D[c + o]); }
                }
            }
            float* dst = a.dx + col0 + (size_t)r * rstride;
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                *reinterpret_cast<float4*>(dst + (size_t)o * C) = dxv[o];
                if (RED) {
                    const float4 psc = cs[16 * FQ + tq], psh = cs[17 * FQ + tq], xa = cs[18 * FQ + tq], xb = cs[19 * FQ + tq];
                    const float d[4] = {dxv[o].x * dm[o][0], dxv[o].y * dm[o][1], dxv[o].z * dm[o][2], dxv[o].w * dm[o][3]};
                    const float yy[4] = {yv[o].x, yv[o].y, yv[o].z, yv[o].w};
                    const float scv[4] = {psc.x, psc.y, psc.z, psc.w}, shv[4] = {psh.x, psh.y, psh.z, psh.w};
                    const float xav[4] = {xa.x, xa.y, xa.z, xa.w}, xbv[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float zz = fmaf(yy[e], scv[e], shv[e]);
                        const float dz = (zz >= 0.f && zz <= 6.f) ? d[e] : 0.f;
                        s[e] += dz; sq[e] = fmaf(dz, fmaf(yy[e], xav[e], xbv[e]), sq[e]);
                    }
                }
            }
        }
        pslot = pslot + G >= NR ? pslot + G - NR : pslot + G;
        cslot = cslot + G >= NR ? cslot + G - NR : cslot + G;
    }
    // ---- CTA reductions: nine taps x 4 channels per thread -> one fp32 atomic per tap/channel; RED sums -> one fp64 atomic pair per channel
    __syncthreads();
    float* red = reinterpret_cast<float*>(smraw);               // [FP][36][FQ]
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        red[((tp * 36) + q * 4 + 0) * FQ + tq] = acc[q].x; red[((tp * 36) + q * 4 + 1) * FQ + tq] = acc[q].y;
        red[((tp * 36) + q * 4 + 2) * FQ + tq] = acc[q].z; red[((tp * 36) + q * 4 + 3) * FQ + tq] = acc[q].w;
    }
    __syncthreads();
    if (cok)
        for (int e = tp; e < 36; e += FP) {
            float sum = 0.f;
            for (int yy = 0; yy < FP; ++yy) sum += red[(yy * 36 + e) * FQ + tq];
            atomicAdd(a.dk + (size_t)(e >> 2) * C + c4 * 4 + (e & 3), sum);
        }
    if (!RED) return;
    __syncthreads();
    double* dsm = reinterpret_cast<double*>(smraw);             // [FP][8][FQ]
#pragma unroll
    for (int e = 0; e < 4; ++e) { dsm[(tp * 8 + e) * FQ + tq] = (double)s[e]; dsm[(tp * 8 + 4 + e) * FQ + tq] = (double)sq[e]; }
    __syncthreads();
    if (cok)
        for (int e = tp; e < 8; e += FP) {
            double t = 0.0;
            for (int i = 0; i < FP; ++i) t += dsm[(i * 8 + e) * FQ + tq];
            atomicAdd(a.pred + (e >> 2) * C + c4 * 4 + (e & 3), t);
        }
}

// ================================================================================================= forward
// (2) forward: the depthwise conv of block i reads the RAW pointwise output y of a non-pooled block i-1 and applies that block's
// BatchNorm + ReLU6 + Dropout (utils.py:53-56) while staging its input rows in shared memory, so the `block{i-1}` tensor is neither
// written (act_pool_fwd launch gone) nor read: two tensor passes (y in, conv out) instead of four.  Same CTA layout and ring as above;
// every input element is transformed exactly once (the register-marching kernel in dwconv_rows.cu re-loads halo columns, which would
// repeat the mask hash 1.67x).  Also accumulates the BatchNorm statistics of its output (training).
struct FwdArgs {
    const float* y; const float* k; float* out; double* stats;
    const float* pscale; const float* pshift; float rate, inv_keep; uint64_t seed; uint32_t layer; const uint64_t* seed_ptr;
    int H, W, C4, NS, RS, nstrips, niter, V, rev;
    BnFin fin;                 // BatchNorm finalize of the output's statistics, done by the last CTA (common.cuh)
};

template <int FQ>
__global__ void __launch_bounds__(FT, 3) dwconv3x3_fwd_fused_kernel(const FwdArgs a)
{ pdl_enter();
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int FP = FT / FQ;
    const int NS = a.NS, W = a.W, H = a.H, C4 = a.C4, C = C4 * 4;
    const int WP = W + 2;
    float4* ring = reinterpret_cast<float4*>(smraw);            // [NR][WP][FQ]
    float4* cs = ring + (size_t)NR * WP * FQ;                   // [9 taps + scale + shift][FQ]
    const int tq = threadIdx.x, tp = threadIdx.y;
    const int g = tp / NS, seg = tp - g * NS, w0 = seg * 3;
    const int c4 = blockIdx.x * FQ + tq;
    const bool cok = c4 < C4, act = cok && g < G;
    // The batch is one column of V = B * (H + 1) VIRTUAL rows: image b owns rows b(H+1) .. b(H+1)+H-1, row b(H+1)+H is an all-zero separator
    // (the bottom padding of image b and the top padding of image b+1 at once).  A strip is any RS consecutive virtual rows, so the launcher
    // can cut the batch into exactly as many equal strips as there are CTA slots.  A row cursor {v, r = v mod (H+1), prow = v - image} is
    // advanced by G rows per iteration without divisions.
    const int HP1 = H + 1, V = a.V;
    const int strip = a.rev ? a.nstrips - 1 - (int)blockIdx.y : (int)blockIdx.y;
    const int hs = strip * a.RS, he = min(V, hs + a.RS);
    struct Cur { int v, r, prow; };
    auto cur_at = [&](int v) { Cur c; const int bb = (v + HP1) / HP1 - 1; c.v = v; c.r = v - bb * HP1; c.prow = v - bb; return c; };
    auto cur_adv = [&](Cur& c) { c.v += G; c.r += G; c.prow += G; if (c.r >= HP1) { c.r -= HP1; c.prow -= 1; } };
    auto cur_ok = [&](const Cur& c) { return c.v >= 0 && c.v < V && c.r != H; };
    if (cok) for (int task = tp; task < 11; task += FP)
        cs[task * FQ + tq] = task < 9 ? ldg4(a.k + (size_t)task * C + c4 * 4) : ldg4((task == 9 ? a.pscale : a.pshift) + c4 * 4);
    for (int r = tp; r < NR; r += FP) { ring[((size_t)r * WP) * FQ + tq] = zero4(); ring[((size_t)r * WP + W + 1) * FQ + tq] = zero4(); }
    const uint64_t rseed = a.seed_ptr ? *a.seed_ptr : a.seed;
    __syncthreads();

    const size_t rstride = (size_t)W * C;
    const size_t col0 = (size_t)w0 * C + (size_t)c4 * 4;
    // input rows are loaded TWO iterations ahead (the conv of one iteration is too short to cover a DRAM round trip: ncu r2s, 17 % barrier +
    // 10 % long-scoreboard stalls with a distance of one)
    float4 py[3], pn[3];
    int rowy = -1, rown = -1;                                 // pixel rows of the two loads in flight (-1: padding row)
    Cur cp = cur_at(hs - 1 + g), co = cur_at(hs - 2 + g);     // next row to load, next output row
    auto load_y = [&](float4 (&d)[3], int& row) {
        const bool ok = cur_ok(cp) && act;
        row = ok ? cp.prow : -1;
        const size_t o = col0 + (size_t)(ok ? cp.prow : 0) * rstride;
#pragma unroll
        for (int t = 0; t < 3; ++t) d[t] = ok ? ldg4(a.y + o + (size_t)t * C) : zero4();
        cur_adv(cp);
    };
    float s[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
    load_y(py, rowy);
    load_y(pn, rown);
    int pslot = g, cslot = g + NR - 2;
    for (int k = 0; k < a.niter; ++k) {
        if (act) {
            const int rho = rowy;
            float4* dst = ring + ((size_t)pslot * WP + w0 + 1) * FQ + tq;
            if (rho >= 0) {
                const float4 sc = cs[9 * FQ + tq], sh = cs[10 * FQ + tq];
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    float4 o;
                    o.x = relu6f(fmaf(py[t].x, sc.x, sh.x)); o.y = relu6f(fmaf(py[t].y, sc.y, sh.y));
                    o.z = relu6f(fmaf(py[t].z, sc.z, sh.z)); o.w = relu6f(fmaf(py[t].w, sc.w, sh.w));
                    if (a.rate > 0.f) {
                        float dm[4];
                        crnn_dropout_mask4(rseed, a.layer, (uint64_t)(((size_t)rho * W + w0 + t) * C4 + c4), a.rate, a.inv_keep, dm);
                        o.x *= dm[0]; o.y *= dm[1]; o.z *= dm[2]; o.w *= dm[3];
                    }
                    dst[t * FQ] = o;
                }
            } else {
#pragma unroll
                for (int t = 0; t < 3; ++t) dst[t * FQ] = zero4();
            }
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) py[t] = pn[t];
        rowy = rown;
        if (k + 2 < a.niter) load_y(pn, rown); else rown = -1;
        __syncthreads();
        const bool o_ok = act && co.v >= hs && co.v < he && cur_ok(co);
        const int r = co.prow;
        cur_adv(co);
        if (o_ok) {
            float4 ov[3] = {zero4(), zero4(), zero4()};
#pragma unroll
            for (int ar = 0; ar < 3; ++ar) {
                const int rs = cslot + ar;
                const float4* src = ring + ((size_t)(rs >= NR ? rs - NR : rs) * WP + w0) * FQ + tq;
                float4 D[5];
#pragma unroll
                for (int c = 0; c < 5; ++c) D[c] = src[c * FQ];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 kt = cs[(ar * 3 + c) * FQ + tq];
#pragma unroll
                    for (int o = 0; o < 3; ++o) fma4(ov[o], D[c + o], kt);
                }
            }
            float* dst = a.out + col0 + (size_t)r * rstride;
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                *reinterpret_cast<float4*>(dst + (size_t)o * C) = ov[o];
                s[0] += ov[o].x; s[1] += ov[o].y; s[2] += ov[o].z; s[3] += ov[o].w;
                sq[0] = fmaf(ov[o].x, ov[o].x, sq[0]); sq[1] = fmaf(ov[o].y, ov[o].y, sq[1]);
                sq[2] = fmaf(ov[o].z, ov[o].z, sq[2]); sq[3] = fmaf(ov[o].w, ov[o].w, sq[3]);
            }
        }
        pslot = pslot + G >= NR ? pslot + G - NR : pslot + G;
        cslot = cslot + G >= NR ? cslot + G - NR : cslot + G;
    }
    if (!a.stats) return;
    __syncthreads();
    double* dsm = reinterpret_cast<double*>(smraw);             // [FP][8][FQ]
#pragma unroll
    for (int e = 0; e < 4; ++e) { dsm[(tp * 8 + e) * FQ + tq] = (double)s[e]; dsm[(tp * 8 + 4 + e) * FQ + tq] = (double)sq[e]; }
    __syncthreads();
    if (cok)
        for (int e = tp; e < 8; e += FP) {
            double t = 0.0;
            for (int i = 0; i < FP; ++i) t += dsm[(i * 8 + e) * FQ + tq];
            atomicAdd(a.stats + (e >> 2) * C + c4 * 4 + (e & 3), t);
        }
    bn_finalize_tail(a.fin);
}

int g_fused_off = -1;

// strips of RS virtual rows (V = B * (H + 1) in total) cost ceil((RS+2)/2) iterations of 2 rows plus a prologue / epilogue worth ~3 iterations
// (constants, first loads, the CTA reductions): choose the strip count that maximises (useful rows per row-time) x (fill of the waves of
// 148 SMs x `occ` resident CTAs); ties -> fewer, longer strips
int plan_strips(int V, int gx, int occ) {
    double best = -1.0; int best_rs = V;
    const long long cap = 148LL * occ;
    for (int n = 1; n <= V; ++n) {
        const int rs = (V + n - 1) / n;
        if (rs < 4 * G && n > 1) break;
        const int d = (V + rs - 1) / rs, it = (rs + 2 + G - 1) / G;
        const long long ctas = (long long)gx * d;
        const double eff = (double)V / ((double)d * (it + 3) * G) * (double)ctas / (double)(((ctas + cap - 1) / cap) * cap);
        if (eff > best + 1e-9) { best = eff; best_rs = rs; }
    }
    return best_rs;
}

// channel quads per CTA for an image width: the largest of 8 / 16 / 32 that leaves 2 rows x W/3 position slots in 192 threads
int fused_fq(int W) { const int ns = W / 3; return ns <= 3 ? 32 : (ns <= 6 ? 16 : 8); }

}  // namespace

// shapes the fused kernel handles (CRNN_DW_FUSED=0: none -> the caller keeps the separate kernels)
int dwconv_bwd_fused_covers(int H, int W, int C) {
    if (g_fused_off < 0) { const char* e = getenv("CRNN_DW_FUSED"); g_fused_off = (e && e[0] == '0') ? 1 : 0; }
    return !(g_fused_off || C % 4 || W % 3 || W < 3 || W > 36 || H < 1);
}

int launch_dwconv_bwd_fused(const float* dA, const float* z, const float* x_or_y, const float* k, float* dx, float* dk,
                            const float* scale, const float* shift, const float* mean, const float* invstd, const float* gamma, const double* red1,
                            int B, int H, int W, int C, int rev, const DwRowsRed* red, double* red_buf, cudaStream_t st)
{
    if (!dwconv_bwd_fused_covers(H, W, C)) { crnn_set_error("dwconv_bwd_fused: shape not covered"); return CRNN_ERR_INVALID; }
    if ((long long)B * H * W * C >= (1LL << 31)) { crnn_set_error("dwconv_bwd_fused: tensor too large"); return CRNN_ERR_INVALID; }
    FusedArgs a = {};
    a.dA = dA; a.z = z; a.x = x_or_y; a.k = k; a.dx = dx; a.dk = dk;
    a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd; a.gamma = gamma; a.red1 = red1; a.invM = 1.0 / ((double)B * H * W);
    a.H = H; a.W = W; a.C4 = C / 4; a.NS = W / 3; a.rev = rev;
    if (red) {
        a.pscale = red->scale; a.pshift = red->shift; a.pmean = red->mean; a.pinvstd = red->invstd; a.pred = red_buf;
        a.rate = red->rate; a.inv_keep = red->rate > 0.f ? 1.f / (1.f - red->rate) : 1.f; a.seed = red->seed; a.layer = red->layer; a.seed_ptr = red->seed_ptr;
    }
    const int FQ = fused_fq(W);
    const int gx = (a.C4 + FQ - 1) / FQ;
    a.V = B * (H + 1);
    a.RS = plan_strips(a.V, gx, 2); a.nstrips = (a.V + a.RS - 1) / a.RS; a.niter = (a.RS + 2 + G - 1) / G;
    const size_t ring = sizeof(float4) * ((size_t)NR * (W + 2) + NCONST) * FQ;
    const size_t sm = std::max(ring, sizeof(float) * 36 * FT);
    static bool attr_done = false;
    if (!attr_done) {
#define FATTR(FQ_, R_) cudaFuncSetAttribute(dwconv3x3_bwd_fused_kernel<FQ_, R_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)
        FATTR(8, false); FATTR(8, true); FATTR(16, false); FATTR(16, true); FATTR(32, false); FATTR(32, true);
#undef FATTR
        attr_done = true;
    }
    g_crnn_family = CRNN_FAM_DWROWS;
    const dim3 grid(gx, (unsigned)a.nstrips), block(FQ, FT / FQ);
#define FLAUNCH(FQ_) do { if (red) (void)crnn_launch(dwconv3x3_bwd_fused_kernel<FQ_, true>, grid, block, sm, st, a); else (void)crnn_launch(dwconv3x3_bwd_fused_kernel<FQ_, false>, grid, block, sm, st, a); } while (0)
    if (FQ == 8) FLAUNCH(8); else if (FQ == 16) FLAUNCH(16); else FLAUNCH(32);
#undef FLAUNCH
    LAUNCH_CHECK();
    return CRNN_OK;
}

// out = depthwise3x3( dropout(relu6(y * pscale + pshift)) ), + per-channel sum / sum of squares of out into stats (optional, pre-zeroed double[2C])
int launch_dwconv_fwd_fused(const float* y, const float* pscale, const float* pshift, float rate, uint64_t seed, uint32_t layer, const uint64_t* seed_ptr,
                            const float* k, float* out, double* stats, int B, int H, int W, int C, int rev, cudaStream_t st)
{
    if (!dwconv_bwd_fused_covers(H, W, C)) { crnn_set_error("dwconv_fwd_fused: shape not covered"); return CRNN_ERR_INVALID; }
    if ((long long)B * H * W * C >= (1LL << 31)) { crnn_set_error("dwconv_fwd_fused: tensor too large"); return CRNN_ERR_INVALID; }
    FwdArgs a = {};
    a.y = y; a.k = k; a.out = out; a.stats = stats; a.pscale = pscale; a.pshift = pshift;
    a.rate = rate; a.inv_keep = rate > 0.f ? 1.f / (1.f - rate) : 1.f; a.seed = seed; a.layer = layer; a.seed_ptr = seed_ptr;
    a.H = H; a.W = W; a.C4 = C / 4; a.NS = W / 3; a.rev = rev;
    if (stats) a.fin = crnn_take_bn_fin();
    const int FQ = fused_fq(W);
    const int gx = (a.C4 + FQ - 1) / FQ;
    a.V = B * (H + 1);
    a.RS = plan_strips(a.V, gx, 3); a.nstrips = (a.V + a.RS - 1) / a.RS; a.niter = (a.RS + 2 + G - 1) / G;
    const size_t sm = std::max(sizeof(float4) * ((size_t)NR * (W + 2) + 11) * FQ, sizeof(double) * 8 * FT);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(dwconv3x3_fwd_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(dwconv3x3_fwd_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(dwconv3x3_fwd_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        attr_done = true;
    }
    g_crnn_family = CRNN_FAM_DWROWS;
    const dim3 grid(gx, (unsigned)a.nstrips), block(FQ, FT / FQ);
    if (FQ == 8) (void)crnn_launch(dwconv3x3_fwd_fused_kernel<8>, grid, block, sm, st, a);
    else if (FQ == 16) (void)crnn_launch(dwconv3x3_fwd_fused_kernel<16>, grid, block, sm, st, a);
    else (void)crnn_launch(dwconv3x3_fwd_fused_kernel<32>, grid, block, sm, st, a);
    LAUNCH_CHECK();
    return CRNN_OK;
}
