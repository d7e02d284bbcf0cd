// dwconv_rows.cu -- row-marching DepthwiseConv2D 3x3 'same' kernels (forward, backward-data, backward-weight) for sm_100a.
// Reference op: utils.py:44-46 (DepthwiseConv2D(3x3, same, stride 1, depth_multiplier 1, no bias)), NHWC fp32.
//
// The channel-block kernel in conv.cu gives every thread one (row, 4-column group) item: 18 input loads for 4 outputs, i.e. every
// input element is fetched 4.5 times (L1/L2 absorb it, but the LSU does not) and ~210 instructions are issued per output float4;
// ncu r1h shows it at 20-26 % of HBM peak, 52 % of its stalls on the long scoreboard, 2 CTAs per SM.  An HBM-bound op should not be
// instruction/latency bound, so here a thread owns 4 channels x 3 adjacent columns and MARCHES DOWN a strip of image rows:
//   * input row r (5 float4: the 3 columns + 1 halo column each side) is loaded once and scattered into THREE output-row
//     accumulators (rows r-1, r, r+1), so vertical reuse lives in registers: 5 loads per 3 outputs (1.67x instead of 4.5x);
//   * the next row's loads are issued before the current row's 108 FFMAs (register double buffer) -> 10 independent 16-byte
//     loads in flight per thread, no shared-memory staging, no barriers;
//   * the nine taps are read from shared memory (conflict-free LDS.128) instead of occupying 36 registers;
//   * ~45 issued instructions per output float4.
// Strips are RS rows of one image (halo re-read (RS+2)/RS, mostly L2 hits because neighbouring strips run concurrently); RS is
// chosen by the launcher so the grid fills whole waves of 148 SMs x resident CTAs.
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

constexpr int SEG = 3;          // output columns per thread
constexpr int NTHR = 256;

// Work item = (image b, strip, column segment); blockDim = (CQ channel quads, PY items).
// y[r][w] = sum_{i,j} x[r+i-1][w+j-1] * kk[i][j], kk = k (forward) or k rotated by 180 degrees (FLIP: backward-data).
// RED (backward-data only): the output is d(block output) of the block below, whose BatchNorm+ReLU6(+Dropout) backward starts with a
// reduction pass over exactly this tensor -- it is accumulated here while the values are still in registers (red[c] += dz,
// red[C+c] += dz*xhat, dz = out * dropmask * 1[0 <= bn(y) <= 6], y = that block's raw pointwise output), saving one read of out and a launch.
struct RowsRed { const float* y; const float* scale; const float* shift; const float* mean; const float* invstd;
                 float rate, inv_keep; uint64_t seed; uint32_t layer; const uint64_t* seed_ptr; };
template <bool FLIP, bool STATS, bool RED>
__global__ void __launch_bounds__(NTHR, 2) dwconv3x3_rows_kernel(const float* __restrict__ x, const float* __restrict__ k, float* __restrict__ y,
                                                                 int H, int W, int C4, int nseg, int nstrips, int RS, int nitems,
                                                                 double* __restrict__ stats, int rev, RowsRed rr, const BnFin fin)
{ pdl_enter();
    extern __shared__ __align__(16) unsigned char smraw[];
    float4* ks = reinterpret_cast<float4*>(smraw);              // [9][CQ] taps of this CTA's channel quads (+ [4][CQ] BN constants with RED)
    const int CQ = blockDim.x, PY = blockDim.y;
    const int c4 = blockIdx.x * CQ + threadIdx.x;
    const int C = C4 * 4;
    const bool cok = c4 < C4;
    if (cok)
        for (int q = threadIdx.y; q < 9; q += PY) ks[q * CQ + threadIdx.x] = ldg4(k + (size_t)(FLIP ? 8 - q : q) * C + c4 * 4);
    if (RED && cok && threadIdx.y == 0) {
        const float4 xa = ldg4(rr.invstd + c4 * 4), mu = ldg4(rr.mean + c4 * 4);
        ks[9 * CQ + threadIdx.x] = ldg4(rr.scale + c4 * 4);
        ks[10 * CQ + threadIdx.x] = ldg4(rr.shift + c4 * 4);
        ks[11 * CQ + threadIdx.x] = xa;                                                   // xhat = y*xa + xb
        ks[12 * CQ + threadIdx.x] = make_float4(-mu.x * xa.x, -mu.y * xa.y, -mu.z * xa.z, -mu.w * xa.w);
    }
    const uint64_t rseed = (RED && rr.seed_ptr) ? *rr.seed_ptr : rr.seed;
    __syncthreads();
    float s[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
    int item = blockIdx.y * PY + threadIdx.y;
    if (cok && item < nitems) {
        if (rev) item = nitems - 1 - item;
        const int seg = item % nseg; const int t1 = item / nseg;
        const int strip = t1 % nstrips; const int b = t1 / nstrips;
        const int w0 = seg * SEG;
        const int hs = strip * RS, he = min(H, hs + RS);
        const bool lok = w0 > 0, rok = w0 + SEG < W;                 // halo columns inside the image?
        const float4* kq = ks + threadIdx.x;
        // pointer to column w0-1 of virtual row hs-1
        const float* xrow = x + (((size_t)b * H + hs - 1) * W + (w0 - 1)) * C + c4 * 4;
        float* yrow = y + (((size_t)b * H + hs) * W + w0) * C + c4 * 4;
        const size_t rstride = (size_t)W * C;
        float4 acc[3][SEG];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int o = 0; o < SEG; ++o) acc[a][o] = zero4();
        float4 xb[2][SEG + 2];
        auto load_row = [&](float4 (&d)[SEG + 2], int r, const float* p) {
            const bool v = r >= 0 && r < H;
            d[0] = (v && lok) ? ldg4(p) : zero4();
#pragma unroll
            for (int t = 1; t <= SEG; ++t) d[t] = v ? ldg4(p + (size_t)t * C) : zero4();
            d[SEG + 1] = (v && rok) ? ldg4(p + (size_t)(SEG + 1) * C) : zero4();
        };
        const int nsteps = he - hs + 2;                              // virtual input rows hs-1 .. he
        load_row(xb[0], hs - 1, xrow);
        for (int base = 0; base < nsteps; base += 6) {
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int q = base + u;
                if (q < nsteps) {
                    const int r = hs - 1 + q;
                    if (q + 1 < nsteps) load_row(xb[(u + 1) & 1], r + 1, xrow + (size_t)(q + 1) * rstride);
                    float4 yv[SEG];                                  // RED: raw activations at the positions stored at the end of this step
                    if (RED && q >= 2) {
                        const float* yp = rr.y + (yrow - y) + (size_t)(q - 2) * rstride;
#pragma unroll
                        for (int o = 0; o < SEG; ++o) yv[o] = ldg4(yp + (size_t)o * C);
                    }
                    float4 (&xc)[SEG + 2] = xb[u & 1];
                    float4 (&A)[SEG] = acc[u % 3];               // output row r-1 (gets tap row 2)
                    float4 (&Bm)[SEG] = acc[(u + 1) % 3];        // output row r   (tap row 1)
                    float4 (&Cn)[SEG] = acc[(u + 2) % 3];        // output row r+1 (tap row 0)
                    if (r >= 0 && r < H) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const float4 k0 = kq[(0 * 3 + j) * CQ], k1 = kq[(1 * 3 + j) * CQ], k2 = kq[(2 * 3 + j) * CQ];
#pragma unroll
                            for (int o = 0; o < SEG; ++o) { fma4(A[o], xc[o + j], k2); fma4(Bm[o], xc[o + j], k1); fma4(Cn[o], xc[o + j], k0); }
                        }
                    }
                    if (q >= 2) {                                    // output row r-1 = hs + q - 2 is complete
                        float* dst = yrow + (size_t)(q - 2) * rstride;
#pragma unroll
                        for (int o = 0; o < SEG; ++o) {
                            *reinterpret_cast<float4*>(dst + (size_t)o * C) = A[o];
                            if (STATS) {
                                s[0] += A[o].x; s[1] += A[o].y; s[2] += A[o].z; s[3] += A[o].w;
                                sq[0] = fmaf(A[o].x, A[o].x, sq[0]); sq[1] = fmaf(A[o].y, A[o].y, sq[1]);
                                sq[2] = fmaf(A[o].z, A[o].z, sq[2]); sq[3] = fmaf(A[o].w, A[o].w, sq[3]);
                            }
                            if (RED) {
                                const float4 sc = kq[9 * CQ], sh = kq[10 * CQ], xa = kq[11 * CQ], xbn = kq[12 * CQ];
                                float d[4] = {A[o].x, A[o].y, A[o].z, A[o].w};
                                if (rr.rate > 0.f) {
                                    float dm[4];
                                    crnn_dropout_mask4(rseed, rr.layer, (uint64_t)(((dst - y) + (size_t)o * C) >> 2), rr.rate, rr.inv_keep, dm);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) d[e] *= dm[e];
                                }
                                const float yy[4] = {yv[o].x, yv[o].y, yv[o].z, yv[o].w};
                                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
                                const float xav[4] = {xa.x, xa.y, xa.z, xa.w}, xbv[4] = {xbn.x, xbn.y, xbn.z, xbn.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float z = fmaf(yy[e], scv[e], shv[e]);
                                    const float dz = (z >= 0.f && z <= 6.f) ? d[e] : 0.f;
                                    s[e] += dz; sq[e] = fmaf(dz, fmaf(yy[e], xav[e], xbv[e]), sq[e]);
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o < SEG; ++o) A[o] = zero4();    // becomes the "row r+2" accumulator of the next step
                }
            }
        }
    }
    if (!STATS && !RED) return;
    // BatchNorm statistics of the outputs (STATS) / reduction pass of the next BatchNorm backward (RED) (fp32 per thread over <= RS x 3 values, fp64 across threads): one atomic pair per channel and CTA
    __syncthreads();                                                 // taps no longer needed: reuse the buffer
    double* dsm = reinterpret_cast<double*>(smraw);                  // [PY][8][CQ]
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        dsm[(threadIdx.y * 8 + e) * CQ + threadIdx.x] = (double)s[e];
        dsm[(threadIdx.y * 8 + 4 + e) * CQ + threadIdx.x] = (double)sq[e];
    }
    __syncthreads();
    for (int e = threadIdx.y; e < 8; e += PY) {
        if (!cok) break;
        double t = 0.0;
        for (int i = 0; i < PY; ++i) t += dsm[(i * 8 + e) * CQ + threadIdx.x];
        atomicAdd(stats + (e >> 2) * C + c4 * 4 + (e & 3), t);
    }
    if (STATS) bn_finalize_tail(fin);
}

// Backward-weight: dk[i][j][c] = sum_{b,r,w} x[r+i-1][w+j-1][c] * dy[r][w][c].  Same marching scheme with 2 output columns per thread:
// input row r of x (4 float4 incl. halo) meets the three dy rows r+1, r, r-1 (tap rows 0, 1, 2) that live in a 4-slot register ring
// (the slot of row r-3 is refilled with row r+2 one step ahead); the nine tap accumulators (36 registers) stay put for the whole strip
// and are reduced across the CTA through shared memory, then one fp32 atomic per tap/channel and CTA.
constexpr int SEGW = 2;
__global__ void __launch_bounds__(NTHR, 2) dwconv3x3_rows_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dk,
                                                                            int H, int W, int C4, int nseg, int nstrips, int RS, int nitems)
{ pdl_enter();
    extern __shared__ __align__(16) unsigned char smraw[];
    float* red = reinterpret_cast<float*>(smraw);               // [PY][36][CQ]
    const int CQ = blockDim.x, PY = blockDim.y;
    const int c4 = blockIdx.x * CQ + threadIdx.x;
    const int C = C4 * 4;
    const bool cok = c4 < C4;
    float4 acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = zero4();
    const int item = blockIdx.y * PY + threadIdx.y;
    if (cok && item < nitems) {
        const int seg = item % nseg; const int t1 = item / nseg;
        const int strip = t1 % nstrips; const int b = t1 / nstrips;
        const int w0 = seg * SEGW;
        const int hs = strip * RS, he = min(H, hs + RS);
        bool xok[SEGW + 2], dok[SEGW];
#pragma unroll
        for (int t = 0; t < SEGW + 2; ++t) xok[t] = (w0 - 1 + t) >= 0 && (w0 - 1 + t) < W;
#pragma unroll
        for (int o = 0; o < SEGW; ++o) dok[o] = (w0 + o) < W;
        const size_t rstride = (size_t)W * C;
        const float* xrow = x + (((size_t)b * H + hs - 1) * W + (w0 - 1)) * C + c4 * 4;   // column w0-1 of virtual row hs-1
        const float* drow = dy + (((size_t)b * H + hs) * W + w0) * C + c4 * 4;            // column w0 of row hs
        float4 xb[2][SEGW + 2], dr[4][SEGW];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int o = 0; o < SEGW; ++o) dr[a][o] = zero4();
        auto load_x = [&](float4 (&d)[SEGW + 2], int r, const float* p) {
            const bool v = r >= 0 && r < H;
#pragma unroll
            for (int t = 0; t < SEGW + 2; ++t) d[t] = (v && xok[t]) ? ldg4(p + (size_t)t * C) : zero4();
        };
        auto load_dy = [&](float4 (&d)[SEGW], int rho, const float* p) {
            const bool v = rho >= hs && rho < he;
#pragma unroll
            for (int o = 0; o < SEGW; ++o) d[o] = (v && dok[o]) ? ldg4(p + (size_t)o * C) : zero4();
        };
        const int nsteps = he - hs + 2;                              // virtual x rows hs-1 .. he
        load_x(xb[0], hs - 1, xrow);
        load_dy(dr[0], hs, drow);
        for (int base = 0; base < nsteps; base += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = base + u;
                if (q < nsteps) {
                    const int r = hs - 1 + q;
                    if (q + 1 < nsteps) {
                        load_x(xb[(u + 1) & 1], r + 1, xrow + (size_t)(q + 1) * rstride);
                        load_dy(dr[(u + 1) & 3], hs + q + 1, drow + (size_t)(q + 1) * rstride);
                    }
                    if (r >= 0 && r < H) {
                        float4 (&xc)[SEGW + 2] = xb[u & 1];
                        float4 (&d0)[SEGW] = dr[u & 3];              // dy row r+1 -> tap row 0
                        float4 (&d1)[SEGW] = dr[(u + 3) & 3];        // dy row r   -> tap row 1
                        float4 (&d2)[SEGW] = dr[(u + 2) & 3];        // dy row r-1 -> tap row 2
#pragma unroll
                        for (int j = 0; j < 3; ++j)
#pragma unroll
                            for (int o = 0; o < SEGW; ++o) { fma4(acc[0 + j], xc[o + j], d0[o]); fma4(acc[3 + j], xc[o + j], d1[o]); fma4(acc[6 + j], xc[o + j], d2[o]); }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        red[((threadIdx.y * 36) + q * 4 + 0) * CQ + threadIdx.x] = acc[q].x; red[((threadIdx.y * 36) + q * 4 + 1) * CQ + threadIdx.x] = acc[q].y;
        red[((threadIdx.y * 36) + q * 4 + 2) * CQ + threadIdx.x] = acc[q].z; red[((threadIdx.y * 36) + q * 4 + 3) * CQ + threadIdx.x] = acc[q].w;
    }
    __syncthreads();
    if (cok)
        for (int e = threadIdx.y; e < 36; e += PY) {
            float sum = 0.f;
            for (int yy = 0; yy < PY; ++yy) sum += red[(yy * 36 + e) * CQ + threadIdx.x];
            atomicAdd(dk + (size_t)(e >> 2) * C + c4 * 4 + (e & 3), sum);
        }
}

bool g_rows_disabled = false, g_rows_env_read = false;
bool rows_enabled() {
    if (!g_rows_env_read) { const char* e = getenv("CRNN_DWCONV_V1"); g_rows_disabled = e && e[0] == '1'; g_rows_env_read = true; }
    return !g_rows_disabled;
}

// strips per image: RS = ceil(H / n) rows (at least 6), n chosen to maximise (fill of the last wave of 148 SMs x `occ` resident CTAs)
// x (useful rows per loaded row) x (balance of the short last strip); ties -> longer strips
void plan_rows(int B, int H, int W, int C4, int occ, int segw, dim3& grid, dim3& block, int& nseg, int& nstrips, int& RS, int& nitems) {
    const int CQ = C4 >= 32 ? 32 : (C4 >= 16 ? 16 : (C4 >= 8 ? 8 : (C4 >= 4 ? 4 : (C4 >= 2 ? 2 : 1))));
    const int PY = NTHR / CQ, gx = (C4 + CQ - 1) / CQ;
    nseg = (W + segw - 1) / segw;
    double best = -1.0; int bestd = 1;
    for (int n = 1; n <= H; ++n) {
        const int rs = (H + n - 1) / n;
        if (rs < 6 && n > 1) break;
        const int d = (H + rs - 1) / rs;                         // strips actually needed with this strip length
        const long long items = (long long)B * d * nseg;
        const long long ctas = (long long)gx * ((items + PY - 1) / PY);
        const long long cap = 148LL * occ;
        const double eff = (double)ctas / (double)(((ctas + cap - 1) / cap) * cap) * ((double)rs / (rs + 2)) * ((double)H / ((double)d * rs));
        if (eff > best + 1e-9) { best = eff; bestd = d; }
    }
    nstrips = bestd; RS = (H + bestd - 1) / bestd;
    nstrips = (H + RS - 1) / RS;
    nitems = B * nstrips * nseg;
    grid = dim3(gx, (unsigned)((nitems + PY - 1) / PY)); block = dim3(CQ, PY);
}

}  // namespace

// returns CRNN_OK when the row-marching kernel ran, 1 when the shape is not covered (caller falls back to the channel-block kernel)
int launch_dwconv_rows(const float* x, const float* k, float* y, int B, int H, int W, int C, int flip, double* stats, int rev, cudaStream_t st,
                       const DwRowsRed* red)
{
    if (!rows_enabled() || C % 4 || W % SEG || W < SEG || (flip && stats && !red) || (red && (!flip || !stats))) return 1;
    g_crnn_family = CRNN_FAM_DWROWS;
    dim3 grid, block; int nseg, nstrips, RS, nitems;
    plan_rows(B, H, W, C / 4, 2, SEG, grid, block, nseg, nstrips, RS, nitems);
    const size_t sm = std::max(sizeof(float4) * 13 * block.x, (stats ? sizeof(double) * 8 * NTHR : (size_t)0));
    RowsRed rr = {};
    BnFin fin = {}; if (stats && !flip && !red) fin = crnn_take_bn_fin();
    if (red) { rr.y = red->y; rr.scale = red->scale; rr.shift = red->shift; rr.mean = red->mean; rr.invstd = red->invstd; rr.rate = red->rate;
               rr.inv_keep = red->rate > 0.f ? 1.f / (1.f - red->rate) : 1.f; rr.seed = red->seed; rr.layer = red->layer; rr.seed_ptr = red->seed_ptr; }
    if (red) (void)crnn_launch(dwconv3x3_rows_kernel<true, false, true>, grid, block, sm, st, x, k, y, H, W, C / 4, nseg, nstrips, RS, nitems, stats, rev, rr, fin);
    else if (flip) (void)crnn_launch(dwconv3x3_rows_kernel<true, false, false>, grid, block, sm, st, x, k, y, H, W, C / 4, nseg, nstrips, RS, nitems, nullptr, rev, rr, fin);
    else if (stats) (void)crnn_launch(dwconv3x3_rows_kernel<false, true, false>, grid, block, sm, st, x, k, y, H, W, C / 4, nseg, nstrips, RS, nitems, stats, rev, rr, fin);
    else (void)crnn_launch(dwconv3x3_rows_kernel<false, false, false>, grid, block, sm, st, x, k, y, H, W, C / 4, nseg, nstrips, RS, nitems, nullptr, rev, rr, fin);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_dwconv_rows_bwd_weight(const float* x, const float* dy, float* dk, int B, int H, int W, int C, cudaStream_t st)
{
    if (!rows_enabled() || C % 4) return 1;
    g_crnn_family = CRNN_FAM_DWROWS;
    dim3 grid, block; int nseg, nstrips, RS, nitems;
    plan_rows(B, H, W, C / 4, 2, SEGW, grid, block, nseg, nstrips, RS, nitems);
    (void)crnn_launch(dwconv3x3_rows_bwd_weight_kernel, grid, block, sizeof(float) * 36 * NTHR, st, x, dy, dk, H, W, C / 4, nseg, nstrips, RS, nitems);
    LAUNCH_CHECK();
    return CRNN_OK;
}
