// stn.cu -- spatial transformer of the reference (utils.py:105-258) for sm_100a.
//   trunk : MaxPool(2,2) -> Conv5x5 valid 1->20 (+bias, linear) -> MaxPool(2,2) -> Conv5x5 valid 20->20 (+bias) -> Flatten(HWC)
//           (utils.py:248-252); the two dense layers (utils.py:253-256) run on the shared GEMM.
//   sampler: BilinearInterpolation._transform/_interpolate (utils.py:140-232) with every quirk kept:
//           scale by size (not size-1), int cast truncates toward zero, corners clipped before the weights are
//           formed, four-term sum left-associated; explicit round-to-nearest fp32 ops (no FMA contraction) so the
//           sample coordinates are bit-identical to the oracle's.  Output is written straight into the
//           ZeroPadding2D((2,2)) buffer (utils.py:63).
#include "common.cuh"
#include "kernels.h"

StnDims stn_dims(int H, int W) {
    StnDims d;
    d.H = H; d.W = W;
    d.P1h = H / 2; d.P1w = W / 2;
    d.C1h = d.P1h - 4; d.C1w = d.P1w - 4;
    d.P2h = d.C1h / 2; d.P2w = d.C1w / 2;
    d.C2h = d.P2h - 4; d.C2w = d.P2w - 4;
    d.H1 = 0; d.W1 = 0;
    d.F = d.C2h * d.C2w * 20;
    return d;
}

namespace {
constexpr int NC = 20;   // locnet conv channels

// one CTA per image (1024 threads: every loop is a latency-bound dependent FMA chain per output, so warps are what hides it)
__global__ void __launch_bounds__(1024)
stn_trunk_fwd_kernel(const float* __restrict__ x, const float* __restrict__ k1, const float* __restrict__ b1,
                     const float* __restrict__ k2, const float* __restrict__ b2,
                     float* __restrict__ p1g, float* __restrict__ p2g, int* __restrict__ p2arg, float* __restrict__ flat, StnDims d)
{ pdl_enter();
    extern __shared__ float sm[];
    float* p1 = sm;                           // P1h*P1w
    float* sk1 = p1 + d.P1h * d.P1w;          // 25*20
    float* p2 = sk1 + 25 * NC;                // P2h*P2w*20
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* xb = x + (size_t)b * d.H * d.W;
    for (int i = tid; i < d.P1h * d.P1w; i += blockDim.x) {
        int h = i / d.P1w, w = i - h * d.P1w;
        const float* s = xb + (size_t)(2 * h) * d.W + 2 * w;
        float v = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[d.W], s[d.W + 1]));
        p1[i] = v; p1g[(size_t)b * d.P1h * d.P1w + i] = v;
    }
    for (int i = tid; i < 25 * NC; i += blockDim.x) sk1[i] = k1[i];
    __syncthreads();
    // conv1 + pool2 (argmax kept for backward)
    const int n2 = d.P2h * d.P2w * NC;
    for (int i = tid; i < n2; i += blockDim.x) {
        int co = i % NC; int r = i / NC; int w2 = r % d.P2w, h2 = r / d.P2w;
        float best = -INFINITY; int arg = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int h = 2 * h2 + (q >> 1), w = 2 * w2 + (q & 1);
            float acc = 0.f;
#pragma unroll
            for (int ii = 0; ii < 5; ++ii)
#pragma unroll
                for (int jj = 0; jj < 5; ++jj) acc = fmaf(p1[(h + ii) * d.P1w + w + jj], sk1[(ii * 5 + jj) * NC + co], acc);
            acc += b1[co];
            if (acc > best) { best = acc; arg = q; }
        }
        p2[i] = best;
        p2g[(size_t)b * n2 + i] = best; p2arg[(size_t)b * n2 + i] = arg;
    }
    __syncthreads();
    // conv2 -> flat (H,W,C order)
    for (int i = tid; i < d.F; i += blockDim.x) {
        int co = i % NC; int r = i / NC; int w = r % d.C2w, h = r / d.C2w;
        float acc = 0.f;
        for (int ii = 0; ii < 5; ++ii)
            for (int jj = 0; jj < 5; ++jj) {
                const float* pp = p2 + ((h + ii) * d.P2w + w + jj) * NC;
                const float* kk = k2 + ((ii * 5 + jj) * NC) * NC + co;
#pragma unroll
                for (int ci = 0; ci < NC; ++ci) acc = fmaf(pp[ci], __ldg(kk + ci * NC), acc);
            }
        flat[(size_t)b * d.F + i] = acc + b2[co];
    }
}

// backward of the trunk for one image: weight/bias grads (atomics into the shared grad buffers).  Two CTAs per image run the two
// independent halves concurrently (ncu r1m: one 1024-thread CTA per image took 133 us at the tail of the step's critical path):
//   blockIdx.y == 0: db2, dk2 (conv2 weight gradient: 10 000 outputs x 52 positions)
//   blockIdx.y == 1: dp2 (conv2 input gradient) -> through pool2's argmax -> db1, dk1 (two threads per dk1 output)
__global__ void __launch_bounds__(1024)
stn_trunk_bwd_kernel(const float* __restrict__ dflat, const float* __restrict__ p1g, const float* __restrict__ p2g,
                     const int* __restrict__ p2arg, const float* __restrict__ k2,
                     float* __restrict__ dk1, float* __restrict__ db1, float* __restrict__ dk2, float* __restrict__ db2, StnDims d)
{ pdl_enter();
    extern __shared__ float sm[];
    const int n2 = d.P2h * d.P2w * NC;
    float* p1 = sm;                       // P1h*P1w
    float* p2 = p1 + d.P1h * d.P1w;       // n2
    float* dc2 = p2 + n2;                 // F
    float* dp2 = dc2 + d.F;               // n2
    float* sk2 = dp2 + n2;                // 25*NC*NC conv2 kernel (read 500x per dp2 output)
    unsigned char* arg = reinterpret_cast<unsigned char*>(sk2 + 25 * NC * NC);   // n2 pool2 argmax codes (0..3)
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < d.F; i += blockDim.x) dc2[i] = dflat[(size_t)b * d.F + i];
    if (blockIdx.y == 0) {
        for (int i = tid; i < n2; i += blockDim.x) p2[i] = p2g[(size_t)b * n2 + i];
        __syncthreads();
        for (int co = tid; co < NC; co += blockDim.x) {
            float s = 0.f;
            for (int r = 0; r < d.C2h * d.C2w; ++r) s += dc2[r * NC + co];
            atomicAdd(db2 + co, s);
        }
        for (int i = tid; i < 25 * NC * NC; i += blockDim.x) {
            int co = i % NC; int r = i / NC; int ci = r % NC; int q = r / NC; int ii = q / 5, jj = q % 5;
            float s = 0.f;
            for (int h = 0; h < d.C2h; ++h)
                for (int w = 0; w < d.C2w; ++w) s = fmaf(p2[((h + ii) * d.P2w + w + jj) * NC + ci], dc2[(h * d.C2w + w) * NC + co], s);
            atomicAdd(dk2 + i, s);
        }
        return;
    }
    for (int i = tid; i < d.P1h * d.P1w; i += blockDim.x) p1[i] = p1g[(size_t)b * d.P1h * d.P1w + i];
    for (int i = tid; i < 25 * NC * NC; i += blockDim.x) sk2[i] = k2[i];
    for (int i = tid; i < n2; i += blockDim.x) arg[i] = (unsigned char)p2arg[(size_t)b * n2 + i];
    __syncthreads();
    // dp2[h'][w'][ci] = sum_{ii,jj,co} dc2[h'-ii][w'-jj][co] * k2[ii][jj][ci][co]
    for (int i = tid; i < n2; i += blockDim.x) {
        int ci = i % NC; int r = i / NC; int w = r % d.P2w, h = r / d.P2w;
        float s = 0.f;
        for (int ii = 0; ii < 5; ++ii) {
            int hh = h - ii; if (hh < 0 || hh >= d.C2h) continue;
            for (int jj = 0; jj < 5; ++jj) {
                int ww = w - jj; if (ww < 0 || ww >= d.C2w) continue;
                const float* g = dc2 + (hh * d.C2w + ww) * NC;
                const float* kk = sk2 + ((ii * 5 + jj) * NC + ci) * NC;
#pragma unroll
                for (int co = 0; co < NC; ++co) s = fmaf(g[co], kk[co], s);
            }
        }
        dp2[i] = s;
    }
    __syncthreads();
    // through pool2 (argmax) into conv1: db1, dk1
    for (int co = tid; co < NC; co += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < d.P2h * d.P2w; ++r) s += dp2[r * NC + co];
        atomicAdd(db1 + co, s);
    }
    const int npos = d.P2h * d.P2w, hpos = (npos + 1) / 2;
    for (int i2 = tid; i2 < 2 * 25 * NC; i2 += blockDim.x) {
        const int i = i2 % (25 * NC), part = i2 / (25 * NC);
        int co = i % NC; int q = i / NC; int ii = q / 5, jj = q % 5;
        float s = 0.f;
        const int r1 = min(npos, (part + 1) * hpos);
        for (int r = part * hpos; r < r1; ++r) {
            const int h2 = r / d.P2w, w2 = r - h2 * d.P2w;
            const int e = r * NC + co;
            const int a = arg[e];
            const int h = 2 * h2 + (a >> 1), w = 2 * w2 + (a & 1);
            s = fmaf(p1[(h + ii) * d.P1w + w + jj], dp2[e], s);
        }
        atomicAdd(dk1 + i, s);
    }
}

struct SamplePt { float xf, yf, wa, wb, wc, wd, x0f, x1f, y0f, y1f; int ia, ib, ic, id; float gx, gy; };

__device__ __forceinline__ SamplePt sample_point(const float* __restrict__ th, int i, int j, int H, int W)
{
    SamplePt s;
    const float stepx = __fdiv_rn(2.f, (float)(W - 1)), stepy = __fdiv_rn(2.f, (float)(H - 1));
    s.gx = __fadd_rn(-1.f, __fmul_rn(stepx, (float)j));       // tf.linspace(-1,1,W)[j] = start + step*j (fp32)
    s.gy = __fadd_rn(-1.f, __fmul_rn(stepy, (float)i));
    float xs = __fadd_rn(__fadd_rn(__fmul_rn(th[0], s.gx), __fmul_rn(th[1], s.gy)), th[2]);
    float ys = __fadd_rn(__fadd_rn(__fmul_rn(th[3], s.gx), __fmul_rn(th[4], s.gy)), th[5]);
    s.xf = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(xs, 1.f)), (float)W);   // utils.py:150  .5*(x+1)*width
    s.yf = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(ys, 1.f)), (float)H);   // utils.py:151
    int x0 = __float2int_rz(s.xf), y0 = __float2int_rz(s.yf);          // K.cast(x,'int32'): truncation (utils.py:153,155)
    x0 = max(min(x0, 1 << 30), -(1 << 30)); y0 = max(min(y0, 1 << 30), -(1 << 30));
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1);          // utils.py:161-164
    y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
    s.ia = y0 * W + x0; s.ib = y1 * W + x0; s.ic = y0 * W + x1; s.id = y1 * W + x1;   // utils.py:179-182
    s.x0f = (float)x0; s.x1f = (float)x1; s.y0f = (float)y0; s.y1f = (float)y1;
    s.wa = __fmul_rn(__fsub_rn(s.x1f, s.xf), __fsub_rn(s.y1f, s.yf));   // utils.py:196-199
    s.wb = __fmul_rn(__fsub_rn(s.x1f, s.xf), __fsub_rn(s.yf, s.y0f));
    s.wc = __fmul_rn(__fsub_rn(s.xf, s.x0f), __fsub_rn(s.y1f, s.yf));
    s.wd = __fmul_rn(__fsub_rn(s.xf, s.x0f), __fsub_rn(s.yf, s.y0f));
    return s;
}

__global__ void stn_sample_fwd_kernel(const float* __restrict__ x, const float* __restrict__ theta, float* __restrict__ out,
                                      int B, int H, int W, int pad, long long total)
{ pdl_enter();
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int jp = (int)(idx % Wp); long long r = idx / Wp;
        int ip = (int)(r % Hp); int b = (int)(r / Hp);
        int i = ip - pad, j = jp - pad;
        float v = 0.f;
        if (i >= 0 && i < H && j >= 0 && j < W) {
            SamplePt s = sample_point(theta + (size_t)b * 6, i, j, H, W);
            const float* xb = x + (size_t)b * H * W;
            float pa = xb[s.ia], pb = xb[s.ib], pc = xb[s.ic], pd = xb[s.id];
            v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.wa, pa), __fmul_rn(s.wb, pb)), __fmul_rn(s.wc, pc)), __fmul_rn(s.wd, pd));  // utils.py:205
        }
        out[idx] = v;
    }
}

// dtheta[b][0..5] = sum_pix dout * d out / d theta ; gradient flows through xf,yf inside the weights only
__global__ void stn_sample_bwd_kernel(const float* __restrict__ x, const float* __restrict__ theta, const float* __restrict__ dout,
                                      float* __restrict__ dtheta, int H, int W, int pad)
{ pdl_enter();
    const int b = blockIdx.y;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const float* xb = x + (size_t)b * H * W;
    float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
        int i = p / W, j = p - i * W;
        SamplePt s = sample_point(theta + (size_t)b * 6, i, j, H, W);
        float go = dout[((size_t)b * Hp + i + pad) * Wp + j + pad];
        float pa = xb[s.ia], pb = xb[s.ib], pc = xb[s.ic], pd = xb[s.id];
        float dx = -(s.y1f - s.yf) * pa - (s.yf - s.y0f) * pb + (s.y1f - s.yf) * pc + (s.yf - s.y0f) * pd;
        float dy = -(s.x1f - s.xf) * pa + (s.x1f - s.xf) * pb - (s.xf - s.x0f) * pc + (s.xf - s.x0f) * pd;
        float gxs = go * dx * (0.5f * (float)W), gys = go * dy * (0.5f * (float)H);
        g[0] += gxs * s.gx; g[1] += gxs * s.gy; g[2] += gxs;
        g[3] += gys * s.gx; g[4] += gys * s.gy; g[5] += gys;
    }
    __shared__ float red[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 6; ++q) { g[q] = warp_sum(g[q]); if (lane == 0) red[warp][q] = g[q]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
        atomicAdd(dtheta + (size_t)b * 6 + threadIdx.x, s);
    }
}
// ---- localisation head (utils.py:253-256): Dense(50) -> relu -> Dense(6).  One CTA per image, 512 threads = 8 k-slices x 64 unit
// lanes (50 active); the 6-wide second layer runs on the relu'd row while it is still in shared memory.
constexpr int ND1 = 50, NTH = 6;
__global__ void __launch_bounds__(512)
stn_head_fwd_kernel(const float* __restrict__ flat, const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                    const float* __restrict__ b2, float* __restrict__ loc_d1, float* __restrict__ theta, int F)
{ pdl_enter();
    extern __shared__ float hsm[];            // flat row [F] | d1 [64] | partial (double) [8][64]
    float* row = hsm; float* d1 = row + ((F + 1) & ~1); double* part = reinterpret_cast<double*>(d1 + 64);
    const int b = blockIdx.x, tid = threadIdx.x, n = tid & 63, ks = tid >> 6;
    for (int k = tid; k < F; k += 512) row[k] = __ldg(flat + (size_t)b * F + k);
    __syncthreads();
    // double accumulation (3.3 MFLOP per batch, free): theta steers every sample coordinate of the bilinear sampler, whose int
    // truncation makes the rest of the network discontinuous in it -- keep it at the rounding floor of fp32
    double acc = 0.0;
    if (n < ND1) {
        const int per = (F + 7) / 8, k0 = ks * per, k1 = min(F, k0 + per);
#pragma unroll 4
        for (int k = k0; k < k1; ++k) acc = fma((double)row[k], (double)__ldg(W1 + (size_t)k * ND1 + n), acc);
    }
    part[ks * 64 + n] = acc;
    __syncthreads();
    if (tid < 64) {
        float v = 0.f;
        if (tid < ND1) {
            double t = (double)__ldg(b1 + tid);
#pragma unroll
            for (int i = 0; i < 8; ++i) t += part[i * 64 + tid];
            v = fmaxf((float)t, 0.f);
            loc_d1[(size_t)b * ND1 + tid] = v;
        }
        d1[tid] = v;
    }
    __syncthreads();
    if (tid < NTH) {
        double v = (double)__ldg(b2 + tid);
        for (int i = 0; i < ND1; ++i) v = fma((double)d1[i], (double)__ldg(W2 + i * NTH + tid), v);
        theta[(size_t)b * NTH + tid] = (float)v;
    }
}
// data gradient of the head: dd1 = 1[loc_d1 > 0] * (dtheta @ W2^T); dflat = dd1 @ W1^T (one CTA per image)
__global__ void __launch_bounds__(256)
stn_head_bwd_kernel(const float* __restrict__ dtheta, const float* __restrict__ loc_d1, const float* __restrict__ W1, const float* __restrict__ W2,
                    float* __restrict__ dd1, float* __restrict__ dflat, int F)
{ pdl_enter();
    __shared__ float g[ND1];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < ND1) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < NTH; ++j) v = fmaf(__ldg(dtheta + (size_t)b * NTH + j), __ldg(W2 + tid * NTH + j), v);
        v = __ldg(loc_d1 + (size_t)b * ND1 + tid) > 0.f ? v : 0.f;
        g[tid] = v; dd1[(size_t)b * ND1 + tid] = v;
    }
    __syncthreads();
    for (int k = tid; k < F; k += 256) {
        const float* wr = W1 + (size_t)k * ND1;
        float v = 0.f;
#pragma unroll 10
        for (int i = 0; i < ND1; ++i) v = fmaf(g[i], __ldg(wr + i), v);
        dflat[(size_t)b * F + k] = v;
    }
}
}  // namespace

int launch_stn_trunk_fwd(const float* x, const float* k1, const float* b1, const float* k2, const float* b2,
                         float* p1, float* p2, int* p2arg, float* flat, int B, int H, int W, cudaStream_t st)
{
    StnDims d = stn_dims(H, W);
    if (d.C2h <= 0 || d.C2w <= 0) { crnn_set_error("stn: image %dx%d too small for the localisation net", H, W); return CRNN_ERR_INVALID; }
    size_t smem = sizeof(float) * ((size_t)d.P1h * d.P1w + 25 * NC + (size_t)d.P2h * d.P2w * NC);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) { CUDA_TRY(cudaFuncSetAttribute(stn_trunk_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = smem; }
    (void)crnn_launch(stn_trunk_fwd_kernel, B, 1024, smem, st, x, k1, b1, k2, b2, p1, p2, p2arg, flat, d);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_stn_trunk_bwd(const float* dflat, const float* p1, const float* p2, const int* p2arg, const float* k2,
                         float* dk1, float* db1, float* dk2, float* db2, float*, int B, int H, int W, cudaStream_t st)
{
    StnDims d = stn_dims(H, W);
    size_t n2 = (size_t)d.P2h * d.P2w * NC;
    size_t smem = sizeof(float) * ((size_t)d.P1h * d.P1w + 2 * n2 + d.F + 25 * NC * NC) + ((n2 + 15) / 16) * 16;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) { CUDA_TRY(cudaFuncSetAttribute(stn_trunk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); configured = smem; }
    (void)crnn_launch(stn_trunk_bwd_kernel, dim3(B, 2), 1024, smem, st, dflat, p1, p2, p2arg, k2, dk1, db1, dk2, db2, d);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_stn_sample_fwd(const float* x, const float* theta, float* out, int B, int H, int W, int pad, cudaStream_t st)
{
    long long total = (long long)B * (H + 2 * pad) * (W + 2 * pad);
    long long blocks = (total + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
    (void)crnn_launch(stn_sample_fwd_kernel, (int)blocks, 256, 0, st, x, theta, out, B, H, W, pad, total);
    LAUNCH_CHECK(); return CRNN_OK;
}
int launch_stn_sample_bwd(const float* x, const float* theta, const float* dout, float* dtheta, int B, int H, int W, int pad, cudaStream_t st)
{
    dim3 grid(ceil_div((long long)H * W, 256 * 4), B);
    (void)crnn_launch(stn_sample_bwd_kernel, grid, 256, 0, st, x, theta, dout, dtheta, H, W, pad);
    LAUNCH_CHECK(); return CRNN_OK;
}

int launch_stn_head_fwd(const float* flat, const float* W1, const float* b1, const float* W2, const float* b2, float* loc_d1, float* theta, int B, int F, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    const size_t smem = sizeof(float) * (((size_t)F + 1) / 2 * 2 + 64) + sizeof(double) * 8 * 64;
    if (smem > 48 * 1024) { crnn_set_error("stn_head: flatten size %d too large", F); return CRNN_ERR_INVALID; }
    (void)crnn_launch(stn_head_fwd_kernel, B, 512, smem, st, flat, W1, b1, W2, b2, loc_d1, theta, F);
    LAUNCH_CHECK();
    return CRNN_OK;
}
int launch_stn_head_bwd(const float* dtheta, const float* loc_d1, const float* W1, const float* W2, float* dd1, float* dflat, int B, int F, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    (void)crnn_launch(stn_head_bwd_kernel, B, 256, 0, st, dtheta, loc_d1, W1, W2, dd1, dflat, F);
    LAUNCH_CHECK();
    return CRNN_OK;
}
