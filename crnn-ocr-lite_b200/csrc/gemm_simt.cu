// gemm_simt.cu -- fp32 SIMT GEMM  C[M,N] (+)= op(A)[M,K] * op(B)[K,N]  for every dense contraction of the path
// (pointwise 1x1 convs utils.py:47, dense1 utils.py:74, RNN input projections utils.py:78-82, dense2 utils.py:85
// and all their backward products).  Exact-fp32 baseline of the tensor-core (tcgen05) pointwise kernel.
//
//   transA=0: A stored [M][K] (lda)      transA=1: A stored [K][M] (lda)   (dW = X^T dY)
//   transB=0: B stored [K][N] (ldb)      transB=1: B stored [N][K] (ldb)   (dX = dY W^T)
//   A-prologue (optional): a <- relu6(a*scale[ch]+shift[ch]), ch = column of the STORED A  (BN+ReLU6 on load)
//   epilogue: + bias[n], relu; or split-K with atomicAdd into a pre-zeroed C.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BK = 16;
constexpr int NT = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// tile of a matrix stored [X][K] (K contiguous): rows x0.., cols k0..k0+15 -> regs (R float4 per thread)
template <int BX>
__device__ __forceinline__ void fetch_kcontig(const float* __restrict__ P, int ld, int X, int x0, int k0, int kend,
                                              bool vec_ok, const float* __restrict__ sc, const float* __restrict__ sh,
                                              float (&v)[BX * 4 / NT][4])
{
#pragma unroll
    for (int r = 0; r < BX * 4 / NT; ++r) {
        int idx = threadIdx.x + r * NT;
        int x = idx >> 2, kq = idx & 3;
        int gx = x0 + x, gk = k0 + kq * 4;
        v[r][0] = v[r][1] = v[r][2] = v[r][3] = 0.f;
        if (gx < X && gk < kend) {
            const float* src = P + (size_t)gx * ld + gk;
            if (vec_ok && gk + 3 < kend) {
                float4 t = ld4(src); v[r][0] = t.x; v[r][1] = t.y; v[r][2] = t.z; v[r][3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (gk + i < kend) v[r][i] = src[i];
            }
            if (sc) {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (gk + i < kend) v[r][i] = relu6f(fmaf(v[r][i], __ldg(sc + gk + i), __ldg(sh + gk + i)));
            }
        }
    }
}
template <int BX, int LD>
__device__ __forceinline__ void store_kcontig(float (*S)[LD], const float (&v)[BX * 4 / NT][4])
{
#pragma unroll
    for (int r = 0; r < BX * 4 / NT; ++r) {
        int idx = threadIdx.x + r * NT;
        int x = idx >> 2, kq = idx & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) S[kq * 4 + i][x] = v[r][i];
    }
}
// tile of a matrix stored [K][X] (X contiguous): rows k0..k0+15, cols x0..
template <int BX>
__device__ __forceinline__ void fetch_xcontig(const float* __restrict__ P, int ld, int X, int x0, int k0, int kend,
                                              bool vec_ok, const float* __restrict__ sc, const float* __restrict__ sh,
                                              float (&v)[BX * 4 / NT][4])
{
    constexpr int QPR = BX / 4;  // float4 per tile row
#pragma unroll
    for (int r = 0; r < BX * 4 / NT; ++r) {
        int idx = threadIdx.x + r * NT;
        int k = idx / QPR, xq = idx % QPR;
        int gk = k0 + k, gx = x0 + xq * 4;
        v[r][0] = v[r][1] = v[r][2] = v[r][3] = 0.f;
        if (gk < kend && gx < X) {
            const float* src = P + (size_t)gk * ld + gx;
            if (vec_ok && gx + 3 < X) {
                float4 t = ld4(src); v[r][0] = t.x; v[r][1] = t.y; v[r][2] = t.z; v[r][3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (gx + i < X) v[r][i] = src[i];
            }
            if (sc) {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (gx + i < X) v[r][i] = relu6f(fmaf(v[r][i], __ldg(sc + gx + i), __ldg(sh + gx + i)));
            }
        }
    }
}
template <int BX, int LD>
__device__ __forceinline__ void store_xcontig(float (*S)[LD], const float (&v)[BX * 4 / NT][4])
{
    constexpr int QPR = BX / 4;
#pragma unroll
    for (int r = 0; r < BX * 4 / NT; ++r) {
        int idx = threadIdx.x + r * NT;
        int k = idx / QPR, xq = idx % QPR;
        *reinterpret_cast<float4*>(&S[k][xq * 4]) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    }
}

template <int BM, int BN, int CM, int CN>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmArgs p)
{ pdl_enter();
    constexpr int LDA = BM + 4, LDB = BN + 4;
    constexpr int TX = BN / (4 * CN);
    static_assert(TX * (BM / (4 * CM)) == NT, "thread grid");
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];

    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // split-K range (multiples of BK)
    const int ktiles = (p.K + BK - 1) / BK;
    const int tiles_per = (ktiles + gridDim.z - 1) / gridDim.z;
    const int kbeg = blockIdx.z * tiles_per * BK;
    int kend = kbeg + tiles_per * BK; if (kend > p.K) kend = p.K;
    if (kbeg >= kend) return;
    const bool vecA = (p.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const bool vecB = (p.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0);

    float ra[BM * 4 / NT][4], rb[BN * 4 / NT][4];
    float acc[CM * 4][CN * 4];
#pragma unroll
    for (int i = 0; i < CM * 4; ++i)
#pragma unroll
        for (int j = 0; j < CN * 4; ++j) acc[i][j] = 0.f;

    auto fetch = [&](int k0) {
        if (p.transA) fetch_xcontig<BM>(p.A, p.lda, p.M, m0, k0, kend, vecA, p.a_scale, p.a_shift, ra);
        else          fetch_kcontig<BM>(p.A, p.lda, p.M, m0, k0, kend, vecA, p.a_scale, p.a_shift, ra);
        if (p.transB) fetch_kcontig<BN>(p.B, p.ldb, p.N, n0, k0, kend, vecB, nullptr, nullptr, rb);
        else          fetch_xcontig<BN>(p.B, p.ldb, p.N, n0, k0, kend, vecB, nullptr, nullptr, rb);
    };
    auto stash = [&](int buf) {
        if (p.transA) store_xcontig<BM, LDA>(As[buf], ra); else store_kcontig<BM, LDA>(As[buf], ra);
        if (p.transB) store_kcontig<BN, LDB>(Bs[buf], rb); else store_xcontig<BN, LDB>(Bs[buf], rb);
    };

    fetch(kbeg);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[CM * 4], b[CN * 4];
#pragma unroll
            for (int c = 0; c < CM; ++c) {
                float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][c * (BM / CM) + ty * 4]);
                a[c * 4 + 0] = t.x; a[c * 4 + 1] = t.y; a[c * 4 + 2] = t.z; a[c * 4 + 3] = t.w;
            }
#pragma unroll
            for (int c = 0; c < CN; ++c) {
                float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][c * (BN / CN) + tx * 4]);
                b[c * 4 + 0] = t.x; b[c * 4 + 1] = t.y; b[c * 4 + 2] = t.z; b[c * 4 + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < CM * 4; ++i)
#pragma unroll
                for (int j = 0; j < CN * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) { stash(buf ^ 1); __syncthreads(); buf ^= 1; }
    }

    // epilogue
    const bool vecC = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
    const bool atomic = gridDim.z > 1;
#pragma unroll
    for (int cm = 0; cm < CM; ++cm)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + cm * (BM / CM) + ty * 4 + i;
            if (m >= p.M) continue;
#pragma unroll
            for (int cn = 0; cn < CN; ++cn) {
                int n = n0 + cn * (BN / CN) + tx * 4;
                if (n >= p.N) continue;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = acc[cm * 4 + i][cn * 4 + j];
                float* dst = p.C + (size_t)m * p.ldc + n;
                if (atomic) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (n + j < p.N) atomicAdd(dst + j, v[j]);
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (p.bias && n + j < p.N) v[j] += __ldg(p.bias + n + j);
                    if (p.relu) v[j] = fmaxf(v[j], 0.f);
                    if (p.accumulate && n + j < p.N) v[j] += dst[j];
                }
                if (vecC && n + 3 < p.N) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (n + j < p.N) dst[j] = v[j];
                }
            }
        }
}

}  // namespace

int launch_gemm_simt(const GemmArgs& a, cudaStream_t st)
{
    if (a.M <= 0 || a.N <= 0) return CRNN_OK;
    if (a.K <= 0) { crnn_set_error("gemm: K<=0"); return CRNN_ERR_INVALID; }
    GemmArgs p = a;
    int split = a.split_k < 1 ? 1 : a.split_k;
    if (split > 1 && (a.bias || a.relu || a.accumulate)) { crnn_set_error("gemm: split-K with epilogue"); return CRNN_ERR_INVALID; }
    const int ktiles = ceil_div(a.K, BK);
    if (split > ktiles) split = ktiles;
    auto run = [&](auto kern, int BM, int BN) {
        dim3 grid(ceil_div(a.N, BN), ceil_div(a.M, BM), split);
        (void)crnn_launch(kern, grid, NT, 0, st, p);
    };
    const long long big_ctas = (long long)ceil_div(a.M, 128) * ceil_div(a.N, 128) * split;
    if (a.N > 64 && big_ctas >= 148) run(gemm_simt_kernel<128, 128, 2, 2>, 128, 128);
    else if (a.N <= 64 && (long long)ceil_div(a.M, 128) * split >= 148) run(gemm_simt_kernel<128, 64, 2, 1>, 128, 64);
    else run(gemm_simt_kernel<64, 64, 1, 1>, 64, 64);
    LAUNCH_CHECK();
    return CRNN_OK;
}
