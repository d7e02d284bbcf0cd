// eval.cu -- evaluation step of predict.py (reference utils.py:262-298: levenshtein / edit_distance / normalized_edit_distance;
// predict.py:183-191).  The reference fills an (n+1) x (m+1) numpy matrix per pair in Python (2.1 s per 8000 pairs); here one
// thread per (prediction, truth) pair runs the same recurrence with a rolling row in registers / local memory.  Integer arithmetic:
// the distances are bit-identical to the reference's; the means are formed on the host in the reference's summation order.
#include "common.cuh"
#include "kernels.h"

namespace {
constexpr int ED_MAXLEN = 128;

__global__ void edit_distance_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ alen, const int32_t* __restrict__ b,
                                     const int32_t* __restrict__ blen, int N, int maxlen, int32_t* __restrict__ out)
{ pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int n = min(max(alen[i], 0), maxlen), m = min(max(blen[i], 0), maxlen);
    const int32_t* sa = a + (size_t)i * maxlen;
    const int32_t* sb = b + (size_t)i * maxlen;
    int32_t row[ED_MAXLEN + 1], tb[ED_MAXLEN];
    for (int j = 0; j < m; ++j) { tb[j] = sb[j]; row[j] = j; }
    row[m] = m;
    for (int x = 1; x <= n; ++x) {
        const int32_t ca = sa[x - 1];
        int diag = row[0];               // matrix[x-1][y-1]
        row[0] = x;
        for (int y = 1; y <= m; ++y) {
            const int up = row[y];       // matrix[x-1][y]
            const int v = min(min(up + 1, row[y - 1] + 1), diag + (ca != tb[y - 1] ? 1 : 0));
            diag = up;
            row[y] = v;
        }
    }
    out[i] = row[m];
}
}  // namespace

// Input pipeline (utils.py:415-416 `norm`, called at utils.py:490): x = (float32(u8) - float32(mean)) / float32(std), IEEE fp32 ops in the
// reference's order -> bit-identical to numpy; lets the host upload the 8-bit line images (4x fewer H2D bytes) instead of float32.
namespace {
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long long n, float mean, float stdv)
{ pdl_enter();
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const uchar4 v = reinterpret_cast<const uchar4*>(in)[i];
        float4 o;
        o.x = __fdiv_rn(__fsub_rn((float)v.x, mean), stdv); o.y = __fdiv_rn(__fsub_rn((float)v.y, mean), stdv);
        o.z = __fdiv_rn(__fsub_rn((float)v.z, mean), stdv); o.w = __fdiv_rn(__fsub_rn((float)v.w, mean), stdv);
        reinterpret_cast<float4*>(out)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const long long i = (n4 << 2) + threadIdx.x; out[i] = __fdiv_rn(__fsub_rn((float)in[i], mean), stdv); }
}
}  // namespace

int launch_normalize_u8(const uint8_t* in, float* out, long long n, float mean, float stdv, cudaStream_t st)
{
    if (n <= 0) return CRNN_OK;
    if ((reinterpret_cast<uintptr_t>(in) & 3) || (reinterpret_cast<uintptr_t>(out) & 15)) { crnn_set_error("normalize_u8: in must be 4-byte, out 16-byte aligned"); return CRNN_ERR_INVALID; }
    long long blocks = ((n >> 2) + 255) / 256; if (blocks < 1) blocks = 1; if (blocks > 148 * 8) blocks = 148 * 8;
    (void)crnn_launch(normalize_u8_kernel, (int)blocks, 256, 0, st, in, out, n, mean, stdv);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_edit_distance(const int32_t* a, const int32_t* alen, const int32_t* b, const int32_t* blen, int N, int maxlen, int32_t* out, cudaStream_t st)
{
    if (N <= 0) return CRNN_OK;
    if (maxlen < 1 || maxlen > ED_MAXLEN) { crnn_set_error("edit_distance: maxlen must be in 1..%d", ED_MAXLEN); return CRNN_ERR_INVALID; }
    (void)crnn_launch(edit_distance_kernel, ceil_div(N, 128), 128, 0, st, a, alen, b, blen, N, maxlen, out);
    LAUNCH_CHECK();
    return CRNN_OK;
}
