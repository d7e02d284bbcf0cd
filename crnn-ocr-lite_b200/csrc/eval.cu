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
{
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

int launch_edit_distance(const int32_t* a, const int32_t* alen, const int32_t* b, const int32_t* blen, int N, int maxlen, int32_t* out, cudaStream_t st)
{
    if (N <= 0) return CRNN_OK;
    if (maxlen < 1 || maxlen > ED_MAXLEN) { crnn_set_error("edit_distance: maxlen must be in 1..%d", ED_MAXLEN); return CRNN_ERR_INVALID; }
    edit_distance_kernel<<<ceil_div(N, 128), 128, 0, st>>>(a, alen, b, blen, N, maxlen, out);
    LAUNCH_CHECK();
    return CRNN_OK;
}
