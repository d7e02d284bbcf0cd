// kernels.h -- internal launcher prototypes (host side) for the CRNN hot-path kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct GemmArgs {
    const float* A = nullptr; const float* B = nullptr; float* C = nullptr;
    int M = 0, N = 0, K = 0;
    int lda = 0, ldb = 0, ldc = 0;
    int transA = 0, transB = 0;
    const float* a_scale = nullptr; const float* a_shift = nullptr;  // A <- relu6(A*scale[ch]+shift[ch])
    const float* bias = nullptr;
    int relu = 0;
    int accumulate = 0;   // C += result (non split-K)
    int split_k = 1;      // >1: atomicAdd into pre-zeroed C
};
int launch_gemm_simt(const GemmArgs& a, cudaStream_t st);

// ---- gemm_tc.cu : tcgen05 / TMEM 3xTF32 kernel, out[m][n] = sum_k f(X[m][k]) * Wop[n][k] ----
size_t tc_weight_image_floats(int N, int K);
int launch_prep_weight_images(const float* W, int ldw, int N, int K, int transposed, float* img, cudaStream_t st);
// the same for up to MAXJ weight matrices in one launch (all images of a training step)
struct TcPrepBatch { static constexpr int MAXJ = 28; const float* W[MAXJ]; float* img[MAXJ]; int ldw[MAXJ], N[MAXJ], K[MAXJ], transposed[MAXJ]; long long start[MAXJ + 1]; int n = 0; };
int tc_prep_batch_add(TcPrepBatch& t, const float* W, int ldw, int N, int K, int transposed, float* img);
int launch_prep_weight_images_batch(const TcPrepBatch& t, cudaStream_t st);
// fused reduction pass of the BatchNorm+ReLU6 backward consuming the GEMM output (see TcArgs::red_y); y has the output's shape / stride
struct TcBnRed { const float* y; const float* scale; const float* shift; const float* mean; const float* invstd; };
int launch_xw_gemm_tc(const float* X, int ldx, const float* Wimg, float* out, int ldo, int M, int N, int K,
                      const float* x_scale, const float* x_shift, double* stats, cudaStream_t st,
                      const float* bias = nullptr, int relu = 0, int accumulate = 0, int rev = 0, int ksplit = 1, long long split_stride = 0,
                      const TcBnRed* red = nullptr);
// split-K for GEMMs with too few 128 x 128 output tiles to fill the SMs: slices the caller should request (1 = none); every slice
// writes its own copy of the output (split_stride floats apart), summed in a fixed order by launch_sum_partials
int xw_gemm_tc_pick_ksplit(int M, int N, int K);
// out[m*ldo + n] = mask(m*N + n) * act(sum_{s<S} part[s*stride + m*N + n] + bias[n]);  act = ReLU if relu; mask = inverted dropout if rate > 0
int launch_sum_partials(const float* part, int S, long long stride, long long M, int N, const float* bias, int relu, float* out, int ldo,
                        float rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr);

// dW[ci][co] += sum_m f(X[m][ci]) * dY[m][co]  (tcgen05, MN-major operands, split-K over pixels, atomics into pre-zeroed dW)
int launch_xty_gemm_tc(const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout, float* dW, int ldw, int M,
                       const float* x_scale, const float* x_shift, cudaStream_t st);

// ---- ctc.cu ----
int launch_ctc_loss_grad(const float* probs, int t_off, const int* labels, int maxL, const int* label_len,
                         const int* input_len, int B, int T, int V, float eps, float* loss, float* grad_u,
                         float* grad_logits, float scale, int* status, cudaStream_t st);
int launch_ctc_greedy(const float* probs, const int* seq_len, int B, int T, int V, float eps,
                      int* out, int* out_len, float* score, cudaStream_t st);
// top_paths P: out (B,P,T), out_len / logprob (B,P), best path first
int launch_ctc_beam(const float* probs, const int* seq_len, int B, int T, int V, float eps, int W, int merge_repeated,
                    int* out, int* out_len, float* logprob, cudaStream_t st, int top_paths = 1);

// ---- conv.cu (depthwise conv, BN statistics, activation/pool, elementwise) ----
// rev (here and below): walk the tensor from its end (serpentine traversal: a kernel starts where its producer finished, in L2)
int launch_dwconv_fwd(const float* x, const float* k33c, float* y, int B, int H, int W, int C, cudaStream_t st, double* stats = nullptr, int rev = 0);
// `red` (backward-data only, needs `stats` = the reduction buffer): fused reduction pass of the BatchNorm+ReLU6(+Dropout) backward that
// consumes the output (see dwconv_rows.cu); y = raw pre-BN activation with the output's shape
struct DwRowsRed { const float* y; const float* scale; const float* shift; const float* mean; const float* invstd;
                   float rate; uint64_t seed; uint32_t layer; const uint64_t* seed_ptr; };
// red_done (optional out): set to 1 when the fused reduction described by `red` / `red_buf` was performed by the kernel that ran
int launch_dwconv_bwd_data(const float* dy, const float* k33c, float* dx, int B, int H, int W, int C, int accumulate, cudaStream_t st, int rev = 0,
                           const DwRowsRed* red = nullptr, double* red_buf = nullptr, int* red_done = nullptr);
// dwconv_rows.cu: row-marching variants; return 1 (not an error) when the shape is not covered and the caller must fall back
int launch_dwconv_rows_bwd_weight(const float* x, const float* dy, float* dk33c, int B, int H, int W, int C, cudaStream_t st);
int launch_dwconv_rows(const float* x, const float* k33c, float* y, int B, int H, int W, int C, int flip, double* stats, int rev, cudaStream_t st,
                       const DwRowsRed* red = nullptr);
int launch_dwconv_bwd_weight(const float* x, const float* dy, float* dk33c, int B, int H, int W, int C, cudaStream_t st);
// dwconv_fused.cu: ReLU6+BN backward apply (needs the finished reduction `red1` = double[2C]) + depthwise backward-data + backward-weight
// in one pass; with `red` the third operand is the RAW pointwise output of the block below (the block input is recomputed from it) and that
// block's BN-backward reduction is accumulated into red_buf.
int dwconv_bwd_fused_covers(int H, int W, int C);
int launch_dwconv_bwd_fused(const float* dA, const float* z, const float* x_or_y, const float* k33c, float* dx, float* dk33c,
                            const float* scale, const float* shift, const float* mean, const float* invstd, const float* gamma, const double* red1,
                            int B, int H, int W, int C, int rev, const DwRowsRed* red, double* red_buf, cudaStream_t st);
// dwconv_fused.cu: out = depthwise3x3(dropout(relu6(y*pscale+pshift))) straight from the RAW pointwise output y of a non-pooled block (its
// activation tensor is never materialised) + BatchNorm statistics of out (optional).  Same shape coverage as the backward kernel.
int launch_dwconv_fwd_fused(const float* y, const float* pscale, const float* pshift, float rate, uint64_t seed, uint32_t layer, const uint64_t* seed_ptr,
                            const float* k33c, float* out, double* stats, int B, int H, int W, int C, int rev, cudaStream_t st);
int launch_bn_param_grads(const double* red, float* dgamma, float* dbeta, int C, cudaStream_t st);
// per-channel sum / sum of squares over rows of y[M][C] -> stats[0..C) , stats[C..2C) (double, pre-zeroed)
int launch_colstats(const float* y, long long M, int C, double* stats, cudaStream_t st);
// training: batch stats -> scale/shift (+ saved mean / inv-std, moving-stat update); inference: moving stats
int launch_bn_finalize(const double* stats, long long M, int C, const float* gamma, const float* beta,
                       float* moving_mean, float* moving_var, float eps, float momentum, int training,
                       float* scale, float* shift, float* save_mean, float* save_invstd, cudaStream_t st);
// a = dropout(pool(relu6(y*scale+shift)))  ; pool (ph,pw) in {(1,1),(2,2),(1,2)}
int launch_act_pool_fwd(const float* y, const float* scale, const float* shift, float* a, int B, int H, int W, int C,
                        int ph, int pw, float drop_rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr = nullptr, int rev = 0);
// fused (ReLU6 + MaxPool + Dropout) backward + BatchNorm-train backward, two passes over (da, y), no dz round trip:
//   dz = unpool(da*dropmask) * 1[0<=z<=6];  dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat));  dgamma += sum(dz*xhat), dbeta += sum(dz)
int launch_act_pool_bn_bwd(const float* da, const float* y, const float* scale, const float* shift, const float* mean, const float* invstd,
                           const float* gamma, float* dy, double* red /*pre-zeroed [2C]*/, float* dgamma, float* dbeta,
                           int B, int H, int W, int C, int ph, int pw, float drop_rate, uint64_t seed, uint32_t layer, cudaStream_t st,
                           const uint64_t* seed_ptr = nullptr, int rev = 0, int emit_param_grads = 1,
                           int reduce_done = 0);   // reduce pass walks `rev`, apply pass the opposite way; reduce_done: `red` already accumulated by the producer of da
// same for the BN after the depthwise conv (no pool / dropout); dy may alias da
// reduce_done != 0: `red` was already accumulated by the producer of `da` (fused into the dX GEMM epilogue), only the apply pass runs
int launch_relu6_bn_bwd(const float* da, const float* y, const float* scale, const float* shift, const float* mean, const float* invstd,
                        const float* gamma, float* dy, double* red, float* dgamma, float* dbeta, long long M, int C, cudaStream_t st, int rev = 0,
                        int reduce_done = 0, int emit_param_grads = 1);
// dgamma / dbeta of up to 16 BatchNorm layers from their reduction buffers in ONE launch (instead of one tiny launch per layer on the critical path)
struct BnGradTable { const double* red[16]; float* dgamma[16]; float* dbeta[16]; int C[16]; int n; };
int launch_bn_param_grads_all(const BnGradTable& t, cudaStream_t st);
// BN training backward: dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)) in place; dgamma = sum(dz*xhat), dbeta = sum(dz)
int launch_bn_bwd_apply(float* dz_inout, const float* y, const double* red, const float* gamma, const float* save_mean,
                        const float* save_invstd, float* dgamma, float* dbeta, long long M, int C, cudaStream_t st);
// misc elementwise
int launch_colsum(const float* y, long long M, int C, int ldy, float* out, cudaStream_t st);            // out[c] = sum_m y[m][c]
int launch_relu_dropout_bwd(float* g_inout, const float* act, long long n, float drop_rate, uint64_t seed, uint32_t layer, cudaStream_t st);
// seed_ptr (optional): device location of the step's dropout seed (CUDA-graph replay); overrides `seed` when non-null
int launch_dropout_fwd(float* x_inout, long long n, float drop_rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr = nullptr);
int launch_dropout_copy(const float* in, float* out, long long n, float drop_rate, uint64_t seed, uint32_t layer, cudaStream_t st, const uint64_t* seed_ptr = nullptr);
int launch_set_u64(uint64_t* p, uint64_t v, cudaStream_t st);
// block 1 of the conv stack (Cin = 1): the pointwise conv is an outer product out[m][co] = f(x[m]) * w[co], f = relu6(x*scale+shift).
// stats (optional, pre-zeroed double[2*Cout]) receives the per-channel sum / sum of squares of out.
int launch_pw1_fwd(const float* x, const float* scale, const float* shift, const float* w, float* out, long long M, int Cout, double* stats, cudaStream_t st, int rev = 0);
// its backward in one pass over dY: dX[m] = sum_co dY[m][co] w[co];  dW[co] += sum_m f(x[m]) dY[m][co]
int launch_pw1_bwd(const float* x, const float* scale, const float* shift, const float* dY, const float* w, float* dX, float* dW, long long M, int Cout, cudaStream_t st, int rev = 0);
int launch_sum_dirs(const float* hs, float* out, long long rows, int U, cudaStream_t st);               // out[r][u] = hs[r][0][u]+hs[r][1][u]
int launch_dup_dirs(const float* g, float* out, long long rows, int U, cudaStream_t st);                // out[r][d][u] = g[r][u]
int launch_softmax_rows(const float* z, float* p, long long rows, int V, cudaStream_t st);

// ---- stn.cu ----
struct StnDims { int H, W, H1, W1, P1h, P1w, C1h, C1w, P2h, P2w, C2h, C2w, F; };
StnDims stn_dims(int H, int W);
// locnet conv trunk: x (B,H,W) -> p1 (B,P1h,P1w) , p2 (B,P2h,P2w,20) [+argmax idx], flat (B,F)
int launch_stn_trunk_fwd(const float* x, const float* k1, const float* b1, const float* k2, const float* b2,
                         float* p1, float* p2, int* p2arg, float* flat, int B, int H, int W, cudaStream_t st);
int launch_stn_trunk_bwd(const float* dflat, const float* p1, const float* p2, const int* p2arg, const float* k2,
                         float* dk1, float* db1, float* dk2, float* db2, float* scratch_dc1, int B, int H, int W, cudaStream_t st);
// localisation head: loc_d1 = relu(flat @ W1 + b1) (B,50); theta = loc_d1 @ W2 + b2 (B,6)      (utils.py:253-256)
int launch_stn_head_fwd(const float* flat, const float* W1, const float* b1, const float* W2, const float* b2, float* loc_d1, float* theta, int B, int F, cudaStream_t st);
// its data gradient: dd1 = 1[loc_d1 > 0] * (dtheta @ W2^T) (B,50); dflat = dd1 @ W1^T (B,F)
int launch_stn_head_bwd(const float* dtheta, const float* loc_d1, const float* W1, const float* W2, float* dd1, float* dflat, int B, int F, cudaStream_t st);
// sampler: x (B,H,W), theta (B,6) -> padded out (B,H+2*pad,W+2*pad) (border zeroed)
int launch_stn_sample_fwd(const float* x, const float* theta, float* out, int B, int H, int W, int pad, cudaStream_t st);
int launch_stn_sample_bwd(const float* x, const float* theta, const float* dout_padded, float* dtheta, int B, int H, int W, int pad, cudaStream_t st);

// ---- rnn_mma.cu : cluster-resident GRU / LSTM recurrence with U held in registers as mma.sync fragments, state via DSMEM ----
// xp (B,T,2,G*U) input projections (+bias), U0/U1 (U,G*U) recurrent kernels of the two directions; hs (B,T,2,U) outputs; gates (B,T,2,GS*U) saved
// for BPTT (may be null).  Backward: dout (B,T,2,U) gradient wrt hs -> dxp (B,T,2,G*U), hprev (B,T,2,U), rh (B,T,2,U) (GRU only)
int launch_gru_fwd_mma(const float* xp, const float* U0, const float* U1, float* hs, float* gates, int B, int T, cudaStream_t st);
int launch_gru_bwd_mma(const float* dout, const float* hs, const float* gates, const float* U0, const float* U1,
                       float* dxp, float* hprev, float* rh, int B, int T, cudaStream_t st);
int launch_lstm_fwd_mma(const float* xp, const float* U0, const float* U1, float* hs, float* gates, int B, int T, cudaStream_t st);
int launch_lstm_bwd_mma(const float* dout, const float* hs, const float* gates, const float* U0, const float* U1,
                        float* dxp, float* hprev, int B, int T, cudaStream_t st);
// ---- optim.cu ----
int launch_sumsq(const float* g, long long n, double* out /*pre-zeroed*/, cudaStream_t st);
int launch_adam(float* w, const float* g, float* m, float* v, long long n, const double* sumsq, float clipnorm,
                float lr_t, float b1, float b2, float eps, float gscale, cudaStream_t st);
int launch_sgd_nesterov(float* w, const float* g, float* vel, long long n, const double* sumsq, float clipnorm,
                        float lr_i, float momentum, float gscale, cudaStream_t st);

// ---- eval.cu : batched Levenshtein distance (utils.py:262-298), one thread per pair, sequences padded to maxlen <= 128 ----
int launch_edit_distance(const int32_t* a, const int32_t* alen, const int32_t* b, const int32_t* blen, int N, int maxlen, int32_t* out, cudaStream_t st);
// input normalisation of 8-bit line images on the device (utils.py:415-416), bit-identical to numpy's float32 arithmetic
int launch_normalize_u8(const uint8_t* in, float* out, long long n, float mean, float stdv, cudaStream_t st);
