/*
 * crnn_b200.h -- C ABI of the B200-native CRNN-OCR hot path (libcrnn_b200.so).
 *
 * The reference (gasparian/CRNN-OCR-lite) has NO FFI / plugin interface for this path: its hot path is the
 * Python surface of utils.py driving Keras 2.2.2 / TensorFlow 1.8.  Each entry point below
 * therefore cites the reference *call site* it replaces; the Python mirror of that surface lives in
 * crnn-ocr-lite_b200/ (utils-compatible names) and binds these symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 (CRNN_OK) or a negative status; crnn_last_error() gives the message
 *     (thread-local); nothing throws across the ABI.
 *   - device pointers are BORROWED: the caller (PyTorch is only the allocator / stream provider) owns all
 *     memory, including the handle's workspace.  `stream` is a cudaStream_t passed as void*; every call is
 *     asynchronous on it, no hidden synchronisation, except the *_host convenience entry points which copy
 *     host<->device and synchronise the stream before returning.
 *   - one handle per GPU; a handle is not thread-safe, different handles are independent.
 *   - tensors are fp32, NHWC; axis 1 of the image is the text-line width (time), as in the reference
 *     (utils.py:370 flips+transposes the image).  Class V-1 is the CTC blank.
 */
#ifndef CRNN_B200_H
#define CRNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRNN_OK 0
#define CRNN_ERR_INVALID (-1)
#define CRNN_ERR_CUDA (-2)
#define CRNN_ERR_NOMEM (-3)
#define CRNN_ERR_UNKNOWN_NAME (-4)
#define CRNN_ERR_INFEASIBLE (-5)   /* "Not enough time for target transition sequence" (TF CTCLoss) */

#define CRNN_CELL_GRU 0            /* what the reference CLI actually builds */
#define CRNN_CELL_LSTM 1           /* utils.py:78-79 */

typedef struct crnn_handle crnn_handle;

/* CRNN(num_classes, max_string_len, shape, time_dense_size, GRU, n_units) -- utils.py:34-41 */
typedef struct {
    int32_t imgh;          /* shape[0]: text-line width / time axis (train.py:105, default 100) */
    int32_t imgw;          /* shape[1]: line height (train.py:106, 32) */
    int32_t num_classes;   /* len(lexicon)+1 (train.py:165) */
    int32_t cell;          /* CRNN_CELL_* */
    int32_t n_units;       /* 256 */
    int32_t time_dense;    /* 128 */
    int32_t max_len;       /* max_string_len (23) */
    int32_t max_batch;     /* largest batch the workspace is sized for */
} crnn_config;

typedef struct {
    int64_t offset;        /* in bytes from the workspace base */
    int64_t numel;
    int32_t is_int;        /* 1: int32 tensor, 0: float32 */
} crnn_tensor_info;

const char* crnn_last_error(void);
const char* crnn_version(void);

/* ---------------------------------------------------------------- model handle (CRNN.get_model, utils.py:58-96) */
int crnn_workspace_bytes(const crnn_config* cfg, size_t* bytes);
/* `workspace` = device memory of >= crnn_workspace_bytes, 256-byte aligned, zero-filled by the caller */
int crnn_create(const crnn_config* cfg, void* workspace, size_t workspace_bytes, crnn_handle** out);
int crnn_destroy(crnn_handle* h);
/* named tensors inside the workspace: weights ("conv2d_3/kernel", ... the names of models/<name>/final_weights.h5,
 * utils.py:328 load_weights), their gradients ("grad/<name>"), Adam state ("adam_m/<name>", "adam_v/<name>"),
 * activations ("act/<name>").  Flat arenas: "arena/params", "arena/grads", "arena/opt_m", "arena/opt_v". */
int crnn_num_tensors(const crnn_handle* h);
const char* crnn_tensor_name(const crnn_handle* h, int index);
int crnn_tensor_lookup(const crnn_handle* h, const char* name, crnn_tensor_info* out);

/* predictor = Model(the_input -> softmax).predict (utils.py:308-312, predict.py:166): x (B,imgh,imgw,1) ->
 * softmax (B,T,V).  BN uses moving statistics, no dropout. */
int crnn_forward(crnn_handle* h, const float* x_dev, int B, float* softmax_dev, void* stream);
/* same with host buffers (H2D + forward + D2H, synchronises) */
int crnn_forward_host(crnn_handle* h, const float* x_host, int B, float* softmax_host, void* stream);

/* one training forward/backward: model.train_on_batch up to the gradients (train.py:187-209):
 * STN+conv+BiRNN forward in training mode, ctc_lambda_func (utils.py:98-103), full backward.  Gradients of the
 * MEAN-over-batch loss land in "arena/grads"; per-sample losses in loss_dev (B).  dropout_seed==0 disables
 * dropout (parity mode); otherwise masks are a stateless hash of (seed, layer, element). */
int crnn_train_fwd_bwd(crnn_handle* h, const float* x_dev, const int32_t* labels_dev, const int32_t* label_len_dev,
                       const int32_t* input_len_dev, int B, float* loss_dev, uint64_t dropout_seed, void* stream);
/* optimizers.Adam(lr, beta_1=.5, beta_2=.999, clipnorm=5) / SGD(nesterov) step over the arenas (train.py:188-190);
 * grad_scale folds 1/world_size after a sum all-reduce of "arena/grads". */
int crnn_adam_step(crnn_handle* h, float lr, float beta1, float beta2, float eps, float clipnorm, float grad_scale, void* stream);
int crnn_sgd_step(crnn_handle* h, float lr, float decay, float momentum, float clipnorm, float grad_scale, void* stream);
int crnn_get_iterations(const crnn_handle* h, int64_t* it);
int crnn_set_iterations(crnn_handle* h, int64_t it);
/* model.train_on_batch on HOST buffers in one call (train.py:201-209: what fit_generator does per batch of the generator's dict,
 * utils.py:495-502): x_host = B*imgh*imgw float32 (already normalised, utils.py:415) or, with x_is_u8, the raw 8-bit images (norm() with
 * mean/std runs on the device); labels (B*max_len), label_len (B), input_len (B) int32.  Copies through an internal pinned buffer (a
 * page-locked x_host is used in place), runs crnn_train_fwd_bwd + the optimiser step, reads back the per-sample losses (optional,
 * B floats), their mean and the CTC status (0, or -(b+1) for the first sample whose labels do not fit its input length), synchronises. */
#define CRNN_OPT_ADAM 0
#define CRNN_OPT_SGD 1
typedef struct crnn_optimizer {
    int kind;                       /* CRNN_OPT_ADAM: lr, beta1, beta2, eps, clipnorm;  CRNN_OPT_SGD (Nesterov): lr, decay, momentum, clipnorm */
    float lr, beta1, beta2, eps, decay, momentum, clipnorm;
} crnn_optimizer;
int crnn_train_on_batch_host(crnn_handle* h, const void* x_host, int x_is_u8, float mean, float std, const int32_t* labels_host,
                             const int32_t* label_len_host, const int32_t* input_len_host, int B, uint64_t dropout_seed,
                             const crnn_optimizer* opt, float grad_scale, float* losses_host, float* mean_loss, int32_t* ctc_status, void* stream);
/* status of the last CTC loss launch (read after a stream sync): 0 or -(b+1) for the first infeasible sample */
int crnn_ctc_status(crnn_handle* h, int32_t* status_host, void* stream);

/* ---------------------------------------------------------------- data parallel (NEW capability: the reference is single-device,
 * train.py:111,116 import multi_gpu_model and never call it -- SURVEY 0.8 / 8e).  One process per GPU, batch sharded by rank, the only
 * exchange of the step is a sum all-reduce of the flat fp32 gradient arena (NCCL over NVLink/NVSwitch); 1/world is applied by the optimiser
 * entry points (`grad_scale`), clip-by-global-norm + update run after the reduce, identically on every rank.  libnccl.so.2 is bound at run
 * time (dlopen; the instance already loaded in the process is preferred), so there is no link-time dependency.
 *   crnn_nccl_unique_id   rank 0: 128-byte ncclUniqueId to hand to the other ranks by any side channel
 *   crnn_comm_init_rank   ncclCommInitRank on the current device; the communicator is owned by the handle
 *   crnn_set_comm         alternatively: borrow the caller's ncclComm_t (NULL detaches)
 *   crnn_allreduce_grads  in-place sum all-reduce of arena/grads on `stream` (comm NULL = the handle's) -- SURVEY 8b's entry point
 *   crnn_set_dp_fused     1: crnn_train_fwd_bwd issues the exchange itself, in two buckets: the head gradients (dense1 .. dense2, ~70 % of
 *                         the bytes) on a private stream as soon as their weight-gradient GEMMs are done -- overlapping the whole conv-stack
 *                         backward -- and the conv-stack + STN bucket at the end; part of the captured step graph.  The caller then skips
 *                         crnn_allreduce_grads and passes grad_scale = 1 / world to the optimiser step. */
int crnn_nccl_unique_id(void* id128_out);
int crnn_comm_init_rank(crnn_handle* h, const void* id128, int nranks, int rank);
int crnn_set_comm(crnn_handle* h, void* nccl_comm, int nranks);
int crnn_set_dp_fused(crnn_handle* h, int on);
int crnn_comm_ranks(const crnn_handle* h);
int crnn_allreduce_grads(crnn_handle* h, void* nccl_comm, void* stream);

/* ---------------------------------------------------------------- stand-alone CTC ops on device buffers */
/* K.ctc_batch_cost (utils.py:103) on probs[:, t_off:, :]; grad_u / grad_logits may be NULL.
 * status_dev: one int32, 0 or -(b+1). */
int crnn_ctc_loss_grad(const float* probs_dev, int B, int T, int V, int t_off, const int32_t* labels_dev, int max_len,
                       const int32_t* label_len_dev, const int32_t* input_len_dev, float eps,
                       float* loss_dev, float* grad_u_dev, float* grad_logits_dev, float scale,
                       int32_t* status_dev, void* stream);
/* K.ctc_decode(greedy=True): out (B,T) padded with -1 */
int crnn_ctc_greedy(const float* probs_dev, const int32_t* seq_len_dev, int B, int T, int V, float eps,
                    int32_t* out_dev, int32_t* out_len_dev, float* score_dev, void* stream);
/* K.ctc_decode(greedy=False, beam_width, top_paths=1) as used by DecodeCTCPred.decode (utils.py:347-357) */
int crnn_ctc_beam(const float* probs_dev, const int32_t* seq_len_dev, int B, int T, int V, float eps, int beam_width,
                  int merge_repeated, int32_t* out_dev, int32_t* out_len_dev, float* logprob_dev, void* stream);
/* K.ctc_decode(greedy=False, beam_width, top_paths = P <= beam_width) (utils.py:353-354): out (B,P,T), out_len / logprob (B,P), best path
 * first (TF BeamSearch::TopPaths order); paths beyond the number of leaves come back empty with logprob = -inf */
int crnn_ctc_beam_topk(const float* probs_dev, const int32_t* seq_len_dev, int B, int T, int V, float eps, int beam_width,
                       int merge_repeated, int top_paths, int32_t* out_dev, int32_t* out_len_dev, float* logprob_dev, void* stream);
/* host-buffer variant: copies probs H2D, decodes, copies labels D2H, synchronises */
int crnn_ctc_beam_host(const float* probs_host, int B, int T, int V, float eps, int beam_width, int merge_repeated,
                       int32_t* out_host, int32_t* out_len_host, float* logprob_host, void* stream);
int crnn_ctc_greedy_host(const float* probs_host, int B, int T, int V, float eps,
                         int32_t* out_host, int32_t* out_len_host, float* score_host, void* stream);

/* BilinearInterpolation.call (utils.py:140-232) as a stand-alone op: x (B,H,W) fp32, theta (B,6) -> out (B,H,W), output size = input size
 * (what STN() asks for, utils.py:257); all sampler quirks of the reference kept */
int crnn_bilinear_sample(const float* x_dev, const float* theta_dev, float* out_dev, int B, int H, int W, void* stream);

/* ---------------------------------------------------------------- input pipeline
 * Device-side `norm` of utils.py:415-416 (called per image at utils.py:490): out = (float32(u8) - mean) / std in fp32, bit-identical to
 * numpy.  The host uploads the 8-bit line images produced by open_img (B*imgh*imgw bytes) instead of float32. */
int crnn_normalize_u8(const uint8_t* x_u8_dev, float* out_dev, long long n, float mean, float std, void* stream);

/* ---------------------------------------------------------------- evaluation step
 * Replaces the Python loops of utils.py:262-298 (levenshtein / edit_distance / normalized_edit_distance, called from predict.py:183-191):
 * Levenshtein distance of N (prediction, truth) pairs.  Sequences are int32 symbols (character codes or class indices) padded to
 * `maxlen` (<= 128) per row; lengths are clamped to [0, maxlen].  dist[i] is the exact integer the reference's matrix fill produces;
 * the two means are formed by the caller in the reference's order (sum_i dist_i / N, sum_i dist_i / (len(truth_i) * N)). */
int crnn_edit_distance(const int32_t* a_dev, const int32_t* alen_dev, const int32_t* b_dev, const int32_t* blen_dev, int N, int maxlen,
                       int32_t* dist_dev, void* stream);
int crnn_edit_distance_host(const int32_t* a_host, const int32_t* alen_host, const int32_t* b_host, const int32_t* blen_host, int N, int maxlen,
                            int32_t* dist_host, void* stream);

/* ---------------------------------------------------------------- building blocks exposed for parity tests */
/* C[M,N] = op(A) op(B) (+bias, relu); see csrc/gemm_simt.cu */
int crnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
              int transA, int transB, const float* a_scale, const float* a_shift, const float* bias, int relu,
              int split_k, void* stream);

/* teacher-forced backward of ONE depthwise-separable block (utils.py:43-56) for the parity tests: after a training-mode
 * forward of batch B, reads d(block output) from dout_dev, zeroes "arena/grads", runs that block's backward alone and
 * writes d(block input) to din_dev. */
/* act/block{i} of the non-pooled blocks 1..6 is not written by the forward pass (recomputed where it is consumed); this fills those
 * tensors in for inspection, from the raw pointwise outputs, BatchNorm constants and dropout seed of the last forward call. */
int crnn_debug_materialize_blocks(crnn_handle* h, void* stream);
int crnn_debug_block_backward(crnn_handle* h, int block, const float* dout_dev, float* din_dev, int B, uint64_t dropout_seed, void* stream);

/* tcgen05 / TMEM 3xTF32 kernel of the pointwise convolutions (csrc/gemm_tc.cu): out[m][n] = sum_k f(X[m][k]) * Wop[n][k],
 * f = optional relu6(x*scale[k]+shift[k]); Wop[n][k] = W[k*ldw+n] if w_transposed else W[n*ldw+k]; K % 32 == 0;
 * stats (optional, pre-zeroed double[2N]) receives per-column sum / sum of squares of out.
 * img_scratch: crnn_gemm_tc_scratch_floats(N,K) floats of device memory for the pre-swizzled hi/lo weight images. */
int crnn_gemm_tc(const float* X, int ldx, const float* W, int ldw, int w_transposed, float* out, int ldo, int M, int N, int K,
                 const float* x_scale, const float* x_shift, double* stats, float* img_scratch, void* stream);
long long crnn_gemm_tc_scratch_floats(int N, int K);
/* weight gradient on the tensor cores: dW[ci][co] += sum_m f(X[m][ci]) * dY[m][co]; dW must be pre-zeroed (atomics) */
int crnn_gemm_tc_dw(const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout, float* dW, int ldw, int M,
                    const float* x_scale, const float* x_shift, void* stream);

/* ---------------------------------------------------------------- measurement hooks (bench.py) */
/* kernels launched by this library since load (bench.py's gpu_launches) */
long long crnn_launch_count(void);
/* per-stage CUDA-event timing of the step, recorded on the launching stream while enabled */
int crnn_profile_enable(crnn_handle* h, int on);
int crnn_profile_num_stages(void);
const char* crnn_profile_stage_name(int stage);
int crnn_profile_report(crnn_handle* h, double* ms, double* work, long long* launches);
/* the same records broken down by kernel family as well (one kernel, e.g. xw_gemm_tc_v2_kernel, serves several stages): fam_* arrays are
 * [num_stages * num_families], row-major (stage, family); family 0 collects the kernels that are not tracked individually */
int crnn_profile_num_families(void);
const char* crnn_profile_family_name(int family);
int crnn_profile_report2(crnn_handle* h, double* ms, double* work, long long* launches, double* fam_ms, double* fam_work, long long* fam_launches);

#ifdef __cplusplus
}
#endif
#endif
