// engine.cu -- the CRNN hot path as a fixed launch sequence over one caller-owned workspace, plus the C ABI
// (include/crnn_b200.h).  Mirrors CRNN.get_model (reference utils.py:58-96): STN -> ZeroPadding2D -> 7 depthwise-
// separable blocks -> dense1 -> 2 x Bidirectional GRU/LSTM -> dense2 -> softmax -> CTC (utils.py:98-103), the full
// backward of that graph and the Keras optimiser step (train.py:187-192).
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/crnn_b200.h"
#include "common.cuh"
#include "kernels.h"

long long g_crnn_launches = 0;
int g_crnn_family = CRNN_FAM_OTHER;
int g_crnn_pdl = []() { const char* e = getenv("CRNN_PDL"); return (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0; }();   // common.cuh: programmatic dependent launch (OFF by default: measured slower in every mode)
int g_crnn_side_open = 0;
cudaStream_t g_crnn_pdl_sparse = nullptr; int g_crnn_pdl_sparse_set = 0;
const BnFin* g_crnn_bn_fin = nullptr;   // common.cuh: BatchNorm finalize offered to the next statistics-producing launch
static thread_local char g_err[512] = "";
void crnn_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
#define TRY(expr) do { int _s = (expr); if (_s != CRNN_OK) return _s; } while (0)

namespace {
struct BlockPlan { int cin, cout, ph, pw; };
const BlockPlan kBlocks[7] = {{1, 64, 1, 1}, {64, 128, 1, 1}, {128, 256, 2, 2}, {256, 256, 1, 1}, {256, 512, 1, 2}, {512, 512, 1, 1}, {512, 512, 1, 1}};
constexpr float kBnEps = 1e-3f, kBnMomentum = 0.99f, kKerasEps = 1e-7f;
constexpr float kDropBlock = 0.1f, kDropDense1 = 0.4f, kDropRnn = 0.2f;

struct Tensor { std::string name; int64_t offset; int64_t numel; int is_int; };

// ---- per-stage profiler: CUDA-event pairs around every stage of the step, on the launching stream ----
enum Stage { ST_STN = 0, ST_DWCONV, ST_BN_STATS, ST_GEMM_PW_FWD, ST_ACT_POOL, ST_GEMM_HEAD_FWD, ST_RNN_FWD, ST_SOFTMAX, ST_CTC,
             ST_GEMM_HEAD_BWD, ST_RNN_BWD, ST_ACT_BWD, ST_BN_BWD, ST_GEMM_PW_DW, ST_GEMM_PW_DX, ST_DWCONV_BWD, ST_STN_BWD,
             ST_OPTIM, ST_MISC, ST_COUNT };
const char* kFamilyNames[CRNN_FAM_COUNT] = {"(other)", "xw_gemm_tc_v2_kernel", "xty_gemm_tc_kernel", "gru/lstm_{fwd,bwd}_mma_kernel", "dwconv3x3_rows_*"};
const char* kStageNames[ST_COUNT] = {"stn_fwd", "dwconv_fwd", "bn_stats", "gemm_pw_fwd", "act_pool_fwd", "gemm_head_fwd", "rnn_fwd", "softmax",
                                     "ctc_loss_grad", "gemm_head_bwd", "rnn_bwd", "act_pool_bwd", "bn_bwd", "gemm_pw_dw", "gemm_pw_dx",
                                     "dwconv_bwd", "stn_bwd", "optimizer", "misc"};
struct Prof {
    bool on = false;
    struct Rec { int stage, fam; cudaEvent_t a, b; double work; long long launches; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    cudaEvent_t ev() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
};

struct Layout {
    std::vector<Tensor> tensors;
    std::map<std::string, int> index;
    int64_t cursor = 0;
    int64_t add(const std::string& name, int64_t numel, int elem = 4, int is_int = 0) {
        cursor = (cursor + 255) & ~(int64_t)255;
        int64_t off = cursor;
        index[name] = (int)tensors.size();
        tensors.push_back({name, off, numel, is_int});
        cursor += numel * elem;
        return off;
    }
    // view into an existing arena (no allocation)
    void alias(const std::string& name, int64_t offset, int64_t numel) {
        index[name] = (int)tensors.size();
        tensors.push_back({name, offset, numel, 0});
    }
};
}  // namespace

// ---- NCCL, bound at run time (no link-time dependency: single-GPU users never load it).  Only the handful of entry points the data-parallel
// step needs; the ABI constants (ncclUniqueId = 128 bytes, ncclFloat32 = 7, ncclSum = 0, ncclSuccess = 0) are stable across NCCL 2.x.
struct crnn_nccl_id { char internal[128]; };      // ncclUniqueId (passed by value to ncclCommInitRank)
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, crnn_nccl_id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
int nccl_api() {
    if (g_nccl.lib) return CRNN_OK;
    // prefer the instance that is already in the process (PyTorch loads its bundled libnccl.so.2): one NCCL per process
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { crnn_set_error("NCCL not found: dlopen(libnccl.so.2) failed (%s)", dlerror()); return CRNN_ERR_INVALID; }
    NcclApi a; a.lib = lib;
    *(void**)&a.GetUniqueId = dlsym(lib, "ncclGetUniqueId");
    *(void**)&a.CommInitRank = dlsym(lib, "ncclCommInitRank");
    *(void**)&a.CommDestroy = dlsym(lib, "ncclCommDestroy");
    *(void**)&a.AllReduce = dlsym(lib, "ncclAllReduce");
    *(void**)&a.GetErrorString = dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce) { crnn_set_error("libnccl lacks a required symbol"); return CRNN_ERR_INVALID; }
    g_nccl = a;
    return CRNN_OK;
}
#define NCCL_TRY(expr) do { int _r = (expr); if (_r != 0) { crnn_set_error("%s:%d: %s -> NCCL error %d (%s)", __FILE__, __LINE__, #expr, _r, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); return CRNN_ERR_CUDA; } } while (0)
}  // namespace

struct crnn_handle {
    crnn_config cfg;
    char* base = nullptr;
    size_t bytes = 0;
    Layout L;
    int H, W, Hp, Wp, T, V, U, G, GS, TD, FEAT, maxB;
    StnDims sd;
    int64_t n_params = 0;
    int64_t iterations = 0;
    std::vector<std::pair<std::string, int64_t>> weights;   // trainable, Keras order
    std::vector<std::pair<std::string, int64_t>> stats;     // BN moving statistics
    Prof prof;
    bool gemm_simt = false;   // CRNN_GEMM_SIMT=1: fp32 SIMT GEMM for the pointwise convs instead of the tcgen05 3xTF32 kernel
    int bn2_red_done[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // [block]: reduction pass of the block's BN2 backward already accumulated by the producer of its dy
    bool defer_bn_grads = false;   // full backward: dgamma/dbeta of all 14 BN layers in one launch at the end instead of 14 tiny ones
    bool fuse_bn_red = true;   // CRNN_FUSE_BN_RED=0: separate reduction pass for the ReLU6+BN backward after the depthwise conv
    // CRNN_FWD_FUSED=0: every block materialises its output.  Default: the output of a non-pooled block 1..6 is NOT written in the forward pass --
    // the next block's depthwise conv (dwconv_fused.cu) and its fused backward kernel recompute it from the raw pointwise output.  block_live[i]
    // says whether act/block{i} holds the values of the last forward; crnn_debug_materialize_blocks() fills in the skipped ones (tests).
    bool fwd_fused = true; bool block_live[8] = {true, true, true, true, true, true, true, true};
    uint64_t last_seed = 0; int last_B = 0; bool last_drop = false;
    bool dw_fused = true;      // CRNN_DW_FUSED=0: separate ReLU6+BN-backward apply / depthwise backward-data / backward-weight kernels
    int prio = 0;              // CRNN_PRIO: stream / graph-node priorities of the critical path vs the side branch (0 = none)
    bool bn_tail = true;       // CRNN_BN_TAIL=0: BatchNorm finalize as a launch of its own instead of the last-CTA tail of the kernel that accumulates the statistics
    bool dw_red = true;        // CRNN_DW_RED=0: the depthwise backward-data kernel does not accumulate the BN2-backward reduction of the block below
    // ---- step scheduling: the train step / the predictor forward are captured once per (batch, buffers) into a CUDA graph and
    // replayed; work that is off the activation-gradient critical path (weight gradients, weight-image preparation) runs on a side
    // stream = a parallel branch of the graph.  CRNN_GRAPH=0 / CRNN_OVERLAP=0 switch either off; profiling runs eager + serial.
    bool use_graph = true, overlap = true, serp = true;
    int flip = 0;   // serpentine traversal: consecutive kernels of the conv-stack chain walk their tensors in alternating directions, so each
                    // one starts on the bytes its producer touched last (still in the 126 MB L2) -- CRNN_SERPENTINE=0 disables
    int rv() { const int r = serp ? flip : 0; flip ^= 1; return r; }
    cudaStream_t side = nullptr, cap = nullptr;
    // ---- data parallel (NEW capability): NCCL communicator (owned when created by crnn_comm_init_rank, else borrowed) and the
    // stream its all-reduces run on.  With `dp_fused` the training step itself issues the exchange, in two buckets: the head gradients
    // (dense1 .. dense2 = the tail of the arena, ~70 % of the bytes) as soon as their weight-gradient GEMMs are done, overlapping the whole
    // conv-stack backward; the conv-stack + STN bucket at the end of the step.
    // pinned host staging of crnn_train_on_batch_host: [x (float or u8) | labels | label_len | input_len | losses | status]
    unsigned char* pin = nullptr; size_t pin_bytes = 0;
    void* comm = nullptr; bool comm_owned = false; int comm_ranks = 1; bool dp_fused = false;
    cudaStream_t comm_stream = nullptr;
    int64_t head_off = 0;     // first float of dense1/kernel inside the arenas
    std::vector<cudaEvent_t> evs; size_t ev_used = 0;
    const uint64_t* seed_ptr = nullptr;   // non-null while capturing: dropout kernels read the seed from device memory
    struct StepGraph { int kind, B; const void* p[6]; int drop; int calls; cudaGraphExec_t exec; long long launches; };
    std::vector<StepGraph> graphs;
    int bn_off[16];   // offset (in doubles) of BN n's [sum | sumsq] slots inside act/stats and act/red
    cudaEvent_t ev() {
        if (ev_used == evs.size()) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); evs.push_back(e); }
        return evs[ev_used++];
    }

    float* f(const std::string& name) const {
        auto it = L.index.find(name);
        if (it == L.index.end()) { crnn_set_error("internal: unknown tensor %s", name.c_str()); return nullptr; }
        return reinterpret_cast<float*>(base + L.tensors[it->second].offset);
    }
    float* w(const std::string& n) const { return f(n); }
    float* g(const std::string& n) const { return f("grad/" + n); }
    float* a(const std::string& n) const { return f("act/" + n); }
    std::string rnn(int layer, int dir) const {
        const char* c = cfg.cell == CRNN_CELL_GRU ? "gru" : "lstm";
        char buf[96]; snprintf(buf, sizeof(buf), "bidirectional_%d/%s_%s_%d", layer, dir ? "backward" : "forward", c, layer);
        return buf;
    }
};

namespace {

struct Scope {   // records a [start, stop] event pair around a stage when profiling is enabled
    crnn_handle* h; cudaStream_t st; int idx = -1; long long l0;
    Scope(crnn_handle* h_, cudaStream_t st_, int stage, double work) : h(h_), st(st_) {
        if (!h->prof.on) return;
        Prof::Rec r; r.stage = stage; r.fam = CRNN_FAM_OTHER; r.a = h->prof.ev(); r.b = h->prof.ev(); r.work = work; r.launches = 0;
        l0 = g_crnn_launches; g_crnn_family = CRNN_FAM_OTHER;
        cudaEventRecord(r.a, st);
        idx = (int)h->prof.recs.size(); h->prof.recs.push_back(r);
    }
    ~Scope() { if (idx >= 0) { cudaEventRecord(h->prof.recs[idx].b, st); h->prof.recs[idx].launches = g_crnn_launches - l0; h->prof.recs[idx].fam = g_crnn_family; } }
};
#define ST(stage, work, call) do { Scope _s(h, st, stage, (double)(work)); int _r = (call); if (_r != CRNN_OK) return _r; } while (0)

// Side branch of the step: returns the stream for work that is off the critical path, ordered after everything issued on `st` so far.
// Under stream capture the event pair becomes a graph edge; eagerly it is a real cross-stream dependency.
cudaStream_t side_after(crnn_handle* h, cudaStream_t st) {
    if (!h->overlap || h->prof.on || !h->side) return st;
    cudaEvent_t e = h->ev();
    cudaEventRecord(e, st); cudaStreamWaitEvent(h->side, e, 0);
    g_crnn_side_open = 1;
    return h->side;
}
// everything issued on the side branch so far completes before what follows on `st`
void side_join(crnn_handle* h, cudaStream_t st) {
    if (!h->overlap || h->prof.on || !h->side) return;
    cudaEvent_t e = h->ev();
    cudaEventRecord(e, h->side); cudaStreamWaitEvent(st, e, 0);
    g_crnn_side_open = 0;
}

// ---- data-parallel exchange inside the step (dp_fused): sum all-reduce of arena/grads[off, off + n) on stream s
int dp_allreduce(crnn_handle* h, int64_t off, int64_t n, cudaStream_t s) {
    if (n <= 0) return CRNN_OK;
    float* g = h->f("arena/grads") + off;
    NCCL_TRY(g_nccl.AllReduce(g, g, (size_t)n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, h->comm, s));
    return CRNN_OK;
}
// head bucket (dense1 .. dense2): everything that writes it was issued on the side branch before this call; the reduce runs on the comm
// stream, concurrently with the conv-stack backward that follows on `st` (and its weight-gradient kernels on the side branch)
int dp_reduce_head(crnn_handle* h, cudaStream_t st) {
    if (!h->dp_fused || !h->comm) return CRNN_OK;
    const bool branch = h->overlap && !h->prof.on && h->side && h->comm_stream;
    if (!branch) return dp_allreduce(h, h->head_off, h->n_params - h->head_off, st);
    cudaEvent_t e = h->ev();
    cudaEventRecord(e, h->side); cudaStreamWaitEvent(h->comm_stream, e, 0);
    return dp_allreduce(h, h->head_off, h->n_params - h->head_off, h->comm_stream);
}
// conv-stack + STN bucket at the end of the backward pass (after the side branch has been joined), then join the comm stream
int dp_reduce_tail(crnn_handle* h, cudaStream_t st) {
    if (!h->dp_fused || !h->comm) return CRNN_OK;
    TRY(dp_allreduce(h, 0, h->head_off, st));
    if (h->overlap && !h->prof.on && h->side && h->comm_stream) {
        cudaEvent_t e = h->ev();
        cudaEventRecord(e, h->comm_stream); cudaStreamWaitEvent(st, e, 0);
    }
    return CRNN_OK;
}

int validate(const crnn_config* c) {
    if (!c) { crnn_set_error("null config"); return CRNN_ERR_INVALID; }
    if (c->imgw != 32) { crnn_set_error("imgW must be 32 (the conv stack reduces it to 9 columns)"); return CRNN_ERR_INVALID; }
    if (c->imgh < 40 || c->imgh % 2) { crnn_set_error("imgh must be even and >= 40"); return CRNN_ERR_INVALID; }
    if (c->n_units != 256) { crnn_set_error("n_units must be 256"); return CRNN_ERR_INVALID; }
    if (c->num_classes < 2 || c->num_classes > 1024) { crnn_set_error("num_classes out of range"); return CRNN_ERR_INVALID; }
    if (c->cell != CRNN_CELL_GRU && c->cell != CRNN_CELL_LSTM) { crnn_set_error("bad cell"); return CRNN_ERR_INVALID; }
    if (c->time_dense < 4 || c->time_dense % 4) { crnn_set_error("time_dense must be a multiple of 4"); return CRNN_ERR_INVALID; }
    if (c->max_batch < 1 || c->max_len < 1) { crnn_set_error("bad max_batch/max_len"); return CRNN_ERR_INVALID; }
    return CRNN_OK;
}

void plan(crnn_handle* h) {
    const crnn_config& c = h->cfg;
    h->H = c.imgh; h->W = c.imgw; h->Hp = c.imgh + 4; h->Wp = c.imgw + 4;
    h->T = h->Hp / 2; h->V = c.num_classes; h->U = c.n_units; h->TD = c.time_dense;
    h->G = c.cell == CRNN_CELL_GRU ? 3 : 4; h->GS = c.cell == CRNN_CELL_GRU ? 3 : 5;
    h->FEAT = (h->Wp / 4) * 512; h->maxB = c.max_batch;
    h->sd = stn_dims(h->H, h->W);
    Layout& L = h->L;
    auto W = [&](const std::string& n, int64_t numel) { h->weights.push_back({n, numel}); };
    auto S = [&](const std::string& n, int64_t numel) { h->stats.push_back({n, numel}); };
    // Keras layer_names / weight_names order (models/<name>/final_weights.h5)
    W("conv2d_1/kernel", 25 * 20); W("conv2d_1/bias", 20);
    W("conv2d_2/kernel", 25 * 20 * 20); W("conv2d_2/bias", 20);
    W("dense_1/kernel", (int64_t)h->sd.F * 50); W("dense_1/bias", 50);
    W("dense_2/kernel", 300); W("dense_2/bias", 6);
    for (int i = 1; i <= 7; ++i) {
        const BlockPlan& b = kBlocks[i - 1];
        char n[96];
        snprintf(n, sizeof(n), "depthwise_conv2d_%d/depthwise_kernel", i); W(n, 9 * b.cin);
        for (int k = 0; k < 2; ++k) {
            int bn = 2 * i - 1 + k, ch = k ? b.cout : b.cin;
            if (k == 1) { snprintf(n, sizeof(n), "conv2d_%d/kernel", i + 2); W(n, (int64_t)b.cin * b.cout); }
            snprintf(n, sizeof(n), "batch_normalization_%d/gamma", bn); W(n, ch);
            snprintf(n, sizeof(n), "batch_normalization_%d/beta", bn); W(n, ch);
            snprintf(n, sizeof(n), "batch_normalization_%d/moving_mean", bn); S(n, ch);
            snprintf(n, sizeof(n), "batch_normalization_%d/moving_variance", bn); S(n, ch);
        }
    }
    W("dense1/kernel", (int64_t)h->FEAT * h->TD); W("dense1/bias", h->TD);
    for (int layer = 1; layer <= 2; ++layer)
        for (int d = 0; d < 2; ++d) {
            int kin = layer == 1 ? h->TD : h->U;
            W(h->rnn(layer, d) + "/kernel", (int64_t)kin * h->G * h->U);
            W(h->rnn(layer, d) + "/recurrent_kernel", (int64_t)h->U * h->G * h->U);
            W(h->rnn(layer, d) + "/bias", h->G * h->U);
        }
    W("dense2/kernel", 2 * h->U * h->V); W("dense2/bias", h->V);

    // arenas: every weight padded to a multiple of 4 floats so that all tensors stay 16-byte aligned
    int64_t n = 0;
    std::vector<int64_t> offs;
    for (auto& p : h->weights) { if (p.first == "dense1/kernel") h->head_off = n; offs.push_back(n); n += (p.second + 3) & ~(int64_t)3; }
    h->n_params = n;
    const char* arenas[4] = {"arena/params", "arena/grads", "arena/opt_m", "arena/opt_v"};
    const char* prefix[4] = {"", "grad/", "adam_m/", "adam_v/"};
    for (int a = 0; a < 4; ++a) {
        int64_t base = L.add(arenas[a], n);
        for (size_t i = 0; i < h->weights.size(); ++i) L.alias(std::string(prefix[a]) + h->weights[i].first, base + offs[i] * 4, h->weights[i].second);
    }
    for (auto& p : h->stats) L.add(p.first, p.second);

    // activations (sized for max_batch)
    const int64_t B = h->maxB;
    auto A = [&](const std::string& nm, int64_t numel) { L.add("act/" + nm, numel); };
    A("x", B * h->H * h->W);
    A("p1", B * h->sd.P1h * h->sd.P1w); A("p2", B * h->sd.P2h * h->sd.P2w * 20);
    L.add("act/p2arg", B * h->sd.P2h * h->sd.P2w * 20, 4, 1);
    A("flat", B * h->sd.F); A("loc_d1", B * 50); A("theta", B * 6);
    A("a0", B * h->Hp * h->Wp);
    int hh = h->Hp, ww = h->Wp;
    int64_t max_act = 0;
    for (int i = 1; i <= 7; ++i) {
        const BlockPlan& b = kBlocks[i - 1];
        char nm[32];
        snprintf(nm, sizeof(nm), "dw%d", i); A(nm, B * hh * ww * b.cin);
        snprintf(nm, sizeof(nm), "pw%d", i); A(nm, B * hh * ww * b.cout);
        if (B * hh * ww * b.cout > max_act) max_act = B * hh * ww * b.cout;
        hh /= b.ph; ww /= b.pw;
        snprintf(nm, sizeof(nm), "block%d", i); A(nm, B * hh * ww * b.cout);
        for (int k = 0; k < 2; ++k) {
            int bn = 2 * i - 1 + k, ch = k ? b.cout : b.cin;
            const char* parts[4] = {"scale", "shift", "mean", "invstd"};
            for (auto pt : parts) { char t[48]; snprintf(t, sizeof(t), "bn%d/%s", bn, pt); A(t, ch); }
        }
    }
    const int64_t M = B * h->T;
    if (M * h->FEAT > max_act) max_act = M * h->FEAT;
    A("dense1", M * h->TD);
    for (int layer = 1; layer <= 2; ++layer) {
        char nm[32];
        snprintf(nm, sizeof(nm), "xp%d", layer); A(nm, M * 2 * h->G * h->U);
        snprintf(nm, sizeof(nm), "hs%d", layer); A(nm, M * 2 * h->U);
        snprintf(nm, sizeof(nm), "gates%d", layer); A(nm, M * 2 * h->GS * h->U);
    }
    A("rnn1", M * h->U); A("rnn2drop", M * 2 * h->U);
    A("logits", M * h->V); A("softmax", M * h->V); A("dlogits", M * h->V); A("loss", B);
    // backward scratch
    // d(block output) ping-pongs between gA/gB; the pointwise-output and depthwise-output gradients of every block get their own
    // buffers so that the weight-gradient branch of the step graph can read them without write-after-read hazards
    A("gA", max_act); A("gB", max_act);
    A("gemm_part", (int64_t)8 * h->maxB * h->T * std::max(h->TD, h->U));      // split-K copies of the small head GEMMs (<= 8 copies)
    hh = h->Hp; ww = h->Wp;
    for (int i = 1; i <= 7; ++i) {
        const BlockPlan& b = kBlocks[i - 1];
        char nm2[32];
        snprintf(nm2, sizeof(nm2), "dpw%d", i); A(nm2, B * hh * ww * b.cout);
        snprintf(nm2, sizeof(nm2), "ddw%d", i); A(nm2, B * hh * ww * b.cin);
        hh /= b.ph; ww /= b.pw;
    }
    for (int layer = 1; layer <= 2; ++layer) {
        char nm2[32];
        snprintf(nm2, sizeof(nm2), "dxp%d", layer); A(nm2, M * 2 * h->G * h->U);
        snprintf(nm2, sizeof(nm2), "hprev%d", layer); A(nm2, M * 2 * h->U);
        snprintf(nm2, sizeof(nm2), "rh%d", layer); A(nm2, M * 2 * h->U);
    }
    A("dtheta", B * 6); A("dd1", B * 50); A("dflat", B * h->sd.F); A("ddense1", M * h->TD);
    // pre-swizzled hi/lo weight images of the tcgen05 kernels (gemm_tc.cu), one per GEMM: prepared once per step on the side branch
    A("wimg_d1f", (int64_t)tc_weight_image_floats(h->TD, h->FEAT)); A("wimg_d1b", (int64_t)tc_weight_image_floats(h->FEAT, h->TD));
    for (int layer = 1; layer <= 2; ++layer)
        for (int d = 0; d < 2; ++d) {
            const int kin = layer == 1 ? h->TD : h->U;
            char nm2[32];
            snprintf(nm2, sizeof(nm2), "wimg_r%d%df", layer, d); A(nm2, (int64_t)tc_weight_image_floats(h->G * h->U, kin));
            snprintf(nm2, sizeof(nm2), "wimg_r%d%db", layer, d); A(nm2, (int64_t)tc_weight_image_floats(kin, h->G * h->U));
        }
    for (int i = 2; i <= 7; ++i) {
        const BlockPlan& b = kBlocks[i - 1];
        char nm2[32];
        snprintf(nm2, sizeof(nm2), "wimg_fwd%d", i); A(nm2, (int64_t)tc_weight_image_floats(b.cout, b.cin));
        snprintf(nm2, sizeof(nm2), "wimg_dx%d", i); A(nm2, (int64_t)tc_weight_image_floats(b.cin, b.cout));
    }
    {   // per-BN [sum | sumsq] slots (forward statistics / backward reductions): one memset per pass instead of one per layer
        int off = 0;
        for (int i = 1; i <= 7; ++i)
            for (int k = 0; k < 2; ++k) { h->bn_off[2 * i - 1 + k] = off; off += 2 * (k ? kBlocks[i - 1].cout : kBlocks[i - 1].cin); }
        h->bn_off[15] = off;
        L.add("act/stats", off + 16, 8);           // + one ticket counter per BN (last-CTA finalize, common.cuh), zeroed with the sums
        L.add("act/red", off, 8);
    }
    L.add("act/sumsq", 1, 8); L.add("act/seed", 1, 8);
    L.add("act/status", 1, 4, 1);
    L.add("act/labels", B * c.max_len, 4, 1); L.add("act/label_len", B, 4, 1); L.add("act/input_len", B, 4, 1);
    L.add("act/x_u8", (B * h->H * h->W + 3) / 4, 4, 1);           // raw 8-bit images of crnn_train_on_batch_host (normalised on the device)
    L.cursor = (L.cursor + 255) & ~(int64_t)255;
}

int pick_split(int M, int N, int K) {
    long long ctas = (long long)ceil_div(M, 64) * ceil_div(N, 64);
    int ktiles = ceil_div(K, 16);
    long long s = (148 * 3 + ctas - 1) / ctas;
    if (s > ktiles / 8) s = ktiles / 8;
    if (s < 1) s = 1;
    if (s > 256) s = 256;
    return (int)s;
}

// C = A(MxK) @ B(KxN) + bias (, relu)
int gemm_nn(crnn_handle* h, int stage, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K, const float* bias, int relu,
            const float* sc, const float* sh, cudaStream_t st) {
    GemmArgs g; g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.bias = bias; g.relu = relu; g.a_scale = sc; g.a_shift = sh;
    Scope _s(h, st, stage, 2.0 * M * N * K);
    return launch_gemm_simt(g, st);
}
// dW(KinxN) += X(MxKin)^T @ dY(MxN)   (reduction over rows M; split-K atomics into pre-zeroed dW)
int gemm_tn(crnn_handle* h, int stage, const float* X, int ldx, const float* dY, int ldy, float* dW, int ldw, int Kin, int N, int M, const float* sc, const float* sh, cudaStream_t st) {
    GemmArgs g; g.A = X; g.B = dY; g.C = dW; g.M = Kin; g.N = N; g.K = M; g.lda = ldx; g.ldb = ldy; g.ldc = ldw;
    g.transA = 1; g.a_scale = sc; g.a_shift = sh; g.split_k = pick_split(Kin, N, M);
    if (g.split_k == 1) g.accumulate = 1;   // grads arena is pre-zeroed; keep += semantics either way
    Scope _s(h, st, stage, 2.0 * Kin * N * M);
    return launch_gemm_simt(g, st);
}
// dX(MxKin) (+)= dY(MxN) @ W(KinxN)^T
int gemm_nt(crnn_handle* h, int stage, const float* dY, int ldy, const float* Wt, int ldw, float* dX, int ldx, int M, int Kin, int N, int accumulate, cudaStream_t st) {
    GemmArgs g; g.A = dY; g.B = Wt; g.C = dX; g.M = M; g.N = Kin; g.K = N; g.lda = ldy; g.ldb = ldw; g.ldc = ldx;
    g.transB = 1; g.accumulate = accumulate;
    Scope _s(h, st, stage, 2.0 * M * Kin * N);
    return launch_gemm_simt(g, st);
}

// out[M][N] = [out +] relu?(X[M][K] @ Wop^T + bias) on the tensor cores; `img` = weight image prepared by prep_images()
int tc_xw(crnn_handle* h, int stage, const float* X, int ldx, const float* img, float* out, int ldo, int M, int N, int K,
          const float* bias, int relu, int accumulate, cudaStream_t st) {
    ST(stage, 2.0 * M * N * K, launch_xw_gemm_tc(X, ldx, img, out, ldo, M, N, K, nullptr, nullptr, nullptr, st, bias, relu, accumulate));
    return CRNN_OK;
}
bool tc_ok(const crnn_handle* h, int N, int K) { return !h->gemm_simt && (K % 32 == 0) && (N % 4 == 0) && N >= 64; }

std::string bnname(int bn, const char* leaf) { char b[64]; snprintf(b, sizeof(b), "batch_normalization_%d/%s", bn, leaf); return b; }
std::string actbn(int bn, const char* leaf) { char b[64]; snprintf(b, sizeof(b), "bn%d/%s", bn, leaf); return b; }
std::string nm(const char* fmt, int i) { char b[64]; snprintf(b, sizeof(b), fmt, i); return b; }

double* bn_stats(crnn_handle* h, int bn) { return reinterpret_cast<double*>(h->a("stats")) + h->bn_off[bn]; }
double* bn_red(crnn_handle* h, int bn) { return reinterpret_cast<double*>(h->a("red")) + h->bn_off[bn]; }

// act/stats is zeroed once at the start of a training forward; stats_ready = the producer kernel already accumulated into the slot
// Offer the finalize of BN `bn` to the kernel that is about to accumulate its batch statistics (training only; CRNN_BN_TAIL=0: never).
// `f` must outlive the producer launch and the bn_forward call that follows.
void bn_offer(crnn_handle* h, int bn, long long M, int C, bool training, BnFin& f) {
    g_crnn_bn_fin = nullptr;
    if (!training || !h->bn_tail) return;
    f.stats = bn_stats(h, bn);
    f.ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<double*>(h->a("stats")) + h->bn_off[15] + bn);
    f.M = (double)M; f.C = C;
    f.gamma = h->w(bnname(bn, "gamma")); f.beta = h->w(bnname(bn, "beta"));
    f.mm = h->w(bnname(bn, "moving_mean")); f.mv = h->w(bnname(bn, "moving_variance"));
    f.eps = kBnEps; f.momentum = kBnMomentum;
    f.scale = h->a(actbn(bn, "scale")); f.shift = h->a(actbn(bn, "shift")); f.save_mean = h->a(actbn(bn, "mean")); f.save_invstd = h->a(actbn(bn, "invstd"));
    g_crnn_bn_fin = &f;
}
int bn_forward(crnn_handle* h, int bn, const float* y, long long M, int C, bool training, cudaStream_t st, bool stats_ready = false, bool offered = false) {
    double* stats = bn_stats(h, bn);
    if (training && !stats_ready) ST(ST_BN_STATS, 4.0 * M * C, launch_colstats(y, M, C, stats, st));
    const bool folded = offered && g_crnn_bn_fin == nullptr;     // a producer took the job: its last CTA finalizes
    g_crnn_bn_fin = nullptr;
    if (folded) return CRNN_OK;
    return launch_bn_finalize(stats, M, C, h->w(bnname(bn, "gamma")), h->w(bnname(bn, "beta")), h->w(bnname(bn, "moving_mean")),
                              h->w(bnname(bn, "moving_variance")), kBnEps, kBnMomentum, training ? 1 : 0,
                              h->a(actbn(bn, "scale")), h->a(actbn(bn, "shift")), h->a(actbn(bn, "mean")), h->a(actbn(bn, "invstd")), st);
}

// hi/lo weight images of every tcgen05 GEMM of the step (forward ones; plus the dX ones when training): ONE launch on `ss`
int prep_images(crnn_handle* h, bool training, cudaStream_t ss) {
    cudaStream_t st = ss;
    if (h->gemm_simt) return CRNN_OK;
    const int U = h->U, G = h->G;
    TcPrepBatch t;
    for (int i = 2; i <= 7; ++i) {
        const BlockPlan& b = kBlocks[i - 1];
        const float* W = h->w(nm("conv2d_%d/kernel", i + 2));
        TRY(tc_prep_batch_add(t, W, b.cout, b.cout, b.cin, 1, h->a(nm("wimg_fwd%d", i))));
        if (training) TRY(tc_prep_batch_add(t, W, b.cout, b.cin, b.cout, 0, h->a(nm("wimg_dx%d", i))));
    }
    if (tc_ok(h, h->TD, h->FEAT)) TRY(tc_prep_batch_add(t, h->w("dense1/kernel"), h->TD, h->TD, h->FEAT, 1, h->a("wimg_d1f")));
    if (training && tc_ok(h, h->FEAT, h->TD)) TRY(tc_prep_batch_add(t, h->w("dense1/kernel"), h->TD, h->FEAT, h->TD, 0, h->a("wimg_d1b")));
    for (int layer = 1; layer <= 2; ++layer)
        for (int d = 0; d < 2; ++d) {
            const int kin = layer == 1 ? h->TD : U;
            const float* W = h->w(h->rnn(layer, d) + "/kernel");
            char nf[32], nb[32]; snprintf(nf, sizeof(nf), "wimg_r%d%df", layer, d); snprintf(nb, sizeof(nb), "wimg_r%d%db", layer, d);
            if (tc_ok(h, G * U, kin)) TRY(tc_prep_batch_add(t, W, G * U, G * U, kin, 1, h->a(nf)));
            if (training && tc_ok(h, kin, G * U)) TRY(tc_prep_batch_add(t, W, G * U, kin, G * U, 0, h->a(nb)));
        }
    ST(ST_MISC, 0, launch_prep_weight_images_batch(t, ss));
    return CRNN_OK;
}

int forward(crnn_handle* h, const float* x, int B, bool training, uint64_t seed, cudaStream_t st) {
    if (B < 1 || B > h->maxB) { crnn_set_error("batch %d outside [1,%d]", B, h->maxB); return CRNN_ERR_INVALID; }
    const bool drop = training && seed != 0;
    const int H = h->H, W = h->W, U = h->U, G = h->G, T = h->T, V = h->V;
    h->ev_used = 0; h->flip = 0; g_crnn_side_open = 0;
    // ---- side branch: weight images of all tensor-core GEMMs of this step (weights only change in the optimiser)
    TRY(prep_images(h, training, side_after(h, st)));
    if (training) CUDA_TRY(cudaMemsetAsync(h->a("stats"), 0, sizeof(double) * (h->bn_off[15] + 16), st));
    // ---- STN (utils.py:247-258)
    ST(ST_STN, 0, launch_stn_trunk_fwd(x, h->w("conv2d_1/kernel"), h->w("conv2d_1/bias"), h->w("conv2d_2/kernel"), h->w("conv2d_2/bias"),
                             h->a("p1"), h->a("p2"), reinterpret_cast<int*>(h->a("p2arg")), h->a("flat"), B, H, W, st));
    ST(ST_STN, 0, launch_stn_head_fwd(h->a("flat"), h->w("dense_1/kernel"), h->w("dense_1/bias"), h->w("dense_2/kernel"), h->w("dense_2/bias"),
                                      h->a("loc_d1"), h->a("theta"), B, h->sd.F, st));
    ST(ST_STN, 0, launch_stn_sample_fwd(x, h->a("theta"), h->a("a0"), B, H, W, 2, st));
    // ---- depthwise-separable stack (utils.py:43-56, 64-70)
    const float* in = h->a("a0");
    int hh = h->Hp, ww = h->Wp;
    for (int i = 1; i <= 7; ++i) {
        const BlockPlan& b = kBlocks[i - 1];
        const long long M = (long long)B * hh * ww;
        float* dw = h->a(nm("dw%d", i)); float* pw = h->a(nm("pw%d", i)); float* out = h->a(nm("block%d", i));
        const bool dw_stats = training && (b.cin % 4 == 0);       // BN statistics of the depthwise output fused into the conv kernel
        BnFin fin1, fin2;
        bn_offer(h, 2 * i - 1, M, b.cin, training, fin1);
        if (i >= 2 && !h->block_live[i - 1]) {
            // the block below did not write its output: BN + ReLU6 + Dropout of its raw pointwise output are applied while the conv stages its rows
            ST(ST_DWCONV, 8.0 * M * b.cin, launch_dwconv_fwd_fused(h->a(nm("pw%d", i - 1)), h->a(actbn(2 * i - 2, "scale")), h->a(actbn(2 * i - 2, "shift")),
                                                                   drop ? kDropBlock : 0.f, seed, (uint32_t)(i - 1), h->seed_ptr, h->w(nm("depthwise_conv2d_%d/depthwise_kernel", i)),
                                                                   dw, dw_stats ? bn_stats(h, 2 * i - 1) : nullptr, B, hh, ww, b.cin, h->rv(), st));
        } else {
            ST(ST_DWCONV, 8.0 * M * b.cin, launch_dwconv_fwd(in, h->w(nm("depthwise_conv2d_%d/depthwise_kernel", i)), dw, B, hh, ww, b.cin, st,
                                                             dw_stats ? bn_stats(h, 2 * i - 1) : nullptr, h->rv()));
        }
        TRY(bn_forward(h, 2 * i - 1, dw, M, b.cin, training, st, dw_stats, training && h->bn_tail));
        bn_offer(h, 2 * i, M, b.cout, training, fin2);
        const bool tc = !h->gemm_simt && (b.cin % 32 == 0);
        bool pw_stats = false;
        if (tc) {
            if (i == 2) side_join(h, st);                          // the weight images are ready
            ST(ST_GEMM_PW_FWD, 2.0 * M * b.cout * b.cin, launch_xw_gemm_tc(dw, b.cin, h->a(nm("wimg_fwd%d", i)), pw, b.cout, (int)M, b.cout, b.cin,
                                                                          h->a(actbn(2 * i - 1, "scale")), h->a(actbn(2 * i - 1, "shift")),
                                                                          training ? bn_stats(h, 2 * i) : nullptr, st, nullptr, 0, 0, h->rv()));
            pw_stats = true;
        } else if (b.cin == 1) {                                   // block 1: the "GEMM" is an outer product
            ST(ST_GEMM_PW_FWD, 2.0 * M * b.cout, launch_pw1_fwd(dw, h->a(actbn(1, "scale")), h->a(actbn(1, "shift")), h->w(nm("conv2d_%d/kernel", i + 2)), pw, M, b.cout,
                                                                training ? bn_stats(h, 2 * i) : nullptr, st, h->rv()));
            pw_stats = true;
        } else {
            TRY(gemm_nn(h, ST_GEMM_PW_FWD, dw, b.cin, h->w(nm("conv2d_%d/kernel", i + 2)), b.cout, pw, b.cout, (int)M, b.cout, b.cin, nullptr, 0,
                        h->a(actbn(2 * i - 1, "scale")), h->a(actbn(2 * i - 1, "shift")), st));
        }
        TRY(bn_forward(h, 2 * i, pw, M, b.cout, training, st, pw_stats, training && h->bn_tail));
        // skip writing this block's output when (a) the next block's conv can recompute it (non-pooled, shapes covered) and (b) the backward pass
        // will not read it: the fused backward kernel with the fused BN2 reduction recomputes it as well (block_backward: `rrp` path)
        h->block_live[i] = !(i <= 6 && h->fwd_fused && b.ph == 1 && b.pw == 1 && dwconv_bwd_fused_covers(hh, ww, b.cout) &&
                             h->dw_fused && h->fuse_bn_red && h->dw_red && !h->gemm_simt && (kBlocks[i].cout % 32 == 0) && (kBlocks[i].cin % 4 == 0));
        if (h->block_live[i])
            ST(ST_ACT_POOL, 4.0 * M * b.cout * (1.0 + 1.0 / (b.ph * b.pw)), launch_act_pool_fwd(pw, h->a(actbn(2 * i, "scale")), h->a(actbn(2 * i, "shift")), out, B, hh, ww, b.cout, b.ph, b.pw,
                                    drop ? kDropBlock : 0.f, seed, (uint32_t)i, st, h->seed_ptr, h->rv()));
        hh /= b.ph; ww /= b.pw; in = out;
    }
    if (h->gemm_simt) side_join(h, st);
    // ---- dense1 (utils.py:72-75): (B,T,9,512) is already (B*T, 4608) with feature = w*512+c
    const int M = B * T;
    const int ks1 = tc_ok(h, h->TD, h->FEAT) ? std::min(8, xw_gemm_tc_pick_ksplit(M, h->TD, h->FEAT)) : 1;
    if (ks1 > 1) {
        // 33 output tiles for 148 SMs: split K = 4608 over `ks1` CTAs per tile, then one pass sums the copies + bias + ReLU + dropout
        float* part = h->a("gemm_part");
        ST(ST_GEMM_HEAD_FWD, 2.0 * M * h->TD * h->FEAT, launch_xw_gemm_tc(in, h->FEAT, h->a("wimg_d1f"), part, h->TD, M, h->TD, h->FEAT, nullptr, nullptr, nullptr, st,
                                                                           nullptr, 0, 0, 0, ks1, (long long)M * h->TD));
        ST(ST_MISC, 0, launch_sum_partials(part, ks1, (long long)M * h->TD, M, h->TD, h->w("dense1/bias"), 1, h->a("dense1"), h->TD,
                                           drop ? kDropDense1 : 0.f, seed, 8, st, h->seed_ptr));
    } else {
        if (tc_ok(h, h->TD, h->FEAT)) TRY(tc_xw(h, ST_GEMM_HEAD_FWD, in, h->FEAT, h->a("wimg_d1f"), h->a("dense1"), h->TD, M, h->TD, h->FEAT, h->w("dense1/bias"), 1, 0, st));
        else TRY(gemm_nn(h, ST_GEMM_HEAD_FWD, in, h->FEAT, h->w("dense1/kernel"), h->TD, h->a("dense1"), h->TD, M, h->TD, h->FEAT, h->w("dense1/bias"), 1, nullptr, nullptr, st));
        if (drop) ST(ST_MISC, 0, launch_dropout_fwd(h->a("dense1"), (long long)M * h->TD, kDropDense1, seed, 8, st, h->seed_ptr));
    }
    // ---- two bidirectional recurrent layers (utils.py:77-82)
    const float* rin = h->a("dense1"); int kin = h->TD;
    for (int layer = 1; layer <= 2; ++layer) {
        float* xp = h->a(nm("xp%d", layer)); float* hs = h->a(nm("hs%d", layer));
        for (int d = 0; d < 2; ++d) {
            char nf[32]; snprintf(nf, sizeof(nf), "wimg_r%d%df", layer, d);
            if (tc_ok(h, G * U, kin)) TRY(tc_xw(h, ST_GEMM_HEAD_FWD, rin, kin, h->a(nf), xp + d * G * U, 2 * G * U, M, G * U, kin,
                                                h->w(h->rnn(layer, d) + "/bias"), 0, 0, st));
            else TRY(gemm_nn(h, ST_GEMM_HEAD_FWD, rin, kin, h->w(h->rnn(layer, d) + "/kernel"), G * U, xp + d * G * U, 2 * G * U, M, G * U, kin,
                             h->w(h->rnn(layer, d) + "/bias"), 0, nullptr, nullptr, st));
        }
        {
            const float* U0 = h->w(h->rnn(layer, 0) + "/recurrent_kernel"); const float* U1 = h->w(h->rnn(layer, 1) + "/recurrent_kernel");
            float* gsave = training ? h->a(nm("gates%d", layer)) : nullptr;
            const double work = 4.0 * B * T * 2 * (G * U + U + (training ? h->GS * U : 0));
            if (h->cfg.cell == CRNN_CELL_GRU) ST(ST_RNN_FWD, work, launch_gru_fwd_mma(xp, U0, U1, hs, gsave, B, T, st));
            else ST(ST_RNN_FWD, work, launch_lstm_fwd_mma(xp, U0, U1, hs, gsave, B, T, st));
        }
        if (layer == 1) { ST(ST_MISC, 0, launch_sum_dirs(hs, h->a("rnn1"), M, U, st)); rin = h->a("rnn1"); kin = U; }   // merge_mode='sum'
    }
    const float* head_in = h->a("hs2");                                                                      // merge_mode='concat'
    if (drop) {
        ST(ST_MISC, 0, launch_dropout_copy(h->a("hs2"), h->a("rnn2drop"), (long long)M * 2 * U, kDropRnn, seed, 9, st, h->seed_ptr));
        head_in = h->a("rnn2drop");
    }
    // ---- dense2 + softmax (utils.py:85-86)
    TRY(gemm_nn(h, ST_GEMM_HEAD_FWD, head_in, 2 * U, h->w("dense2/kernel"), V, h->a("logits"), V, M, V, 2 * U, h->w("dense2/bias"), 0, nullptr, nullptr, st));
    ST(ST_SOFTMAX, 0, launch_softmax_rows(h->a("logits"), h->a("softmax"), M, V, st));
    return CRNN_OK;
}

// Recurrent layer backward.  Critical path (on st): BPTT kernel -> dX of the input projections.  Side branch: the three weight
// gradients per direction + the bias column sums.
int rnn_backward(crnn_handle* h, int layer, const float* dout /*(M,2,U)*/, const float* rin, int kin, float* dx /*(M,kin)*/, int B, cudaStream_t st) {
    const int U = h->U, G = h->G, T = h->T, M = B * T;
    float* dxp = h->a(nm("dxp%d", layer)); float* hprev = h->a(nm("hprev%d", layer)); float* rh = h->a(nm("rh%d", layer));
    const double bwork = 4.0 * B * T * 2 * (2 * U + h->GS * U + G * U + 2 * U);
    if (h->cfg.cell == CRNN_CELL_GRU) {
        ST(ST_RNN_BWD, bwork, launch_gru_bwd_mma(dout, h->a(nm("hs%d", layer)), h->a(nm("gates%d", layer)), h->w(h->rnn(layer, 0) + "/recurrent_kernel"),
                                                 h->w(h->rnn(layer, 1) + "/recurrent_kernel"), dxp, hprev, rh, B, T, st));
    } else {
        ST(ST_RNN_BWD, bwork, launch_lstm_bwd_mma(dout, h->a(nm("hs%d", layer)), h->a(nm("gates%d", layer)), h->w(h->rnn(layer, 0) + "/recurrent_kernel"),
                                                  h->w(h->rnn(layer, 1) + "/recurrent_kernel"), dxp, hprev, B, T, st));
    }
    cudaStream_t ss = side_after(h, st);
    for (int d = 0; d < 2; ++d) {
        const std::string base = h->rnn(layer, d);
        const float* dxd = dxp + d * G * U;
        float* gU = h->g(base + "/recurrent_kernel");
        const bool tc = !h->gemm_simt;
        auto dwgemm = [&](const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout, float* dW, int ldw) -> int {
            if (tc) { ST(ST_GEMM_HEAD_BWD, 2.0 * M * Cin * Cout, launch_xty_gemm_tc(X, ldx, Cin, dY, ldy, Cout, dW, ldw, M, nullptr, nullptr, ss)); return CRNN_OK; }
            return gemm_tn(h, ST_GEMM_HEAD_BWD, X, ldx, dY, ldy, dW, ldw, Cin, Cout, M, nullptr, nullptr, ss);
        };
        if (h->cfg.cell == CRNN_CELL_GRU) {
            TRY(dwgemm(hprev + d * U, 2 * U, U, dxd, 2 * G * U, 2 * U, gU, G * U));
            TRY(dwgemm(rh + d * U, 2 * U, U, dxd + 2 * U, 2 * G * U, U, gU + 2 * U, G * U));
        } else {
            TRY(dwgemm(hprev + d * U, 2 * U, U, dxd, 2 * G * U, G * U, gU, G * U));
        }
        TRY(dwgemm(rin, kin, kin, dxd, 2 * G * U, G * U, h->g(base + "/kernel"), G * U));
        ST(ST_MISC, 0, launch_colsum(dxd, M, G * U, 2 * G * U, h->g(base + "/bias"), ss));
    }
    const int ksx = tc_ok(h, kin, G * U) ? std::min(4, xw_gemm_tc_pick_ksplit(M, kin, G * U)) : 1;
    if (ksx > 1) {
        // 33-66 output tiles: both directions x `ksx` k-slices write disjoint copies, one pass adds them up (fixed order)
        float* part = h->a("gemm_part");
        const long long stride = (long long)M * kin;
        for (int d = 0; d < 2; ++d) {
            char nb[32]; snprintf(nb, sizeof(nb), "wimg_r%d%db", layer, d);
            ST(ST_GEMM_HEAD_BWD, 2.0 * M * kin * G * U, launch_xw_gemm_tc(dxp + d * G * U, 2 * G * U, h->a(nb), part + (size_t)d * ksx * stride, kin, M, kin, G * U,
                                                                          nullptr, nullptr, nullptr, st, nullptr, 0, 0, 0, ksx, stride));
        }
        ST(ST_MISC, 0, launch_sum_partials(part, 2 * ksx, stride, M, kin, nullptr, 0, dx, kin, 0.f, 0, 0, st, nullptr));
        return CRNN_OK;
    }
    for (int d = 0; d < 2; ++d) {
        const std::string base = h->rnn(layer, d);
        const float* dxd = dxp + d * G * U;
        char nb[32]; snprintf(nb, sizeof(nb), "wimg_r%d%db", layer, d);
        if (tc_ok(h, kin, G * U)) TRY(tc_xw(h, ST_GEMM_HEAD_BWD, dxd, 2 * G * U, h->a(nb), dx, kin, M, kin, G * U, nullptr, 0, d, st));
        else TRY(gemm_nt(h, ST_GEMM_HEAD_BWD, dxd, 2 * G * U, h->w(base + "/kernel"), G * U, dx, kin, M, kin, G * U, d, st));
    }
    return CRNN_OK;
}

// Backward of depthwise-separable block i: `cur` holds d(block output) on entry, d(block input) lands in `other`.
// Critical path (st): BN/ReLU6/pool backward -> dX GEMM -> BN/ReLU6 backward -> depthwise conv backward-data.
// Side branch: pointwise dW GEMM and depthwise dW (they read the block's own dpw/ddw buffers, which nobody overwrites this step).
int block_backward(crnn_handle* h, int i, int hh, int ww, float* cur, float* other, int B, bool drop, uint64_t seed, cudaStream_t st) {
    const BlockPlan& b = kBlocks[i - 1];
    const long long Mi = (long long)B * hh * ww;
    const float* dw = h->a(nm("dw%d", i)); const float* pw = h->a(nm("pw%d", i));
    float* dpw = h->a(nm("dpw%d", i)); float* ddw = h->a(nm("ddw%d", i));
    const int bn1 = 2 * i - 1, bn2 = 2 * i;
    bool fused_red = false;            // reduction pass of the BN1 backward done by the dX GEMM epilogue
    ST(ST_ACT_BWD, 4.0 * Mi * b.cout * (3.0 + 2.0 / (b.ph * b.pw)),
       launch_act_pool_bn_bwd(cur, pw, h->a(actbn(bn2, "scale")), h->a(actbn(bn2, "shift")), h->a(actbn(bn2, "mean")), h->a(actbn(bn2, "invstd")),
                              h->w(bnname(bn2, "gamma")), dpw, bn_red(h, bn2), h->g(bnname(bn2, "gamma")), h->g(bnname(bn2, "beta")),
                              B, hh, ww, b.cout, b.ph, b.pw, drop ? kDropBlock : 0.f, seed, (uint32_t)i, st, h->seed_ptr, h->rv(), h->defer_bn_grads ? 0 : 1, h->bn2_red_done[i]));
    h->bn2_red_done[i] = 0;
    h->rv();   // two kernels (reduce, apply): two direction flips
    if (b.cin == 1) {
        // block 1: dW[co] = sum_m f(x[m]) dY[m][co] and dX[m] = sum_co dY[m][co] W[co] in ONE pass over dY
        ST(ST_GEMM_PW_DX, 4.0 * Mi * b.cout, launch_pw1_bwd(dw, h->a(actbn(bn1, "scale")), h->a(actbn(bn1, "shift")), dpw, h->w(nm("conv2d_%d/kernel", i + 2)),
                                                            ddw, h->g(nm("conv2d_%d/kernel", i + 2)), Mi, b.cout, st, h->rv()));
    } else {
        cudaStream_t ss = side_after(h, st);
        if (!h->gemm_simt && (b.cin % 4 == 0)) {
            ST(ST_GEMM_PW_DW, 2.0 * Mi * b.cin * b.cout, launch_xty_gemm_tc(dw, b.cin, b.cin, dpw, b.cout, b.cout, h->g(nm("conv2d_%d/kernel", i + 2)), b.cout, (int)Mi,
                                                                           h->a(actbn(bn1, "scale")), h->a(actbn(bn1, "shift")), ss));
        } else {
            TRY(gemm_tn(h, ST_GEMM_PW_DW, dw, b.cin, dpw, b.cout, h->g(nm("conv2d_%d/kernel", i + 2)), b.cout, b.cin, b.cout, (int)Mi,
                        h->a(actbn(bn1, "scale")), h->a(actbn(bn1, "shift")), ss));
        }
        if (!h->gemm_simt && (b.cout % 32 == 0) && (b.cin % 4 == 0)) {
            // the epilogue also accumulates the reduction pass of the ReLU6+BN backward below (sum dz, sum dz*xhat per channel)
            const TcBnRed rr = {dw, h->a(actbn(bn1, "scale")), h->a(actbn(bn1, "shift")), h->a(actbn(bn1, "mean")), h->a(actbn(bn1, "invstd"))};
            fused_red = h->fuse_bn_red;
            ST(ST_GEMM_PW_DX, 2.0 * Mi * b.cin * b.cout, launch_xw_gemm_tc(dpw, b.cout, h->a(nm("wimg_dx%d", i)), ddw, b.cin, (int)Mi, b.cin, b.cout, nullptr, nullptr,
                                                                          fused_red ? bn_red(h, bn1) : nullptr, st, nullptr, 0, 0, h->rv(), 1, 0, fused_red ? &rr : nullptr));
        } else {
            TRY(gemm_nt(h, ST_GEMM_PW_DX, dpw, b.cout, h->w(nm("conv2d_%d/kernel", i + 2)), b.cout, ddw, b.cin, (int)Mi, b.cin, b.cout, 0, st));
        }
    }
    const float* bin = i == 1 ? h->a("a0") : h->a(nm("block%d", i - 1));
    // `other` becomes d(output of block i-1): when that block is not pooled, its BN2-backward reduction pass is accumulated by the kernel that writes it
    DwRowsRed rr; const DwRowsRed* rrp = nullptr; double* rbuf = nullptr;
    const bool prev_plain = i >= 2 && h->fuse_bn_red && h->dw_red && kBlocks[i - 2].ph == 1 && kBlocks[i - 2].pw == 1;
    if (prev_plain) {
        const int pbn = 2 * (i - 1);
        rr.y = h->a(nm("pw%d", i - 1)); rr.scale = h->a(actbn(pbn, "scale")); rr.shift = h->a(actbn(pbn, "shift")); rr.mean = h->a(actbn(pbn, "mean"));
        rr.invstd = h->a(actbn(pbn, "invstd")); rr.rate = drop ? kDropBlock : 0.f; rr.seed = seed; rr.layer = (uint32_t)(i - 1); rr.seed_ptr = h->seed_ptr;
        rrp = &rr; rbuf = bn_red(h, pbn);
    }
    if (fused_red && h->dw_fused && dwconv_bwd_fused_covers(hh, ww, b.cin)) {
        // one pass: ReLU6+BN backward apply, depthwise backward-data and backward-weight (dwconv_bwd_fused.cu); with `rrp` the block input is
        // recomputed from the raw activation of the block below, whose BN2-backward reduction is accumulated on the way
        ST(ST_DWCONV_BWD, 16.0 * Mi * b.cin,
           launch_dwconv_bwd_fused(ddw, dw, rrp ? rr.y : bin, h->w(nm("depthwise_conv2d_%d/depthwise_kernel", i)), other, h->g(nm("depthwise_conv2d_%d/depthwise_kernel", i)),
                                   h->a(actbn(bn1, "scale")), h->a(actbn(bn1, "shift")), h->a(actbn(bn1, "mean")), h->a(actbn(bn1, "invstd")), h->w(bnname(bn1, "gamma")),
                                   bn_red(h, bn1), B, hh, ww, b.cin, h->rv(), rrp, rbuf, st));
        if (!h->defer_bn_grads) ST(ST_BN_BWD, 0, launch_bn_param_grads(bn_red(h, bn1), h->g(bnname(bn1, "gamma")), h->g(bnname(bn1, "beta")), b.cin, st));
        if (i >= 2) h->bn2_red_done[i - 1] = rrp ? 1 : 0;
        return CRNN_OK;
    }
    if (i >= 2 && !h->block_live[i - 1]) { crnn_set_error("internal: block %d output was not materialised by the forward pass", i - 1); return CRNN_ERR_INVALID; }
    if (!h->defer_bn_grads) { rrp = nullptr; rbuf = nullptr; }      // the separate kernels only fuse that reduction inside the full backward
    ST(ST_BN_BWD, 20.0 * Mi * b.cin,
       launch_relu6_bn_bwd(ddw, dw, h->a(actbn(bn1, "scale")), h->a(actbn(bn1, "shift")), h->a(actbn(bn1, "mean")), h->a(actbn(bn1, "invstd")),
                           h->w(bnname(bn1, "gamma")), ddw, bn_red(h, bn1), h->g(bnname(bn1, "gamma")), h->g(bnname(bn1, "beta")), Mi, b.cin, st, h->rv(), fused_red ? 1 : 0, h->defer_bn_grads ? 0 : 1));
    h->rv();
    {
        cudaStream_t ss = side_after(h, st);
        ST(ST_DWCONV_BWD, 8.0 * Mi * b.cin, launch_dwconv_bwd_weight(bin, ddw, h->g(nm("depthwise_conv2d_%d/depthwise_kernel", i)), B, hh, ww, b.cin, ss));
    }
    int done = 0;
    ST(ST_DWCONV_BWD, (rrp ? 12.0 : 8.0) * Mi * b.cin /* + one read of the block-below's raw activation for the fused reduction */, launch_dwconv_bwd_data(ddw, h->w(nm("depthwise_conv2d_%d/depthwise_kernel", i)), other, B, hh, ww, b.cin, 0, st, h->rv(), rrp, rbuf, &done));
    if (i >= 2) h->bn2_red_done[i - 1] = done;
    return CRNN_OK;
}

int backward(crnn_handle* h, const float* x, const int* labels, const int* label_len, const int* input_len, int B,
             float* loss, uint64_t seed, cudaStream_t st) {
    const bool drop = seed != 0;
    const int U = h->U, T = h->T, V = h->V, M = B * T;
    CUDA_TRY(cudaMemsetAsync(h->f("arena/grads"), 0, sizeof(float) * (size_t)h->n_params, st));
    CUDA_TRY(cudaMemsetAsync(h->a("red"), 0, sizeof(double) * h->bn_off[15], st));
    // ---- CTC (utils.py:98-103); mean over the batch (identity Keras loss, train.py:192) => scale 1/B
    ST(ST_CTC, 0, launch_ctc_loss_grad(h->a("softmax"), 2, labels, h->cfg.max_len, label_len, input_len, B, T, V, kKerasEps, loss, nullptr,
                             h->a("dlogits"), 1.f / (float)B, reinterpret_cast<int*>(h->a("status")), st));
    float* gA = h->a("gA"); float* gB = h->a("gB");
    // ---- dense2
    const float* head_in = drop ? h->a("rnn2drop") : h->a("hs2");
    {
        cudaStream_t ss = side_after(h, st);
        TRY(gemm_tn(h, ST_GEMM_HEAD_BWD, head_in, 2 * U, h->a("dlogits"), V, h->g("dense2/kernel"), V, 2 * U, V, M, nullptr, nullptr, ss));
        ST(ST_MISC, 0, launch_colsum(h->a("dlogits"), M, V, V, h->g("dense2/bias"), ss));
    }
    TRY(gemm_nt(h, ST_GEMM_HEAD_BWD, h->a("dlogits"), V, h->w("dense2/kernel"), V, gA, 2 * U, M, 2 * U, V, 0, st));
    if (drop) ST(ST_MISC, 0, launch_dropout_fwd(gA, (long long)M * 2 * U, kDropRnn, seed, 9, st, h->seed_ptr));
    // ---- recurrent layers
    TRY(rnn_backward(h, 2, gA, h->a("rnn1"), U, gB, B, st));            // gB = d rnn1 (M,U)
    ST(ST_MISC, 0, launch_dup_dirs(gB, gA, M, U, st));                              // 'sum' merge: same gradient to both directions
    float* dd = h->a("ddense1");                                         // own buffer: the side branch reads it while gB is recycled
    TRY(rnn_backward(h, 1, gA, h->a("dense1"), h->TD, dd, B, st));       // dd = d dense1 (M,TD)
    // ---- dense1
    ST(ST_MISC, 0, launch_relu_dropout_bwd(dd, h->a("dense1"), (long long)M * h->TD, drop ? kDropDense1 : 0.f, seed, 8, st));
    {
        cudaStream_t ss = side_after(h, st);
        if (!h->gemm_simt) ST(ST_GEMM_HEAD_BWD, 2.0 * M * h->FEAT * h->TD, launch_xty_gemm_tc(h->a("block7"), h->FEAT, h->FEAT, dd, h->TD, h->TD, h->g("dense1/kernel"), h->TD, M, nullptr, nullptr, ss));
        else TRY(gemm_tn(h, ST_GEMM_HEAD_BWD, h->a("block7"), h->FEAT, dd, h->TD, h->g("dense1/kernel"), h->TD, h->FEAT, h->TD, M, nullptr, nullptr, ss));
        ST(ST_MISC, 0, launch_colsum(dd, M, h->TD, h->TD, h->g("dense1/bias"), ss));
    }
    TRY(dp_reduce_head(h, st));                                          // data parallel: the head gradients are complete on the side branch
    if (tc_ok(h, h->FEAT, h->TD)) TRY(tc_xw(h, ST_GEMM_HEAD_BWD, dd, h->TD, h->a("wimg_d1b"), gA, h->FEAT, M, h->FEAT, h->TD, nullptr, 0, 0, st));
    else TRY(gemm_nt(h, ST_GEMM_HEAD_BWD, dd, h->TD, h->w("dense1/kernel"), h->TD, gA, h->FEAT, M, h->FEAT, h->TD, 0, st));
    // ---- conv stack, reverse
    float* cur = gA; float* other = gB;
    int dims_h[8], dims_w[8];
    dims_h[1] = h->Hp; dims_w[1] = h->Wp;
    for (int i = 1; i < 7; ++i) { dims_h[i + 1] = dims_h[i] / kBlocks[i - 1].ph; dims_w[i + 1] = dims_w[i] / kBlocks[i - 1].pw; }
    h->defer_bn_grads = true;
    for (int i = 0; i < 8; ++i) h->bn2_red_done[i] = 0;     // no stale "reduction already done" flag from a step that failed half-way
    for (int i = 7; i >= 1; --i) {
        const int rc = block_backward(h, i, dims_h[i], dims_w[i], cur, other, B, drop, seed, st);
        if (rc != CRNN_OK) { h->defer_bn_grads = false; return rc; }
        float* t = cur; cur = other; other = t;
    }
    h->defer_bn_grads = false;
    {
        BnGradTable tb; tb.n = 14;
        for (int bn = 1; bn <= 14; ++bn) {
            tb.red[bn - 1] = bn_red(h, bn); tb.dgamma[bn - 1] = h->g(bnname(bn, "gamma")); tb.dbeta[bn - 1] = h->g(bnname(bn, "beta"));
            tb.C[bn - 1] = (bn & 1) ? kBlocks[(bn - 1) / 2].cin : kBlocks[(bn - 1) / 2].cout;
        }
        ST(ST_BN_BWD, 0, launch_bn_param_grads_all(tb, side_after(h, st)));
    }
    // ---- STN: sampler -> theta -> localisation net
    CUDA_TRY(cudaMemsetAsync(h->a("dtheta"), 0, sizeof(float) * 6 * B, st));
    ST(ST_STN_BWD, 0, launch_stn_sample_bwd(x, h->a("theta"), cur, h->a("dtheta"), B, h->H, h->W, 2, st));
    // critical path: dd1 = relu'(loc_d1) * (dtheta @ W2^T), dflat = dd1 @ W1^T in one kernel; parameter gradients on the side branch
    ST(ST_STN_BWD, 0, launch_stn_head_bwd(h->a("dtheta"), h->a("loc_d1"), h->w("dense_1/kernel"), h->w("dense_2/kernel"), h->a("dd1"), h->a("dflat"), B, h->sd.F, st));
    {
        cudaStream_t ss = side_after(h, st);
        TRY(gemm_tn(h, ST_STN_BWD, h->a("loc_d1"), 50, h->a("dtheta"), 6, h->g("dense_2/kernel"), 6, 50, 6, B, nullptr, nullptr, ss));
        ST(ST_STN_BWD, 0, launch_colsum(h->a("dtheta"), B, 6, 6, h->g("dense_2/bias"), ss));
        TRY(gemm_tn(h, ST_STN_BWD, h->a("flat"), h->sd.F, h->a("dd1"), 50, h->g("dense_1/kernel"), 50, h->sd.F, 50, B, nullptr, nullptr, ss));
        ST(ST_STN_BWD, 0, launch_colsum(h->a("dd1"), B, 50, 50, h->g("dense_1/bias"), ss));
    }
    ST(ST_STN_BWD, 0, launch_stn_trunk_bwd(h->a("dflat"), h->a("p1"), h->a("p2"), reinterpret_cast<const int*>(h->a("p2arg")), h->w("conv2d_2/kernel"),
                             h->g("conv2d_1/kernel"), h->g("conv2d_1/bias"), h->g("conv2d_2/kernel"), h->g("conv2d_2/bias"), nullptr, B, h->H, h->W, st));
    side_join(h, st);
    TRY(dp_reduce_tail(h, st));                                          // data parallel: conv-stack + STN bucket, then wait for the head bucket
    return CRNN_OK;
}

// Runs `body(stream)` -- a fixed launch sequence over workspace buffers -- through a CUDA graph: the first call with a given
// (kind, batch, buffer pointers, dropout on/off) runs eagerly (lazy kernel attributes get configured), the second is captured on
// the handle's private capture stream and instantiated, every later one is a single cudaGraphLaunch on the caller's stream.
// The only per-step scalar, the dropout seed, lives in device memory ("act/seed") and is refreshed before each launch.
template <class F>
int run_graphed(crnn_handle* h, int kind, int B, const void* const (&ptrs)[6], int drop, uint64_t seed, cudaStream_t st, F&& body) {
    if (!h->use_graph || h->prof.on || !h->cap) { h->seed_ptr = nullptr; return body(st); }
    crnn_handle::StepGraph* g = nullptr;
    for (auto& e : h->graphs)
        if (e.kind == kind && e.B == B && e.drop == drop && !memcmp(e.p, ptrs, sizeof(e.p))) { g = &e; break; }
    if (!g) {
        if (h->graphs.size() >= 16) {   // callers that pass fresh buffers every step would otherwise re-capture forever
            for (auto& e : h->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
            h->graphs.clear();
        }
        crnn_handle::StepGraph e; e.kind = kind; e.B = B; e.drop = drop; e.calls = 0; e.exec = nullptr; e.launches = 0;
        memcpy(e.p, ptrs, sizeof(e.p));
        h->graphs.push_back(e); g = &h->graphs.back();
    }
    uint64_t* seed_dev = reinterpret_cast<uint64_t*>(h->a("seed"));
    if (!g->exec) {
        if (g->calls <= 0) { if (g->calls == 0) g->calls = 1; h->seed_ptr = nullptr; return body(st); }   // first call (or capture disabled: -1)
        const long long l0 = g_crnn_launches;
        CUDA_TRY(cudaStreamBeginCapture(h->cap, cudaStreamCaptureModeThreadLocal));
        h->seed_ptr = seed_dev;
        const int rc = body(h->cap);
        h->seed_ptr = nullptr;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(h->cap, &graph);
        g->launches = g_crnn_launches - l0; g_crnn_launches = l0;
        cudaError_t ie = cudaSuccess;
        if (rc == CRNN_OK && ce == cudaSuccess && graph) ie = cudaGraphInstantiate(&g->exec, graph, h->prio ? (unsigned long long)cudaGraphInstantiateFlagUseNodePriority : 0ull);
        if (graph) cudaGraphDestroy(graph);
        if (rc != CRNN_OK) { g->calls = -1; cudaGetLastError(); return rc; }
        if (ce != cudaSuccess || ie != cudaSuccess || !g->exec) {   // could not capture: stay eager for this key
            g->calls = -1; g->exec = nullptr; cudaGetLastError();
            return body(st);
        }
    }
    if (drop) TRY(launch_set_u64(seed_dev, seed, st));
    CUDA_TRY(cudaGraphLaunch(g->exec, st));
    g_crnn_launches += g->launches;
    return CRNN_OK;
}
}  // namespace

// =============================================================================================== C ABI
extern "C" {

const char* crnn_last_error(void) { return g_err; }
const char* crnn_version(void) { return "crnn_b200 0.1 (sm_100a)"; }

int crnn_workspace_bytes(const crnn_config* cfg, size_t* bytes) {
    TRY(validate(cfg));
    crnn_handle tmp; tmp.cfg = *cfg; plan(&tmp);
    *bytes = (size_t)tmp.L.cursor;
    return CRNN_OK;
}

int crnn_create(const crnn_config* cfg, void* workspace, size_t workspace_bytes, crnn_handle** out) {
    TRY(validate(cfg));
    if (!workspace || !out) { crnn_set_error("null workspace/out"); return CRNN_ERR_INVALID; }
    if (reinterpret_cast<uintptr_t>(workspace) & 255) { crnn_set_error("workspace must be 256-byte aligned"); return CRNN_ERR_INVALID; }
    crnn_handle* h = new crnn_handle(); h->cfg = *cfg; plan(h);
    if ((size_t)h->L.cursor > workspace_bytes) { crnn_set_error("workspace too small: need %lld bytes", (long long)h->L.cursor); delete h; return CRNN_ERR_NOMEM; }
    h->base = static_cast<char*>(workspace); h->bytes = workspace_bytes;
    { const char* e = getenv("CRNN_FUSE_BN_RED"); h->fuse_bn_red = !(e && e[0] == '0'); }
    { const char* e = getenv("CRNN_DW_RED"); h->dw_red = !(e && e[0] == '0'); }
    { const char* e = getenv("CRNN_BN_TAIL"); h->bn_tail = !(e && e[0] == '0'); }
    { const char* e = getenv("CRNN_DW_FUSED"); h->dw_fused = !(e && e[0] == '0'); }
    { const char* e = getenv("CRNN_FWD_FUSED"); h->fwd_fused = !(e && e[0] == '0'); }
    { const char* e = getenv("CRNN_GEMM_SIMT"); h->gemm_simt = e && e[0] == '1'; }
    { const char* e = getenv("CRNN_GRAPH"); h->use_graph = !(e && e[0] == '0'); }
    { const char* e = getenv("CRNN_OVERLAP"); h->overlap = !(e && e[0] == '0'); }
    { const char* e = getenv("CRNN_SERPENTINE"); h->serp = !(e && e[0] == '0'); }
    // CRNN_PRIO=1: the critical path (capture stream) at the highest stream priority, the side branch at the lowest; =2 the other way round;
    // the step graph is then instantiated with per-node priorities (kernel nodes inherit the priority of the stream they were captured on)
    // Measured (B200, bench.py, 40 steps, twice): 0 (none, default) 5.000 / 4.999 ms, 1 5.099 / 5.099 ms, 2 5.036 ms: the plain launch-order
    // arbitration between the two branches is the best of the three -- favouring the critical path starves the weight-gradient GEMMs until
    // the end of the backward pass, where nothing is left to overlap them with.
    { const char* e = getenv("CRNN_PRIO"); h->prio = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0; }
    int p_least = 0, p_greatest = 0;
    if (h->prio && cudaDeviceGetStreamPriorityRange(&p_least, &p_greatest) != cudaSuccess) { h->prio = 0; cudaGetLastError(); }
    const int p_side = h->prio == 1 ? p_least : p_greatest, p_cap = h->prio == 1 ? p_greatest : p_least;
    if ((h->prio ? cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, p_side) : cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking)) != cudaSuccess) { h->side = nullptr; cudaGetLastError(); }
    if ((h->prio ? cudaStreamCreateWithPriority(&h->cap, cudaStreamNonBlocking, p_cap) : cudaStreamCreateWithFlags(&h->cap, cudaStreamNonBlocking)) != cudaSuccess) { h->cap = nullptr; cudaGetLastError(); }
    *out = h;
    return CRNN_OK;
}
int crnn_destroy(crnn_handle* h) {
    if (!h) return CRNN_OK;
    for (auto& e : h->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
    for (auto e : h->evs) cudaEventDestroy(e);
    for (auto e : h->prof.pool) cudaEventDestroy(e);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->cap) cudaStreamDestroy(h->cap);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    if (h->comm && h->comm_owned && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    if (h->pin) cudaFreeHost(h->pin);
    delete h;
    return CRNN_OK;
}
int crnn_num_tensors(const crnn_handle* h) { return h ? (int)h->L.tensors.size() : 0; }
const char* crnn_tensor_name(const crnn_handle* h, int i) { return (h && i >= 0 && i < (int)h->L.tensors.size()) ? h->L.tensors[i].name.c_str() : nullptr; }
int crnn_tensor_lookup(const crnn_handle* h, const char* name, crnn_tensor_info* out) {
    if (!h || !name || !out) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    auto it = h->L.index.find(name);
    if (it == h->L.index.end()) { crnn_set_error("unknown tensor '%s'", name); return CRNN_ERR_UNKNOWN_NAME; }
    const Tensor& t = h->L.tensors[it->second];
    out->offset = t.offset; out->numel = t.numel; out->is_int = t.is_int;
    return CRNN_OK;
}

static int forward_graphed(crnn_handle* h, const float* x_dev, int B, float* softmax_dev, cudaStream_t st) {
    if (B < 1 || B > h->maxB) { crnn_set_error("batch %d outside [1,%d]", B, h->maxB); return CRNN_ERR_INVALID; }
    const void* const key[6] = {x_dev, softmax_dev, nullptr, nullptr, nullptr, nullptr};
    h->last_B = B; h->last_seed = 0; h->last_drop = false;
    return run_graphed(h, 0, B, key, 0, 0, st, [&](cudaStream_t s) -> int {
        TRY(forward(h, x_dev, B, false, 0, s));
        if (softmax_dev && softmax_dev != h->a("softmax"))
            CUDA_TRY(cudaMemcpyAsync(softmax_dev, h->a("softmax"), sizeof(float) * (size_t)B * h->T * h->V, cudaMemcpyDeviceToDevice, s));
        return CRNN_OK;
    });
}
int crnn_forward(crnn_handle* h, const float* x_dev, int B, float* softmax_dev, void* stream) {
    if (!h || !x_dev) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    return forward_graphed(h, x_dev, B, softmax_dev, static_cast<cudaStream_t>(stream));
}
int crnn_forward_host(crnn_handle* h, const float* x_host, int B, float* softmax_host, void* stream) {
    if (!h || !x_host || !softmax_host) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    if (B < 1 || B > h->maxB) { crnn_set_error("batch %d outside [1,%d]", B, h->maxB); return CRNN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemcpyAsync(h->a("x"), x_host, sizeof(float) * (size_t)B * h->H * h->W, cudaMemcpyHostToDevice, st));
    TRY(forward_graphed(h, h->a("x"), B, nullptr, st));
    CUDA_TRY(cudaMemcpyAsync(softmax_host, h->a("softmax"), sizeof(float) * (size_t)B * h->T * h->V, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return CRNN_OK;
}

int crnn_train_fwd_bwd(crnn_handle* h, const float* x_dev, const int32_t* labels_dev, const int32_t* label_len_dev,
                       const int32_t* input_len_dev, int B, float* loss_dev, uint64_t dropout_seed, void* stream) {
    if (!h || !x_dev || !labels_dev || !label_len_dev || !input_len_dev) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    if (B < 1 || B > h->maxB) { crnn_set_error("batch %d outside [1,%d]", B, h->maxB); return CRNN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* loss = loss_dev ? loss_dev : h->a("loss");
    const void* const key[6] = {x_dev, labels_dev, label_len_dev, input_len_dev, loss, nullptr};
    h->last_B = B; h->last_seed = dropout_seed; h->last_drop = dropout_seed != 0;
    return run_graphed(h, 1, B, key, dropout_seed != 0, dropout_seed, st, [&](cudaStream_t s) -> int {
        TRY(forward(h, x_dev, B, true, dropout_seed, s));
        return backward(h, x_dev, labels_dev, label_len_dev, input_len_dev, B, loss, dropout_seed, s);
    });
}

int crnn_adam_step(crnn_handle* h, float lr, float b1, float b2, float eps, float clipnorm, float grad_scale, void* stream) {
    if (!h) { crnn_set_error("null handle"); return CRNN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* ss = reinterpret_cast<double*>(h->a("sumsq"));
    CUDA_TRY(cudaMemsetAsync(ss, 0, sizeof(double), st));
    ST(ST_OPTIM, 4.0 * h->n_params * 1, launch_sumsq(h->f("arena/grads"), h->n_params, ss, st));
    const double t = (double)(h->iterations + 1);
    const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
    ST(ST_OPTIM, 4.0 * h->n_params * 7, launch_adam(h->f("arena/params"), h->f("arena/grads"), h->f("arena/opt_m"), h->f("arena/opt_v"), h->n_params, ss, clipnorm, lr_t, b1, b2, eps, grad_scale, st));
    h->iterations += 1;
    return CRNN_OK;
}
int crnn_sgd_step(crnn_handle* h, float lr, float decay, float momentum, float clipnorm, float grad_scale, void* stream) {
    if (!h) { crnn_set_error("null handle"); return CRNN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* ss = reinterpret_cast<double*>(h->a("sumsq"));
    CUDA_TRY(cudaMemsetAsync(ss, 0, sizeof(double), st));
    ST(ST_OPTIM, 4.0 * h->n_params * 1, launch_sumsq(h->f("arena/grads"), h->n_params, ss, st));
    const float lr_i = (float)((double)lr * (1.0 / (1.0 + (double)decay * (double)h->iterations)));
    ST(ST_OPTIM, 4.0 * h->n_params * 5, launch_sgd_nesterov(h->f("arena/params"), h->f("arena/grads"), h->f("arena/opt_m"), h->n_params, ss, clipnorm, lr_i, momentum, grad_scale, st));
    h->iterations += 1;
    return CRNN_OK;
}
// model.train_on_batch (train.py:201-209 through fit_generator) in ONE call on host buffers: staging into pinned memory (skipped for a
// buffer that already is pinned / registered), H2D, the graphed forward + backward (+ in-step gradient exchange under data parallel),
// the optimiser step, D2H of the per-sample losses and the CTC status, one stream synchronisation.
int crnn_train_on_batch_host(crnn_handle* h, const void* x_host, int x_is_u8, float mean, float stdv, const int32_t* labels_host,
                             const int32_t* label_len_host, const int32_t* input_len_host, int B, uint64_t dropout_seed,
                             const crnn_optimizer* opt, float grad_scale, float* losses_host, float* mean_loss, int32_t* ctc_status, void* stream) {
    if (!h || !x_host || !labels_host || !label_len_host || !input_len_host || !opt) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    if (B < 1 || B > h->maxB) { crnn_set_error("batch %d outside [1,%d]", B, h->maxB); return CRNN_ERR_INVALID; }
    if (opt->kind != CRNN_OPT_ADAM && opt->kind != CRNN_OPT_SGD) { crnn_set_error("unknown optimizer kind %d", opt->kind); return CRNN_ERR_INVALID; }
    if (x_is_u8 && !(stdv != 0.f)) { crnn_set_error("std must be non-zero"); return CRNN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t npx = (size_t)h->H * h->W, ML = (size_t)h->cfg.max_len, mB = (size_t)h->maxB;
    const size_t off_lab = (mB * npx * 4 + 255) / 256 * 256, off_ll = off_lab + mB * ML * 4, off_il = off_ll + mB * 4, off_loss = off_il + mB * 4,
                 off_st = off_loss + mB * 4, total = off_st + 64;
    if (!h->pin) { CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&h->pin), total, cudaHostAllocDefault)); h->pin_bytes = total; }
    const size_t xbytes = (size_t)B * npx * (x_is_u8 ? 1 : 4);
    const void* xsrc = x_host;
    {   // large operand: copy through the pinned buffer unless the caller's memory already is page-locked
        cudaPointerAttributes at; const cudaError_t e = cudaPointerGetAttributes(&at, x_host);
        if (e != cudaSuccess) cudaGetLastError();
        if (e != cudaSuccess || at.type != cudaMemoryTypeHost) { memcpy(h->pin, x_host, xbytes); xsrc = h->pin; }
    }
    memcpy(h->pin + off_lab, labels_host, (size_t)B * ML * 4); memcpy(h->pin + off_ll, label_len_host, (size_t)B * 4); memcpy(h->pin + off_il, input_len_host, (size_t)B * 4);
    int32_t* lab = reinterpret_cast<int32_t*>(h->f("act/labels")); int32_t* ll = reinterpret_cast<int32_t*>(h->f("act/label_len"));
    int32_t* il = reinterpret_cast<int32_t*>(h->f("act/input_len"));
    if (x_is_u8) {
        uint8_t* xu = reinterpret_cast<uint8_t*>(h->f("act/x_u8"));
        CUDA_TRY(cudaMemcpyAsync(xu, xsrc, xbytes, cudaMemcpyHostToDevice, st));
        TRY(launch_normalize_u8(xu, h->a("x"), (long long)B * npx, mean, stdv, st));
    } else {
        CUDA_TRY(cudaMemcpyAsync(h->a("x"), xsrc, xbytes, cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(lab, h->pin + off_lab, (size_t)B * ML * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(ll, h->pin + off_ll, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(il, h->pin + off_il, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    TRY(crnn_train_fwd_bwd(h, h->a("x"), lab, ll, il, B, nullptr, dropout_seed, stream));
    if (opt->kind == CRNN_OPT_ADAM) TRY(crnn_adam_step(h, opt->lr, opt->beta1, opt->beta2, opt->eps, opt->clipnorm, grad_scale, stream));
    else TRY(crnn_sgd_step(h, opt->lr, opt->decay, opt->momentum, opt->clipnorm, grad_scale, stream));
    CUDA_TRY(cudaMemcpyAsync(h->pin + off_loss, h->a("loss"), (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h->pin + off_st, h->a("status"), 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const float* lh = reinterpret_cast<const float*>(h->pin + off_loss);
    double acc = 0.0;
    for (int b = 0; b < B; ++b) acc += (double)lh[b];
    if (losses_host) memcpy(losses_host, lh, (size_t)B * 4);
    if (mean_loss) *mean_loss = (float)(acc / B);
    if (ctc_status) *ctc_status = *reinterpret_cast<const int32_t*>(h->pin + off_st);
    return CRNN_OK;
}

// ---------------------------------------------------------------- data parallel (NEW capability; the reference is single-device, train.py:111,116)
static void drop_graphs(crnn_handle* h) {
    for (auto& e : h->graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
    h->graphs.clear();
}
int crnn_nccl_unique_id(void* id128_out) {
    if (!id128_out) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    TRY(nccl_api());
    NCCL_TRY(g_nccl.GetUniqueId(id128_out));
    return CRNN_OK;
}
int crnn_comm_init_rank(crnn_handle* h, const void* id128, int nranks, int rank) {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { crnn_set_error("bad argument"); return CRNN_ERR_INVALID; }
    TRY(nccl_api());
    if (h->comm && h->comm_owned) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
    crnn_nccl_id id; memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    NCCL_TRY(g_nccl.CommInitRank(&comm, nranks, id, rank));
    h->comm = comm; h->comm_owned = true; h->comm_ranks = nranks;
    if (!h->comm_stream && cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking) != cudaSuccess) { h->comm_stream = nullptr; cudaGetLastError(); }
    drop_graphs(h);
    return CRNN_OK;
}
int crnn_set_comm(crnn_handle* h, void* nccl_comm, int nranks) {
    if (!h || nranks < 1) { crnn_set_error("bad argument"); return CRNN_ERR_INVALID; }
    if (nccl_comm) TRY(nccl_api());
    if (h->comm && h->comm_owned) g_nccl.CommDestroy(h->comm);
    h->comm = nccl_comm; h->comm_owned = false; h->comm_ranks = nccl_comm ? nranks : 1;
    if (nccl_comm && !h->comm_stream && cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking) != cudaSuccess) { h->comm_stream = nullptr; cudaGetLastError(); }
    if (!nccl_comm) h->dp_fused = false;
    drop_graphs(h);
    return CRNN_OK;
}
int crnn_set_dp_fused(crnn_handle* h, int on) {
    if (!h) { crnn_set_error("null handle"); return CRNN_ERR_INVALID; }
    if (on && !h->comm) { crnn_set_error("crnn_set_dp_fused: no communicator (crnn_comm_init_rank / crnn_set_comm first)"); return CRNN_ERR_INVALID; }
    h->dp_fused = on != 0;
    drop_graphs(h);
    return CRNN_OK;
}
int crnn_comm_ranks(const crnn_handle* h) { return (h && h->comm) ? h->comm_ranks : 1; }
int crnn_allreduce_grads(crnn_handle* h, void* nccl_comm, void* stream) {
    if (!h) { crnn_set_error("null handle"); return CRNN_ERR_INVALID; }
    void* comm = nccl_comm ? nccl_comm : h->comm;
    if (!comm) { crnn_set_error("crnn_allreduce_grads: no communicator"); return CRNN_ERR_INVALID; }
    TRY(nccl_api());
    float* g = h->f("arena/grads");
    NCCL_TRY(g_nccl.AllReduce(g, g, (size_t)h->n_params, 7, 0, comm, static_cast<cudaStream_t>(stream)));
    return CRNN_OK;
}

int crnn_get_iterations(const crnn_handle* h, int64_t* it) { if (!h || !it) return CRNN_ERR_INVALID; *it = h->iterations; return CRNN_OK; }
int crnn_set_iterations(crnn_handle* h, int64_t it) { if (!h) return CRNN_ERR_INVALID; h->iterations = it; return CRNN_OK; }
int crnn_ctc_status(crnn_handle* h, int32_t* status_host, void* stream) {
    if (!h || !status_host) return CRNN_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaMemcpyAsync(status_host, h->a("status"), sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return CRNN_OK;
}

int crnn_ctc_loss_grad(const float* probs_dev, int B, int T, int V, int t_off, const int32_t* labels_dev, int max_len,
                       const int32_t* label_len_dev, const int32_t* input_len_dev, float eps, float* loss_dev, float* grad_u_dev,
                       float* grad_logits_dev, float scale, int32_t* status_dev, void* stream) {
    if (!probs_dev || !labels_dev || !label_len_dev || !input_len_dev || !loss_dev || !status_dev) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    return launch_ctc_loss_grad(probs_dev, t_off, labels_dev, max_len, label_len_dev, input_len_dev, B, T, V, eps, loss_dev, grad_u_dev,
                                grad_logits_dev, scale, status_dev, static_cast<cudaStream_t>(stream));
}
int crnn_ctc_greedy(const float* probs_dev, const int32_t* seq_len_dev, int B, int T, int V, float eps, int32_t* out_dev, int32_t* out_len_dev,
                    float* score_dev, void* stream) {
    if (!probs_dev || !out_dev || !out_len_dev) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    return launch_ctc_greedy(probs_dev, seq_len_dev, B, T, V, eps, out_dev, out_len_dev, score_dev, static_cast<cudaStream_t>(stream));
}
int crnn_ctc_beam(const float* probs_dev, const int32_t* seq_len_dev, int B, int T, int V, float eps, int beam_width, int merge_repeated,
                  int32_t* out_dev, int32_t* out_len_dev, float* logprob_dev, void* stream) {
    if (!probs_dev || !out_dev || !out_len_dev) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    return launch_ctc_beam(probs_dev, seq_len_dev, B, T, V, eps, beam_width, merge_repeated, out_dev, out_len_dev, logprob_dev, static_cast<cudaStream_t>(stream));
}

int crnn_ctc_beam_topk(const float* probs_dev, const int32_t* seq_len_dev, int B, int T, int V, float eps, int beam_width, int merge_repeated,
                       int top_paths, int32_t* out_dev, int32_t* out_len_dev, float* logprob_dev, void* stream) {
    if (!probs_dev || !out_dev || !out_len_dev) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    return launch_ctc_beam(probs_dev, seq_len_dev, B, T, V, eps, beam_width, merge_repeated, out_dev, out_len_dev, logprob_dev, static_cast<cudaStream_t>(stream), top_paths);
}

// Host-buffer decode (DecodeCTCPred.decode, utils.py:347-357, takes host softmax rows): grow-only device scratch owned by the
// library (cudaMallocAsync's default pool hands its memory back at every synchronise: 39 MB re-allocated per call cost ~10 ms),
// and the batch is cut into chunks whose H2D copies (copy stream) overlap the decode of the previous chunk (caller's stream).
namespace {
struct DecodeScratch { void* p = nullptr; size_t cap = 0; cudaStream_t copy = nullptr; cudaEvent_t ev[8] = {}; int dev = -1; };
DecodeScratch g_dec;
}
static int decode_host(bool beam, const float* probs_host, int B, int T, int V, float eps, int beam_width, int merge_repeated,
                       int32_t* out_host, int32_t* out_len_host, float* score_host, void* stream) {
    if (!probs_host || !out_host || !out_len_host) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    if (B <= 0) return CRNN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
    const size_t np = (size_t)B * T * V, no = (size_t)B * T;
    const size_t off_out = (sizeof(float) * np + 255) & ~(size_t)255, off_len = off_out + ((sizeof(int32_t) * no + 255) & ~(size_t)255);
    const size_t off_sc = off_len + ((sizeof(int32_t) * B + 255) & ~(size_t)255), need = off_sc + sizeof(float) * B;
    DecodeScratch& D = g_dec;
    if (D.dev != dev || D.cap < need) {
        if (D.p) { cudaDeviceSynchronize(); cudaFree(D.p); D.p = nullptr; D.cap = 0; }
        CUDA_TRY(cudaMalloc(&D.p, need)); D.cap = need; D.dev = dev;
        if (!D.copy) {
            CUDA_TRY(cudaStreamCreateWithFlags(&D.copy, cudaStreamNonBlocking));
            for (auto& e : D.ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
    }
    char* base = static_cast<char*>(D.p);
    float* dp = reinterpret_cast<float*>(base); int32_t* dout = reinterpret_cast<int32_t*>(base + off_out);
    int32_t* dlen = reinterpret_cast<int32_t*>(base + off_len); float* dsc = reinterpret_cast<float*>(base + off_sc);
    const int nchunk = B >= 2048 ? 4 : (B >= 512 ? 2 : 1);
    const int per = (B + nchunk - 1) / nchunk;
    // the copy stream must not overwrite the scratch while an earlier call on `st` still reads it
    CUDA_TRY(cudaEventRecord(D.ev[7], st)); CUDA_TRY(cudaStreamWaitEvent(D.copy, D.ev[7], 0));
    int rc = CRNN_OK;
    for (int c = 0; c < nchunk && rc == CRNN_OK; ++c) {
        const int b0 = c * per, nb = (b0 + per <= B ? per : B - b0);
        if (nb <= 0) break;
        CUDA_TRY(cudaMemcpyAsync(dp + (size_t)b0 * T * V, probs_host + (size_t)b0 * T * V, sizeof(float) * (size_t)nb * T * V, cudaMemcpyHostToDevice, D.copy));
        CUDA_TRY(cudaEventRecord(D.ev[c], D.copy)); CUDA_TRY(cudaStreamWaitEvent(st, D.ev[c], 0));
        rc = beam ? launch_ctc_beam(dp + (size_t)b0 * T * V, nullptr, nb, T, V, eps, beam_width, merge_repeated, dout + (size_t)b0 * T, dlen + b0, dsc + b0, st)
                  : launch_ctc_greedy(dp + (size_t)b0 * T * V, nullptr, nb, T, V, eps, dout + (size_t)b0 * T, dlen + b0, dsc + b0, st);
    }
    if (rc == CRNN_OK) {
        CUDA_TRY(cudaMemcpyAsync(out_host, dout, sizeof(int32_t) * no, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(out_len_host, dlen, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, st));
        if (score_host) CUDA_TRY(cudaMemcpyAsync(score_host, dsc, sizeof(float) * B, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return rc;
}
int crnn_ctc_beam_host(const float* probs_host, int B, int T, int V, float eps, int beam_width, int merge_repeated,
                       int32_t* out_host, int32_t* out_len_host, float* logprob_host, void* stream) {
    return decode_host(true, probs_host, B, T, V, eps, beam_width, merge_repeated, out_host, out_len_host, logprob_host, stream);
}
int crnn_ctc_greedy_host(const float* probs_host, int B, int T, int V, float eps, int32_t* out_host, int32_t* out_len_host, float* score_host, void* stream) {
    return decode_host(false, probs_host, B, T, V, eps, 0, 1, out_host, out_len_host, score_host, stream);
}

// Stand-alone bilinear sampler (BilinearInterpolation.call, utils.py:140-232) on device buffers: x (B,H,W) fp32, theta (B,6) -> out (B,H,W);
// the same kernel the engine runs inside the step (there it writes the zero-padded buffer of the conv stack directly)
int crnn_bilinear_sample(const float* x_dev, const float* theta_dev, float* out_dev, int B, int H, int W, void* stream) {
    if (!x_dev || !theta_dev || !out_dev || B < 1 || H < 2 || W < 2) { crnn_set_error("bad argument"); return CRNN_ERR_INVALID; }
    return launch_stn_sample_fwd(x_dev, theta_dev, out_dev, B, H, W, 0, static_cast<cudaStream_t>(stream));
}

// Input pipeline (utils.py:415-416): out[i] = (float32(in[i]) - mean) / std on the device; `out` is then fed to crnn_forward / crnn_train_fwd_bwd.
int crnn_normalize_u8(const uint8_t* x_u8_dev, float* out_dev, long long n, float mean, float std, void* stream) {
    if (n < 0 || (n > 0 && (!x_u8_dev || !out_dev)) || !(std != 0.f)) { crnn_set_error("bad argument"); return CRNN_ERR_INVALID; }
    return launch_normalize_u8(x_u8_dev, out_dev, n, mean, std, static_cast<cudaStream_t>(stream));
}
// Evaluation step (utils.py:262-298): Levenshtein distance of N (prediction, truth) pairs of int32 symbol sequences padded to maxlen.
int crnn_edit_distance(const int32_t* a_dev, const int32_t* alen_dev, const int32_t* b_dev, const int32_t* blen_dev, int N, int maxlen,
                       int32_t* dist_dev, void* stream) {
    if (N < 0 || (N > 0 && (!a_dev || !alen_dev || !b_dev || !blen_dev || !dist_dev))) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    return launch_edit_distance(a_dev, alen_dev, b_dev, blen_dev, N, maxlen, dist_dev, static_cast<cudaStream_t>(stream));
}
// host buffers in, host buffer out (H2D + kernel + D2H + stream synchronisation)
int crnn_edit_distance_host(const int32_t* a, const int32_t* alen, const int32_t* b, const int32_t* blen, int N, int maxlen, int32_t* dist, void* stream) {
    if (N < 0 || (N > 0 && (!a || !alen || !b || !blen || !dist))) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    if (N == 0) return CRNN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t seq = sizeof(int32_t) * (size_t)N * maxlen, len = sizeof(int32_t) * (size_t)N;
    char* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 2 * seq + 3 * len));
    int32_t* da = reinterpret_cast<int32_t*>(d); int32_t* db = reinterpret_cast<int32_t*>(d + seq);
    int32_t* dal = reinterpret_cast<int32_t*>(d + 2 * seq); int32_t* dbl = dal + N; int32_t* dout = dbl + N;
    int rc = CRNN_OK;
    if (cudaMemcpyAsync(da, a, seq, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaMemcpyAsync(db, b, seq, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(dal, alen, len, cudaMemcpyHostToDevice, st) != cudaSuccess || cudaMemcpyAsync(dbl, blen, len, cudaMemcpyHostToDevice, st) != cudaSuccess) {
        crnn_set_error("edit_distance_host: H2D copy failed"); rc = CRNN_ERR_CUDA;
    }
    if (rc == CRNN_OK) rc = launch_edit_distance(da, dal, db, dbl, N, maxlen, dout, st);
    if (rc == CRNN_OK && (cudaMemcpyAsync(dist, dout, len, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)) {
        crnn_set_error("edit_distance_host: D2H copy failed"); rc = CRNN_ERR_CUDA;
    }
    cudaFree(d);
    return rc;
}
int crnn_gemm_tc(const float* X, int ldx, const float* W, int ldw, int w_transposed, float* out, int ldo, int M, int N, int K,
                 const float* x_scale, const float* x_shift, double* stats, float* img_scratch, void* stream) {
    if (!X || !W || !out || !img_scratch) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TRY(launch_prep_weight_images(W, ldw, N, K, w_transposed, img_scratch, st));
    return launch_xw_gemm_tc(X, ldx, img_scratch, out, ldo, M, N, K, x_scale, x_shift, stats, st);
}
// Test / debug hook: act/block{i} of the non-pooled blocks is not written by the forward pass (dwconv_fused.cu recomputes it where it is
// consumed); this fills those tensors in from the raw pointwise outputs with the BatchNorm constants and dropout seed of the LAST forward.
int crnn_debug_materialize_blocks(crnn_handle* h, void* stream) {
    if (!h) { crnn_set_error("null handle"); return CRNN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->last_B < 1) return CRNN_OK;
    int hh = h->Hp, ww = h->Wp;
    for (int i = 1; i <= 7; ++i) {
        const BlockPlan& b = kBlocks[i - 1];
        if (!h->block_live[i])
            TRY(launch_act_pool_fwd(h->a(nm("pw%d", i)), h->a(actbn(2 * i, "scale")), h->a(actbn(2 * i, "shift")), h->a(nm("block%d", i)), h->last_B, hh, ww, b.cout, b.ph, b.pw,
                                    h->last_drop ? kDropBlock : 0.f, h->last_seed, (uint32_t)i, st, nullptr, 0));
        hh /= b.ph; ww /= b.pw;
    }
    return CRNN_OK;
}
// teacher-forced backward of ONE conv block (parity tests): after a training-mode forward of batch B, takes d(block_i output)
// (numel of act/block{i}) from dout_dev, zeroes the gradient arena, runs the block's backward and copies d(block_i input) to din_dev.
int crnn_debug_block_backward(crnn_handle* h, int block, const float* dout_dev, float* din_dev, int B, uint64_t dropout_seed, void* stream) {
    if (!h || !dout_dev || !din_dev || block < 1 || block > 7 || B < 1 || B > h->maxB) { crnn_set_error("bad argument"); return CRNN_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int hh = h->Hp, ww = h->Wp;
    for (int i = 1; i < block; ++i) { hh /= kBlocks[i - 1].ph; ww /= kBlocks[i - 1].pw; }
    const BlockPlan& b = kBlocks[block - 1];
    const size_t n_out = (size_t)B * (hh / b.ph) * (ww / b.pw) * b.cout, n_in = (size_t)B * hh * ww * b.cin;
    CUDA_TRY(cudaMemsetAsync(h->f("arena/grads"), 0, sizeof(float) * (size_t)h->n_params, st));
    CUDA_TRY(cudaMemsetAsync(h->a("red"), 0, sizeof(double) * h->bn_off[15], st));
    CUDA_TRY(cudaMemcpyAsync(h->a("gA"), dout_dev, sizeof(float) * n_out, cudaMemcpyDeviceToDevice, st));
    h->ev_used = 0;
    for (int i = 0; i < 8; ++i) h->bn2_red_done[i] = 0;
    TRY(block_backward(h, block, hh, ww, h->a("gA"), h->a("gB"), B, dropout_seed != 0, dropout_seed, st));
    side_join(h, st);
    CUDA_TRY(cudaMemcpyAsync(din_dev, h->a("gB"), sizeof(float) * n_in, cudaMemcpyDeviceToDevice, st));
    return CRNN_OK;
}

int crnn_gemm_tc_dw(const float* X, int ldx, int Cin, const float* dY, int ldy, int Cout, float* dW, int ldw, int M,
                    const float* x_scale, const float* x_shift, void* stream) {
    if (!X || !dY || !dW) { crnn_set_error("null argument"); return CRNN_ERR_INVALID; }
    return launch_xty_gemm_tc(X, ldx, Cin, dY, ldy, Cout, dW, ldw, M, x_scale, x_shift, static_cast<cudaStream_t>(stream));
}
long long crnn_gemm_tc_scratch_floats(int N, int K) { return (long long)tc_weight_image_floats(N, K); }

long long crnn_launch_count(void) { return g_crnn_launches; }

int crnn_profile_enable(crnn_handle* h, int on) {
    if (!h) return CRNN_ERR_INVALID;
    h->prof.on = on != 0; h->prof.recs.clear(); h->prof.used = 0;
    return CRNN_OK;
}
int crnn_profile_num_stages(void) { return ST_COUNT; }
const char* crnn_profile_stage_name(int stage) { return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : nullptr; }
// synchronises the device, aggregates the event pairs recorded since the last enable/report, then clears them
int crnn_profile_report(crnn_handle* h, double* ms, double* work, long long* launches) {
    return crnn_profile_report2(h, ms, work, launches, nullptr, nullptr, nullptr);
}
int crnn_profile_num_families(void) { return CRNN_FAM_COUNT; }
const char* crnn_profile_family_name(int fam) { return (fam >= 0 && fam < CRNN_FAM_COUNT) ? kFamilyNames[fam] : nullptr; }
// per stage AND per (stage, kernel family): fam_* are [num_stages * num_families] row-major (stage, family); family 0 = kernels not tracked
int crnn_profile_report2(crnn_handle* h, double* ms, double* work, long long* launches, double* fam_ms, double* fam_work, long long* fam_launches) {
    if (!h || !ms || !work || !launches) return CRNN_ERR_INVALID;
    CUDA_TRY(cudaDeviceSynchronize());
    for (int i = 0; i < ST_COUNT; ++i) { ms[i] = 0; work[i] = 0; launches[i] = 0; }
    if (fam_ms) for (int i = 0; i < ST_COUNT * CRNN_FAM_COUNT; ++i) { fam_ms[i] = 0; if (fam_work) fam_work[i] = 0; if (fam_launches) fam_launches[i] = 0; }
    for (auto& r : h->prof.recs) {
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.stage] += t; work[r.stage] += r.work; launches[r.stage] += r.launches;
        if (fam_ms) {
            const int k = r.stage * CRNN_FAM_COUNT + r.fam;
            fam_ms[k] += t; if (fam_work) fam_work[k] += r.work; if (fam_launches) fam_launches[k] += r.launches;
        }
    }
    h->prof.recs.clear(); h->prof.used = 0;
    return CRNN_OK;
}

int crnn_gemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
              const float* a_scale, const float* a_shift, const float* bias, int relu, int split_k, void* stream) {
    GemmArgs g; g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.transA = transA; g.transB = transB; g.a_scale = a_scale; g.a_shift = a_shift; g.bias = bias; g.relu = relu; g.split_k = split_k;
    return launch_gemm_simt(g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
