// ctc.cu -- CTC loss/gradient, greedy decode and beam-search decode for sm_100a.
//
// Replaces (reference gasparian/CRNN-OCR-lite):
//   utils.py:98-103   ctc_lambda_func -> K.ctc_batch_cost   (TF 1.8 CTCLoss, a CPU-only op)
//   utils.py:347-357  DecodeCTCPred.decode -> K.ctc_decode  (TF 1.8 CTCBeamSearchDecoder, CPU-only,
//                                                            README.md:75 ">95 % of wall time")
// Semantics restated in SURVEY.md Appendix A.1 / A.2; blank = V-1; all log-space arithmetic in fp32 with the
// same LogSumExp form as TF (max + log1p(exp(min-max))).
#include "common.cuh"

// =================================================================================================
// CTC loss + gradient.  One CTA (128 threads) per sequence:
//   phase 0  all threads : u=log(p+eps), re-softmax (TF does), log y          -> smem logy[T'][V]
//   phase 1  warp 0      : alpha sweep (lanes over the 2L+1 extended-label states) -> smem alpha[T'][S]
//            warp 1      : beta sweep  (independent of alpha)                     -> smem beta[T'][S]
//   phase 2  all threads : per (t,k) gradient  y - exp(LSE_{s:l'_s=k}(alpha+beta) - log p)
//                          optionally chained through u=log(p+eps) and the dense2 softmax to d/d logits.
// =================================================================================================
#define CTC_MAX_NS 8   // states per lane  -> S <= 256, L <= 127

__global__ void __launch_bounds__(128)
ctc_loss_grad_kernel(const float* __restrict__ probs,   // (B, T, V) softmax output, full T
                     int t_off,                          // frames dropped at the front (reference: 2)
                     const int* __restrict__ labels, int maxL,
                     const int* __restrict__ label_len, const int* __restrict__ input_len,
                     int B, int T, int V, float eps,
                     float* __restrict__ loss,           // (B)
                     float* __restrict__ grad_u,         // (B, T-t_off, V) or null : d loss_b / d u
                     float* __restrict__ grad_logits,    // (B, T, V) or null : scale * d loss_b / d dense2-logits
                     float scale, int* __restrict__ status)
{
    extern __shared__ float sm[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Tp = T - t_off;
    const int L = label_len[b];
    const int Tb = input_len[b];
    const int S = 2 * L + 1;
    const int Smax = 2 * maxL + 1;
    const int blank = V - 1;
    float* logy = sm;                      // Tp*V
    float* alpha = logy + (size_t)Tp * V;  // Tp*Smax
    float* beta = alpha + (size_t)Tp * Smax;
    int* lp = (int*)(beta + (size_t)Tp * Smax);  // Smax
    __shared__ float s_logp;
    __shared__ int s_bad;

    const float* pb = probs + ((size_t)b * T + t_off) * V;

    if (tid == 0) {
        int rep = 0;
        for (int i = 1; i < L; ++i) rep += labels[b * maxL + i] == labels[b * maxL + i - 1];
        s_bad = (Tb > Tp) || (L + rep > Tb) || (L > maxL) || (L < 0);
    }
    for (int s = tid; s < S && s < Smax; s += blockDim.x) lp[s] = (s & 1) ? labels[b * maxL + (s >> 1)] : blank;
    __syncthreads();
    if (s_bad) {   // TF: InvalidArgument "Not enough time for target transition sequence"
        if (tid == 0) { loss[b] = INFINITY; atomicMin(status, -(b + 1)); }
        if (grad_u) for (int i = tid; i < Tp * V; i += blockDim.x) grad_u[(size_t)b * Tp * V + i] = 0.f;
        if (grad_logits) for (int i = tid; i < T * V; i += blockDim.x) grad_logits[(size_t)b * T * V + i] = 0.f;
        return;
    }

    // ---- phase 0: log y (thread per frame, serial over classes: same summation order as the oracle) ----
    for (int t = tid; t < Tb; t += blockDim.x) {
        const float* p = pb + (size_t)t * V;
        float* ly = logy + (size_t)t * V;
        float mx = -INFINITY;
        for (int k = 0; k < V; ++k) { float u = logf(p[k] + eps); ly[k] = u; mx = fmaxf(mx, u); }
        float sum = 0.f;
        for (int k = 0; k < V; ++k) sum += expf(ly[k] - mx);
        for (int k = 0; k < V; ++k) ly[k] = logf(expf(ly[k] - mx) / sum);
    }
    for (int i = tid; i < Tb * S; i += blockDim.x) {
        int t = i / S, s = i - t * S;
        alpha[t * Smax + s] = NEG_INF; beta[t * Smax + s] = NEG_INF;
    }
    __syncthreads();

    // ---- phase 1: alpha (warp 0) and beta (warp 1) ----
    if (warp == 0) {
        if (lane == 0) { alpha[0] = logy[blank]; if (S > 1) alpha[1] = logy[lp[1]]; }
        __syncwarp();
        for (int t = 1; t < Tb; ++t) {
            int lo = S - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > S) hi = S;
            const float* ap = alpha + (size_t)(t - 1) * Smax;
            float* an = alpha + (size_t)t * Smax;
            const float* ly = logy + (size_t)t * V;
            for (int s = lo + lane; s < hi; s += 32) {
                float a = ap[s];
                if (s > 0) a = lse2(a, ap[s - 1]);
                int l = lp[s];
                if (s > 1 && l != blank && l != lp[s - 2]) a = lse2(a, ap[s - 2]);
                an[s] = ly[l] + a;
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) { beta[(size_t)(Tb - 1) * Smax + S - 1] = 0.f; if (S > 1) beta[(size_t)(Tb - 1) * Smax + S - 2] = 0.f; }
        __syncwarp();
        for (int t = Tb - 2; t >= 0; --t) {
            int lo = S - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > S) hi = S;
            const float* bn = beta + (size_t)(t + 1) * Smax;
            float* bc = beta + (size_t)t * Smax;
            const float* ly = logy + (size_t)(t + 1) * V;
            for (int s = lo + lane; s < hi; s += 32) {
                int l = lp[s];
                float v = bn[s] + ly[l];
                if (s + 1 < S) v = lse2(v, bn[s + 1] + ly[lp[s + 1]]);
                if (s + 2 < S) { int l2 = lp[s + 2]; if (l2 != blank && l2 != l) v = lse2(v, bn[s + 2] + ly[l2]); }
                bc[s] = v;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid == 0) {
        float lpv = NEG_INF;
        for (int s = 0; s < S; ++s) lpv = lse2(lpv, alpha[s] + beta[s]);
        s_logp = lpv;
        loss[b] = -lpv;
    }
    __syncthreads();
    const float logp = s_logp;

    // ---- phase 2a: g_u[t][k] overwrites logy[t][k]  (thread per (t,k)) ----
    for (int i = tid; i < Tb * V; i += blockDim.x) {
        int t = i / V, k = i - t * V;
        float acc = NEG_INF;
        const float* a = alpha + (size_t)t * Smax;
        const float* be = beta + (size_t)t * Smax;
        if (k == blank) { for (int s = 0; s < S; s += 2) acc = lse2(acc, a[s] + be[s]); }
        else            { for (int s = 1; s < S; s += 2) if (lp[s] == k) acc = lse2(acc, a[s] + be[s]); }
        float y = expf(logy[i]);
        float g = (acc == NEG_INF || logp == NEG_INF) ? y : y - expf(acc - logp);
        logy[i] = g;
    }
    __syncthreads();
    if (grad_u) {
        float* g = grad_u + (size_t)b * Tp * V;
        for (int i = tid; i < Tp * V; i += blockDim.x) g[i] = (i < Tb * V) ? logy[i] : 0.f;
    }
    // ---- phase 2b: chain to the dense2 logits: u=log(p+eps), p=softmax(z)  (thread per frame) ----
    if (grad_logits) {
        float* gz = grad_logits + (size_t)b * T * V;
        for (int i = tid; i < t_off * V; i += blockDim.x) gz[i] = 0.f;          // frames 0,1: no gradient (utils.py:102)
        for (int t = tid; t < Tp; t += blockDim.x) {
            float* o = gz + (size_t)(t + t_off) * V;
            if (t >= Tb) { for (int k = 0; k < V; ++k) o[k] = 0.f; continue; }
            const float* p = pb + (size_t)t * V;
            const float* g = logy + (size_t)t * V;
            float dot = 0.f;
            for (int k = 0; k < V; ++k) dot += (g[k] / (p[k] + eps)) * p[k];
            for (int k = 0; k < V; ++k) o[k] = scale * p[k] * (g[k] / (p[k] + eps) - dot);
        }
    }
}

// =================================================================================================
// Greedy decode: one warp per sequence.  argmax over classes of u=log(p+eps) (first max wins), merge
// repeats, drop blanks.  out (B,T) padded with -1.
// =================================================================================================
__global__ void ctc_greedy_kernel(const float* __restrict__ probs, const int* __restrict__ seq_len,
                                  int B, int T, int V, float eps,
                                  int* __restrict__ out, int* __restrict__ out_len, float* __restrict__ score)
{
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * warps + (threadIdx.x >> 5);
    if (b >= B) return;
    const int Tb = seq_len ? seq_len[b] : T;
    const int blank = V - 1;
    int n = 0, prev = -1;
    float acc = 0.f;
    for (int t = 0; t < Tb; ++t) {
        const float* p = probs + ((size_t)b * T + t) * V;
        float bv = -INFINITY; int bk = 0x7fffffff;
        for (int k = lane; k < V; k += 32) { float u = logf(p[k] + eps); if (u > bv) { bv = u; bk = k; } }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o); int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
        }
        acc -= bv;
        if (bk != blank && bk != prev) { if (lane == 0) out[(size_t)b * T + n] = bk; ++n; }
        prev = bk;
    }
    for (int t = n + lane; t < T; t += 32) out[(size_t)b * T + t] = -1;
    if (lane == 0) { out_len[b] = n; if (score) score[b] = acc; }
}

// =================================================================================================
// Beam search: one warp per sequence, prefix trie + beam state in shared memory, the sorted list of leaves
// distributed over the lanes (lane i holds the i-th best leaf).
//
// This reproduces TF 1.8's CTCBeamSearchDecoder::Step *including its sequential side effects* -- it is NOT a plain
// "score all W*(V-1) children, keep the global top-W":
//   phase 1  every survivor gets the standard prefix-beam update (parent's t-1 probabilities iff the parent is in
//            the beam);
//   phase 2  parents are visited best-first (beam order).  A parent is skipped if its t-1 total does not beat the
//            current bottom leaf, or if it was BLOCKED: TF resets the t-1 probabilities ("Deactivate child") of a
//            survivor that has already been evicted from the leaves when its own parent reaches its label in the
//            children loop, so that survivor never expands children in this step.  For each parent the children
//            that are not already survivors are inserted into the leaves iff they beat the bottom (strictly, when the
//            list is full); insertion order does not matter for the resulting set, only for the blocking test, which
//            is evaluated in closed form: survivor c (child k_c of parent b) is blocked iff
//            rank(c in leaves) + #{eligible children of b with label < k_c and total > total(c)} >= W.
// The trie gives every prefix one node id for its whole life, so a prefix that leaves the beam and re-enters later as
// a child re-links to children of it that stayed -- what TF's persistent BeamEntry tree does.
// (oracle/beam_reference_py.py states the same formulation in Python; tests check both against the literal restatement.)
// =================================================================================================
struct TrieNode { short parent, label, first_child, next_sib; };

#define BEAM_MAX_W 32

__global__ void ctc_beam_kernel(const float* __restrict__ probs, const int* __restrict__ seq_len,
                                int B, int T, int V, float eps, int W, int merge_repeated,
                                int* __restrict__ out, int* __restrict__ out_len, float* __restrict__ logprob,
                                int smem_per_warp_bytes)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int b = blockIdx.x * warps + wib;
    if (b >= B) return;
    const unsigned FULL = 0xffffffffu;
    const int blank = V - 1, NC = V - 1;
    const int Tb = seq_len ? seq_len[b] : T;
    const int NOID = 0x7fffffff;

    unsigned char* base = smraw + (size_t)wib * smem_per_warp_bytes;
    float* in_ = (float*)base;                                           // V (padded)
    TrieNode* nodes = (TrieNode*)(in_ + ((V + 3) & ~3));                 // 1 + W*T
    int* bnode = (int*)(nodes + (((size_t)1 + (size_t)W * T + 1) & ~1)); // [2][32]
    float* bpb = (float*)(bnode + 64);                                   // [2][32] blank
    float* bpl = bpb + 64;                                               // [2][32] label
    float* bpt = bpl + 64;                                               // [2][32] total

    int nn = 1;        // nodes in the trie (uniform)
    int nb = 1;        // beam entries (uniform); slots are in descending-total order
    int cur = 0;
    if (lane == 0) {
        nodes[0].parent = -1; nodes[0].label = -1; nodes[0].first_child = -1; nodes[0].next_sib = -1;
        bnode[0] = 0; bpb[0] = 0.f; bpl[0] = NEG_INF; bpt[0] = 0.f;
    }
    __syncwarp();

    for (int t = 0; t < Tb; ++t) {
        const float* p = probs + ((size_t)b * T + t) * V;
        // (a) normalised step input: u - max(u)   (TF 1.8: max-subtraction only)
        float mx = -INFINITY;
        for (int k = lane; k < V; k += 32) { float u = logf(p[k] + eps); in_[k] = u; mx = fmaxf(mx, u); }
        mx = warp_max(mx);
        __syncwarp();
        for (int k = lane; k < V; k += 32) in_[k] -= mx;
        __syncwarp();

        const int* on = bnode + cur * 32; const float* opb = bpb + cur * 32; const float* opl = bpl + cur * 32; const float* opt = bpt + cur * 32;
        // ---- phase 1: survivors (lane e < nb)
        float s_nb = NEG_INF, s_nl = NEG_INF, s_nt = NEG_INF;
        int s_node = -1, s_q = -1, s_k = -1;
        if (lane < nb) {
            s_node = on[lane];
            const TrieNode nd = nodes[s_node];
            float nl = opl[lane];
            if (nd.parent >= 0) {
                for (int i = 0; i < nb; ++i) if (on[i] == nd.parent) s_q = i;
                if (s_q >= 0) {
                    float prev = (nd.label == nodes[nd.parent].label) ? opb[s_q] : opt[s_q];
                    nl = lse2(nl, prev);
                }
                nl += in_[nd.label];
                s_k = nd.label;
            }
            s_nl = nl;
            s_nb = opt[lane] + in_[blank];
            s_nt = lse2(s_nb, nl);
        }
        // ---- leaves list L over the lanes: (Lv, Lid); id<0: survivor -1-e, id>=0: child q*NC+k
        float Lv = NEG_INF; int Lid = NOID; int nL = 0;
        auto insert = [&](float v, int id) {     // warp-uniform call; keeps L sorted (descending), capacity W
            int pos = __popc(__ballot_sync(FULL, Lv > v));
            float upv = __shfl_up_sync(FULL, Lv, 1); int upi = __shfl_up_sync(FULL, Lid, 1);
            if (lane > pos) { Lv = upv; Lid = upi; } else if (lane == pos) { Lv = v; Lid = id; }
            if (lane >= W) { Lv = NEG_INF; Lid = NOID; }
            if (nL < W) ++nL;
        };
        for (int e = 0; e < nb; ++e) insert(__shfl_sync(FULL, s_nt, e), -1 - e);

        // ---- phase 2: parents in beam order
        bool blocked = false;     // lane e: survivor e lost its t-1 probabilities (cannot expand children)
        for (int r = 0; r < nb; ++r) {
            const bool r_blocked = __shfl_sync(FULL, (int)blocked, r) != 0;
            const float ot = opt[r], ob = opb[r];
            float bottom = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
            if (r_blocked || !(ot > bottom)) continue;     // is_candidate(b->oldp)
            const int lab_r = nodes[on[r]].label;
            const unsigned km = __ballot_sync(FULL, lane < nb && s_q == r);   // survivors that are children of r
            // blocking test for every such survivor
            for (unsigned m = km; m; m &= m - 1) {
                const int c = __ffs(m) - 1;
                const int k_c = __shfl_sync(FULL, s_k, c);
                const float s_c = __shfl_sync(FULL, s_nt, c);
                const unsigned pm = __ballot_sync(FULL, Lid == -1 - c);
                bool blk;
                if (pm == 0) blk = true;                     // already evicted from the leaves
                else {
                    int cnt = __ffs(pm) - 1;                 // rank of c in L
                    for (int k0 = 0; k0 < k_c; k0 += 32) {
                        const int k = k0 + lane;
                        bool ok = k < k_c;
                        for (unsigned m2 = km; m2; m2 &= m2 - 1) { const int kk = __shfl_sync(FULL, s_k, __ffs(m2) - 1); if (k == kk) ok = false; }
                        float x = ok ? in_[k] + ((k == lab_r) ? ob : ot) : NEG_INF;
                        cnt += __popc(__ballot_sync(FULL, x > s_c));
                    }
                    blk = cnt >= W;
                }
                if (blk && lane == c) blocked = true;
            }
            // insert the eligible children of r
            for (int k0 = 0; k0 < NC; k0 += 32) {
                const int k = k0 + lane;
                float x = (k < NC) ? in_[k] + ((k == lab_r) ? ob : ot) : NEG_INF;
                for (unsigned m2 = km; m2; m2 &= m2 - 1) { int kk = __shfl_sync(FULL, s_k, __ffs(m2) - 1); if (k == kk) x = NEG_INF; }
                bottom = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
                for (unsigned m = __ballot_sync(FULL, x > bottom); m; m &= m - 1) {
                    const int src = __ffs(m) - 1;
                    const float v = __shfl_sync(FULL, x, src);
                    bottom = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
                    if (v > bottom) insert(v, r * NC + k0 + src);
                }
            }
        }

        // ---- build the new beam (slot i = i-th best leaf) in the other buffer
        const int nxt = cur ^ 1;
        int* nnode = bnode + nxt * 32; float* npb = bpb + nxt * 32; float* npl = bpl + nxt * 32; float* npt = bpt + nxt * 32;
        const int esrc = (Lid < 0) ? (-1 - Lid) : 0;
        const int g_node = __shfl_sync(FULL, s_node, esrc);
        const float g_nb = __shfl_sync(FULL, s_nb, esrc), g_nl = __shfl_sync(FULL, s_nl, esrc), g_nt = __shfl_sync(FULL, s_nt, esrc);
        if (lane < nL) {
            if (Lid < 0) { nnode[lane] = g_node; npb[lane] = g_nb; npl[lane] = g_nl; npt[lane] = g_nt; }
            else { npb[lane] = NEG_INF; npl[lane] = Lv; npt[lane] = Lv; }
        }
        for (int i = 0; i < nL; ++i) {     // trie find-or-create for the new children (lane 0, sequential)
            const int id = __shfl_sync(FULL, Lid, i);
            if (id >= 0 && lane == 0) {
                const int q = id / NC, k = id - q * NC;
                const int par = on[q];
                int c = nodes[par].first_child;
                while (c >= 0 && nodes[c].label != k) c = nodes[c].next_sib;
                if (c < 0) {
                    c = nn;
                    nodes[c].parent = (short)par; nodes[c].label = (short)k;
                    nodes[c].first_child = -1; nodes[c].next_sib = nodes[par].first_child;
                    nodes[par].first_child = (short)c;
                    ++nn;
                }
                nnode[i] = c;
            }
        }
        nn = __shfl_sync(FULL, nn, 0);
        nb = nL;
        cur = nxt;
        __syncwarp();
    }

    // top path = best total = slot 0 (slots are sorted); before the first step it is the root
    const float* fpt = bpt + cur * 32; const int* fn = bnode + cur * 32;
    int n = 0;
    if (lane == 0) {
        // TF BeamEntry::LabelSeq: walk leaf -> root, drop a label equal to the previously visited one
        int prev = -1;
        for (int c = fn[0]; nodes[c].parent >= 0; c = nodes[c].parent) {
            int l = nodes[c].label;
            if (!merge_repeated || l != prev) ++n;
            prev = l;
        }
        int i = n; prev = -1;
        for (int c = fn[0]; nodes[c].parent >= 0; c = nodes[c].parent) {
            int l = nodes[c].label;
            if (!merge_repeated || l != prev) out[(size_t)b * T + (--i)] = l;
            prev = l;
        }
        out_len[b] = n;
        if (logprob) logprob[b] = fpt[0];
    }
    n = __shfl_sync(FULL, n, 0);
    for (int t = n + lane; t < T; t += 32) out[(size_t)b * T + t] = -1;
}

// =================================================================================================
// host launchers
// =================================================================================================
size_t ctc_loss_smem_bytes(int T, int t_off, int V, int maxL) {
    int Tp = T - t_off, Smax = 2 * maxL + 1;
    return sizeof(float) * ((size_t)Tp * V + 2 * (size_t)Tp * Smax) + sizeof(int) * Smax;
}

int launch_ctc_loss_grad(const float* probs, int t_off, const int* labels, int maxL, const int* label_len,
                         const int* input_len, int B, int T, int V, float eps, float* loss, float* grad_u,
                         float* grad_logits, float scale, int* status, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    if (T - t_off <= 0 || V < 2 || maxL < 0) { crnn_set_error("ctc_loss: bad shape"); return CRNN_ERR_INVALID; }
    size_t smem = ctc_loss_smem_bytes(T, t_off, V, maxL);
    if (smem > 227 * 1024) { crnn_set_error("ctc_loss: T'*(V+2S) too large for shared memory (%zu B)", smem); return CRNN_ERR_INVALID; }
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        CUDA_TRY(cudaFuncSetAttribute(ctc_loss_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
    ctc_loss_grad_kernel<<<B, 128, smem, st>>>(probs, t_off, labels, maxL, label_len, input_len, B, T, V, eps,
                                               loss, grad_u, grad_logits, scale, status);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_ctc_greedy(const float* probs, const int* seq_len, int B, int T, int V, float eps,
                      int* out, int* out_len, float* score, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    const int warps = 4;
    ctc_greedy_kernel<<<ceil_div(B, warps), warps * 32, 0, st>>>(probs, seq_len, B, T, V, eps, out, out_len, score);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_ctc_beam(const float* probs, const int* seq_len, int B, int T, int V, float eps, int W, int merge_repeated,
                    int* out, int* out_len, float* logprob, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    if (W < 1 || W > BEAM_MAX_W) { crnn_set_error("ctc_beam: beam width %d not in [1,%d]", W, BEAM_MAX_W); return CRNN_ERR_INVALID; }
    if ((size_t)1 + (size_t)W * T > 32000) { crnn_set_error("ctc_beam: W*T too large"); return CRNN_ERR_INVALID; }
    size_t per = sizeof(float) * ((V + 3) & ~3) + sizeof(TrieNode) * (((size_t)1 + (size_t)W * T + 1) & ~1)
               + sizeof(int) * 64 + sizeof(float) * 64 * 3;
    per = (per + 15) & ~(size_t)15;
    int warps = 8;
    while (warps > 1 && per * warps > 200 * 1024) warps >>= 1;
    size_t smem = per * warps;
    if (smem > 227 * 1024) { crnn_set_error("ctc_beam: per-sequence state %zu B exceeds shared memory", per); return CRNN_ERR_INVALID; }
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        CUDA_TRY(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    ctc_beam_kernel<<<ceil_div(B, warps), warps * 32, smem, st>>>(probs, seq_len, B, T, V, eps, W, merge_repeated,
                                                                  out, out_len, logprob, (int)per);
    LAUNCH_CHECK();
    return CRNN_OK;
}
