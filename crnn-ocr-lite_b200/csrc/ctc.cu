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
// CTC loss + gradient.  One CTA (128 threads = 4 warps) per sequence, everything in shared memory:
//   phase 0  warp per frame : u=log(p+eps), re-softmax (TF does), log y (lanes over classes)        -> smem logy[T'][V]
//   phase 1  warp 0         : alpha sweep, warp 1: beta sweep (independent).  A lane OWNS the extended-label states s = lane + 32 j
//                             in registers; neighbours s-1, s-2 (s+1, s+2) arrive by shuffles, so a time step is one dependent
//                             chain of two LogSumExp per lane with no shared-memory round trip.              -> smem alpha/beta[T'][S]
//   phase 2  warp per frame : per-class LSE_{s: l'_s = k}(alpha+beta): states grouped by label with __match_any (lanes over label
//                             positions, the blank states by a warp max/sum reduction), then lanes over classes form
//                             y - exp(acc - log p) and chain it through u=log(p+eps) and the dense2 softmax to d/d logits.
// (ncu r1f on the previous version -- thread per (t,k) looping over all states: 188 us per launch, the 1/38 of the items that were
//  blanks serialised 24 LogSumExp each and the divergence made every warp iteration that slow.)
// =================================================================================================
#define CTC_MAX_NS 8   // states per lane  -> S <= 256, L <= 127
#define CTC_WARPS 8

// LogSumExp on the sequential alpha/beta chains: max + log(1 + exp(min - max)) with the hardware ex2/lg2 approximations
// (|error| <= ~3e-7 per call; the chain is 2 calls per time step, so <= ~4e-5 on log p(l|x) ~ 1e2 -- inside the stated 1e-5 relative
// tolerance).  The accurate expf/log1pf pair is ~6x longer and this chain is pure latency (one warp, ncu r1g: 11 cycles / instruction).
__device__ __forceinline__ float lse2_fast(float a, float b) {
    const float mx = fmaxf(a, b), mn = fminf(a, b);
    if (mx == NEG_INF) return NEG_INF;
    return mx + __logf(1.f + __expf(mn - mx));
}

template <int NS>
__global__ void __launch_bounds__(CTC_WARPS * 32)
ctc_loss_grad_kernel(const float* __restrict__ probs,   // (B, T, V) softmax output, full T
                     int t_off,                          // frames dropped at the front (reference: 2)
                     const int* __restrict__ labels, int maxL,
                     const int* __restrict__ label_len, const int* __restrict__ input_len,
                     int B, int T, int V, float eps,
                     float* __restrict__ loss,           // (B)
                     float* __restrict__ grad_u,         // (B, T-t_off, V) or null : d loss_b / d u
                     float* __restrict__ grad_logits,    // (B, T, V) or null : scale * d loss_b / d dense2-logits
                     float scale, int* __restrict__ status)
{ pdl_enter();
    extern __shared__ float sm[];
    constexpr unsigned FULL = 0xffffffffu;
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
    float* acc = beta + (size_t)Tp * Smax; // CTC_WARPS x V
    int* lp = (int*)(acc + CTC_WARPS * (size_t)V); // Smax
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

    // ---- phase 0: log y, warp per frame, lanes over classes ----
    for (int t = warp; t < Tb; t += CTC_WARPS) {
        const float* p = pb + (size_t)t * V;
        float* ly = logy + (size_t)t * V;
        float mx = -INFINITY;
        for (int k = lane; k < V; k += 32) { const float u = logf(__ldg(p + k) + eps); ly[k] = u; mx = fmaxf(mx, u); }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int k = lane; k < V; k += 32) sum += expf(ly[k] - mx);
        sum = warp_sum(sum);
        const float lsum = logf(sum);
        for (int k = lane; k < V; k += 32) ly[k] = (ly[k] - mx) - lsum;      // = log(exp(u - mx) / sum)
    }
    __syncthreads();

    // ---- phase 1: alpha (warp 0) and beta (warp 1), states in registers ----
    const int nj = (S + 31) >> 5;
    if (warp == 0) {
        float a[NS]; int lab[NS]; bool skip[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int s = lane + 32 * j;
            const bool ok = s < S;
            lab[j] = ok ? lp[s] : blank;
            skip[j] = ok && s > 1 && lab[j] != blank && lab[j] != lp[s - 2];
            a[j] = NEG_INF;
        }
        if (lane == 0) a[0] = logy[blank];
        if (lane == 1 && S > 1) a[0] = logy[lab[0]];
#pragma unroll
        for (int j = 0; j < NS; ++j) { const int s = lane + 32 * j; if (j < nj && s < S) alpha[s] = a[j]; }
        for (int t = 1; t < Tb; ++t) {
            int lo = S - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > S) hi = S;
            const float* ly = logy + (size_t)t * V;
            float* an = alpha + (size_t)t * Smax;
            float c31 = NEG_INF, c30 = NEG_INF;          // previous 32-state block's (old) values at lanes 31 / 30
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                if (j < nj) {
                    const int s = lane + 32 * j;
                    const float cur = a[j];
                    float up1 = __shfl_up_sync(FULL, cur, 1), up2 = __shfl_up_sync(FULL, cur, 2);
                    if (lane == 0) { up1 = c31; up2 = c30; } else if (lane == 1) up2 = c31;
                    c31 = __shfl_sync(FULL, cur, 31); c30 = __shfl_sync(FULL, cur, 30);
                    float v = cur;
                    if (s > 0) v = lse2_fast(v, up1);
                    if (skip[j]) v = lse2_fast(v, up2);
                    v = (s >= lo && s < hi) ? ly[lab[j]] + v : NEG_INF;
                    a[j] = v;
                    if (s < S) an[s] = v;
                }
            }
        }
    } else if (warp == 1) {
        float bt[NS]; int lab[NS]; bool skip2[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int s = lane + 32 * j;
            const bool ok = s < S;
            lab[j] = ok ? lp[s] : blank;
            skip2[j] = (s + 2 < S) && lp[s + 2] != blank && lp[s + 2] != lab[j];
            bt[j] = (ok && s >= S - 2) ? 0.f : NEG_INF;     // beta[Tb-1][S-1] = beta[Tb-1][S-2] = 0
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) { const int s = lane + 32 * j; if (j < nj && s < S) beta[(size_t)(Tb - 1) * Smax + s] = bt[j]; }
        for (int t = Tb - 2; t >= 0; --t) {
            int lo = S - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > S) hi = S;
            const float* ly = logy + (size_t)(t + 1) * V;
            float* bc = beta + (size_t)t * Smax;
            float c[NS + 1];
#pragma unroll
            for (int j = 0; j < NS; ++j) { const int s = lane + 32 * j; c[j] = (j < nj && s < S) ? bt[j] + ly[lab[j]] : NEG_INF; }
            c[NS] = NEG_INF;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                if (j < nj) {
                    const int s = lane + 32 * j;
                    float dn1 = __shfl_down_sync(FULL, c[j], 1), dn2 = __shfl_down_sync(FULL, c[j], 2);
                    const float n0 = __shfl_sync(FULL, c[j + 1], 0), n1 = __shfl_sync(FULL, c[j + 1], 1);
                    if (lane == 31) { dn1 = n0; dn2 = n1; } else if (lane == 30) dn2 = n0;
                    float v = c[j];
                    if (s + 1 < S) v = lse2_fast(v, dn1);
                    if (skip2[j]) v = lse2_fast(v, dn2);
                    v = (s >= lo && s < hi) ? v : NEG_INF;
                    bt[j] = v;
                    if (s < S) bc[s] = v;
                }
            }
        }
    }
    __syncthreads();
    if (warp == 0) {   // log p(l|x) = LSE_s(alpha_0(s) + beta_0(s))
        float m = NEG_INF;
        for (int s = lane; s < S; s += 32) m = fmaxf(m, alpha[s] + beta[s]);
        m = warp_max(m);
        float sum = 0.f;
        if (m != NEG_INF) for (int s = lane; s < S; s += 32) sum += expf(alpha[s] + beta[s] - m);
        sum = warp_sum(sum);
        if (lane == 0) { const float lpv = (m == NEG_INF) ? NEG_INF : m + logf(sum); s_logp = lpv; loss[b] = -lpv; }
    }
    if (grad_logits) for (int i = tid; i < t_off * V; i += blockDim.x) grad_logits[(size_t)b * T * V + i] = 0.f;   // frames 0,1: no gradient (utils.py:102)
    __syncthreads();
    const float logp = s_logp;

    // ---- phase 2: warp per frame ----
    float* accw = acc + (size_t)warp * V;
    for (int t = warp; t < Tp; t += CTC_WARPS) {
        float* gu = grad_u ? grad_u + ((size_t)b * Tp + t) * V : nullptr;
        float* gz = grad_logits ? grad_logits + ((size_t)b * T + t + t_off) * V : nullptr;
        if (t >= Tb) {
            for (int k = lane; k < V; k += 32) { if (gu) gu[k] = 0.f; if (gz) gz[k] = 0.f; }
            continue;
        }
        const float* a = alpha + (size_t)t * Smax;
        const float* be = beta + (size_t)t * Smax;
        for (int k = lane; k < V; k += 32) accw[k] = NEG_INF;
        // blank states s = 2i, i = 0..L
        float mB = NEG_INF;
        for (int i = lane; i <= L; i += 32) mB = fmaxf(mB, a[2 * i] + be[2 * i]);
        mB = warp_max(mB);
        float sB = 0.f;
        if (mB != NEG_INF) for (int i = lane; i <= L; i += 32) sB += expf(a[2 * i] + be[2 * i] - mB);
        sB = warp_sum(sB);
        __syncwarp();
        // label states s = 2i+1 in chunks of 32 positions; equal labels inside a chunk are combined with shuffles, across chunks via accw
        for (int i0 = 0; i0 < L; i0 += 32) {
            const int i = i0 + lane;
            const bool on = i < L;
            const int l = on ? lp[2 * i + 1] : -1 - lane;                 // inactive lanes form singleton groups
            const float v = on ? a[2 * i + 1] + be[2 * i + 1] : NEG_INF;
            const unsigned grp = __match_any_sync(FULL, l);
            float m = v;
            unsigned rem = grp & ~(1u << lane);
            while (__any_sync(FULL, rem != 0)) {
                const int src = rem ? __ffs(rem) - 1 : lane;
                const float o = __shfl_sync(FULL, v, src);
                if (rem) { m = fmaxf(m, o); rem &= rem - 1; }
            }
            float sum = (m == NEG_INF) ? 0.f : expf(v - m);
            rem = grp & ~(1u << lane);
            while (__any_sync(FULL, rem != 0)) {
                const int src = rem ? __ffs(rem) - 1 : lane;
                const float o = __shfl_sync(FULL, v, src);
                if (rem) { if (m != NEG_INF) sum += expf(o - m); rem &= rem - 1; }
            }
            if (on && (__ffs(grp) - 1) == lane && l >= 0 && l < V) {
                const float r = (m == NEG_INF) ? NEG_INF : m + logf(sum);
                accw[l] = lse2(accw[l], r);
            }
            __syncwarp();
        }
        if (lane == 0) accw[blank] = lse2(accw[blank], (mB == NEG_INF) ? NEG_INF : mB + logf(sB));
        __syncwarp();
        // gradient wrt u (overwrites logy[t]) and the chain to the dense2 logits: u=log(p+eps), p=softmax(z)
        const float* p = pb + (size_t)t * V;
        float* ly = logy + (size_t)t * V;
        float dot = 0.f;
        for (int k = lane; k < V; k += 32) {
            const float y = expf(ly[k]);
            const float ac = accw[k];
            const float g = (ac == NEG_INF || logp == NEG_INF) ? y : y - expf(ac - logp);
            ly[k] = g;
            if (gu) gu[k] = g;
            const float pk = __ldg(p + k);
            dot += __fdividef(g, pk + eps) * pk;
        }
        dot = warp_sum(dot);
        if (gz)
            for (int k = lane; k < V; k += 32) {
                const float pk = __ldg(p + k);
                gz[k] = scale * pk * (__fdividef(ly[k], pk + eps) - dot);
            }
        __syncwarp();
    }
}

// =================================================================================================
// Greedy decode: one warp per sequence.  argmax over classes of u=log(p+eps) (first max wins), merge
// repeats, drop blanks.  out (B,T) padded with -1.
// =================================================================================================
__global__ void ctc_greedy_kernel(const float* __restrict__ probs, const int* __restrict__ seq_len,
                                  int B, int T, int V, float eps,
                                  int* __restrict__ out, int* __restrict__ out_len, float* __restrict__ score)
{ pdl_enter();
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

__global__ void __launch_bounds__(256, 4) ctc_beam_kernel(const float* __restrict__ probs, const int* __restrict__ seq_len,
                                int B, int T, int V, float eps, int W, int merge_repeated, int P,
                                int* __restrict__ out, int* __restrict__ out_len, float* __restrict__ logprob,
                                int smem_per_warp_bytes)
{ pdl_enter();
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int b = blockIdx.x * warps + wib;
    if (b >= B) return;
    const unsigned FULL = 0xffffffffu;
    const int blank = V - 1, NC = V - 1;
    const int Tb = seq_len ? seq_len[b] : T;
    const int NOID = 0x7fffffff;
    const bool fast = W <= 16;     // sorted-candidate walk (see (a') below); wider beams scan all labels per parent

    unsigned char* base = smraw + (size_t)wib * smem_per_warp_bytes;
    float* in_ = (float*)base;                                           // V (padded)
    TrieNode* nodes = (TrieNode*)(in_ + ((V + 3) & ~3));                 // 1 + W*T
    int* bnode = (int*)(nodes + (((size_t)1 + (size_t)W * T + 1) & ~1)); // [2][32]
    float* bpb = (float*)(bnode + 64);                                   // [2][32] blank
    float* bpl = bpb + 64;                                               // [2][32] label
    float* bpt = bpl + 64;                                               // [2][32] total
    float* sv = bpt + 64;                                                // [32] + int [32]: scratch for the compactions / permutations of a step
    int* sk = (int*)(sv + 32);

    int nn = 1;        // nodes in the trie (uniform)
    int nb = 1;        // beam entries (uniform); slots are in descending-total order
    int cur = 0;
    if (lane == 0) {
        nodes[0].parent = -1; nodes[0].label = -1; nodes[0].first_child = -1; nodes[0].next_sib = -1;
        bnode[0] = 0; bpb[0] = 0.f; bpl[0] = NEG_INF; bpt[0] = 0.f;
    }
    __syncwarp();

    // the probabilities of step t + 1 are fetched while step t is processed (V <= 128: four registers per lane); the global-load
    // latency sat at the top of every step's dependency chain (ncu r2f: most-sampled line of the kernel)
    const bool pre = V <= 128;
    float nx[4] = {0.f, 0.f, 0.f, 0.f};
    if (pre && Tb > 0) {
        const float* p0 = probs + (size_t)b * T * V;
#pragma unroll
        for (int i = 0; i < 4; ++i) if (lane + 32 * i < V) nx[i] = __ldg(p0 + lane + 32 * i);
    }
    for (int t = 0; t < Tb; ++t) {
        const float* p = probs + ((size_t)b * T + t) * V;
        // (a) normalised step input: u - max(u)   (TF 1.8: max-subtraction only)
        float mx = -INFINITY;
        if (pre) {
            float cur4[4] = {nx[0], nx[1], nx[2], nx[3]};
            if (t + 1 < Tb) {
#pragma unroll
                for (int i = 0; i < 4; ++i) if (lane + 32 * i < V) nx[i] = __ldg(p + V + lane + 32 * i);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) if (lane + 32 * i < V) { float u = logf(cur4[i] + eps); in_[lane + 32 * i] = u; mx = fmaxf(mx, u); }
        } else {
            for (int k = lane; k < V; k += 32) { float u = logf(p[k] + eps); in_[k] = u; mx = fmaxf(mx, u); }
        }
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
        // ---- leaves list L over the lanes: (Lv, Lid); id<0: survivor -1-e, id>=0: child (q << 10) | k  (V <= 1024)
        float Lv = NEG_INF; int Lid = NOID; int nL = 0;
        auto insert = [&](float v, int id) {     // warp-uniform call; keeps L sorted (descending), capacity W
            int pos = __popc(__ballot_sync(FULL, Lv > v));
            float upv = __shfl_up_sync(FULL, Lv, 1); int upi = __shfl_up_sync(FULL, Lid, 1);
            if (lane > pos) { Lv = upv; Lid = upi; } else if (lane == pos) { Lv = v; Lid = id; }
            if (lane >= W) { Lv = NEG_INF; Lid = NOID; }
            if (nL < W) ++nL;
        };
        {   // the survivors enter in slot order; insert() puts a new leaf BEFORE the leaves of equal total, so sequential insertion = sort by
            // (total descending, slot descending): every survivor counts the ones that rank before it and the leaves are permuted through
            // shared memory in one go (ncu r2u: 10 one-by-one insertions were 7 % of the instructions)
            int rk = 0;
            for (int e = 0; e < nb; ++e) { const float o = __shfl_sync(FULL, s_nt, e); rk += (o > s_nt || (o == s_nt && e > lane)) ? 1 : 0; }
            if (lane < nb) { sv[rk] = s_nt; sk[rk] = -1 - lane; }
            __syncwarp();
            nL = nb < W ? nb : W;
            if (lane < nL) { Lv = sv[lane]; Lid = sk[lane]; }
            __syncwarp();
        }

        // (a') the best non-blank labels of this step, sorted (score descending, label ascending): lane j holds the j-th best.  Every
        // parent scores its children as in_[k] + const, so ONE sorted list serves all parents: a parent then walks its candidates best
        // first and stops at the first one that does not beat the bottom leaf -- the work per parent is (#insertions + 1) instead of a
        // scan of all V-1 labels with an insertion attempt for every label above the bottom it started with (ncu r2a: 53 % of the
        // kernel's instructions were that scan; candidates inserted early were evicted again by better siblings).
        // Pruning (ncu r2u: the three 32-wide bitonic sorts + merges were 21 % of the instructions): once the leaves are full, a child scores at
        // most in_[k] + opt[0] (best parent, and its blank probability is below its total) while insertion and the blocking count both need a
        // score above a leaf, i.e. above the bottom after phase 1 (the bottom never drops).  Labels that fail in_[k] + opt[0] > bottom0 are
        // dead for every parent; the survivors of that test are compacted through shared memory and sorted in ONE network of just their
        // power-of-two size.  More than 32 survivors: the chunked sort below -- on configs[3] (N(0,9) logits) that is still the common case
        // (ncu r2v: chunked sort 12 % of the instructions, pruned sort 1.5 %): the bottom after phase 1 carries this step's blank / repeat
        // probability of a survivor, which is far below the best new label, so the test is weak until the first parent has inserted its
        // children; peaked inputs (a trained model's softmax) are where it pays.
        float tv = NEG_INF; int tk = 0x7fff0000 + lane;
        float list_floor = NEG_INF;      // the sorted list holds every label whose child score can reach this value (see the second floor below)
        if (fast) {
            auto before = [](float va, int ka, float vb, int kb) { return va > vb || (va == vb && ka < kb); };
            const float bottom0 = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
            const float ot0 = opt[0];
            // labels whose best possible child score passes `thr` (strictly above, or >= when `incl`), compacted into sv / sk; returns their number
            auto compact = [&](float thr, bool incl) {
                int n = 0;
                for (int k0 = 0; k0 < NC; k0 += 32) {
                    const int k = k0 + lane;
                    const float cv = (k < NC) ? in_[k] : NEG_INF;
                    const float x0 = cv + ot0;
                    const bool pass = k < NC && (incl ? x0 >= thr : x0 > thr);
                    const unsigned bal = __ballot_sync(FULL, pass);
                    if (pass) { const int pos = n + __popc(bal & ((1u << lane) - 1u)); if (pos < 32) { sv[pos] = cv; sk[pos] = k; } }
                    n += __popc(bal);
                }
                return n;
            };
            int cnt = compact(bottom0, false);
            if (cnt > 32 && nL == W) {
                // Second, stronger floor B <= the bottom leaf AFTER parent 0 (the best entry, visited first, never blocked) has inserted its
                // children: the leaves are then the top W of {survivors} U {its new children}, and the W-th largest of any SUBSET of that union
                // bounds it from below.  Subset = the survivors + one new child per lane (the lane's best label, not the parent's own label --
                // scored apart with the blank probability -- and not a label of one of its surviving children, whose total is already a leaf).
                // B = W-th largest of two sorted lists a (leaves) and b (the 32 lane-best child scores, sorted here by value only)
                //   = max over i of min(a[i-1], b[W-1-i]).  Labels with in_[k] + opt[0] < B are dead: parent 0 never inserts a child below the
                // bottom it ends with (children are inserted best first and never evicted by a later, smaller one), later parents need to beat
                // a bottom >= B with a smaller total, and the blocking count is taken on the list only when the survivor's total is >= B
                // (list_floor below; otherwise the full scan).
                const unsigned km0 = __ballot_sync(FULL, lane < nb && s_q == 0);
                const int lab0 = nodes[on[0]].label;
                float lm = NEG_INF;
                for (int k0 = 0; k0 < NC; k0 += 32) {
                    const int k = k0 + lane;
                    float v = (k < NC && k != lab0) ? in_[k] : NEG_INF;
                    for (unsigned m2 = km0; m2; m2 &= m2 - 1) { const int kk = __shfl_sync(FULL, s_k, __ffs(m2) - 1); if (k == kk) v = NEG_INF; }
                    lm = fmaxf(lm, v);
                }
                float c = lm + ot0;
#pragma unroll
                for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
                    for (int stride = size >> 1; stride > 0; stride >>= 1) {
                        const float o = __shfl_xor_sync(FULL, c, stride);
                        c = (((lane & size) == 0) == ((lane & stride) == 0)) ? fmaxf(c, o) : fminf(c, o);      // descending
                    }
                }
                const float a1 = __shfl_up_sync(FULL, Lv, 1);
                const float b1 = __shfl_sync(FULL, c, lane < W ? W - 1 - lane : 0);
                float t = (lane <= W) ? fminf(lane == 0 ? INFINITY : a1, lane < W ? b1 : INFINITY) : NEG_INF;
                const float Bf = warp_max(t);
                if (Bf > bottom0) {
                    const int cnt2 = compact(Bf, true);
                    if (cnt2 <= 32) { cnt = cnt2; list_floor = Bf; }
                }
            }
            if (cnt <= 32) {
                __syncwarp();
                float cv = (lane < cnt) ? sv[lane] : NEG_INF; int ck = (lane < cnt) ? sk[lane] : 0x7ffe0000 + lane;
                // bitonic sort (descending) of the first 2^ceil(log2 cnt) lanes; the padding beyond them is already in place
                for (int size = 2; (size >> 1) < cnt; size <<= 1) {
                    for (int stride = size >> 1; stride > 0; stride >>= 1) {
                        const float ov = __shfl_xor_sync(FULL, cv, stride); const int ok = __shfl_xor_sync(FULL, ck, stride);
                        const bool desc = (lane & size) == 0, low = (lane & stride) == 0;
                        const bool other_first = before(ov, ok, cv, ck);
                        if ((low == desc) ? other_first : !other_first) { cv = ov; ck = ok; }
                    }
                }
                tv = cv; tk = ck;
            } else
            for (int k0 = 0; k0 < NC; k0 += 32) {
                float cv = (k0 + lane < NC) ? in_[k0 + lane] : NEG_INF; int ck = (k0 + lane < NC) ? k0 + lane : 0x7ffe0000 + lane;
                // bitonic sort of the chunk, descending
#pragma unroll
                for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
                    for (int stride = size >> 1; stride > 0; stride >>= 1) {
                        const float ov = __shfl_xor_sync(FULL, cv, stride); const int ok = __shfl_xor_sync(FULL, ck, stride);
                        const bool desc = (lane & size) == 0, low = (lane & stride) == 0;
                        const bool other_first = before(ov, ok, cv, ck);
                        if ((low == desc) ? other_first : !other_first) { cv = ov; ck = ok; }
                    }
                }
                // merge with the running top 32: max(A_i, B_{31-i}) is bitonic and holds the 32 best of the union; 5 merge stages sort it
                const float rv = __shfl_sync(FULL, cv, 31 - lane); const int rk = __shfl_sync(FULL, ck, 31 - lane);
                if (before(rv, rk, tv, tk)) { tv = rv; tk = rk; }
#pragma unroll
                for (int stride = 16; stride > 0; stride >>= 1) {
                    const float ov = __shfl_xor_sync(FULL, tv, stride); const int ok = __shfl_xor_sync(FULL, tk, stride);
                    const bool low = (lane & stride) == 0;
                    const bool other_first = before(ov, ok, tv, tk);
                    if (low ? other_first : !other_first) { tv = ov; tk = ok; }
                }
            }
        }

        // ---- phase 2: parents in beam order
        bool blocked = false;     // lane e: survivor e lost its t-1 probabilities (cannot expand children)
        for (int r = 0; r < nb; ++r) {
            const bool r_blocked = __shfl_sync(FULL, (int)blocked, r) != 0;
            const float ot = opt[r], ob = opb[r];
            float bottom = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
            if (r_blocked) continue;
            if (!(ot > bottom)) break;     // is_candidate(b->oldp) fails; the parents after r have smaller totals and the bottom never drops
            const int lab_r = nodes[on[r]].label;
            const unsigned km = __ballot_sync(FULL, lane < nb && s_q == r);   // survivors that are children of r
            // this parent's view of the sorted candidates (fast path): lane j's label is a child still to be created iff it is not the
            // parent's own label (scored apart, with the blank probability) and not one of its surviving children
            bool lab_ok = lab_r >= 0;
            bool valid = tk < NC && tk != lab_r;
            if (fast)
                for (unsigned m2 = km; m2; m2 &= m2 - 1) { const int kk = __shfl_sync(FULL, s_k, __ffs(m2) - 1); if (tk == kk) valid = false; if (kk == lab_r) lab_ok = false; }
            const float x = valid ? tv + ot : NEG_INF;
            // blocking test for every such survivor
            if (km) {
                const float vlab = (fast && lab_ok) ? in_[lab_r] + ob : NEG_INF;
                const float x31 = __shfl_sync(FULL, tv, 31) + ot;           // no label outside the sorted 32 scores above this
                for (unsigned m = km; m; m &= m - 1) {
                    const int c = __ffs(m) - 1;
                    const int k_c = __shfl_sync(FULL, s_k, c);
                    const float s_c = __shfl_sync(FULL, s_nt, c);
                    const unsigned pm = __ballot_sync(FULL, Lid == -1 - c);
                    bool blk;
                    if (pm == 0) blk = true;                     // already evicted from the leaves
                    else {
                        const int rank = __ffs(pm) - 1;          // rank of c in L
                        int cnt = W;
                        bool counted = false;
                        if (fast) {
                            // eligible children with a smaller label and a larger total, counted on the sorted list: exact unless the list's
                            // last entry still beats s_c (then labels beyond it might, too) and the count has not reached W yet
                            cnt = rank + __popc(__ballot_sync(FULL, x > s_c && tk < k_c)) + ((lab_r < k_c && vlab > s_c) ? 1 : 0);
                            counted = cnt >= W || (s_c >= list_floor && (NC <= 32 || !(x31 > s_c)));
                        }
                        if (!counted) {
                            cnt = rank;
                            for (int k0 = 0; k0 < k_c; k0 += 32) {
                                const int k = k0 + lane;
                                bool ok = k < k_c;
                                for (unsigned m2 = km; m2; m2 &= m2 - 1) { const int kk = __shfl_sync(FULL, s_k, __ffs(m2) - 1); if (k == kk) ok = false; }
                                float xk = ok ? in_[k] + ((k == lab_r) ? ob : ot) : NEG_INF;
                                cnt += __popc(__ballot_sync(FULL, xk > s_c));
                            }
                        }
                        blk = cnt >= W;
                    }
                    if (blk && lane == c) blocked = true;
                }
            }
            // insert the eligible children of r
            if (fast) {
                // the child that repeats the parent's own label is scored with the parent's blank probability: handled apart
                if (lab_ok) {
                    const float v = in_[lab_r] + ob;
                    bottom = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
                    if (v > bottom) insert(v, (r << 10) + lab_r);
                }
                // all other children, best first (lane order); equal scores keep the label order TF visits them in.  With W <= 16 the
                // 32 sorted labels always suffice: at most W - 1 of them are survivors of this parent, one is its own label, W get in.
                // How many get in is known up front: candidates only ever evict old leaves, from the end (a candidate cannot beat the bottom if
                // the bottom is an earlier, larger candidate), so after j insertions the bottom is old leaf W-1-j (or candidate j-1 if that is
                // smaller, and then candidate j fails both tests): candidate j gets in iff it beats old leaf W-1-j, and once one fails all
                // later ones do.  No bottom shuffle / compare / break per candidate (ncu r2v: this loop was 20 % of the instructions).
                // Measured alone: 12.2 -> 12.8 M lines/s on configs[3] (when first tried together with 64-bit packed sort keys the pair was 5 % slower --
                // the packed keys were the loss).
                bottom = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
                const unsigned vm = __ballot_sync(FULL, x > bottom);
                if (vm) {
                    const int j = __popc(vm & ((1u << lane) - 1u));
                    const float thr = __shfl_sync(FULL, Lv, j < W ? W - 1 - j : 0);
                    unsigned im = __ballot_sync(FULL, ((vm >> lane) & 1u) && j < W && x > thr);
                    const unsigned fail = vm & ~im;
                    if (fail) im &= (fail & (0u - fail)) - 1u;            // keep the prefix (it is one by the argument above)
                    for (unsigned m = im; m; m &= m - 1) {
                        const int src = __ffs(m) - 1;
                        insert(__shfl_sync(FULL, x, src), (r << 10) + __shfl_sync(FULL, tk, src));
                    }
                }
            } else {
                for (int k0 = 0; k0 < NC; k0 += 32) {
                    const int k = k0 + lane;
                    float x = (k < NC) ? in_[k] + ((k == lab_r) ? ob : ot) : NEG_INF;
                    for (unsigned m2 = km; m2; m2 &= m2 - 1) { int kk = __shfl_sync(FULL, s_k, __ffs(m2) - 1); if (k == kk) x = NEG_INF; }
                    bottom = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
                    for (unsigned m = __ballot_sync(FULL, x > bottom); m; m &= m - 1) {
                        const int src = __ffs(m) - 1;
                        const float v = __shfl_sync(FULL, x, src);
                        bottom = (nL == W) ? __shfl_sync(FULL, Lv, W - 1) : NEG_INF;
                        if (v > bottom) insert(v, (r << 10) + k0 + src);
                    }
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
        {   // trie find-or-create for the new children, one lane per leaf: the search of the parent's sibling list is read-only; the nodes
            // that have to be created get consecutive indices (prefix count over the creating lanes) and are pushed at the head of their
            // parent's list -- lanes that share a parent chain their nodes in lane order, the last of them becomes the new head
            const bool child = lane < nL && Lid >= 0;
            int par = -1, k = 0, c = -1;
            if (child) {
                par = on[Lid >> 10]; k = Lid & 1023;
                c = nodes[par].first_child;
                while (c >= 0 && nodes[c].label != k) c = nodes[c].next_sib;
            }
            const unsigned mk = __ballot_sync(FULL, child && c < 0);
            if (mk) {
                const unsigned lt = (1u << lane) - 1u;
                const bool mine = (mk >> lane) & 1u;
                const unsigned grp = __match_any_sync(FULL, mine ? par : -2 - lane) & mk;     // creating lanes with the same parent
                int old_head = -1;
                if (mine) old_head = nodes[par].first_child;
                __syncwarp();
                if (mine) {
                    c = nn + __popc(mk & lt);
                    const unsigned below = grp & lt;
                    const int prev = below ? nn + __popc(mk & ((1u << (31 - __clz(below))) - 1u)) : old_head;
                    TrieNode nd; nd.parent = (short)par; nd.label = (short)k; nd.first_child = -1; nd.next_sib = (short)prev;
                    nodes[c] = nd;
                    if (!(grp >> lane >> 1)) nodes[par].first_child = (short)c;                 // highest lane of the group
                }
                nn += __popc(mk);
            }
            if (child) nnode[lane] = c;
        }
        nb = nL;
        cur = nxt;
        __syncwarp();
    }

    // top paths = slots 0..P-1 (slots are sorted by total, best first = TF BeamSearch::TopPaths); before the first step slot 0 is the root.
    // out (B,P,T), out_len / logprob (B,P); a path beyond the number of leaves comes back empty with logprob = -inf.
    const float* fpt = bpt + cur * 32; const int* fn = bnode + cur * 32;
    for (int pth = 0; pth < P; ++pth) {
        int* o = out + ((size_t)b * P + pth) * T;
        int n = 0;
        if (lane == 0) {
            if (pth < nb) {
                // TF BeamEntry::LabelSeq: walk leaf -> root, drop a label equal to the previously visited one
                int prev = -1;
                for (int c = fn[pth]; nodes[c].parent >= 0; c = nodes[c].parent) {
                    int l = nodes[c].label;
                    if (!merge_repeated || l != prev) ++n;
                    prev = l;
                }
                int i = n; prev = -1;
                for (int c = fn[pth]; nodes[c].parent >= 0; c = nodes[c].parent) {
                    int l = nodes[c].label;
                    if (!merge_repeated || l != prev) o[--i] = l;
                    prev = l;
                }
            }
            out_len[(size_t)b * P + pth] = n;
            if (logprob) logprob[(size_t)b * P + pth] = pth < nb ? fpt[pth] : NEG_INF;
        }
        n = __shfl_sync(FULL, n, 0);
        for (int t = n + lane; t < T; t += 32) o[t] = -1;
    }
}

// =================================================================================================
// host launchers
// =================================================================================================
size_t ctc_loss_smem_bytes(int T, int t_off, int V, int maxL) {
    int Tp = T - t_off, Smax = 2 * maxL + 1;
    return sizeof(float) * ((size_t)Tp * V + 2 * (size_t)Tp * Smax + CTC_WARPS * (size_t)V) + sizeof(int) * Smax;
}

int launch_ctc_loss_grad(const float* probs, int t_off, const int* labels, int maxL, const int* label_len,
                         const int* input_len, int B, int T, int V, float eps, float* loss, float* grad_u,
                         float* grad_logits, float scale, int* status, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    if (T - t_off <= 0 || V < 2 || maxL < 0) { crnn_set_error("ctc_loss: bad shape"); return CRNN_ERR_INVALID; }
    if (2 * maxL + 1 > 32 * CTC_MAX_NS) { crnn_set_error("ctc_loss: max label length %d exceeds %d", maxL, (32 * CTC_MAX_NS - 1) / 2); return CRNN_ERR_INVALID; }
    size_t smem = ctc_loss_smem_bytes(T, t_off, V, maxL);
    if (smem > 227 * 1024) { crnn_set_error("ctc_loss: T'*(V+2S) too large for shared memory (%zu B)", smem); return CRNN_ERR_INVALID; }
    const int ns = (2 * maxL + 1 + 31) / 32;
    CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int), st));
#define CTC_LAUNCH(NS_)                                                                                                                   \
    do {                                                                                                                                  \
        static size_t configured = 0;                                                                                                     \
        if (smem > 48 * 1024 && smem > configured) {                                                                                      \
            CUDA_TRY(cudaFuncSetAttribute(ctc_loss_grad_kernel<NS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
            configured = smem;                                                                                                            \
        }                                                                                                                                 \
        (void)crnn_launch(ctc_loss_grad_kernel<NS_>, B, CTC_WARPS * 32, smem, st, probs, t_off, labels, maxL, label_len, input_len, B, T, V, eps,        \
                                                                   loss, grad_u, grad_logits, scale, status);                             \
    } while (0)
    if (ns <= 1) CTC_LAUNCH(1); else if (ns <= 2) CTC_LAUNCH(2); else if (ns <= 4) CTC_LAUNCH(4); else CTC_LAUNCH(8);
#undef CTC_LAUNCH
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_ctc_greedy(const float* probs, const int* seq_len, int B, int T, int V, float eps,
                      int* out, int* out_len, float* score, cudaStream_t st)
{
    if (B <= 0) return CRNN_OK;
    const int warps = 4;
    (void)crnn_launch(ctc_greedy_kernel, ceil_div(B, warps), warps * 32, 0, st, probs, seq_len, B, T, V, eps, out, out_len, score);
    LAUNCH_CHECK();
    return CRNN_OK;
}

int launch_ctc_beam(const float* probs, const int* seq_len, int B, int T, int V, float eps, int W, int merge_repeated,
                    int* out, int* out_len, float* logprob, cudaStream_t st, int top_paths)
{
    if (B <= 0) return CRNN_OK;
    if (top_paths < 1 || top_paths > W) { crnn_set_error("ctc_beam: top_paths %d not in [1, beam width %d]", top_paths, W); return CRNN_ERR_INVALID; }
    if (W < 1 || W > BEAM_MAX_W) { crnn_set_error("ctc_beam: beam width %d not in [1,%d]", W, BEAM_MAX_W); return CRNN_ERR_INVALID; }
    if (V < 2 || V > 1024) { crnn_set_error("ctc_beam: %d classes not in [2,1024]", V); return CRNN_ERR_INVALID; }
    if ((size_t)1 + (size_t)W * T > 32000) { crnn_set_error("ctc_beam: W*T too large"); return CRNN_ERR_INVALID; }
    size_t per = sizeof(float) * ((V + 3) & ~3) + sizeof(TrieNode) * (((size_t)1 + (size_t)W * T + 1) & ~1)
               + sizeof(int) * 64 + sizeof(float) * 64 * 3 + sizeof(float) * 64;
    per = (per + 15) & ~(size_t)15;
    // 4 warps per CTA: the kernel is bound by instruction issue, i.e. by the SM with the most resident warps; 4096 sequences in CTAs of 8 put
    // 32 warps on some SMs and 24 on others (512 CTAs over 148 SMs), in CTAs of 4 they spread 28 / 24.  The register cap of 64
    // (__launch_bounds__(256, 4)) keeps the whole batch resident in one wave: at 67 registers the last 68 CTAs ran as a second wave
    // (ncu r2u: 554 us instead of 441 us with 17 % FEWER instructions)
    int warps = 4;
    while (warps > 1 && per * warps > 200 * 1024) warps >>= 1;
    size_t smem = per * warps;
    if (smem > 227 * 1024) { crnn_set_error("ctc_beam: per-sequence state %zu B exceeds shared memory", per); return CRNN_ERR_INVALID; }
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        CUDA_TRY(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    (void)crnn_launch(ctc_beam_kernel, ceil_div(B, warps), warps * 32, smem, st, probs, seq_len, B, T, V, eps, W, merge_repeated, top_paths,
                                                                  out, out_len, logprob, (int)per);
    LAUNCH_CHECK();
    return CRNN_OK;
}
