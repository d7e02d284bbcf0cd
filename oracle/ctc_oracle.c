/*
 * oracle/ctc_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or imported by the product path).
 *
 * CPU restatement of the CTC stages of the reference hot path:
 *   - K.ctc_batch_cost      (call site: reference utils.py:98-103)   -> ctc_oracle_loss_grad
 *   - K.ctc_decode greedy   (implied API, BASELINE configs[0,1])     -> ctc_oracle_greedy
 *   - K.ctc_decode beam     (call site: reference utils.py:347-357)  -> ctc_oracle_beam
 *
 * The arithmetic of those calls lives in keras==2.2.2 / tensorflow==1.8.0 (pinned by the reference's
 * Dockerfile:61-63), which are NOT vendored under /root/reference and cannot be installed here.  This
 * file restates their published algorithms as summarised in SURVEY.md Appendix A.1/A.2
 * (TF core/util/ctc/ctc_loss_calculator.{h,cc}, ctc_beam_search.h, ctc_beam_entry.h, lib/gtl/top_n.h;
 * Keras backend/tensorflow_backend.py ctc_batch_cost / ctc_decode, epsilon()=1e-7).
 *
 * PARITY UNPINNED for the loss/gradient: the reference ships no tests / golden vectors for this path and its
 * own implementation cannot run in this image; the restatement is pinned only by independent cross-checks
 * (torch.nn.functional.ctc_loss in fp64, brute-force most-probable-labelling enumeration) -- see
 * tests/test_oracle_ctc.py.  The beam decoder is additionally exercised by the reference's own example
 * predictions (tests/test_golden.py: shipped weights + README-figure inputs -> the labels the reference printed).
 *
 * All arithmetic is float32 in log space, exactly like the TF CPU kernels.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LOGZERO (-INFINITY)

static inline float lse2(float a, float b) {
    /* TF ctc_loss_util.h LogSumExp */
    if (a == LOGZERO && b == LOGZERO) return LOGZERO;
    return (a > b) ? a + log1pf(expf(b - a)) : b + log1pf(expf(a - b));
}

/* ------------------------------------------------------------------------------------------------
 * CTC loss + gradient wrt u = log(p + eps), one sequence at a time.
 *   probs   (B, T, V)  softmax output of the network AFTER the [:, 2:, :] slice (utils.py:102)
 *   labels  (B, maxL)  int32, only the first label_len[b] entries are read (Keras dense->sparse)
 *   loss    (B)        -log p(l|x)
 *   grad_u  (B, T, V)  d loss / d u  (TF CTCLoss op gradient output); rows t >= input_len[b] are 0
 * returns 0, or -(b+1) if sequence b has "not enough time for target transition sequence".
 * ------------------------------------------------------------------------------------------------ */
int ctc_oracle_loss_grad(const float* probs, int B, int T, int V,
                         const int* labels, int maxL, const int* label_len, const int* input_len,
                         float eps, float* loss, float* grad_u)
{
    const int blank = V - 1;
    const int Umax = 2 * maxL + 1;
    float* logy = (float*)malloc(sizeof(float) * (size_t)T * V);
    float* y = (float*)malloc(sizeof(float) * (size_t)T * V);
    float* alpha = (float*)malloc(sizeof(float) * (size_t)T * Umax);
    float* beta = (float*)malloc(sizeof(float) * (size_t)T * Umax);
    int* lp = (int*)malloc(sizeof(int) * Umax);
    int status = 0;

    for (int b = 0; b < B; ++b) {
        const int L = label_len[b], Tb = input_len[b], U = 2 * L + 1;
        float* g = grad_u + (size_t)b * T * V;
        memset(g, 0, sizeof(float) * (size_t)T * V);
        int repeats = 0;
        for (int i = 1; i < L; ++i) repeats += labels[b * maxL + i] == labels[b * maxL + i - 1];
        if (Tb > T || L + repeats > Tb) { status = -(b + 1); loss[b] = INFINITY; continue; }
        for (int s = 0; s < U; ++s) lp[s] = (s & 1) ? labels[b * maxL + (s >> 1)] : blank;

        /* u = log(p + eps); TF re-softmaxes u (max-subtracted, float) */
        for (int t = 0; t < Tb; ++t) {
            const float* p = probs + ((size_t)b * T + t) * V;
            float mx = -INFINITY;
            for (int k = 0; k < V; ++k) { float u = logf(p[k] + eps); logy[t * V + k] = u; if (u > mx) mx = u; }
            float sum = 0.f;
            for (int k = 0; k < V; ++k) { float e = expf(logy[t * V + k] - mx); y[t * V + k] = e; sum += e; }
            for (int k = 0; k < V; ++k) { y[t * V + k] /= sum; logy[t * V + k] = logf(y[t * V + k]); }
        }
        for (int i = 0; i < Tb * U; ++i) alpha[i] = beta[i] = LOGZERO;
        /* forward */
        alpha[0] = logy[blank];
        if (U > 1) alpha[1] = logy[lp[1]];
        for (int t = 1; t < Tb; ++t) {
            int lo = U - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > U) hi = U;
            for (int s = lo; s < hi; ++s) {
                float a = alpha[(t - 1) * U + s];
                if (s > 0) a = lse2(a, alpha[(t - 1) * U + s - 1]);
                if (s > 1 && lp[s] != blank && lp[s] != lp[s - 2]) a = lse2(a, alpha[(t - 1) * U + s - 2]);
                alpha[t * U + s] = logy[t * V + lp[s]] + a;
            }
        }
        /* backward (beta excludes the emission at t) */
        beta[(Tb - 1) * U + U - 1] = 0.f;
        if (U > 1) beta[(Tb - 1) * U + U - 2] = 0.f;
        for (int t = Tb - 2; t >= 0; --t) {
            int lo = U - 2 * (Tb - t); if (lo < 0) lo = 0;
            int hi = 2 * (t + 1); if (hi > U) hi = U;
            for (int s = lo; s < hi; ++s) {
                float v = beta[(t + 1) * U + s] + logy[(t + 1) * V + lp[s]];
                if (s + 1 < U) v = lse2(v, beta[(t + 1) * U + s + 1] + logy[(t + 1) * V + lp[s + 1]]);
                if (s + 2 < U && lp[s + 2] != blank && lp[s + 2] != lp[s])
                    v = lse2(v, beta[(t + 1) * U + s + 2] + logy[(t + 1) * V + lp[s + 2]]);
                beta[t * U + s] = v;
            }
        }
        float logp = LOGZERO;
        for (int s = 0; s < U; ++s) logp = lse2(logp, alpha[s] + beta[s]);
        loss[b] = -logp;
        /* gradient */
        for (int t = 0; t < Tb; ++t) {
            for (int k = 0; k < V; ++k) {
                float acc = LOGZERO;
                for (int s = 0; s < U; ++s) if (lp[s] == k) acc = lse2(acc, alpha[t * U + s] + beta[t * U + s]);
                float yv = y[t * V + k];
                g[t * V + k] = (acc == LOGZERO || logp == LOGZERO) ? yv : yv - expf(acc - logp);
            }
        }
    }
    free(logy); free(y); free(alpha); free(beta); free(lp);
    return status;
}

/* ------------------------------------------------------------------------------------------------
 * Greedy decode: argmax per frame on u = log(p+eps) (first max wins), merge repeated,
 * drop blanks; neg_sum_logits[b] = -sum_t max_k u.
 *   out (B, T) int32 padded with -1, out_len (B)
 * ------------------------------------------------------------------------------------------------ */
void ctc_oracle_greedy(const float* probs, int B, int T, int V, const int* seq_len, float eps,
                       int* out, int* out_len, float* neg_sum_logits)
{
    const int blank = V - 1;
    for (int b = 0; b < B; ++b) {
        int n = 0, prev = -1; float acc = 0.f;
        for (int t = 0; t < T; ++t) out[b * T + t] = -1;
        for (int t = 0; t < seq_len[b]; ++t) {
            const float* p = probs + ((size_t)b * T + t) * V;
            int best = 0; float bu = logf(p[0] + eps);
            for (int k = 1; k < V; ++k) { float u = logf(p[k] + eps); if (u > bu) { bu = u; best = k; } }
            acc += -bu;
            if (best != blank && best != prev) out[b * T + n++] = best;
            prev = best;
        }
        out_len[b] = n;
        if (neg_sum_logits) neg_sum_logits[b] = acc;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Beam search, literal restatement of TF 1.8 CTCBeamSearchDecoder with the default
 * scorer: prefix tree of beam entries with (blank,label,total) log-probs for t-1 ("old") and t
 * ("new"), a TopN of `beam_width` leaves ordered by new.total, sequential insert / evict.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int parent, label;
    int children;      /* index into Tree.kids of this entry's (V-1) child ids, -1 until populated (TF PopulateChildren) */
    float ob, ol, ot;  /* oldp: blank, label, total */
    float nb, nl, nt;  /* newp */
} Entry;

typedef struct {
    Entry* e; int n, cap;
    int* kids; int nk, capk;   /* pool of children-id arrays, (V-1) ints each */
    int nc;                    /* V-1 */
} Tree;

static int tree_new(Tree* tr, int parent, int label) {
    if (tr->n == tr->cap) { tr->cap *= 2; tr->e = (Entry*)realloc(tr->e, sizeof(Entry) * tr->cap); }
    Entry* x = &tr->e[tr->n];
    x->parent = parent; x->label = label; x->children = -1;
    x->ob = x->ol = x->ot = x->nb = x->nl = x->nt = LOGZERO;
    return tr->n++;
}
/* BeamEntry::GetChild: children are created lazily with oldp=newp=-inf */
static int tree_child(Tree* tr, int parent, int label) {
    if (tr->e[parent].children < 0) {
        if (tr->nk + tr->nc > tr->capk) { while (tr->nk + tr->nc > tr->capk) tr->capk *= 2; tr->kids = (int*)realloc(tr->kids, sizeof(int) * tr->capk); }
        tr->e[parent].children = tr->nk;
        for (int k = 0; k < tr->nc; ++k) tr->kids[tr->nk + k] = -1;
        tr->nk += tr->nc;
    }
    int* slot = &tr->kids[tr->e[parent].children + label];
    if (*slot < 0) { int id = tree_new(tr, parent, label); slot = &tr->kids[tr->e[parent].children + label]; *slot = id; }
    return *slot;
}

/* leaves: unordered array of <= W entry ids; "bottom" = smallest new.total */
static int leaves_bottom(const Tree* tr, const int* leaves, int n) {
    int bi = 0;
    for (int i = 1; i < n; ++i) if (tr->e[leaves[i]].nt < tr->e[leaves[bi]].nt) bi = i;
    return bi;
}

/*   probs (B,T,V) -> out (B,P,T) int32 padded -1 (the P = top_paths best paths, best first; merge_repeated as given), out_len (B,P),
 *   log_prob (B,P) = new.total of each path (TF 1.8: max-subtracted, un-normalised).  TF BeamSearch::TopPaths: leaves sorted by
 *   total, descending.  Paths beyond the number of leaves come back empty with log_prob = -inf (TF raises an error there).       */
void ctc_oracle_beam_topk(const float* probs, int B, int T, int V, const int* seq_len, float eps,
                          int beam_width, int merge_repeated, int top_paths, int* out, int* out_len, float* log_prob)
{
    const int P = top_paths;
    const int blank = V - 1, W = beam_width;
    float* in = (float*)malloc(sizeof(float) * V);
    int* leaves = (int*)malloc(sizeof(int) * (W + 1));
    int* branches = (int*)malloc(sizeof(int) * (W + 1));
    Tree tr; tr.cap = 1024; tr.e = (Entry*)malloc(sizeof(Entry) * tr.cap);
    tr.capk = 4096; tr.kids = (int*)malloc(sizeof(int) * tr.capk); tr.nc = V - 1;

    for (int b = 0; b < B; ++b) {
        tr.n = 0; tr.nk = 0;
        int root = tree_new(&tr, -1, -1);
        tr.e[root].nt = 0.f; tr.e[root].nb = 0.f; tr.e[root].nl = LOGZERO;
        int nleaves = 1; leaves[0] = root;

        for (int t = 0; t < seq_len[b]; ++t) {
            const float* p = probs + ((size_t)b * T + t) * V;
            float mx = -INFINITY;
            for (int k = 0; k < V; ++k) { in[k] = logf(p[k] + eps); if (in[k] > mx) mx = in[k]; }
            for (int k = 0; k < V; ++k) in[k] -= mx;

            /* branches = leaves.Extract(): sorted best-first (stable insertion sort) */
            int nbr = nleaves;
            for (int i = 0; i < nbr; ++i) branches[i] = leaves[i];
            for (int i = 1; i < nbr; ++i) {
                int x = branches[i], j = i - 1;
                while (j >= 0 && tr.e[branches[j]].nt < tr.e[x].nt) { branches[j + 1] = branches[j]; --j; }
                branches[j + 1] = x;
            }
            nleaves = 0;
            for (int i = 0; i < nbr; ++i) { Entry* e = &tr.e[branches[i]]; e->ob = e->nb; e->ol = e->nl; e->ot = e->nt; }
            /* (1) update survivors */
            for (int i = 0; i < nbr; ++i) {
                Entry* e = &tr.e[branches[i]];
                if (e->parent >= 0) {
                    const Entry* par = &tr.e[e->parent];
                    if (par->nt != LOGZERO) { /* Active(parent) */
                        float prev = (e->label == par->label) ? par->ob : par->ot;
                        e->nl = lse2(e->nl, prev);
                    }
                    e->nl += in[e->label];
                }
                e->nb = e->ot + in[blank];
                e->nt = lse2(e->nb, e->nl);
                leaves[nleaves++] = branches[i];
            }
            /* (2) grow children, best parent first */
            for (int i = 0; i < nbr; ++i) {
                const int bi = branches[i];
                {
                    const Entry* e = &tr.e[bi];
                    int cand = e->ot > LOGZERO &&
                               (nleaves < W || e->ot > tr.e[leaves[leaves_bottom(&tr, leaves, nleaves)]].nt);
                    if (!cand) continue;
                }
                for (int k = 0; k < V; ++k) {
                    if (k == blank) continue;
                    int ci = tree_child(&tr, bi, k);   /* may realloc: re-fetch pointers after */
                    Entry* c = &tr.e[ci];
                    const Entry* e = &tr.e[bi];
                    if (c->nt != LOGZERO) continue;   /* already active: handled in (1) */
                    c->nb = LOGZERO;
                    float prev = (k == e->label) ? e->ob : e->ot;
                    c->nl = in[k] + prev;
                    c->nt = c->nl;
                    int is_cand = c->nt > LOGZERO;
                    int bot = -1;
                    if (is_cand && nleaves >= W) {
                        bot = leaves_bottom(&tr, leaves, nleaves);
                        is_cand = c->nt > tr.e[leaves[bot]].nt;
                    }
                    if (is_cand) {
                        if (nleaves >= W) {
                            Entry* z = &tr.e[leaves[bot]];
                            z->nb = z->nl = z->nt = LOGZERO;
                            leaves[bot] = ci;
                        } else {
                            leaves[nleaves++] = ci;
                        }
                    } else {
                        c->ob = c->ol = c->ot = c->nb = c->nl = c->nt = LOGZERO;
                    }
                }
            }
        }
        /* top paths: leaves sorted by total, best first (stable insertion sort) */
        for (int i = 1; i < nleaves; ++i) {
            int x = leaves[i], j = i - 1;
            while (j >= 0 && tr.e[leaves[j]].nt < tr.e[x].nt) { leaves[j + 1] = leaves[j]; --j; }
            leaves[j + 1] = x;
        }
        int* tmp = (int*)malloc(sizeof(int) * (T + 1));
        for (int pth = 0; pth < P; ++pth) {
            int* o = out + ((size_t)b * P + pth) * T;
            for (int t = 0; t < T; ++t) o[t] = -1;
            if (pth >= nleaves) { out_len[b * P + pth] = 0; if (log_prob) log_prob[b * P + pth] = LOGZERO; continue; }
            int n = 0, prev = -1;
            /* walk leaf -> root collecting labels, dropping a label equal to the previously visited one */
            for (int c = leaves[pth]; tr.e[c].parent >= 0; c = tr.e[c].parent) {
                if (!merge_repeated || tr.e[c].label != prev) tmp[n++] = tr.e[c].label;
                prev = tr.e[c].label;
            }
            for (int i = 0; i < n; ++i) o[i] = tmp[n - 1 - i];
            out_len[b * P + pth] = n;
            if (log_prob) log_prob[b * P + pth] = tr.e[leaves[pth]].nt;
        }
        free(tmp);
    }
    free(in); free(leaves); free(branches); free(tr.e); free(tr.kids);
}

/* top_paths = 1 (what DecodeCTCPred.decode uses, utils.py:353-356) */
void ctc_oracle_beam(const float* probs, int B, int T, int V, const int* seq_len, float eps,
                     int beam_width, int merge_repeated, int* out, int* out_len, float* log_prob)
{
    ctc_oracle_beam_topk(probs, B, T, V, seq_len, eps, beam_width, merge_repeated, 1, out, out_len, log_prob);
}
