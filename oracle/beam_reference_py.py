"""Second, independent statement of the TF-1.8 prefix beam search (TEST INFRASTRUCTURE ONLY).

This is the *parent-sequential* formulation the CUDA kernel implements (csrc/ctc.cu), written over prefix tuples in
float32 with TF's LogSumExp.  It is NOT the naive "score everything, keep the global top-W": TF's sequential
insert/evict has an observable quirk -- a survivor that is evicted from the leaves during the step and is then
re-visited as a child of its own (better-ranked) parent fails the candidate test and gets its t-1 probabilities
reset ("Deactivate child", ctc_beam_search.h), so it no longer expands children in that step.  The formulation:

  per step: update survivors (phase 1); then visit the parents in beam order (best t-1 total first).  For parent b
  (skipped if blocked or its t-1 total does not beat the current bottom): every survivor c that is a child of b is
  tested -- if it has already fallen out of the top-W (or would, counting b's eligible children with a smaller
  label that beat it), it is BLOCKED; then L <- top-W(L U eligible children of b), insertion needing a strictly
  larger total than the bottom when L is full.
Used to check the formulation against the literal sequential restatement in ctc_oracle.c.
"""
import numpy as np

NEG = np.float32(-np.inf)


def _lse(a, b):
    a = np.float32(a); b = np.float32(b)
    if a == NEG and b == NEG:
        return NEG
    hi, lo = (a, b) if a > b else (b, a)
    return np.float32(hi + np.log1p(np.exp(np.float32(lo - hi), dtype=np.float32), dtype=np.float32))


def beam_parent_sequential(p, W, merge_repeated, eps=1e-7):
    T, V = p.shape
    blank = V - 1
    # beam: list of dicts in rank order (best total first)
    beam = [{"pre": (), "b": np.float32(0), "l": NEG, "t": np.float32(0)}]
    for t in range(T):
        u = np.log(p[t].astype(np.float32) + np.float32(eps), dtype=np.float32)
        u = (u - u.max()).astype(np.float32)
        slot = {e["pre"]: i for i, e in enumerate(beam)}
        # phase 1: survivors
        surv = []
        for e in beam:
            nl = e["l"]
            pre = e["pre"]
            q = -1
            if pre:
                q = slot.get(pre[:-1], -1)
                if q >= 0:
                    par = beam[q]
                    prev = par["b"] if (len(par["pre"]) and par["pre"][-1] == pre[-1]) else par["t"]
                    nl = _lse(nl, prev)
                nl = np.float32(nl + u[pre[-1]])
            nb = np.float32(e["t"] + u[blank])
            surv.append({"pre": pre, "b": nb, "l": nl, "t": _lse(nb, nl), "q": q})
        # L: sorted list of (total, kind, payload)
        L = sorted([(s["t"], "s", i) for i, s in enumerate(surv)], key=lambda x: -x[0])
        blocked = [False] * len(beam)
        for bi, b in enumerate(beam):       # beam order == descending t-1 total
            if blocked[bi]:
                continue
            if len(L) >= W and not (b["t"] > L[-1][0]):
                continue
            kids = {surv[i]["pre"][-1]: i for i in range(len(beam)) if surv[i]["q"] == bi}
            x = {}
            for k in range(V - 1):
                if k in kids:
                    continue
                prev = b["b"] if (b["pre"] and b["pre"][-1] == k) else b["t"]
                s = np.float32(u[k] + prev)
                if s > NEG:
                    x[k] = s
            for k_c, ci in kids.items():
                in_L = [j for j, it in enumerate(L) if it[1] == "s" and it[2] == ci]
                if not in_L:
                    blocked[ci] = True
                    continue
                rank = in_L[0]
                s_c = surv[ci]["t"]
                cnt = sum(1 for k, v in x.items() if k < k_c and v > s_c)
                if rank + cnt >= W:
                    blocked[ci] = True
            for k in sorted(x):
                v = x[k]
                if len(L) < W or v > L[-1][0]:
                    pos = sum(1 for it in L if it[0] > v)
                    L.insert(pos, (v, "c", (bi, k)))
                    if len(L) > W:
                        L.pop()
        nbeam = []
        for tot, kind, pay in L:
            if kind == "s":
                s = surv[pay]
                nbeam.append({"pre": s["pre"], "b": s["b"], "l": s["l"], "t": s["t"]})
            else:
                bi, k = pay
                nbeam.append({"pre": beam[bi]["pre"] + (k,), "b": NEG, "l": tot, "t": tot})
        beam = nbeam
    best = max(beam, key=lambda e: e["t"])["pre"]
    out, prev = [], -1
    for k in reversed(best):     # TF LabelSeq walks leaf -> root
        if not merge_repeated or k != prev:
            out.append(int(k))
        prev = k
    return out[::-1]


# kept under the old name for the tests
beam_global_topk = beam_parent_sequential
