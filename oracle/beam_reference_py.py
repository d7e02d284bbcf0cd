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


from collections import Counter
STATS = Counter()      # how often each short cut of the fast formulation was taken (tests assert that all of them were exercised)


def beam_parent_sequential_fast(p, W, merge_repeated, eps=1e-7, trace=None):
    """The same formulation with the SHORT CUTS of the CUDA kernel's fast path (csrc/ctc.cu, W <= 16), restated so that their exactness can be
    checked on the CPU against beam_parent_sequential over far more inputs than the GPU tests see:
      * one list of candidate labels per step, sorted (score descending, label ascending), at most 32 long, pruned by two floors: labels
        whose best possible child score u[k] + t_best cannot beat the bottom leaf after phase 1 (floor 1), or is below the W-th largest of
        {leaves} U {one new child per lane of the best parent} (floor 2, used when floor 1 leaves more than 32);
      * the blocking count of a surviving child taken on that list (full scan only when the list may be incomplete for that child);
      * the number of children a parent inserts in closed form: candidate j gets in iff it beats old leaf W-1-j.
    `trace`, if a list, receives the beam (prefix, total) after every step."""
    assert W <= 16
    T, V = p.shape
    blank, NC = V - 1, V - 1
    beam = [{"pre": (), "b": np.float32(0), "l": NEG, "t": np.float32(0)}]
    for t in range(T):
        u = np.log(p[t].astype(np.float32) + np.float32(eps), dtype=np.float32)
        u = (u - u.max()).astype(np.float32)
        slot = {e["pre"]: i for i, e in enumerate(beam)}
        surv = []
        STATS["steps"] += 1
        for e in beam:
            nl, pre, q = e["l"], e["pre"], -1
            if pre:
                q = slot.get(pre[:-1], -1)
                if q >= 0:
                    par = beam[q]
                    prev = par["b"] if (len(par["pre"]) and par["pre"][-1] == pre[-1]) else par["t"]
                    nl = _lse(nl, prev)
                nl = np.float32(nl + u[pre[-1]])
            nb = np.float32(e["t"] + u[blank])
            surv.append({"pre": pre, "b": nb, "l": nl, "t": _lse(nb, nl), "q": q})
        nb_ = len(beam)
        # leaves: sequential insertion puts a new leaf BEFORE leaves of equal total -> (total desc, slot desc)
        order = sorted(range(nb_), key=lambda e: (-surv[e]["t"], -e))
        L = [(surv[e]["t"], "s", e) for e in order]
        # ---- candidate list
        ot0 = beam[0]["t"]
        bottom0 = L[W - 1][0] if len(L) == W else NEG
        cand = [k for k in range(NC) if np.float32(u[k] + ot0) > bottom0]
        list_floor = NEG
        if len(cand) > 32 and len(L) == W:
            kids0 = {surv[i]["pre"][-1] for i in range(nb_) if surv[i]["q"] == 0}
            lab0 = beam[0]["pre"][-1] if beam[0]["pre"] else -1
            lm = []
            for lane in range(32):
                vals = [u[k] for k in range(lane, NC, 32) if k != lab0 and k not in kids0]
                lm.append(np.float32(max(vals) + ot0) if vals else NEG)
            c = sorted(lm, reverse=True)
            a = [it[0] for it in L] + [NEG] * 32
            Bf = NEG
            for i in range(W + 1):
                ai = np.float32(np.inf) if i == 0 else a[i - 1]
                bi_ = np.float32(np.inf) if i == W else c[W - 1 - i]
                Bf = max(Bf, min(ai, bi_))
            if Bf > bottom0:
                cand2 = [k for k in range(NC) if np.float32(u[k] + ot0) >= Bf]
                if len(cand2) <= 32:
                    cand, list_floor = cand2, Bf
                    STATS["floor2"] += 1
        STATS["floor1" if (len(cand) <= 32 and list_floor == NEG and len(L) == W) else "other"] += 1
        STATS["chunked"] += 1 if len(cand) > 32 else 0
        cand.sort(key=lambda k: (-u[k], k))
        cand = cand[:32]                                   # the chunked top-32 (only reached with > 32 candidates left)
        tv31 = u[cand[31]] if len(cand) == 32 else NEG     # lane 31 of the sorted list
        blocked = [False] * nb_
        for bi, b in enumerate(beam):
            if blocked[bi]:
                continue
            bottom = L[W - 1][0] if len(L) == W else NEG
            if not (b["t"] > bottom):
                break
            lab_r = b["pre"][-1] if b["pre"] else -1
            kids = {surv[i]["pre"][-1]: i for i in range(nb_) if surv[i]["q"] == bi}
            lab_ok = lab_r >= 0 and lab_r not in kids
            valid = [k for k in cand if k != lab_r and k not in kids]
            xs = {k: np.float32(u[k] + b["t"]) for k in valid}
            vlab = np.float32(u[lab_r] + b["b"]) if lab_ok else NEG
            x31 = np.float32(tv31 + b["t"])
            for k_c, ci in kids.items():
                in_L = [j for j, it in enumerate(L) if it[1] == "s" and it[2] == ci]
                if not in_L:
                    blocked[ci] = True
                    continue
                rank, s_c = in_L[0], surv[ci]["t"]
                cnt = rank + sum(1 for k in valid if k < k_c and xs[k] > s_c) + (1 if (lab_r < k_c and vlab > s_c) else 0)
                counted = cnt >= W or (s_c >= list_floor and (NC <= 32 or not (x31 > s_c)))
                STATS["block_tests"] += 1
                if not counted:
                    STATS["block_fallback"] += 1
                    cnt = rank
                    for k in range(k_c):
                        if k in kids:
                            continue
                        xk = np.float32(u[k] + (b["b"] if k == lab_r else b["t"]))
                        cnt += 1 if xk > s_c else 0
                if cnt >= W:
                    blocked[ci] = True

            def insert(v, pay):
                pos = sum(1 for it in L if it[0] > v)
                L.insert(pos, (v, "c", pay))
                if len(L) > W:
                    L.pop()
            if lab_ok:
                bottom = L[W - 1][0] if len(L) == W else NEG
                if vlab > bottom:
                    insert(vlab, (bi, lab_r))
            bottom = L[W - 1][0] if len(L) == W else NEG
            live = [k for k in valid if xs[k] > bottom]                   # list order = sorted order
            old = [it[0] for it in L] + [NEG] * (W + 1)
            n_ins = 0
            for j, k in enumerate(live):
                if j < W and xs[k] > old[W - 1 - j]:
                    n_ins += 1
                else:
                    break
            STATS["inserted"] += n_ins
            for k in live[:n_ins]:
                insert(xs[k], (bi, k))
        nbeam = []
        for tot, kind, pay in L:
            if kind == "s":
                s = surv[pay]
                nbeam.append({"pre": s["pre"], "b": s["b"], "l": s["l"], "t": s["t"]})
            else:
                bi, k = pay
                nbeam.append({"pre": beam[bi]["pre"] + (k,), "b": NEG, "l": tot, "t": tot})
        beam = nbeam
        if trace is not None:
            trace.append([(e["pre"], float(e["t"])) for e in beam])
    best = beam[0]["pre"]
    out, prev = [], -1
    for k in reversed(best):
        if not merge_repeated or k != prev:
            out.append(int(k))
        prev = k
    return out[::-1]
