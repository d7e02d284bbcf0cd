"""Second, independent statement of the prefix beam search (TEST INFRASTRUCTURE ONLY).

Pure-Python 'score all candidates, keep the global top-W' formulation over prefix tuples -- the
formulation the CUDA kernel uses -- in float32 with the same LogSumExp as TF (SURVEY A.2 note ii).
Used to check that it agrees with the literal sequential restatement in ctc_oracle.c.
"""
import numpy as np

NEG = np.float32(-np.inf)


def _lse(a, b):
    a = np.float32(a); b = np.float32(b)
    if a == NEG and b == NEG:
        return NEG
    hi, lo = (a, b) if a > b else (b, a)
    return np.float32(hi + np.log1p(np.exp(np.float32(lo - hi), dtype=np.float32), dtype=np.float32))


def beam_global_topk(p, W, merge_repeated, eps=1e-7):
    T, V = p.shape
    blank = V - 1
    beam = {(): (np.float32(0), NEG, np.float32(0))}  # prefix -> (blank, label, total) at t-1
    for t in range(T):
        u = np.log(p[t].astype(np.float32) + np.float32(eps), dtype=np.float32)
        u = (u - u.max()).astype(np.float32)
        cand = {}
        for pre, (pb, pl, pt) in beam.items():
            nl = pl
            if pre:
                par = pre[:-1]
                if par in beam:
                    qb, ql, qt = beam[par]
                    prev = qb if (len(par) and par[-1] == pre[-1]) else qt
                    nl = _lse(pl, prev)
                nl = np.float32(nl + u[pre[-1]])
            nb = np.float32(pt + u[blank])
            cand[pre] = (nb, nl, _lse(nb, nl))
        for pre, (pb, pl, pt) in beam.items():
            for k in range(V - 1):
                ch = pre + (k,)
                if ch in beam:
                    continue
                prev = pb if (pre and pre[-1] == k) else pt
                s = np.float32(u[k] + prev)
                if s > NEG:
                    cand[ch] = (NEG, s, s)
        top = sorted(cand.items(), key=lambda kv: -kv[1][2])[:W]
        beam = dict(top)
    best = max(beam.items(), key=lambda kv: kv[1][2])[0]
    out, prev = [], -1
    for k in best:
        if not merge_repeated or k != prev:
            out.append(int(k))
        prev = k
    return out
