#!/usr/bin/env python3
"""Numerical model of the products used on the tensor cores: relative error of a K-term dot product for
  fp32 FMA accumulation | 1xTF32 | 3xTF32 (hi.hi + hi.lo + lo.hi) | the 2-instruction scheme (tf32 hi.hi + bf16 cross terms)
with exact (fp64) accumulation of the rounded products, so that only the operand roundings are compared.  CPU only (numpy)."""
import numpy as np


def tf32(x):                                   # round to nearest (ties away), 10 explicit mantissa bits -- cvt.rna.tf32.f32
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return b.view(np.float32)


def bf16_rn(x):                                # round to nearest even, 7 explicit mantissa bits -- cvt.rn.bf16x2.f32
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return b.view(np.float32)


def bf16_trunc(x):                             # upper 16 bits of the (tf32-rounded) register -- the PRMT used for a_hi'
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def run(K, trials=2000, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((trials, K)).astype(np.float32)
    b = rng.standard_normal((trials, K)).astype(np.float32)
    exact = (a.astype(np.float64) * b.astype(np.float64)).sum(1)
    scale = np.sqrt((a.astype(np.float64) ** 2 * b.astype(np.float64) ** 2).sum(1))          # ~ |a||b| sqrt(K): size of the summands
    ah, bh = tf32(a), tf32(b)
    al, bl = (a - ah).astype(np.float32), (b - bh).astype(np.float32)
    d = lambda x, y: (x.astype(np.float64) * y.astype(np.float64)).sum(1)
    res = {}
    acc = np.zeros(trials, np.float32)
    for k in range(K):
        acc = (acc + a[:, k] * b[:, k]).astype(np.float32)                                       # fp32 multiply-add chain (FFMA baseline, roughly)
    res["fp32 accumulate (FFMA chain)"] = acc.astype(np.float64)
    res["1xTF32"] = d(ah, bh)
    res["3xTF32 (hi.hi + hi.lo + lo.hi, lo truncated to tf32 by the MMA)"] = d(ah, bh) + d(ah, tf32(bl)) + d(tf32(al), bh)
    res["2 MMAs (tf32 hi.hi + bf16 {a_hi'.b_lo + a_lo.b_hi'})"] = d(ah, bh) + d(bf16_trunc(ah), bf16_rn(bl)) + d(bf16_rn(al), bf16_rn(bh))
    out = {}
    for k, v in res.items():
        e = np.abs(v - exact) / scale
        out[k] = (float(np.sqrt((e ** 2).mean())), float(e.max()))
    return out


if __name__ == "__main__":
    for K in (256, 512, 4608):
        print("K = %d   (error relative to |a||b|sqrt(K), rms / max over 2000 random N(0,1) dot products)" % K)
        for k, (rms, mx) in run(K).items():
            print("   %-72s rms %.2e   max %.2e" % (k, rms, mx))
