#!/usr/bin/env python
"""Generates the committed golden fixtures (run in the authoring container, where /root/reference is mounted):

  shipped_mjsynth_weights.npz   the reference's trained weights models/OCR_mjsynth_FULL_2/final_weights.h5, read with
                                our own HDF5 reader (94 tensors, 2 831 027 fp32) -- the GPU box has no /root/reference
  shipped_mjsynth_golden.npz    oracle outputs for 8 seeded synthetic 100x32 images through those weights:
                                theta, STN output, per-block checksums, softmax (8,52,38), greedy / beam-10 label ids
  ctc_golden.npz                CTC loss / gradient for (8,50,38) random probs with labels of length 1,2,5,23 (repeats),
                                beam-10 / greedy decode of (64,25,96) random softmax

The oracle itself is "parity unpinned" (no reference-run outputs exist); these fixtures pin the CUDA path AND the oracle
against silent drift, and carry the only trained weights available.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import crnn_b200 as cb  # noqa: E402
from oracle import crnn_oracle as N  # noqa: E402
from oracle import ctc_oracle as O  # noqa: E402

REF = "/root/reference/models/OCR_mjsynth_FULL_2/final_weights.h5"


def main():
    w = cb.hdf5_lite.load_keras_weights(REF)
    np.savez_compressed(os.path.join(HERE, "shipped_mjsynth_weights.npz"), **{k.replace("/", "__"): v for k, v in w.items()})
    cfg = N.Cfg(imgh=100, cell="gru")
    rng = np.random.default_rng(2024)
    # text-like synthetic lines: blurred strokes on a dark background, normalised like the reference generator (utils.py:421)
    x = np.zeros((8, 100, 32), np.float32)
    for b in range(8):
        img = np.zeros((100, 32), np.float32)
        for _ in range(14):
            c0, r0 = rng.integers(4, 96), rng.integers(6, 26)
            img[max(0, c0 - 2):c0 + 3, max(0, r0 - 5):r0 + 6] = rng.uniform(120, 255)
        x[b] = img
    x = ((x + rng.uniform(0, 30, x.shape).astype(np.float32) - 118.24236953981779) / 36.72835353999682).astype(np.float32)[..., None]
    keep = N.forward(N.to_torch(w), torch.tensor(x), cfg, training=False)
    sm = keep["softmax"].numpy()
    g_out, g_n, _ = O.greedy(sm)
    b_out, b_n, b_lp = O.beam(sm, beam_width=10, merge_repeated=True)
    gold = {"x": x, "theta": keep["theta"].numpy(), "stn": keep["stn"].numpy(), "softmax": sm,
            "greedy": g_out, "greedy_len": g_n, "beam": b_out, "beam_len": b_n, "beam_logprob": b_lp}
    for i in range(1, 8):
        a = keep[f"block{i}"].numpy()
        gold[f"block{i}_sum"] = np.array([a.sum(dtype=np.float64), np.abs(a).sum(dtype=np.float64), (a.astype(np.float64) ** 2).sum()])
    np.savez_compressed(os.path.join(HERE, "shipped_mjsynth_golden.npz"), **gold)
    lex = N.__dict__.get("LEX", "0123456789abcdefghijklmnopqrstuvwxyz-")
    print("beam strings:", ["".join(lex[k] for k in b_out[i, :b_n[i]]) for i in range(8)])

    # CTC fixtures
    z = rng.standard_normal((8, 50, 38)).astype(np.float32) * 2
    p = torch.softmax(torch.tensor(z), -1).numpy()
    lens = np.array([1, 2, 5, 23, 1, 2, 5, 23], np.int32)
    labels = np.full((8, 23), 37, np.int32)
    for b in range(8):
        lab = rng.integers(0, 37, lens[b])
        if lens[b] >= 2:
            lab[1] = lab[0]
        labels[b, :lens[b]] = lab
    in_len = np.full(8, 50, np.int32)
    loss, grad = O.ctc_loss_grad(p, labels, lens, in_len)
    z2 = rng.standard_normal((64, 25, 96)).astype(np.float32) * 3
    p2 = torch.softmax(torch.tensor(z2), -1).numpy()
    bo, bn, bl = O.beam(p2, beam_width=10)
    go, gn, gs = O.greedy(p2)
    np.savez_compressed(os.path.join(HERE, "ctc_golden.npz"), probs=p, labels=labels, label_len=lens, input_len=in_len, loss=loss, grad_u=grad,
                        probs2=p2, beam=bo, beam_len=bn, beam_logprob=bl, greedy=go, greedy_len=gn, greedy_score=gs)
    print("written:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
