#!/usr/bin/env python
"""Tiny driver for ncu: builds the B=64, 128x32 GRU model and runs a few training steps (no timing, no oracle)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import crnn_b200 as cb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cell = sys.argv[2] if len(sys.argv) > 2 else "gru"
m = cb.CRNN(bench.V, bench.MAXLEN, (bench.IMGH, bench.IMGW, 1), 128, cell == "gru", 256, max_batch=bench.BATCH, seed=1).get_model()
m.compile(optimizer=cb.Adam(lr=1e-4, beta_1=0.5, beta_2=0.999, epsilon=1e-7, clipnorm=5.0))
x, lab, L, il = bench.synth_batch(bench.BATCH, 2)
d = "cuda"
args = [torch.tensor(a, device=d) for a in (x, lab, L, il)]
for s in range(steps):
    m.train_fwd_bwd_device(*args, dropout_seed=100 + s)
    m.optimizer_step()
torch.cuda.synchronize()
sm = m.forward_device(args[0])
cb.ctc_decode_device(sm, greedy=False, beam_width=10)
torch.cuda.synchronize()
print("done")
