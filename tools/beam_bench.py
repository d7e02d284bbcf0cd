#!/usr/bin/env python
"""Device-timed beam-10 decode of BASELINE configs[3] ((4096,25,96) softmax of N(0,1)*3 logits) and of the model-shaped case (4096,66,38)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import crnn_b200 as cb

for (B, T, V, scale) in ((4096, 25, 96, 3.0), (4096, 66, 38, 3.0), (4096, 52, 38, 8.0)):
    rng = np.random.default_rng(3)
    p = torch.softmax(torch.tensor(rng.standard_normal((B, T, V)).astype(np.float32) * scale, device="cuda"), -1).contiguous()
    for _ in range(3):
        cb.ctc_decode_device(p, greedy=False, beam_width=10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        cb.ctc_decode_device(p, greedy=False, beam_width=10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("beam-10 (%d,%d,%d) scale %.0f: %.3f ms  %.2f M lines/s" % (B, T, V, scale, ms, B / ms / 1e3))
