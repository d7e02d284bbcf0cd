#!/usr/bin/env python
"""Tiny driver for ncu: BASELINE configs[3] beam-10 decode, (4096,25,96) softmax of N(0,1)*3 logits, on the device."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import crnn_b200 as cb

rng = np.random.default_rng(3)
z = torch.tensor(rng.standard_normal((4096, 25, 96)).astype(np.float32) * 3, device="cuda")
p = torch.softmax(z, -1).contiguous()
for _ in range(3):
    out = cb.ctc_decode_device(p, greedy=False, beam_width=10)
torch.cuda.synchronize()
print("done")
