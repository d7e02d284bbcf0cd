#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM kernels at the conv-stack shapes of the B=64, 128x32 step (CUDA events, rotating buffers
larger than L2).  Each crnn_gemm_tc call = weight-image prep kernel + GEMM; crnn_gemm_tc_dw = split-K dW kernel (dW pre-zeroed)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import crnn_b200 as cb

lib = cb._lib.load()
PIX = [64 * 132 * 36, 64 * 132 * 36, 64 * 66 * 18, 64 * 66 * 18, 64 * 66 * 9, 64 * 66 * 9]
CH = [(64, 128), (128, 256), (256, 256), (256, 512), (512, 512), (512, 512)]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10


def timeit(fn, n):
    fn(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000


tot = {"fwd": 0.0, "dx": 0.0, "dw": 0.0}
for li, (M, (ci, co)) in enumerate(zip(PIX, CH), 2):
    nb = max(2, int(300e6 // (M * max(ci, co) * 4)) + 1)
    X = [torch.randn(M, ci, device="cuda") for _ in range(nb)]
    dY = [torch.randn(M, co, device="cuda") for _ in range(nb)]
    out = torch.empty(M, co, device="cuda"); dX = torch.empty(M, ci, device="cuda")
    W = torch.randn(ci, co, device="cuda") * 0.05
    sc = torch.rand(ci, device="cuda") + 0.5; sh = torch.randn(ci, device="cuda")
    stats = torch.zeros(2 * co, dtype=torch.float64, device="cuda")
    scr = torch.empty(lib.crnn_gemm_tc_scratch_floats(max(ci, co), max(ci, co)), device="cuda")
    dW = torch.zeros(ci, co, device="cuda")
    fwd = lambda i: cb._lib.check(lib.crnn_gemm_tc(X[i % nb].data_ptr(), ci, W.data_ptr(), co, 1, out.data_ptr(), co, M, co, ci,
                                                   sc.data_ptr(), sh.data_ptr(), stats.data_ptr(), scr.data_ptr(), st))
    dx = lambda i: cb._lib.check(lib.crnn_gemm_tc(dY[i % nb].data_ptr(), co, W.data_ptr(), co, 0, dX.data_ptr(), ci, M, ci, co,
                                                  None, None, None, scr.data_ptr(), st))
    dw = lambda i: cb._lib.check(lib.crnn_gemm_tc_dw(X[i % nb].data_ptr(), ci, ci, dY[i % nb].data_ptr(), co, co, dW.data_ptr(), co, M,
                                                     sc.data_ptr(), sh.data_ptr(), st))
    t = {k: timeit(f, reps) for k, f in (("fwd", fwd), ("dx", dx), ("dw", dw))}
    fl = 2.0 * M * ci * co
    hbm_f = (M * (ci + co) * 4) / 6.5e12 * 1e6
    print(f"block{li}: M={M} {ci}->{co}  fwd {t['fwd']:.1f} us  dx {t['dx']:.1f} us  dw {t['dw']:.1f} us   | 3xTF32 floor {3 * fl / 1.1e15 * 1e6:.0f} us, HBM floor fwd {hbm_f:.0f} us")
    for k in t:
        tot[k] += t[k]
    del X, dY
print("total", {k: round(v) for k, v in tot.items()})
