#!/usr/bin/env python3
"""Host input pipeline throughput: Readf.run_generator on synthetic mjsynth-like JPEG word crops, sequential (the
reference's behaviour: one Python thread doing cv2 read / pad / resize per image) vs the threaded loader (workers=N; bit-identical batches).
CPU only."""
import os, sys, tempfile, time
import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crnn_b200 as cb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = tempfile.mkdtemp()
rng = np.random.default_rng(0)
names = []
for i in range(N):
    h, w = int(rng.integers(24, 40)), int(rng.integers(60, 200))          # mjsynth crops are ~32 px high, 60-200 px wide
    img = np.full((h, w, 3), 255, np.uint8)
    cv2.putText(img, "word%d" % (i % 97), (3, h - 8), cv2.FONT_HERSHEY_SIMPLEX, 0.7, (int(rng.integers(0, 90)),) * 3, 2)
    p = os.path.join(d, "%d_word%d_%d.jpg" % (i, i % 97, i)); cv2.imwrite(p, img); names.append(p)
classes = {c: i for i, c in enumerate(cb.get_lexicon())}
for workers in (0, 2, 4, 8, 16):
    np.random.seed(1)
    g = cb.Readf(img_size=(100, 32, 1), max_len=23, normed=True, batch_size=64, classes=classes, transform_p=0.7, workers=workers, device_norm=True).run_generator(names)
    next(g)                                                                 # warm-up batch (thread pool start, file cache)
    t0 = time.perf_counter()
    for _ in range(N // 64 - 1):
        next(g)
    dt = time.perf_counter() - t0
    print("workers=%2d  %7.0f images/s" % (workers, (N - 64) / dt))
