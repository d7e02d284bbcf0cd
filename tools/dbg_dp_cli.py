#!/usr/bin/env python
"""Debug driver for tests/test_gpu_dp.py::test_dp_train_cli_uneven_shards_and_early_stopping: runs the same torchrun train.py job N times and
reports where the two replicas' final parameters differ (tensor name, count, max |diff|)."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
extra = dict(a.split("=", 1) for a in sys.argv[2:])
for it in range(n):
    tmp = tempfile.mkdtemp()
    img_dir = os.path.join(tmp, "imgs"); os.makedirs(img_dir)
    rng = np.random.default_rng(0)
    words = ["hello", "world", "ocr", "lite", "b200", "crnn", "text", "line"]
    for i in range(19):
        w = words[i % len(words)]
        img = np.full((32, 100), 255, np.uint8)
        cv2.putText(img, w, (2, 24), cv2.FONT_HERSHEY_SIMPLEX, 0.8, int(rng.integers(0, 80)), 2)
        cv2.imwrite(os.path.join(img_dir, "%d_%s_%d.png" % (i, w, i)), img)
    import torch
    env = dict(os.environ, CRNN_DP_DUMP_PARAMS=os.path.join(tmp, "params"), **extra)
    if torch.cuda.device_count() < 2:
        env["CRNN_DIST_BACKEND"] = "gloo"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(29700 + it),
                        "train.py", "--path", img_dir, "--save_path", tmp, "--model_name", "m", "--nbepochs", "2", "--batch_size", "8", "--opt", "sgd", "--lr", "0.001",
                        "--imgh", "100", "--imgW", "32", "--train_portion", "0.9", "--early_stopping", "3"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    if r.returncode != 0:
        print("run", it, "rc", r.returncode, r.stdout[-800:], r.stderr[-1500:]); continue
    a = np.load(os.path.join(tmp, "params.rank0.npy")); b = np.load(os.path.join(tmp, "params.rank1.npy"))
    print("\n".join(l for l in (r.stdout + r.stderr).splitlines() if "dbg-nan" in l)[:3000])
    d = np.flatnonzero(a != b)
    print("run %d: %d of %d parameters differ, max |diff| %.3e%s" % (it, d.size, a.size, np.abs(a - b).max(), "" if d.size == 0 else ", first at %d, last at %d" % (d[0], d[-1])))
    if d.size:
        print("  sample:", [(int(i), float(a[i]), float(b[i])) for i in d[:5]])
        print("  losses rank0 tail:", [l for l in r.stdout.splitlines() if "loss" in l][-3:])
