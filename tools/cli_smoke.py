#!/usr/bin/env python
"""End-to-end smoke of the drop-in command lines on a GPU box: synthetic word images -> `train.py` (1 epoch, Adam) -> the files the reference
writes (train.py:125-127,169,182-183,211,215-216) -> `predict.py --validate` on the trained directory -> prediction.csv + edit distances."""
import os, subprocess, sys, tempfile
import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = tempfile.mkdtemp()
img_dir = os.path.join(d, "imgs"); os.makedirs(img_dir)
rng = np.random.default_rng(0)
words = ["hello", "world", "ocr", "lite", "b200", "crnn", "text", "line"]
for i in range(40):
    w = words[i % len(words)]
    img = np.full((32, 100), 255, np.uint8)
    cv2.putText(img, w, (2, 24), cv2.FONT_HERSHEY_SIMPLEX, 0.8, int(rng.integers(0, 80)), 2)
    cv2.imwrite(os.path.join(img_dir, "%d_%s_%d.png" % (i, w, i)), img)
env = dict(os.environ, PYTHONPATH=ROOT)
run = lambda cmd: subprocess.run([sys.executable] + cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=240)
r = run(["train.py", "--G", "0", "--path", img_dir, "--save_path", d, "--model_name", "m", "--nbepochs", "1", "--batch_size", "8", "--opt", "adam",
         "--lr", "0.0001", "--norm", "--GRU", "--imgh", "100", "--imgW", "32", "--train_portion", "0.8"])
print(r.stdout[-600:]); print(r.stderr[-600:])
assert r.returncode == 0, "train.py failed"
out = os.path.join(d, "m")
for f in ("arguments.txt", "model.json", "model_summary.txt", "loss_history.pickle.dat", "final_weights.h5", "final_model.h5"):
    assert os.path.exists(os.path.join(out, f)), f
r = run(["predict.py", "--G", "0", "--model_path", out, "--image_path", img_dir, "--result_path", d, "--validate", "--batch_size", "8", "--max_len", "23"])
print(r.stdout[-700:]); print(r.stderr[-400:])
assert r.returncode == 0, "predict.py failed"
assert os.path.exists(os.path.join(d, "prediction.csv")) and "mean edit distance" in r.stdout
print("CLI SMOKE OK")
