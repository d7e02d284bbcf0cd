#!/usr/bin/env python3
"""Golden vectors for the host input pipeline: runs the REFERENCE's own open_img / norm / get_lexicon / parse_mjsynth
(utils.py:359-416, extracted by AST from /root/reference/utils.py; cv2, numpy and PIL are available here, keras/tensorflow are not) on
seeded synthetic word crops with seeded np.random, and writes tests/golden/pipeline_golden.npz.  Run in the build container only."""
import ast, os, string
import cv2
import numpy as np
from PIL import Image

REF = "/root/reference/utils.py"
mod = ast.parse(open(REF).read())
from tqdm import tqdm
want = {"read_img", "open_img", "norm", "get_lexicon", "parse_mjsynth", "labels_to_text", "get_lengths", "make_ohe"}
ns = {"np": np, "cv2": cv2, "Image": Image, "os": os, "string": string, "tqdm": tqdm}
body = [n for n in mod.body if (isinstance(n, ast.FunctionDef) and n.name in want) or (isinstance(n, ast.ClassDef) and n.name == "DecodeCTCPred")]
exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)

rng = np.random.default_rng(5)
out = {}
n = 0
for k in range(14):
    h = int(rng.integers(8, 40)); w = int(rng.integers(10, 120))
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    if k % 2:
        img = (255 - (img // 4)).astype(np.uint8)            # bright background -> exercises the inversion branch
    if k == 13:
        img = rng.integers(0, 256, (12, 40), dtype=np.uint8)  # small crop -> exercises the 1.5x up-scaling branch
    for p in (0.0, 0.7):
        for size in ((100, 32), (128, 32)):
            np.random.seed(1000 + k)
            res, _ = ns["open_img"](img.copy(), size, p=p)
            out["in_%d" % n] = img; out["p_%d" % n] = np.float64(p); out["size_%d" % n] = np.array(size); out["seed_%d" % n] = np.int64(1000 + k)
            out["out_%d" % n] = np.asarray(res)
            n += 1
out["n"] = np.int64(n)
x = rng.integers(0, 256, (7, 9), dtype=np.uint8)
out["norm_in"] = x
out["norm_out"] = ns["norm"](x, 118.24236953981779, 36.72835353999682)
out["lexicon_default"] = np.array(sorted(ns["get_lexicon"]()))
out["lexicon_non_intersecting"] = np.array(sorted(ns["get_lexicon"](non_intersecting_chars=True)))
out["mjsynth"] = np.array(ns["parse_mjsynth"]("/data/mj", ["./2194/2/334_EFFLORESCENT_24742.jpg 24742", "./3000/7/1_a_1.jpg 1"]))
# label <-> text helpers (utils.py:314-345, 518-522): blank (= len(inverse_classes)) and -1 padding are dropped
lex = ns["get_lexicon"]()
inv = {i: c for i, c in enumerate(lex)}
lab = np.array([[17, 14, 21, 21, 24, 37, -1, -1], [37, 37, 0, 9, 36, 37, 1, -1], [-1, -1, -1, -1, -1, -1, -1, -1]])
out["l2t_labels"] = lab
out["l2t_text_fn"] = np.array([ns["labels_to_text"](r, inverse_classes=inv) for r in lab])
out["l2t_text_cls"] = np.array([ns["DecodeCTCPred"](top_paths=1, beam_width=3, inverse_classes=inv).labels_to_text(r) for r in lab])
gl = ns["get_lengths"](["/a/b/12_hello_3.png", "7_x_1.jpg", "/q/0_abcdefghij_99.png"])
out["get_lengths_keys"] = np.array(list(gl.keys())); out["get_lengths_vals"] = np.array(list(gl.values()))
out["ohe_in"] = np.array([3, 0, 5, 5, 1]); out["ohe_out"] = ns["make_ohe"](out["ohe_in"], 6)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pipeline_golden.npz"), **out)
print(n, "open_img cases")
