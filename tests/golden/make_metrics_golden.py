#!/usr/bin/env python3
"""Golden vectors for the evaluation step: runs the REFERENCE's own levenshtein / edit_distance /
normalized_edit_distance (extracted from /root/reference/utils.py:262-298 by AST -- the module itself cannot be imported here, it
pulls in keras/tensorflow at import time) on seeded string pairs and writes tests/golden/metrics_golden.json.
Run in the build container only (the GPU box has no /root/reference); the JSON is committed."""
import ast, json, os, random
import numpy as np

REF = "/root/reference/utils.py"
src = open(REF).read()
mod = ast.parse(src)
want = {"levenshtein", "edit_distance", "normalized_edit_distance"}
code = ast.Module(body=[n for n in mod.body if isinstance(n, ast.FunctionDef) and n.name in want], type_ignores=[])
ns = {"np": np}
exec(compile(code, REF, "exec"), ns)

rng = random.Random(20260117)
alphabet = "abcdefghijklmnopqrstuvwxyz0123456789"
def word(lo, hi): return "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))
def mutate(w):
    w = list(w)
    for _ in range(rng.randint(0, 4)):
        op = rng.random()
        if op < 0.34 and w: w[rng.randrange(len(w))] = rng.choice(alphabet)
        elif op < 0.67 and w: del w[rng.randrange(len(w))]
        else: w.insert(rng.randint(0, len(w)), rng.choice(alphabet))
    return "".join(w)[:23]
pairs = []
for _ in range(300):
    t = word(1, 23)
    pairs.append([mutate(t), t])                # near misses, like OCR output
for _ in range(60):
    pairs.append([word(0, 23), word(1, 23)])    # unrelated strings, empty predictions
pairs += [["", "a"], ["abc", "abc"], ["kitten", "sitting"], ["celist", "cellist"], ["a" * 23, "b" * 23]]
pred = [p for p, _ in pairs]; true = [t for _, t in pairs]
out = {"pairs": pairs, "levenshtein": [float(ns["levenshtein"](p, t)) for p, t in pairs],
       "edit_distance": float(ns["edit_distance"](pred, true)), "normalized_edit_distance": float(ns["normalized_edit_distance"](pred, true)),
       "source": "reference utils.py:262-298 executed by tests/golden/make_metrics_golden.py"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics_golden.json"), "w"))
print(len(pairs), "pairs; ed", out["edit_distance"], "ned", out["normalized_edit_distance"])
