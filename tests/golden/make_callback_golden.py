#!/usr/bin/env python3
"""Golden traces for the training-driver policy: runs the REFERENCE's own EarlyStoppingIter (utils.py:535-614), extracted by
AST from /root/reference/utils.py and executed with a stub keras Callback base class (keras itself is not installable here) and a numpy
shim that still has `np.Inf` (removed in numpy 2), on seeded loss sequences.  Writes tests/golden/callback_golden.json.
Run in the build container only; the JSON is committed."""
import ast, json, os, random, types, warnings
import numpy as np

REF = "/root/reference/utils.py"
mod = ast.parse(open(REF).read())
cls = [n for n in mod.body if isinstance(n, ast.ClassDef) and n.name == "EarlyStoppingIter"]

class Callback:                       # what keras.callbacks.Callback provides to this class: nothing but the attribute slot
    def __init__(self): self.model = None

npx = types.SimpleNamespace(less=np.less, greater=np.greater, Inf=np.inf)
ns = {"np": npx, "Callback": Callback, "warnings": warnings}
exec(compile(ast.Module(body=cls, type_ignores=[]), REF, "exec"), ns)
Ref = ns["EarlyStoppingIter"]

class Model:
    def __init__(self): self.stop_training = False; self.w = 0; self.sets = []
    def get_weights(self): return self.w
    def set_weights(self, w): self.sets.append(w); self.w = w

rng = random.Random(4711)
cases = []
def run(kwargs, losses, key="loss"):
    cb = Ref(**kwargs); m = Model(); cb.model = m
    cb.on_train_begin()
    trace = []
    for i, v in enumerate(losses):
        m.w = i                                          # "weights" = iteration index, so restore_best_weights is observable
        cb.on_batch_end(i, {key: v} if v is not None else {})
        trace.append([bool(m.stop_training), float(cb.best), int(cb.stopped_iter), int(cb.cycle_iterations), float(cb.sum_monitor)])
        if m.stop_training:
            break
    cb.on_train_end()
    cases.append({"kwargs": kwargs, "key": key, "losses": losses, "trace": trace, "sets": m.sets, "final_w": m.w})

def seq(n, trend, noise): return [max(0.0, 10.0 * (trend ** i) + rng.uniform(-noise, noise)) for i in range(n)]
run({"patience": 5}, seq(60, 0.97, 0.05))
run({"patience": 5}, seq(60, 1.00, 0.5))
run({"patience": 7, "min_delta": 0.05}, seq(80, 0.99, 0.2))
run({"patience": 3, "restore_best_weights": True}, seq(40, 0.95, 0.1) + seq(20, 1.05, 0.1))
run({"patience": 4, "mode": "max", "monitor": "acc"}, [0.1 * i for i in range(10)] + [0.5] * 20, key="acc")
run({"patience": 4, "monitor": "val_acc"}, [0.3, 0.4, 0.5, 0.6, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2], key="val_acc")
run({"patience": 2, "baseline": 5.0}, [6.0, 6.0, 6.0, 4.0, 4.0])
run({"patience": 5}, [None, None, 3.0, 2.0, None, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 5.0, 5.0, 5.0, 5.0, 5.0, 9.0])   # missing monitor key
run({"patience": 1}, [3.0, 2.0, 1.5, 1.6, 1.0])
json.dump({"cases": cases, "source": "reference utils.py:535-614 executed by tests/golden/make_callback_golden.py"},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "callback_golden.json"), "w"))
print(len(cases), "cases;", [len(c["trace"]) for c in cases])
