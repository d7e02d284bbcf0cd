#!/usr/bin/env python3
"""Golden graph structure: executes the REFERENCE's own graph-assembly source -- CRNN.__init__ / depthwise_conv_block /
get_model, STN, get_initial_weights (utils.py:32-96, 239-258), extracted by AST from /root/reference/utils.py -- over a RECORDING shim of the
Keras layer classes it instantiates (keras is not installable here).  Every shim layer only does Keras' shape inference and appends
(layer class, constructor arguments, output shape) to a trace; no arithmetic.  Writes tests/golden/graph_golden.json with the traces for
both geometries, both cells and two vocabulary sizes.  Run in the build container only."""
import ast, json, os
import numpy as np

TRACE = []


class T:                                     # symbolic tensor: batch-less shape
    def __init__(self, shape): self.shape = tuple(shape)
    def __getitem__(self, idx): raise NotImplementedError


def _rec(kind, cfg, out):
    TRACE.append({"layer": kind, "config": cfg, "shape": list(out.shape)})
    return out


def _pair(v): return (v, v) if isinstance(v, int) else tuple(v)


class Layer:
    def __init__(self, **kw): self.kw = kw


def Input(name=None, shape=None, dtype=None): return _rec("Input", {"name": name, "dtype": dtype}, T(shape))


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1)): self.p = _pair(padding)
    def __call__(self, x): h, w, c = x.shape; return _rec("ZeroPadding2D", {"padding": list(self.p)}, T((h + 2 * self.p[0], w + 2 * self.p[1], c)))


class DepthwiseConv2D(Layer):
    def __init__(self, kernel_size, padding="valid", strides=(1, 1), depth_multiplier=1, use_bias=True):
        self.k, self.pad, self.s, self.dm, self.bias = _pair(kernel_size), padding, _pair(strides), depth_multiplier, use_bias
    def __call__(self, x):
        h, w, c = x.shape
        assert self.pad == "same" and self.s == (1, 1)
        return _rec("DepthwiseConv2D", {"kernel_size": list(self.k), "padding": self.pad, "strides": list(self.s), "depth_multiplier": self.dm, "use_bias": self.bias}, T((h, w, c * self.dm)))


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", use_bias=True):
        self.f, self.k, self.s, self.pad, self.bias = filters, _pair(kernel_size), _pair(strides), padding, use_bias
    def __call__(self, x):
        h, w, c = x.shape
        if self.pad == "valid": h, w = (h - self.k[0]) // self.s[0] + 1, (w - self.k[1]) // self.s[1] + 1
        return _rec("Conv2D", {"filters": self.f, "kernel_size": list(self.k), "strides": list(self.s), "padding": self.pad, "use_bias": self.bias}, T((h, w, self.f)))


class BatchNormalization(Layer):
    def __init__(self, axis=-1): self.axis = axis
    def __call__(self, x): return _rec("BatchNormalization", {"axis": self.axis}, x)


class ReLU(Layer):
    def __init__(self, max_value=None): self.m = max_value
    def __call__(self, x): return _rec("ReLU", {"max_value": self.m}, x)


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2)): self.p = _pair(pool_size)
    def __call__(self, x): h, w, c = x.shape; return _rec("MaxPooling2D", {"pool_size": list(self.p)}, T((h // self.p[0], w // self.p[1], c)))


MaxPool2D = MaxPooling2D


class Dropout(Layer):
    def __init__(self, rate): self.r = rate
    def __call__(self, x): return _rec("Dropout", {"rate": self.r}, x)


class Reshape(Layer):
    def __init__(self, target_shape=None, name=None): self.t, self.name = tuple(target_shape), name
    def __call__(self, x):
        assert int(np.prod(x.shape)) == int(np.prod(self.t)), (x.shape, self.t)
        return _rec("Reshape", {"target_shape": list(self.t), "name": self.name}, T(self.t))


class Flatten(Layer):
    def __init__(self): pass
    def __call__(self, x): return _rec("Flatten", {}, T((int(np.prod(x.shape)),)))


class Dense(Layer):
    def __init__(self, units, activation=None, name=None, kernel_initializer="glorot_uniform", weights=None):
        self.u, self.act, self.name, self.init, self.w = units, activation, name, kernel_initializer, weights
    def __call__(self, x):
        cfg = {"units": self.u, "activation": self.act, "name": self.name, "kernel_initializer": self.init}
        if self.w is not None: cfg["weights"] = [np.asarray(a).tolist() for a in self.w]
        return _rec("Dense", cfg, T(x.shape[:-1] + (self.u,)))


class Activation(Layer):
    def __init__(self, activation, name=None): self.a, self.name = activation, name
    def __call__(self, x): return _rec("Activation", {"activation": self.a, "name": self.name}, x)


class _RNN(Layer):
    kind = "RNN"
    def __init__(self, units, return_sequences=False, kernel_initializer="glorot_uniform"): self.u, self.rs, self.init = units, return_sequences, kernel_initializer


class LSTM(_RNN): kind = "LSTM"
class GRU(_RNN): kind = "GRU"


class Bidirectional(Layer):
    def __init__(self, layer, merge_mode="concat", weights=None): self.l, self.m = layer, merge_mode
    def __call__(self, x):
        t, _ = x.shape
        assert self.l.rs
        out = self.l.u * (2 if self.m == "concat" else 1)
        return _rec("Bidirectional", {"cell": self.l.kind, "units": self.l.u, "return_sequences": self.l.rs, "kernel_initializer": self.l.init, "merge_mode": self.m}, T((t, out)))


class BilinearInterpolation(Layer):
    def __init__(self, output_size=(100, 32)): self.o = tuple(output_size)
    def __call__(self, xs): img, th = xs; assert th.shape == (6,); return _rec("BilinearInterpolation", {"output_size": list(self.o)}, T(self.o + (img.shape[-1],)))


class Lambda(Layer):
    def __init__(self, fn, output_shape=None, name=None): self.fn, self.o, self.name = fn, output_shape, name
    def __call__(self, xs): return _rec("Lambda", {"function": self.fn.__name__, "name": self.name, "inputs": [list(t.shape) for t in xs]}, T(self.o))


def Model(inputs=None, outputs=None): return {"inputs": inputs, "outputs": outputs}


def main():
    REF = "/root/reference/utils.py"
    mod = ast.parse(open(REF).read())
    body = [n for n in mod.body if (isinstance(n, ast.FunctionDef) and n.name in {"STN", "get_initial_weights", "ctc_lambda_func"}) or
            (isinstance(n, ast.ClassDef) and n.name == "CRNN")]
    ns = dict(globals())
    ns["np"] = np
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    out = {"source": "reference utils.py:32-103, 239-258 executed over the recording Keras shim of tests/golden/make_graph_golden.py", "cases": []}
    for shape in ((100, 32, 1), (128, 32, 1)):
        for gru in (True, False):
            for V in (38, 97):
                TRACE.clear()
                c = ns["CRNN"](num_classes=V, max_string_len=23, shape=shape, time_dense_size=128, GRU=gru, n_units=256)
                c.get_model()
                out["cases"].append({"shape": list(shape), "GRU": gru, "num_classes": V, "pooling_counter_h": c.pooling_counter_h,
                                     "pooling_counter_w": c.pooling_counter_w, "trace": [dict(t) for t in TRACE]})
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "graph_golden.json"), "w"))
    t = out["cases"][0]["trace"]
    print(len(out["cases"]), "cases;", len(t), "layers;", [(x["layer"], x["shape"]) for x in t if x["layer"] in ("Reshape", "Flatten", "Lambda", "BilinearInterpolation")])


if __name__ == "__main__":
    main()
