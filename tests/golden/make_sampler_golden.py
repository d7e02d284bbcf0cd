#!/usr/bin/env python3
"""Golden vectors for the STN sampler: executes the REFERENCE's own BilinearInterpolation layer source
(utils.py:116-232: _make_regular_grids, _transform, _interpolate -- extracted by AST from /root/reference/utils.py) over a small NUMPY
shim of the keras-backend / tensorflow ops it calls (keras and tensorflow are not installable here), and writes
tests/golden/sampler_golden.npz.  What the shim assumes about those ops is stated next to each one; all arithmetic is float32.
Run in the build container only; the JSON/NPZ is committed."""
import ast, os, types
import numpy as np

f32 = np.float32


def _cast(x, dtype):
    x = np.asarray(x)
    if dtype == "int32":
        return np.trunc(x).astype(np.int32)          # tf.cast float->int32 truncates toward zero
    return x.astype(dtype)


def _batch_dot(a, b):
    # K.batch_dot((B,2,3), (B,3,N)) = per-sample matmul; K = 3 terms accumulated left to right in float32, no FMA contraction
    a = a.astype(f32); b = b.astype(f32)
    out = (a[:, :, 0:1] * b[:, 0:1, :]).astype(f32)
    out = (out + (a[:, :, 1:2] * b[:, 1:2, :]).astype(f32)).astype(f32)
    out = (out + (a[:, :, 2:3] * b[:, 2:3, :]).astype(f32)).astype(f32)
    return out


def _linspace(start, stop, num):
    # tf.linspace (LinSpace op): start + step * i with step = (stop - start) / (num - 1), float32
    step = f32((f32(stop) - f32(start)) / f32(num - 1))
    return (f32(start) + step * np.arange(num, dtype=f32)).astype(f32)


K = types.SimpleNamespace(
    shape=lambda x: np.asarray(x).shape, int_shape=lambda x: np.asarray(x).shape, cast=_cast,
    flatten=lambda x: np.asarray(x).reshape(-1), clip=lambda x, lo, hi: np.clip(x, lo, hi),
    arange=lambda a, b: np.arange(a, b, dtype=np.int32), expand_dims=lambda x, axis: np.expand_dims(x, axis),
    repeat_elements=lambda x, rep, axis: np.repeat(x, rep, axis=axis), reshape=lambda x, shape: np.reshape(x, shape),
    gather=lambda ref, idx: np.asarray(ref)[np.asarray(idx)], ones_like=lambda x: np.ones_like(x),
    concatenate=lambda xs, axis: np.concatenate(xs, axis), tile=lambda x, n: np.tile(x, n), stack=lambda xs: np.array(xs),
    batch_dot=_batch_dot)
tf = types.SimpleNamespace(meshgrid=lambda x, y: np.meshgrid(x, y), linspace=_linspace)


class Layer:
    def __init__(self, **kw): pass


REF = "/root/reference/utils.py"
mod = ast.parse(open(REF).read())
body = [n for n in mod.body if (isinstance(n, ast.FunctionDef) and n.name in {"K_meshgrid", "K_linspace"}) or
        (isinstance(n, ast.ClassDef) and n.name == "BilinearInterpolation")]
ns = {"K": K, "tf": tf, "Layer": Layer, "np": np}
exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)

if __name__ == "__main__":
    rng = np.random.default_rng(77)
    thetas = [[1, 0, 0, 0, 1, 0], [1, 0, 0.25, 0, 1, -0.1], [0.8, 0.15, -0.05, -0.1, 1.1, 0.2], [1.6, 0.4, 0.6, 0.3, 1.5, -0.7], [0.5, 0, -0.9, 0, 0.5, 0.9]]
    out = {"n": np.int64(0)}
    k = 0
    for H in (100, 128):
        layer = ns["BilinearInterpolation"](output_size=(H, 32))
        x = ((rng.integers(0, 256, (len(thetas), H, 32, 1)).astype(f32) - f32(118.24236953981779)) / f32(36.72835353999682)).astype(f32)
        th = np.array(thetas, f32) + (rng.standard_normal((len(thetas), 6)) * 0.01).astype(f32) * np.array([0] + [1] * (len(thetas) - 1), f32)[:, None]
        y = layer.call([x, th.astype(f32)])
        assert y.dtype == np.float32 and y.shape == (len(thetas), H, 32, 1)
        out["x_%d" % k] = x; out["theta_%d" % k] = th.astype(f32); out["y_%d" % k] = y
        k += 1
    out["n"] = np.int64(k)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_golden.npz"), **out)
    print(k, "cases; identity invariants:", float(np.abs(out["y_0"][0, -1]).max()), float(np.abs(out["y_0"][0, :, -1]).max()), bool(out["y_0"][0, 0, 0, 0] == out["x_0"][0, 0, 0, 0]))
