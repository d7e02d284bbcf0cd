"""oracle/crnn_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (PyTorch-CPU tensors, fp32 by default, fp64 optional) of the reference's CRNN graph
(gasparian/CRNN-OCR-lite utils.py:32-258) with the layer hyper-parameters of models/*/model.json and the
Keras 2.2.2 / TF 1.8 semantics summarised in SURVEY.md Appendix A.  Backward = torch.autograd over this
restated forward (CTC gradient = the TF CTCLoss op gradient from oracle/ctc_oracle.c).

PARITY: Keras/TF cannot run in this image and the reference has no tests.  forward() + beam decode are PINNED
against the reference's own example predictions (7 figure-extracted inputs through the shipped weights give the labels
the reference printed, e.g. "cellist" -> "celist"; tests/test_golden.py, tests/golden/make_reference_examples.py).
bilinear_sampler() is PINNED bit-exactly on the reference's own BilinearInterpolation source (utils.py:116-232) executed over a numpy
shim of its backend ops (tests/golden/make_sampler_golden.py, tests/test_golden.py).
The training half (CTC gradient, backward, Adam) is PARITY UNPINNED: pinned by structure (parameter counts /
shapes of the shipped weight files, model_summary.txt), closed-form sampler invariants and torch cross-checks only.

Tensor layout follows the reference: NHWC, axis 1 ("H") = text-line width = time, axis 2 ("W") = 32.
Weight names are "<keras layer>/<weight>" exactly as stored in models/<name>/final_weights.h5.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import ctc_oracle

BLOCK_PLAN = [  # (Cin, Cout, pool) -- utils.py:64-70
    (1, 64, None), (64, 128, None), (128, 256, (2, 2)), (256, 256, None),
    (256, 512, (1, 2)), (512, 512, None), (512, 512, None)]
BN_EPS = 1e-3       # model.json: "epsilon": 0.001
BN_MOMENTUM = 0.99  # model.json: "momentum": 0.99
K_EPS = 1e-7        # keras.backend.epsilon()


@dataclass
class Cfg:
    imgh: int = 100          # axis-1 length (text-line width), train.py:105
    imgw: int = 32           # axis-2 length (line height), train.py:106
    num_classes: int = 38    # len(lexicon)+1, train.py:165
    cell: str = "gru"        # CLI/shipped weights are GRU; "lstm" = utils.py:78-79
    n_units: int = 256
    time_dense: int = 128
    max_len: int = 23

    @property
    def T(self):
        return (self.imgh + 4) // 2

    @property
    def feat(self):
        return ((self.imgw + 4) // 4) * 512

    @property
    def gates(self):
        return 3 if self.cell == "gru" else 4

    def loc_flat(self):
        h1, w1 = self.imgh // 2 - 4, self.imgw // 2 - 4
        return (h1 // 2 - 4) * (w1 // 2 - 4) * 20


def rnn_names(cfg: Cfg, layer: int):
    c = cfg.cell
    return [f"bidirectional_{layer}/forward_{c}_{layer}", f"bidirectional_{layer}/backward_{c}_{layer}"]


def weight_shapes(cfg: Cfg) -> "OrderedDict[str, tuple]":
    """Every weight of the graph in Keras `layer_names`/`weight_names` order."""
    s = OrderedDict()
    s["conv2d_1/kernel"] = (5, 5, 1, 20); s["conv2d_1/bias"] = (20,)
    s["conv2d_2/kernel"] = (5, 5, 20, 20); s["conv2d_2/bias"] = (20,)
    s["dense_1/kernel"] = (cfg.loc_flat(), 50); s["dense_1/bias"] = (50,)
    s["dense_2/kernel"] = (50, 6); s["dense_2/bias"] = (6,)
    for i, (cin, cout, _) in enumerate(BLOCK_PLAN, 1):
        s[f"depthwise_conv2d_{i}/depthwise_kernel"] = (3, 3, cin, 1)
        for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"batch_normalization_{2 * i - 1}/{nm}"] = (cin,)
        s[f"conv2d_{i + 2}/kernel"] = (1, 1, cin, cout)
        for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"batch_normalization_{2 * i}/{nm}"] = (cout,)
    s["dense1/kernel"] = (cfg.feat, cfg.time_dense); s["dense1/bias"] = (cfg.time_dense,)
    g, u = cfg.gates, cfg.n_units
    for layer, cin in ((1, cfg.time_dense), (2, u)):
        for nm in rnn_names(cfg, layer):
            s[nm + "/kernel"] = (cin, g * u); s[nm + "/recurrent_kernel"] = (u, g * u); s[nm + "/bias"] = (g * u,)
    s["dense2/kernel"] = (2 * u, cfg.num_classes); s["dense2/bias"] = (cfg.num_classes,)
    return s


def is_trainable(name):
    return not (name.endswith("moving_mean") or name.endswith("moving_variance"))


def init_weights(cfg: Cfg, seed=0) -> "OrderedDict[str, np.ndarray]":
    """Keras default initialisers: glorot_uniform convs/dense, he_normal RNN kernels + dense2,
    orthogonal recurrent kernels, zero biases (LSTM forget bias 1), BN (1,0,0,1), STN dense_2 = identity."""
    rng = np.random.default_rng(seed)
    w = OrderedDict()
    for name, shp in weight_shapes(cfg).items():
        leaf = name.split("/")[-1]
        if leaf in ("bias", "beta", "moving_mean"):
            a = np.zeros(shp, np.float32)
            if leaf == "bias" and "lstm" in name:
                a[cfg.n_units:2 * cfg.n_units] = 1.0
        elif leaf in ("gamma", "moving_variance"):
            a = np.ones(shp, np.float32)
        elif leaf == "recurrent_kernel":
            u = shp[0]
            blocks = []
            for _ in range(shp[1] // u):
                q, r = np.linalg.qr(rng.standard_normal((u, u)))
                blocks.append(q * np.sign(np.diag(r)))
            a = np.concatenate(blocks, 1).astype(np.float32)
        else:
            if leaf == "depthwise_kernel":
                fan_in, fan_out = 9 * shp[2], 9
            elif len(shp) == 4:
                rf = shp[0] * shp[1]; fan_in, fan_out = rf * shp[2], rf * shp[3]
            else:
                fan_in, fan_out = shp
            if name.startswith("bidirectional") or name.startswith("dense2"):
                std = math.sqrt(2.0 / fan_in) / 0.87962566103423978  # he_normal = truncated normal
                a = np.clip(rng.standard_normal(shp), -2, 2) * std
            else:
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                a = rng.uniform(-lim, lim, shp)
            a = a.astype(np.float32)
        w[name] = a
    # STN final layer: W=0, b=[1,0,0,0,1,0] (utils.py:239-245)
    w["dense_2/kernel"][:] = 0
    w["dense_2/bias"][:] = np.array([1, 0, 0, 0, 1, 0], np.float32)
    return w


def randomize_for_test(w, seed=0):
    """Perturb the 'boring' initial values (BN stats, biases, STN head) so parity tests exercise them."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for k, v in w.items():
        leaf = k.split("/")[-1]
        v = v.copy()
        if leaf in ("bias", "beta"):
            v += rng.standard_normal(v.shape).astype(np.float32) * 0.05
        elif leaf == "gamma":
            v *= (1 + 0.1 * rng.standard_normal(v.shape)).astype(np.float32)
        elif leaf == "moving_mean":
            v += rng.standard_normal(v.shape).astype(np.float32) * 0.1
        elif leaf == "moving_variance":
            v *= rng.uniform(0.5, 1.5, v.shape).astype(np.float32)
        out[k] = v
    out["dense_2/kernel"] = (rng.standard_normal((50, 6)) * 0.01).astype(np.float32)
    out["dense_2/bias"] = (np.array([1, 0, 0, 0, 1, 0]) + rng.standard_normal(6) * 0.03).astype(np.float32)
    return out


# ------------------------------------------------------------------------------------------
# ops
# ------------------------------------------------------------------------------------------
def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def hard_sigmoid(v):
    return torch.clamp(0.2 * v + 0.5, 0.0, 1.0)


def stn_locnet(w, x):
    """utils.py:247-256.  x (B,H,W,1) -> theta (B,6).  Convs are 'valid', bias, NO activation."""
    y = F.max_pool2d(_nchw(x), 2)
    y = F.conv2d(y, w["conv2d_1/kernel"].permute(3, 2, 0, 1), w["conv2d_1/bias"])
    y = F.max_pool2d(y, 2)
    y = F.conv2d(y, w["conv2d_2/kernel"].permute(3, 2, 0, 1), w["conv2d_2/bias"])
    y = _nhwc(y).reshape(x.shape[0], -1)            # Flatten in (H,W,C) order
    y = torch.relu(y @ w["dense_1/kernel"] + w["dense_1/bias"])
    return y @ w["dense_2/kernel"] + w["dense_2/bias"]


def linspace_tf(n, dtype):
    """tf.linspace(-1., 1., n) as TF 1.8 computes it in float32: start + step*i."""
    step = np.float32(2.0) / np.float32(n - 1)
    v = (np.float32(-1.0) + step * np.arange(n, dtype=np.float32)).astype(np.float32)
    return torch.tensor(v).to(dtype)


def bilinear_sampler(x, theta):
    """BilinearInterpolation._transform/_interpolate, utils.py:140-232, quirks kept verbatim:
    scale by size (not size-1), int cast truncates toward zero, corners clipped BEFORE the weights are
    formed, 4-term sum left-associated.  fp32 op order (shared with the CUDA kernel, no FMA contraction):
        xs = (t0*gx + t1*gy) + t2 ;  xf = (0.5*(xs+1))*W."""
    B, H, W, _ = x.shape
    dt = x.dtype
    gx = linspace_tf(W, dt).view(1, 1, W).expand(1, H, W)
    gy = linspace_tf(H, dt).view(1, H, 1).expand(1, H, W)
    t = theta.view(B, 6, 1, 1)
    xs = (t[:, 0] * gx + t[:, 1] * gy) + t[:, 2]
    ys = (t[:, 3] * gx + t[:, 4] * gy) + t[:, 5]
    xf = (0.5 * (xs + 1.0)) * float(W)
    yf = (0.5 * (ys + 1.0)) * float(H)
    big = 2.0 ** 30
    x0 = torch.nan_to_num(xf.detach(), nan=0.0).clamp(-big, big).to(torch.int32)  # trunc toward zero
    y0 = torch.nan_to_num(yf.detach(), nan=0.0).clamp(-big, big).to(torch.int32)
    x1, y1 = x0 + 1, y0 + 1
    x0 = x0.clamp(0, W - 1); x1 = x1.clamp(0, W - 1)
    y0 = y0.clamp(0, H - 1); y1 = y1.clamp(0, H - 1)
    img = x.reshape(B, H * W)
    def gat(yy, xx):
        return torch.gather(img, 1, (yy.long() * W + xx.long()).reshape(B, -1)).reshape(B, H, W)
    pa, pb, pc, pd = gat(y0, x0), gat(y1, x0), gat(y0, x1), gat(y1, x1)
    x0f, x1f, y0f, y1f = x0.to(dt), x1.to(dt), y0.to(dt), y1.to(dt)
    wa = (x1f - xf) * (y1f - yf)
    wb = (x1f - xf) * (yf - y0f)
    wc = (xf - x0f) * (y1f - yf)
    wd = (xf - x0f) * (yf - y0f)
    out = ((wa * pa + wb * pb) + wc * pc) + wd * pd
    return out.unsqueeze(-1)


def batchnorm(w, idx, y, training, new_stats):
    """BatchNormalization(axis=-1, eps 1e-3, momentum .99) on NHWC."""
    g, b = w[f"batch_normalization_{idx}/gamma"], w[f"batch_normalization_{idx}/beta"]
    if training:
        mean = y.mean((0, 1, 2))
        var = y.var((0, 1, 2), unbiased=False)
        n = float(y.shape[0] * y.shape[1] * y.shape[2])
        if new_stats is not None:
            mm, mv = w[f"batch_normalization_{idx}/moving_mean"].detach(), w[f"batch_normalization_{idx}/moving_variance"].detach()
            var_mov = var.detach() * (n / (n - 1.0)) * (n / (n - (1.0 + BN_EPS)))
            new_stats[f"batch_normalization_{idx}/moving_mean"] = mm - (mm - mean.detach()) * (1.0 - BN_MOMENTUM)
            new_stats[f"batch_normalization_{idx}/moving_variance"] = mv - (mv - var_mov) * (1.0 - BN_MOMENTUM)
    else:
        mean, var = w[f"batch_normalization_{idx}/moving_mean"], w[f"batch_normalization_{idx}/moving_variance"]
    return (y - mean) * (g / torch.sqrt(var + BN_EPS)) + b


def relu6(y):
    return torch.clamp(y, 0.0, 6.0)


def _forced_gate(z, act, gate):
    """Teacher-forced piecewise-linear activation for the isolated backward tests: same forward value as act(z) (up to the rounding-level
    difference at the elements whose decision differs), but the derivative is 1 exactly where `gate` (the decision the device path took in
    ITS forward, read back by the test) is set -- so a fp32-vs-fp64 rounding flip of a ReLU gate cannot masquerade as a kernel error."""
    return torch.where(gate, z, act(z).detach())


def conv_block(w, i, x, pool, training, masks, new_stats, keep, gate1=None):
    """CRNN.depthwise_conv_block, utils.py:43-56 (i = 1..7).  gate1 (tests only): forced ReLU6 pass-through mask of the FIRST activation."""
    c = x.shape[-1]
    y1 = _nhwc(F.conv2d(_nchw(x), w[f"depthwise_conv2d_{i}/depthwise_kernel"].permute(2, 3, 0, 1), padding=1, groups=c))
    keep[f"dw{i}"] = y1
    z1 = batchnorm(w, 2 * i - 1, y1, training, new_stats)
    a1 = relu6(z1) if gate1 is None else _forced_gate(z1, relu6, gate1)
    y2 = a1 @ w[f"conv2d_{i + 2}/kernel"][0, 0]
    keep[f"pw{i}"] = y2
    a2 = relu6(batchnorm(w, 2 * i, y2, training, new_stats))
    if pool is not None:
        a2 = _nhwc(F.max_pool2d(_nchw(a2), pool))
    if training and masks is not None and f"dropout_{i}" in masks:
        a2 = a2 * masks[f"dropout_{i}"]
    keep[f"block{i}"] = a2
    return a2


def gru_dir(x, Wk, U, b, reverse):
    """Keras 2.2.2 GRUCell reset_after=False.  x (B,T,in) -> (B,T,u) in time order."""
    B, T, _ = x.shape
    u = U.shape[0]
    xp = x @ Wk + b
    h = x.new_zeros(B, u)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        xz, xr, xh = xp[:, t, :u], xp[:, t, u:2 * u], xp[:, t, 2 * u:]
        z = hard_sigmoid(xz + h @ U[:, :u])
        r = hard_sigmoid(xr + h @ U[:, u:2 * u])
        hh = torch.tanh(xh + (r * h) @ U[:, 2 * u:])
        h = z * h + (1 - z) * hh
        outs[t] = h
    return torch.stack(outs, 1)


def lstm_dir(x, Wk, U, b, reverse):
    """Keras 2.2.2 LSTMCell, gate order [i,f,c,o], hard_sigmoid recurrent activation."""
    B, T, _ = x.shape
    u = U.shape[0]
    xp = x @ Wk + b
    h = x.new_zeros(B, u); c = x.new_zeros(B, u)
    outs = [None] * T
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        g = xp[:, t] + h @ U
        i = hard_sigmoid(g[:, :u]); f = hard_sigmoid(g[:, u:2 * u])
        c = f * c + i * torch.tanh(g[:, 2 * u:3 * u])
        o = hard_sigmoid(g[:, 3 * u:])
        h = o * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, 1)


def birnn(w, cfg, layer, x, keep):
    fn = gru_dir if cfg.cell == "gru" else lstm_dir
    nf, nb = rnn_names(cfg, layer)
    yf = fn(x, w[nf + "/kernel"], w[nf + "/recurrent_kernel"], w[nf + "/bias"], False)
    yb = fn(x, w[nb + "/kernel"], w[nb + "/recurrent_kernel"], w[nb + "/bias"], True)
    keep[f"rnn{layer}_f"], keep[f"rnn{layer}_b"] = yf, yb
    return yf + yb if layer == 1 else torch.cat([yf, yb], -1)   # merge 'sum' then 'concat', utils.py:78-82


def forward(w, x, cfg: Cfg, training=False, masks=None, new_stats=None):
    """the_input (B,imgh,imgw,1) -> dict of intermediates; 'softmax' (B,T,V) is the predictor output
    (utils.py:308-312).  `masks`: optional dict dropout_k -> pre-scaled keep mask (training only)."""
    keep = OrderedDict()
    theta = stn_locnet(w, x)
    keep["theta"] = theta
    s = bilinear_sampler(x, theta)
    keep["stn"] = s
    h = F.pad(s, (0, 0, 2, 2, 2, 2))              # ZeroPadding2D((2,2)), utils.py:63
    for i, (_, _, pool) in enumerate(BLOCK_PLAN, 1):
        h = conv_block(w, i, h, pool, training, masks, new_stats, keep)
    return head(w, h, cfg, training, masks, keep)


def head(w, h, cfg: Cfg, training=False, masks=None, keep=None, dense1_gate=None):
    """Everything after the conv stack (utils.py:72-86): block-7 output (B,T,9,512) -> reshape -> dense1 -> 2 x Bi-RNN -> dense2 -> softmax.
    dense1_gate (tests only): forced ReLU pass-through mask of dense1, see _forced_gate."""
    keep = OrderedDict() if keep is None else keep
    B, T = h.shape[0], h.shape[1]
    h = h.reshape(B, T, -1)                        # feature index = w*512 + c, utils.py:72-73
    h = h @ w["dense1/kernel"] + w["dense1/bias"]
    h = torch.relu(h) if dense1_gate is None else _forced_gate(h, torch.relu, dense1_gate)
    if training and masks is not None and "dropout_8" in masks:
        h = h * masks["dropout_8"]
    keep["dense1"] = h
    h = birnn(w, cfg, 1, h, keep)
    keep["rnn1"] = h
    h = birnn(w, cfg, 2, h, keep)
    if training and masks is not None and "dropout_9" in masks:
        h = h * masks["dropout_9"]
    keep["rnn2"] = h
    z = h @ w["dense2/kernel"] + w["dense2/bias"]
    keep["logits"] = z
    keep["softmax"] = torch.softmax(z, -1)
    return keep


class _CTCLossTF(torch.autograd.Function):
    """K.ctc_batch_cost on y_pred[:,2:,:] (utils.py:98-103): loss and d loss/d p via the TF CTCLoss gradient."""

    @staticmethod
    def forward(ctx, probs, labels, label_len, input_len):
        p = probs.detach().to(torch.float32).numpy()
        loss, grad_u = ctc_oracle.ctc_loss_grad(p, labels, label_len, input_len, eps=K_EPS)
        ctx.save_for_backward(probs, torch.tensor(grad_u))
        return torch.tensor(loss).to(probs.dtype)

    @staticmethod
    def backward(ctx, gout):
        probs, grad_u = ctx.saved_tensors
        return gout.view(-1, 1, 1) * (grad_u.to(probs.dtype) / (probs + K_EPS)), None, None, None


def ctc_batch_cost(softmax, labels, label_len, input_len, exact64=False):
    """softmax (B,T,V) full-T network output; returns per-sample loss (B,)."""
    y = softmax[:, 2:, :]
    if exact64:
        u = torch.log(y + K_EPS)
        lsm = torch.log_softmax(u, -1).transpose(0, 1)
        return F.ctc_loss(lsm, torch.as_tensor(np.asarray(labels), dtype=torch.long), torch.as_tensor(np.asarray(input_len).reshape(-1), dtype=torch.long),
                          torch.as_tensor(np.asarray(label_len).reshape(-1), dtype=torch.long), blank=softmax.shape[-1] - 1, reduction="none")
    return _CTCLossTF.apply(y, np.asarray(labels), np.asarray(label_len), np.asarray(input_len))


def to_torch(w, dtype=torch.float32, grad=False):
    out = OrderedDict()
    for k, v in w.items():
        t = torch.tensor(np.asarray(v)).to(dtype)
        if grad and is_trainable(k):
            t.requires_grad_(True)
        out[k] = t
    return out


def loss_and_grads(w_np, x, labels, label_len, input_len, cfg, masks=None, dtype=torch.float32):
    """One training forward/backward: mean-over-batch CTC loss (identity Keras loss, train.py:192) and
    d loss / d every trainable weight; also the BN moving-stat updates and all intermediates."""
    w = to_torch(w_np, dtype, grad=True)
    new_stats = OrderedDict()
    xt = torch.tensor(np.asarray(x)).to(dtype)
    keep = forward(w, xt, cfg, training=True, masks=masks, new_stats=new_stats)
    per = ctc_batch_cost(keep["softmax"], labels, label_len, input_len, exact64=(dtype == torch.float64))
    loss = per.mean()
    names = [k for k in w if is_trainable(k)]
    grads = torch.autograd.grad(loss, [w[k] for k in names], allow_unused=True)
    g = OrderedDict((k, (gi if gi is not None else torch.zeros_like(w[k])).detach().to(torch.float32).numpy()) for k, gi in zip(names, grads))
    return float(loss.detach()), per.detach().numpy(), g, OrderedDict((k, v.to(torch.float32).numpy()) for k, v in new_stats.items()), keep


def clip_by_global_norm(grads, clipnorm=5.0):
    """Keras optimizers.clip_norm with the GLOBAL norm over all gradients."""
    norm = math.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values()))
    if norm >= clipnorm:
        return OrderedDict((k, (g * np.float32(clipnorm) / np.float32(norm)).astype(np.float32)) for k, g in grads.items()), norm
    return grads, norm


def adam_step(w, grads, state, lr=1e-4, b1=0.5, b2=0.999, eps=1e-7, clipnorm=5.0):
    """Keras 2.2.2 Adam (train.py:188; values in models/*/final_model.h5 training_config)."""
    grads, norm = clip_by_global_norm(grads, clipnorm)
    t = state.get("iterations", 0) + 1
    lr_t = np.float32(lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))
    out = OrderedDict(w)
    for k, g in grads.items():
        m = state.get("m/" + k, np.zeros_like(g)); v = state.get("v/" + k, np.zeros_like(g))
        m = np.float32(b1) * m + np.float32(1 - b1) * g
        v = np.float32(b2) * v + np.float32(1 - b2) * g * g
        out[k] = (w[k] - lr_t * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32)
        state["m/" + k], state["v/" + k] = m.astype(np.float32), v.astype(np.float32)
    state["iterations"] = t
    return out, norm


def sgd_step(w, grads, state, lr=1e-3, decay=1e-6, momentum=0.9, clipnorm=5.0):
    """Keras 2.2.2 SGD(nesterov=True) (train.py:190)."""
    grads, norm = clip_by_global_norm(grads, clipnorm)
    it = state.get("iterations", 0)
    lr_i = np.float32(lr * (1.0 / (1.0 + decay * it)))
    out = OrderedDict(w)
    for k, g in grads.items():
        vel = state.get("vel/" + k, np.zeros_like(g))
        vel = np.float32(momentum) * vel - lr_i * g
        out[k] = (w[k] + np.float32(momentum) * vel - lr_i * g).astype(np.float32)
        state["vel/" + k] = vel.astype(np.float32)
    state["iterations"] = it + 1
    return out, norm


def synth_batch(cfg: Cfg, B, seed):
    """Synthetic inputs of SURVEY 8d: normalised U{0..255} pixels, labels L~U{3..max_len}."""
    rng = np.random.default_rng(seed)
    x = ((rng.integers(0, 256, (B, cfg.imgh, cfg.imgw, 1)).astype(np.float32) - np.float32(118.24236953981779))
         / np.float32(36.72835353999682)).astype(np.float32)
    L = rng.integers(3, cfg.max_len + 1, B).astype(np.int32)
    labels = np.full((B, cfg.max_len), cfg.num_classes - 1, np.int32)
    for b in range(B):
        labels[b, :L[b]] = rng.integers(0, cfg.num_classes - 1, L[b])
    input_len = np.full(B, cfg.T - 2, np.int32)
    return x, labels, L, input_len
