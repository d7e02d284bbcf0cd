"""Host-side mirror of the reference's model surface (gasparian/CRNN-OCR-lite utils.py:32-96, 300-329) over the
B200 C ABI (include/crnn_b200.h).  PyTorch is used only as the device allocator / stream provider / NCCL plumbing.

    CRNN(num_classes, max_string_len, shape, time_dense_size, GRU, n_units).get_model() -> CRNNModel
    CRNNModel: load_weights / save_weights / get_weights / set_weights / to_json / summary / compile /
               predict_on_batch / predict_generator / train_on_batch / fit_generator / save
"""
from __future__ import annotations

import ctypes
import json
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from . import hdf5_lite

BLOCK_PLAN = [(1, 64, None), (64, 128, None), (128, 256, (2, 2)), (256, 256, None), (256, 512, (1, 2)), (512, 512, None), (512, 512, None)]


def _cell_name(gru):
    return "gru" if gru else "lstm"


def weight_shapes(imgh, imgw, num_classes, cell, n_units=256, time_dense=128):
    """Keras `layer_names`/`weight_names` order and shapes of models/<name>/final_weights.h5."""
    s = OrderedDict()
    p1h, p1w = imgh // 2, imgw // 2
    c2h, c2w = (p1h - 4) // 2 - 4, (p1w - 4) // 2 - 4
    s["conv2d_1/kernel"] = (5, 5, 1, 20); s["conv2d_1/bias"] = (20,)
    s["conv2d_2/kernel"] = (5, 5, 20, 20); s["conv2d_2/bias"] = (20,)
    s["dense_1/kernel"] = (c2h * c2w * 20, 50); s["dense_1/bias"] = (50,)
    s["dense_2/kernel"] = (50, 6); s["dense_2/bias"] = (6,)
    for i, (cin, cout, _) in enumerate(BLOCK_PLAN, 1):
        s[f"depthwise_conv2d_{i}/depthwise_kernel"] = (3, 3, cin, 1)
        for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"batch_normalization_{2 * i - 1}/{nm}"] = (cin,)
        s[f"conv2d_{i + 2}/kernel"] = (1, 1, cin, cout)
        for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[f"batch_normalization_{2 * i}/{nm}"] = (cout,)
    feat = ((imgw + 4) // 4) * 512
    s["dense1/kernel"] = (feat, time_dense); s["dense1/bias"] = (time_dense,)
    g = 3 if cell == "gru" else 4
    for layer, cin in ((1, time_dense), (2, n_units)):
        for d in ("forward", "backward"):
            base = f"bidirectional_{layer}/{d}_{cell}_{layer}"
            s[base + "/kernel"] = (cin, g * n_units); s[base + "/recurrent_kernel"] = (n_units, g * n_units); s[base + "/bias"] = (g * n_units,)
    s["dense2/kernel"] = (2 * n_units, num_classes); s["dense2/bias"] = (num_classes,)
    return s


def keras_initial_weights(shapes, cell, n_units=256, seed=None):
    """Keras default initialisers for a freshly built model (train.py:168 before load_weights)."""
    rng = np.random.default_rng(seed)
    w = OrderedDict()
    for name, shp in shapes.items():
        leaf = name.split("/")[-1]
        if leaf in ("bias", "beta", "moving_mean"):
            a = np.zeros(shp, np.float32)
            if leaf == "bias" and "lstm" in name:
                a[n_units:2 * n_units] = 1.0  # unit_forget_bias
        elif leaf in ("gamma", "moving_variance"):
            a = np.ones(shp, np.float32)
        elif leaf == "recurrent_kernel":
            blocks = []
            for _ in range(shp[1] // shp[0]):
                q, r = np.linalg.qr(rng.standard_normal((shp[0], shp[0])))
                blocks.append(q * np.sign(np.diag(r)))
            a = np.concatenate(blocks, 1).astype(np.float32)
        else:
            if leaf == "depthwise_kernel":
                fan_in, fan_out = 9 * shp[2], 9
            elif len(shp) == 4:
                fan_in, fan_out = shp[0] * shp[1] * shp[2], shp[0] * shp[1] * shp[3]
            else:
                fan_in, fan_out = shp
            if name.startswith("bidirectional") or name.startswith("dense2"):   # he_normal (utils.py:78-85)
                # Keras 2.2.2 VarianceScaling(scale=2, fan_in, 'normal') = K.truncated_normal(stddev=sqrt(2/fan_in)): TF re-draws every
                # sample beyond 2 stddev (no clipping), and 2.2.2 predates the 1/0.8796 stddev correction of later Keras releases
                a = rng.standard_normal(shp)
                out = np.abs(a) > 2
                while out.any():
                    a[out] = rng.standard_normal(int(out.sum()))
                    out = np.abs(a) > 2
                a = a * math.sqrt(2.0 / fan_in)
            else:                                                                 # glorot_uniform
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                a = rng.uniform(-lim, lim, shp)
            a = a.astype(np.float32)
        w[name] = a
    w["dense_2/kernel"][:] = 0                                                    # get_initial_weights, utils.py:239-245
    w["dense_2/bias"][:] = np.array([1, 0, 0, 0, 1, 0], np.float32)
    return w


class Adam:
    """keras.optimizers.Adam subset used by train.py:188."""
    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, clipnorm=0.0, **_):
        self.kind, self.lr, self.beta_1, self.beta_2, self.epsilon, self.clipnorm = "adam", lr, beta_1, beta_2, epsilon, clipnorm


class SGD:
    """keras.optimizers.SGD subset used by train.py:190."""
    def __init__(self, lr=0.01, decay=0.0, momentum=0.0, nesterov=False, clipnorm=0.0, **_):
        if not nesterov:
            raise NotImplementedError("only the reference's nesterov=True SGD is implemented")
        self.kind, self.lr, self.decay, self.momentum, self.clipnorm = "sgd", lr, decay, momentum, clipnorm


class _History:
    def __init__(self):
        self.history = {"loss": []}


class CRNNModel:
    def __init__(self, num_classes, max_string_len, shape, time_dense_size, gru, n_units, max_batch=64, device=None, seed=None):
        if not torch.cuda.is_available():
            raise _lib.CrnnError("CRNNModel needs a CUDA device: this path has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.imgh, self.imgw = int(shape[0]), int(shape[1])
        self.num_classes, self.max_len = int(num_classes), int(max_string_len)
        self.cell = _cell_name(gru)
        self.n_units, self.time_dense = int(n_units), int(time_dense_size)
        self.max_batch = int(max_batch)
        self.T = (self.imgh + 4) // 2
        self.cfg = _lib.CrnnConfig(self.imgh, self.imgw, self.num_classes, _lib.CRNN_CELL_GRU if gru else _lib.CRNN_CELL_LSTM,
                                   self.n_units, self.time_dense, self.max_len, self.max_batch)
        nbytes = ctypes.c_size_t()
        _lib.check(self.lib.crnn_workspace_bytes(ctypes.byref(self.cfg), ctypes.byref(nbytes)))
        with torch.cuda.device(self.device):
            self.workspace = torch.zeros(nbytes.value, dtype=torch.uint8, device=self.device)
        self.handle = ctypes.c_void_p()
        _lib.check(self.lib.crnn_create(ctypes.byref(self.cfg), self.workspace.data_ptr(), nbytes.value, ctypes.byref(self.handle)))
        self.shapes = weight_shapes(self.imgh, self.imgw, self.num_classes, self.cell, self.n_units, self.time_dense)
        self.optimizer = None
        self.stop_training = False
        self.dropout = True
        self._step_seed = np.random.SeedSequence(seed).generate_state(1, np.uint64)[0] | 1
        self._pinned = {}
        self.set_weights(keras_initial_weights(self.shapes, self.cell, self.n_units, seed))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.crnn_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ tensors
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def tensor(self, name) -> torch.Tensor:
        """Torch view (no copy) of a named tensor of the workspace: weights, 'grad/<w>', 'act/<x>', 'arena/<a>'."""
        cache = self.__dict__.setdefault("_tensor_cache", {})
        t = cache.get(name)
        if t is None:
            info = _lib.TensorInfo()
            _lib.check(self.lib.crnn_tensor_lookup(self.handle, name.encode(), ctypes.byref(info)))
            raw = self.workspace[info.offset: info.offset + info.numel * 4]
            t = cache[name] = raw.view(torch.int32 if info.is_int else torch.float32)
        return t

    def tensor_names(self):
        return [self.lib.crnn_tensor_name(self.handle, i).decode() for i in range(self.lib.crnn_num_tensors(self.handle))]

    def set_weights(self, weights):
        for name, shp in self.shapes.items():
            if name not in weights:
                raise KeyError(f"missing weight {name}")
            a = np.ascontiguousarray(weights[name], np.float32)
            if tuple(a.shape) != tuple(shp):
                raise ValueError(f"{name}: shape {a.shape} != {shp}")
            self.tensor(name).copy_(torch.from_numpy(a.reshape(-1)))

    def get_weights(self):
        return OrderedDict((n, self.tensor(n).cpu().numpy().reshape(s).copy()) for n, s in self.shapes.items())

    def get_grads(self):
        return OrderedDict((n, self.tensor("grad/" + n).cpu().numpy().reshape(s).copy()) for n, s in self.shapes.items()
                           if not n.endswith(("moving_mean", "moving_variance")))

    def activation(self, name, B=None):
        """Host copy of an activation tensor of the last forward (tests / inspection).  The outputs of the non-pooled blocks are not written
        by the forward pass (csrc/dwconv_fused.cu recomputes them where they are consumed): crnn_debug_materialize_blocks fills them in."""
        if name.startswith("block"):
            _lib.check(self.lib.crnn_debug_materialize_blocks(self.handle, self._stream()))
        return self.tensor("act/" + name).cpu().numpy()

    # ------------------------------------------------------------------ Keras-compatible weight I/O
    def load_weights(self, path):
        """model.load_weights(path) of the reference (utils.py:305,328; train.py:171): Keras-2.2.2 HDF5 layout."""
        self.set_weights(hdf5_lite.load_keras_weights(path))

    def _keras_layers(self):
        layers = OrderedDict()
        w = self.get_weights()
        for name, arr in w.items():
            layer, leaf = name.split("/", 1)
            layers.setdefault(layer, OrderedDict())[f"{layer}/{leaf}:0"] = arr
        return layers

    def save_weights(self, path):
        """model.save_weights (train.py:215), readable by Keras 2.2.2 / h5py and by load_weights above."""
        hdf5_lite.save_keras_weights(path, self._keras_layers())

    def save(self, path):
        """model.save (train.py:216): weights + optimizer slots + training_config in one file, laid out like the reference's
        models/*/final_model.h5 (/model_weights, /optimizer_weights with Keras' slot names and order; hdf5_lite.save_keras_model), root
        attribute `model_config` = the Keras-2.2.2 model JSON (to_json).  Errors propagate: there is no second layout to fall back to."""
        extra = {"model_config": self.to_json(), "training_config": json.dumps(self._training_config())}
        tr = [(n, s) for n, s in self.shapes.items() if not n.endswith(("moving_mean", "moving_variance"))]
        adam = sgd = None
        if self.optimizer is not None and self.optimizer.kind == "adam":
            adam = (int(self.iterations()), [self.tensor("adam_m/" + n).cpu().numpy().reshape(s) for n, s in tr],
                    [self.tensor("adam_v/" + n).cpu().numpy().reshape(s) for n, s in tr])
        elif self.optimizer is not None and self.optimizer.kind == "sgd":      # the velocities live in the first optimiser arena
            sgd = (int(self.iterations()), [self.tensor("adam_m/" + n).cpu().numpy().reshape(s) for n, s in tr])
        hdf5_lite.save_keras_model(path, self._keras_layers(), adam=adam, sgd=sgd, root_attrs=extra)

    def load_optimizer_state(self, path):
        """Resume Adam from a Keras-2.2.2 `final_model.h5` (the reference's own files or ours): iterations and the m / v slots of every
        trainable weight (hdf5_lite.load_keras_adam_state).  The reference itself only reloads weights (train.py:170-171)."""
        it, m, v = hdf5_lite.load_keras_adam_state(path)
        for n, a in m.items():
            self.tensor("adam_m/" + n).copy_(torch.from_numpy(np.ascontiguousarray(a, np.float32).reshape(-1)))
            self.tensor("adam_v/" + n).copy_(torch.from_numpy(np.ascontiguousarray(v[n], np.float32).reshape(-1)))
        _lib.check(self.lib.crnn_set_iterations(self.handle, int(it)))
        return it

    def _training_config(self):
        """`training_config` attribute as Keras 2.2.2 `model.save` writes it (models/*/final_model.h5)."""
        o = self.optimizer
        if o is None:
            return {}
        f32 = lambda v: float(np.float32(v))            # Keras stores the float32 value of its backend variables
        if o.kind == "adam":
            oc = OrderedDict([("clipnorm", o.clipnorm), ("lr", f32(o.lr)), ("beta_1", f32(o.beta_1)), ("beta_2", f32(o.beta_2)), ("decay", 0.0),
                              ("epsilon", o.epsilon), ("amsgrad", False)])
        else:
            oc = OrderedDict([("clipnorm", o.clipnorm), ("lr", f32(o.lr)), ("momentum", f32(o.momentum)), ("decay", f32(o.decay)), ("nesterov", True)])
        return OrderedDict([("optimizer_config", OrderedDict([("class_name", "Adam" if o.kind == "adam" else "SGD"), ("config", oc)])),
                            ("loss", {"ctc": "<lambda>"}), ("metrics", []), ("sample_weight_mode", None), ("loss_weights", None)])

    def to_json(self):
        """Keras-2.2.2 functional-model JSON of this graph: byte-identical to the reference's models/*/model.json for the same
        hyper-parameters (keras_json.py), so directories written by either side load on the other."""
        from . import keras_json
        return json.dumps(keras_json.keras_model_config(self.imgh, self.imgw, self.num_classes, self.max_len, self.time_dense, self.n_units, self.cell))

    def summary(self, print_fn=print):
        tot = sum(int(np.prod(s)) for s in self.shapes.values())
        nt = sum(int(np.prod(s)) for n, s in self.shapes.items() if n.endswith(("moving_mean", "moving_variance")))
        print_fn("_" * 65)
        for n, s in self.shapes.items():
            print_fn(f"{n:<58}{str(tuple(s)):>20}")
        print_fn("=" * 65)
        print_fn(f"Total params: {tot:,}\nTrainable params: {tot - nt:,}\nNon-trainable params: {nt:,}")

    # ------------------------------------------------------------------ inference (predict.py:166)
    def forward_device(self, x_dev: torch.Tensor) -> torch.Tensor:
        """x_dev (B,imgh,imgw,1) float32 CUDA tensor -> view of the softmax (B,T,V) inside the workspace."""
        B = int(x_dev.shape[0])
        assert x_dev.is_cuda and x_dev.dtype == torch.float32 and x_dev.is_contiguous()
        _lib.check(self.lib.crnn_forward(self.handle, x_dev.data_ptr(), B, None, self._stream()))
        return self.tensor("act/softmax")[: B * self.T * self.num_classes].view(B, self.T, self.num_classes)

    def predict_on_batch(self, x) -> np.ndarray:
        """Host numpy (B,imgh,imgw,1) -> host numpy softmax (B,T,V): H2D + forward + D2H through the C ABI."""
        if np.asarray(x).dtype == np.uint8:                      # raw 8-bit line images: upload bytes, normalise + forward on the device
            x = np.asarray(x)
            out = []
            for s in range(0, x.shape[0], self.max_batch):
                sm = self.forward_device(self._stage_input(x[s:s + self.max_batch]))
                out.append(sm.cpu().numpy().copy())
            return np.concatenate(out, 0)
        x = np.ascontiguousarray(x, np.float32)
        out = []
        for s in range(0, x.shape[0], self.max_batch):
            xb = x[s:s + self.max_batch]
            o = np.empty((xb.shape[0], self.T, self.num_classes), np.float32)
            _lib.check(self.lib.crnn_forward_host(self.handle, xb.ctypes.data, xb.shape[0], o.ctypes.data, self._stream()))
            out.append(o)
        return np.concatenate(out, 0)

    predict = predict_on_batch

    def predict_generator(self, generator, steps, **_):
        outs = []
        for _i in range(steps):
            inputs, _t = next(generator)
            outs.append(self.predict_on_batch(inputs["the_input"] if isinstance(inputs, dict) else inputs))
        return np.concatenate(outs, 0)

    # ------------------------------------------------------------------ training (train.py:187-209)
    def compile(self, loss=None, optimizer=None, **_):
        self.optimizer = optimizer

    def iterations(self):
        it = ctypes.c_int64()
        _lib.check(self.lib.crnn_get_iterations(self.handle, ctypes.byref(it)))
        return it.value

    def train_fwd_bwd_device(self, x_dev, labels_dev, label_len_dev, input_len_dev, dropout_seed=0):
        """Forward (training mode) + CTC + backward on device tensors; returns a view of the per-sample losses."""
        B = int(x_dev.shape[0])
        loss = self.tensor("act/loss")[:B]
        _lib.check(self.lib.crnn_train_fwd_bwd(self.handle, x_dev.data_ptr(), labels_dev.data_ptr(), label_len_dev.data_ptr(),
                                               input_len_dev.data_ptr(), B, loss.data_ptr(), ctypes.c_uint64(int(dropout_seed)), self._stream()))
        return loss

    def optimizer_step(self, grad_scale=1.0):
        o = self.optimizer
        if o is None:
            raise RuntimeError("compile(optimizer=...) first")
        if o.kind == "adam":
            _lib.check(self.lib.crnn_adam_step(self.handle, o.lr, o.beta_1, o.beta_2, o.epsilon, o.clipnorm or 0.0, grad_scale, self._stream()))
        else:
            _lib.check(self.lib.crnn_sgd_step(self.handle, o.lr, o.decay, o.momentum, o.clipnorm or 0.0, grad_scale, self._stream()))

    def enable_native_dp(self, fused=True):
        """Data parallel through the engine's own NCCL communicator (parallel.native_comm_for): with `fused` the training step reduces its
        gradients itself -- head bucket overlapped with the conv-stack backward -- and allreduce_grads() only returns the 1/world scale."""
        from . import parallel
        ok = parallel.native_comm_for(self, fused)
        self._native_dp = ("fused" if fused else "call") if ok else None
        return self._native_dp

    def allreduce_grads(self):
        """Data-parallel exchange (NEW capability): one sum all-reduce of the flat gradient arena.  Returns the scale (1/world)
        the optimiser must apply.  Fused native mode: already done inside the step; native call mode: crnn_allreduce_grads (C ABI, NCCL);
        otherwise torch.distributed (gloo in the CPU / shared-GPU tests)."""
        from . import parallel
        w = parallel.world_size()
        if w <= 1:
            return 1.0
        mode = self.__dict__.get("_native_dp")
        if mode == "fused":
            return 1.0 / w
        if mode == "call":
            _lib.check(self.lib.crnn_allreduce_grads(self.handle, None, self._stream()))
            return 1.0 / w
        return parallel.allreduce_sum_(self.tensor("arena/grads"))

    def _stage(self, key, arr, dtype):
        """host numpy -> pinned staging -> device (async on the current stream)."""
        arr = np.ascontiguousarray(arr)
        t = self._pinned.get(key)
        if t is None or t[0].numel() < arr.size or t[0].dtype != dtype:
            pin = torch.empty(max(arr.size, 1), dtype=dtype).pin_memory()
            dev = torch.empty(max(arr.size, 1), dtype=dtype, device=self.device)
            self._pinned[key] = t = (pin, dev)
        pin, dev = t
        pin[:arr.size].copy_(torch.from_numpy(arr.reshape(-1)).to(dtype))
        dev[:arr.size].copy_(pin[:arr.size], non_blocking=True)
        return dev[:arr.size]

    input_mean, input_std = 118.24236953981779, 36.72835353999682      # utils.py:421 (norm constants of the mjsynth runs)

    def _stage_input(self, x):
        """Host images -> device float32 (B,imgh,imgw,1).  float input: as the reference's generator yields it (already normalised).
        uint8 input (NEW): the raw 8-bit line images are uploaded and `norm` (utils.py:415-416) runs on the device."""
        x = np.asarray(x)
        B = x.shape[0]
        if x.dtype == np.uint8:
            xu = self._stage("x_u8", x, torch.uint8)
            t = self._pinned.get("x_f32")
            if t is None or t.numel() < xu.numel():
                t = self._pinned["x_f32"] = torch.empty(max(xu.numel(), self.max_batch * self.imgh * self.imgw), dtype=torch.float32, device=self.device)
            _lib.check(self.lib.crnn_normalize_u8(xu.data_ptr(), t.data_ptr(), xu.numel(), float(np.float32(self.input_mean)), float(np.float32(self.input_std)), self._stream()))
            return t[:xu.numel()].view(B, self.imgh, self.imgw, 1)
        return self._stage("x", x.astype(np.float32, copy=False), torch.float32).view(B, self.imgh, self.imgw, 1)

    def train_on_batch(self, inputs, outputs=None):
        """Keras train_on_batch on the generator's dict (utils.py:495-502): host buffers in, scalar mean loss out."""
        from . import parallel
        if parallel.world_size() <= 1:
            return self._train_on_batch_host(inputs)
        return self._train_on_batch_staged(inputs)

    def _train_on_batch_staged(self, inputs):
        """The same step assembled from the device-level entry points (torch pinned staging, crnn_train_fwd_bwd, gradient exchange,
        optimiser step): the data-parallel path, where the monitored loss is all-reduced between the step and the read-back."""
        xd = self._stage_input(inputs["the_input"])
        B = xd.shape[0]
        lab = self._stage("labels", np.asarray(inputs["the_labels"]).astype(np.int32), torch.int32)
        ll = self._stage("label_len", np.asarray(inputs["label_length"]).reshape(-1).astype(np.int32), torch.int32)
        il = self._stage("input_len", np.asarray(inputs["input_length"]).reshape(-1).astype(np.int32), torch.int32)
        self._step_seed = (int(self._step_seed) * 6364136223846793005 + 1442695040888963407) % (1 << 64) | 1
        loss = self.train_fwd_bwd_device(xd, lab, ll, il, dropout_seed=self._step_seed if self.dropout else 0)
        if os.environ.get("CRNN_DBG_NAN"):                  # debug hook (tools/dbg_dp_cli.py): where does a non-finite gradient first appear
            self._dbg_step = self.__dict__.get("_dbg_step", 0) + 1
            bad = [n for n in self.shapes if not n.endswith(("moving_mean", "moving_variance")) and not bool(torch.isfinite(self.tensor("grad/" + n)).all())]
            if bad:
                names = ("act/dtheta", "act/dd1", "act/dflat", "act/theta", "act/loc_d1", "act/flat", "act/bn1/scale", "act/bn1/shift", "act/bn1/mean", "act/bn1/invstd",
                         "act/bn2/scale", "act/bn2/shift", "act/bn2/mean", "act/bn2/invstd", "act/dw1", "act/pw1", "act/ddw1", "act/dpw1", "act/ddw2", "act/dpw2", "act/gA", "act/gB", "act/a0")
                acts = []
                for a in names:
                    t = self.tensor(a)
                    nf = int((~torch.isfinite(t)).sum())
                    if nf:
                        acts.append("%s:%d/%d@%d" % (a, nf, t.numel(), int(torch.nonzero(~torch.isfinite(t))[0])))
                b1 = [float(self.tensor("act/bn1/" + k)[0]) for k in ("scale", "shift", "mean", "invstd")]
                print("[dbg-nan] step %d B %d loss %s: non-finite grads in %d tensors %s ; non-finite acts %s ; bn1 %s" % (self._dbg_step, B, float(loss.mean()), len(bad), bad[:12], acts, b1), flush=True)
        scale = self.allreduce_grads()
        self.optimizer_step(scale)
        # D2H of the step's result: per-sample losses + the CTC feasibility status in one pinned buffer, ONE stream synchronisation
        pin = self.__dict__.get("_result_pin")
        if pin is None or pin[0].numel() < B:
            pin = self._result_pin = (torch.empty(max(B, self.max_batch), dtype=torch.float32).pin_memory(), torch.empty(1, dtype=torch.int32).pin_memory())
        from . import parallel
        if parallel.world_size() > 1:
            # data parallel: every rank reports the GLOBAL mean loss, so loss-driven callbacks (EarlyStoppingIter, ModelCheckpoint) decide
            # identically on all ranks and nobody leaves the loop while the others wait in the next gradient all-reduce
            gl = self.__dict__.get("_global_loss")
            if gl is None:
                gl = self._global_loss = torch.empty(1, dtype=torch.float32, device=self.device)
            gl.copy_(loss.mean(dtype=torch.float32).reshape(1))
            parallel.allreduce_mean_(gl)
            loss = gl.expand(B)
        pin[0][:B].copy_(loss, non_blocking=True)
        pin[1].copy_(self.tensor("act/status")[:1].view(torch.int32), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        st = int(pin[1][0])
        if st != 0:
            raise ValueError(f"Not enough time for target transition sequence (batch element {-st - 1})")
        return float(pin[0][:B].numpy().mean(dtype=np.float64))

    def _train_on_batch_host(self, inputs):
        """Single process: the whole step is ONE C-ABI call on the host arrays (crnn_train_on_batch_host: staging, H2D, forward + backward,
        optimiser, D2H of losses + CTC status, one synchronisation)."""
        o = self.optimizer
        if o is None:
            raise RuntimeError("compile(optimizer=...) first")
        x = np.ascontiguousarray(inputs["the_input"])
        if x.dtype != np.uint8:
            x = np.ascontiguousarray(x, dtype=np.float32)
        B = int(x.shape[0])
        if x.size != B * self.imgh * self.imgw:
            raise ValueError(f"the_input has {x.size} elements, expected {B}x{self.imgh}x{self.imgw}")
        lab = np.ascontiguousarray(inputs["the_labels"], dtype=np.int32).reshape(B, -1)
        if lab.shape[1] != self.max_len:
            raise ValueError(f"the_labels must be (B, {self.max_len}), got {lab.shape}")
        ll = np.ascontiguousarray(inputs["label_length"], dtype=np.int32).reshape(-1)
        il = np.ascontiguousarray(inputs["input_length"], dtype=np.int32).reshape(-1)
        if ll.size != B or il.size != B:
            raise ValueError("label_length / input_length must have one entry per sample")
        self._step_seed = (int(self._step_seed) * 6364136223846793005 + 1442695040888963407) % (1 << 64) | 1
        so = self.__dict__.get("_opt_struct")
        if so is None:
            so = self._opt_struct = _lib.Optimizer()
        if o.kind == "adam":
            so.kind, so.lr, so.beta1, so.beta2, so.eps, so.clipnorm = 0, o.lr, o.beta_1, o.beta_2, o.epsilon, o.clipnorm or 0.0
        else:
            so.kind, so.lr, so.decay, so.momentum, so.clipnorm = 1, o.lr, o.decay, o.momentum, o.clipnorm or 0.0
        mean, st = ctypes.c_float(), ctypes.c_int32()
        _lib.check(self.lib.crnn_train_on_batch_host(
            self.handle, x.ctypes.data, 1 if x.dtype == np.uint8 else 0, float(np.float32(self.input_mean)), float(np.float32(self.input_std)),
            lab.ctypes.data, ll.ctypes.data, il.ctypes.data, B, ctypes.c_uint64(self._step_seed if self.dropout else 0), ctypes.byref(so), 1.0, None,
            ctypes.byref(mean), ctypes.byref(st), self._stream()))
        if st.value != 0:
            raise ValueError(f"Not enough time for target transition sequence (batch element {-st.value - 1})")
        return float(mean.value)

    def test_on_batch(self, inputs, outputs=None):
        """Validation loss: inference-mode forward + CTC loss (no gradient)."""
        xd = self._stage_input(inputs["the_input"])
        B = xd.shape[0]
        sm = self.forward_device(xd)
        lab = self._stage("labels", np.asarray(inputs["the_labels"]).astype(np.int32), torch.int32)
        ll = self._stage("label_len", np.asarray(inputs["label_length"]).reshape(-1).astype(np.int32), torch.int32)
        il = self._stage("input_len", np.asarray(inputs["input_length"]).reshape(-1).astype(np.int32), torch.int32)
        from .ctc import ctc_batch_cost_device
        return float(ctc_batch_cost_device(sm, lab.view(B, -1), ll, il, t_off=2).mean().item())

    def fit_generator(self, generator, steps_per_epoch, epochs=1, validation_data=None, validation_steps=None,
                      shuffle=False, verbose=1, callbacks=None, **_):
        """Subset of Keras fit_generator that train.py:201-209 relies on."""
        hist = _History()
        callbacks = callbacks or []
        for cb in callbacks:
            cb.model = self
            getattr(cb, "on_train_begin", lambda logs=None: None)({})
        self.stop_training = False
        for epoch in range(epochs):
            run = 0.0
            for step in range(steps_per_epoch):
                inputs, targets = next(generator)
                loss = self.train_on_batch(inputs, targets)
                run += loss
                for cb in callbacks:
                    getattr(cb, "on_batch_end", lambda b, logs=None: None)(step, {"loss": loss})
                if verbose and (step % 50 == 0 or step == steps_per_epoch - 1):
                    print(f"Epoch {epoch + 1}/{epochs} step {step + 1}/{steps_per_epoch} - loss: {run / (step + 1):.4f}", flush=True)
                if self.stop_training:
                    break
            logs = {"loss": run / max(1, step + 1)}
            hist.history["loss"].append(logs["loss"])
            if validation_data is not None and validation_steps and not self.stop_training:
                v = sum(self.test_on_batch(*next(validation_data)) for _ in range(validation_steps)) / validation_steps
                logs["val_loss"] = v
                hist.history.setdefault("val_loss", []).append(v)
            for cb in callbacks:
                getattr(cb, "on_epoch_end", lambda e, logs=None: None)(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            getattr(cb, "on_train_end", lambda logs=None: None)({})
        return hist


class CRNN:
    """CRNN(num_classes=97, max_string_len=23, shape=(40,40,1), time_dense_size=128, GRU=False, n_units=256)
    -- same constructor as the reference (utils.py:34-41); get_model() returns the B200 engine."""

    def __init__(self, num_classes=97, max_string_len=23, shape=(40, 40, 1), time_dense_size=128, GRU=False, n_units=256,
                 max_batch=64, seed=None):
        self.num_classes, self.shape, self.max_string_len = num_classes, shape, max_string_len
        self.n_units, self.GRU, self.time_dense_size = n_units, GRU, time_dense_size
        self.max_batch, self.seed = max_batch, seed

    def get_model(self):
        self.pooling_counter_h, self.pooling_counter_w = 1, 2      # utils.py:52-55 for the fixed block plan
        return CRNNModel(self.num_classes, self.max_string_len, self.shape, self.time_dense_size, bool(self.GRU), self.n_units,
                         max_batch=self.max_batch, seed=self.seed)
