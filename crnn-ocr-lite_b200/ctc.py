"""CTC stage of the reference behind its own names (gasparian/CRNN-OCR-lite utils.py:98-103, 314-321, 331-357):
ctc_batch_cost / ctc_decode on the B200 kernels (csrc/ctc.cu) + DecodeCTCPred / labels_to_text."""
import ctypes

import numpy as np
import torch

from . import _lib

K_EPS = 1e-7   # keras.backend.epsilon()


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def ctc_batch_cost_device(probs, labels, label_len, input_len, t_off=0, want_grad_u=False, want_grad_logits=False, scale=1.0, eps=K_EPS):
    """K.ctc_batch_cost on probs[:, t_off:, :] (utils.py:102-103).  probs (B,T,V) CUDA f32; labels (B,maxL) i32."""
    lib = _lib.load()
    B, T, V = probs.shape
    assert probs.is_cuda and probs.dtype == torch.float32 and probs.is_contiguous()
    labels = labels.to(torch.int32).contiguous(); label_len = label_len.to(torch.int32).contiguous().view(-1)
    input_len = input_len.to(torch.int32).contiguous().view(-1)
    dev = probs.device
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    gu = torch.empty(B, T - t_off, V, dtype=torch.float32, device=dev) if want_grad_u else None
    gz = torch.empty(B, T, V, dtype=torch.float32, device=dev) if want_grad_logits else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.crnn_ctc_loss_grad(_ptr(probs), B, T, V, t_off, _ptr(labels), labels.shape[1], _ptr(label_len), _ptr(input_len),
                                      float(eps), _ptr(loss), _ptr(gu), _ptr(gz), float(scale), _ptr(status), _stream(dev)))
    st = int(status.item())
    if st != 0:
        raise ValueError(f"Not enough time for target transition sequence (batch element {-st - 1})")
    if want_grad_u or want_grad_logits:
        return loss, gu, gz
    return loss


def ctc_decode_device(probs, seq_len=None, greedy=True, beam_width=100, merge_repeated=True, top_paths=1):
    """K.ctc_decode(y_pred, input_length, greedy, beam_width, top_paths): returns (dense (B,T) padded -1, lengths (B,), score (B,)); with
    top_paths = P > 1 (beam search only) the shapes are (B,P,T), (B,P), (B,P), best path first."""
    lib = _lib.load()
    B, T, V = probs.shape
    assert probs.is_cuda and probs.dtype == torch.float32 and probs.is_contiguous()
    dev = probs.device
    out = torch.empty(B, T, dtype=torch.int32, device=dev)
    n = torch.empty(B, dtype=torch.int32, device=dev)
    score = torch.empty(B, dtype=torch.float32, device=dev)
    sl = seq_len.to(torch.int32).contiguous() if seq_len is not None else None
    if greedy:
        _lib.check(lib.crnn_ctc_greedy(_ptr(probs), _ptr(sl), B, T, V, K_EPS, _ptr(out), _ptr(n), _ptr(score), _stream(dev)))
    elif int(top_paths) > 1:
        P = int(top_paths)
        out = torch.empty(B, P, T, dtype=torch.int32, device=dev)
        n = torch.empty(B, P, dtype=torch.int32, device=dev)
        score = torch.empty(B, P, dtype=torch.float32, device=dev)
        _lib.check(lib.crnn_ctc_beam_topk(_ptr(probs), _ptr(sl), B, T, V, K_EPS, max(int(beam_width), P), int(bool(merge_repeated)), P,
                                          _ptr(out), _ptr(n), _ptr(score), _stream(dev)))
    else:
        _lib.check(lib.crnn_ctc_beam(_ptr(probs), _ptr(sl), B, T, V, K_EPS, int(beam_width), int(bool(merge_repeated)),
                                     _ptr(out), _ptr(n), _ptr(score), _stream(dev)))
    return out, n, score


def ctc_decode_host(probs, greedy=False, beam_width=10, merge_repeated=True):
    """Host numpy (N,T,V) softmax -> host numpy labels: H2D + decode + D2H inside the C ABI call."""
    lib = _lib.load()
    if not torch.cuda.is_available():
        raise _lib.CrnnError("ctc_decode needs a CUDA device: no CPU fallback on this path")
    probs = np.ascontiguousarray(probs, np.float32)
    B, T, V = probs.shape
    out = np.empty((B, T), np.int32); n = np.empty(B, np.int32); score = np.empty(B, np.float32)
    st = _stream(torch.device("cuda", torch.cuda.current_device()))
    if greedy:
        _lib.check(lib.crnn_ctc_greedy_host(probs.ctypes.data, B, T, V, K_EPS, out.ctypes.data, n.ctypes.data, score.ctypes.data, st))
    else:
        _lib.check(lib.crnn_ctc_beam_host(probs.ctypes.data, B, T, V, K_EPS, int(beam_width), int(bool(merge_repeated)),
                                          out.ctypes.data, n.ctypes.data, score.ctypes.data, st))
    return out, n, score


def labels_to_text(labels, inverse_classes=None):
    """utils.py:314-321."""
    ret = []
    for c in labels:
        if c == len(inverse_classes) or c == -1:
            ret.append("")
        else:
            ret.append(str(inverse_classes[c]))
    return "".join(ret)


class DecodeCTCPred:
    """utils.py:331-357: same constructor / decode(result) contract (top-1 strings), one batched GPU beam search
    instead of one TF graph + CPU op per sample.  `greedy=True` is an added switch (BASELINE configs[0,1])."""

    def __init__(self, top_paths=1, beam_width=5, inverse_classes=None, greedy=False):
        self.top_paths = top_paths
        self.beam_width = beam_width
        self.inverse_classes = inverse_classes
        self.greedy = greedy

    def labels_to_text(self, labels):
        return labels_to_text(labels, self.inverse_classes)

    def decode(self, result):
        # top_paths > 1 only widens the beam in the reference: decode() still keeps `[0][0]`, the best path (utils.py:353-356)
        if self.beam_width < self.top_paths:
            self.beam_width = self.top_paths
        result = np.asarray(result, np.float32)
        if result.ndim == 2:
            result = result[None]
        labels, n, _ = ctc_decode_host(result, greedy=self.greedy, beam_width=self.beam_width, merge_repeated=True)
        return [self.labels_to_text(labels[i, :n[i]]) for i in range(labels.shape[0])]


class BilinearInterpolation:
    """utils.py:116-237 as a callable on arrays: `BilinearInterpolation(output_size)([image, theta])` -> sampled image, run by the engine's
    sampler kernel (csrc/stn.cu through crnn_bilinear_sample).  image (B,H,W,1) or (B,H,W), theta (B,6); torch CUDA tensors are processed in
    place on the device, numpy arrays are copied there and back.  output_size must equal the image size (the only use in the reference:
    STN(image, sampling_size) with sampling_size = the input size, utils.py:59-62,257)."""

    def __init__(self, output_size, **_):
        self.output_size = tuple(int(v) for v in output_size)

    def compute_output_shape(self, input_shapes):
        return (None, self.output_size[0], self.output_size[1], input_shapes[0][-1])

    def get_config(self):
        return {"output_size": self.output_size}

    def __call__(self, tensors):
        X, theta = tensors
        if not torch.cuda.is_available():
            raise _lib.CrnnError("BilinearInterpolation needs a CUDA device: no CPU fallback on this path")
        lib = _lib.load()
        as_numpy = not torch.is_tensor(X)
        x = torch.as_tensor(np.ascontiguousarray(X, np.float32) if as_numpy else X, dtype=torch.float32, device="cuda")
        th = torch.as_tensor(np.ascontiguousarray(theta, np.float32) if not torch.is_tensor(theta) else theta, dtype=torch.float32, device=x.device).reshape(-1, 6).contiguous()
        shp = tuple(x.shape)
        if len(shp) == 4:
            if shp[-1] != 1:
                raise ValueError("BilinearInterpolation: single-channel images only (the CRNN input is grayscale)")
            x3 = x.reshape(shp[:3])
        else:
            x3 = x
        B, H, W = x3.shape
        if (H, W) != self.output_size:
            raise ValueError("BilinearInterpolation: output_size %s must equal the image size %s" % (self.output_size, (H, W)))
        if th.shape[0] != B:
            raise ValueError("BilinearInterpolation: theta must be (B, 6)")
        x3 = x3.contiguous()
        out = torch.empty_like(x3)
        _lib.check(lib.crnn_bilinear_sample(_ptr(x3), _ptr(th), _ptr(out), B, H, W, _stream(x3.device)))
        out = out.reshape(shp)
        return out.cpu().numpy() if as_numpy else out


def STN(image, sampling_size):
    """utils.py:247-258 at construction time: the localisation head's last Dense layer starts as W = 0, b = identity affine
    (get_initial_weights), so a freshly built STN resamples with theta = [1,0,0,0,1,0] whatever its conv weights are -- which, with the
    reference's sampler (scale by size, not size - 1), is NOT the identity map.  Inside a model the trained localisation net is part of
    the engine's forward (csrc/stn.cu)."""
    B = int(image.shape[0])
    theta = np.tile(np.array([1, 0, 0, 0, 1, 0], np.float32), (B, 1))
    return BilinearInterpolation(tuple(sampling_size))([image, theta])
