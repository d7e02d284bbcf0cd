"""ctypes wrapper over oracle/libctc_oracle.so (oracle/ctc_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libctc_oracle.so")
    src = os.path.join(_HERE, "ctc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libctc_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.ctc_oracle_loss_grad.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ctc_loss_grad(probs, labels, label_len, input_len, eps=1e-7):
    """probs (B,T,V) f32 (already sliced [:,2:,:]); labels (B,maxL) int; -> loss (B,), grad_u (B,T,V)."""
    probs = np.ascontiguousarray(probs, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    label_len = np.ascontiguousarray(label_len, np.int32).reshape(-1)
    input_len = np.ascontiguousarray(input_len, np.int32).reshape(-1)
    B, T, V = probs.shape
    loss = np.empty(B, np.float32)
    grad = np.empty_like(probs)
    st = lib().ctc_oracle_loss_grad(_p(probs), B, T, V, _p(labels), labels.shape[1], _p(label_len), _p(input_len),
                                    ctypes.c_float(eps), _p(loss), _p(grad))
    if st != 0:
        raise ValueError(f"Not enough time for target transition sequence (batch element {-st - 1})")
    return loss, grad


def greedy(probs, seq_len=None, eps=1e-7):
    probs = np.ascontiguousarray(probs, np.float32)
    B, T, V = probs.shape
    seq_len = np.full(B, T, np.int32) if seq_len is None else np.ascontiguousarray(seq_len, np.int32)
    out = np.empty((B, T), np.int32)
    out_len = np.empty(B, np.int32)
    score = np.empty(B, np.float32)
    lib().ctc_oracle_greedy(_p(probs), B, T, V, _p(seq_len), ctypes.c_float(eps), _p(out), _p(out_len), _p(score))
    return out, out_len, score


def beam(probs, seq_len=None, beam_width=10, merge_repeated=True, eps=1e-7):
    probs = np.ascontiguousarray(probs, np.float32)
    B, T, V = probs.shape
    seq_len = np.full(B, T, np.int32) if seq_len is None else np.ascontiguousarray(seq_len, np.int32)
    out = np.empty((B, T), np.int32)
    out_len = np.empty(B, np.int32)
    lp = np.empty(B, np.float32)
    lib().ctc_oracle_beam(_p(probs), B, T, V, _p(seq_len), ctypes.c_float(eps), int(beam_width), int(bool(merge_repeated)),
                          _p(out), _p(out_len), _p(lp))
    return out, out_len, lp


def beam_topk(probs, top_paths, seq_len=None, beam_width=10, merge_repeated=True, eps=1e-7):
    """K.ctc_decode(greedy=False, beam_width, top_paths): out (B,P,T), out_len (B,P), log_prob (B,P), best path first."""
    probs = np.ascontiguousarray(probs, np.float32)
    B, T, V = probs.shape
    P = int(top_paths)
    seq_len = np.full(B, T, np.int32) if seq_len is None else np.ascontiguousarray(seq_len, np.int32)
    out = np.empty((B, P, T), np.int32)
    out_len = np.empty((B, P), np.int32)
    lp = np.empty((B, P), np.float32)
    lib().ctc_oracle_beam_topk(_p(probs), B, T, V, _p(seq_len), ctypes.c_float(eps), int(beam_width), int(bool(merge_repeated)), P,
                               _p(out), _p(out_len), _p(lp))
    return out, out_len, lp


def beam_threaded(probs, n_threads, **kw):
    """Batch-parallel driver (ctypes releases the GIL): the oracle run on `n_threads` host threads."""
    from concurrent.futures import ThreadPoolExecutor
    probs = np.ascontiguousarray(probs, np.float32)
    chunks = np.array_split(np.arange(probs.shape[0]), n_threads)
    with ThreadPoolExecutor(n_threads) as ex:
        parts = list(ex.map(lambda idx: beam(probs[idx[0]:idx[-1] + 1], **kw) if len(idx) else None, chunks))
    parts = [p for p in parts if p is not None]
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
