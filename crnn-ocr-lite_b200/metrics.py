"""Edit-distance evaluation of the reference (utils.py:262-298; predict.py:183-191)."""
import numpy as np


def levenshtein(seq1, seq2):
    """Classic O(n*m) edit distance with a rolling row (the reference fills the full numpy matrix)."""
    n, m = len(seq1), len(seq2)
    prev = list(range(m + 1))
    for i in range(1, n + 1):
        cur = [i] + [0] * m
        a = seq1[i - 1]
        for j in range(1, m + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a != seq2[j - 1]))
        prev = cur
    return float(prev[m])


def edit_distance(y_pred, y_true):
    # plain left-to-right accumulation like utils.py:289-293 (the builtin sum() is compensated since CPython 3.12 and differs in the last
    # ulp -- caught by the golden vectors produced with the reference's own function)
    n = len(y_true)
    mean = 0
    for a, b in zip(y_pred, y_true):
        mean += levenshtein(a, b) / n
    return mean


def normalized_edit_distance(y_pred, y_true):
    n = len(y_true)
    mean = 0
    for a, b in zip(y_pred, y_true):
        mean += levenshtein(a, b) / (len(b) * n)
    return mean


# ---------------------------------------------------------------- CUDA evaluation step (include/crnn_b200.h: crnn_edit_distance_host)
def _encode(seqs, maxlen):
    """list of strings / symbol sequences -> (int32 [N, maxlen] padded with -1, int32 [N] lengths)."""
    n = len(seqs)
    a = np.full((n, maxlen), -1, np.int32)
    ln = np.zeros(n, np.int32)
    for i, q in enumerate(seqs):
        v = [ord(c) for c in q] if isinstance(q, str) else [int(c) for c in q]
        ln[i] = len(v)
        a[i, :len(v)] = v
    return a, ln


def levenshtein_batch_cuda(y_pred, y_true):
    """Levenshtein distance of every (prediction, truth) pair on the GPU (one thread per pair); int32 array, bit-identical to
    levenshtein() / the reference's utils.py:262-287.  No CPU fallback: raises CrnnError without the CUDA library or a device."""
    import ctypes
    import torch
    from . import _lib
    y_pred, y_true = list(y_pred), list(y_true)
    if len(y_pred) != len(y_true):
        raise ValueError("y_pred and y_true must have the same length")
    if not y_true:
        return np.zeros(0, np.int32)
    if not torch.cuda.is_available():
        raise _lib.CrnnError("no CUDA device: the batched evaluation has no CPU fallback (use levenshtein() for host-side checks)")
    lib = _lib.load()
    maxlen = max(1, max(len(q) for q in y_pred), max(len(q) for q in y_true))
    if maxlen > 128:
        raise ValueError("sequences longer than 128 symbols are not supported by crnn_edit_distance")
    a, al = _encode(y_pred, maxlen)
    b, bl = _encode(y_true, maxlen)
    out = np.empty(len(y_true), np.int32)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.crnn_edit_distance_host(p(a), p(al), p(b), p(bl), len(y_true), maxlen, p(out), st))
    return out


def edit_distance_cuda(y_pred, y_true):
    """utils.py:289-293 with the distances from the GPU; same summation order and float64 arithmetic as the reference."""
    d = levenshtein_batch_cuda(y_pred, y_true)
    n = len(y_true)
    mean = 0
    for v in d:
        mean += float(v) / n
    return mean


def normalized_edit_distance_cuda(y_pred, y_true):
    """utils.py:295-299 with the distances from the GPU."""
    d = levenshtein_batch_cuda(y_pred, y_true)
    n = len(y_true)
    mean = 0
    for v, y in zip(d, y_true):
        mean += float(v) / (len(y) * n)
    return mean
