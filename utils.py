"""Drop-in for the reference's `utils` module (gasparian/CRNN-OCR-lite utils.py): the same public names, backed by the
B200-native path in crnn-ocr-lite_b200/.  `from utils import *` (reference train.py:119) keeps working, including the
names the reference's train.py picks up implicitly through that star import (`re`, `optimizers`, `np`, `os`)."""
import os  # noqa: F401
import re  # noqa: F401
import types

import numpy as np  # noqa: F401

import crnn_b200 as _cb

CRNN = _cb.CRNN
DecodeCTCPred = _cb.DecodeCTCPred
labels_to_text = _cb.labels_to_text
load_custom_model = _cb.load_custom_model
load_model_custom = _cb.load_model_custom
init_predictor = _cb.init_predictor
save_model_json = _cb.save_model_json
Readf, open_img, read_img, norm = _cb.Readf, _cb.open_img, _cb.read_img, _cb.norm
parse_mjsynth, get_lexicon, get_lengths, make_ohe = _cb.parse_mjsynth, _cb.get_lexicon, _cb.get_lengths, _cb.make_ohe
levenshtein, edit_distance, normalized_edit_distance = _cb.levenshtein, _cb.edit_distance, _cb.normalized_edit_distance
# GPU-batched evaluation step (include/crnn_b200.h crnn_edit_distance_host): same numbers, one kernel instead of O(N L^2) Python
levenshtein_batch_cuda, edit_distance_cuda, normalized_edit_distance_cuda = _cb.levenshtein_batch_cuda, _cb.edit_distance_cuda, _cb.normalized_edit_distance_cuda
EarlyStoppingIter, ModelCheckpoint = _cb.EarlyStoppingIter, _cb.ModelCheckpoint
optimizers = types.SimpleNamespace(Adam=_cb.Adam, SGD=_cb.SGD)   # keras.optimizers subset (train.py:188-190)


def ctc_lambda_func(args):
    """utils.py:98-103 on device tensors: y_pred[:, 2:, :] -> K.ctc_batch_cost."""
    y_pred, labels, input_length, label_length = args
    return _cb.ctc_batch_cost_device(y_pred, labels, label_length, input_length, t_off=2).view(-1, 1)


BilinearInterpolation, STN = _cb.BilinearInterpolation, _cb.STN    # utils.py:116-258 as callables on arrays (csrc/stn.cu sampler kernel)


def get_initial_weights(output_size):
    """utils.py:239-245: [W = 0, b = identity affine] of the localisation head's last Dense layer."""
    b = np.zeros((2, 3), dtype="float32")
    b[0, 0] = 1
    b[1, 1] = 1
    return [np.zeros((output_size, 6), dtype="float32"), b.flatten()]


def K_linspace(start, stop, num):
    """utils.py:113-114 (tf.linspace) as the sampler kernels evaluate it: start + step * i in float32."""
    step = np.float32((np.float32(stop) - np.float32(start)) / np.float32(num - 1))
    return (np.float32(start) + step * np.arange(num, dtype=np.float32)).astype(np.float32)


def K_meshgrid(x, y):
    """utils.py:110-111 (tf.meshgrid)."""
    return np.meshgrid(x, y)


class GRU:       # noqa: D401 - names that `from utils import *` used to bring in from keras.layers (utils.py:20): the reference's train.py
    """passes `GRU` to CRNN(GRU=...) AFTER the star import, i.e. this (truthy) class object and never the --GRU flag."""


class LSTM:
    """see GRU."""

