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


def STN(*_a, **_k):
    raise NotImplementedError("STN / BilinearInterpolation are fused stages of the engine (csrc/stn.cu); build the model with CRNN(...).get_model()")


BilinearInterpolation = STN
