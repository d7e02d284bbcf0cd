"""Keras-2.2.2 functional-model JSON of the CRNN graph (what `model.to_json()` returns in the reference, written to
`models/<name>/model.json` by utils.py:530-533 / train.py:169 and stored as the `model_config` attribute of `final_model.h5`).

The graph is fixed by CRNN.get_model (utils.py:58-96), so the JSON is a function of its hyper-parameters only.  `keras_model_config`
rebuilds it layer by layer, with Keras' own auto-generated layer names (one counter per class, in construction order) and the key order
Keras 2.2.2 emits, so that for the shipped hyper-parameters `json.dumps()` of the result is byte-identical to the reference's
`models/OCR_mjsynth_FULL_2/model.json` (tests/test_host_logic.py::test_to_json_is_the_reference_model_json).  A directory written by this
repo's train.py is therefore loadable by the reference's predict.py (Keras `model_from_json` + `load_weights`), and vice versa."""
from collections import OrderedDict

BLOCKS = [(64, None), (128, None), (256, (2, 2)), (256, None), (512, (1, 2)), (512, None), (512, None)]     # utils.py:64-70

# Keras serialises a Lambda layer's Python function as base64(marshal(code object)) of the interpreter that built the model.  The payload
# below is the one every model.json shipped with the reference carries for `ctc_lambda_func` (utils.py:98-103, CPython 3.6 bytecode); it is
# an opaque constant of the file format here -- this package never executes it (the CTC stage is csrc/ctc.cu) -- but Keras 2.2.2 needs it
# to deserialise the `ctc` layer when it loads the full training graph.
_CTC_LAMBDA_PY36 = ("4wEAAAAAAAAABQAAAAUAAABDAAAAczYAAAB8AFwEfQF9An0DfQR8AWQAZACFAmQBZACFAmQAZACF\n"
                    "AmYDGQB9AXQAagF8AnwBfAN8BIMEUwApAk7pAgAAACkC2gFL2g5jdGNfYmF0Y2hfY29zdCkF2gRh\n"
                    "cmdz2gZ5X3ByZWTaBmxhYmVsc9oMaW5wdXRfbGVuZ3Ro2gxsYWJlbF9sZW5ndGipAHIJAAAA+iIv\n"
                    "ZGF0YS9kYXRhL0NSTk5fT0NSX2tlcmFzL3V0aWxzLnB52g9jdGNfbGFtYmRhX2Z1bmNiAAAAcwYA\n"
                    "AAAAAQwDGgE=\n")


def _od(*pairs):
    return OrderedDict(pairs)


def _glorot():
    return _od(("class_name", "VarianceScaling"), ("config", _od(("scale", 1.0), ("mode", "fan_avg"), ("distribution", "uniform"), ("seed", None))))


def _he_normal():
    return _od(("class_name", "VarianceScaling"), ("config", _od(("scale", 2.0), ("mode", "fan_in"), ("distribution", "normal"), ("seed", None))))


def _init(name):
    return _od(("class_name", name), ("config", _od()))


def _layer(name, cls, config, inbound):
    return _od(("name", name), ("class_name", cls), ("config", config), ("inbound_nodes", [[[src, 0, 0, {}] for src in inbound]] if inbound else []))


def _input(name, shape, dtype):
    return _layer(name, "InputLayer", _od(("batch_input_shape", [None] + list(shape)), ("dtype", dtype), ("sparse", False), ("name", name)), None)


def _conv2d(name, filters, k, padding, use_bias, src):
    return _layer(name, "Conv2D", _od(("name", name), ("trainable", True), ("filters", filters), ("kernel_size", [k, k]), ("strides", [1, 1]), ("padding", padding),
                                      ("data_format", "channels_last"), ("dilation_rate", [1, 1]), ("activation", "linear"), ("use_bias", use_bias),
                                      ("kernel_initializer", _glorot()), ("bias_initializer", _init("Zeros")), ("kernel_regularizer", None), ("bias_regularizer", None),
                                      ("activity_regularizer", None), ("kernel_constraint", None), ("bias_constraint", None)), [src])


def _dense(name, units, activation, kernel_init, src):
    return _layer(name, "Dense", _od(("name", name), ("trainable", True), ("units", units), ("activation", activation), ("use_bias", True), ("kernel_initializer", kernel_init),
                                     ("bias_initializer", _init("Zeros")), ("kernel_regularizer", None), ("bias_regularizer", None), ("activity_regularizer", None),
                                     ("kernel_constraint", None), ("bias_constraint", None)), [src])


def _maxpool(name, pool, src):
    return _layer(name, "MaxPooling2D", _od(("name", name), ("trainable", True), ("pool_size", list(pool)), ("padding", "valid"), ("strides", list(pool)),
                                            ("data_format", "channels_last")), [src])


def _dropout(name, rate, src):
    return _layer(name, "Dropout", _od(("name", name), ("trainable", True), ("rate", rate), ("noise_shape", None), ("seed", None)), [src])


def _bn(name, src):
    return _layer(name, "BatchNormalization", _od(("name", name), ("trainable", True), ("axis", -1), ("momentum", 0.99), ("epsilon", 0.001), ("center", True), ("scale", True),
                                                  ("beta_initializer", _init("Zeros")), ("gamma_initializer", _init("Ones")), ("moving_mean_initializer", _init("Zeros")),
                                                  ("moving_variance_initializer", _init("Ones")), ("beta_regularizer", None), ("gamma_regularizer", None),
                                                  ("beta_constraint", None), ("gamma_constraint", None)), [src])


def _rnn_cell(cell, idx, units):
    name = "%s_%d" % (cell, idx)
    cfg = [("name", name), ("trainable", True), ("return_sequences", True), ("return_state", False), ("go_backwards", False), ("stateful", False), ("unroll", False),
           ("units", units), ("activation", "tanh"), ("recurrent_activation", "hard_sigmoid"), ("use_bias", True), ("kernel_initializer", _he_normal()),
           ("recurrent_initializer", _od(("class_name", "Orthogonal"), ("config", _od(("gain", 1.0), ("seed", None))))), ("bias_initializer", _init("Zeros"))]
    if cell == "lstm":
        cfg.append(("unit_forget_bias", True))
    cfg += [("kernel_regularizer", None), ("recurrent_regularizer", None), ("bias_regularizer", None), ("activity_regularizer", None), ("kernel_constraint", None),
            ("recurrent_constraint", None), ("bias_constraint", None), ("dropout", 0.0), ("recurrent_dropout", 0.0), ("implementation", 1)]
    if cell == "gru":
        cfg.append(("reset_after", False))
    return _od(("class_name", "GRU" if cell == "gru" else "LSTM"), ("config", _od(*cfg)))


def keras_model_config(imgh, imgw, num_classes, max_string_len, time_dense_size=128, n_units=256, cell="gru"):
    """The dict Keras 2.2.2 `Model.get_config()`-wraps into `to_json()` for CRNN(...).get_model() (utils.py:58-96)."""
    L = []
    L.append(_input("the_input", (imgh, imgw, 1), "float32"))
    # STN(), utils.py:247-258
    L.append(_maxpool("max_pooling2d_1", (2, 2), "the_input"))
    L.append(_conv2d("conv2d_1", 20, 5, "valid", True, "max_pooling2d_1"))
    L.append(_maxpool("max_pooling2d_2", (2, 2), "conv2d_1"))
    L.append(_conv2d("conv2d_2", 20, 5, "valid", True, "max_pooling2d_2"))
    L.append(_layer("flatten_1", "Flatten", _od(("name", "flatten_1"), ("trainable", True), ("data_format", "channels_last")), ["conv2d_2"]))
    L.append(_dense("dense_1", 50, "linear", _glorot(), "flatten_1"))
    L.append(_layer("activation_1", "Activation", _od(("name", "activation_1"), ("trainable", True), ("activation", "relu")), ["dense_1"]))
    L.append(_dense("dense_2", 6, "linear", _glorot(), "activation_1"))
    L.append(_layer("bilinear_interpolation_1", "BilinearInterpolation", _od(("name", "bilinear_interpolation_1"), ("trainable", True), ("output_size", [imgh, imgw])),
                    ["the_input", "dense_2"]))
    L.append(_layer("zero_padding2d_1", "ZeroPadding2D", _od(("name", "zero_padding2d_1"), ("trainable", True), ("padding", [[2, 2], [2, 2]]), ("data_format", "channels_last")),
                    ["bilinear_interpolation_1"]))
    src, pools = "zero_padding2d_1", 2
    for i, (cout, pool) in enumerate(BLOCKS, 1):                         # depthwise_conv_block, utils.py:43-56
        dw = "depthwise_conv2d_%d" % i
        L.append(_layer(dw, "DepthwiseConv2D", _od(("name", dw), ("trainable", True), ("kernel_size", [3, 3]), ("strides", [1, 1]), ("padding", "same"),
                                                    ("data_format", "channels_last"), ("dilation_rate", [1, 1]), ("activation", "linear"), ("use_bias", False),
                                                    ("bias_initializer", _init("Zeros")), ("bias_regularizer", None), ("activity_regularizer", None), ("bias_constraint", None),
                                                    ("depth_multiplier", 1), ("depthwise_initializer", _glorot()), ("depthwise_regularizer", None), ("depthwise_constraint", None)),
                        [src]))
        L.append(_bn("batch_normalization_%d" % (2 * i - 1), dw))
        L.append(_layer("re_lu_%d" % (2 * i - 1), "ReLU", _od(("name", "re_lu_%d" % (2 * i - 1)), ("trainable", True), ("max_value", 6.0)), ["batch_normalization_%d" % (2 * i - 1)]))
        L.append(_conv2d("conv2d_%d" % (i + 2), cout, 1, "same", False, "re_lu_%d" % (2 * i - 1)))
        L.append(_bn("batch_normalization_%d" % (2 * i), "conv2d_%d" % (i + 2)))
        L.append(_layer("re_lu_%d" % (2 * i), "ReLU", _od(("name", "re_lu_%d" % (2 * i)), ("trainable", True), ("max_value", 6.0)), ["batch_normalization_%d" % (2 * i)]))
        src = "re_lu_%d" % (2 * i)
        if pool is not None:
            pools += 1
            L.append(_maxpool("max_pooling2d_%d" % pools, pool, src))
            src = "max_pooling2d_%d" % pools
        L.append(_dropout("dropout_%d" % i, 0.1, src))
        src = "dropout_%d" % i
    T, feat = (imgh + 4) // 2, ((imgw + 4) // 4) * 512                  # conv_to_rnn_dims, utils.py:72
    L.append(_layer("reshape", "Reshape", _od(("name", "reshape"), ("trainable", True), ("target_shape", [T, feat])), [src]))
    L.append(_dense("dense1", time_dense_size, "relu", _glorot(), "reshape"))
    L.append(_dropout("dropout_8", 0.4, "dense1"))
    L.append(_layer("bidirectional_1", "Bidirectional", _od(("name", "bidirectional_1"), ("trainable", True), ("layer", _rnn_cell(cell, 1, n_units)), ("merge_mode", "sum")), ["dropout_8"]))
    L.append(_layer("bidirectional_2", "Bidirectional", _od(("name", "bidirectional_2"), ("trainable", True), ("layer", _rnn_cell(cell, 2, n_units)), ("merge_mode", "concat")), ["bidirectional_1"]))
    L.append(_dropout("dropout_9", 0.2, "bidirectional_2"))
    L.append(_dense("dense2", num_classes, "linear", _he_normal(), "dropout_9"))
    L.append(_layer("softmax", "Activation", _od(("name", "softmax"), ("trainable", True), ("activation", "softmax")), ["dense2"]))
    L.append(_input("the_labels", (max_string_len,), "float32"))
    L.append(_input("input_length", (1,), "int64"))
    L.append(_input("label_length", (1,), "int64"))
    L.append(_layer("ctc", "Lambda", _od(("name", "ctc"), ("trainable", True), ("function", [_CTC_LAMBDA_PY36, None, None]), ("function_type", "lambda"), ("output_shape", [1]),
                                         ("output_shape_type", "raw"), ("arguments", {})), ["softmax", "the_labels", "input_length", "label_length"]))
    return _od(("class_name", "Model"),
               ("config", _od(("name", "model_1"), ("layers", L),
                              ("input_layers", [["the_input", 0, 0], ["the_labels", 0, 0], ["input_length", 0, 0], ["label_length", 0, 0]]),
                              ("output_layers", [["ctc", 0, 0]]))),
               ("keras_version", "2.2.2"), ("backend", "tensorflow"))
