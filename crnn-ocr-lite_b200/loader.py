"""Model topology / weight loading with the reference's file layout (utils.py:300-329, 530-533):
`models/<name>/model.json` (Keras 2.2.2 functional-model JSON, or this repo's own compact JSON) + `final_weights.h5`."""
import json

from .model import CRNN


def _config_from_json(text):
    """Pull the CRNN hyper-parameters out of a Keras model.json (or of CRNNModel.to_json())."""
    d = json.loads(text)
    cfg = d.get("config", {})
    if "layers" not in cfg:      # our own compact format
        return dict(num_classes=cfg["num_classes"], max_string_len=cfg["max_string_len"], shape=tuple(cfg["shape"]),
                    time_dense_size=cfg["time_dense_size"], GRU=cfg["GRU"], n_units=cfg["n_units"])
    out = dict(max_string_len=23, time_dense_size=128, n_units=256, GRU=True, num_classes=None, shape=None)
    for layer in cfg["layers"]:
        name, c = layer.get("name"), layer.get("config", {})
        if name == "the_input":
            out["shape"] = tuple(c["batch_input_shape"][1:])
        elif name == "the_labels":
            out["max_string_len"] = c["batch_input_shape"][1]
        elif name == "dense1":
            out["time_dense_size"] = c["units"]
        elif name == "dense2":
            out["num_classes"] = c["units"]
        elif layer.get("class_name") == "Bidirectional":
            out["GRU"] = c["layer"]["class_name"] == "GRU"
            out["n_units"] = c["layer"]["config"]["units"]
    if out["shape"] is None or out["num_classes"] is None:
        raise ValueError("model.json does not describe a CRNN-OCR-lite graph")
    return out


def model_from_json(text, max_batch=64):
    return CRNN(max_batch=max_batch, **_config_from_json(text)).get_model()


def load_custom_model(model_path, model_name="/model.json", weights="/final_weights.h5", max_batch=64):
    """utils.py:323-329."""
    with open(model_path + model_name, "r") as f:
        model = model_from_json(f.read(), max_batch=max_batch)
    model.load_weights(model_path + weights)
    return model


def load_model_custom(path, weights="model", max_batch=64):
    """utils.py:300-306."""
    with open(path + "/model.json", "r") as f:
        model = model_from_json(f.read(), max_batch=max_batch)
    model.load_weights(path + "/%s.h5" % weights)
    return model


def init_predictor(model):
    """utils.py:308-312: the sub-model the_input -> softmax.  CRNNModel.predict* already is that sub-graph."""
    return model


def save_model_json(model, save_path, model_name):
    """utils.py:530-533."""
    with open(save_path + "/" + model_name + "/model.json", "w") as f:
        f.write(model.to_json())
