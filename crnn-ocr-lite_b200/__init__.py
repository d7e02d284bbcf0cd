"""crnn-ocr-lite_b200: B200-native (sm_100a) implementation of the CRNN-OCR hot path of gasparian/CRNN-OCR-lite
behind the reference's own Python surface (utils.py names).  Import as `crnn_ocr_lite_b200` via the repo-root
`crnn_b200.py` loader (the directory name carries a hyphen)."""
from . import _lib, hdf5_lite, keras_json, parallel  # noqa: F401
from .ctc import DecodeCTCPred, ctc_batch_cost_device, ctc_decode_device, ctc_decode_host, labels_to_text, BilinearInterpolation, STN  # noqa: F401
from .model import CRNN, CRNNModel, Adam, SGD, weight_shapes, keras_initial_weights  # noqa: F401
from .data import Readf, open_img, read_img, norm, parse_mjsynth, get_lexicon, get_lengths, make_ohe  # noqa: F401,E402
from .metrics import levenshtein, edit_distance, normalized_edit_distance  # noqa: F401,E402
from .metrics import levenshtein_batch_cuda, edit_distance_cuda, normalized_edit_distance_cuda  # noqa: F401,E402
from .callbacks import EarlyStoppingIter, ModelCheckpoint, Callback  # noqa: F401,E402
from .loader import load_custom_model, load_model_custom, init_predictor, save_model_json, model_from_json  # noqa: F401,E402
