#!/usr/bin/env python
"""tests/golden/reference_mjsynth_model.json = the reference's own models/OCR_mjsynth_FULL_2/model.json (the Keras-2.2.2 `model.to_json()` of
CRNN(num_classes=38, max_string_len=23, shape=(100,32,1), GRU) written by utils.py:530-533), copied verbatim as the known answer for
crnn-ocr-lite_b200/keras_json.py.  The IAM / Stickies files differ only in max_string_len (21 / 20)."""
import shutil
shutil.copy("/root/reference/models/OCR_mjsynth_FULL_2/model.json", __file__.replace("make_model_json_golden.py", "reference_mjsynth_model.json"))
