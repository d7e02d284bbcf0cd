#!/usr/bin/env python
"""Extracts the reference's OWN example inputs/outputs from its README figures (run in the authoring container, needs PIL and
/root/reference) -> tests/golden/reference_examples.npz.

imgs/STN_examples/mjsynth_{1..7}.png of the reference are matplotlib figures with two imshow panels at the model resolution
(100 x 32, ~3.34 screen px per image px, nearest-neighbour): the top panel is the pre-processed network input titled with the
TRUE label, the bottom panel the output of the spatial transformer titled with the label the reference PREDICTED.  These are the
only reference-generated input/output pairs that exist offline (TF 1.8 / Keras 2.2.2 cannot run here), so they pin the oracle
(and through it the CUDA path) against the reference itself -- including its mistakes ("cellist" -> "celist").

What the figure loses: imshow autoscaling maps each panel's min..max to 0..255, i.e. the absolute intensity window (2 numbers per
image) of the uint8 input is unknown.  The fixture stores the displayed 8-bit values; the tests state which window they assume.
Titles were transcribed by eye (they are text rendered in the figure).
"""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/imgs/STN_examples/mjsynth_%d.png"
SRC_IAM = "/root/reference/imgs/STN_examples/IAM_%d.png"            # same figure layout, produced with models/OCR_IAM_ver1
TRUE_IAM = ["expressed", "mr.", "to", "intelligence", "surfaces", "effected"]
PRED_IAM = ["expresed", "mr-", "to", "inteligence", "surfaces", "efected"]
TRUE = ["cellist", "conduction", "gropes", "breeziest", "alisha", "mapmakers", "trojan"]
PRED = ["celist", "conduction", "cropes", "breziest", "alisha", "mapmakers", "trojan"]     # "Pred. label" titles


def panels(path):
    """The two imshow panels of a figure, resampled to (32, 100) by reading the centre screen pixel of every image pixel."""
    im = np.array(Image.open(path).convert("RGB")).astype(int)
    nonwhite = im.sum(2) < 740
    rows = nonwhite.sum(1)
    r = [i for i in range(len(rows)) if rows[i] > 200]
    seg, s, p = [], r[0], r[0]
    for i in r[1:]:
        if i != p + 1:
            seg.append((s, p)); s = i
        p = i
    seg.append((s, p))
    out = []
    for y0, y1 in seg:
        cols = nonwhite[y0:y1 + 1].sum(0)
        c = [i for i in range(len(cols)) if cols[i] > (y1 - y0) * 0.8]
        iy0, iy1, ix0, ix1 = y0 + 1, y1 - 1, c[0] + 1, c[-1] - 1          # inside the 1-px axes frame
        H, W = iy1 - iy0 + 1, ix1 - ix0 + 1
        a = np.zeros((32, 100), np.uint8)
        for rr in range(32):
            for cc in range(100):
                a[rr, cc] = im[iy0 + int((rr + 0.5) * H / 32), ix0 + int((cc + 0.5) * W / 100), 0]
        out.append(a)
    assert len(out) == 2, path
    return out


def main():
    inp, stn = [], []
    for i in range(1, 8):
        a, b = panels(SRC % i)
        inp.append(a); stn.append(b)
    inp2, stn2 = [], []
    for i in range(1, 7):
        a, b = panels(SRC_IAM % i)
        inp2.append(a); stn2.append(b)
    np.savez_compressed(os.path.join(HERE, "reference_examples.npz"), input_panel=np.stack(inp), stn_panel=np.stack(stn),
                        true_label=np.array(TRUE), ref_pred=np.array(PRED),
                        iam_input_panel=np.stack(inp2), iam_stn_panel=np.stack(stn2), iam_true_label=np.array(TRUE_IAM), iam_ref_pred=np.array(PRED_IAM))
    # the IAM examples need the IAM-finetuned weights (models/OCR_IAM_ver1/final_weights.h5), read with our own HDF5 reader
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import crnn_b200 as cb
    w = cb.hdf5_lite.load_keras_weights("/root/reference/models/OCR_IAM_ver1/final_weights.h5")
    np.savez_compressed(os.path.join(HERE, "shipped_iam_weights.npz"), **{k.replace("/", "__"): v for k, v in w.items()})
    print("written reference_examples.npz", np.stack(inp).shape, np.stack(inp2).shape, "+ shipped_iam_weights.npz")


if __name__ == "__main__":
    main()
