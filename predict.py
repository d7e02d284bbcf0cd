#!/usr/bin/env python
"""predict.py -- same command line as the reference's predict.py (gasparian/CRNN-OCR-lite predict.py:62-82) on the
B200-native engine: forward pass of `models/<name>/{model.json,final_weights.h5}` + batched GPU beam-10 decode
(`DecodeCTCPred`).  `--greedy` (new) switches to the greedy decoder of BASELINE configs[0,1].  There is no CPU mode:
`--G -1` (the reference's CPU-only switch) is mapped to device 0 and reported."""
import argparse
import os
import pickle
import re
import time

import numpy as np
from numpy.random import RandomState


# The reference's command line (predict.py:62-82), one row per flag: (name, type or None for a store_true switch, default, required).
_FLAGS = [
    ("--model_path", str, None, True), ("--image_path", str, None, True), ("--result_path", str, None, False), ("--max_len", int, 23, False),
    ("--boxes", str, None, False), ("--val_fname", str, None, False), ("--num_instances", int, None, False), ("--G", int, -1, False),
    ("--batch_size", int, 64, False), ("--random_state", int, 42, False), ("--train_portion", float, .9, False),
    ("--validate", None, False, False), ("--mjsynth", None, False, False), ("--imgh", int, 100, False), ("--imgW", int, 32, False),
]


def build_parser():
    parser = argparse.ArgumentParser()
    for name, typ, default, required in _FLAGS:
        if typ is None:
            parser.add_argument(name, action="store_true")
        elif required:
            parser.add_argument(name, type=typ, required=True)
        else:
            parser.add_argument(name, type=typ, default=default)
    parser.add_argument("--greedy", action="store_true", help="extension: greedy CTC decode instead of beam width 10")
    return parser


def main():
    parser = build_parser()
    args = parser.parse_args()

    import torch
    import utils as U
    if args.G < 0:
        print(" [INFO] --G -1 (CPU-only in the reference) is not available on this engine: using cuda:0")
    torch.cuda.set_device(max(args.G, 0))

    prng = RandomState(args.random_state)
    model = U.init_predictor(U.load_custom_model(args.model_path, model_name="/model.json", weights="/final_weights.h5", max_batch=args.batch_size))
    classes = {ch: i for i, ch in enumerate(U.get_lexicon())}
    inverse_classes = {v: k for k, v in classes.items()}
    decoder = U.DecodeCTCPred(top_paths=1, beam_width=10, inverse_classes=inverse_classes, greedy=args.greedy)
    img_size = (model.imgh, model.imgw, 1)           # fixed by model.json, as in Keras

    def walk():
        return np.array([os.path.join(dp, f) for dp, _dn, fs in os.walk(args.image_path) for f in fs if re.search("png|jpeg|jpg", f)])
    if args.validate and args.mjsynth:
        fnames = np.array(U.parse_mjsynth(args.image_path, open(os.path.join(args.image_path, args.val_fname)).readlines()))
    elif args.validate:
        fnames = walk()
        prng.shuffle(fnames)
        fnames = fnames[int(len(fnames) * args.train_portion):]
    else:
        fnames = walk()
    if args.num_instances is not None:
        fnames = fnames[np.random.randint(0, len(fnames), min(args.num_instances, len(fnames)))]

    reader = U.Readf(img_size=img_size, normed=True, batch_size=args.batch_size, transform_p=0., classes=classes, max_len=args.max_len)
    length = len(fnames)
    bboxs = {}
    y_true = None
    if args.boxes is not None:
        bboxs = pickle.load(open(args.boxes, "rb"))     # {image: [(word|None, x0, y0, x1, y1), ...]}
        part = len(bboxs) // 2
        bboxs = {os.path.join(args.image_path, k): v for i, (k, v) in enumerate(bboxs.items()) if i <= part}
        length = sum(len(v) for v in bboxs.values())
        fnames = list(bboxs.keys())
        if args.validate:
            y_true = [reader.make_target(el[0]) for v in bboxs.values() for el in v]
    else:
        y_true = reader.get_labels(fnames)
    steps = -(-length // args.batch_size)

    print(" [INFO] Predicting... ")
    start = time.time()
    predicted = model.predict_generator(reader.run_generator(fnames, bboxs=bboxs, downsample_factor=2), steps=steps)
    print(f" [INFO] {len(fnames)} images processed in {round(time.time() - start, 2)} sec. ")
    start = time.time()
    predicted_text = decoder.decode(predicted)[:length]
    print(f" [INFO] {len(predicted)} predictions decoded in {round(time.time() - start, 2)} sec. ")

    if args.result_path is not None:
        import pandas as pd
        if len(fnames) != len(predicted_text):
            fnames = [f for f in bboxs for _ in range(len(bboxs[f]))]
        out_name = os.path.join(args.result_path, "prediction.csv")
        pd.DataFrame({"fname": fnames, "prediction": predicted_text}).to_csv(out_name)
        print(" [INFO] Prediction example: \n", predicted_text[:10])
        print(" [INFO] Result store in: ", out_name)
    if args.validate:
        print(" [INFO] Computing edit distance metric... ")
        start = time.time()
        true_text = [decoder.labels_to_text(y) for y in y_true]
        print(" [INFO] Example pairs (predicted, true): \n", list(zip(predicted_text[:10], true_text[:10])))
        ed = U.edit_distance_cuda(predicted_text, true_text)                 # utils.py:289-299 on the GPU (bit-identical distances)
        ned = U.normalized_edit_distance_cuda(predicted_text, true_text)
        print(f" [INFO] edit distances calculated in {round(time.time() - start, 2)} sec. ")
        print(f" [INFO] mean edit distance: {ed} ")
        print(f" [INFO] mean normalized edit distance: {ned} ")


if __name__ == "__main__":
    main()
