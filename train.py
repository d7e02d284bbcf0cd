#!/usr/bin/env python
"""train.py -- same command line as the reference's train.py (gasparian/CRNN-OCR-lite train.py:83-106), running on
the B200-native engine.  Files written under <save_path>/<model_name>/ match the reference: arguments.txt, model.json,
model_summary.txt, checkpoint_weights.h5 (best val_loss), loss_history.pickle.dat, final_weights.h5, final_model.h5.

Deliberate, documented differences:
  * the reference's `--GRU` / `--norm` flags are dead (shadowed by `from utils import *`): every model the
    CLI ever produced is a normalised Bi-GRU.  That behaviour is kept as the default; `--cell lstm` (new) selects the
    LSTM graph of utils.py:78-79.
  * `--G` selects the CUDA device; launching under torchrun adds data-parallel training (new capability).
"""
import argparse
import os
import pickle
import re
import time
from shutil import rmtree

import numpy as np
from numpy.random import RandomState


# The reference's command line (train.py:83-106), one row per flag: (names, type or None for a store_true switch, default, required).
_FLAGS = [
    (("-p", "--path"), str, None, True), (("--training_fname",), str, None, False), (("--val_fname",), str, "", False),
    (("--save_path",), str, None, True), (("--model_name",), str, None, True), (("--pretrained_path",), str, None, False),
    (("--nbepochs",), int, 20, False), (("--G",), str, "1", False), (("--random_state",), int, 42, False),
    (("--train_portion",), float, 0.9, False), (("--time_dense_size",), int, 128, False), (("--n_units",), int, 256, False),
    (("--batch_size",), int, 64, False), (("--opt",), str, "sgd", False), (("--lr",), float, 0.001, False),
    (("--early_stopping",), int, 0, False), (("--norm",), None, False, False), (("--mjsynth",), None, False, False),
    (("--GRU",), None, False, False), (("--imgh",), int, 100, False), (("--imgW",), int, 32, False),
]


def build_parser():
    parser = argparse.ArgumentParser(description="crnn_ctc_loss")
    for names, typ, default, required in _FLAGS:
        if typ is None:
            parser.add_argument(*names, action="store_true")
        elif required:
            parser.add_argument(*names, type=typ, required=True)
        else:
            parser.add_argument(*names, type=typ, default=default)
    parser.add_argument("--cell", choices=["gru", "lstm"], default="gru", help="extension: recurrent cell (reference CLI always builds GRU)")
    return parser


def main():
    parser = build_parser()
    args = parser.parse_args()

    if "LOCAL_RANK" not in os.environ:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.G)
    import torch
    import utils as U
    import crnn_b200 as cb
    local = int(os.environ.get("LOCAL_RANK", "0")) % max(1, torch.cuda.device_count())   # several ranks may share a GPU (gloo test mode)
    torch.cuda.set_device(local)
    cb.parallel.init_distributed(device=torch.device("cuda", local))
    rank, world = cb.parallel.rank(), cb.parallel.world_size()

    out_dir = os.path.join(args.save_path, args.model_name)
    if rank == 0:
        rmtree(out_dir, ignore_errors=True)
        os.makedirs(out_dir)
        with open(os.path.join(out_dir, "arguments.txt"), "w") as f:
            f.write(str(args))
    prng = RandomState(args.random_state)
    lexicon = U.get_lexicon()
    classes = {ch: i for i, ch in enumerate(lexicon)}
    print(" [INFO] %s" % classes)

    if args.mjsynth:
        train = U.parse_mjsynth(args.path, open(os.path.join(args.path, args.training_fname)).readlines())
        prng.shuffle(train)
        val = U.parse_mjsynth(args.path, open(os.path.join(args.path, args.val_fname)).readlines())
    else:
        train = [os.path.join(dp, f) for dp, _dn, fs in os.walk(args.path) for f in fs if re.search("png|jpeg|jpg", f)]
        prng.shuffle(train)
        cut = int(len(train) * args.train_portion)
        train, val = train[:cut], train[cut:]
    max_len = max(U.get_lengths(train).values())
    n_train_global = len(train)
    if world > 1:                                   # text lines are independent: shard the file list, no data-path collective
        lo, hi = cb.parallel.shard_batch(len(train))
        train = train[lo:hi]
    print(f" [INFO] {len(train)} train and {len(val)} validation images loaded ")

    reader = U.Readf(img_size=(args.imgh, args.imgW, 1), normed=True, batch_size=args.batch_size, classes=classes, max_len=max_len, transform_p=0.7)
    print(" [INFO] Number of classes: {}; Max. string length: {} ".format(len(classes) + 1, max_len))
    init_model = U.CRNN(num_classes=len(classes) + 1, shape=(args.imgh, args.imgW, 1), GRU=(args.cell == "gru"),
                        time_dense_size=args.time_dense_size, n_units=args.n_units, max_string_len=max_len, max_batch=args.batch_size)
    model = init_model.get_model()
    if rank == 0:
        U.save_model_json(model, args.save_path, args.model_name)
    if args.pretrained_path is not None:
        model.load_weights(args.pretrained_path)
    cb.parallel.broadcast_(model.tensor("arena/params"))
    model.enable_native_dp()                        # NCCL process group: the step reduces its gradients itself (bucketed, overlapped); no-op otherwise

    # identical on every rank (from the largest shard); the generator wraps around its list, so a shorter shard just re-uses its first files
    train_steps = cb.parallel.steps_per_epoch(n_train_global, args.batch_size, world)
    test_steps = -(-len(val) // args.batch_size)
    start_time = time.time()
    if rank == 0:
        with open(os.path.join(out_dir, "model_summary.txt"), "w") as f:
            model.summary(print_fn=lambda x: f.write(x + "\n"))
        model.summary()
    if args.opt == "adam":
        optimizer = U.optimizers.Adam(lr=args.lr, beta_1=0.5, beta_2=0.999, clipnorm=5)
    else:
        optimizer = U.optimizers.SGD(lr=args.lr, decay=1e-6, momentum=0.9, nesterov=True, clipnorm=5)
    model.compile(loss={"ctc": lambda y_true, y_pred: y_pred}, optimizer=optimizer)
    callbacks = []
    if rank == 0:
        callbacks.append(U.ModelCheckpoint(filepath=os.path.join(out_dir, "checkpoint_weights.h5"), verbose=1, save_best_only=True, save_weights_only=True))
    if args.early_stopping:
        callbacks.append(U.EarlyStoppingIter(monitor="loss", min_delta=.0001, patience=args.early_stopping, verbose=1, restore_best_weights=True, mode="auto"))
    ds = 2 ** init_model.pooling_counter_h
    H = model.fit_generator(generator=reader.run_generator(train, downsample_factor=ds), steps_per_epoch=train_steps, epochs=args.nbepochs,
                            validation_data=reader.run_generator(val, downsample_factor=ds) if test_steps else None, validation_steps=test_steps,
                            shuffle=False, verbose=1 if rank == 0 else 0, callbacks=callbacks)
    if os.environ.get("CRNN_DP_DUMP_PARAMS"):          # test hook: every rank dumps its final parameters (replicas must be identical)
        np.save("%s.rank%d.npy" % (os.environ["CRNN_DP_DUMP_PARAMS"], rank), model.tensor("arena/params").cpu().numpy())
    if rank == 0:
        pickle.dump(H.history, open(os.path.join(out_dir, "loss_history.pickle.dat"), "wb"))
        print(" [INFO] Training finished in %i sec.!" % (round(time.time() - start_time, 2)))
        model.save_weights(os.path.join(out_dir, "final_weights.h5"))
        model.save(os.path.join(out_dir, "final_model.h5"))
        print(" [INFO] Models and history saved! ")


if __name__ == "__main__":
    main()
