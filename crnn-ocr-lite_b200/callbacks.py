"""Training-driver policy of the reference: EarlyStoppingIter (utils.py:535-614) and the ModelCheckpoint it uses
(train.py:194-195).  Plain Python objects driven by CRNNModel.fit_generator."""
import warnings

import numpy as np


class Callback:
    model = None

    def on_train_begin(self, logs=None): pass
    def on_train_end(self, logs=None): pass
    def on_batch_end(self, batch, logs=None): pass
    def on_epoch_end(self, epoch, logs=None): pass


class EarlyStoppingIter(Callback):
    """Every `patience` iterations compare the cumulative mean of `monitor` with the best one seen; stop when it
    did not improve by min_delta (utils.py:587-610).  np.Inf / missing `warnings` of the reference are fixed."""

    def __init__(self, monitor="loss", min_delta=0, patience=5000, verbose=0, mode="auto", baseline=None, restore_best_weights=False):
        self.monitor, self.baseline, self.patience, self.verbose = monitor, baseline, patience, verbose
        self.restore_best_weights = restore_best_weights
        self.stopped_iter = self.cycle_iterations = 0
        self.sum_monitor = 0
        self.best_weights = None
        if mode not in ("auto", "min", "max"):
            warnings.warn("EarlyStopping mode %s is unknown, fallback to auto mode." % mode, RuntimeWarning)
            mode = "auto"
        maximise = mode == "max" or (mode == "auto" and "acc" in monitor)
        self.monitor_op = np.greater if maximise else np.less
        self.min_delta = min_delta if maximise else -min_delta

    def on_train_begin(self, logs=None):
        self.stopped_iter = 0
        self.best = self.baseline if self.baseline is not None else (np.inf if self.monitor_op == np.less else -np.inf)

    def on_batch_end(self, batch, logs=None):
        self.cycle_iterations += 1
        logs = logs or {}
        if self.monitor not in logs:
            return
        self.sum_monitor += logs[self.monitor]
        if (self.cycle_iterations - 1) % self.patience:
            return
        current = self.sum_monitor / self.cycle_iterations
        if self.monitor_op(current - self.min_delta, self.best):
            self.best = current
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
        else:
            self.stopped_iter = self.cycle_iterations
            self.model.stop_training = True
            if self.restore_best_weights and self.best_weights is not None:
                if self.verbose > 0:
                    print("\nRestoring model weights from the end of the best epoch")
                self.model.set_weights(self.best_weights)

    def on_train_end(self, logs=None):
        if self.stopped_iter > 0 and self.verbose > 0:
            print("\nIteration %i: early stopping\nBest metric value: %.4f" % (self.stopped_iter + 1, self.best))


class ModelCheckpoint(Callback):
    """keras.callbacks.ModelCheckpoint(filepath, save_best_only=True, save_weights_only=True) on val_loss."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False, **_):
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.save_weights_only = save_best_only, save_weights_only
        self.best = np.inf

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        cur = logs.get(self.monitor)
        if self.save_best_only:
            if cur is None:
                warnings.warn("Can save best model only with %s available, skipping." % self.monitor, RuntimeWarning)
                return
            if not cur < self.best:
                return
            if self.verbose:
                print("\nEpoch %05d: %s improved from %0.5f to %0.5f, saving model to %s" % (epoch + 1, self.monitor, self.best, cur, self.filepath))
            self.best = cur
        (self.model.save_weights if self.save_weights_only else self.model.save)(self.filepath)
