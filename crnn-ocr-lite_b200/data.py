"""Host-side input pipeline with the behaviour of the reference's generator (gasparian/CRNN-OCR-lite utils.py:359-528):
image loading / padding / inversion / resize (`open_img`), normalisation, label encoding and the batch generator that
feeds `fit_generator` / `predict_generator`.  CPU code (numpy + OpenCV); it is NOT on the accelerated path
(SURVEY.md 8f-2 ranks it "next").  Same names, arguments and dict keys as the reference."""
import os
import string

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

MJ_MEAN, MJ_STD = 118.24236953981779, 36.72835353999682   # utils.py:421


def get_lexicon(non_intersecting_chars=False):
    base = string.digits + string.ascii_lowercase
    if non_intersecting_chars:
        return list(set(base + "AaBbDdEeFfGgHhLlMmNnQqRrTt" + "-"))
    return list(base + "-")


def read_img(name):
    return cv2.cvtColor(np.asarray(cv2.imread(name), dtype=np.uint8), cv2.COLOR_BGR2GRAY)


def _modal_value(a):
    vals, counts = np.unique(a, return_counts=True)
    return vals[np.flatnonzero(counts == counts.max())[0]]


def _pad_axis(img, axis, total, front):
    """pad `front` fill-lines before and total-front after along `axis`."""
    fill = _pad_axis.fill
    shape_a = list(img.shape); shape_a[axis] = front
    shape_b = list(img.shape); shape_b[axis] = total - front
    return np.concatenate([np.full(shape_a, fill), img, np.full(shape_b, fill)], axis=axis)


def open_img(img, img_size, p=.7):
    """utils.py:364-410.  Returns (uint8 image of shape img_size[:2] = (line width, line height), label-or-False).
    Note: the reference passes PIL's LANCZOS constant in cv2.resize's `dst` slot, so the effective interpolation is
    cv2's default INTER_LINEAR -- reproduced here."""
    name = None
    if isinstance(img, str):
        name = img
        img = read_img(name)
    img = img[::-1].T                                   # text-line width becomes axis 0
    fill = _modal_value(img)
    H, W = int(img_size[0]), int(img_size[1])
    if img.shape[0] <= H // 2 and img.shape[1] <= W // 2:
        img = cv2.resize(img, (int(img.shape[1] * 1.5), int(img.shape[0] * 1.5)))
    _pad_axis.fill = fill
    room = W - img.shape[1]
    if room > 2:
        r = round(np.random.uniform(0, 1), 1)
        if r < p and p > 0.:                            # random horizontal placement
            c = np.random.choice(list(range(2, room)))
            img = _pad_axis(img, 1, room - 1, c - 1)
        else:
            img = _pad_axis(img, 1, room, 0)
    room = H - img.shape[0]
    if room > 2:
        r = round(np.random.uniform(0, 1), 1)
        if r <= p and p > 0.:
            c = np.random.choice(list(range(2, room)))
            img = _pad_axis(img, 0, room - 1, c - 1)
        else:
            img = _pad_axis(img, 0, 2 * (room // 2), room // 2)
    binar = cv2.threshold(img, 255 // 2, 255, cv2.THRESH_BINARY)[1]
    if _modal_value(binar) == 255:                      # dark text on white -> invert
        img = cv2.bitwise_not(img)
    img = cv2.resize(img, (W, H))
    if name is not None:
        return img, os.path.basename(name).split("_")[1].lower()
    return img, False


def parse_mjsynth(path, names):
    return [os.path.join(path, line.split()[0][2:]) for line in names]


def norm(image, mean, std):
    return (image.astype("float32") - mean) / std


def make_ohe(y, nclasses):
    out = np.zeros((len(y), nclasses))
    out[np.arange(len(y)), np.asarray(y).astype("int64")] = 1
    return out


def get_lengths(names):
    return {n: len(os.path.basename(n).split("_")[1]) for n in names}


class Readf:
    """utils.py:418-511: batch generator yielding ({'the_input','the_labels','input_length','label_length',
    'source_str'}, {'ctc'}) with X float64 (B,H,W,1) and labels padded with blank=len(classes)."""

    def __init__(self, img_size=(40, 40), max_len=30, normed=False, batch_size=32, classes={}, mean=MJ_MEAN, std=MJ_STD, transform_p=0.7,
                 device_norm=False):
        # device_norm (NEW): with normed=True, yield the 8-bit crops ('the_input' uint8) and let the model normalise them on the
        # GPU (crnn_normalize_u8, bit-identical to norm()): a quarter of the host->device bytes and no float64 batch on the host
        self.device_norm = bool(device_norm) and bool(normed)
        self.batch_size, self.transform_p, self.img_size, self.normed = batch_size, transform_p, img_size, normed
        self.classes, self.max_len, self.mean, self.std = classes, max_len, mean, std
        self.voc = list(classes.keys())
        if isinstance(classes, dict):
            self.blank = len(classes)

    def make_target(self, text):
        dash = self.classes["-"]
        return np.array([self.classes.get(ch, dash) for ch in text])

    def get_labels(self, names):
        Y = np.full([len(names), self.max_len], self.blank)
        for i, name in enumerate(names):
            _, word = open_img(name, self.img_size, p=self.transform_p)
            t = self.make_target(word)
            Y[i, :len(t)] = t
        return Y

    def get_blank_matrices(self):
        X = np.empty((self.batch_size,) + tuple(self.img_size), np.uint8 if getattr(self, "device_norm", False) else np.float64)
        Y = np.full([self.batch_size, self.max_len], self.blank)
        return X, Y, np.ones((self.batch_size, 1)), np.zeros((self.batch_size, 1))

    def run_generator(self, names, downsample_factor=2, bboxs={}):
        if bboxs:
            n_items = sum(len(v) for v in bboxs.values())
        else:
            bboxs = {name: [name] for name in names}
            n_items = len(names)
        full, rem = divmod(n_items, self.batch_size)
        in_len = (self.img_size[0] + 4) // downsample_factor - 2          # utils.py:487
        # State machine of utils.py:462-511, kept as is: the counters and buffers are NOT reset between passes over `names`, so only the
        # first pass ends with a partial batch (tail rows stale, trimmed by the caller); from then on the stream is full batches that
        # wrap around the end of the list (the first one after the wrap starts with the `rem` samples of the partial batch again).
        # Pinned against the reference's own class in tests/test_host_logic.py::test_readf_generator_reproduces_reference.
        done, i, words = 0, 0, []
        X, Y, il, ll = self.get_blank_matrices()
        while True:
            for name in names:
                whole = bboxs[name][0] == name
                if whole:
                    crop, word = open_img(name, self.img_size, p=self.transform_p)
                else:
                    page = read_img(name)
                for box in bboxs[name]:
                    if not whole:
                        crop, _ = open_img(page[box[1]:box[3], box[2]:box[4]], self.img_size, p=self.transform_p)
                        word = box[0] if box[0] is not None else "-"
                    words.append(word)
                    t = self.make_target(word)
                    Y[i, :len(t)] = t
                    ll[i] = len(t)
                    il[i] = in_len
                    X[i] = (norm(crop, self.mean, self.std) if (self.normed and not self.device_norm) else crop)[:, :, np.newaxis]
                    i += 1
                    batch = ({"the_input": X, "the_labels": Y, "input_length": il, "label_length": ll, "source_str": np.array(words)},
                             {"ctc": np.zeros([self.batch_size])})
                    if done == full and i == rem:
                        yield batch            # last, partially filled batch of the FIRST pass
                    elif i == self.batch_size:
                        done += 1; i = 0; words = []
                        X, Y, il, ll = self.get_blank_matrices()
                        yield batch
