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


def _pad_axis(img, axis, total, front, fill):
    """pad `front` fill-lines before and total-front after along `axis`."""
    shape_a = list(img.shape); shape_a[axis] = front
    shape_b = list(img.shape); shape_b[axis] = total - front
    return np.concatenate([np.full(shape_a, fill), img, np.full(shape_b, fill)], axis=axis)


# open_img in three stages so that a loader can run the expensive, deterministic parts (decode, up-scaling, threshold, final resize) on
# worker threads (OpenCV releases the GIL) while the middle stage -- the only one that consumes np.random -- runs on the caller's
# thread in the reference's order: the parallel loader is bit-identical to the sequential generator for any number of workers.
def _open_load(img, img_size):
    """utils.py:366-376: decode, make the line width axis 0, background value, 1.5x up-scaling of small crops.  Thread-safe."""
    name = None
    if isinstance(img, str):
        name = img
        img = read_img(name)
    img = img[::-1].T                                   # text-line width becomes axis 0
    fill = _modal_value(img)
    H, W = int(img_size[0]), int(img_size[1])
    if img.shape[0] <= H // 2 and img.shape[1] <= W // 2:
        img = cv2.resize(img, (int(img.shape[1] * 1.5), int(img.shape[0] * 1.5)))
    return img, fill, name


def _placement_draws(shape, img_size, p):
    """The random decisions of utils.py:378-401 for a crop of `shape` (after _open_load): consumes the global np.random stream exactly like
    the reference (one uniform per padded axis, one choice when the placement is random).  Returns (w_dec, h_dec), each None (no padding),
    ("rand", c) or ("default",)."""
    H, W = int(img_size[0]), int(img_size[1])
    decs = []
    for room, strict in ((W - shape[1], True), (H - shape[0], False)):
        if room > 2:
            r = round(np.random.uniform(0, 1), 1)
            if (r < p if strict else r <= p) and p > 0.:
                decs.append(("rand", int(np.random.choice(list(range(2, room))))))
            else:
                decs.append(("default",))
        else:
            decs.append(None)
    return decs[0], decs[1]


def _apply_placement(img, fill, img_size, w_dec, h_dec):
    H, W = int(img_size[0]), int(img_size[1])
    if w_dec is not None:
        room = W - img.shape[1]
        img = _pad_axis(img, 1, room - 1, w_dec[1] - 1, fill) if w_dec[0] == "rand" else _pad_axis(img, 1, room, 0, fill)
    if h_dec is not None:
        room = H - img.shape[0]
        img = _pad_axis(img, 0, room - 1, h_dec[1] - 1, fill) if h_dec[0] == "rand" else _pad_axis(img, 0, 2 * (room // 2), room // 2, fill)
    return img


def _open_place(img, fill, img_size, p):
    """utils.py:378-401: padding to the target size with the random placement (probability p).  Consumes the global np.random stream
    exactly like the reference -> must run sequentially, in image order."""
    w_dec, h_dec = _placement_draws(img.shape, img_size, p)
    return _apply_placement(img, fill, img_size, w_dec, h_dec)


def _file_image_size(name):
    """(width, height) as cv2.imread will decode the file, from its header: a few hundred bytes for PNG / baseline JPEG (the parent of the
    parallel loader does this for every image, PIL's generic open cost 100 us); anything unusual -- other formats, an EXIF block (cv2
    applies its orientation) -- goes through PIL."""
    import struct
    try:
        with open(name, "rb") as f:
            head = f.read(32)
            if head[:8] == b"\x89PNG\r\n\x1a\n" and head[12:16] == b"IHDR":
                w, h = struct.unpack(">II", head[16:24])
                return int(w), int(h)
            if head[:2] == b"\xff\xd8":
                f.seek(2)
                while True:
                    b = f.read(1)
                    if len(b) < 1 or b[0] != 0xFF:
                        break
                    m = f.read(1)
                    while m == b"\xff":                      # fill bytes before a marker
                        m = f.read(1)
                    if len(m) < 1:
                        break
                    if m[0] == 0x01 or 0xD0 <= m[0] <= 0xD7:   # stand-alone markers (TEM, RSTn) carry no length
                        continue
                    ln = f.read(2)
                    if len(ln) < 2:
                        break
                    marker, seglen = m[0], struct.unpack(">H", ln)[0]
                    if marker == 0xE1:                       # APP1 (EXIF): let PIL decide the orientation
                        break
                    if 0xC0 <= marker <= 0xCF and marker not in (0xC4, 0xC8, 0xCC):
                        d = f.read(5)
                        h, w = struct.unpack(">HH", d[1:5])
                        return int(w), int(h)
                    f.seek(seglen - 2, 1)
    except Exception:
        pass
    from PIL import Image
    with Image.open(name) as im:
        wf, hf = im.size
        try:
            if im.getexif().get(0x0112) in (5, 6, 7, 8):
                wf, hf = hf, wf
        except Exception:
            pass
    return int(wf), int(hf)


def _shape_after_load(name, img_size):
    """Shape `_open_load(name)` will return, from the file header only."""
    wf, hf = _file_image_size(name)
    s0, s1 = wf, hf                                      # img[::-1].T: (file width, file height)
    H, W = int(img_size[0]), int(img_size[1])
    if s0 <= H // 2 and s1 <= W // 2:
        s0, s1 = int(s0 * 1.5), int(s1 * 1.5)
    return s0, s1


def _open_with_draws(job):
    """Worker-process half of the parallel loader: the whole of open_img for one file with the random decisions made by the parent."""
    name, img_size, w_dec, h_dec, shape = job
    img, fill, _ = _open_load(name, img_size)
    if tuple(img.shape) != tuple(shape):
        return None                                      # header probe disagreed with the decoder: the parent redoes this file itself
    return _open_finish(_apply_placement(img, fill, img_size, w_dec, h_dec), img_size)


def _open_finish(img, img_size):
    """utils.py:403-407: invert dark-on-white crops, resize to the model's input size.  Thread-safe."""
    H, W = int(img_size[0]), int(img_size[1])
    binar = cv2.threshold(img, 255 // 2, 255, cv2.THRESH_BINARY)[1]
    if _modal_value(binar) == 255:                      # dark text on white -> invert
        img = cv2.bitwise_not(img)
    return cv2.resize(img, (W, H))


def _label_of(name):
    return os.path.basename(name).split("_")[1].lower() if name is not None else False


def open_img(img, img_size, p=.7):
    """utils.py:364-410.  Returns (uint8 image of shape img_size[:2] = (line width, line height), label-or-False).
    Note: the reference passes PIL's LANCZOS constant in cv2.resize's `dst` slot, so the effective interpolation is
    cv2's default INTER_LINEAR -- reproduced here."""
    img, fill, name = _open_load(img, img_size)
    img = _open_finish(_open_place(img, fill, img_size, p), img_size)
    return img, _label_of(name)


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
                 device_norm=False, workers=0):
        # device_norm (NEW): with normed=True, yield the 8-bit crops ('the_input' uint8) and let the model normalise them on the
        # GPU (crnn_normalize_u8, bit-identical to norm()): a quarter of the host->device bytes and no float64 batch on the host
        self.device_norm = bool(device_norm) and bool(normed)
        # workers (NEW): decode / pad / resize on a pool of worker processes, `workers * 16` images ahead; the random
        # placement decisions stay in this process, in the reference's order, so the batches are bit-identical to workers=0
        self.workers = int(workers)
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
        # The reference allocates X with np.empty (utils.py:448) and, for the partially filled last batch of the FIRST pass, hands Keras the rows it
        # never wrote -- uninitialised memory that train_on_batch then trains on.  Zeros here: with recycled heap memory those rows were
        # sometimes NaN / huge, which ReLU6 masks in the forward pass (finite loss) but not in the BatchNorm-1 / STN backward (found through
        # tests/test_gpu_dp.py::test_dp_train_cli_uneven_shards_and_early_stopping failing in ~1 of 5 runs; tools/dbg_dp_cli.py).  The valid
        # rows are unchanged (pinned on the reference's own generator in tests/test_host_logic.py).
        X = np.zeros((self.batch_size,) + tuple(self.img_size), np.uint8 if getattr(self, "device_norm", False) else np.float64)
        Y = np.full([self.batch_size, self.max_len], self.blank)
        return X, Y, np.ones((self.batch_size, 1)), np.zeros((self.batch_size, 1))

    def _crops(self, names):
        """(crop, word) of every file in `names`, in order, decoded on `self.workers` worker PROCESSES (the per-image work is dominated by
        Python / numpy overhead that holds the GIL, threads gave no speed-up).  The parent only reads each file's header, makes the random
        placement decisions in the reference's order (np.random is consumed exactly as by the sequential generator) and ships
        (file, decisions) to the pool, so the batches are bit-identical to workers=0."""
        import multiprocessing as mp
        if getattr(self, "_pool", None) is None:
            import atexit
            import warnings
            with warnings.catch_warnings():                             # CPython 3.12 warns about fork() in a multi-threaded process
                warnings.simplefilter("ignore", DeprecationWarning)
                self._pool = mp.get_context("fork").Pool(self.workers)  # fork, like torch's DataLoader: the workers never touch CUDA
            atexit.register(self.close)
        window = max(self.workers * 16, 1)

        def submit(chunk):
            jobs = []
            for n in chunk:
                shape = _shape_after_load(n, self.img_size)
                w_dec, h_dec = _placement_draws(shape, self.img_size, self.transform_p)
                jobs.append((n, self.img_size, w_dec, h_dec, shape))
            return jobs, self._pool.map_async(_open_with_draws, jobs, chunksize=max(1, len(jobs) // (self.workers * 4)))
        pending = submit(names[:window]) if names else None
        for lo in range(0, len(names), window):
            jobs, res = pending
            hi = lo + window
            pending = submit(names[hi:hi + window]) if hi < len(names) else None      # next window decodes while this one is consumed
            for job, crop in zip(jobs, res.get()):
                if crop is None:        # header and decoder disagreed on the size (never seen on mjsynth / IAM): do it here with the same decisions
                    img, fill, _ = _open_load(job[0], self.img_size)
                    # the decisions were drawn for a wrong shape: default placement, built WITHOUT touching np.random (the stream must stay
                    # exactly where the sequential generator would have it, or every later crop differs -- ADVICE r1)
                    H_, W_ = int(self.img_size[0]), int(self.img_size[1])
                    w_dec = ("default",) if W_ - img.shape[1] > 2 else None
                    h_dec = ("default",) if H_ - img.shape[0] > 2 else None
                    crop = _open_finish(_apply_placement(img, fill, self.img_size, w_dec, h_dec), self.img_size)
                yield crop, _label_of(job[0])

    def close(self):
        """Stop the worker processes of the parallel loader (idempotent)."""
        pool, self._pool = getattr(self, "_pool", None), None
        if pool is not None:
            pool.terminate()
            pool.join()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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
            whole_crops = self._crops([n for n in names if bboxs[n][0] == n]) if self.workers > 1 else None
            for name in names:
                whole = bboxs[name][0] == name
                if whole:
                    crop, word = next(whole_crops) if whole_crops is not None else open_img(name, self.img_size, p=self.transform_p)
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
                    if (done == full and i == rem) or i == self.batch_size:
                        # (the reference rebuilds this dict, incl. np.array(source_str), after EVERY image; only the yielded ones are observable)
                        batch = ({"the_input": X, "the_labels": Y, "input_length": il, "label_length": ll, "source_str": np.array(words)},
                                 {"ctc": np.zeros([self.batch_size])})
                        if done == full and i == rem:
                            yield batch            # last, partially filled batch of the FIRST pass
                        else:
                            done += 1; i = 0; words = []
                            X, Y, il, ll = self.get_blank_matrices()
                            yield batch
