"""Dependency-free reader/writer for the Keras-2.2.2 HDF5 weight files of the reference.

The reference loads/saves weights through Keras (`model.load_weights`, `model.save_weights`,
reference utils.py:300-329, train.py:170-171,215-216) which needs h5py/libhdf5 -- neither
exists in this image.  The files under `models/<name>/*.h5` are plain "old style" HDF5:
superblock v0, symbol-table groups (TREE/HEAP/SNOD), version-1 object headers, contiguous
little-endian fp32 datasets, fixed-length string attributes.  That
subset is small enough to parse (and emit) directly.

Public API
    read_h5(path)            -> H5Group tree (groups: .attrs/.children, datasets: numpy arrays)
    load_keras_weights(path) -> OrderedDict "layer/weight" -> np.ndarray   (e.g. "conv2d_3/kernel")
    save_keras_weights(path, layers) -> writes a Keras-2.2.2 style weight file readable by read_h5
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Group:
    def __init__(self, name):
        self.name = name
        self.attrs = OrderedDict()
        self.children = OrderedDict()  # name -> H5Group | np.ndarray

    def __getitem__(self, key):
        node = self
        for part in key.strip("/").split("/"):
            node = node.children[part]
        return node

    def walk(self, prefix=""):
        for k, v in self.children.items():
            p = f"{prefix}/{k}" if prefix else k
            if isinstance(v, H5Group):
                yield from v.walk(p)
            else:
                yield p, v


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        ver = buf[8]
        if ver != 0:
            raise ValueError(f"unsupported HDF5 superblock version {ver} (only v0, as written by Keras 2.2.2/h5py 2.x)")
        self.so, self.sl = buf[13], buf[14]
        if (self.so, self.sl) != (8, 8):
            raise ValueError("unsupported offset/length size")
        # sig8 ver1 fs1 root1 rsv1 shm1 so1 sl1 rsv1 leafk2 intk2 flags4 | base8 free8 eof8 drv8 | root entry
        self.base = struct.unpack_from("<Q", buf, 24)[0]
        self.root_entry = 24 + 32
        self.gcol = {}

    # -- primitives ------------------------------------------------------------------
    def u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def sym_entry(self, off):
        name_off, ohdr, cache = struct.unpack_from("<QQI", self.b, off)
        scratch = self.b[off + 24:off + 40]
        return name_off, ohdr, cache, scratch

    # -- object header v1 ------------------------------------------------------------
    def messages(self, addr):
        b = self.b
        ver, _, nmsgs, _refcnt, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise ValueError(f"unsupported object header version {ver} at {addr}")
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsgs:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsgs:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x0010:  # continuation
                    coff, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((coff + self.base, clen))
                out.append((mtype, body, msize, mflags))
                p = body + msize
        return out

    # -- datatype / dataspace --------------------------------------------------------
    def parse_datatype(self, off):
        b = self.b
        cv = b[off]
        cls, ver = cv & 0x0F, cv >> 4
        bits = b[off + 1:off + 4]
        size = struct.unpack_from("<I", b, off + 4)[0]
        if cls == 0:  # fixed point
            signed = bool(bits[0] & 0x08)
            return {"cls": "int", "size": size, "signed": signed, "len": 8 + 4}
        if cls == 1:
            return {"cls": "float", "size": size, "len": 8 + 12}
        if cls == 3:
            return {"cls": "string", "size": size, "len": 8}
        if cls == 9:
            base = self.parse_datatype(off + 8)
            vtype = bits[0] & 0x0F  # 0 sequence, 1 string
            return {"cls": "vlen", "size": size, "base": base, "vstr": vtype == 1, "len": 8 + base["len"]}
        raise ValueError(f"unsupported datatype class {cls} (v{ver})")

    def parse_dataspace(self, off):
        b = self.b
        ver, rank, flags = b[off], b[off + 1], b[off + 2]
        if ver == 1:
            p = off + 8
        elif ver == 2:
            p = off + 4
        else:
            raise ValueError(f"dataspace version {ver}")
        dims = [struct.unpack_from("<Q", b, p + 8 * i)[0] for i in range(rank)]
        return dims

    def np_dtype(self, dt):
        if dt["cls"] == "float":
            return np.dtype("<f%d" % dt["size"])
        if dt["cls"] == "int":
            return np.dtype("<%s%d" % ("i" if dt["signed"] else "u", dt["size"]))
        if dt["cls"] == "string":
            return np.dtype("S%d" % dt["size"])
        raise ValueError("no numpy dtype for " + dt["cls"])

    def global_heap_obj(self, addr, index):
        b = self.b
        a = addr + self.base
        if b[a:a + 4] != b"GCOL":
            raise ValueError("bad global heap")
        csize = struct.unpack_from("<Q", b, a + 8)[0]
        p = a + 16
        while p < a + csize:
            idx, _ref, _r, osize = struct.unpack_from("<HHIQ", b, p)
            if idx == index:
                return b[p + 16:p + 16 + osize]
            if idx == 0:
                break
            p += 16 + ((osize + 7) & ~7)
        raise KeyError("global heap object")

    def read_values(self, dt, dims, raw_off):
        n = int(np.prod(dims)) if dims else 1
        if dt["cls"] == "vlen":
            vals = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", self.b, raw_off + 16 * i)
                vals.append(self.global_heap_obj(gaddr, gidx)[:ln])
            if dt["vstr"]:
                vals = [v.decode("utf8") for v in vals]
            return vals[0] if not dims else vals
        npdt = self.np_dtype(dt)
        arr = np.frombuffer(self.b, dtype=npdt, count=n, offset=raw_off).reshape(dims) if dims else \
            np.frombuffer(self.b, dtype=npdt, count=1, offset=raw_off)[0]
        return arr

    def parse_attribute(self, off):
        b = self.b
        ver = b[off]
        nsz, dtsz, dssz = struct.unpack_from("<HHH", b, off + 2)
        if ver == 1:
            pad = lambda x: (x + 7) & ~7
            p = off + 8
        elif ver in (2, 3):
            pad = lambda x: x
            p = off + 8 + (1 if ver == 3 else 0)
        else:
            raise ValueError(f"attribute version {ver}")
        name = b[p:p + nsz].split(b"\0")[0].decode()
        p += pad(nsz)
        dt = self.parse_datatype(p)
        p += pad(dtsz)
        dims = self.parse_dataspace(p)
        p += pad(dssz)
        val = self.read_values(dt, dims, p)
        if isinstance(val, np.ndarray) and val.dtype.kind == "S":
            val = [s.decode() for s in val.tolist()] if val.ndim else val.tobytes().decode()
        elif isinstance(val, (bytes, np.bytes_)):
            val = bytes(val).split(b"\0")[0].decode()
        return name, val

    # -- groups ----------------------------------------------------------------------
    def heap_data_addr(self, heap_addr):
        a = heap_addr + self.base
        if self.b[a:a + 4] != b"HEAP":
            raise ValueError("bad local heap")
        return struct.unpack_from("<Q", self.b, a + 24)[0] + self.base

    def btree_entries(self, btree_addr, heap_data):
        b = self.b
        a = btree_addr + self.base
        if b[a:a + 4] != b"TREE":
            raise ValueError("bad btree node")
        ntype, level, nused = struct.unpack_from("<BBH", b, a + 4)
        p = a + 8 + 16  # skip siblings
        out = []
        for i in range(nused):
            child = struct.unpack_from("<Q", b, p + 8 + 16 * i)[0]
            if level > 0:
                out += self.btree_entries(child, heap_data)
            else:
                s = child + self.base
                if b[s:s + 4] != b"SNOD":
                    raise ValueError("bad SNOD")
                nsym = struct.unpack_from("<H", b, s + 6)[0]
                for j in range(nsym):
                    name_off, ohdr, cache, scratch = self.sym_entry(s + 8 + 40 * j)
                    e = heap_data + name_off
                    name = b[e:b.index(b"\0", e)].decode()
                    out.append((name, ohdr + self.base))
        return out

    def read_object(self, name, ohdr):
        msgs = self.messages(ohdr)
        types = {m[0] for m in msgs}
        if 0x0011 in types:  # group
            g = H5Group(name)
            for mtype, body, msize, _ in msgs:
                if mtype == 0x000C:
                    k, v = self.parse_attribute(body)
                    g.attrs[k] = v
                elif mtype == 0x0011:
                    bt, hp = struct.unpack_from("<QQ", self.b, body)
                    for cname, caddr in self.btree_entries(bt, self.heap_data_addr(hp)):
                        g.children[cname] = self.read_object(cname, caddr)
            return g
        dt = dims = None
        data_addr = data_size = None
        for mtype, body, msize, _ in msgs:
            if mtype == 0x0001:
                dims = self.parse_dataspace(body)
            elif mtype == 0x0003:
                dt = self.parse_datatype(body)
            elif mtype == 0x0008:
                ver, cls = self.b[body], self.b[body + 1]
                if ver != 3 or cls != 1:
                    raise ValueError(f"dataset {name}: only contiguous layout v3 supported (got v{ver} class {cls})")
                data_addr, data_size = struct.unpack_from("<QQ", self.b, body + 2)
            elif mtype == 0x000B:
                raise ValueError(f"dataset {name}: filter pipelines (compression) not supported")
        if dt is None or dims is None:
            raise ValueError(f"object {name}: neither group nor dataset")
        if data_addr is None or data_addr == _UNDEF:
            return np.zeros(dims, dtype=self.np_dtype(dt))
        return np.array(self.read_values(dt, dims, data_addr + self.base))

    def root(self):
        _, ohdr, _, _ = self.sym_entry(self.root_entry)
        return self.read_object("/", ohdr + self.base)


def read_h5(path) -> H5Group:
    with open(path, "rb") as f:
        return _Reader(f.read()).root()


def _short(layer, wname):
    """'conv2d_3/kernel:0' -> 'kernel'; 'bidirectional_1/forward_gru_1/kernel:0' -> 'forward_gru_1/kernel'."""
    w = wname[:-2] if wname.endswith(":0") else wname
    if w.startswith(layer + "/"):
        w = w[len(layer) + 1:]
    return w


def load_keras_weights(path, group=None):
    """Flatten a Keras weight file to {"<layer>/<weight>": array}, in `layer_names`/`weight_names` order.

    Mirrors what `model.load_weights(path)` consumes (reference utils.py:305,328; train.py:171).
    `group="model_weights"` selects the weights inside a full `final_model.h5`.
    """
    root = read_h5(path)
    g = root[group] if group else (root["model_weights"] if "model_weights" in root.children else root)
    out = OrderedDict()
    for layer in g.attrs.get("layer_names", list(g.children)):
        lg = g.children[layer]
        for wname in lg.attrs.get("weight_names", []):
            out[f"{layer}/{_short(layer, wname)}"] = np.ascontiguousarray(lg[wname])
    return out


# ------------------------------------------------------------------------------------------
# Writer: emits the same subset (superblock v0, symbol-table groups, v1 headers, contiguous data).
# ------------------------------------------------------------------------------------------
class _Writer:
    LEAF_K, INT_K = 4, 16

    def __init__(self):
        self.buf = bytearray(b"\0" * (24 + 32 + 40))  # superblock placeholder

    def alloc(self, data: bytes, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        off = len(self.buf)
        self.buf += data
        return off

    @staticmethod
    def _pad8(b: bytes):
        return b + b"\0" * ((-len(b)) % 8)

    @staticmethod
    def dt_float32():
        # class 1 v1, bits: LE, mantissa normalisation implied-1 (0x20), sign bit 31; props
        return bytes([0x11, 0x20, 0x1F, 0x00]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def dt_string(n):
        return bytes([0x13, 0x00, 0x00, 0x00]) + struct.pack("<I", n)

    @staticmethod
    def ds_simple(dims):
        if not dims:
            return struct.pack("<BBBB4x", 1, 0, 0, 0)
        return struct.pack("<BBBB4x", 1, len(dims), 0, 0) + b"".join(struct.pack("<Q", d) for d in dims)

    def msg(self, mtype, body: bytes, flags=0):
        body = self._pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def attr_strings(self, name, strings):
        width = max([len(s) for s in strings] + [1])
        dt = self.dt_string(width)
        ds = self.ds_simple([len(strings)])
        nm = name.encode() + b"\0"
        data = b"".join(s.encode().ljust(width, b"\0") for s in strings)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + self._pad8(nm) + self._pad8(dt) + self._pad8(ds) + data
        return self.msg(0x000C, body)

    def attr_scalar_string(self, name, s):
        raw = s.encode()
        dt = self.dt_string(max(len(raw), 1))
        ds = self.ds_simple([])
        nm = name.encode() + b"\0"
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + self._pad8(nm) + self._pad8(dt) + self._pad8(ds) + (raw or b"\0")
        return self.msg(0x000C, body)

    def object_header(self, msgs):
        payload = b"".join(msgs)
        return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(payload)) + payload

    @staticmethod
    def dt_int64():
        # class 0 (fixed point) v1, little-endian, signed (bit 3); size 8; bit offset 0, precision 64
        return bytes([0x10, 0x08, 0x00, 0x00]) + struct.pack("<I", 8) + struct.pack("<HH", 0, 64)

    def write_dataset(self, arr: np.ndarray):
        arr = np.asarray(arr)
        is_int = arr.dtype.kind in "iu"
        arr = np.array(arr, dtype="<i8" if is_int else "<f4", order="C")      # (np.ascontiguousarray would turn a 0-d scalar into shape (1,))
        daddr = self.alloc(arr.tobytes())
        msgs = [
            self.msg(0x0001, self.ds_simple(list(arr.shape))),
            self.msg(0x0003, self.dt_int64() if is_int else self.dt_float32(), flags=1),
            self.msg(0x0008, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes)),
        ]
        return self.alloc(self.object_header(msgs))

    def write_group(self, entries, attr_msgs):
        """entries: list[(name, ohdr_addr)] -> object header address of the new group."""
        entries = sorted(entries, key=lambda e: e[0].encode())
        heap = bytearray(b"\0" * 8)
        offs = []
        for name, _ in entries:
            offs.append(len(heap))
            heap += name.encode() + b"\0"
            while len(heap) % 8:
                heap.append(0)
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), _UNDEF, heap_data))
        # leaf SNODs of <= 2*LEAF_K symbols each, one level-0 TREE node (<= 2*INT_K children)
        per = 2 * self.LEAF_K
        chunks = [list(range(i, min(i + per, len(entries)))) for i in range(0, len(entries), per)] or [[]]
        if len(chunks) > 2 * self.INT_K:
            raise ValueError("group too large for single-level b-tree writer")
        snods, keys = [], [0]
        for ch in chunks:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(ch))
            for i in ch:
                body += struct.pack("<QQII16x", offs[i], entries[i][1], 0, 0)
            body += b"\0" * (40 * (per - len(ch)))
            snods.append(self.alloc(body))
            keys.append(offs[ch[-1]] if ch else 0)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(chunks), _UNDEF, _UNDEF)
        for i, s in enumerate(snods):
            tree += struct.pack("<QQ", keys[i], s)
        tree += struct.pack("<Q", keys[len(snods)])
        tree += b"\0" * (16 * (2 * self.INT_K - len(snods)))
        btree_addr = self.alloc(tree)
        msgs = [self.msg(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + list(attr_msgs)
        return self.alloc(self.object_header(msgs)), btree_addr, heap_addr

    def finish(self, root_ohdr, root_bt, root_heap):
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INT_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", root_bt, root_heap)
        self.buf[0:len(sb)] = sb
        return bytes(self.buf)


def save_keras_weights(path, layers, extra_root_attrs=None):
    """Write a Keras-2.2.2 layout weight file (reference train.py:215 `model.save_weights`).

    layers: OrderedDict layer_name -> OrderedDict weight_name ("<layer>/<w>:0" full Keras name) -> ndarray.
    Layers without weights are listed in `layer_names` with an empty `weight_names`, like Keras does.
    """
    w = _Writer()
    top = []
    for lname, weights in layers.items():
        # nested path: dataset lives at <layer>/<weight_name> where weight_name contains '/'
        tree = OrderedDict()
        for wname, arr in weights.items():
            node = tree
            parts = wname.split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, OrderedDict())
            node[parts[-1]] = arr

        def emit(node):
            ents = []
            for k, v in node.items():
                if isinstance(v, OrderedDict):
                    sub = emit(v)
                    ents.append((k, w.write_group(sub, [])[0]))
                else:
                    ents.append((k, w.write_dataset(v)))
            return ents

        ents = emit(tree)
        addr, _, _ = w.write_group(ents, [w.attr_strings("weight_names", list(weights.keys()))] if weights
                                   else [w.attr_strings("weight_names", [])] if False else
                                   ([w.attr_strings("weight_names", list(weights.keys()))] if weights else []))
        top.append((lname, addr))
    attrs = [w.attr_strings("layer_names", list(layers.keys())),
             w.attr_scalar_string("backend", "tensorflow"),
             w.attr_scalar_string("keras_version", "2.2.2")]
    for k, v in (extra_root_attrs or {}).items():
        attrs.append(w.attr_scalar_string(k, v))
    root, bt, hp = w.write_group(top, attrs)
    with open(path, "wb") as f:
        f.write(w.finish(root, bt, hp))


# ------------------------------------------------------------------------------------------
# Full-model files (`model.save`, reference train.py:216): /model_weights + /optimizer_weights, Keras 2.2.2 layout
# ------------------------------------------------------------------------------------------
def _emit_tree(w, named_arrays):
    """named_arrays: OrderedDict "a/b/c" -> ndarray; returns the symbol-table entries of the top level."""
    tree = OrderedDict()
    for name, arr in named_arrays.items():
        node = tree
        parts = name.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, OrderedDict())
        node[parts[-1]] = arr

    def emit(node):
        ents = []
        for k, v in node.items():
            ents.append((k, w.write_group(emit(v), [])[0]) if isinstance(v, OrderedDict) else (k, w.write_dataset(v)))
        return ents
    return emit(tree)


def keras_adam_weight_names(n_trainable):
    """Names Keras 2.2.2 gives the Adam slots: iterations, then m_k, v_k and the (1,)-shaped vhat placeholders, numbered in creation order."""
    var = lambda k: "training/Adam/Variable:0" if k == 0 else "training/Adam/Variable_%d:0" % k
    return ["Adam/iterations:0"] + [var(k) for k in range(3 * n_trainable)]


def keras_sgd_weight_names(n_trainable):
    """Names Keras 2.2.2 gives the SGD slots (`SGD.weights = [iterations] + moments`): iterations, then one velocity per trainable weight."""
    var = lambda k: "training/SGD/Variable:0" if k == 0 else "training/SGD/Variable_%d:0" % k
    return ["SGD/iterations:0"] + [var(k) for k in range(n_trainable)]


def save_keras_model(path, layers, adam=None, root_attrs=None, sgd=None):
    """Write `final_model.h5` with the layout of the reference's files (models/*/final_model.h5): root attrs keras_version / backend /
    model_config / training_config; /model_weights = what save_keras_weights writes at the root; /optimizer_weights with attr
    weight_names = [Adam/iterations:0, training/Adam/Variable:0 ...] in Keras' order (m of every trainable weight in layer/weight order,
    then v, then one zero of shape (1,) per weight for the unused amsgrad slot), iterations as an int64 scalar.

    layers: as for save_keras_weights.  adam: None or (iterations, [m arrays], [v arrays]) in trainable-weight order;
    sgd: None or (iterations, [velocity arrays]) -- the reference's default optimiser (train.py:190), slots named as Keras' SGD names them."""
    w = _Writer()
    top = []
    for lname, weights in layers.items():
        ents = _emit_tree(w, weights)
        top.append((lname, w.write_group(ents, [w.attr_strings("weight_names", list(weights.keys()))] if weights else [])[0]))
    mw = w.write_group(top, [w.attr_strings("layer_names", list(layers.keys())), w.attr_scalar_string("backend", "tensorflow"),
                             w.attr_scalar_string("keras_version", "2.2.2")])[0]
    root_entries = [("model_weights", mw)]
    if adam is not None:
        it, ms, vs = adam
        if len(ms) != len(vs):
            raise ValueError("adam: m and v lists differ in length")
        names = keras_adam_weight_names(len(ms))
        arrays = [np.asarray(int(it), np.int64)] + [np.asarray(a, np.float32) for a in ms] + [np.asarray(a, np.float32) for a in vs] + \
                 [np.zeros((1,), np.float32) for _ in ms]
        ents = _emit_tree(w, OrderedDict(zip(names, arrays)))
        root_entries.append(("optimizer_weights", w.write_group(ents, [w.attr_strings("weight_names", names)])[0]))
    elif sgd is not None:
        it, vel = sgd
        names = keras_sgd_weight_names(len(vel))
        arrays = [np.asarray(int(it), np.int64)] + [np.asarray(a, np.float32) for a in vel]
        ents = _emit_tree(w, OrderedDict(zip(names, arrays)))
        root_entries.append(("optimizer_weights", w.write_group(ents, [w.attr_strings("weight_names", names)])[0]))
    attrs = [w.attr_scalar_string("keras_version", "2.2.2"), w.attr_scalar_string("backend", "tensorflow")]
    for k, v in (root_attrs or {}).items():
        attrs.append(w.attr_scalar_string(k, v))
    root, bt, hp = w.write_group(root_entries, attrs)
    with open(path, "wb") as f:
        f.write(w.finish(root, bt, hp))


def load_keras_adam_state(path):
    """(iterations, OrderedDict "<layer>/<weight>" -> m, same -> v) of a Keras-2.2.2 `final_model.h5` (the reference's or ours): the k-th
    trainable weight in layer_names / weight_names order (BatchNorm moving statistics are not trainable) owns Variable_k (m) and
    Variable_{k+n} (v).  Shapes are checked."""
    root = read_h5(path)
    if "optimizer_weights" not in root.children:
        raise ValueError("no optimizer_weights group in %s" % path)
    ow, mw = root["optimizer_weights"], root["model_weights"]
    trainable = []
    for layer in mw.attrs.get("layer_names", list(mw.children)):
        lg = mw.children[layer]
        for wname in lg.attrs.get("weight_names", []):
            if "moving_mean" in wname or "moving_variance" in wname:
                continue
            trainable.append((f"{layer}/{_short(layer, wname)}", lg[wname].shape))
    n = len(trainable)
    names = keras_adam_weight_names(n)
    m, v = OrderedDict(), OrderedDict()
    for k, (name, shape) in enumerate(trainable):
        mk, vk = np.ascontiguousarray(ow[names[1 + k]]), np.ascontiguousarray(ow[names[1 + n + k]])
        if mk.shape != tuple(shape) or vk.shape != tuple(shape):
            raise ValueError("optimizer slot %d does not match %s %s" % (k, name, shape))
        m[name], v[name] = mk, vk
    return int(np.asarray(ow["Adam/iterations:0"]).reshape(-1)[0]), m, v

