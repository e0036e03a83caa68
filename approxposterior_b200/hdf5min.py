"""A minimal HDF5 writer (and matching reader) in pure NumPy, for the chain files the reference leaves behind.

The reference hands ``emcee.backends.HDFBackend(runName + ".h5")`` to its sampler (approx.py:828-840), so a run with
``cache=True`` ends with ``apRunii.h5`` files that users open with emcee / h5py.  h5py is not installable in this
image (no wheel, no libhdf5), so the engine writes the same layout itself:

    /mcmc                      group, attrs: version, nwalkers, ndim, has_blobs, iteration
    /mcmc/accepted             float64 [nwalkers]
    /mcmc/chain                float64 [iteration][nwalkers][ndim]
    /mcmc/log_prob             float64 [iteration][nwalkers]
    /mcmc/blobs                float64 [iteration][nwalkers]        (when the log-probability returns blobs)

File format (HDF5 File Format Specification, version 0 superblock -- the "classic" layout every libhdf5 since 1.0 reads):
superblock v0 -> root symbol-table entry -> per group: version-1 object header with a Symbol Table message, a version-1
group B-tree with one leaf, one symbol-table node (SNOD) and a local heap for the link names -> per dataset: version-1
object header with Dataspace (v1), Datatype (v1), Fill Value (v2), Data Layout (v3, contiguous) and version-1 Attribute
messages; raw data contiguous, little endian.  Differences from what h5py itself would write, none of which a reader of
the chain notices: datasets are contiguous, not chunked (the file is written once, when the chain is complete -- it cannot
be extended in place by a later emcee run); ``has_blobs`` is an int8 (h5py stores NumPy bools as an enum of int8);
``version`` is a fixed-length ASCII string.

``read_hdf5`` parses exactly this subset back (tests round-trip every file through it).  There is no libhdf5 in this
image to cross-check against: the layout follows the specification field by field (cited inline).
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K = 16           # symbol-table node holds up to 2 * LEAF_K entries
INTERNAL_K = 16       # B-tree node holds up to 2 * INTERNAL_K children
HEAP_FREE_NULL = 1    # H5HL_FREE_NULL: "no free block" in a local heap's free-list head


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# --------------------------------------------------------------------------------------------------- messages
def _datatype(dt):
    """Datatype message body (spec IV.A.2.d), version 1.  Fixed point, IEEE float and fixed-length strings."""
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        # class 1; bit field 0: byte order LE (bit 0 = 0), mantissa normalisation 2 (bits 4-5); bits 8-15: sign location
        exp_bits, man_bits, bias = (11, 52, 1023) if dt.itemsize == 8 else (8, 23, 127)
        head = struct.pack("<BBBBI", 0x11, 0x20, dt.itemsize * 8 - 1, 0, dt.itemsize)
        prop = struct.pack("<HHBBBBI", 0, dt.itemsize * 8, man_bits, exp_bits, 0, man_bits, bias)
        return head + prop
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        # class 3; bit field: padding 1 (null pad) in bits 0-3, character set 0 (ASCII) in bits 4-7
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)
    raise TypeError("hdf5min: unsupported dtype %r" % (dt,))


def _dataspace(shape):
    """Dataspace message body (spec IV.A.2.b), version 1, no maximum dimensions."""
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _message(mtype, body):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _attribute(name, value):
    """Attribute message (spec IV.A.2.m), version 1: name, datatype and dataspace each padded to 8 bytes."""
    arr = _as_array(value)
    nm = name.encode("ascii") + b"\0"
    dtb, dsb = _datatype(arr.dtype), _dataspace(arr.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtb), len(dsb)) + _pad8(nm) + _pad8(dtb) + _pad8(dsb) + arr.tobytes()
    return _message(0x000C, body)


def _as_array(value):
    if isinstance(value, (bytes, str)):
        b = value.encode("ascii") if isinstance(value, str) else value
        return np.array(b, dtype="S%d" % max(len(b), 1))
    arr = np.asarray(value)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.int8)
    if arr.dtype.kind == "U":
        arr = arr.astype("S")
    if arr.dtype.byteorder == ">":
        arr = arr.astype(arr.dtype.newbyteorder("<"))
    return arr if arr.flags.c_contiguous else arr.copy(order="C")   # (ascontiguousarray would turn scalars into [1])


def _object_header(messages):
    """Version-1 object header (spec IV.A.1.a): 12-byte prefix + 4 bytes of alignment, then the messages."""
    data = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data


# --------------------------------------------------------------------------------------------------- tree model
class Group(object):
    def __init__(self, attrs=None, children=None):
        self.attrs = dict(attrs or {})
        self.children = dict(children or {})          # name -> Group | array-like | Dataset


class Dataset(object):
    def __init__(self, data, attrs=None):
        self.data = _as_array(data)
        self.attrs = dict(attrs or {})


class _Writer(object):
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, nbytes):
        off = len(self.buf)
        self.buf += b"\0" * (nbytes + (-nbytes % 8))
        return off

    def put(self, off, b):
        self.buf[off:off + len(b)] = b

    # a dataset: header first (its size is known), then the raw data
    def write_dataset(self, ds):
        attrs = [_attribute(k, v) for k, v in ds.attrs.items()]
        fill = _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0))                  # v2: late alloc, write if set, undefined
        fixed = [_message(0x0001, _dataspace(ds.data.shape)), _message(0x0003, _datatype(ds.data.dtype)), fill]
        layout_len = len(_message(0x0008, struct.pack("<BBQQ", 3, 1, 0, 0)))
        hdr_len = 16 + sum(len(m) for m in fixed + attrs) + layout_len
        hdr_off = self.alloc(hdr_len)
        data_off = self.alloc(max(ds.data.nbytes, 1)) if ds.data.nbytes else UNDEF
        layout = _message(0x0008, struct.pack("<BBQQ", 3, 1, data_off, ds.data.nbytes))   # v3, class 1 = contiguous
        self.put(hdr_off, _object_header(fixed + [layout] + attrs))
        if ds.data.nbytes:
            self.put(data_off, ds.data.tobytes())
        return hdr_off

    # a group: children first, then local heap, symbol-table node, B-tree and the group's own header
    def write_group(self, grp):
        names = sorted(grp.children)
        if len(names) > 2 * LEAF_K:
            raise ValueError("hdf5min: at most %d links per group" % (2 * LEAF_K))
        entries = []
        for nm in names:
            child = grp.children[nm]
            if isinstance(child, Group):
                hdr, bt, hp = self.write_group(child)
                entries.append((nm, hdr, 1, struct.pack("<QQ", bt, hp)))           # cache type 1: B-tree + heap addresses
            else:
                ds = child if isinstance(child, Dataset) else Dataset(child)
                entries.append((nm, self.write_dataset(ds), 0, b"\0" * 16))
        # local heap (spec III.D): data segment = "" at offset 0, then the names, each NUL-terminated and padded to 8
        heap_data = bytearray(b"\0" * 8)
        name_off = {}
        for nm in names:
            name_off[nm] = len(heap_data)
            heap_data += _pad8(nm.encode("ascii") + b"\0")
        heap_off = self.alloc(32)
        heap_data_off = self.alloc(len(heap_data))
        self.put(heap_off, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), HEAP_FREE_NULL, heap_data_off))
        self.put(heap_data_off, bytes(heap_data))
        # symbol-table node (spec III.C): all 2K slots allocated, the used ones sorted by name
        snod_off = self.alloc(8 + 2 * LEAF_K * 40)
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for nm, hdr, ctype, scratch in entries:
            snod += struct.pack("<QQII", name_off[nm], hdr, ctype, 0) + scratch
        self.put(snod_off, snod)
        # group B-tree (spec III.A.1), one leaf-level node with one child; keys are heap offsets of names:
        # key[0] = "" (smaller than every name), key[1] = the largest name in the child
        bt_off = self.alloc(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8)
        bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        if names:
            bt += struct.pack("<QQQ", 0, snod_off, name_off[names[-1]])
        self.put(bt_off, bt)
        msgs = [_message(0x0011, struct.pack("<QQ", bt_off, heap_off))] + [_attribute(k, v) for k, v in grp.attrs.items()]
        hdr = _object_header(msgs)
        hdr_off = self.alloc(len(hdr))
        self.put(hdr_off, hdr)
        return hdr_off, bt_off, heap_off


def write_hdf5(path, root):
    """Write ``root`` (a Group) to ``path``."""
    w = _Writer()
    w.alloc(96)                                                                     # superblock v0 + root entry
    hdr, bt, hp = w.write_group(root)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(w.buf), UNDEF)                         # base, free-space, EOF, driver info
    sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", bt, hp)           # root symbol-table entry
    w.put(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(w.buf))


# --------------------------------------------------------------------------------------------------- reader (same subset)
def _read_datatype(b):
    cls, b0, b1, _b2, size = struct.unpack_from("<BBBBI", b, 0)
    c = cls & 0x0F
    if c == 1:
        return np.dtype("<f%d" % size)
    if c == 0:
        return np.dtype("<%s%d" % ("i" if b0 & 0x08 else "u", size))
    if c == 3:
        return np.dtype("S%d" % size)
    raise TypeError("hdf5min: datatype class %d not supported" % c)


def _read_dataspace(b):
    ver, rank, flags = struct.unpack_from("<BBB", b, 0)
    if ver != 1:
        raise ValueError("hdf5min: dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, b, 8)) if rank else ()


def _read_messages(buf, off):
    ver, _r, nmsg, _ref, size = struct.unpack_from("<BBHII", buf, off)
    if ver != 1:
        raise ValueError("hdf5min: object header version %d" % ver)
    p, end, out = off + 16, off + 16 + size, []
    while p < end and len(out) < nmsg:
        mtype, msize, _fl = struct.unpack_from("<HHB", buf, p)
        out.append((mtype, bytes(buf[p + 8:p + 8 + msize])))
        p += 8 + msize
    return out


def _read_attrs(msgs):
    attrs = {}
    for mtype, body in msgs:
        if mtype != 0x000C:
            continue
        _ver, _r, nlen, dtlen, dslen = struct.unpack_from("<BBHHH", body, 0)
        p = 8
        name = body[p:p + nlen].split(b"\0")[0].decode("ascii"); p += nlen + (-nlen % 8)
        dt = _read_datatype(body[p:p + dtlen]); p += dtlen + (-dtlen % 8)
        shape = _read_dataspace(body[p:p + dslen]); p += dslen + (-dslen % 8)
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(body, dtype=dt, count=n, offset=p)
        attrs[name] = val.reshape(shape).copy() if shape else (val[0].decode("ascii") if dt.kind == "S" else val[0])
    return attrs


def _read_object(buf, off):
    msgs = _read_messages(buf, off)
    kinds = dict(msgs)
    attrs = _read_attrs(msgs)
    if 0x0011 in kinds:                                                             # group
        bt, hp = struct.unpack_from("<QQ", kinds[0x0011], 0)
        assert buf[hp:hp + 4] == b"HEAP" and buf[bt:bt + 4] == b"TREE"
        heap_data = struct.unpack_from("<Q", buf, hp + 24)[0]
        _t, level, used = struct.unpack_from("<BBH", buf, bt + 4)
        assert level == 0, "hdf5min reads single-level group trees"
        children = {}
        for k in range(used):
            snod = struct.unpack_from("<Q", buf, bt + 24 + 8 + 16 * k)[0]
            assert buf[snod:snod + 4] == b"SNOD"
            nsym = struct.unpack_from("<H", buf, snod + 6)[0]
            for e in range(nsym):
                noff, hdr = struct.unpack_from("<QQ", buf, snod + 8 + 40 * e)
                nm = bytes(buf[heap_data + noff:]).split(b"\0")[0].decode("ascii")
                children[nm] = _read_object(buf, hdr)
        return Group(attrs, children)
    dt, shape = _read_datatype(kinds[0x0003]), _read_dataspace(kinds[0x0001])
    ver, cls, addr, size = struct.unpack_from("<BBQQ", kinds[0x0008], 0)
    assert ver == 3 and cls == 1, "hdf5min reads contiguous version-3 layouts"
    n = int(np.prod(shape)) if shape else 1
    data = np.frombuffer(buf, dtype=dt, count=n, offset=addr).reshape(shape).copy() if size else np.empty(shape, dt)
    return Dataset(data, attrs)


def read_hdf5(path):
    """Read a file written by ``write_hdf5`` back into a Group tree (Dataset leaves carry ``.data`` and ``.attrs``)."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != SIGNATURE or buf[8] != 0:
        raise ValueError("hdf5min: not a version-0 superblock HDF5 file")
    so, sl = buf[13], buf[14]
    if (so, sl) != (8, 8):
        raise ValueError("hdf5min: 8-byte offsets and lengths only")
    eof = struct.unpack_from("<Q", buf, 40)[0]
    if eof != len(buf):
        raise ValueError("hdf5min: end-of-file address %d != file size %d" % (eof, len(buf)))
    root_hdr = struct.unpack_from("<Q", buf, 56 + 8)[0]
    return _read_object(buf, root_hdr)


# --------------------------------------------------------------------------------------------------- emcee layout
def write_emcee_backend(path, chain, log_prob, blobs=None, accepted=None, name="mcmc", version="3.0.2"):
    """``chain`` [iteration][nwalkers][ndim], ``log_prob`` [iteration][nwalkers] (+ ``blobs``): the group emcee's
    HDFBackend reads with get_chain / get_log_prob / get_blobs / accepted / iteration / shape (emcee/backends/hdf.py)."""
    chain = np.ascontiguousarray(chain, dtype=np.float64)
    log_prob = np.ascontiguousarray(log_prob, dtype=np.float64)
    it, nw, nd = chain.shape
    kids = {"accepted": np.zeros(nw) if accepted is None else np.asarray(accepted, dtype=np.float64),
            "chain": chain, "log_prob": log_prob}
    if blobs is not None:
        kids["blobs"] = np.ascontiguousarray(blobs, dtype=np.float64)
    attrs = {"version": version, "nwalkers": np.int64(nw), "ndim": np.int64(nd), "has_blobs": blobs is not None,
             "iteration": np.int64(it)}
    write_hdf5(path, Group(children={name: Group(attrs, kids)}))


def read_emcee_backend(path, name="mcmc"):
    g = read_hdf5(path).children[name]
    out = {k: v.data for k, v in g.children.items()}
    out["attrs"] = g.attrs
    return out
