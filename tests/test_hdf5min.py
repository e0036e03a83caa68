"""The minimal HDF5 writer behind the chain cache (reference approx.py:829-833 hands emcee.backends.HDFBackend(runName +
".h5") to its sampler).  No libhdf5 exists in this image, so the file is checked (a) through the package's own reader and
(b) field by field against the HDF5 File Format Specification (version-0 superblock, version-1 object headers / group
B-tree / symbol-table node / local heap) at fixed byte offsets computed here independently of the writer."""
import struct

import numpy as np
import pytest

from approxposterior_b200 import hdf5min as h5


def _file(tmp_path, with_blobs=True):
    rng = np.random.default_rng(3)
    chain = rng.standard_normal((7, 6, 2))
    logp = rng.standard_normal((7, 6))
    blobs = rng.standard_normal((7, 6)) if with_blobs else None
    acc = np.arange(6, dtype=float)
    path = str(tmp_path / "chain.h5")
    h5.write_emcee_backend(path, chain, logp, blobs=blobs, accepted=acc)
    return path, chain, logp, blobs, acc


@pytest.mark.parametrize("with_blobs", [True, False])
def test_emcee_layout_round_trip(tmp_path, with_blobs):
    path, chain, logp, blobs, acc = _file(tmp_path, with_blobs)
    out = h5.read_emcee_backend(path)
    assert np.array_equal(out["chain"], chain) and np.array_equal(out["log_prob"], logp) and np.array_equal(out["accepted"], acc)
    assert ("blobs" in out) == with_blobs and (not with_blobs or np.array_equal(out["blobs"], blobs))
    a = out["attrs"]
    # what emcee/backends/hdf.py reads: iteration, nwalkers, ndim, has_blobs (+ version)
    assert int(a["iteration"]) == 7 and int(a["nwalkers"]) == 6 and int(a["ndim"]) == 2 and bool(a["has_blobs"]) == with_blobs
    assert a["version"] == "3.0.2"
    root = h5.read_hdf5(path)
    assert list(root.children) == ["mcmc"] and sorted(root.children["mcmc"].children) == sorted(
        ["accepted", "chain", "log_prob"] + (["blobs"] if with_blobs else []))


def test_superblock_and_root_entry_follow_the_specification(tmp_path):
    path, *_ = _file(tmp_path)
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    ver_sb, ver_fs, ver_root, _r, ver_shm, size_off, size_len, _r2 = struct.unpack_from("<8B", b, 8)
    assert (ver_sb, ver_fs, ver_root, ver_shm, size_off, size_len) == (0, 0, 0, 0, 8, 8)
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", b, 16)
    assert leaf_k == h5.LEAF_K and internal_k == h5.INTERNAL_K and flags == 0
    base, freespace, eof, driver = struct.unpack_from("<4Q", b, 24)
    assert base == 0 and freespace == h5.UNDEF and driver == h5.UNDEF and eof == len(b) and len(b) % 8 == 0
    name_off, hdr, cache, _res, bt, hp = struct.unpack_from("<QQIIQQ", b, 56)
    assert name_off == 0 and cache == 1 and hdr % 8 == 0
    # root object header: version 1, one Symbol Table message (type 0x11, 16 bytes) pointing at the same B-tree and heap
    ver, _r, nmsg, refs, hsize = struct.unpack_from("<BBHII", b, hdr)
    assert (ver, nmsg, refs, hsize) == (1, 1, 1, 24)
    mtype, msize, mflags = struct.unpack_from("<HHB", b, hdr + 16)
    assert (mtype, msize, mflags) == (0x11, 16, 0) and struct.unpack_from("<QQ", b, hdr + 24) == (bt, hp)
    # group B-tree node: one child; keys are heap offsets: "" first, then the (only, hence largest) name "mcmc"
    assert b[bt:bt + 4] == b"TREE"
    ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, bt + 4)
    assert (ntype, level, used, left, right) == (0, 0, 1, h5.UNDEF, h5.UNDEF)
    key0, snod, key1 = struct.unpack_from("<QQQ", b, bt + 24)
    # local heap: version 0, free list empty (H5HL_FREE_NULL = 1), data segment starts with the empty name
    assert b[hp:hp + 4] == b"HEAP" and b[hp + 4] == 0
    dsize, free_head, daddr = struct.unpack_from("<QQQ", b, hp + 8)
    assert free_head == 1 and dsize % 8 == 0 and b[daddr] == 0
    assert key0 == 0 and b[daddr + key1:daddr + key1 + 5] == b"mcmc\0"
    # symbol-table node: version 1, one symbol, entry = (name offset, header address, cache type 1, B-tree, heap of /mcmc)
    assert b[snod:snod + 4] == b"SNOD" and b[snod + 4] == 1 and struct.unpack_from("<H", b, snod + 6)[0] == 1
    e_name, e_hdr, e_cache, _res, e_bt, e_hp = struct.unpack_from("<QQIIQQ", b, snod + 8)
    assert e_name == key1 and e_cache == 1 and b[e_bt:e_bt + 4] == b"TREE" and b[e_hp:e_hp + 4] == b"HEAP"


def test_dataset_header_messages_follow_the_specification(tmp_path):
    path, chain, *_ = _file(tmp_path)
    b = open(path, "rb").read()
    # walk root -> /mcmc -> its symbol-table node, find "chain" (entries are sorted: accepted, blobs, chain, log_prob)
    root_hdr = struct.unpack_from("<Q", b, 64)[0]
    bt = struct.unpack_from("<Q", b, root_hdr + 24)[0]
    snod = struct.unpack_from("<Q", b, bt + 32)[0]
    mcmc_hdr = struct.unpack_from("<Q", b, snod + 16)[0]
    mbt, mhp = struct.unpack_from("<QQ", b, mcmc_hdr + 24)
    msnod = struct.unpack_from("<Q", b, mbt + 32)[0]
    heap_data = struct.unpack_from("<Q", b, mhp + 24)[0]
    nsym = struct.unpack_from("<H", b, msnod + 6)[0]
    names, hdrs = [], []
    for e in range(nsym):
        noff, hdr, cache = struct.unpack_from("<QQI", b, msnod + 8 + 40 * e)
        names.append(b[heap_data + noff:].split(b"\0")[0].decode()); hdrs.append(hdr)
        assert cache == 0
    assert names == sorted(names) == ["accepted", "blobs", "chain", "log_prob"]
    hdr = hdrs[names.index("chain")]
    ver, _r, nmsg, refs, hsize = struct.unpack_from("<BBHII", b, hdr)
    assert ver == 1 and nmsg == 4 and refs == 1
    p = hdr + 16
    # 1: dataspace v1, rank 3, no max dims, dims (7, 6, 2)
    assert struct.unpack_from("<HHB", b, p) == (0x0001, 32, 0)
    assert struct.unpack_from("<BBB", b, p + 8) == (1, 3, 0) and struct.unpack_from("<3Q", b, p + 16) == (7, 6, 2)
    p += 8 + 32
    # 2: datatype v1 class 1 (floating point), little endian, IEEE binary64: sign bit 63, exponent 11 bits at 52 (bias 1023),
    #    mantissa 52 bits at 0 with an implied leading bit
    assert struct.unpack_from("<HHB", b, p) == (0x0003, 24, 0)
    assert struct.unpack_from("<BBBBI", b, p + 8) == (0x11, 0x20, 63, 0, 8)
    assert struct.unpack_from("<HHBBBBI", b, p + 16) == (0, 64, 52, 11, 0, 52, 1023)
    p += 8 + 24
    # 3: fill value v2: late allocation, written if set, undefined
    assert struct.unpack_from("<HHB", b, p) == (0x0005, 8, 0) and struct.unpack_from("<4B", b, p + 8) == (2, 2, 2, 0)
    p += 8 + 8
    # 4: data layout v3, contiguous, address + size; the raw little-endian doubles sit there
    assert struct.unpack_from("<HHB", b, p) == (0x0008, 24, 0)
    lver, lclass, addr, size = struct.unpack_from("<BBQQ", b, p + 8)
    assert (lver, lclass, size) == (3, 1, chain.nbytes) and addr % 8 == 0
    assert np.array_equal(np.frombuffer(b, dtype="<f8", count=chain.size, offset=addr).reshape(chain.shape), chain)
    assert p + 8 + 24 == hdr + 16 + hsize


def test_attribute_message_layout(tmp_path):
    path, *_ = _file(tmp_path)
    b = open(path, "rb").read()
    root_hdr = struct.unpack_from("<Q", b, 64)[0]
    bt = struct.unpack_from("<Q", b, root_hdr + 24)[0]
    snod = struct.unpack_from("<Q", b, bt + 32)[0]
    mcmc_hdr = struct.unpack_from("<Q", b, snod + 16)[0]
    nmsg = struct.unpack_from("<H", b, mcmc_hdr + 2)[0]
    assert nmsg == 1 + 5                                  # symbol table + version, nwalkers, ndim, has_blobs, iteration
    p = mcmc_hdr + 16 + 8 + 16                            # past the symbol-table message
    seen = {}
    for _ in range(5):
        mtype, msize, _fl = struct.unpack_from("<HHB", b, p)
        assert mtype == 0x000C and msize % 8 == 0
        ver, _r, nlen, dtlen, dslen = struct.unpack_from("<BBHHH", b, p + 8)
        assert ver == 1 and dslen == 8                    # scalar dataspace: version 1, rank 0
        q = p + 16
        name = b[q:q + nlen - 1].decode(); q += nlen + (-nlen % 8)
        cls = b[q] & 0x0F; dsize = struct.unpack_from("<I", b, q + 4)[0]; q += dtlen + (-dtlen % 8)
        assert struct.unpack_from("<BBB", b, q) == (1, 0, 0); q += 8
        seen[name] = (cls, dsize, b[q:q + dsize])
        p += 8 + msize
    assert seen["iteration"][:2] == (0, 8) and struct.unpack("<q", seen["iteration"][2])[0] == 7
    assert seen["nwalkers"][:2] == (0, 8) and seen["ndim"][:2] == (0, 8)
    assert seen["has_blobs"][:2] == (0, 1) and seen["has_blobs"][2] == b"\x01"
    assert seen["version"][0] == 3 and seen["version"][2] == b"3.0.2"


def test_generic_tree_and_limits(tmp_path):
    path = str(tmp_path / "t.h5")
    tree = h5.Group(attrs={"note": "x"}, children={
        "a": h5.Group(children={"i4": np.arange(5, dtype=np.int32), "u1": np.arange(3, dtype=np.uint8)}),
        "empty": np.zeros((0, 4)), "scalar": h5.Dataset(np.float64(2.5), attrs={"unit": "s", "vec": np.arange(3.0)})})
    h5.write_hdf5(path, tree)
    back = h5.read_hdf5(path)
    assert back.attrs["note"] == "x" and np.array_equal(back.children["a"].children["i4"].data, np.arange(5, dtype=np.int32))
    assert back.children["a"].children["u1"].data.dtype == np.uint8 and back.children["empty"].data.shape == (0, 4)
    assert float(back.children["scalar"].data) == 2.5 and back.children["scalar"].attrs["unit"] == "s"
    assert np.array_equal(back.children["scalar"].attrs["vec"], np.arange(3.0))
    with pytest.raises(ValueError):
        h5.write_hdf5(path, h5.Group(children={"d%02d" % i: np.zeros(1) for i in range(2 * h5.LEAF_K + 1)}))
    with pytest.raises(TypeError):
        h5.write_hdf5(path, h5.Group(children={"c": np.zeros(2, dtype=complex)}))
