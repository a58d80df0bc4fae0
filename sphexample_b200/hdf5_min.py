"""A minimal HDF5 *writer* — just enough of the file format for the reference's VTKHDF output
(src/ProduceHDFVTK.jl) — because neither libhdf5 nor h5py exists in this image (SURVEY §8f, N3:
"needs a hand-rolled HDF5 writer").

What is written is the classic (HDF5 1.0-1.6 compatible, "earliest" libver) layout that every HDF5
library reads: superblock version 0, version-1 object headers, groups as symbol tables (one v1
B-tree node of level 0, symbol-table nodes of up to 8 entries, one local heap for the names),
contiguous datasets of little-endian fixed-point / IEEE floating-point numbers, and version-1
attribute messages (numeric arrays and fixed-length ASCII strings).  No chunking, filters, links,
references, compound or variable-length types.  Layout follows the "HDF5 File Format Specification
Version 2.0" (sections II.A superblock, III.A B-trees v1, III.B symbol table nodes, III.D local
heaps, IV.A object headers and IV.A.2.* messages 0x01, 0x03, 0x05, 0x08, 0x0C, 0x11).

    root = Group()
    g = root.group("VTKHDF"); g.attrs["Version"] = np.array([2, 3], np.int64); g.attrs["Type"] = b"PolyData"
    g.dataset("Points", xyz)                 # a numpy array, or a Spool (data streamed from a temp file)
    write_file(path, root)
"""
from __future__ import annotations

import os
import struct
import tempfile
from typing import Dict, Optional, Tuple, Union

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 4, 16            # superblock defaults: 2K = 8 symbols per SNOD, 2K = 32 children per B-tree node
SNOD_CAP, BTREE_CAP = 2 * LEAF_K, 2 * INTERNAL_K
SNOD_SIZE = 8 + SNOD_CAP * 40
BTREE_SIZE = 24 + BTREE_CAP * 8 + (BTREE_CAP + 1) * 8


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


class Spool:
    """Row-major data of one dataset streamed to a temporary file step by step (the transient VTKHDF
    file appends every output step to the same datasets; without chunked storage the final size must
    be known before the dataset is laid out, so the rows wait here until the file is closed)."""

    def __init__(self, dtype, row_shape: Tuple[int, ...] = ()):
        self.dtype = np.dtype(dtype).newbyteorder("<")
        self.row_shape = tuple(int(x) for x in row_shape)
        self.rows = 0
        self._fh = tempfile.TemporaryFile()

    def append(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype).reshape((-1,) + self.row_shape)
        self._fh.write(a.tobytes())
        self.rows += a.shape[0]

    @property
    def shape(self):
        return (self.rows,) + self.row_shape

    @property
    def nbytes(self):
        return self.rows * int(np.prod(self.row_shape, dtype=np.int64)) * self.dtype.itemsize

    def copy_to(self, out, chunk=1 << 24):
        self._fh.seek(0)
        while True:
            b = self._fh.read(chunk)
            if not b:
                break
            out.write(b)

    def close(self):
        self._fh.close()


class Dataset:
    def __init__(self, data: Union[np.ndarray, Spool]):
        if not isinstance(data, Spool):
            data = np.asarray(data)
            if data.dtype.kind not in "iuf":
                raise TypeError(f"hdf5_min: unsupported dataset dtype {data.dtype}")
            data = np.ascontiguousarray(data, dtype=data.dtype.newbyteorder("<"))
        self.data = data
        self.attrs: Dict[str, object] = {}


class Group:
    def __init__(self):
        self.children: Dict[str, Union["Group", Dataset]] = {}
        self.attrs: Dict[str, object] = {}

    def group(self, name: str) -> "Group":
        g = self.children.setdefault(name, Group())
        assert isinstance(g, Group)
        return g

    def dataset(self, name: str, data) -> Dataset:
        if len(self.children) >= SNOD_CAP * BTREE_CAP and name not in self.children:
            raise ValueError("hdf5_min: a group holds at most 256 links (one level-0 B-tree node)")
        d = Dataset(data)
        self.children[name] = d
        return d


# ---- messages -----------------------------------------------------------------------------------
def _datatype(dt: np.dtype) -> bytes:
    """IV.A.2.d datatype message, version 1: class 0 fixed-point / class 1 floating-point, little-endian"""
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00          # bit 3: signed (two's complement)
        return struct.pack("<BBBBIHH", 0x10 | 0, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        e_bits, m_bits, bias = (8, 23, 127) if dt.itemsize == 4 else (11, 52, 1023)
        # bits 4-5 of byte 0 = 2: mantissa normalisation "msb implied"; byte 1 = position of the sign bit
        return struct.pack("<BBBBIHHBBBBI", 0x10 | 1, 0x20, 8 * dt.itemsize - 1, 0, dt.itemsize, 0, 8 * dt.itemsize,
                           m_bits, e_bits, 0, m_bits, bias)
    raise TypeError(f"hdf5_min: unsupported dtype {dt}")


def _string_type(n: int) -> bytes:
    """fixed-length string: class 3; byte 0 bits 0-3 = 1 (null-padded), bits 4-7 = 0 (ASCII)"""
    return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, n)


def _dataspace(shape: Optional[Tuple[int, ...]]) -> bytes:
    """IV.A.2.b dataspace message, version 1 (rank 0 = scalar); no maximum dimensions"""
    shape = () if shape is None else tuple(shape)
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attribute(name: str, value) -> bytes:
    """IV.A.2.m attribute message, version 1: name, datatype and dataspace each padded to 8 bytes"""
    if isinstance(value, (bytes, str)):
        raw = value.encode("ascii") if isinstance(value, str) else value
        dt, ds, data = _string_type(max(len(raw), 1)), _dataspace(None), raw or b"\0"
    else:
        a = np.asarray(value)
        a = np.ascontiguousarray(a, dtype=a.dtype.newbyteorder("<"))
        dt, ds, data = _datatype(a.dtype), _dataspace(a.shape if a.ndim else None), a.tobytes()
    nm = name.encode("ascii") + b"\0"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + data
    return _message(0x000C, body)


def _object_header(messages) -> bytes:
    """IV.A.1.a version-1 object header prefix (16 bytes incl. alignment) + the messages of chunk 0"""
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


# ---- file assembly ------------------------------------------------------------------------------
class _File:
    def __init__(self, path):
        self.fh = open(path, "wb")
        self.end = 96                                     # the version-0 superblock with 8-byte offsets / lengths

    def alloc(self, size: int) -> int:
        addr = self.end
        self.end += size + (-size % 8)
        return addr

    def put(self, addr: int, data: bytes):
        self.fh.seek(addr)
        self.fh.write(data)


def _write_dataset(f: _File, d: Dataset) -> int:
    data = d.data
    nbytes = int(data.nbytes)
    addr = f.alloc(nbytes) if nbytes else UNDEF
    if nbytes:
        f.fh.seek(addr)
        if isinstance(data, Spool):
            data.copy_to(f.fh)
        else:
            f.fh.write(data.tobytes())
    msgs = [_message(0x0001, _dataspace(data.shape)),
            _message(0x0003, _datatype(data.dtype), flags=0x01),
            # fill value (new), version 2: allocation time late, write time "if set", no fill value defined
            _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0), flags=0x01),
            # data layout, version 3, class 1 = contiguous: address, size
            _message(0x0008, struct.pack("<BBQQ", 3, 1, addr, nbytes))]
    msgs += [_attribute(k, v) for k, v in d.attrs.items()]
    hdr = _object_header(msgs)
    haddr = f.alloc(len(hdr))
    f.put(haddr, hdr)
    return haddr


def _write_group(f: _File, g: Group) -> Tuple[int, int, int]:
    """children first (their header addresses go into the symbol table), then heap, symbol-table nodes,
    B-tree node and the group's own header; returns (header, btree, heap) addresses"""
    names = sorted(g.children, key=lambda s: s.encode("ascii"))        # strcmp order, as H5G_node_cmp wants
    entries = []
    for nm in names:
        child = g.children[nm]
        if isinstance(child, Group):
            h, bt, hp = _write_group(f, child)
            entries.append((nm, h, 1, struct.pack("<QQ", bt, hp)))      # cache type 1: B-tree / heap addresses
        else:
            entries.append((nm, _write_dataset(f, child), 0, b"\0" * 16))
    # local heap: offset 0 holds the empty string (the B-tree's first key), names 8-byte aligned, no free block
    seg, offs = bytearray(8), {}
    for nm in names:
        offs[nm] = len(seg)
        seg += _pad8(nm.encode("ascii") + b"\0")
    seg_addr = f.alloc(len(seg))
    f.put(seg_addr, bytes(seg))
    heap_addr = f.alloc(32)
    f.put(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), 1, seg_addr))     # free-list head 1 = H5HL_FREE_NULL
    # symbol-table nodes of <= 8 entries; key[i+1] of the B-tree = heap offset of the last name in node i
    keys, kids = [0], []
    for k in range(0, len(entries), SNOD_CAP):
        part = entries[k:k + SNOD_CAP]
        node = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
        for nm, haddr, cache, scratch in part:
            node += struct.pack("<QQII", offs[nm], haddr, cache, 0) + scratch
        node += b"\0" * (SNOD_SIZE - len(node))
        a = f.alloc(SNOD_SIZE)
        f.put(a, node)
        kids.append(a)
        keys.append(offs[part[-1][0]])
    assert len(kids) <= BTREE_CAP
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(kids), UNDEF, UNDEF)
    for i, a in enumerate(kids):
        bt += struct.pack("<QQ", keys[i], a)
    bt += struct.pack("<Q", keys[len(kids)])
    bt += b"\0" * (BTREE_SIZE - len(bt))
    bt_addr = f.alloc(BTREE_SIZE)
    f.put(bt_addr, bt)
    msgs = [_message(0x0011, struct.pack("<QQ", bt_addr, heap_addr))] + [_attribute(k, v) for k, v in g.attrs.items()]
    hdr = _object_header(msgs)
    haddr = f.alloc(len(hdr))
    f.put(haddr, hdr)
    return haddr, bt_addr, heap_addr


def write_file(path: str, root: Group) -> int:
    """Serialise the tree under `root` to `path`; returns the file size."""
    tmp = path + ".part"
    f = _File(tmp)
    try:
        haddr, bt, hp = _write_group(f, root)
        f.alloc(0)
        # pad the tail so that a speculative metadata read past the last header stays inside the file
        f.put(f.end, b"")
        f.fh.truncate(f.end)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, f.end, UNDEF)
        sb += struct.pack("<QQII", 0, haddr, 1, 0) + struct.pack("<QQ", bt, hp)       # root symbol-table entry
        assert len(sb) == 96
        f.put(0, sb)
    finally:
        f.fh.close()
    os.replace(tmp, path)
    return f.end
