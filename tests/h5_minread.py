"""A minimal HDF5 *reader* for the tests of sphexample_b200/hdf5_min.py — test infrastructure, written
against the HDF5 File Format Specification (superblock v0, v1 object headers, symbol-table groups,
local heaps, contiguous datasets, v1 attributes) the way a foreign library would walk the file:
from the superblock, through the root symbol-table entry, B-tree nodes, symbol-table nodes and the
local heap to the object headers.  It validates every signature, size and alignment it passes.
(No libhdf5 / h5py exists in this image; where h5py is importable the tests use it as well.)"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


def _need(cond, msg):
    if not cond:
        raise H5Error(msg)


class Reader:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        _need(b[:8] == b"\x89HDF\r\n\x1a\n", "signature")
        ver, fsver, rgver, _, shver, so, sl, _, self.leaf_k, self.int_k, flags = struct.unpack_from("<BBBBBBBBHHI", b, 8)
        _need((ver, fsver, rgver, shver) == (0, 0, 0, 0), "superblock version fields")
        _need((so, sl) == (8, 8), "offset / length sizes")
        base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
        _need(base == 0 and free == UNDEF and drv == UNDEF, "base / free-space / driver addresses")
        _need(eof == len(b), f"end-of-file address {eof} != file size {len(b)}")
        self.root_entry = self._entry(56)
        _need(self.root_entry["cache"] == 1, "root entry caches the symbol table")

    def _entry(self, off):
        name_off, hdr, cache, _ = struct.unpack_from("<QQII", self.b, off)
        bt, hp = struct.unpack_from("<QQ", self.b, off + 24)
        return {"name_off": name_off, "header": hdr, "cache": cache, "btree": bt, "heap": hp}

    # ---- object headers ----
    def messages(self, addr):
        b = self.b
        _need(addr % 8 == 0, "object header alignment")
        ver, _, nmsg, refc, size = struct.unpack_from("<BBHII", b, addr)
        _need(ver == 1 and refc == 1, "object header version / reference count")
        _need(size % 8 == 0 and addr + 16 + size <= len(b), "object header size")
        out, p, end = [], addr + 16, addr + 16 + size
        for _ in range(nmsg):
            mtype, msize, flags = struct.unpack_from("<HHB", b, p)
            _need(msize % 8 == 0 and p + 8 + msize <= end, "message size")
            out.append((mtype, b[p + 8:p + 8 + msize]))
            p += 8 + msize
        _need(p == end, "messages fill the header chunk exactly")
        return out

    @staticmethod
    def _dtype(m):
        cls, ver = m[0] & 15, m[0] >> 4
        _need(ver == 1, "datatype version")
        size = struct.unpack_from("<I", m, 4)[0]
        if cls == 0:
            _need(m[1] & 1 == 0, "little-endian")
            off, prec = struct.unpack_from("<HH", m, 8)
            _need(off == 0 and prec == 8 * size, "fixed-point precision")
            return np.dtype(("<i" if m[1] & 8 else "<u") + str(size)), 12
        if cls == 1:
            _need(m[1] & 1 == 0 and (m[1] >> 4) & 3 == 2, "IEEE little-endian, implied mantissa msb")
            off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", m, 8)
            _need(m[2] == 8 * size - 1 and off == 0 and prec == 8 * size, "sign position / precision")
            _need((eloc, esize, mloc, msize, bias) == ((23, 8, 0, 23, 127) if size == 4 else (52, 11, 0, 52, 1023)), "IEEE fields")
            return np.dtype("<f" + str(size)), 20
        if cls == 3:
            _need(m[1] >> 4 == 0, "ASCII character set")
            return np.dtype("S" + str(size)), 8
        raise H5Error(f"datatype class {cls}")

    @staticmethod
    def _space(m):
        ver, rank, flags = m[0], m[1], m[2]
        _need(ver == 1 and flags == 0, "dataspace version / flags")
        return tuple(struct.unpack_from("<Q", m, 8 + 8 * k)[0] for k in range(rank)), 8 + 8 * rank

    def attrs(self, addr):
        out = {}
        for mtype, m in self.messages(addr):
            if mtype != 0x000C:
                continue
            ver, _, nsz, dsz, ssz = struct.unpack_from("<BBHHH", m, 0)
            _need(ver == 1, "attribute version")
            p = 8
            name = m[p:p + nsz]
            _need(name.endswith(b"\0"), "attribute name terminator")
            p += nsz + (-nsz % 8)
            dt, used = self._dtype(m[p:p + dsz])
            _need(used == dsz, "attribute datatype size")
            p += dsz + (-dsz % 8)
            shape, used = self._space(m[p:p + ssz])
            _need(used == ssz, "attribute dataspace size")
            p += ssz + (-ssz % 8)
            cnt = int(np.prod(shape, dtype=np.int64)) if shape else 1
            val = np.frombuffer(m, dt, cnt, p)
            out[name[:-1].decode()] = val.reshape(shape) if shape else val[0]
        return out

    # ---- groups ----
    def _heap_name(self, heap_addr, off):
        b = self.b
        _need(b[heap_addr:heap_addr + 4] == b"HEAP" and b[heap_addr + 4] == 0, "local heap signature / version")
        size, free, seg = struct.unpack_from("<QQQ", b, heap_addr + 8)
        _need(size % 8 == 0 and seg + size <= len(b), "heap data segment")
        _need(free == 1 or free + 16 <= size, "heap free list")
        _need(b[seg] == 0, "heap offset 0 is the empty string")
        _need(off < size and off % 8 == 0, "name offset")
        end = b.index(b"\0", seg + off)
        _need(end < seg + size, "name terminator inside the segment")
        return b[seg + off:end].decode()

    def links(self, entry_or_addr):
        """{name: symbol-table entry} of a group given its entry (or its object header address)"""
        if isinstance(entry_or_addr, dict):
            hdr = entry_or_addr["header"]
        else:
            hdr = entry_or_addr
        stab = [m for t, m in self.messages(hdr) if t == 0x0011]
        _need(len(stab) == 1, "one symbol table message")
        bt, hp = struct.unpack_from("<QQ", stab[0], 0)
        if isinstance(entry_or_addr, dict) and entry_or_addr["cache"] == 1:
            _need((bt, hp) == (entry_or_addr["btree"], entry_or_addr["heap"]), "cached B-tree / heap addresses")
        b = self.b
        _need(b[bt:bt + 4] == b"TREE", "B-tree signature")
        ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, bt + 4)
        _need(ntype == 0 and level == 0 and left == UNDEF and right == UNDEF, "group B-tree leaf-level node")
        _need(used <= 2 * self.int_k and bt + 24 + (4 * self.int_k + 1) * 8 <= len(b), "B-tree node size")
        out, prev_key = {}, None
        for i in range(used):
            key_lo, child, key_hi = struct.unpack_from("<QQQ", b, bt + 24 + 16 * i)
            lo, hi = self._heap_name(hp, key_lo), self._heap_name(hp, key_hi)
            _need(i > 0 or lo == "", "first key is the empty string")
            _need(b[child:child + 4] == b"SNOD" and b[child + 4] == 1, "symbol table node signature / version")
            nsym = struct.unpack_from("<H", b, child + 6)[0]
            _need(1 <= nsym <= 2 * self.leaf_k and child + 8 + 2 * self.leaf_k * 40 <= len(b), "symbol table node size")
            names = []
            for k in range(nsym):
                e = self._entry(child + 8 + 40 * k)
                names.append(self._heap_name(hp, e["name_off"]))
                out[names[-1]] = e
            enc = [n.encode() for n in names]
            _need(enc == sorted(enc), "entries sorted by name")
            _need(lo.encode() < enc[0] and enc[-1] == hi.encode(), "B-tree keys bracket the node (key[i] < names <= key[i+1])")
            _need(prev_key is None or prev_key == lo, "keys chain")
            prev_key = hi
        return out

    def is_group(self, entry):
        return any(t == 0x0011 for t, _ in self.messages(entry["header"]))

    # ---- datasets ----
    def dataset(self, entry):
        msgs = dict()
        for t, m in self.messages(entry["header"]):
            msgs.setdefault(t, m)
        _need({0x0001, 0x0003, 0x0008} <= set(msgs), "dataset needs dataspace, datatype and layout messages")
        shape, _ = self._space(msgs[0x0001])
        dt, _ = self._dtype(msgs[0x0003])
        if 0x0005 in msgs:
            f = msgs[0x0005]
            _need(f[0] == 2 and f[1] in (1, 2, 3) and f[2] in (0, 1, 2), "fill value message")
        ver, cls, addr, size = struct.unpack_from("<BBQQ", msgs[0x0008], 0)
        _need(ver == 3 and cls == 1, "contiguous layout, version 3")
        cnt = int(np.prod(shape, dtype=np.int64)) if shape else 1
        _need(size == cnt * dt.itemsize, "layout size = elements x element size")
        if size == 0:
            return np.zeros(shape, dt)
        _need(addr != UNDEF and addr % 8 == 0 and addr + size <= len(self.b), "data address")
        return np.frombuffer(self.b, dt, cnt, addr).reshape(shape)

    def tree(self, entry=None, prefix=""):
        """{path: ndarray} of every dataset, {path + '@': attrs} of every group / dataset with attributes"""
        entry = self.root_entry if entry is None else entry
        out = {}
        a = self.attrs(entry["header"])
        if a:
            out[prefix + "@"] = a
        for name, e in self.links(entry).items():
            path = prefix + "/" + name
            if self.is_group(e):
                out.update(self.tree(e, path))
            else:
                out[path] = self.dataset(e)
                a = self.attrs(e["header"])
                if a:
                    out[path + "@"] = a
        return out
