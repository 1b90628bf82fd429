"""A minimal pure-Python HDF5 subset: exactly what Keras-3 `model.weights.h5` files need (h5py is not installable in
this image).  Written from the HDF5 File Format Specification, version 0 superblock ("earliest" format, the h5py /
libhdf5 default that Keras' H5IOStore uses):

  read   superblock v0 -> root symbol-table entry -> version-1 object headers (with continuation blocks) ->
         old-style groups (symbol-table message: v1 B-tree of SNOD nodes + local heap) -> datasets with a simple
         dataspace, fixed-point / IEEE float datatype and CONTIGUOUS (or compact) layout;
  write  the same structures: one object header per group / dataset, one single-level B-tree per group (<= 256
         links), names in a local heap with a proper free block, contiguous little-endian raw data.

Not supported (raises NotImplementedError): chunked / compressed datasets, new-style (link-message / fractal-heap)
groups, version-2 object headers, superblock v2/v3, strings, compound types.  Keras writes none of these.

    tree = read_hdf5(bytes_or_path)        # nested dict  {name: dict | numpy.ndarray}
    blob = write_hdf5(tree)                # bytes of a file holding the same tree
"""
from __future__ import annotations

import struct
from typing import Union

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16              # group B-tree parameters of the superblock (libhdf5 defaults)
SNOD_ENTRIES = 2 * LEAF_K
SNOD_SIZE = 8 + SNOD_ENTRIES * 40
BTREE_SIZE = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT, MSG_CONT, MSG_STAB = 0, 1, 3, 4, 5, 8, 0x10, 0x11

Tree = dict


# ------------------------------------------------------------------------------------------------ reader
class _Reader:
    def __init__(self, buf: bytes):
        self.b = memoryview(buf)
        if bytes(self.b[:8]) != SIGNATURE:
            raise ValueError("not an HDF5 file (bad signature)")
        ver = self.b[8]
        if ver != 0 and ver != 1:
            raise NotImplementedError(f"HDF5 superblock version {ver} is not supported (Keras/h5py write version 0)")
        if self.b[13] != 8 or self.b[14] != 8:
            raise NotImplementedError("only 8-byte offsets / lengths are supported")
        off = 24 if ver == 0 else 28          # v1 adds indexed-storage K + 2 reserved bytes
        self.base = self.u64(off)
        root_entry = off + 32
        self.root_ohdr = self.u64(root_entry + 8)

    def u8(self, o): return self.b[o]
    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def u64(self, o): return struct.unpack_from("<Q", self.b, o)[0]

    def messages(self, addr):
        """[(type, flags, offset of the message data, size)] of a version-1 object header, continuations followed."""
        addr += self.base
        if bytes(self.b[addr:addr + 4]) == b"OHDR":
            raise NotImplementedError("version-2 object headers (libver='latest' files) are not supported")
        if self.u8(addr) != 1:
            raise ValueError(f"bad object header version {self.u8(addr)} at {addr}")
        nmsgs, size = self.u16(addr + 2), self.u32(addr + 8)
        blocks, out = [(addr + 16, size)], []
        while blocks and len(out) < nmsgs:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsgs:
                mtype, msize, flags = self.u16(pos), self.u16(pos + 2), self.u8(pos + 4)
                data = pos + 8
                out.append((mtype, flags, data, msize))
                if mtype == MSG_CONT:
                    blocks.append((self.base + self.u64(data), self.u64(data + 8)))
                pos = data + msize
        return out

    def read_object(self, addr):
        msgs = self.messages(addr)
        types = {m[0] for m in msgs}
        if MSG_STAB in types:
            m = next(m for m in msgs if m[0] == MSG_STAB)
            return self.read_group(self.u64(m[2]), self.u64(m[2] + 8))
        if MSG_LAYOUT in types:
            return self.read_dataset(msgs)
        if types & {2, 6}:            # link info / link messages
            raise NotImplementedError("new-style groups (link messages) are not supported")
        return {}

    def read_group(self, btree, heap) -> Tree:
        heap += self.base
        if bytes(self.b[heap:heap + 4]) != b"HEAP":
            raise ValueError("bad local heap signature")
        hdata = self.base + self.u64(heap + 24)
        out: Tree = {}
        for name_off, ohdr in self.btree_entries(btree):
            end = hdata + name_off
            while self.b[end] != 0:
                end += 1
            name = bytes(self.b[hdata + name_off:end]).decode("utf-8")
            out[name] = self.read_object(ohdr)
        return out

    def btree_entries(self, addr):
        addr += self.base
        if bytes(self.b[addr:addr + 4]) != b"TREE":
            raise ValueError("bad B-tree node signature")
        if self.u8(addr + 4) != 0:
            raise ValueError("not a group B-tree node")
        level, used = self.u8(addr + 5), self.u16(addr + 6)
        for i in range(used):
            child = self.u64(addr + 24 + 8 + i * 16)
            if level > 0:
                yield from self.btree_entries(child)
            else:
                s = self.base + child
                if bytes(self.b[s:s + 4]) != b"SNOD":
                    raise ValueError("bad symbol-table node signature")
                for k in range(self.u16(s + 6)):
                    e = s + 8 + 40 * k
                    yield self.u64(e), self.u64(e + 8)

    def read_dataset(self, msgs) -> np.ndarray:
        shape, dtype, layout = None, None, None
        for mtype, _flags, o, size in msgs:
            if mtype == MSG_DATASPACE:
                ver, rank, flags = self.u8(o), self.u8(o + 1), self.u8(o + 2)
                dims_at = o + 8 if ver == 1 else o + 4
                shape = tuple(self.u64(dims_at + 8 * i) for i in range(rank))
            elif mtype == MSG_DATATYPE:
                dtype = self.dtype_of(o)
            elif mtype == MSG_LAYOUT:
                layout = (o, size)
        if shape is None or dtype is None or layout is None:
            raise ValueError("dataset object header lacks dataspace / datatype / layout")
        o, _ = layout
        ver, cls = self.u8(o), self.u8(o + 1)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if ver != 3:
            raise NotImplementedError(f"data layout message version {ver} is not supported")
        if cls == 1:
            daddr, dsize = self.u64(o + 2), self.u64(o + 10)
            if daddr == UNDEF or n == 0:
                return np.zeros(shape, dtype)
            raw = self.b[self.base + daddr:self.base + daddr + dsize]
        elif cls == 0:
            dsize = self.u16(o + 2)
            raw = self.b[o + 4:o + 4 + dsize]
        else:
            raise NotImplementedError("chunked datasets are not supported (Keras weight files are contiguous)")
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()

    def dtype_of(self, o) -> np.dtype:
        cv, b0 = self.u8(o), self.u8(o + 1)
        cls, size = cv & 0x0F, self.u32(o + 4)
        order = ">" if (b0 & 1) else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if (b0 & 8) else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        raise NotImplementedError(f"HDF5 datatype class {cls} is not supported")


def read_hdf5(src: Union[bytes, bytearray, str]) -> Tree:
    """Nested dict {name: dict | ndarray} of every group / dataset of the file."""
    if isinstance(src, (str,)) or hasattr(src, "__fspath__"):
        with open(src, "rb") as f:
            src = f.read()
    r = _Reader(bytes(src))
    return r.read_object(r.root_ohdr)


# ------------------------------------------------------------------------------------------------ writer
def _pad8(n: int) -> int:
    return (n + 7) & ~7


def _msg(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = data + b"\0" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _ohdr(msgs: list) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


def _datatype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (2, 4, 8):
        exp_bits, mant_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[dt.itemsize]
        bits = 8 * dt.itemsize
        # class 1 (float), version 1; bit field: little-endian, mantissa normalisation 2 (implied msb), sign bit position
        head = struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0, dt.itemsize)
        prop = struct.pack("<HHBBBBI", 0, bits, mant_bits, exp_bits, 0, mant_bits, (1 << (exp_bits - 1)) - 1)
        return head + prop
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise NotImplementedError(f"dtype {dt} cannot be written")


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)            # superblock, filled in last

    def alloc(self, data: bytes) -> int:
        addr = _pad8(len(self.buf))
        self.buf.extend(b"\0" * (addr - len(self.buf)))
        self.buf.extend(data)
        return addr

    def dataset(self, arr) -> int:
        a = np.asarray(arr)
        if not a.flags.c_contiguous:         # (np.ascontiguousarray would turn a 0-d scalar into shape (1,))
            a = a.copy(order="C")
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        raw = a.tobytes()
        daddr = self.alloc(raw) if raw else UNDEF
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", d) for d in a.shape)
        msgs = [_msg(MSG_DATASPACE, space, 1), _msg(MSG_DATATYPE, _datatype_msg(a.dtype), 1),
                _msg(MSG_FILL, struct.pack("<BBBBI", 2, 2, 2, 1, 0), 1),
                _msg(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, daddr, len(raw)))]
        return self.alloc(_ohdr(msgs))

    def group(self, tree: Tree):
        """Returns (object header address, B-tree address, heap address)."""
        children = []
        for name in sorted(tree, key=lambda s: s.encode("utf-8")):        # strcmp order, as the B-tree search expects
            v = tree[name]
            if isinstance(v, dict):
                oh, bt, hp = self.group(v)
                children.append((name, oh, 1, bt, hp))
            else:
                children.append((name, self.dataset(v), 0, 0, 0))
        if len(children) > 2 * INTERNAL_K * SNOD_ENTRIES:
            raise NotImplementedError(f"groups with more than {2 * INTERNAL_K * SNOD_ENTRIES} links are not supported")
        # local heap data segment: "" at offset 0, then the names (8-byte aligned), then one free block
        heap_data = bytearray(8)
        offs = []
        for name, *_ in children:
            offs.append(len(heap_data))
            nb = name.encode("utf-8") + b"\0"
            heap_data.extend(nb + b"\0" * (_pad8(len(nb)) - len(nb)))
        free_off = len(heap_data)
        heap_data.extend(struct.pack("<QQ", 1, 32) + b"\0" * 16)       # free block: next = 1 (H5HL_FREE_NULL), size 32
        data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        # symbol-table nodes of <= 8 entries, then the single B-tree node that points at them
        snods, keys = [], [0]
        for i in range(0, len(children), SNOD_ENTRIES):
            part = list(zip(children[i:i + SNOD_ENTRIES], offs[i:i + SNOD_ENTRIES]))
            body = b""
            for (name, oh, is_group, bt, hp), off in part:
                body += struct.pack("<QQII", off, oh, 1 if is_group else 0, 0) + (struct.pack("<QQ", bt, hp) if is_group else b"\0" * 16)
            node = b"SNOD" + struct.pack("<BxH", 1, len(part)) + body
            snods.append(self.alloc(node + b"\0" * (SNOD_SIZE - len(node))))
            keys.append(part[-1][1])
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for k, c in zip(keys, snods):
            node += struct.pack("<QQ", k, c)
        node += struct.pack("<Q", keys[-1])
        btree_addr = self.alloc(node + b"\0" * (BTREE_SIZE - len(node)))
        ohdr_addr = self.alloc(_ohdr([_msg(MSG_STAB, struct.pack("<QQ", btree_addr, heap_addr))]))
        return ohdr_addr, btree_addr, heap_addr

    def finish(self, tree: Tree) -> bytes:
        oh, bt, hp = self.group(tree)
        eof = _pad8(len(self.buf))
        self.buf.extend(b"\0" * (eof - len(self.buf)))
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)      # root symbol-table entry (cached group)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_hdf5(tree: Tree) -> bytes:
    """Serialise a nested dict of numpy arrays (float16/32/64, signed / unsigned integers) into HDF5 bytes."""
    return _Writer().finish(tree)
