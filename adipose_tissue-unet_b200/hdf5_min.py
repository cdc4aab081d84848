"""Minimal HDF5 reader/writer for Keras weight files (no h5py / libhdf5 in this image).

What the reference does with these files (paths relative to the reference root):
  * writes them with ``net.save_weights('*.weights.h5')`` / ``ModelCheckpoint`` (train_adipose_unet_v3.py:918-922, 984-1053),
  * reads them with ``net.load_weights`` and, on failure, ``hdf5_format.load_weights_from_hdf5_group[_by_name]``
    (full_evaluation_enhanced.py:1266-1301, train_adipose_unet_v3.py:881-916).

Two on-disk layouts exist:
  1. legacy Keras HDF5: ``/[model_weights/]<layer>/<layer>/kernel:0`` and ``bias:0`` plus the attributes
     ``layer_names``, ``backend``, ``keras_version`` on the root and ``weight_names`` on each layer group;
  2. Keras-2.13 "v3" weights-only: ``.../<snake_case_class>[_k]/vars/{0,1}`` groups numbered by creation order.
``read_keras_weights`` accepts both (layout 2 is matched by creation order and tensor shape because the exact
container path depends on the Keras patch level and no genuine file is available here — unpinned);
``write_keras_legacy_weights`` emits layout 1, which the reference's own by-name fallback loads.

Format subset (HDF5 File Format Specification v3): superblock v0/v1 (v2/v3 read-only), v1 object headers with
continuation blocks (v2 read-only), symbol-table groups (B-tree v1 + local heap + SNOD) and compact link-message
groups, contiguous / compact / chunked (B-tree v1, optional shuffle+deflate) datasets, fixed-length string and
IEEE float / integer datatypes, v1-v3 attribute messages.  Dense (fractal-heap) groups, variable-length strings and
external links are not supported and raise ``Hdf5Error``.
"""
from __future__ import annotations

import re
import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


# =====================================================================================================
# Reader
# =====================================================================================================
class _Dataset:
    def __init__(self):
        self.shape: Tuple[int, ...] = ()
        self.dtype: Optional[np.dtype] = None
        self.layout = None          # ("contiguous", addr, size) | ("compact", bytes) | ("chunked", btree, chunk_dims)
        self.filters: List[int] = []
        self.attrs: Dict[str, object] = {}


class Hdf5Reader:
    """Walks every group of a file; ``datasets`` maps the full path to a NumPy array, ``attrs`` maps the
    object path to its attribute dict (string / numeric attributes only)."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        self.datasets: Dict[str, np.ndarray] = {}
        self.attrs: Dict[str, Dict[str, object]] = {}
        self.groups: List[str] = []
        self._parse_superblock()
        self._walk(self.root_addr, "", set())

    # ---- primitives
    def _u(self, off: int, n: int) -> int:
        return int.from_bytes(self.buf[off:off + n], "little")

    def _addr(self, off: int) -> int:
        v = self._u(off, self.so)
        return UNDEF if v == (1 << (8 * self.so)) - 1 else v + self.base

    def _parse_superblock(self):
        b = self.buf
        off = 0
        while True:      # the superblock may sit at 0, 512, 1024, ...
            if b[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off >= len(b):
                raise Hdf5Error("not an HDF5 file (signature not found)")
        ver = b[off + 8]
        self.base = 0
        if ver in (0, 1):
            self.so, self.sl = b[off + 13], b[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = self._u(p, self.so)
            p += 4 * self.so
            # root symbol table entry: name offset, object header address, cache type, reserved, scratch
            self.root_addr = self._addr(p + self.so)
        elif ver in (2, 3):
            self.so, self.sl = b[off + 9], b[off + 10]
            p = off + 12
            self.base = self._u(p, self.so)
            self.root_addr = self._addr(p + 3 * self.so)
        else:
            raise Hdf5Error(f"unsupported superblock version {ver}")

    # ---- object headers
    def _messages(self, addr: int) -> List[Tuple[int, int, bytes]]:
        """(type, flags, body) of every header message of the object at addr (v1 and v2 headers)."""
        b = self.buf
        out = []
        if b[addr:addr + 4] == b"OHDR":
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            csize_len = 1 << (flags & 3)
            chunk0 = self._u(p, csize_len)
            p += csize_len
            track_order = bool(flags & 0x04)
            blocks = [(p, chunk0)]
            i = 0
            while i < len(blocks):
                q, size = blocks[i]
                end = q + size
                while q + 4 <= end:
                    mtype = b[q]
                    msize = self._u(q + 1, 2)
                    mflags = b[q + 3]
                    q += 4 + (2 if track_order else 0)
                    body = b[q:q + msize]
                    if mtype == 0x10:
                        caddr, clen = self._addr(q), self._u(q + self.so, self.sl)
                        blocks.append((caddr + 4, clen - 8))      # skip "OCHK", drop the checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, bytes(body)))
                    q += msize
                i += 1
            return out
        if b[addr] != 1:
            raise Hdf5Error(f"unsupported object header version {b[addr]} at {addr}")
        nmsg = self._u(addr + 2, 2)
        hsize = self._u(addr + 8, 4)
        blocks = [(addr + 16, hsize)]
        i = 0
        while i < len(blocks) and len(out) < nmsg + 64:
            q, size = blocks[i]
            end = q + size
            while q + 8 <= end:
                mtype = self._u(q, 2)
                msize = self._u(q + 2, 2)
                mflags = b[q + 4]
                body = b[q + 8:q + 8 + msize]
                if mtype == 0x10:
                    blocks.append((self._addr(q + 8), self._u(q + 8 + self.so, self.sl)))
                elif mtype != 0:
                    out.append((mtype, mflags, bytes(body)))
                q += 8 + msize
            i += 1
        return out

    # ---- datatype / dataspace / attribute decoding
    def _dtype(self, body: bytes) -> Tuple[Optional[np.dtype], int]:
        cls = body[0] & 0x0F
        bits0 = body[1]
        size = int.from_bytes(body[4:8], "little")
        order = ">" if (bits0 & 1) else "<"
        if cls == 1:
            return np.dtype(f"{order}f{size}"), size
        if cls == 0:
            signed = bool(bits0 & 0x08)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), size
        if cls == 3:
            return np.dtype(f"S{size}"), size
        return None, size      # compound, vlen, reference, ...: skipped by the callers

    @staticmethod
    def _dataspace(body: bytes, sl: int) -> Tuple[int, ...]:
        ver = body[0]
        rank = body[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            p = 4
            if body[3] == 2:          # null dataspace
                return (0,)
        else:
            raise Hdf5Error(f"unsupported dataspace version {ver}")
        return tuple(int.from_bytes(body[p + i * sl:p + (i + 1) * sl], "little") for i in range(rank))

    def _attribute(self, body: bytes):
        ver = body[0]
        nsz, tsz, ssz = (int.from_bytes(body[2:4], "little"), int.from_bytes(body[4:6], "little"),
                         int.from_bytes(body[6:8], "little"))
        p = 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = body[p:p + nsz].split(b"\x00")[0].decode("utf8", "replace"); p += pad(nsz)
        dt, esz = self._dtype(body[p:p + tsz]); p += pad(tsz)
        shape = self._dataspace(body[p:p + ssz], self.sl) if ssz else (); p += pad(ssz)
        if dt is None:
            return name, None
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(body[p:p + n * esz], dtype=dt, count=n).reshape(shape)
        return name, (arr if shape else arr.reshape(())[()])

    # ---- groups
    def _walk(self, addr: int, path: str, seen):
        if addr == UNDEF:
            return
        if not isinstance(seen, dict):       # object address -> first path it was reached by
            seen = {a: None for a in seen}
        if addr in seen:
            # hard link to an object already visited: a dataset is available under this path too (same array), a group
            # is not descended twice (cycles)
            first = seen[addr]
            if first in self.datasets:
                self.datasets[path] = self.datasets[first]
            return
        seen[addr] = path
        msgs = self._messages(addr)
        attrs = {}
        ds = _Dataset()
        links: List[Tuple[str, int]] = []
        is_dataset = False
        for mtype, _flags, body in msgs:
            if mtype == 0x11:                                  # symbol table (old-style group)
                btree = int.from_bytes(body[0:self.so], "little") + self.base
                heap = int.from_bytes(body[self.so:2 * self.so], "little") + self.base
                links += self._symbol_table(btree, heap)
            elif mtype == 0x06:                                # link message (compact new-style group)
                links.append(self._link(body))
            elif mtype == 0x02:                                # link info: dense storage unsupported
                q = 2 + (8 if body[1] & 1 else 0)
                fh = int.from_bytes(body[q:q + self.so], "little")
                if fh != (1 << (8 * self.so)) - 1:
                    raise Hdf5Error(f"group {path or '/'} uses dense link storage (fractal heap): not supported")
            elif mtype == 0x01:
                ds.shape = self._dataspace(body, self.sl); is_dataset = True
            elif mtype == 0x03:
                ds.dtype, _ = self._dtype(body)
            elif mtype == 0x08:
                ds.layout = self._layout(body)
            elif mtype == 0x0B:
                ds.filters = self._filters(body)
            elif mtype == 0x0C:
                try:
                    k, v = self._attribute(body)
                    if v is not None:
                        attrs[k] = v
                except Exception:
                    pass
        self.attrs[path or "/"] = attrs
        if is_dataset and ds.layout is not None:
            if ds.dtype is not None:
                self.datasets[path] = self._read(ds)
            return
        self.groups.append(path or "/")
        for name, child in links:
            if child is not None:
                self._walk(child, f"{path}/{name}", seen)

    def _link(self, body: bytes) -> Tuple[str, Optional[int]]:
        flags = body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]; p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        nlen_size = 1 << (flags & 3)
        nlen = int.from_bytes(body[p:p + nlen_size], "little"); p += nlen_size
        name = body[p:p + nlen].decode("utf8", "replace"); p += nlen
        if ltype != 0:
            return name, None      # soft / external link
        return name, int.from_bytes(body[p:p + self.so], "little") + self.base

    def _heap_name(self, heap: int, off: int) -> str:
        if self.buf[heap:heap + 4] != b"HEAP":
            raise Hdf5Error("bad local heap signature")
        data = self._addr(heap + 8 + 2 * self.sl)
        end = self.buf.index(b"\x00", data + off)
        return self.buf[data + off:end].decode("utf8", "replace")

    def _symbol_table(self, btree: int, heap: int) -> List[Tuple[str, int]]:
        b = self.buf
        out = []
        if b[btree:btree + 4] == b"SNOD":
            n = self._u(btree + 6, 2)
            p = btree + 8
            for _ in range(n):
                out.append((self._heap_name(heap, self._u(p, self.so)), self._addr(p + self.so)))
                p += 2 * self.so + 24
            return out
        if b[btree:btree + 4] != b"TREE":
            raise Hdf5Error("bad B-tree signature in group")
        n = self._u(btree + 6, 2)
        p = btree + 8 + 2 * self.so
        for i in range(n):
            child = self._addr(p + self.sl + i * (self.sl + self.so))
            out += self._symbol_table(child, heap)
        return out

    # ---- datasets
    def _layout(self, body: bytes):
        ver = body[0]
        if ver == 3:
            cls = body[1]
            if cls == 0:
                size = int.from_bytes(body[2:4], "little")
                return ("compact", bytes(body[4:4 + size]))
            if cls == 1:
                a = int.from_bytes(body[2:2 + self.so], "little")
                return ("contiguous", UNDEF if a == (1 << (8 * self.so)) - 1 else a + self.base,
                        int.from_bytes(body[2 + self.so:2 + self.so + self.sl], "little"))
            if cls == 2:
                nd = body[2]
                a = int.from_bytes(body[3:3 + self.so], "little") + self.base
                dims = [int.from_bytes(body[3 + self.so + 4 * i:7 + self.so + 4 * i], "little") for i in range(nd)]
                return ("chunked", a, dims)
        elif ver in (1, 2):
            nd, cls = body[1], body[2]
            p = 8
            a = None
            if cls != 0:
                a = int.from_bytes(body[p:p + self.so], "little") + self.base; p += self.so
            dims = [int.from_bytes(body[p + 4 * i:p + 4 * i + 4], "little") for i in range(nd)]
            if cls == 1:
                return ("contiguous", a, None)
            if cls == 2:
                return ("chunked", a, dims)
        raise Hdf5Error(f"unsupported data layout version {ver}")

    @staticmethod
    def _filters(body: bytes) -> List[int]:
        ver, n = body[0], body[1]
        p = 8 if ver == 1 else 2
        ids = []
        for _ in range(n):
            fid = int.from_bytes(body[p:p + 2], "little")
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(body[p + 2:p + 4], "little"); p += 4
            else:
                nlen = 0; p += 2
            p += 2
            ncd = int.from_bytes(body[p:p + 2], "little"); p += 2
            p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            ids.append(fid)
        return ids

    def _read(self, ds: _Dataset) -> np.ndarray:
        n = int(np.prod(ds.shape)) if ds.shape else 1
        esz = ds.dtype.itemsize
        kind = ds.layout[0]
        if kind == "compact":
            raw = ds.layout[1]
        elif kind == "contiguous":
            a = ds.layout[1]
            raw = b"\x00" * (n * esz) if a == UNDEF else self.buf[a:a + n * esz]
        else:
            return self._read_chunked(ds)
        arr = np.frombuffer(raw, dtype=ds.dtype, count=n).reshape(ds.shape)
        return arr.astype(ds.dtype.newbyteorder("=")) if ds.dtype.kind != "S" else arr.copy()

    def _read_chunked(self, ds: _Dataset) -> np.ndarray:
        _, btree, dims = ds.layout
        rank = len(ds.shape)
        cdims = dims[:rank]
        esz = ds.dtype.itemsize
        out = np.zeros(ds.shape, dtype=ds.dtype.newbyteorder("="))
        for f in ds.filters:
            if f not in (1, 2):
                raise Hdf5Error(f"unsupported filter id {f}")

        def visit(addr):
            b = self.buf
            if b[addr:addr + 4] != b"TREE":
                raise Hdf5Error("bad chunk B-tree signature")
            level = b[addr + 5]
            nent = self._u(addr + 6, 2)
            p = addr + 8 + 2 * self.so
            ksz = 8 + 8 * (rank + 1)
            for i in range(nent):
                k = p + i * (ksz + self.so)
                csize = self._u(k, 4)
                offs = [self._u(k + 8 + 8 * d, 8) for d in range(rank)]
                child = self._addr(k + ksz)
                if level > 0:
                    visit(child)
                    continue
                raw = bytes(b[child:child + csize])
                for f in reversed(ds.filters):
                    if f == 1:
                        raw = zlib.decompress(raw)
                    elif f == 2:      # shuffle
                        a = np.frombuffer(raw, np.uint8).reshape(esz, -1)
                        raw = a.T.tobytes()
                chunk = np.frombuffer(raw, dtype=ds.dtype, count=int(np.prod(cdims))).reshape(cdims)
                sl_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, ds.shape))
                sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
                out[sl_out] = chunk[sl_in]

        visit(btree)
        return out


# =====================================================================================================
# Writer (superblock v0, v1 object headers, symbol-table groups, contiguous datasets, fixed-string attributes)
# =====================================================================================================
LEAF_K = 16          # SNOD capacity 2*LEAF_K = 32 entries
INTERNAL_K = 16      # B-tree node capacity 2*INTERNAL_K = 32 children


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize == 4:
        return struct.pack("<B3BI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "f" and dt.itemsize == 8:
        return struct.pack("<B3BI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<B3BI", 0x10, bits0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<B3BI", 0x13, 0x00, 0, 0, dt.itemsize)      # null-terminated ASCII, fixed length
    raise Hdf5Error(f"cannot write dtype {dt}")


def _dataspace_msg(shape: Tuple[int, ...]) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _attr_msg(name: str, value) -> bytes:
    arr = np.asarray(value)
    if arr.dtype.kind == "U":
        arr = np.char.encode(arr, "utf8")
    if arr.dtype.kind == "S":
        arr = arr.astype(f"S{max(1, arr.dtype.itemsize)}")
    nm = name.encode("utf8") + b"\x00"
    dt = _dtype_msg(arr.dtype)
    sp = _dataspace_msg(arr.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + arr.tobytes()
    return _msg(0x0C, body)


class _Node:
    def __init__(self, name: str):
        self.name = name
        self.children: Dict[str, "_Node"] = {}
        self.data: Optional[np.ndarray] = None
        self.chunks: Optional[Tuple[int, ...]] = None
        self.deflate = False
        self.attrs: Dict[str, object] = {}
        self.addr = 0
        self.btree = 0
        self.heap = 0


class Hdf5Writer:
    """Build a tree of groups/datasets/attributes in memory, then ``save(path)``."""

    def __init__(self):
        self.root = _Node("")

    def _node(self, path: str, create: bool = True) -> _Node:
        n = self.root
        for part in [p for p in path.split("/") if p]:
            if part not in n.children:
                if not create:
                    raise KeyError(path)
                n.children[part] = _Node(part)
            n = n.children[part]
        return n

    def create_group(self, path: str):
        self._node(path)

    def create_dataset(self, path: str, data: np.ndarray, chunks: Optional[Tuple[int, ...]] = None, deflate: bool = False):
        n = self._node(path)
        n.data = np.ascontiguousarray(data)
        n.chunks = tuple(chunks) if chunks else None
        n.deflate = bool(deflate)

    def set_attr(self, path: str, name: str, value):
        self._node(path).attrs[name] = value

    def link(self, existing: str, new_path: str):
        """Hard link: `new_path` names the SAME object as `existing` (one object header, one copy of the data)."""
        node = self._node(existing, create=False)
        parts = [p for p in new_path.split("/") if p]
        parent = self._node("/".join(parts[:-1]))
        parent.children[parts[-1]] = node

    def save(self, path: str):
        out = bytearray(96)      # superblock v0 is 96 bytes with 8-byte offsets

        def alloc(b: bytes) -> int:
            while len(out) % 8:
                out.append(0)
            a = len(out)
            out.extend(b)
            return a

        def header(messages: List[bytes], refs: int = 1) -> bytes:
            body = b"".join(messages)
            return struct.pack("<BxHII4x", 1, len(messages), refs, len(body)) + body

        # object reference counts (hard links name one object from several groups)
        refs: Dict[int, int] = {}

        def count(node: _Node):
            for c in node.children.values():
                refs[id(c)] = refs.get(id(c), 0) + 1
                if refs[id(c)] == 1:
                    count(c)
        count(self.root)
        done = set()

        def emit(node: _Node):
            if id(node) in done:         # reached again through a hard link: already written
                return
            done.add(id(node))
            attr_msgs = [_attr_msg(k, v) for k, v in node.attrs.items()]
            if node.data is not None:
                arr = node.data
                msgs = [_msg(0x01, _dataspace_msg(arr.shape)), _msg(0x03, _dtype_msg(arr.dtype), flags=1),
                        _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0))]
                if node.chunks:      # chunked layout: one level-0 B-tree node over the chunks (test coverage of the reader)
                    import itertools
                    cd, rank, esz = node.chunks, arr.ndim, arr.dtype.itemsize
                    grid = [range(0, s, c) for s, c in zip(arr.shape, cd)]
                    ents = []
                    for offs in itertools.product(*grid):
                        blk = np.zeros(cd, arr.dtype)
                        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cd, arr.shape))
                        blk[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
                        raw = blk.astype(arr.dtype.newbyteorder("<")).tobytes()
                        if node.deflate:
                            raw = zlib.compress(raw, 4)
                        ents.append((offs, len(raw), alloc(raw)))
                    if len(ents) > 64:
                        raise Hdf5Error("too many chunks for this writer")
                    tree = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(ents), UNDEF, UNDEF)
                    for offs, size, a in ents:
                        tree += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
                        tree += struct.pack("<Q", a)
                    tree += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
                    baddr = alloc(tree)
                    if node.deflate:
                        name = _pad8(b"deflate\x00")
                        msgs.append(_msg(0x0B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, len(name), 1, 1) + name +
                                         struct.pack("<II", 4, 0)))
                    msgs.append(_msg(0x08, struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", baddr) +
                                     b"".join(struct.pack("<I", c) for c in cd) + struct.pack("<I", esz)))
                else:
                    daddr = alloc(arr.astype(arr.dtype.newbyteorder("<")).tobytes()) if arr.size else UNDEF
                    msgs.append(_msg(0x08, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes)))
                node.addr = alloc(header(msgs + attr_msgs, refs.get(id(node), 1)))
                return
            for c in node.children.values():
                emit(c)
            # local heap: "" at offset 0, then the child names in sorted order
            names = sorted(node.children)       # bytewise order == strcmp order for ASCII names
            heap_data = bytearray(b"\x00" * 8)
            offs = {}
            for nm in names:
                offs[nm] = len(heap_data)
                heap_data += _pad8(nm.encode("utf8") + b"\x00")
            data_addr = alloc(bytes(heap_data))
            node.heap = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, data_addr))
            # SNODs of up to 2*LEAF_K entries, one level-0 B-tree node over them
            cap = 2 * LEAF_K
            snods, keys = [], [0]
            for i in range(0, max(len(names), 1), cap):
                chunk = names[i:i + cap]
                ents = b""
                for nm in chunk:
                    c = node.children[nm]
                    if c.data is None:
                        ents += struct.pack("<QQII", offs[nm], c.addr, 1, 0) + struct.pack("<QQ", c.btree, c.heap)
                    else:
                        ents += struct.pack("<QQII16x", offs[nm], c.addr, 0, 0)
                ents += b"\x00" * (40 * (cap - len(chunk)))
                snods.append(alloc(b"SNOD" + struct.pack("<BxH", 1, len(chunk)) + ents))
                keys.append(offs[chunk[-1]] if chunk else 0)
            if len(snods) > 2 * INTERNAL_K:
                raise Hdf5Error("too many entries in one group for this writer")
            tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
            for i, s in enumerate(snods):
                tree += struct.pack("<QQ", keys[i], s)
            tree += struct.pack("<Q", keys[len(snods)])
            tree += b"\x00" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(tree))
            node.btree = alloc(tree)
            node.addr = alloc(header([_msg(0x11, struct.pack("<QQ", node.btree, node.heap))] + attr_msgs, refs.get(id(node), 1)))

        emit(self.root)
        while len(out) % 8:
            out.append(0)
        sb = SIGNATURE + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(out), UNDEF)
        sb += struct.pack("<QQII", 0, self.root.addr, 1, 0) + struct.pack("<QQ", self.root.btree, self.root.heap)
        assert len(sb) == 96
        out[0:96] = sb
        with open(path, "wb") as f:
            f.write(bytes(out))


# =====================================================================================================
# Keras layouts
# =====================================================================================================
def write_keras_legacy_weights(path: str, weights: Dict[str, np.ndarray], layer_names: Optional[List[str]] = None,
                               keras_version: str = "2.13.1", backend: str = "tensorflow"):
    """Legacy Keras HDF5 weights file (hdf5_format.save_weights_to_hdf5_group): loadable by the reference's
    load_legacy_weights fallback (full_evaluation_enhanced.py:1285-1301)."""
    if layer_names is None:
        layer_names = []
        for k in weights:
            n = k.rsplit("/", 1)[0]
            if n not in layer_names:
                layer_names.append(n)
    w = Hdf5Writer()
    w.set_attr("/", "layer_names", np.array([n.encode("utf8") for n in layer_names]))
    w.set_attr("/", "backend", np.bytes_(backend.encode("utf8")))
    w.set_attr("/", "keras_version", np.bytes_(keras_version.encode("utf8")))
    for n in layer_names:
        names = [f"{n}/kernel:0".encode(), f"{n}/bias:0".encode()]
        w.create_group(f"/{n}")
        w.set_attr(f"/{n}", "weight_names", np.array(names))
        w.create_dataset(f"/{n}/{n}/kernel:0", np.asarray(weights[n + "/kernel"], np.float32))
        w.create_dataset(f"/{n}/{n}/bias:0", np.asarray(weights[n + "/bias"], np.float32))
    w.save(path)


def write_keras_v3_like_weights(path: str, weights: Dict[str, np.ndarray], layer_names: List[str],
                                container: str = "layers"):
    """Synthetic Keras-2.13 "v3" layout (``<container>/conv2d[_k]/vars/{0,1}`` in creation order).  UNPINNED: used by
    the tests of the tolerant reader only; real checkpoints are written in the legacy layout."""
    w = Hdf5Writer()
    for i, n in enumerate(layer_names):
        g = f"/{container}/conv2d" + (f"_{i}" if i else "")
        w.create_dataset(f"{g}/vars/0", np.asarray(weights[n + "/kernel"], np.float32))
        w.create_dataset(f"{g}/vars/1", np.asarray(weights[n + "/bias"], np.float32))
    w.create_group("/optimizer/vars")
    w.save(path)


# model.layers of the reference graph in creation order (train_adipose_unet_v3.py:664-750): (Keras class, layer name or None)
def _reference_layer_classes(deep_supervision: bool) -> List[Tuple[str, Optional[str]]]:
    L: List[Tuple[str, Optional[str]]] = [("input_layer", None), ("reshape", None)]
    for lvl in (1, 2, 3):
        L += [("conv2d", f"down{lvl}_conv1"), ("conv2d", f"down{lvl}_conv2"), ("max_pooling2d", None)]
    L += [("conv2d", "dilate1"), ("dropout", None)] + [("conv2d", f"dilate{i}") for i in range(2, 7)] + [("add", None)]
    for lvl in (3, 2, 1):
        L += [("up_sampling2d", None), ("conv2d", f"up{lvl}_conv1"), ("concatenate", None), ("conv2d", f"up{lvl}_conv2"),
              ("conv2d", f"up{lvl}_conv3"), ("dropout", None)]
    L += [("conv2d", "output_softmax")]
    if deep_supervision:
        L += [("conv2d", "aux_out1"), ("lambda", None), ("conv2d", "aux_out2"), ("lambda", None), ("lambda", None), ("lambda", None)]
    L += [("lambda", None), ("lambda", None)]
    return L


def write_keras_weights_hybrid(path: str, weights: Dict[str, np.ndarray], keras_version: str = "2.13.1"):
    """`*.weights.h5` as written by this repo: ONE file that every loader of the reference accepts.

    * legacy part (root attrs `layer_names` / per-layer `weight_names`, datasets `/<layer>/<layer>/kernel:0`): what
      `hdf5_format.load_weights_from_hdf5_group[_by_name]` reads (train_adipose_unet_v3.py:881-916,
      full_evaluation_enhanced.py:1285-1301);
    * Keras-2.13 saving_lib part (`save_weights` / `ModelCheckpoint` on a `.weights.h5` name, train:918-922, 1271-1278): root
      `vars`, then per layer of `model.layers` a group `<container>/<snake_case class>[_k]/vars/{0,1}` numbered per class in
      creation order (kernel = 0, bias = 1; weightless layers get empty `vars` groups).  The container is `layers` in Keras
      2.13+ and `_layer_checkpoint_dependencies` in 2.12 (Keras' own reader maps one to the other): both are written.
      With deep supervision the 1x1 heads follow `output_softmax` (conv2d_21), so that the plain inference graph
      (segmentation_inference.py:150) finds its 22 convolutions under the indices it asks for.
    The tensors exist once: the saving_lib paths are HDF5 hard links to the legacy datasets (the file is no larger).
    UNPINNED against a genuine TF-2.13 file (none is shipped with the reference, h5py / TF are not installable here)."""
    names = []
    for k in weights:
        n = k.rsplit("/", 1)[0]
        if n not in names:
            names.append(n)
    w = Hdf5Writer()
    w.set_attr("/", "layer_names", np.array([n.encode("utf8") for n in names]))
    w.set_attr("/", "backend", np.bytes_(b"tensorflow"))
    w.set_attr("/", "keras_version", np.bytes_(keras_version.encode("utf8")))
    for n in names:
        w.create_group(f"/{n}")
        w.set_attr(f"/{n}", "weight_names", np.array([f"{n}/kernel:0".encode(), f"{n}/bias:0".encode()]))
        w.create_dataset(f"/{n}/{n}/kernel:0", np.asarray(weights[n + "/kernel"], np.float32))
        w.create_dataset(f"/{n}/{n}/bias:0", np.asarray(weights[n + "/bias"], np.float32))
    w.create_group("/vars")
    ds = "aux_out1" in names
    for container in ("layers", "_layer_checkpoint_dependencies"):
        counter: Dict[str, int] = {}
        for cls, lname in _reference_layer_classes(ds):
            k = counter.get(cls, 0)
            counter[cls] = k + 1
            g = f"/{container}/{cls}" + (f"_{k}" if k else "")
            w.create_group(g + "/vars")
            if lname is not None:
                if lname not in names:
                    raise Hdf5Error(f"weights of layer {lname} missing")
                w.link(f"/{lname}/{lname}/kernel:0", g + "/vars/0")
                w.link(f"/{lname}/{lname}/bias:0", g + "/vars/1")
    w.save(path)


_V3_RE = re.compile(r"^(?P<prefix>.*)/(?P<cls>[A-Za-z0-9_]*?)(?:_(?P<idx>\d+))?/vars/(?P<var>\d+)$")


def read_keras_weights(path: str, init_nb: int = 44, layout: str = "auto") -> Dict[str, np.ndarray]:
    """'<layer>/kernel' (HWIO) and '<layer>/bias' for the 22 conv layers (+ aux heads when present).
    layout: 'auto' (by-name datasets first, then saving_lib `vars` groups) or 'vars' (saving_lib groups only)."""
    from .layers import conv_layers
    r = Hdf5Reader(path)
    ds = r.datasets
    layers = conv_layers(init_nb)
    out: Dict[str, np.ndarray] = {}
    # layout 1: by name
    by_name = layout != "vars"
    for name, *_ in (layers if by_name else []):
        k = [p for p in ds if p.endswith(f"/{name}/kernel:0") or p.endswith(f"/{name}/kernel")]
        b = [p for p in ds if p.endswith(f"/{name}/bias:0") or p.endswith(f"/{name}/bias")]
        if not k or not b:
            by_name = False
            break
        out[name + "/kernel"] = np.asarray(ds[min(k, key=len)], np.float32)
        out[name + "/bias"] = np.asarray(ds[min(b, key=len)], np.float32)
    if by_name:
        for aux in ("aux_out1", "aux_out2"):
            k = [p for p in ds if p.endswith(f"/{aux}/kernel:0")]
            b = [p for p in ds if p.endswith(f"/{aux}/bias:0")]
            if k and b:
                out[aux + "/kernel"] = np.asarray(ds[k[0]], np.float32); out[aux + "/bias"] = np.asarray(ds[b[0]], np.float32)
        return out
    # layout 2: '<prefix>/<class>[_k]/vars/<i>' in creation order, optimizer state ignored
    groups: Dict[Tuple[str, str, int], Dict[int, np.ndarray]] = {}
    prefixes = sorted({m.group("prefix") for m in (_V3_RE.match(p) for p in ds) if m and "optimizer" not in m.group("prefix")})
    for p, a in ds.items():
        m = _V3_RE.match(p)
        if not m or "optimizer" in m.group("prefix"):
            continue
        if len(prefixes) > 1 and m.group("prefix") != ("/layers" if "/layers" in prefixes else prefixes[0]):
            continue        # a hybrid file names the same variables under two containers: read one
        key = (m.group("prefix"), m.group("cls"), int(m.group("idx") or 0))
        groups.setdefault(key, {})[int(m.group("var"))] = a
    convs = [(k, v) for k, v in sorted(groups.items()) if 0 in v and 1 in v and v[0].ndim == 4 and v[1].ndim == 1]
    if not convs:
        raise Hdf5Error(f"{path}: neither '<layer>/kernel:0' datasets nor '<class>/vars/<i>' groups found "
                        f"(datasets: {sorted(ds)[:8]} ...)")
    out = {}
    three = [(k, v) for k, v in convs if v[0].shape[0] == 3]
    want3 = [l for l in layers if l[3] == 3]
    if len(three) != len(want3):
        raise Hdf5Error(f"{path}: expected {len(want3)} 3x3 conv variable groups, found {len(three)}")
    for (name, ci, co, k, _), (_, v) in zip(want3, three):
        if v[0].shape != (3, 3, ci, co):
            raise Hdf5Error(f"{path}: variable group order does not match the graph at {name}: {v[0].shape}")
        out[name + "/kernel"] = np.asarray(v[0], np.float32); out[name + "/bias"] = np.asarray(v[1], np.float32)
    by_shape = {(1, 1, init_nb, 2): "output_softmax", (1, 1, 4 * init_nb, 1): "aux_out1", (1, 1, 2 * init_nb, 1): "aux_out2"}
    for _, v in convs:
        nm = by_shape.get(tuple(v[0].shape))
        if nm:
            out[nm + "/kernel"] = np.asarray(v[0], np.float32); out[nm + "/bias"] = np.asarray(v[1], np.float32)
    if "output_softmax/kernel" not in out:
        raise Hdf5Error(f"{path}: 1x1 softmax head not found")
    return out
