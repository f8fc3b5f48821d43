"""Minimal pure-Python HDF5 reader (and a tiny writer) — just enough for Keras weight files.

The reference loads its model through ``keras.saving.load_model`` / ``model.load_weights`` (``io.py:386-404``), i.e.
through h5py/libhdf5.  Neither exists in this image, so the subset of the HDF5 file format that h5py produces for Keras
weight files is read here directly:

* superblock versions 0/1 (what h5py writes by default) and 2/3, with an optional user block (signature searched at
  0, 512, 1024, ...);
* object headers version 1 and version 2 (``OHDR`` / ``OCHK``), header continuation blocks;
* groups stored as symbol tables (v1 B-tree + ``SNOD`` nodes + local heap) or as compact link messages; densely stored
  groups (fractal heap) are rejected with a clear error;
* datasets with contiguous, compact or chunked (v1 chunk B-tree) layout, optional deflate and shuffle filters;
* fixed-point and IEEE floating-point datatypes of either byte order; fixed-length strings are returned as bytes.

Attributes are not interpreted.  ``write_h5`` emits the classic layout (superblock 0, symbol-table groups, contiguous
little-endian datasets) and exists for tests and for ``tools``; files from libhdf5 itself are the reader's real target
(``tests/test_keras_weights.py`` parses a libhdf5-written file that ships with scipy's test data).
"""

from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(ValueError):
    pass


class Dataset:
    def __init__(self, f: "H5File", shape, dtype, layout, filters):
        self._f, self.shape, self.dtype, self._layout, self._filters = f, tuple(shape), dtype, layout, filters

    def read(self) -> np.ndarray:
        return self._f._read_dataset(self)


class H5File:
    """``H5File(path_or_bytes)``; ``.datasets()`` -> {"/group/name": Dataset}; ``.read(path)`` -> ndarray."""

    def __init__(self, src):
        self.buf = src if isinstance(src, (bytes, bytearray, memoryview)) else Path(src).read_bytes()
        self.buf = bytes(self.buf)
        self._parse_superblock()
        self._datasets: dict[str, Dataset] | None = None

    # ---- primitive readers ---------------------------------------------------------------------
    def _u(self, off: int, n: int) -> int:
        return int.from_bytes(self.buf[off : off + n], "little")

    def _addr(self, off: int) -> int:
        v = self._u(off, self.so)
        return UNDEF if v == (1 << (8 * self.so)) - 1 else v + self.base

    def _len(self, off: int) -> int:
        return self._u(off, self.sl)

    # ---- superblock ----------------------------------------------------------------------------
    def _parse_superblock(self):
        off = 0
        while True:
            if self.buf[off : off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(self.buf):
                raise Hdf5Error("not an HDF5 file (signature not found)")
        self.sb_off = off
        ver = self.buf[off + 8]
        self.base = 0
        if ver in (0, 1):
            self.so, self.sl = self.buf[off + 13], self.buf[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = self._u(p, self.so)  # every address is relative to it; with a user block libhdf5 stores its size here
            p += 4 * self.so  # base, free-space info, end of file, driver info
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            self.root_oh = self._addr(p + self.so)
        elif ver in (2, 3):
            self.so, self.sl = self.buf[off + 9], self.buf[off + 10]
            p = off + 12
            self.base = self._u(p, self.so)
            self.root_oh = self._addr(p + 3 * self.so)
        else:
            raise Hdf5Error(f"unsupported HDF5 superblock version {ver}")

    # ---- object headers ------------------------------------------------------------------------
    def _messages(self, oh: int):
        """[(type, flags, data_offset, data_size)] of the object header at `oh` (continuations followed)."""
        out = []
        if self.buf[oh : oh + 4] == b"OHDR":
            if self.buf[oh + 4] != 2:
                raise Hdf5Error("unsupported object header version")
            flags = self.buf[oh + 5]
            p = oh + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            n = 1 << (flags & 3)
            size0 = self._u(p, n)
            p += n
            blocks = [(p, size0)]
            track_order = bool(flags & 0x04)
            while blocks:
                p, size = blocks.pop(0)
                end = p + size
                while p + 4 <= end:
                    mtype, msize, mflags = self.buf[p], self._u(p + 1, 2), self.buf[p + 3]
                    p += 4 + (2 if track_order else 0)
                    if mtype == 0x10:
                        coff, clen = self._addr(p), self._len(p + self.so)
                        if self.buf[coff : coff + 4] != b"OCHK":
                            raise Hdf5Error("bad object header continuation block")
                        blocks.append((coff + 4, clen - 8))  # minus signature and checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, p, msize))
                    p += msize
            return out
        ver = self.buf[oh]
        if ver != 1:
            raise Hdf5Error(f"unsupported object header (version byte {ver} at {oh})")
        nmsg = self._u(oh + 2, 2)
        size = self._u(oh + 8, 4)
        blocks = [(oh + 16, size)]
        while blocks and len(out) < nmsg + 64:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end:
                mtype, msize, mflags = self._u(p, 2), self._u(p + 2, 2), self.buf[p + 4]
                p += 8
                if mtype == 0x10:
                    blocks.append((self._addr(p), self._len(p + self.so)))
                elif mtype != 0:
                    out.append((mtype, mflags, p, msize))
                p += msize
        return out

    # ---- groups --------------------------------------------------------------------------------
    def _heap_string(self, heap: int, off: int) -> str:
        if self.buf[heap : heap + 4] != b"HEAP":
            raise Hdf5Error("bad local heap")
        data = self._addr(heap + 8 + 2 * self.sl)
        end = self.buf.index(b"\0", data + off)
        return self.buf[data + off : end].decode("utf-8")

    def _btree_group_entries(self, node: int, heap: int, out: list):
        if self.buf[node : node + 4] == b"SNOD":
            n = self._u(node + 6, 2)
            p = node + 8
            for _ in range(n):
                name = self._heap_string(heap, self._u(p, self.so))
                out.append((name, self._addr(p + self.so)))
                p += 2 * self.so + 24
            return
        if self.buf[node : node + 4] != b"TREE" or self.buf[node + 4] != 0:
            raise Hdf5Error("bad group B-tree node")
        n = self._u(node + 6, 2)
        p = node + 8 + 2 * self.so
        for i in range(n):
            p += self.sl  # key i
            self._btree_group_entries(self._addr(p), heap, out)
            p += self.so

    def _children(self, oh: int):
        """[(name, object header address)] of a group, or None if the object is not a group."""
        msgs = self._messages(oh)
        links = []
        is_group = False
        for mtype, _, p, size in msgs:
            if mtype == 0x11:  # symbol table
                is_group = True
                self._btree_group_entries(self._addr(p), self._addr(p + self.so), links)
            elif mtype == 0x02:  # link info
                is_group = True
                flags = self.buf[p + 1]
                q = p + 2 + (8 if flags & 1 else 0)
                if self._addr(q) != UNDEF:
                    raise Hdf5Error("densely stored groups (fractal heap) are not supported by this reader")
            elif mtype == 0x06:  # link
                is_group = True
                flags = self.buf[p + 1]
                q = p + 2
                ltype = 0
                if flags & 0x08:
                    ltype = self.buf[q]
                    q += 1
                if flags & 0x04:
                    q += 8
                if flags & 0x10:
                    q += 1
                n = 1 << (flags & 3)
                ln = self._u(q, n)
                q += n
                name = self.buf[q : q + ln].decode("utf-8")
                q += ln
                if ltype == 0:
                    links.append((name, self._addr(q)))
        return links if is_group else None

    # ---- datasets ------------------------------------------------------------------------------
    def _dataset(self, oh: int) -> Dataset | None:
        shape = dtype = layout = None
        filters = []
        for mtype, _, p, size in self._messages(oh):
            if mtype == 0x01:
                ver, rank = self.buf[p], self.buf[p + 1]
                q = p + (8 if ver == 1 else 4)
                shape = [self._len(q + i * self.sl) for i in range(rank)]
            elif mtype == 0x03:
                cls, bits0, dsize = self.buf[p] & 0x0F, self.buf[p + 1], self._u(p + 4, 4)
                order = ">" if (bits0 & 1) else "<"
                if cls == 0:
                    dtype = np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{dsize}")
                elif cls == 1:
                    dtype = np.dtype(f"{order}f{dsize}")
                elif cls == 3:
                    dtype = np.dtype(f"S{dsize}")
                else:
                    dtype = ("unsupported", cls, dsize)
            elif mtype == 0x08:
                ver = self.buf[p]
                if ver == 3:
                    cls = self.buf[p + 1]
                    if cls == 0:
                        n = self._u(p + 2, 2)
                        layout = ("compact", p + 4, n)
                    elif cls == 1:
                        layout = ("contiguous", self._addr(p + 2), self._len(p + 2 + self.so))
                    elif cls == 2:
                        nd = self.buf[p + 2]
                        bt = self._addr(p + 3)
                        dims = [self._u(p + 3 + self.so + 4 * i, 4) for i in range(nd)]
                        layout = ("chunked", bt, dims)
                    else:
                        raise Hdf5Error(f"unsupported data layout class {cls}")
                elif ver in (1, 2):
                    nd, cls = self.buf[p + 1], self.buf[p + 2]
                    q = p + 8
                    addr = None
                    if cls != 0:
                        addr = self._addr(q)
                        q += self.so
                    dims = [self._u(q + 4 * i, 4) for i in range(nd)]
                    q += 4 * nd
                    if cls == 1:
                        layout = ("contiguous", addr, None)
                    elif cls == 2:
                        layout = ("chunked", addr, dims + [self._u(q, 4)])
                    else:
                        layout = ("compact", q + 4, self._u(q, 4))
                else:
                    raise Hdf5Error(f"unsupported data layout message version {ver} (written with a newer HDF5 library format)")
            elif mtype == 0x0B:
                ver, nf = self.buf[p], self.buf[p + 1]
                q = p + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self._u(q, 2)
                    if ver == 1 or fid >= 256:
                        nlen = self._u(q + 2, 2)
                        ncd = self._u(q + 6, 2)
                        q += 8 + (((nlen + 7) // 8) * 8 if ver == 1 else nlen)
                    else:
                        ncd = self._u(q + 4, 2)
                        q += 6
                    cd = [self._u(q + 4 * i, 4) for i in range(ncd)]
                    q += 4 * ncd
                    if ver == 1 and ncd % 2:
                        q += 4
                    filters.append((fid, cd))
        if shape is None or dtype is None or layout is None:
            return None
        return Dataset(self, shape, dtype, layout, filters)

    def _chunk_entries(self, node: int, nd: int, out: list):
        if self.buf[node : node + 4] != b"TREE" or self.buf[node + 4] != 1:
            raise Hdf5Error("bad chunk B-tree node")
        level, n = self.buf[node + 5], self._u(node + 6, 2)
        p = node + 8 + 2 * self.so
        ksize = 8 + 8 * nd
        for _ in range(n):
            csize, mask = self._u(p, 4), self._u(p + 4, 4)
            offs = [self._u(p + 8 + 8 * i, 8) for i in range(nd)]
            child = self._addr(p + ksize)
            if level == 0:
                out.append((offs, csize, mask, child))
            else:
                self._chunk_entries(child, nd, out)
            p += ksize + self.so

    def _read_dataset(self, d: Dataset) -> np.ndarray:
        if not isinstance(d.dtype, np.dtype):
            raise Hdf5Error(f"unsupported datatype class {d.dtype[1]}")
        count = int(np.prod(d.shape)) if d.shape else 1
        kind = d._layout[0]
        if kind in ("contiguous", "compact"):
            off = d._layout[1]
            if off == UNDEF:
                return np.zeros(d.shape, d.dtype.newbyteorder("="))
            a = np.frombuffer(self.buf, dtype=d.dtype, count=count, offset=off)
            return a.reshape(d.shape).astype(d.dtype.newbyteorder("="))
        bt, dims = d._layout[1], d._layout[2]
        nd = len(dims)  # rank + 1 (last entry = element size)
        cshape = dims[:-1]
        out = np.zeros(d.shape, d.dtype.newbyteorder("="))
        if bt == UNDEF:
            return out
        entries: list = []
        self._chunk_entries(bt, nd, entries)
        for offs, csize, mask, addr in entries:
            raw = self.buf[addr : addr + csize]
            for i, (fid, cd) in reversed(list(enumerate(d._filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else d.dtype.itemsize
                    n = len(raw) // es
                    raw = np.frombuffer(raw, np.uint8, n * es).reshape(es, n).T.tobytes()
                elif fid == 3:
                    raw = raw[:-4]  # fletcher32 checksum
                else:
                    raise Hdf5Error(f"unsupported HDF5 filter {fid}")
            chunk = np.frombuffer(raw, d.dtype, int(np.prod(cshape))).reshape(cshape)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:-1], cshape, d.shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out

    # ---- public --------------------------------------------------------------------------------
    def datasets(self) -> dict[str, Dataset]:
        if self._datasets is None:
            found: dict[str, Dataset] = {}
            seen = set()

            def walk(oh: int, path: str):
                if oh in seen or oh == UNDEF:
                    return
                seen.add(oh)
                kids = self._children(oh)
                if kids is None:
                    ds = self._dataset(oh)
                    if ds is not None:
                        found[path or "/"] = ds
                    return
                for name, addr in kids:
                    walk(addr, f"{path}/{name}")

            walk(self.root_oh, "")
            self._datasets = found
        return self._datasets

    def read(self, path: str) -> np.ndarray:
        return self.datasets()["/" + path.strip("/")].read()


# -------------------------------------------------------------------------------------------------
# writer (classic layout) — test / tooling support
# -------------------------------------------------------------------------------------------------
def write_h5(path, arrays: dict[str, np.ndarray]) -> None:
    """Write {"/a/b/name": ndarray} as an HDF5 file: superblock 0, symbol-table groups, contiguous little-endian data."""
    tree: dict = {}
    for key, arr in arrays.items():
        node = tree
        parts = [p for p in key.split("/") if p]
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise ValueError(f"{key}: a dataset is used as a group")
        node[parts[-1]] = np.asarray(arr).copy(order="C")   # (ascontiguousarray would turn a 0-d array into 1-d)
    buf = bytearray(96)  # superblock: 8 sig + 16 + 4*8 addresses + 40 root symbol table entry

    def align():
        while len(buf) % 8:
            buf.append(0)

    def alloc(data: bytes) -> int:
        align()
        off = len(buf)
        buf.extend(data)
        return off

    def header(messages: list[tuple[int, bytes]]) -> int:
        body = bytearray()
        for mtype, data in messages:
            data = data + b"\0" * (-len(data) % 8)
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        return alloc(struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + bytes(body))

    def dataset(a: np.ndarray) -> int:
        if a.dtype.kind == "f":
            dt = a.dtype.newbyteorder("<")
            size = dt.itemsize
            # IEEE float: bit field (LE, pad 0, mantissa normalisation = implied msb, sign position), properties
            props = {4: (0, 32, 23, 8, 0, 23, 127), 8: (0, 64, 52, 11, 0, 52, 1023)}[size]
            sign = size * 8 - 1
            dtmsg = struct.pack("<BBBBI", 0x11, 0x20, sign, 0, size) + struct.pack("<HHBBBBI", *props)
        elif a.dtype.kind in "iu":
            dt = a.dtype.newbyteorder("<")
            size = dt.itemsize
            dtmsg = struct.pack("<BBBBI", 0x10, 0x08 if a.dtype.kind == "i" else 0, 0, 0, size) + struct.pack("<HH", 0, size * 8)
        else:
            raise ValueError(f"unsupported dtype {a.dtype}")
        data = a.astype(dt).tobytes()
        daddr = alloc(data) if data else UNDEF
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", s) for s in a.shape)
        layout = struct.pack("<BBQQ", 3, 1, daddr, len(data))
        return header([(0x01, space), (0x03, dtmsg), (0x08, layout)])

    def group(node: dict) -> tuple[int, int, int]:
        """-> (object header, B-tree, heap) addresses"""
        entries = []
        for name in sorted(node):  # symbol table nodes are sorted by name
            child = node[name]
            if isinstance(child, dict):
                oh, bt, hp = group(child)
                entries.append((name, oh, 1, bt, hp))
            else:
                entries.append((name, dataset(child), 0, 0, 0))
        heap_data = bytearray(b"\0" * 8)
        offs = []
        for name, *_ in entries:
            offs.append(len(heap_data))
            heap_data += name.encode() + b"\0"
            heap_data += b"\0" * (-len(heap_data) % 8)
        hd = alloc(bytes(heap_data))
        heap = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, hd))
        # symbol table nodes of at most 2K = 8 entries hang off ONE level-0 B-tree node (group leaf K = 4, internal K = 16)
        if len(entries) > 8 * 32:
            raise ValueError("too many links in one group for this writer")
        snods = []
        for i in range(0, max(len(entries), 1), 8):
            part = entries[i : i + 8]
            body = bytearray(b"SNOD" + struct.pack("<BxH", 1, len(part)))
            for (name, oh, ctype, bt, hp), o in zip(part, offs[i : i + 8]):
                body += struct.pack("<QQI4x", o, oh, ctype) + (struct.pack("<QQ", bt, hp) if ctype == 1 else b"\0" * 16)
            body += b"\0" * (8 + 40 * 8 - len(body))
            snods.append((alloc(bytes(body)), offs[i + len(part) - 1] if part else 0))
        node_b = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF))
        node_b += struct.pack("<Q", 0)
        for addr, last_off in snods:
            node_b += struct.pack("<QQ", addr, last_off)
        node_b += b"\0" * (24 + 8 + 2 * 32 * 16 - len(node_b))
        bt = alloc(bytes(node_b))
        oh = header([(0x11, struct.pack("<QQ", bt, heap))])
        return oh, bt, heap

    root_oh, root_bt, root_heap = group(tree)
    align()
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(buf), UNDEF)
    sb += struct.pack("<QQI4xQQ", 0, root_oh, 1, root_bt, root_heap)
    buf[: len(sb)] = sb
    Path(path).write_bytes(bytes(buf))
