"""Minimal HDF5 writer and reader in pure Python / numpy (no libhdf5, no h5py), for the embedding files the three stages of the
reference hand to each other (``extract_embeddings.py:50-55,106-119``, ``inference.py:94-112``, ``TFAM/data/dataset.py:25-73``).

Scope (written from the HDF5 File Format Specification, version 0 structures -- what libhdf5 / h5py emit by default):
  * superblock v0, "old-style" groups (symbol table message -> v1 B-tree of symbol-table nodes + local heap), any depth and any
    number of links per group (multi-level B-trees);
  * object headers v1 with dataspace (v1), datatype (v1: IEEE float 32 / 64, signed / unsigned integers, fixed-length and
    variable-length UTF-8 / ASCII strings), fill value, data layout and attribute (v1) messages, continuation blocks;
  * WRITE: contiguous datasets, and chunked datasets (v1 chunk B-tree, any number of chunks) with the deflate filter -- the
    storage the reference asks for: ``embeddings`` chunked ``(1, D)`` with gzip, ``labels`` contiguous, ``video_ids`` as
    variable-length UTF-8 strings in a global heap collection;
  * READ: contiguous, compact and chunked datasets (v1 chunk B-tree) with the deflate (gzip) and shuffle filters, i.e. the files
    the reference itself writes with the default ``libver``; files written with ``libver='latest'`` (superblock v2+, fractal-heap
    groups) are rejected with a clear error.

Pinning.  No HDF5 library exists in the build image, but a genuine libhdf5-written file does: scipy ships
``scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat`` (a MATLAB v7.3 file = HDF5 behind a 512-byte user block).  The reader
parses it (tests/test_host_cpu.py: the dataset, its values' shape / dtype and its string attribute), which pins the reader's
reading of the specification to the real library for everything that file contains (superblock, symbol-table group, local heap,
v1 object header, dataspace / IEEE-float datatype / fill value / layout / attribute messages, fixed-length string); the writer is
then checked by round trips through that reader.  The chunked + deflate path and the global heap have no library-written
sample in the image: they follow the specification and are checked against each other only ("unpinned").  Host-side I/O only:
nothing here touches the GPU path.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, NODE_K = 4, 16  # symbols per symbol-table node = 2 * LEAF_K, children per B-tree node = 2 * NODE_K
ISTORE_K = 32           # chunk B-tree: children per node = 2 * ISTORE_K (the library default; not stored in a v0 superblock)


# =====================================================================================================================
# writer
# =====================================================================================================================
def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _dtype_msg(dt: np.dtype) -> bytes:
    """Datatype message body (version 1)."""
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        exp_bits, man_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
        head = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)  # class 1; mantissa normalisation = implied msb; sign bit
        prop = struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, (1 << (exp_bits - 1)) - 1)
        return head + prop
    if dt.kind in "iub":
        size = dt.itemsize
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, size)  # class 0; bit 3 = signed
        return head + struct.pack("<HH", 0, size * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x10, 0, 0, dt.itemsize)  # class 3; null-terminated, UTF-8
    raise TypeError(f"unsupported dtype {dt}")


_VLEN_STR = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x10, 0x00, 0, 0, 1) + struct.pack("<HH", 0, 8)
# class 9 (variable length), type = string (1), padding null-terminate, charset UTF-8 (bit 8 -> second bitfield byte = 1);
# base type: 1-byte unsigned integer.  Element in the dataset: length (4) + global heap collection address (8) + object index (4).


def _dataspace_msg(shape) -> bytes:
    shape = tuple(int(s) for s in shape)
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


class _File:
    def __init__(self):
        self.buf = bytearray(96)  # superblock, filled in at the end

    def alloc(self, data: bytes) -> int:
        pad = -len(self.buf) % 8
        self.buf += b"\x00" * pad
        addr = len(self.buf)
        self.buf += data
        return addr


def _attr_value(v):
    """-> (datatype body, dataspace body, raw bytes)"""
    if isinstance(v, (str, bytes)):
        raw = v.encode("utf-8") if isinstance(v, str) else v
        raw = raw + b"\x00"
        return _dtype_msg(np.dtype(f"S{len(raw)}")), _dataspace_msg(()), raw
    arr = np.asarray(v)
    if arr.dtype.kind == "b":
        arr = arr.astype(np.int8)
    if arr.dtype.kind in "US":
        raw = str(arr.item()).encode("utf-8") + b"\x00" if arr.ndim == 0 else None
        if raw is None:
            raise TypeError("string attributes must be scalars")
        return _dtype_msg(np.dtype(f"S{len(raw)}")), _dataspace_msg(()), raw
    shape = arr.shape  # (np.ascontiguousarray would turn a 0-d scalar into shape (1,))
    arr = arr.astype(arr.dtype.newbyteorder("<"))
    return _dtype_msg(arr.dtype), _dataspace_msg(shape), arr.tobytes()


def _attr_messages(attrs: dict) -> list:
    out = []
    for name, v in attrs.items():
        nm = name.encode("utf-8") + b"\x00"
        dt, ds, raw = _attr_value(v)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + raw
        out.append(_message(0x000C, body))
    return out


def _object_header(f: _File, messages: list) -> int:
    body = b"".join(messages)
    head = struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\x00" * 4
    return f.alloc(head + body)


def _chunk_btree(f: _File, chunks: list, rank: int, end_key: tuple) -> int:
    """chunks = [(offsets tuple (rank + 1), nbytes, address)] in row-major order -> address of the root of a v1 chunk B-tree."""
    fan = 2 * ISTORE_K
    level = [(c[0], c[1], c[2]) for c in chunks]  # (first key offsets, first key nbytes, child address)
    depth = 0
    while True:
        nxt, addrs = [], []
        for i in range(0, len(level), fan):
            grp = level[i:i + fan]
            body = b"TREE" + struct.pack("<BBHQQ", 1, depth, len(grp), UNDEF, UNDEF)
            for offs, nbytes, addr in grp:
                body += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", addr)
            nxt_key = level[i + fan][0] if i + fan < len(level) else end_key
            body += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", o) for o in nxt_key)
            ksz = 8 + 8 * (rank + 1)
            body += b"\x00" * ((ksz + 8) * (fan - len(grp)))
            addrs.append(f.alloc(body))
            nxt.append((grp[0][0], grp[0][1], addrs[-1]))
        for i, a in enumerate(addrs):
            f.buf[a + 8:a + 24] = struct.pack("<QQ", addrs[i - 1] if i else UNDEF, addrs[i + 1] if i + 1 < len(addrs) else UNDEF)
        level = nxt
        depth += 1
        if len(level) == 1:
            return level[0][2]


def _write_chunked(f: _File, a: np.ndarray, chunk, gzip_level, attrs: dict) -> int:
    """Chunked dataset with an optional deflate filter: the storage the reference asks h5py for
    (``create_dataset("embeddings", data=..., compression="gzip", chunks=(1, D))``, extract_embeddings.py:107)."""
    rank = a.ndim
    chunk = tuple(int(min(max(c, 1), max(s, 1))) for c, s in zip(chunk, a.shape))
    grid = [range(0, s, c) for s, c in zip(a.shape, chunk)]
    chunks = []
    for idx in np.ndindex(*[len(g) for g in grid]):
        offs = tuple(g[i] for g, i in zip(grid, idx))
        blk = np.zeros(chunk, dtype=a.dtype)
        sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, a.shape))
        blk[tuple(slice(0, x.stop - x.start) for x in sl)] = a[sl]
        raw = blk.tobytes()
        if gzip_level is not None:
            raw = zlib.compress(raw, gzip_level)
        chunks.append((offs + (0,), len(raw), f.alloc(raw)))
    end_key = tuple(((s + c - 1) // c) * c for s, c in zip(a.shape, chunk)) + (0,)
    bt = _chunk_btree(f, chunks, rank, end_key)
    layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", bt) + b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", a.dtype.itemsize)
    msgs = [
        _message(0x0001, _dataspace_msg(a.shape)),
        _message(0x0003, _dtype_msg(a.dtype), flags=1),
        _message(0x0005, struct.pack("<BBBBI", 1, 3, 2, 1, 0), flags=1),  # fill value v1: incremental allocation
    ]
    if gzip_level is not None:  # filter pipeline v1: one filter, deflate (id 1), one client value = the level
        msgs.append(_message(0x000B, struct.pack("<BB", 1, 1) + b"\x00" * 6 + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<II", gzip_level, 0)))
    msgs.append(_message(0x0008, layout))
    return _object_header(f, msgs + _attr_messages(attrs))


def _write_dataset(f: _File, arr, attrs: dict, chunks=None, compression=None) -> int:
    if chunks is not None and isinstance(arr, np.ndarray) and arr.dtype.kind in "fiu" and arr.size > 0 and len(chunks) == arr.ndim:
        a = np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder("<")))
        return _write_chunked(f, a, chunks, 4 if compression in ("gzip", True) else (int(compression) if isinstance(compression, int) else None), attrs)
    if isinstance(arr, (list, tuple)) or (isinstance(arr, np.ndarray) and arr.dtype.kind in "OU"):
        # variable-length UTF-8 strings (h5py.string_dtype(), extract_embeddings.py:118): one global heap collection
        strs = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in np.asarray(arr, dtype=object).ravel()]
        objs = b"".join(struct.pack("<HHIQ", i + 1, 0, 0, len(s)) + _pad8(s) for i, s in enumerate(strs))
        size = 16 + len(objs) + 16  # header + objects + the free-space object (index 0)
        size = max(4096, size + (-size % 8))
        free = size - 16 - len(objs)
        coll = b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, size) + objs + struct.pack("<HHIQ", 0, 0, 0, free - 16) + b"\x00" * (free - 16)
        gaddr = f.alloc(coll)
        raw = b"".join(struct.pack("<IQI", len(s), gaddr, i + 1) for i, s in enumerate(strs))
        dt_body, shape = _VLEN_STR, (len(strs),)
    else:
        a = np.asarray(arr)
        if a.dtype.kind == "b":
            a = a.astype(np.uint8)
        a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<")))
        raw, dt_body, shape = a.tobytes(), _dtype_msg(a.dtype), a.shape
    daddr = f.alloc(raw) if raw else UNDEF
    msgs = [
        _message(0x0001, _dataspace_msg(shape)),
        _message(0x0003, dt_body, flags=1),
        _message(0x0005, struct.pack("<BBBBI", 1, 2, 2, 1, 0), flags=1),  # fill value v1: late allocation, write if set, size 0
        _message(0x0008, struct.pack("<BBQQ", 3, 1, daddr, len(raw)), flags=0),  # layout v3, contiguous
    ] + _attr_messages(attrs)
    return _object_header(f, msgs)


def _write_group(f: _File, node: dict) -> tuple:
    """node = {"attrs": {...}, "children": {name: node | ("dataset", array, attrs)}} -> (object header, B-tree, heap) addresses."""
    entries = []  # (name bytes, object header address, cache type, scratch)
    for name, child in node["children"].items():
        nb = name.encode("utf-8")
        if isinstance(child, tuple):
            storage = child[3] if len(child) > 3 and child[3] else {}
            entries.append((nb, _write_dataset(f, child[1], child[2], storage.get("chunks"), storage.get("compression")), 0, b"\x00" * 16))
        else:
            oh, bt, hp = _write_group(f, child)
            entries.append((nb, oh, 1, struct.pack("<QQ", bt, hp)))
    entries.sort(key=lambda e: e[0])  # strcmp order of the link names
    # local heap: offset 0 = the empty string, then the names, then one free block
    seg = bytearray(8)
    offs = []
    for nb, *_ in entries:
        offs.append(len(seg))
        seg += _pad8(nb + b"\x00")
    free_off = len(seg)
    seg += struct.pack("<QQ", 1, 32) + b"\x00" * 16  # last free block: next = H5HL_FREE_NULL (1), size 32
    seg_addr = f.alloc(bytes(seg))
    heap_addr = f.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(seg), free_off, seg_addr))
    # symbol-table nodes of <= 2 * LEAF_K entries
    level = []  # (address, heap offset of the largest name below)
    per = 2 * LEAF_K
    for i in range(0, max(len(entries), 1), per):
        chunk = list(zip(offs[i:i + per], entries[i:i + per]))
        body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
        for off, (nb, oh, ctype, scratch) in chunk:
            body += struct.pack("<QQII", off, oh, ctype, 0) + scratch
        body += b"\x00" * (40 * (per - len(chunk)))
        level.append((f.alloc(body), chunk[-1][0] if chunk else 0))
    # v1 B-tree over them, as many levels as needed
    depth = 0
    while True:
        nxt = []
        fan = 2 * NODE_K
        groups = [level[i:i + fan] for i in range(0, len(level), fan)]
        addrs = []
        for grp in groups:
            body = b"TREE" + struct.pack("<BBHQQ", 0, depth, len(grp), UNDEF, UNDEF)
            body += struct.pack("<Q", 0)  # key 0: the empty string sorts before every name
            for addr, maxoff in grp:
                body += struct.pack("<QQ", addr, maxoff)
            body += b"\x00" * (16 * (fan - len(grp)))
            addrs.append(f.alloc(body))
            nxt.append((addrs[-1], grp[-1][1]))
        # sibling pointers
        for i, a in enumerate(addrs):
            left = addrs[i - 1] if i > 0 else UNDEF
            right = addrs[i + 1] if i + 1 < len(addrs) else UNDEF
            f.buf[a + 8:a + 24] = struct.pack("<QQ", left, right)
        level = nxt
        depth += 1
        if len(level) == 1:
            break
    btree_addr = level[0][0]
    oh = _object_header(f, [_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + _attr_messages(node.get("attrs", {})))
    return oh, btree_addr, heap_addr


def write_hdf5(path, root: dict) -> None:
    """root = {"attrs": {...}, "children": {name: group dict | ("dataset", array-or-list-of-str, attrs dict[, storage])}} with
    storage = {"chunks": (..), "compression": "gzip" | level | None} for chunked datasets (contiguous otherwise)."""
    f = _File()
    oh, bt, hp = _write_group(f, root)
    eof = len(f.buf) + (-len(f.buf) % 8)
    f.buf += b"\x00" * (eof - len(f.buf))
    sb = SIG + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, NODE_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)
    assert len(sb) == 96
    f.buf[0:96] = sb
    with open(path, "wb") as fh:
        fh.write(bytes(f.buf))


# =====================================================================================================================
# reader
# =====================================================================================================================
class Hdf5FormatError(ValueError):
    pass


class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        base = -1
        off = 0
        while off < len(data):  # the superblock sits at 0, 512, 1024, 2048 ... (user block)
            if data[off:off + 8] == SIG:
                base = off
                break
            off = 512 if off == 0 else off * 2
        if base < 0:
            raise Hdf5FormatError("not an HDF5 file (no superblock signature)")
        ver = data[base + 8]
        if ver > 1:
            raise Hdf5FormatError(f"superblock version {ver} (libver='latest' files: fractal-heap groups) is outside this reader's scope; "
                                  "re-save with the default libver or use h5py")
        if data[base + 13] != 8 or data[base + 14] != 8:
            raise Hdf5FormatError("only 8-byte offsets / lengths are supported")
        p = base + 16
        self.leaf_k, self.node_k = struct.unpack_from("<HH", data, p)
        p += 8 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", data, p)[0]
        if self.base == 0 and base != 0:
            self.base = base
        p += 32
        _, self.root_oh, ctype, _ = struct.unpack_from("<QQII", data, p)

    def at(self, addr: int) -> int:
        return self.base + addr

    # ---- datatypes ----
    def parse_dtype(self, b: bytes):
        cls, ver = b[0] & 0x0F, b[0] >> 4
        f0, f1 = b[1], b[2]
        size = struct.unpack_from("<I", b, 4)[0]
        order = ">" if f0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if f0 & 8 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 9:
            if (f0 & 0x0F) == 1:
                return "vlen_str"
            raise Hdf5FormatError("variable-length sequences are not supported")
        raise Hdf5FormatError(f"datatype class {cls} (version {ver}) is not supported")

    def parse_dataspace(self, b: bytes):
        ver, rank, flags = b[0], b[1], b[2]
        p = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", b, p + 8 * i)[0] for i in range(rank))

    # ---- object headers ----
    def messages(self, addr: int):
        d = self.d
        a = self.at(addr)
        if d[a:a + 4] == b"OHDR":
            raise Hdf5FormatError("version-2 object headers (libver='latest') are outside this reader's scope")
        ver, _, nmsgs, _, hsize = struct.unpack_from("<BBHII", d, a)
        if ver != 1:
            raise Hdf5FormatError(f"object header version {ver}")
        blocks = [(a + 16, hsize)]
        out = []
        while blocks and len(out) < nmsgs:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsgs:
                mtype, msize, mflags = struct.unpack_from("<HHB", d, p)
                body = d[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.at(caddr), clen))
                out.append((mtype, body))
        return out

    def attrs(self, msgs):
        res = {}
        for mtype, b in msgs:
            if mtype != 0x000C:
                continue
            ver, _, nsz, dsz, ssz = struct.unpack_from("<BBHHH", b, 0)
            if ver != 1:
                raise Hdf5FormatError(f"attribute message version {ver}")
            p = 8
            name = b[p:p + nsz].split(b"\x00")[0].decode("utf-8")
            p += nsz + (-nsz % 8)
            dt = self.parse_dtype(b[p:p + dsz])
            p += dsz + (-dsz % 8)
            shape = self.parse_dataspace(b[p:p + ssz])
            p += ssz + (-ssz % 8)
            if isinstance(dt, str):
                ln, gaddr, idx = struct.unpack_from("<IQI", b, p)
                res[name] = self.global_heap_object(gaddr, idx)[:ln].decode("utf-8")
                continue
            n = int(np.prod(shape)) if shape else 1
            arr = np.frombuffer(b[p:p + n * dt.itemsize], dtype=dt).reshape(shape)
            if dt.kind == "S":
                res[name] = arr.item().split(b"\x00")[0].decode("utf-8") if not shape else arr
            else:
                res[name] = arr.item() if not shape else arr.copy()
        return res

    def global_heap_object(self, gaddr: int, idx: int) -> bytes:
        d = self.d
        a = self.at(gaddr)
        if d[a:a + 4] != b"GCOL":
            raise Hdf5FormatError("bad global heap collection")
        size = struct.unpack_from("<Q", d, a + 8)[0]
        p, end = a + 16, a + size
        while p + 16 <= end:
            oi, _, _, osz = struct.unpack_from("<HHIQ", d, p)
            if oi == idx:
                return d[p + 16:p + 16 + osz]
            if oi == 0:
                break
            p += 16 + osz + (-osz % 8)
        raise Hdf5FormatError("global heap object not found")

    # ---- groups ----
    def group_entries(self, btree: int, heap: int):
        d = self.d
        h = self.at(heap)
        if d[h:h + 4] != b"HEAP":
            raise Hdf5FormatError("bad local heap")
        seg = self.at(struct.unpack_from("<Q", d, h + 24)[0])
        out = []

        def name_at(off):
            e = d.index(b"\x00", seg + off)
            return d[seg + off:e].decode("utf-8")

        def walk(addr):
            a = self.at(addr)
            if d[a:a + 4] == b"SNOD":
                n = struct.unpack_from("<H", d, a + 6)[0]
                for i in range(n):
                    off, oh, ctype, _ = struct.unpack_from("<QQII", d, a + 8 + 40 * i)
                    out.append((name_at(off), oh))
                return
            if d[a:a + 4] != b"TREE":
                raise Hdf5FormatError("bad group B-tree node")
            ntype, level, used = struct.unpack_from("<BBH", d, a + 4)
            if ntype != 0:
                raise Hdf5FormatError("expected a group B-tree")
            for i in range(used):
                walk(struct.unpack_from("<Q", d, a + 24 + 8 + 16 * i)[0])

        walk(btree)
        return out

    # ---- datasets ----
    def dataset(self, msgs):
        shape = dt = layout = None
        filters = []
        for mtype, b in msgs:
            if mtype == 0x0001:
                shape = self.parse_dataspace(b)
            elif mtype == 0x0003:
                dt = self.parse_dtype(b)
            elif mtype == 0x0008:
                layout = b
            elif mtype == 0x000B:
                filters = self.parse_filters(b)
        if shape is None or dt is None or layout is None:
            raise Hdf5FormatError("dataset without dataspace / datatype / layout message")
        esize = 16 if isinstance(dt, str) else dt.itemsize
        n = int(np.prod(shape)) if shape else 1
        ver = layout[0]
        if ver in (1, 2):
            rank, cls = layout[1], layout[2]
            addr = struct.unpack_from("<Q", layout, 8)[0] if cls != 0 else None
            dims = struct.unpack_from(f"<{rank}I", layout, 16 if cls != 0 else 8)
            if cls == 1:
                raw = self.d[self.at(addr):self.at(addr) + n * esize] if addr != UNDEF else b"\x00" * (n * esize)
            elif cls == 2:
                raw = self.read_chunked(addr, shape, dims[:-1], esize, filters)
            else:
                csize = struct.unpack_from("<I", layout, 8 + 4 * rank)[0]
                raw = layout[12 + 4 * rank:12 + 4 * rank + csize]
        elif ver == 3:
            cls = layout[1]
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", layout, 2)
                raw = self.d[self.at(addr):self.at(addr) + n * esize] if addr != UNDEF else b"\x00" * (n * esize)
            elif cls == 2:
                rank = layout[2]
                addr = struct.unpack_from("<Q", layout, 3)[0]
                dims = struct.unpack_from(f"<{rank}I", layout, 11)
                raw = self.read_chunked(addr, shape, dims[:-1], esize, filters)
            elif cls == 0:
                size = struct.unpack_from("<H", layout, 2)[0]
                raw = layout[4:4 + size]
            else:
                raise Hdf5FormatError(f"layout class {cls}")
        else:
            raise Hdf5FormatError(f"data layout message version {ver} (libver='latest') is outside this reader's scope")
        if isinstance(dt, str):
            vals = []
            for i in range(n):
                ln, gaddr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self.global_heap_object(gaddr, idx)[:ln].decode("utf-8") if ln else "")
            return np.array(vals, dtype=object).reshape(shape)
        return np.frombuffer(raw, dtype=dt, count=n).reshape(shape).copy()

    @staticmethod
    def parse_filters(b: bytes):
        ver, nf = b[0], b[1]
        if ver != 1:
            raise Hdf5FormatError(f"filter pipeline version {ver}")
        p, out = 8, []
        for _ in range(nf):
            fid, nlen, _, ncd = struct.unpack_from("<HHHH", b, p)
            p += 8 + nlen + (-nlen % 8) + 4 * ncd + (4 if ncd % 2 else 0)
            out.append(fid)
        return out

    def read_chunked(self, btree: int, shape, chunk, esize: int, filters):
        rank = len(shape)
        out = np.zeros(tuple(shape) + (esize,), dtype=np.uint8)
        if btree == UNDEF:
            return out.tobytes()
        d = self.d
        csize = int(np.prod(chunk)) * esize

        def walk(addr):
            a = self.at(addr)
            if d[a:a + 4] != b"TREE":
                raise Hdf5FormatError("bad chunk B-tree node")
            ntype, level, used = struct.unpack_from("<BBH", d, a + 4)
            if ntype != 1:
                raise Hdf5FormatError("expected a chunk B-tree")
            ksz = 8 + 8 * (rank + 1)
            p = a + 24
            for i in range(used):
                nbytes, mask = struct.unpack_from("<II", d, p)
                offs = struct.unpack_from(f"<{rank + 1}Q", d, p + 8)
                child = struct.unpack_from("<Q", d, p + ksz)[0]
                p += ksz + 8
                if level > 0:
                    walk(child)
                    continue
                raw = d[self.at(child):self.at(child) + nbytes]
                for k, fid in reversed(list(enumerate(filters))):
                    if mask & (1 << k):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:  # shuffle: byte planes -> elements
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(esize, -1).T.tobytes()
                    else:
                        raise Hdf5FormatError(f"filter {fid} is not supported (only deflate and shuffle)")
                blk = np.frombuffer(raw[:csize], dtype=np.uint8).reshape(tuple(chunk) + (esize,))
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs[:rank], chunk, shape))
                out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]

        walk(btree)
        return out.tobytes()

    # ---- tree ----
    def read_object(self, oh: int):
        msgs = self.messages(oh)
        stab = [b for t, b in msgs if t == 0x0011]
        if stab:
            bt, hp = struct.unpack_from("<QQ", stab[0], 0)
            node = {"attrs": self.attrs(msgs), "children": {}}
            for name, child in self.group_entries(bt, hp):
                node["children"][name] = self.read_object(child)
            return node
        if any(t == 0x0002 for t, _ in msgs):
            raise Hdf5FormatError("link-info groups (libver='latest') are outside this reader's scope")
        return ("dataset", self.dataset(msgs), self.attrs(msgs))


def read_hdf5(path) -> dict:
    """-> {"attrs": {...}, "children": {name: group dict | ("dataset", ndarray, attrs)}} (the structure write_hdf5 takes)."""
    with open(path, "rb") as fh:
        data = fh.read()
    r = _Reader(data)
    return r.read_object(r.root_oh)
