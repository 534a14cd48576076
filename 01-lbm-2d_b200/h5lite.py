"""h5lite -- a dependency-free writer (and reader) for the subset of HDF5 the case file needs.

Why: the reference's case file is HDF5 (`io/lbm_writer.py:69-133, 212-251`, written through h5py), and h5py / libhdf5 are
absent from this image.  Rather than fall back to a private container, `DeviceLBMCaseWriter` writes a real HDF5 file
itself when h5py is missing.  The subset is the oldest, most widely readable on-disk format ("HDF5 File Format
Specification" version 1.1: superblock version 0, symbol-table groups, version-1 object headers and B-trees), i.e. what
every libhdf5 since 1.6 -- and therefore h5py, h5dump, MATLAB, netCDF-4 -- opens:

* contiguous datasets (`create_dataset`): `static_mask`, `mean_vel_field`, `mean_vel_sq_field`, `sum_vor`;
* ONE-frame-per-chunk extensible datasets (`create_appendable`): `turbulence`, shape (T, 9, H, W), maxshape
  (None, 9, H, W), chunks (1, 9, H, W) exactly as the reference declares it; every appended frame goes straight to the end
  of the file (streaming: host memory stays bounded, the chunk index -- a version-1 B-tree -- is written at close);
* root attributes: numeric arrays (`stats_min/max/mean`) and variable-length UTF-8 strings (`config_json`, stored like
  h5py stores a Python str: a global-heap object).

Compression: `compression: gzip` (HDF5's standard deflate filter) is honoured per frame; the reference's default `lzf` is
an h5py-only plug-in filter (id 32000, not part of HDF5), so such frames are stored uncompressed, which any reader
handles.  Not written: sub-groups, anything the case file does not use.

The reader (`read`) is written from the same specification and exists for `read_case()` on hosts without h5py and for the
tests: it is pinned on a file produced by libhdf5 itself (a MATLAB 7.3 file shipped with scipy's test data), so the
structures both sides share -- superblock, symbol table, local heap, B-tree, object header, dataspace / datatype / layout
messages, attributes -- are checked against the real library's output, not only against this writer.
"""
from __future__ import annotations

import json
import os
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
GROUP_LEAF_K, GROUP_INTERNAL_K, CHUNK_K = 4, 16, 32   # the library defaults a version-0 superblock implies
DATA_START = 512

MSG_DATASPACE, MSG_DATATYPE, MSG_FILL, MSG_LAYOUT, MSG_FILTERS, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMTAB = \
    0x0001, 0x0003, 0x0005, 0x0008, 0x000B, 0x000C, 0x0010, 0x0011


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------------ writer: messages
def _dt_message(dtype: np.dtype) -> bytes:
    dtype = np.dtype(dtype)
    if dtype.byteorder == ">":
        raise ValueError("big-endian arrays are not supported")
    if dtype.kind == "f" and dtype.itemsize in (4, 8):
        if dtype.itemsize == 4:
            sign, props = 31, struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        else:
            sign, props = 63, struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        # class 1 (floating point), version 1; little endian, mantissa normalisation "msb implied" (bits 4-5 = 2)
        return struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dtype.itemsize) + props
    if dtype.kind in "iu" and dtype.itemsize in (1, 2, 4, 8):
        return struct.pack("<BBBBI", 0x10, 0x08 if dtype.kind == "i" else 0x00, 0, 0, dtype.itemsize) + \
            struct.pack("<HH", 0, dtype.itemsize * 8)
    raise TypeError(f"h5lite cannot store dtype {dtype}")


_DT_VLEN_UTF8 = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0x00, 16) + struct.pack("<BBBBI", 0x13, 0, 0, 0, 1)


def _space_message(shape, maxshape=None) -> bytes:
    out = struct.pack("<BBB5x", 1, len(shape), 0 if maxshape is None else 1)
    out += b"".join(struct.pack("<Q", int(d)) for d in shape)
    if maxshape is not None:
        out += b"".join(struct.pack("<Q", UNDEF if d is None else int(d)) for d in maxshape)
    return out


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages) -> bytes:
    blob = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(blob)) + blob


def _attribute_message(name: str, dt: bytes, space: bytes, data: bytes) -> bytes:
    nm = name.encode("utf-8") + b"\0"
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(space)) + _pad8(nm) + _pad8(dt) + _pad8(space) + data
    if len(body) > 65000:
        raise ValueError(f"attribute {name!r} is too large for a version-1 object header message")
    return _message(MSG_ATTRIBUTE, body)


class _Appendable:
    def __init__(self, owner, name, frame_shape, dtype, gzip=None):
        self.owner, self.name = owner, name
        self.frame_shape, self.dtype = tuple(int(d) for d in frame_shape), np.dtype(dtype)
        self.frame_bytes = int(np.prod(self.frame_shape)) * self.dtype.itemsize
        self.gzip = None if gzip is None else int(gzip)    # deflate level: HDF5's standard filter 1
        self.addresses, self.sizes = [], []

    def append(self, frame):
        a = np.ascontiguousarray(frame, self.dtype)
        if a.shape != self.frame_shape:
            raise ValueError(f"{self.name}: frame of shape {a.shape}, expected {self.frame_shape}")
        if self.gzip is not None:
            a = np.frombuffer(zlib.compress(a.tobytes(), self.gzip), np.uint8)
        self.addresses.append(self.owner._write_raw(a))
        self.sizes.append(a.nbytes)

    def __len__(self):
        return len(self.addresses)


class Writer:
    """Streaming writer.  Raw data go to the file as they arrive; all metadata are written by `close()`."""

    def __init__(self, path):
        self.path = path
        self.fh = open(path, "wb")
        # a superblock without a root group yet: a killed run leaves a file no reader mistakes for a finished case
        self.fh.write(self._superblock(UNDEF, UNDEF, UNDEF, DATA_START).ljust(DATA_START, b"\0"))
        self.pos = DATA_START
        self.datasets = {}     # name -> ("contiguous", shape, dtype, address, nbytes) | ("chunked", _Appendable)
        self.attrs = {}
        self.closed = False

    # -- raw data ---------------------------------------------------------------------------------------------
    def _align(self):
        pad = -self.pos % 8
        if pad:
            self.fh.write(b"\0" * pad)
            self.pos += pad

    def _write_raw(self, arr) -> int:
        self._align()
        addr = self.pos
        arr.tofile(self.fh)
        self.pos += arr.nbytes
        return addr

    def _write_meta(self, blob: bytes) -> int:
        self._align()
        addr = self.pos
        self.fh.write(blob)
        self.pos += len(blob)
        return addr

    def create_dataset(self, name, data, dtype=None):
        self._check_name(name)
        a = np.asarray(data, dtype, order="C")   # (np.ascontiguousarray would turn a scalar into shape (1,))
        _dt_message(a.dtype)
        self.datasets[name] = ("contiguous", a.shape, a.dtype, self._write_raw(a) if a.nbytes else UNDEF, a.nbytes)

    def create_appendable(self, name, frame_shape, dtype="f4", gzip=None) -> _Appendable:
        """`gzip`: deflate level 0-9 per frame (the filter h5py calls compression="gzip"); None = stored as is."""
        self._check_name(name)
        _dt_message(np.dtype(dtype))
        ap = _Appendable(self, name, frame_shape, dtype, gzip)
        self.datasets[name] = ("chunked", ap)
        return ap

    def set_attr(self, name, value):
        self.attrs[name] = value

    def flush(self):
        self.fh.flush()

    def _check_name(self, name):
        if self.closed:
            raise ValueError("the file is closed")
        if not name or "/" in name or name in self.datasets:
            raise ValueError(f"bad or duplicate dataset name {name!r}")

    # -- metadata ---------------------------------------------------------------------------------------------
    @staticmethod
    def _superblock(root_header, btree, heap, eof) -> bytes:
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        # root group symbol table entry; cache type 1: the scratch pad holds the group's B-tree and heap addresses
        sb += struct.pack("<QQII", 0, root_header, 1 if root_header != UNDEF else 0, 0) + struct.pack("<QQ", btree, heap)
        return sb

    def _chunk_index(self, ap: _Appendable) -> int:
        """Version-1 B-tree (node type 1) over the chunks, built bottom-up; returns the root node's address."""
        rank1 = len(ap.frame_shape) + 2      # dataset rank + 1 (the element-size pseudo dimension)
        key_size = 8 + 8 * rank1
        node_size = 24 + (2 * CHUNK_K + 1) * key_size + 2 * CHUNK_K * 8

        def key(frame_index, nbytes):
            return struct.pack("<II", nbytes, 0) + struct.pack("<Q", frame_index) + b"\0" * (8 * (rank1 - 1))

        n = len(ap)
        # level 0: (first frame index, child address) per entry
        entries = [(i, ap.addresses[i], ap.sizes[i]) for i in range(n)]
        level = 0
        while True:
            groups = [entries[i:i + 2 * CHUNK_K] for i in range(0, len(entries), 2 * CHUNK_K)] or [[]]
            self._align()
            addrs = [self.pos + j * node_size for j in range(len(groups))]   # node_size is a multiple of 8
            nxt = []
            for j, g in enumerate(groups):
                left = addrs[j - 1] if j > 0 else UNDEF
                right = addrs[j + 1] if j + 1 < len(groups) else UNDEF
                blob = b"TREE" + struct.pack("<BBH", 1, level, len(g)) + struct.pack("<QQ", left, right)
                for first, child, nbytes in g:
                    blob += key(first, nbytes) + struct.pack("<Q", child)
                # the closing key: the first chunk of the next node, or one past the last frame
                end = groups[j + 1][0][0] if j + 1 < len(groups) else n
                blob += key(end, groups[j + 1][0][2] if j + 1 < len(groups) else 0)
                self._write_meta(blob.ljust(node_size, b"\0"))
                nxt.append((g[0][0] if g else 0, addrs[j], g[0][2] if g else 0))
            if len(nxt) == 1:
                return nxt[0][1]
            entries, level = nxt, level + 1

    def _dataset_header(self, name) -> bytes:
        d, extra = self.datasets[name], []
        if d[0] == "contiguous":
            _, shape, dtype, addr, nbytes = d
            space = _space_message(shape)
            layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", addr, nbytes)
            fill = struct.pack("<BBBBI", 2, 2, 2, 1, 0)      # allocate late, fill if set, the default (zero) fill value
        else:
            ap = d[1]
            dtype = ap.dtype
            space = _space_message((len(ap),) + ap.frame_shape, (None,) + ap.frame_shape)
            dims = (1,) + ap.frame_shape + (dtype.itemsize,)
            layout = struct.pack("<BBB", 3, 2, len(dims)) + struct.pack("<Q", self._chunk_index(ap)) + \
                b"".join(struct.pack("<I", v) for v in dims)
            fill = struct.pack("<BBBBI", 2, 3, 2, 1, 0)      # allocate incrementally
            if ap.gzip is not None:   # version-1 filter pipeline: one filter, id 1 (deflate), one client value (the level) + padding
                extra = [_message(MSG_FILTERS, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I4x", ap.gzip), 1)]
        return _object_header([_message(MSG_DATASPACE, space), _message(MSG_DATATYPE, _dt_message(dtype), 1),
                               _message(MSG_FILL, fill), _message(MSG_LAYOUT, layout)] + extra)

    def _attribute(self, name, value, strings) -> bytes:
        if isinstance(value, (str, bytes)):
            raw = value.encode("utf-8") if isinstance(value, str) else value
            strings.append(raw)          # global heap object index = position + 1; the address is patched in by close()
            return name, raw, len(strings)
        a = np.asarray(value, order="C")
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        return _attribute_message(name, _dt_message(a.dtype), _space_message(a.shape), a.tobytes())

    def close(self):
        if self.closed:
            return
        self.closed = True
        names = sorted(self.datasets, key=lambda s: s.encode("utf-8"))   # symbol table order: strcmp
        headers = {nm: self._write_meta(self._dataset_header(nm)) for nm in names}

        # variable-length strings of the attributes: one global heap collection
        strings, attr_msgs = [], []
        pending = [self._attribute(k, v, strings) for k, v in self.attrs.items()]
        gcol_addr = UNDEF
        if strings:
            body = b""
            for idx, raw in enumerate(strings, start=1):
                body += struct.pack("<HH4xQ", idx, 0, len(raw)) + _pad8(raw)
            size = max(4096, (16 + len(body) + 16 + 7) // 8 * 8)
            free = size - 16 - len(body)
            blob = b"GCOL" + struct.pack("<B3xQ", 1, size) + body + struct.pack("<HH4xQ", 0, 0, free)
            gcol_addr = self._write_meta(blob.ljust(size, b"\0"))
        for item in pending:
            if isinstance(item, tuple):
                nm, raw, idx = item
                item = _attribute_message(nm, _DT_VLEN_UTF8, _space_message(()), struct.pack("<IQI", len(raw), gcol_addr, idx))
            attr_msgs.append(item)

        # local heap with the link names; offset 0 is the empty string every group B-tree's first key points at
        seg, offsets = b"\0" * 8, {}
        for nm in names:
            offsets[nm] = len(seg)
            seg += _pad8(nm.encode("utf-8") + b"\0")
        heap_addr = self._write_meta(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), 1, 0))   # free list: none (1)
        seg_addr = self._write_meta(seg)
        self.fh.seek(heap_addr + 24)
        self.fh.write(struct.pack("<Q", seg_addr))
        self.fh.seek(self.pos)

        # symbol table nodes (<= 2 * leaf K entries each) under one B-tree node
        per = 2 * GROUP_LEAF_K
        groups = [names[i:i + per] for i in range(0, len(names), per)]
        if len(groups) > 2 * GROUP_INTERNAL_K:
            raise ValueError("too many datasets for a single-level group B-tree")
        snod_size = 8 + per * 40
        tree = b"TREE" + struct.pack("<BBH", 0, 0, len(groups)) + struct.pack("<QQ", UNDEF, UNDEF) + struct.pack("<Q", 0)
        for g in groups:
            blob = b"SNOD" + struct.pack("<BxH", 1, len(g))
            for nm in g:
                blob += struct.pack("<QQII16x", offsets[nm], headers[nm], 0, 0)
            tree += struct.pack("<QQ", self._write_meta(blob.ljust(snod_size, b"\0")), offsets[g[-1]])
        tree_size = 24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8
        btree_addr = self._write_meta(tree.ljust(tree_size, b"\0"))

        root = self._write_meta(_object_header([_message(MSG_SYMTAB, struct.pack("<QQ", btree_addr, heap_addr))] + attr_msgs))
        self._align()
        self.fh.seek(0)
        self.fh.write(self._superblock(root, btree_addr, heap_addr, self.pos))
        self.fh.close()

    def abort(self):
        if not self.closed:
            self.closed = True
            self.fh.close()


# ------------------------------------------------------------------------------------------------------- reader
class _Reader:
    def __init__(self, path):
        # the file is mapped, not read: a case file is mostly `turbulence` frames, which `read` hands out as views
        self.buf = np.memmap(path, np.uint8, "r") if os.path.getsize(path) else b""
        base = 0
        while bytes(self.buf[base:base + 8]) != SIGNATURE:     # a user block pushes the superblock to 512, 1024, 2048 ...
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self.buf):
                raise ValueError(f"{path}: no HDF5 signature")
        b = self.buf
        ver = b[base + 8]
        if ver not in (0, 1) or b[base + 13] != 8 or b[base + 14] != 8:
            raise ValueError(f"{path}: superblock version {ver} / offset size {b[base + 13]} not supported by h5lite.read")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, base + 16)
        p = base + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", b, p)   # every address is relative to the base address
        _, self.root_header, cache, _ = struct.unpack_from("<QQII", b, p + 32)
        if self.root_header == UNDEF:
            raise ValueError(f"{path}: unfinished file (no root group): the writer was not closed")
        if self.eof > len(b):   # the end-of-file address is absolute (it includes a user block)
            raise ValueError(f"{path}: truncated (end-of-file address {self.eof} beyond {len(b)} bytes)")

    def at(self, addr):
        return self.base + addr

    # -- object headers ---------------------------------------------------------------------------------------
    def messages(self, addr):
        b, p = self.buf, self.at(addr)
        ver, nmsg, _, size = struct.unpack_from("<BxHII", b, p)
        if ver != 1:
            raise ValueError(f"object header version {ver} not supported")
        blocks, out = [(p + 16, size)], []
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, q)
                body = bytes(b[q + 8:q + 8 + msize])
                if mtype == MSG_CONTINUATION:
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.at(off), ln))
                out.append((mtype, flags, body))
                q += 8 + msize
                left -= 8 + msize
        return out

    # -- groups -----------------------------------------------------------------------------------------------
    def heap_name(self, heap_addr, offset):
        p = self.at(heap_addr)
        if bytes(self.buf[p:p + 4]) != b"HEAP":
            raise ValueError("bad local heap signature")
        seg = self.at(struct.unpack_from("<Q", self.buf, p + 24)[0])
        raw = bytes(self.buf[seg + offset:seg + offset + 1024])
        return raw[:raw.index(b"\0")].decode("utf-8")

    def group_entries(self, btree_addr, heap_addr):
        b, p = self.buf, self.at(btree_addr)
        if bytes(b[p:p + 4]) != b"TREE":
            raise ValueError("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
        if ntype != 0:
            raise ValueError("not a group B-tree")
        out, q = [], p + 24
        for i in range(used):
            child = struct.unpack_from("<Q", b, q + 8)[0]
            q += 16
            if level > 0:
                out += self.group_entries(child, heap_addr)
                continue
            s = self.at(child)
            if bytes(b[s:s + 4]) != b"SNOD":
                raise ValueError("bad symbol table node signature")
            for j in range(struct.unpack_from("<H", b, s + 6)[0]):
                name_off, header, cache = struct.unpack_from("<QQI", b, s + 8 + 40 * j)
                out.append((self.heap_name(heap_addr, name_off), header))
        return out

    # -- datatypes / dataspaces -------------------------------------------------------------------------------
    @staticmethod
    def dtype_of(body):
        cls, ver = body[0] & 0x0F, body[0] >> 4
        bits0, bits1 = body[1], body[2]
        size = struct.unpack_from("<I", body, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 9 and (bits0 & 0x0F) == 1:
            return "vlen_str"
        raise ValueError(f"datatype class {cls} (version {ver}) not supported by h5lite.read")

    @staticmethod
    def shape_of(body):
        ver, rank, flags = body[0], body[1], body[2]
        off = 8 if ver == 1 else 4
        dims = struct.unpack_from(f"<{rank}Q", body, off)
        maxd = struct.unpack_from(f"<{rank}Q", body, off + 8 * rank) if flags & 1 else None
        return tuple(int(d) for d in dims), maxd

    def global_heap_object(self, addr, index):
        b, p = self.buf, self.at(addr)
        if bytes(b[p:p + 4]) != b"GCOL":
            raise ValueError("bad global heap signature")
        size = struct.unpack_from("<Q", b, p + 8)[0]
        q = p + 16
        while q + 16 <= p + size:
            idx, _, osize = struct.unpack_from("<HH4xQ", b, q)
            if idx == 0:
                break
            if idx == index:
                return bytes(b[q + 16:q + 16 + osize])
            q += 16 + (osize + 7) // 8 * 8
        raise ValueError(f"global heap object {index} not found")

    def decode(self, dtype, shape, raw):
        if isinstance(dtype, str):   # variable-length strings: (length, collection address, object index) per element
            n = int(np.prod(shape)) if shape else 1
            vals = []
            for i in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self.global_heap_object(addr, idx)[:ln].decode("utf-8"))
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        a = np.frombuffer(raw, dtype, count=int(np.prod(shape)) if shape else 1)   # a view when `raw` is the mapped file
        return a.reshape(shape) if shape else a[0]

    # -- datasets ---------------------------------------------------------------------------------------------
    def chunks(self, btree_addr, rank1):
        """[(offsets, nbytes, filter_mask, address)] of a chunk B-tree, depth first."""
        b, p = self.buf, self.at(btree_addr)
        if bytes(b[p:p + 4]) != b"TREE":
            raise ValueError("bad chunk B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
        if ntype != 1:
            raise ValueError("not a chunk B-tree")
        key_size, out, q = 8 + 8 * rank1, [], p + 24
        for i in range(used):
            nbytes, fmask = struct.unpack_from("<II", b, q)
            offs = struct.unpack_from(f"<{rank1}Q", b, q + 8)
            child = struct.unpack_from("<Q", b, q + key_size)[0]
            q += key_size + 8
            out += self.chunks(child, rank1) if level > 0 else [(offs, nbytes, fmask, child)]
        return out

    def dataset(self, msgs):
        dtype = shape = maxshape = None
        layout, filters = None, []
        for mtype, _, body in msgs:
            if mtype == MSG_DATATYPE:
                dtype = self.dtype_of(body)
            elif mtype == MSG_DATASPACE:
                shape, maxshape = self.shape_of(body)
            elif mtype == MSG_LAYOUT:
                layout = body
            elif mtype == MSG_FILTERS:
                filters = self.filters_of(body)
        if layout is None or dtype is None or shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        itemsize = 16 if isinstance(dtype, str) else np.dtype(dtype).itemsize
        if layout[0] in (1, 2):   # libhdf5 < 1.6.3: version, dimensionality, class, 5 reserved, [address], 4-byte dimensions
            ndim, cls = layout[1], layout[2]
            addr = struct.unpack_from("<Q", layout, 8)[0] if cls != 0 else UNDEF
            q = 8 + (8 if cls != 0 else 0)
            dims = struct.unpack_from(f"<{ndim}I", layout, q)
            q += 4 * ndim
            if cls == 0:
                size = struct.unpack_from("<I", layout, q)[0]
                layout = struct.pack("<BBH", 3, 0, size) + layout[q + 4:q + 4 + size]
            elif cls == 1:
                layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", addr, n * itemsize)
            else:
                layout = struct.pack("<BBB", 3, 2, ndim) + struct.pack("<Q", addr) + b"".join(struct.pack("<I", v) for v in dims)
        if layout[0] != 3:
            raise ValueError(f"data layout message version {layout[0]} not supported")
        if layout[1] == 0:      # compact
            size = struct.unpack_from("<H", layout, 2)[0]
            return self.decode(dtype, shape, layout[4:4 + size])
        if layout[1] == 1:      # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF or n == 0:
                return np.zeros(shape, dtype)
            return self.decode(dtype, shape, self.buf[self.at(addr):self.at(addr) + n * itemsize])   # a view of the map
        rank1 = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from(f"<{rank1}I", layout, 11)[:-1]
        if btree == UNDEF or n == 0:
            return np.zeros(shape, dtype)
        chunks = self.chunks(btree, rank1)
        cbytes = int(np.prod(cdims)) * itemsize
        if (not filters and tuple(cdims) == (1,) + tuple(shape[1:]) and len(chunks) == shape[0] and
                all(c[0][0] == i and c[3] == chunks[0][3] + i * cbytes for i, c in enumerate(chunks))):
            # one frame per chunk, stored back to back in arrival order (what the Writer produces): a view, no copy
            return np.frombuffer(self.buf, dtype, count=n, offset=self.at(chunks[0][3])).reshape(shape)
        out = np.zeros(shape, dtype)
        for offs, nbytes, fmask, addr in chunks:
            raw = bytes(self.buf[self.at(addr):self.at(addr) + nbytes])
            for k, (fid, _) in reversed(list(enumerate(filters))):
                if fmask >> k & 1:
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:   # shuffle
                    raw = np.frombuffer(raw, np.uint8).reshape(itemsize, -1).T.tobytes()
                else:
                    raise ValueError(f"filter {fid} not supported by h5lite.read")
            block = np.frombuffer(raw, dtype, count=int(np.prod(cdims))).reshape(cdims)
            sel = tuple(slice(int(o), min(int(o) + c, s)) for o, c, s in zip(offs[:-1], cdims, shape))
            out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out

    @staticmethod
    def filters_of(body):
        ver, n = body[0], body[1]
        q, out = (8 if ver == 1 else 2), []
        for _ in range(n):
            fid, name_len, flags, ncv = struct.unpack_from("<HHHH", body, q)
            q += 8
            if ver == 1 or fid >= 256:
                q += (name_len + 7) // 8 * 8 if ver == 1 else name_len
            else:
                q -= 2   # version 2 omits the name length for library filters
            vals = struct.unpack_from(f"<{ncv}I", body, q)
            q += 4 * ncv + (4 if ver == 1 and ncv % 2 else 0)
            out.append((fid, vals))
        return out

    def attributes(self, msgs):
        out = {}
        for mtype, _, body in msgs:
            if mtype != MSG_ATTRIBUTE:
                continue
            ver, name_size, dt_size, ds_size = struct.unpack_from("<BxHHH", body, 0)
            if ver != 1:
                raise ValueError(f"attribute message version {ver} not supported")
            r8 = lambda v: (v + 7) // 8 * 8  # noqa: E731
            q = 8
            name = body[q:q + name_size].split(b"\0")[0].decode("utf-8")
            q += r8(name_size)
            dtype = self.dtype_of(body[q:q + dt_size])
            q += r8(dt_size)
            shape, _ = self.shape_of(body[q:q + ds_size]) if ds_size >= 8 and body[q + 1] else ((), None)
            q += r8(ds_size)
            v = self.decode(dtype, shape, body[q:])
            out[name] = v.copy() if isinstance(v, np.ndarray) else v
        return out

    def group(self, header_addr):
        msgs = self.messages(header_addr)
        out = {"attrs": self.attributes(msgs)}
        for mtype, _, body in msgs:
            if mtype == MSG_SYMTAB:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                for name, header in self.group_entries(btree, heap):
                    sub = self.messages(header)
                    if any(t == MSG_SYMTAB for t, _, _ in sub):
                        out[name] = self.group(header)
                    else:
                        out[name] = self.dataset(sub)
                        at = self.attributes(sub)
                        if at:
                            out.setdefault("dataset_attrs", {})[name] = at
        return out


def read(path):
    """{dataset name: array, sub-group name: dict, "attrs": {...}} of the root group.  Uncompressed datasets come back
    as read-only views of the memory-mapped file (like the raw container's memmap): nothing is copied until it is used."""
    r = _Reader(path)
    return r.group(r.root_header)


def is_hdf5(path):
    try:
        with open(path, "rb") as f:
            return f.read(8) == SIGNATURE
    except OSError:
        return False


if __name__ == "__main__":   # python h5lite.py FILE: list the contents (a poor man's h5ls)
    import sys

    def show(d, indent=""):
        for k, v in d.items():
            if isinstance(v, dict):
                print(f"{indent}{k}/")
                show(v, indent + "  ")
            elif isinstance(v, np.ndarray):
                print(f"{indent}{k}: {v.dtype} {v.shape}")
            else:
                s = json.dumps(v) if isinstance(v, str) else repr(v)
                print(f"{indent}{k}: {s[:100]}")

    show(read(sys.argv[1]))
    print(f"{os.path.getsize(sys.argv[1])} bytes")
