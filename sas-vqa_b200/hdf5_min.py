"""Minimal HDF5 writer/reader for the one artefact the sampling stage persists: a contiguous float32
dataset ``sampled_frames`` of shape ``(N_videos, K, 3*IMG*IMG)``.

The reference writes it with ``h5py.File(path, 'w').create_dataset("sampled_frames", shape)``
(src/preprocessing/extract_features.py:77-79) and its training side opens it with
``h5py.File(path, 'r')['sampled_frames']`` (src/datasets/dataset_base.py:104).  h5py / libhdf5 are not in this
image, so the file is laid out by hand in the oldest, most widely readable form of the format (HDF5 File
Format Specification, version 1.x structures -- what libhdf5 1.6 wrote and every later release reads):

    superblock v0 -> root group object header (v1) with a symbol-table message -> local heap (link names),
    group B-tree node (v1, "TREE") -> symbol-table node ("SNOD") -> dataset object header (v1) carrying
    fill-value, datatype (IEEE little-endian float32), dataspace (v1, simple, fixed dims) and data-layout
    (v3, contiguous) messages -> raw row-major data, 4 KiB aligned.

OUTPUT CONTRACT: the writer has been checked field by field against the format specification and is read back by this module's
own reader, by the reference's unmodified ``VideoQADataset`` / ``GITVideoQACollator`` through an ``h5py.File(p, 'r')[name]`` shim
over that reader (tests/test_consumer_ref.py), and by real h5py wherever h5py is importable (the same test then uses it) -- but
libhdf5 / h5py are absent from the build image, so a file written HERE has not been opened by libhdf5 itself.  ``writer.py`` uses
h5py for ``.h5`` paths whenever it is importable and records which backend wrote a file (``SampledFramesWriter.backend_used``); the
``.npy`` backend is the format tested end to end.

The reader below understands exactly those structures (plus the v1/v2 layout messages and header continuation
blocks libhdf5 1.6 emitted); tests pin it against a genuine libhdf5-written file that ships with scipy and then
use it to read what the writer produced.  Rows are exposed as a ``numpy.memmap`` so the extraction loop
(``sampled_frames_h5[i] = frames``, extract_features.py:96-97) streams into the file in place.
"""
from __future__ import annotations

import os
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16
DATA_ALIGN = 4096


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _message(msg_type: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", msg_type, len(body), flags) + body


def _object_header(messages: list[bytes]) -> bytes:
    payload = b"".join(messages)
    # version 1 | reserved | number of messages | reference count | header size | pad to 8
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(payload)) + payload


_DTYPES = {
    np.dtype("<f4"): struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127),
    np.dtype("<f8"): struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023),
}


def create_dataset_file(path: str, name: str, shape: tuple, dtype=np.float32) -> np.memmap:
    """Creates ``path`` holding one contiguous dataset ``name`` and returns a writable memmap of its data."""
    dtype = np.dtype(dtype).newbyteorder("<")
    if dtype not in _DTYPES:
        raise TypeError(f"unsupported dtype {dtype}")
    shape = tuple(int(s) for s in shape)
    if not shape or any(s <= 0 for s in shape):
        raise ValueError("dataset shape must have positive extents")
    name_b = name.encode("ascii") + b"\x00"
    n_bytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize

    # ---- fixed layout of the metadata blocks (addresses relative to base address 0)
    a_root = 96                                                   # root group object header (16 + 24 + 8)
    root_len = 16 + 24
    a_heap = a_root + root_len                                    # local heap header (32) + data segment
    heap_data_len = 8 + len(_pad8(name_b)) + 16                   # "" | name | one free block
    heap_data_len += -heap_data_len % 8
    a_heap_data = a_heap + 32
    a_tree = a_heap_data + heap_data_len                          # group B-tree node
    tree_len = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    a_snod = a_tree + tree_len                                    # symbol table node
    snod_len = 8 + 2 * LEAF_K * 40
    a_dset = a_snod + snod_len                                    # dataset object header
    dataspace = struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)
    a_data_placeholder = 0
    fill = struct.pack("<BBBBI", 1, 2, 2, 1, 0)                   # v1: alloc late, write if-set, defined, size 0

    def dataset_header(a_data: int) -> bytes:
        layout = struct.pack("<BBQQ", 3, 1, a_data, n_bytes)      # v3, contiguous: address, size
        return _object_header([_message(0x0005, fill, 1), _message(0x0003, _DTYPES[dtype], 1),
                               _message(0x0001, dataspace), _message(0x0008, layout)])

    dset_len = len(dataset_header(a_data_placeholder))
    a_data = a_dset + dset_len
    a_data += -a_data % DATA_ALIGN
    eof = a_data + n_bytes

    root_entry = struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_tree, a_heap)
    superblock = (SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0) +
                  struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + root_entry)
    assert len(superblock) == a_root
    root = _object_header([_message(0x0011, struct.pack("<QQ", a_tree, a_heap))])
    assert len(root) == root_len
    name_off = 8
    free_off = name_off + len(_pad8(name_b))
    heap_data = (b"\x00" * 8 + _pad8(name_b) + struct.pack("<QQ", 1, heap_data_len - free_off)).ljust(heap_data_len, b"\x00")
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, heap_data_len, free_off, a_heap_data) + heap_data
    tree = (b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, a_snod, name_off)).ljust(tree_len, b"\x00")
    entry = struct.pack("<QQII16x", name_off, a_dset, 0, 0)
    snod = (b"SNOD" + struct.pack("<BBH", 1, 0, 1) + entry).ljust(snod_len, b"\x00")
    meta = superblock + root + heap + tree + snod + dataset_header(a_data)
    assert len(meta) == a_dset + dset_len
    with open(path, "wb") as f:
        f.write(meta)
        f.truncate(eof)                                           # sparse: rows are filled through the memmap
    return np.memmap(path, dtype=dtype, mode="r+", offset=a_data, shape=shape)


# ------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf, base: int):
        self.buf, self.base = buf, base

    def u(self, off: int, n: int) -> int:
        return int.from_bytes(self.buf[off:off + n], "little")

    def messages(self, addr: int) -> list[tuple[int, bytes]]:
        o = self.base + addr
        if self.buf[o] != 1:
            raise ValueError(f"object header version {self.buf[o]} is not supported by this minimal reader")
        n_msg, size = self.u(o + 2, 2), self.u(o + 8, 4)
        out: list[tuple[int, bytes]] = []
        blocks = [(o + 16, o + 16 + size)]
        while blocks and len(out) < n_msg:
            p, end = blocks.pop(0)
            while p + 8 <= end and len(out) < n_msg:
                t, sz = self.u(p, 2), self.u(p + 2, 2)
                body = bytes(self.buf[p + 8:p + 8 + sz])
                out.append((t, body))
                if t == 0x0010:                                   # header continuation
                    c_addr, c_len = struct.unpack("<QQ", body[:16])
                    blocks.append((self.base + c_addr, self.base + c_addr + c_len))
                p += 8 + sz
        return out

    def heap_name(self, heap_addr: int, off: int) -> str:
        h = self.base + heap_addr
        assert bytes(self.buf[h:h + 4]) == b"HEAP"
        data = self.base + self.u(h + 24, 8)
        end = data + off
        while self.buf[end] != 0:
            end += 1
        return bytes(self.buf[data + off:end]).decode("ascii")

    def group_entries(self, tree_addr: int, heap_addr: int) -> dict[str, int]:
        t = self.base + tree_addr
        assert bytes(self.buf[t:t + 4]) == b"TREE" and self.buf[t + 4] == 0
        level, used = self.buf[t + 5], self.u(t + 6, 2)
        out: dict[str, int] = {}
        for i in range(used):
            child = self.u(t + 24 + 8 + 16 * i, 8)
            if level > 0:
                out.update(self.group_entries(child, heap_addr))
                continue
            s = self.base + child
            assert bytes(self.buf[s:s + 4]) == b"SNOD"
            for j in range(self.u(s + 6, 2)):
                e = s + 8 + 40 * j
                out[self.heap_name(heap_addr, self.u(e, 8))] = self.u(e + 8, 8)
        return out


def _parse_float_type(body: bytes) -> np.dtype:
    cls, size = body[0] & 0x0F, struct.unpack("<I", body[4:8])[0]
    if cls != 1 or body[1] & 1:
        raise TypeError("only little-endian IEEE floating point datasets are supported")
    return np.dtype(f"<f{size}")


def open_datasets(path: str, mode: str = "r") -> dict[str, np.memmap]:
    """{name: memmap} of the contiguous floating-point datasets in the root group of ``path``."""
    with open(path, "rb") as f:
        head = f.read(1 << 20)
    base = next((o for o in (0, 512, 1024, 2048, 4096) if head[o:o + 8] == SIGNATURE), None)
    if base is None or head[base + 8] != 0:
        raise ValueError("not an HDF5 file with a version-0 superblock")
    if head[base + 13] != 8 or head[base + 14] != 8:
        raise ValueError("only 8-byte offsets and lengths are supported")
    r = _Reader(head, base)
    assert r.u(base + 24, 8) in (0, base)                          # base address
    root = base + 56
    tree_addr, heap_addr = None, None
    for t, body in r.messages(r.u(root + 8, 8)):
        if t == 0x0011:
            tree_addr, heap_addr = struct.unpack("<QQ", body[:16])
    if tree_addr is None:
        raise ValueError("root group has no symbol table")
    out = {}
    for name, addr in r.group_entries(tree_addr, heap_addr).items():
        shape = dtype = data_addr = None
        for t, body in r.messages(addr):
            if t == 0x0001:
                rank = body[1]
                shape = tuple(struct.unpack("<Q", body[8 + 8 * i:16 + 8 * i])[0] for i in range(rank))
            elif t == 0x0003:
                dtype = _parse_float_type(body)
            elif t == 0x0008:
                if body[0] == 3 and body[1] == 1:
                    data_addr = struct.unpack("<Q", body[2:10])[0]
                elif body[0] in (1, 2) and body[2] == 1:
                    data_addr = struct.unpack("<Q", body[8:16])[0]
        if shape is None or dtype is None or data_addr is None or data_addr == UNDEF:
            continue
        out[name] = np.memmap(path, dtype=dtype, mode=mode, offset=base + data_addr, shape=shape)
    return out
