"""Minimal pure-Python HDF5 reader (h5py is not part of this image, and the reference's multigroup
cross-sections -- material.py:36-46 -- live in HDF5 files).

Covers what classic h5py-written files of plain numeric arrays use: superblock version 0 or 1, groups stored
as symbol tables (v1 B-tree + local heap + SNOD nodes), version-1 object headers with continuation blocks,
simple dataspaces, little- or big-endian fixed-point and IEEE floating-point datatypes, and contiguous storage
(what the reference's files use).  Anything else (compact or chunked storage, filters / compression,
new-style groups, variable-length or compound types) raises Hdf5Error naming the feature, so a file this reader
cannot handle never yields silently wrong numbers.

    with File(path) as hf:          # same spelling as h5py for the calls material.py makes
        a = hf["scat_xs"][:]
        names = hf.keys()
"""
from __future__ import annotations


import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5Error(RuntimeError):
    pass


class Dataset:
    def __init__(self, name, array):
        self.name, self._a = name, array
        self.shape, self.dtype = array.shape, array.dtype

    def __getitem__(self, key):
        return self._a[key].copy() if isinstance(self._a[key], np.ndarray) else self._a[key]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._a, dtype=dtype)


class File:
    def __init__(self, path, mode="r"):
        if mode != "r":
            raise Hdf5Error("hdf5_lite is read-only")
        with open(path, "rb") as f:
            self._b = f.read()
        self.filename = path
        self._parse_superblock()
        self._links = self._read_group(self._root_btree, self._root_heap)
        self._cache = {}

    # ---- context manager / mapping protocol (the part of h5py.File the reference uses) -------------------
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        self._b = b""

    def keys(self):
        return list(self._links)

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, name):
        if name not in self._links:
            raise KeyError(f"{name!r} not in {self.filename} (has {sorted(self._links)})")
        if name not in self._cache:
            self._cache[name] = Dataset(name, self._read_dataset(self._links[name]))
        return self._cache[name]

    # ---- low level ------------------------------------------------------------------------------------------
    def _u(self, off, size):
        return int.from_bytes(self._b[off:off + size], "little")

    def _parse_superblock(self):
        b = self._b
        if b[:8] != _SIG:
            raise Hdf5Error(f"{self.filename}: not an HDF5 file (no signature at offset 0)")
        ver = b[8]
        if ver not in (0, 1):
            raise Hdf5Error(f"superblock version {ver} (new-style file) is not supported")
        self._so, self._sl = b[13], b[14]               # size of offsets / lengths
        if self._so != 8 or self._sl != 8:
            raise Hdf5Error(f"offset/length sizes {self._so}/{self._sl} are not supported")
        pos = 24 if ver == 0 else 28                    # v1 adds indexed-storage K + reserved
        self._base = self._u(pos, 8)
        root = pos + 32                                 # base, free-space, EOF, driver-info addresses
        cache_type = self._u(root + 16, 4)
        if cache_type != 1:
            raise Hdf5Error("root group without cached symbol-table addresses is not supported")
        self._root_btree = self._u(root + 24, 8)
        self._root_heap = self._u(root + 32, 8)

    def _heap_string(self, heap_addr, off):
        b = self._b
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise Hdf5Error("local heap signature missing")
        data = self._u(heap_addr + 24, 8) + self._base
        end = b.index(b"\0", data + off)
        return b[data + off:end].decode()

    def _read_group(self, btree_addr, heap_addr):
        """name -> object header address for every link of a symbol-table group."""
        links = {}
        stack = [btree_addr + self._base]
        while stack:
            a = stack.pop()
            b = self._b
            if b[a:a + 4] == b"SNOD":
                n = self._u(a + 6, 2)
                for i in range(n):
                    e = a + 8 + 40 * i
                    links[self._heap_string(heap_addr + self._base, self._u(e, 8))] = self._u(e + 8, 8) + self._base
                continue
            if b[a:a + 4] != b"TREE" or b[a + 4] != 0:
                raise Hdf5Error("group B-tree node expected")
            used = self._u(a + 6, 2)
            p = a + 24 + 8                              # past the header and key 0
            for i in range(used):
                stack.append(self._u(p, 8) + self._base)
                p += 16                                 # child address + next key
        return links

    def _messages(self, addr):
        """(type, payload bytes) of a version-1 object header, following continuation blocks."""
        b = self._b
        if b[addr] != 1:
            raise Hdf5Error(f"object header version {b[addr]} is not supported")
        count = self._u(addr + 2, 2)
        blocks = [(addr + 16, self._u(addr + 8, 4))]
        out = []
        while blocks and len(out) < count:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < count:
                mtype, msize = self._u(p, 2), self._u(p + 2, 2)
                body = b[p + 8:p + 8 + msize]
                if mtype == 0x10:
                    blocks.append((self._u(p + 8, 8) + self._base, self._u(p + 16, 8)))
                out.append((mtype, body))
                p += 8 + msize
        return out

    def _read_dataset(self, addr):
        shape = dtype = layout = None
        for mtype, m in self._messages(addr):
            if mtype == 0x01:
                shape = self._dataspace(m)
            elif mtype == 0x03:
                dtype = self._datatype(m)
            elif mtype == 0x08:
                layout = m
            elif mtype == 0x0B:
                raise Hdf5Error("filtered (compressed) datasets are not supported")
        if shape is None or dtype is None or layout is None:
            raise Hdf5Error("object is not a simple dataset")
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        raw = self._layout_bytes(layout, shape, dtype.itemsize, n)
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).astype(dtype.newbyteorder("="))

    @staticmethod
    def _dataspace(m):
        ver, rank = m[0], m[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            if m[3] == 2:
                raise Hdf5Error("null dataspace")
            p = 4
        else:
            raise Hdf5Error(f"dataspace version {ver}")
        return tuple(int.from_bytes(m[p + 8 * i:p + 8 * i + 8], "little") for i in range(rank))

    @staticmethod
    def _datatype(m):
        cls, ver = m[0] & 0x0F, m[0] >> 4
        bits0 = m[1]
        size = int.from_bytes(m[4:8], "little")
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5Error(f"{size}-byte floating point")
            return np.dtype(f"{order}f{size}")
        raise Hdf5Error(f"datatype class {cls} (version {ver}) is not supported")

    def _layout_bytes(self, m, shape, itemsize, n):
        """Bytes of a version-3 contiguous layout message -- the only storage the reference's files use and therefore
        the only one this reader has fixtures for; everything else is refused rather than guessed at."""
        ver = m[0]
        if ver != 3:
            raise Hdf5Error(f"data layout message version {ver} is not supported")
        cls = m[1]
        if cls != 1:
            raise Hdf5Error({0: "compact", 2: "chunked"}.get(cls, f"class-{cls}") + " dataset storage is not supported")
        addr = int.from_bytes(m[2:10], "little")
        if addr == _UNDEF:
            return bytes(n * itemsize)                  # allocated but never written: fill value 0
        addr += self._base
        return self._b[addr:addr + n * itemsize]
