"""Minimal read-only HDF5 parser (TEST INFRASTRUCTURE ONLY): enough of the format to list the groups and read the
contiguous / compact little-endian numeric datasets of a Keras 3 `model.weights.h5` (h5py and libhdf5 are not in
this image).  Supports superblock v0-v3, object headers v1 and v2, old-style groups (symbol table: B-tree v1 + local
heap) and new-style compact groups (link messages), dataspace v1/v2, fixed-point / floating-point datatypes, data
layout v3 classes compact and contiguous.  Chunked / filtered datasets raise NotImplementedError.
"""

from __future__ import annotations

import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, data: bytes):
        self.b = data
        if data[:8] != SIG:
            raise ValueError("not an HDF5 file")
        ver = data[8]
        if ver in (0, 1):
            self.so, self.sl = data[13], data[14]
            p = 24 + (4 if ver == 1 else 0)
            self.base = self._off(p)
            root_entry = p + 4 * self.so
            self.root = self._off(root_entry + self.so)          # object header address of the root group
        elif ver in (2, 3):
            self.so, self.sl = data[9], data[10]
            self.base = self._off(12)
            self.root = self._off(12 + 3 * self.so)
        else:
            raise NotImplementedError(f"superblock version {ver}")
        if self.so != 8 or self.sl != 8:
            raise NotImplementedError("only 8-byte offsets / lengths")

    def _off(self, p):
        return struct.unpack_from("<Q", self.b, p)[0]

    # ---- object headers ------------------------------------------------------------------------------
    def messages(self, addr):
        """[(type, payload bytes)] of the object header at `addr` (continuations followed)."""
        b = self.b
        out = []
        if b[addr:addr + 4] == b"OHDR":
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            szf = 1 << (flags & 3)
            chunk0 = int.from_bytes(b[p:p + szf], "little")
            p += szf
            blocks = [(p, p + chunk0)]
            track = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, mflags = b[p], struct.unpack_from("<H", b, p + 1)[0], b[p + 3]
                    p += 4 + (2 if track else 0)
                    body = b[p:p + msize]
                    p += msize
                    if mtype == 0x10:
                        o, ln = struct.unpack_from("<QQ", body, 0)
                        blocks.append((o + 4, o + ln - 4))      # "OCHK" signature ... checksum
                    elif mtype != 0:
                        out.append((mtype, body))
            return out
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks = [(addr + 16, addr + 16 + hsize)]
        while blocks and len(out) < 4096:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    o, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((o, o + ln))
                elif mtype != 0:
                    out.append((mtype, body))
        return out

    # ---- groups ----------------------------------------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        b = self.b
        assert b[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = self._off(heap_addr + 8 + 2 * self.sl)
        s = data_addr + off
        e = b.index(b"\0", s)
        return b[s:e].decode()

    def _btree_group(self, addr, heap, out):
        b = self.b
        if b[addr:addr + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, addr + 6)[0]
            p = addr + 8
            for _ in range(n):
                name_off, ohdr = struct.unpack_from("<QQ", b, p)
                out[self._heap_name(heap, name_off)] = ohdr
                p += 2 * self.so + 24
            return
        assert b[addr:addr + 4] == b"TREE", b[addr:addr + 4]
        level, used = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        p = addr + 8 + 2 * self.so
        for i in range(used):
            child = self._off(p + self.sl)                       # key_i, child_i
            self._btree_group(child, heap, out)
            p += self.sl + self.so

    def links(self, addr):
        """{name: object header address} of the group at `addr`."""
        out = {}
        for mtype, body in self.messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                self._btree_group(btree, heap, out)
            elif mtype == 0x06:                                  # link message
                ver, flags = body[0], body[1]
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = body[p]; p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                lsz = 1 << (flags & 3)
                ln = int.from_bytes(body[p:p + lsz], "little"); p += lsz
                name = body[p:p + ln].decode(); p += ln
                if ltype == 0:
                    out[name] = struct.unpack_from("<Q", body, p)[0]
            elif mtype == 0x02:
                raise NotImplementedError("dense link storage (fractal heap)")
        return out

    # ---- datasets --------------------------------------------------------------------------------------
    def dataset(self, addr):
        """numpy array of the dataset at `addr`, or None if the object is not a dataset."""
        shape = dtype = layout = None
        for mtype, body in self.messages(addr):
            if mtype == 0x01:
                ver, rank, flags = body[0], body[1], body[2]
                p = 8 if ver == 1 else 4
                shape = tuple(struct.unpack_from("<Q", body, p + 8 * i)[0] for i in range(rank))
            elif mtype == 0x03:
                cls = body[0] & 0x0F
                bits0 = body[1]
                size = struct.unpack_from("<I", body, 4)[0]
                if bits0 & 1:
                    raise NotImplementedError("big-endian datatype")
                if cls == 1:
                    dtype = {2: np.float16, 4: np.float32, 8: np.float64}[size]
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = {(1, True): np.int8, (1, False): np.uint8, (2, True): np.int16, (2, False): np.uint16, (4, True): np.int32,
                             (4, False): np.uint32, (8, True): np.int64, (8, False): np.uint64}[(size, signed)]
                else:
                    dtype = ("raw", size)
            elif mtype == 0x08:
                layout = body
        if shape is None or dtype is None or layout is None:
            return None
        if isinstance(dtype, tuple):
            return None
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * np.dtype(dtype).itemsize
        ver, cls = layout[0], layout[1]
        if ver != 3:
            raise NotImplementedError(f"data layout version {ver}")
        if cls == 1:
            a = struct.unpack_from("<Q", layout, 2)[0]
            if a == UNDEF:
                return np.zeros(shape, dtype=dtype)
            raw = self.b[self.base + a:self.base + a + nbytes]
        elif cls == 0:
            sz = struct.unpack_from("<H", layout, 2)[0]
            raw = layout[4:4 + sz][:nbytes]
        else:
            raise NotImplementedError("chunked dataset")
        return np.frombuffer(raw, dtype=dtype).reshape(shape).copy()

    def walk(self, addr=None, prefix=""):
        """Yield (path, ndarray) for every dataset below `addr` (depth first, names sorted)."""
        addr = self.root if addr is None else addr
        for name, child in sorted(self.links(addr).items()):
            path = f"{prefix}/{name}"
            arr = None
            try:
                arr = self.dataset(child)
            except NotImplementedError:
                raise
            if arr is not None:
                yield path, arr
            else:
                yield from self.walk(child, path)


def read_keras_weights(keras_path: str) -> tuple[dict, dict]:
    """(config dict, {h5 path: array}) of a Keras 3 `.keras` archive."""
    import json
    import zipfile

    with zipfile.ZipFile(keras_path) as z:
        cfg = json.loads(z.read("config.json"))
        h5 = H5File(z.read("model.weights.h5"))
    return cfg, dict(h5.walk())
