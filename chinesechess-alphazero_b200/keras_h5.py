"""Minimal pure-Python reader for Keras 2.0.x `save_weights` files (HDF5, superblock v0).

The reference stores its networks as `model_*_weight.h5` (agent/model.py:99-100,112-113 via Keras/h5py).  h5py is not
available here, so this module reads exactly the subset of HDF5 that such files use: version-0 superblock, old-style
groups (v1 B-trees + local heaps + symbol nodes), version-1 object headers, contiguous little-endian float32 datasets.
Anything else (chunked / compressed data, new-style groups) raises NotImplementedError.

`read_keras_weights(path)` -> dict  "<layer>/<weight>" -> np.float32 array in Keras layout, e.g.
"res3_conv1-3-192/kernel" (HWIO), "policy_out/bias", "input_batchnorm/moving_variance".
"""
import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class _H5:
    def __init__(self, data):
        self.d = data
        if data[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        if data[8] != 0:
            raise NotImplementedError(f"HDF5 superblock version {data[8]} (only 0 is supported)")
        if data[13] != 8 or data[14] != 8:
            raise NotImplementedError("only 8-byte offsets / lengths")
        self.base = struct.unpack_from("<Q", data, 24)[0]
        self.root = self._entry(56)

    # symbol-table entry: name offset, object header address, cache type, scratch (btree, heap for groups)
    def _entry(self, off):
        name_off, ohdr, cache = struct.unpack_from("<QQI", self.d, off)
        btree = heap = None
        if cache == 1:
            btree, heap = struct.unpack_from("<QQ", self.d, off + 24)
        return {"name_off": name_off, "ohdr": ohdr, "btree": btree, "heap": heap}

    def _heap_name(self, heap_addr, name_off):
        d = self.d
        if d[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        data_addr = struct.unpack_from("<Q", d, heap_addr + 24)[0]
        start = data_addr + name_off
        end = d.index(b"\x00", start)
        return d[start:end].decode()

    def _group_entries(self, btree, heap):
        """Yield (name, entry) of a group by walking its v1 B-tree down to the symbol nodes."""
        d = self.d
        if d[btree:btree + 4] != b"TREE":
            raise ValueError("bad B-tree node")
        node_type, level, used = struct.unpack_from("<BBH", d, btree + 4)
        if node_type != 0:
            raise NotImplementedError("non-group B-tree")
        pos = btree + 24                      # signature 4, type 1, level 1, used 2, left 8, right 8
        children = []
        for i in range(used):
            pos += 8                          # key i
            children.append(struct.unpack_from("<Q", d, pos)[0])
            pos += 8
        for c in children:
            if level > 0:
                yield from self._group_entries(c, heap)
            else:
                if d[c:c + 4] != b"SNOD":
                    raise ValueError("bad symbol node")
                n = struct.unpack_from("<H", d, c + 6)[0]
                for k in range(n):
                    e = self._entry(c + 8 + 40 * k)
                    yield self._heap_name(heap, e["name_off"]), e

    def _messages(self, ohdr):
        """Yield (type, payload bytes) of a version-1 object header, following continuation blocks."""
        d = self.d
        version, _, nmsg, _, hsize = struct.unpack_from("<BBHII", d, ohdr)
        if version != 1:
            raise NotImplementedError(f"object header version {version}")
        blocks = [(ohdr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", d, pos)
                payload = d[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                seen += 1
                if mtype == 0x0010:
                    off, length = struct.unpack_from("<QQ", payload, 0)
                    blocks.append((off, length))
                else:
                    yield mtype, payload

    def _dataset(self, ohdr):
        shape = dtype_size = addr = nbytes = None
        is_float = False
        for mtype, p in self._messages(ohdr):
            if mtype == 0x0001:
                ver, rank, flags = p[0], p[1], p[2]
                off = 8 if ver == 1 else 4
                shape = struct.unpack_from("<%dQ" % rank, p, off) if rank else ()
            elif mtype == 0x0003:
                cls = p[0] & 0x0F
                dtype_size = struct.unpack_from("<I", p, 4)[0]
                is_float = cls == 1 and (p[1] & 1) == 0          # floating point, little endian
            elif mtype == 0x0008:
                if p[0] != 3:
                    raise NotImplementedError(f"data layout message version {p[0]}")
                if p[1] != 1:
                    raise NotImplementedError("only contiguous datasets (no chunking / compression)")
                addr, nbytes = struct.unpack_from("<QQ", p, 2)
        if shape is None or addr is None:
            return None
        if not is_float or dtype_size != 4:
            raise NotImplementedError("only little-endian float32 datasets")
        n = int(np.prod(shape)) if shape else 1
        if addr == _UNDEF or n == 0:
            return np.zeros(shape, np.float32)
        return np.frombuffer(self.d, dtype="<f4", count=n, offset=self.base + addr).reshape(shape).copy()

    def walk(self, entry=None, prefix=""):
        """Yield (path, array) for every dataset below `entry`."""
        entry = entry or self.root
        btree, heap = entry["btree"], entry["heap"]
        if btree is None:                      # group without cached scratch: read its symbol-table message
            for mtype, p in self._messages(entry["ohdr"]):
                if mtype == 0x0011:
                    btree, heap = struct.unpack_from("<QQ", p, 0)
        if btree is None:
            arr = self._dataset(entry["ohdr"])
            if arr is not None:
                yield prefix.rstrip("/"), arr
            return
        for name, e in self._group_entries(btree, heap):
            is_group = e["btree"] is not None or any(t == 0x0011 for t, _ in self._messages(e["ohdr"]))
            if is_group:
                yield from self.walk(e, prefix + name + "/")
            else:
                arr = self._dataset(e["ohdr"])
                if arr is not None:
                    yield prefix + name, arr


def read_keras_weights(path):
    with open(path, "rb") as f:
        h5 = _H5(f.read())
    out = {}
    for p, arr in h5.walk():
        parts = p.split("/")
        layer, weight = parts[0], parts[-1].split(":")[0]       # "<layer>/<layer>/<weight>:0"
        out[f"{layer}/{weight}"] = arr.astype(np.float32)
    return out
