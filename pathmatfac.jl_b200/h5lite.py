"""h5lite -- the subset of the HDF5 file format the reference's scripts read and write, in NumPy.

The reference keeps its inputs and its exported parameters in HDF5 (analyses/scripts/julia/fit_matfac.jl:60-117,
bson_to_hdf.jl:18-71, script_util.jl:86-180, through HDF5.jl = libhdf5 with its default "earliest" format bounds).
The image has no HDF5 library (h5py, PyTables, libhdf5: all absent) and no network, so this module restates the on-disk
format from the published HDF5 File Format Specification, version 1.1 / 2.0 structures only:

  reads   superblock v0 / v1 (with a user block), v1 object headers with continuation blocks, symbol-table groups
          (v1 B-tree + local heap + SNOD), dataspace v1 / v2, datatypes fixed-point / IEEE float / fixed string /
          variable-length string (global heap), layouts compact / contiguous / chunked (v1 chunk B-tree, deflate and
          shuffle filters), attributes v1;
  writes  superblock v0, v1 object headers, symbol-table groups (any number of members: multi-level B-trees), contiguous
          datasets of fixed-point / float / variable-length UTF-8 string elements -- what libhdf5 itself writes for
          ``f["name"] = array`` with default properties, so HDF5.jl / h5py read it back.

The reader is pinned to a file libhdf5 produced (tests/golden/matlab73_testdouble.mat, tests/test_h5lite.py); the writer
is checked through the reader and by a structural walk of the bytes.  Host-only code, off the hot path (SURVEY.md
section 8f rank 4).

Array layout: HDF5 is row-major, Julia is column-major, and HDF5.jl stores a Julia array with its dimensions reversed
(the bytes are the Julia memory).  ``julia=True`` (the default of ``read`` / ``write``) applies that convention: a
NumPy array of shape (K, M) is written as the Julia K x M matrix (HDF5 dims (M, K)) and read back as (K, M)."""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# ------------------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------------------

class _VlenStr:
    """Marker datatype: variable-length string (class 9, type 1)."""

    def __init__(self, utf8: bool):
        self.utf8 = utf8


def _parse_datatype(b: bytes) -> Tuple[Union[np.dtype, _VlenStr], int]:
    """Datatype message (spec IV.A.2.d) -> (dtype, element size in the file)."""
    cls, ver = b[0] & 0x0F, b[0] >> 4
    bits = b[1] | (b[2] << 8) | (b[3] << 16)
    size = struct.unpack_from("<I", b, 4)[0]
    if ver not in (1, 2, 3):
        raise H5Error(f"datatype version {ver}")
    order = ">" if bits & 1 else "<"
    if cls == 0:                                            # fixed point
        return np.dtype(f"{order}{'i' if bits & 8 else 'u'}{size}"), size
    if cls == 1:                                            # IEEE float (the only floats libhdf5 writes natively)
        if size not in (2, 4, 8):
            raise H5Error(f"float of {size} bytes")
        return np.dtype(f"{order}f{size}"), size
    if cls == 3:                                            # fixed-length string (null-terminated / null- / space-padded)
        return np.dtype(f"S{size}"), size
    if cls == 9:                                            # variable length
        if bits & 0x0F != 1:
            raise H5Error("variable-length sequences are not supported (strings only)")
        return _VlenStr(utf8=((bits >> 8) & 0x0F) == 1), size
    raise H5Error(f"datatype class {cls} is not supported")


def _parse_dataspace(b: bytes) -> Optional[Tuple[int, ...]]:
    ver, rank, flags = b[0], b[1], b[2]
    if ver == 1:
        off = 8
    elif ver == 2:
        if b[3] == 2:                                       # null dataspace
            return None
        off = 4
    else:
        raise H5Error(f"dataspace version {ver}")
    return tuple(struct.unpack_from(f"<{rank}Q", b, off)) if rank else ()


class Dataset:
    def __init__(self, f: "File", msgs: List[Tuple[int, bytes]], name: str):
        self.file, self.name = f, name
        self.shape: Optional[Tuple[int, ...]] = ()
        self.dtype = None
        self.esize = 0
        self.layout = None
        self.filters: List[Tuple[int, Tuple[int, ...]]] = []
        self.attrs: Dict[str, object] = {}
        for t, b in msgs:
            if t == 0x0001:
                self.shape = _parse_dataspace(b)
            elif t == 0x0003:
                self.dtype, self.esize = _parse_datatype(b)
            elif t == 0x0008:
                self.layout = b
            elif t == 0x000B:
                self.filters = _parse_filters(b)
            elif t == 0x000C:
                k, v = f._parse_attribute(b)
                self.attrs[k] = v

    # -- raw bytes of the dataset in file (row-major) order
    def _raw(self) -> bytes:
        f, b = self.file, self.layout
        n = int(np.prod(self.shape, dtype=np.int64)) * self.esize if self.shape is not None else 0
        if b is None:
            raise H5Error(f"{self.name}: no data layout message")
        ver = b[0]
        if ver == 3:
            cls = b[1]
            if cls == 0:                                    # compact
                sz = struct.unpack_from("<H", b, 2)[0]
                return bytes(b[4:4 + sz])
            if cls == 1:                                    # contiguous
                addr, sz = struct.unpack_from("<QQ", b, 2)
                if addr == UNDEF:
                    return bytes(n)                         # never written: fill value 0
                return f.view(addr, sz)                     # no copy: a data matrix can be gigabytes
            if cls == 2:
                rank = b[2]
                bt = struct.unpack_from("<Q", b, 3)[0]
                cdims = struct.unpack_from(f"<{rank}I", b, 11)
                return self._read_chunked(bt, cdims[:-1], n)
            raise H5Error(f"layout class {cls}")
        if ver in (1, 2):
            rank, cls = b[1], b[2]
            off = 8
            addr = UNDEF
            if cls != 0:
                addr = struct.unpack_from("<Q", b, off)[0]
                off += 8
            dims = struct.unpack_from(f"<{rank}I", b, off)
            off += 4 * rank
            if cls == 1:
                return f.view(addr, n) if addr != UNDEF else bytes(n)
            if cls == 2:
                return self._read_chunked(addr, dims[:-1] if len(dims) == len(self.shape) + 1 else dims, n)
            sz = struct.unpack_from("<I", b, off)[0]
            return bytes(b[off + 4:off + 4 + sz])
        raise H5Error(f"data layout version {ver}")

    def _read_chunked(self, btree: int, cdims: Tuple[int, ...], n: int) -> bytes:
        shape = self.shape
        out = np.zeros(shape, dtype=np.dtype(f"V{self.esize}"))
        if btree != UNDEF:
            for offsets, mask, addr, size in self.file._chunks(btree, len(shape)):
                raw = self.file.at(addr, size)
                for i in reversed(range(len(self.filters))):
                    if mask & (1 << i):
                        continue
                    fid, cd = self.filters[i]
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:                          # shuffle
                        es = cd[0] if cd else self.esize
                        raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                    elif fid == 3:                          # fletcher32: checksum trails the data
                        raw = raw[:-4]
                    else:
                        raise H5Error(f"filter {fid} is not supported")
                chunk = np.frombuffer(raw, dtype=out.dtype, count=int(np.prod(cdims))).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offsets, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out.tobytes()

    def read(self, julia: bool = True):
        if self.shape is None:
            return None
        raw = self._raw()
        if isinstance(self.dtype, _VlenStr):
            cnt = int(np.prod(self.shape, dtype=np.int64))
            vals = []
            for i in range(cnt):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                s = self.file._global_heap_object(addr, idx)[:ln] if ln else b""
                vals.append(s.decode("utf-8", "replace"))
            arr = np.array(vals, dtype=object).reshape(self.shape)
        else:
            arr = np.frombuffer(raw, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)
            if arr.dtype.kind == "S":
                arr = np.array([x.split(b"\0")[0].decode("utf-8", "replace") for x in arr.ravel()],
                               dtype=object).reshape(self.shape)
            else:
                arr = arr.astype(arr.dtype.newbyteorder("="))
        return arr.T if julia and arr.ndim > 1 else arr


def _parse_filters(b: bytes) -> List[Tuple[int, Tuple[int, ...]]]:
    ver, nf = b[0], b[1]
    off = 8 if ver == 1 else 2
    out = []
    for _ in range(nf):
        fid = struct.unpack_from("<H", b, off)[0]
        off += 2
        nlen = 0
        if ver == 1 or fid >= 256:
            nlen = struct.unpack_from("<H", b, off)[0]
            off += 2
        _flags, ncd = struct.unpack_from("<HH", b, off)
        off += 4
        if nlen:
            off += (nlen + 7) // 8 * 8 if ver == 1 else nlen
        cd = struct.unpack_from(f"<{ncd}I", b, off)
        off += 4 * ncd
        if ver == 1 and ncd % 2:
            off += 4
        out.append((fid, tuple(cd)))
    return out


class Group:
    def __init__(self, f: "File", members: Dict[str, int], msgs, name: str):
        self.file, self._members, self.name = f, members, name
        self.attrs: Dict[str, object] = {}
        for t, b in msgs:
            if t == 0x000C:
                k, v = f._parse_attribute(b)
                self.attrs[k] = v

    def keys(self) -> List[str]:
        return sorted(self._members)

    def __contains__(self, name: str) -> bool:
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._members:
                raise KeyError(path)
            node = node.file._object(node._members[part], f"{node.name.rstrip('/')}/{part}")
        return node


class File(Group):
    """Read-only view of an HDF5 file held in memory.  ``f["omic_data/data"].read()``, ``f.read("X")``."""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as fh:
                self.buf = fh.read()
        sb = -1
        off = 0
        while off < len(self.buf):                          # the superblock sits at 0, 512, 1024, ... (user block)
            if self.buf[off:off + 8] == SIGNATURE:
                sb = off
                break
            off = 512 if off == 0 else off * 2
        if sb < 0:
            raise H5Error("not an HDF5 file (no superblock signature)")
        b = self.buf
        ver = b[sb + 8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} is not supported (the reference's files are written with libhdf5's "
                          f"default format bounds: version 0)")
        so, sl = b[sb + 13], b[sb + 14]
        if (so, sl) != (8, 8):
            raise H5Error(f"size of offsets / lengths {so} / {sl} (only 8 / 8)")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, sb + 16)
        p = sb + 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", b, p)
        # libhdf5 stores the user block size as the base address and keeps every address relative to it
        if self.base == 0 and sb != 0:
            self.base = sb
        root = p + 32
        _name_off, ohdr, cache, _r = struct.unpack_from("<QQII", b, root)
        self._cache: Dict[int, object] = {}
        obj = self._object(ohdr, "/")
        if not isinstance(obj, Group):
            raise H5Error("the root object is not a group")
        super().__init__(self, obj._members, [], "/")
        self.attrs = obj.attrs

    def at(self, addr: int, n: int) -> bytes:
        a = self.base + addr
        if a + n > len(self.buf):
            raise H5Error(f"address {addr} + {n} bytes runs past the end of the file")
        return self.buf[a:a + n]

    def view(self, addr: int, n: int) -> memoryview:
        a = self.base + addr
        if a + n > len(self.buf):
            raise H5Error(f"address {addr} + {n} bytes runs past the end of the file")
        return memoryview(self.buf)[a:a + n]

    def read(self, name: str, julia: bool = True):
        obj = self[name]
        if not isinstance(obj, Dataset):
            raise KeyError(f"{name} is a group")
        return obj.read(julia=julia)

    def visit(self, group: Optional[Group] = None, prefix: str = "") -> List[str]:
        """Every dataset path below ``group`` (depth first, sorted)."""
        out = []
        g = group if group is not None else self
        for k in g.keys():
            o = g[k]
            if isinstance(o, Group):
                out += self.visit(o, f"{prefix}{k}/")
            else:
                out.append(prefix + k)
        return out

    # -- object headers ------------------------------------------------------------------------------------
    def _messages(self, addr: int) -> List[Tuple[int, bytes]]:
        hd = self.at(addr, 16)
        if hd[:4] == b"OHDR":
            raise H5Error("version 2 object headers are not supported (file written with libver='latest')")
        ver, _r, nmsg, _refs, hsize = struct.unpack_from("<BBHII", hd, 0)
        if ver != 1:
            raise H5Error(f"object header version {ver}")
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            a, n = blocks.pop(0)
            blk = self.at(a, n)
            off = 0
            while off + 8 <= n and len(msgs) < nmsg:
                t, sz, _fl = struct.unpack_from("<HHB", blk, off)
                body = blk[off + 8:off + 8 + sz]
                off += 8 + sz
                if t == 0x0010:
                    ca, cn = struct.unpack_from("<QQ", body, 0)
                    blocks.append((ca, cn))
                msgs.append((t, body))
        return msgs

    def _object(self, addr: int, name: str):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        st = [b for t, b in msgs if t == 0x0011]
        if st:
            bt, heap = struct.unpack_from("<QQ", st[0], 0)
            obj = Group(self, self._group_members(bt, heap), msgs, name)
        elif any(t == 0x0002 for t, _ in msgs):
            raise H5Error(f"{name}: link-info (new-style) groups are not supported")
        else:
            obj = Dataset(self, msgs, name)
        self._cache[addr] = obj
        return obj

    # -- groups: v1 B-tree over symbol-table nodes, names in a local heap --------------------------------------
    def _heap_data(self, heap: int) -> bytes:
        h = self.at(heap, 32)
        if h[:4] != b"HEAP":
            raise H5Error("local heap signature")
        size, _free, daddr = struct.unpack_from("<QQQ", h, 8)
        return self.at(daddr, size)

    def _group_members(self, btree: int, heap: int) -> Dict[str, int]:
        names = self._heap_data(heap)
        out: Dict[str, int] = {}

        def walk(a: int):
            nd = self.at(a, 24)
            if nd[:4] == b"SNOD":
                nsym = struct.unpack_from("<H", nd, 6)[0]
                ent = self.at(a + 8, 40 * nsym)
                for i in range(nsym):
                    noff, oaddr = struct.unpack_from("<QQ", ent, 40 * i)
                    end = names.index(b"\0", noff)
                    out[names[noff:end].decode("utf-8")] = oaddr
                return
            if nd[:4] != b"TREE" or nd[4] != 0:
                raise H5Error("group B-tree node signature / type")
            used = struct.unpack_from("<H", nd, 6)[0]
            body = self.at(a + 24, 16 * used + 8)
            for i in range(used):
                walk(struct.unpack_from("<Q", body, 16 * i + 8)[0])

        if btree != UNDEF:
            walk(btree)
        return out

    # -- chunked datasets: v1 B-tree of raw-data chunks --------------------------------------------------------
    def _chunks(self, btree: int, rank: int):
        ksize = 8 + 8 * (rank + 1)

        def walk(a: int):
            nd = self.at(a, 24)
            if nd[:4] != b"TREE" or nd[4] != 1:
                raise H5Error("chunk B-tree node signature / type")
            level, used = nd[5], struct.unpack_from("<H", nd, 6)[0]
            body = self.at(a + 24, used * (ksize + 8) + ksize)
            for i in range(used):
                o = i * (ksize + 8)
                size, mask = struct.unpack_from("<II", body, o)
                offs = struct.unpack_from(f"<{rank}Q", body, o + 8)
                child = struct.unpack_from("<Q", body, o + ksize)[0]
                if level == 0:
                    yield offs, mask, child, size
                else:
                    yield from walk(child)

        yield from walk(btree)

    # -- global heap (variable-length data) --------------------------------------------------------------------
    def _global_heap_object(self, addr: int, index: int) -> bytes:
        key = ("gcol", addr)
        if key not in self._cache:
            hd = self.at(addr, 16)
            if hd[:4] != b"GCOL":
                raise H5Error("global heap signature")
            size = struct.unpack_from("<Q", hd, 8)[0]
            blk = self.at(addr, size)
            objs = {}
            off = 16
            while off + 16 <= size:
                idx, _rc, _r, osz = struct.unpack_from("<HHIQ", blk, off)
                if idx == 0:
                    break
                objs[idx] = blk[off + 16:off + 16 + osz]
                off += 16 + (osz + 7) // 8 * 8
            self._cache[key] = objs
        return self._cache[key][index]

    # -- attributes ------------------------------------------------------------------------------------------
    def _parse_attribute(self, b: bytes):
        ver = b[0]
        if ver != 1:
            return f"<attribute v{ver}>", None
        nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
        pad = lambda n: (n + 7) // 8 * 8
        off = 8
        name = b[off:off + nsz].split(b"\0")[0].decode("utf-8")
        off += pad(nsz)
        dt, es = _parse_datatype(b[off:off + tsz])
        off += pad(tsz)
        shape = _parse_dataspace(b[off:off + ssz])
        off += pad(ssz)
        if shape is None or isinstance(dt, _VlenStr):
            return name, None
        cnt = int(np.prod(shape, dtype=np.int64))
        arr = np.frombuffer(b[off:off + cnt * es], dtype=dt, count=cnt).reshape(shape)
        if arr.dtype.kind == "S":
            vals = [x.split(b"\0")[0].decode("utf-8", "replace") for x in arr.ravel()]
            return name, (vals[0] if shape == () or cnt == 1 else vals)
        return name, (arr.item() if cnt == 1 else arr.copy())


def read(path, name: str, julia: bool = True):
    """``h5read(path, name)`` of HDF5.jl."""
    return File(path).read(name, julia=julia)


# ------------------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------------------

def _pad8(b: bytes) -> bytes:
    return b + bytes(-len(b) % 8)


def _dtype_message(dt: np.dtype) -> bytes:
    """Datatype message, version 1, little endian (what libhdf5 writes for the native types of x86-64)."""
    s = dt.itemsize
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits, 0, 0, s, 0, 8 * s)
    if dt.kind == "f":
        exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[s]
        # bit field: little endian, mantissa normalisation 2 (implied leading one), sign bit position in byte 1
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 8 * s - 1, 0, s, 0, 8 * s, man_bits, exp_bits, 0, man_bits, bias)
    raise H5Error(f"cannot write dtype {dt}")


# class 9 version 1: variable-length STRING, null-terminated, UTF-8; 16-byte elements (length, heap address, index);
# base type = unsigned 8-bit integer (what libhdf5's H5Tset_size(H5T_VARIABLE) gives a string type)
_VLEN_STR_MESSAGE = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBIHH", 0x10, 0, 0, 0, 1, 0, 8)


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHBBBB", mtype, len(body), flags, 0, 0, 0) + body


def _object_header(messages: List[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + bytes(4) + body


class Writer:
    """Collects datasets under slash-separated names, then lays the file out in one pass.

        w = Writer(); w.write("X", X); w.write("theta/values_1", v); w.write("feature_ids", ["a", "b"]); w.save(path)
    """
    LEAF_K, INTERNAL_K = 4, 16                    # libhdf5 defaults (symbol-table node: 2 * 4 entries, B-tree node: 2 * 16 children)
    GCOL_OBJECTS = 4096                           # strings per global-heap collection (the object index is 16 bits wide)

    def __init__(self):
        self.tree: Dict[str, object] = {}

    def write(self, name: str, value, julia: bool = True):
        parts = [p for p in name.split("/") if p]
        if not parts:
            raise H5Error("empty dataset name")
        node = self.tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise H5Error(f"{name}: {p} is a dataset")
        if parts[-1] in node:
            raise H5Error(f"{name} exists")                 # like HDF5.jl: a name is written once
        if isinstance(value, (list, tuple)) and all(isinstance(v, str) for v in value):
            value = np.array(list(value), dtype=object)
        arr = np.asarray(value)
        if arr.dtype.kind in "OUS":
            arr = np.array([v.decode("utf-8") if isinstance(v, bytes) else str(v) for v in arr.ravel()],
                           dtype=object).reshape(arr.shape)
        elif arr.dtype.kind == "b":
            arr = arr.astype(np.uint8)                      # HDF5.jl stores Bool as a 1-byte integer
        elif arr.dtype.kind not in "iuf":
            raise H5Error(f"{name}: dtype {arr.dtype} cannot be written")
        if julia and arr.ndim > 1:
            arr = arr.T
        node[parts[-1]] = np.require(arr, requirements="C") if arr.dtype != object else arr   # (keeps 0-d arrays 0-d)
        return self

    # -- layout ----------------------------------------------------------------------------------------------
    def _alloc(self, data: bytes) -> int:
        self.buf += bytes(-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def _dataset(self, arr: np.ndarray) -> int:
        shape = arr.shape
        if arr.dtype == object:                             # variable-length UTF-8 strings: one global-heap collection
            flat = [s.encode("utf-8") for s in arr.ravel()]
            raw = b""
            for c0 in range(0, len(flat), self.GCOL_OBJECTS):               # one global-heap collection per block of strings
                part = flat[c0:c0 + self.GCOL_OBJECTS]
                objs = b"".join(struct.pack("<HHIQ", i + 1, 0, 0, len(s)) + _pad8(s) for i, s in enumerate(part))
                size = max(4096, 16 + len(objs) + 16)
                free = size - 16 - len(objs)
                col = b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, size) + objs
                col += struct.pack("<HHIQ", 0, 0, 0, free) + bytes(free - 16)   # object 0: the free space (its size counts its header)
                gaddr = self._alloc(col)
                raw += b"".join(struct.pack("<IQI", len(s), gaddr if s else 0, i + 1 if s else 0) for i, s in enumerate(part))
            tmsg = _VLEN_STR_MESSAGE
        else:
            arr = arr.astype(arr.dtype.newbyteorder("<"))
            raw = arr.tobytes()
            tmsg = _dtype_message(arr.dtype)
        daddr = self._alloc(raw) if raw else UNDEF
        space = struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)
        msgs = [_message(0x0001, space),
                _message(0x0003, tmsg, flags=1),                                     # constant
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=1),     # fill value v2: late allocation, written if set, the default value
                _message(0x0008, struct.pack("<BBQQ", 3, 1, daddr, len(raw)))]       # layout v3, contiguous
        return self._alloc(_object_header(msgs))

    def _group(self, members: Dict[str, object]) -> Tuple[int, int, int]:
        """Writes the members, then the group's local heap, symbol-table nodes, B-tree and object header.
        Returns (object header, B-tree, heap) addresses."""
        entries = []                                        # (name bytes, object header address, cache type, scratch)
        for name in members:
            v = members[name]
            if isinstance(v, dict):
                oh, bt, hp = self._group(v)
                entries.append((name.encode("utf-8"), oh, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((name.encode("utf-8"), self._dataset(v), 0, bytes(16)))
        entries.sort(key=lambda e: e[0])                    # strcmp order, as the B-tree keys require
        heap = bytearray(8)                                 # offset 0: the empty string (the B-tree's left-most key)
        offs = []
        for nm, *_ in entries:
            offs.append(len(heap))
            heap += _pad8(nm + b"\0")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16) + bytes(0)        # one free block closing the segment: next = 1 (none), size 16
        daddr = self._alloc(bytes(heap))
        haddr = self._alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap), free_off, daddr))
        # symbol-table nodes of at most 2 * LEAF_K entries, each padded to its full size
        cap = 2 * self.LEAF_K
        level: List[Tuple[int, int]] = []                   # (address of child, heap offset of its largest name)
        for i in range(0, len(entries), cap):
            part = entries[i:i + cap]
            nd = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for j, (nm, oh, cache, scratch) in enumerate(part):
                nd += struct.pack("<QQII", offs[i + j], oh, cache, 0) + scratch
            nd += bytes(40 * (cap - len(part)))
            level.append((self._alloc(nd), offs[i + len(part) - 1]))
        # B-tree levels of at most 2 * INTERNAL_K children
        fan = 2 * self.INTERNAL_K
        depth = 0
        while True:
            nxt = []
            for i in range(0, max(len(level), 1), fan):
                part = level[i:i + fan]
                nd = b"TREE" + struct.pack("<BBHQQ", 0, depth, len(part), UNDEF, UNDEF)
                # key 0 bounds the node from the left: the empty string, or the largest name of the left sibling
                nd += struct.pack("<Q", level[i - 1][1] if i > 0 else 0)
                for child, key in part:
                    nd += struct.pack("<QQ", child, key)
                nd += bytes(16 * (fan - len(part)))
                nxt.append((self._alloc(nd), part[-1][1] if part else 0))
            if len(nxt) == 1:
                btree = nxt[0][0]
                break
            # siblings of one level are chained left to right
            for j, (a, _k) in enumerate(nxt):
                left = nxt[j - 1][0] if j > 0 else UNDEF
                right = nxt[j + 1][0] if j + 1 < len(nxt) else UNDEF
                self.buf[a + 8:a + 24] = struct.pack("<QQ", left, right)
            level, depth = nxt, depth + 1
        oh = self._alloc(_object_header([_message(0x0011, struct.pack("<QQ", btree, haddr))]))
        return oh, btree, haddr

    def tobytes(self) -> bytes:
        self.buf = bytearray(96)                            # superblock v0 with its root symbol-table entry
        oh, bt, hp = self._group(self.tree)
        self.buf += bytes(-len(self.buf) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)

    def save(self, path):
        data = self.tobytes()
        with open(path, "wb") as fh:
            fh.write(data)
        return path
