"""Minimal MAT v7.3 (HDF5) reader/writer for the prior fixtures and salmap outputs.

The reference loads its prior inputs with ``hdf5storage.loadmat(path)["PriorMaps"]``
(/root/reference/utils_data.py:459, :587) and writes ``{'salmap': uint8 (H,W,1,F)}`` with
``hdf5storage.savemat`` (/root/reference/Demo_Test.py:94-95).  Neither hdf5storage nor h5py is
available offline, so this module parses exactly the HDF5 subset those files use:

  * 512-byte MATLAB user block, superblock version 0, 8-byte offsets/lengths
  * root group through a v1 B-tree (node type 0) + local heap + SNOD symbol nodes
  * version-1 object headers (with continuation blocks)
  * dataspace v1/v2, fixed-point / floating-point datatypes (little endian)
  * data layout v3: contiguous or chunked (v1 B-tree node type 1)
  * filter pipeline v1/v2 with shuffle (2), deflate (1) and fletcher32 (3)

MATLAB stores arrays column-major, so the HDF5 dataset dims are the reverse of the numpy shape:
``loadmat`` returns ``dataset.T`` exactly as hdf5storage does.

The writer emits an uncompressed, contiguous-layout file of the same structure (one or more
numeric datasets in the root group, ``MATLAB_class`` attribute attached) that this reader, h5py,
hdf5storage and MATLAB can open.
"""
from __future__ import annotations

import struct
import time
import zlib
from typing import Dict

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b"\x89HDF\r\n\x1a\n"


class Mat73Error(ValueError):
    pass


class _File:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.base = buf.find(_SIG)
        if self.base < 0 or self.base % 512:
            raise Mat73Error("not an HDF5 / MAT v7.3 file")
        sb = self.base
        ver = buf[sb + 8]
        if ver != 0:
            raise Mat73Error(f"unsupported superblock version {ver}")
        if buf[sb + 13] != 8 or buf[sb + 14] != 8:
            raise Mat73Error("only 8-byte offsets/lengths supported")
        # sig 8 | versions 5 | sizes 2 | rsvd 1 | leafK 2 | internalK 2 | flags 4 | 4 addresses
        (self.base_addr,) = struct.unpack_from("<Q", buf, sb + 24)
        root_entry = sb + 24 + 32
        self.root = self._symtab_entry(root_entry)

    # -- primitives -------------------------------------------------------------------------
    def at(self, addr: int) -> int:
        return self.base_addr + addr

    def _symtab_entry(self, off: int):
        name_off, ohdr, cache, _ = struct.unpack_from("<QQII", self.buf, off)
        scratch = self.buf[off + 24: off + 40]
        return dict(name_off=name_off, ohdr=ohdr, cache=cache, scratch=scratch)

    # -- groups -----------------------------------------------------------------------------
    def _heap_data(self, heap_addr: int) -> int:
        o = self.at(heap_addr)
        if self.buf[o:o + 4] != b"HEAP":
            raise Mat73Error("bad local heap")
        _, _, data_addr = struct.unpack_from("<QQQ", self.buf, o + 8)
        return self.at(data_addr)

    def _cstr(self, off: int) -> str:
        end = self.buf.index(b"\0", off)
        return self.buf[off:end].decode("ascii")

    def _walk_group(self, btree_addr: int, heap_data: int, out: Dict[str, int]):
        o = self.at(btree_addr)
        if self.buf[o:o + 4] != b"TREE":
            raise Mat73Error("bad group b-tree")
        ntype, level, used = struct.unpack_from("<BBH", self.buf, o + 4)
        if ntype != 0:
            raise Mat73Error("expected group b-tree node")
        p = o + 24
        for i in range(used):
            (child,) = struct.unpack_from("<Q", self.buf, p + 8 + 16 * i)
            if level > 0:
                self._walk_group(child, heap_data, out)
            else:
                s = self.at(child)
                if self.buf[s:s + 4] != b"SNOD":
                    raise Mat73Error("bad symbol node")
                (nsym,) = struct.unpack_from("<H", self.buf, s + 6)
                for j in range(nsym):
                    e = self._symtab_entry(s + 8 + 40 * j)
                    out[self._cstr(heap_data + e["name_off"])] = e["ohdr"]

    def root_members(self) -> Dict[str, int]:
        if self.root["cache"] == 1:
            bt, hp = struct.unpack_from("<QQ", self.root["scratch"], 0)
        else:
            bt = hp = None
            for mtype, body in self._messages(self.root["ohdr"]):
                if mtype == 0x11:
                    bt, hp = struct.unpack_from("<QQ", body, 0)
            if bt is None:
                raise Mat73Error("root group has no symbol table")
        out: Dict[str, int] = {}
        self._walk_group(bt, self._heap_data(hp), out)
        return out

    # -- object headers ---------------------------------------------------------------------
    def _messages(self, ohdr_addr: int):
        o = self.at(ohdr_addr)
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.buf, o)
        if ver != 1:
            raise Mat73Error(f"unsupported object header version {ver}")
        blocks = [(o + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and seen < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.buf, p)
                body = self.buf[p + 8: p + 8 + msize]
                p += 8 + msize
                seen += 1
                if mtype == 0x10:
                    coff, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.at(coff), clen))
                else:
                    yield mtype, body

    # -- datasets ---------------------------------------------------------------------------
    def read_dataset(self, ohdr_addr: int) -> np.ndarray:
        dims = dtype = layout = None
        filters = []
        for mtype, body in self._messages(ohdr_addr):
            if mtype == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", body, 0)
                off = 8 if ver == 1 else 4
                dims = struct.unpack_from("<%dQ" % rank, body, off)
            elif mtype == 0x03:
                cls = body[0] & 0x0F
                bits0 = body[1]
                (size,) = struct.unpack_from("<I", body, 4)
                if bits0 & 1:
                    raise Mat73Error("big-endian data not supported")
                if cls == 0:
                    dtype = np.dtype("<%s%d" % ("i" if bits0 & 8 else "u", size))
                elif cls == 1:
                    dtype = np.dtype("<f%d" % size)
                else:
                    raise Mat73Error(f"unsupported datatype class {cls}")
            elif mtype == 0x08:
                ver, lclass = body[0], body[1]
                if ver != 3:
                    raise Mat73Error(f"unsupported layout version {ver}")
                if lclass == 1:
                    addr, size = struct.unpack_from("<QQ", body, 2)
                    layout = ("contiguous", addr, size)
                elif lclass == 2:
                    nd = body[2]
                    (addr,) = struct.unpack_from("<Q", body, 3)
                    cdims = struct.unpack_from("<%dI" % nd, body, 11)
                    layout = ("chunked", addr, cdims)
                else:
                    raise Mat73Error("compact layout not supported")
            elif mtype == 0x0B:
                ver, nf = body[0], body[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid, nlen, _fl, ncv = struct.unpack_from("<HHHH", body, p)
                    p += 8
                    if ver == 1:
                        p += (nlen + 7) // 8 * 8
                    elif fid >= 256:
                        p += nlen
                    cvals = struct.unpack_from("<%dI" % ncv, body, p)
                    p += 4 * ncv
                    if ver == 1 and ncv % 2:
                        p += 4
                    filters.append((fid, cvals))
        if dims is None or dtype is None or layout is None:
            raise Mat73Error("dataset header incomplete")
        if layout[0] == "contiguous":
            _, addr, size = layout
            n = int(np.prod(dims)) if dims else 1
            if addr == _UNDEF:
                return np.zeros(dims, dtype)
            o = self.at(addr)
            return np.frombuffer(self.buf, dtype, n, o).reshape(dims).copy()
        _, addr, cdims = layout
        out = np.zeros(dims, dtype)
        self._read_chunks(addr, len(cdims), cdims[:-1], filters, dtype, out)
        return out

    def _read_chunks(self, addr, nd, cshape, filters, dtype, out):
        o = self.at(addr)
        if self.buf[o:o + 4] != b"TREE":
            raise Mat73Error("bad chunk b-tree")
        ntype, level, used = struct.unpack_from("<BBH", self.buf, o + 4)
        if ntype != 1:
            raise Mat73Error("expected chunk b-tree node")
        keysz = 8 + 8 * nd
        p = o + 24
        for i in range(used):
            csize, fmask = struct.unpack_from("<II", self.buf, p)
            offs = struct.unpack_from("<%dQ" % nd, self.buf, p + 8)
            (child,) = struct.unpack_from("<Q", self.buf, p + keysz)
            p += keysz + 8
            if level > 0:
                self._read_chunks(child, nd, cshape, filters, dtype, out)
                continue
            c = self.at(child)
            raw = self.buf[c:c + csize]
            for k in range(len(filters) - 1, -1, -1):
                fid, _cv = filters[k]
                if fmask & (1 << k):
                    continue
                if fid == 3:      # fletcher32: trailing 4-byte checksum
                    raw = raw[:-4]
                elif fid == 1:    # deflate
                    raw = zlib.decompress(raw)
                elif fid == 2:    # byte shuffle
                    es = dtype.itemsize
                    a = np.frombuffer(raw, np.uint8)
                    n = a.size // es
                    raw = a[: n * es].reshape(es, n).T.tobytes() + a[n * es:].tobytes()
                else:
                    raise Mat73Error(f"unsupported filter {fid}")
            chunk = np.frombuffer(raw, dtype, int(np.prod(cshape))).reshape(cshape)
            sl_out, sl_in = [], []
            for d, (o0, cs) in enumerate(zip(offs[:-1], cshape)):
                hi = min(o0 + cs, out.shape[d])
                sl_out.append(slice(o0, hi))
                sl_in.append(slice(0, hi - o0))
            out[tuple(sl_out)] = chunk[tuple(sl_in)]


def loadmat(path: str) -> Dict[str, np.ndarray]:
    """``hdf5storage.loadmat`` replacement for plain numeric variables (utils_data.py:459,587)."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    out = {}
    for name, ohdr in f.root_members().items():
        if name.startswith("#"):
            continue
        try:
            arr = f.read_dataset(ohdr)
        except Mat73Error:
            continue
        out[name] = arr.T  # MATLAB column-major → numpy shape
    return out


# ---------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------
_MATLAB_CLASS = {"uint8": b"uint8", "int8": b"int8", "uint16": b"uint16", "int16": b"int16",
                 "uint32": b"uint32", "int32": b"int32", "uint64": b"uint64", "int64": b"int64",
                 "float32": b"single", "float64": b"double"}


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dtype_msg_body(dt: np.dtype) -> bytes:
    size = dt.itemsize
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10, bits, 0, 0, size) + struct.pack("<HH", 0, size * 8)
    if dt.kind == "f":
        if size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            b1 = 31
        else:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            b1 = 63
        return struct.pack("<BBBBI", 0x11, 0x20, b1, 0, size) + props
    raise Mat73Error(f"unsupported dtype {dt}")


def _string_dtype_body(n: int) -> bytes:
    return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, n)  # class 3 string, null-terminated, ASCII


def savemat(path: str, variables: Dict[str, np.ndarray]) -> None:
    """``hdf5storage.savemat`` replacement for numeric arrays (Demo_Test.py:94-95)."""
    names = sorted(variables)
    if len(names) > 8:
        raise Mat73Error("at most 8 variables per file")
    base = 512
    hdr = ("MATLAB 7.3 MAT-file, Platform: uavsal-b200, Created on: %s HDF5 schema 1.00 ."
           % time.strftime("%a %b %d %H:%M:%S %Y")).encode("ascii")
    user = hdr.ljust(116, b" ") + b"\0" * 8 + struct.pack("<H", 0x0200) + b"IM"
    user = user.ljust(512, b"\0")

    # layout of the HDF5 part (addresses relative to base)
    sb_size = 96
    root_ohdr = sb_size                       # 0x60
    root_ohdr_size = 16 + len(_msg(0x11, b"\0" * 16))
    btree = root_ohdr + root_ohdr_size
    btree_size = 24 + (2 * 16 + 1) * 8 * 2    # generous: K=16
    heap = btree + btree_size
    heap_data_size = 8 + sum(len(n) + 1 + 7 for n in names) // 8 * 8 + 8
    heap_hdr = 32
    heap_data = heap + heap_hdr
    snod = heap_data + heap_data_size
    snod_size = 8 + 40 * 8                    # 2*leafK(4) entries
    cursor = snod + snod_size

    # heap strings
    hd = bytearray(b"\0" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(hd)
        hd += _pad8(n.encode("ascii") + b"\0")
    free_off = len(hd)
    hd = hd.ljust(heap_data_size, b"\0")
    if heap_data_size - free_off >= 16:
        struct.pack_into("<QQ", hd, free_off, 1, heap_data_size - free_off)
    else:
        free_off = _UNDEF

    objs = []
    entries = []
    for n in names:
        arr = np.asarray(variables[n])
        if arr.dtype.name not in _MATLAB_CLASS:
            raise Mat73Error(f"unsupported dtype {arr.dtype}")
        data = np.ascontiguousarray(arr.T).astype(arr.dtype.newbyteorder("<"), copy=False)
        dims = data.shape if data.ndim else (1,)
        cls = _MATLAB_CLASS[arr.dtype.name]
        msgs = b""
        msgs += _msg(0x01, struct.pack("<BBB5x", 1, len(dims), 0) + struct.pack("<%dQ" % len(dims), *dims))
        msgs += _msg(0x03, _dtype_msg_body(data.dtype), flags=1)
        msgs += _msg(0x05, struct.pack("<BBBB", 2, 2, 2, 0))          # fill value v2: never written
        attr_name = b"MATLAB_class\0"
        attr_dt = _string_dtype_body(len(cls))
        attr_ds = struct.pack("<BBB5x", 1, 0, 0)                      # scalar dataspace
        attr = struct.pack("<BxHHH", 1, len(attr_name), len(attr_dt), len(attr_ds))
        attr += _pad8(attr_name) + _pad8(attr_dt) + _pad8(attr_ds) + cls
        msgs += _msg(0x0C, attr)
        lay_len = len(_msg(0x08, struct.pack("<BBQQ", 3, 1, 0, 0)))
        ohdr_size = 16 + len(msgs) + lay_len
        ohdr_addr = cursor
        data_addr = (ohdr_addr + ohdr_size + 7) // 8 * 8
        msgs += _msg(0x08, struct.pack("<BBQQ", 3, 1, data_addr, data.nbytes))
        ohdr = struct.pack("<BBHII4x", 1, 0, 5, 1, len(msgs)) + msgs
        blob = ohdr.ljust(data_addr - ohdr_addr, b"\0") + data.tobytes()
        blob = _pad8(blob)
        objs.append(blob)
        entries.append(struct.pack("<QQII16x", name_off[n], ohdr_addr, 0, 0))
        cursor = ohdr_addr + len(blob)
    eof = cursor

    sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", base, _UNDEF, base + eof, _UNDEF)
    sb += struct.pack("<QQII", 0, root_ohdr, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == sb_size
    root = struct.pack("<BBHII4x", 1, 0, 1, 1, root_ohdr_size - 16) + _msg(0x11, struct.pack("<QQ", btree, heap))
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, _UNDEF, _UNDEF)
    bt += struct.pack("<QQQ", 0, snod, name_off[names[-1]] if names else 0)
    bt = bt.ljust(btree_size, b"\0")
    hp = b"HEAP" + struct.pack("<B3xQQQ", 0, heap_data_size, free_off, heap_data)
    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names)) + b"".join(entries)
    sn = sn.ljust(snod_size, b"\0")
    with open(path, "wb") as fh:
        fh.write(user + sb + root + bt + hp + bytes(hd) + sn + b"".join(objs))
