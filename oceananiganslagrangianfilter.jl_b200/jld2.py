"""Minimal reader for the JLD2 (HDF5-subset) files Oceananigans' ``JLD2Writer`` produces.

Host-side I/O for the offline filter: the reference reads its input through ``jldopen`` /
``FieldTimeSeries`` (reference src/Utils/lagrangian_filter_utils.jl:117-124, 201-215).  Neither
Julia nor h5py exists in this image, so this module decodes the small HDF5 subset those files
use: superblock v2/v3 behind the 512-byte JLD2 text header, v2 object
headers with continuation blocks, compact link messages for groups, and contiguous or compact
dataset layouts holding IEEE floats or integers.  Julia-serialised structs (``serialized/*``)
are reported as opaque and skipped.

Arrays come back in *Julia order*: ``timeseries/u/0`` of a field whose Julia ``parent`` has size
``(Nx+2Hx, Ny+2Hy, Nz+2Hz)`` is returned as a Fortran-ordered numpy array of that same shape
(HDF5 lists the dimensions reversed).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

_SIGNATURE = b"\x89HDF\r\n\x1a\n"

MSG_DATASPACE = 0x01
MSG_DATATYPE = 0x03
MSG_LINK = 0x06
MSG_LAYOUT = 0x08
MSG_CONTINUATION = 0x10


class JLD2FormatError(ValueError):
    """Raised when the file uses an HDF5 feature outside the supported subset."""


@dataclass
class _Dataset:
    dims: Optional[Tuple[int, ...]]
    type_class: Optional[int]
    type_size: Optional[int]
    layout: Optional[Tuple]


class JLD2File:
    """Read-only view of a JLD2 file.  ``f["timeseries/u/0"]`` returns a numpy array (or a
    Python scalar for rank-0 datasets); ``f.keys("timeseries/t")`` lists a group."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as fh:
            self._buf = fh.read()
        sb = self._buf.find(_SIGNATURE)
        if sb < 0:
            raise JLD2FormatError(f"{path}: no HDF5 superblock signature")
        version = self._buf[sb + 8]
        if version not in (2, 3):
            raise JLD2FormatError(f"{path}: superblock version {version} unsupported")
        if self._buf[sb + 9] != 8 or self._buf[sb + 10] != 8:
            raise JLD2FormatError(f"{path}: only 8-byte offsets/lengths supported")
        self._base, _ext, _eof, self._root = struct.unpack_from("<QQQQ", self._buf, sb + 12)
        self._group_cache: Dict[int, Dict[str, int]] = {}

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr: int) -> Iterator[Tuple[int, bytes]]:
        buf = self._buf
        pos = addr + self._base
        if buf[pos:pos + 4] != b"OHDR":
            raise JLD2FormatError(f"object header v2 expected at {pos}")
        flags = buf[pos + 5]
        pos += 6
        if flags & 0x20:
            pos += 16          # access/modification/change/birth times
        if flags & 0x10:
            pos += 4           # max compact / min dense attribute counts
        width = 1 << (flags & 0x03)
        chunk0 = int.from_bytes(buf[pos:pos + width], "little")
        pos += width
        pending: List[Tuple[int, int]] = [(pos, pos + chunk0)]
        tracked = bool(flags & 0x04)
        while pending:
            pos, end = pending.pop(0)
            while pos + 4 <= end:
                mtype = buf[pos]
                msize = struct.unpack_from("<H", buf, pos + 1)[0]
                pos += 4
                if tracked:
                    pos += 2
                body = buf[pos:pos + msize]
                pos += msize
                if mtype == MSG_CONTINUATION:
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    caddr += self._base
                    if buf[caddr:caddr + 4] != b"OCHK":
                        raise JLD2FormatError("bad continuation block")
                    pending.append((caddr + 4, caddr + clen - 4))
                elif mtype != 0:
                    yield mtype, body

    def _links(self, addr: int) -> Dict[str, int]:
        cached = self._group_cache.get(addr)
        if cached is not None:
            return cached
        out: Dict[str, int] = {}
        for mtype, body in self._messages(addr):
            if mtype != MSG_LINK:
                continue
            flags = body[1]
            pos = 2
            link_type = 0
            if flags & 0x08:
                link_type = body[pos]
                pos += 1
            if flags & 0x04:
                pos += 8
            if flags & 0x10:
                pos += 1
            width = 1 << (flags & 0x03)
            nlen = int.from_bytes(body[pos:pos + width], "little")
            pos += width
            name = body[pos:pos + nlen].decode()
            pos += nlen
            if link_type == 0:   # hard link
                out[name] = struct.unpack_from("<Q", body, pos)[0]
        self._group_cache[addr] = out
        return out

    def _describe(self, addr: int) -> _Dataset:
        dims = tclass = tsize = layout = None
        for mtype, body in self._messages(addr):
            if mtype == MSG_DATASPACE:
                version, rank = body[0], body[1]
                offset = 4 if version == 2 else 8
                dims = struct.unpack_from(f"<{rank}Q", body, offset) if rank else ()
            elif mtype == MSG_DATATYPE:
                tclass = body[0] & 0x0F
                tsize = struct.unpack_from("<I", body, 4)[0]
            elif mtype == MSG_LAYOUT:
                lclass = body[1]
                if lclass == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", body, 2)
                elif lclass == 0:
                    n = struct.unpack_from("<H", body, 2)[0]
                    layout = ("compact", bytes(body[4:4 + n]))
                else:
                    layout = ("chunked",)
        return _Dataset(dims, tclass, tsize, layout)

    # ------------------------------------------------------------------ public API
    def _resolve(self, path: str) -> int:
        addr = self._root
        for part in [p for p in path.split("/") if p]:
            links = self._links(addr)
            if part not in links:
                raise KeyError(f"{path!r}: no entry {part!r} in {self.path}")
            addr = links[part]
        return addr

    def __contains__(self, path: str) -> bool:
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def keys(self, path: str = "") -> List[str]:
        return list(self._links(self._resolve(path)).keys())

    def is_group(self, path: str) -> bool:
        return bool(self._links(self._resolve(path)))

    def __getitem__(self, path: str) -> Union[np.ndarray, float, int]:
        ds = self._describe(self._resolve(path))
        if ds.layout is None or ds.type_class is None:
            raise JLD2FormatError(f"{path!r} is not a plain dataset")
        if ds.layout[0] == "contiguous":
            start = ds.layout[1] + self._base
            raw = self._buf[start:start + ds.layout[2]]
        elif ds.layout[0] == "compact":
            raw = ds.layout[1]
        else:
            raise JLD2FormatError(f"{path!r}: chunked layout unsupported")
        if ds.type_class == 1 and ds.type_size in (4, 8):
            dtype = np.dtype("<f4" if ds.type_size == 4 else "<f8")
        elif ds.type_class == 0 and ds.type_size in (1, 2, 4, 8):
            dtype = np.dtype(f"<i{ds.type_size}")
        else:
            raise JLD2FormatError(f"{path!r}: opaque/compound datatype (class {ds.type_class}, size {ds.type_size})")
        flat = np.frombuffer(raw, dtype=dtype)
        if not ds.dims:
            return flat[0].item()
        # HDF5 dims are the reverse of the Julia dims; hand back the Julia-shaped array.
        return flat.reshape(tuple(reversed(ds.dims)), order="F")

    def iterations(self, name: str = "t") -> List[int]:
        """Iteration keys of ``timeseries/<name>`` in file order (the reference does
        ``parse.(Int, keys(file["timeseries/t"]))``, utils:120)."""
        return [int(k) for k in self.keys(f"timeseries/{name}") if k != "serialized"]

    def times(self) -> np.ndarray:
        return np.array([self[f"timeseries/t/{it}"] for it in self.iterations("t")], dtype=np.float64)

    def series(self, name: str, dtype=np.float64) -> List[np.ndarray]:
        """All snapshots of ``timeseries/<name>`` (parent arrays, halos included), file order."""
        return [np.asarray(self[f"timeseries/{name}/{it}"], dtype=dtype, order="F") for it in self.iterations(name)]


# ---------------------------------------------------------------------------------------------- writer
_M32 = 0xFFFFFFFF


def _rot(x: int, k: int) -> int:
    return ((x << k) | (x >> (32 - k))) & _M32


def lookup3(data: bytes, initval: int = 0) -> int:
    """Bob Jenkins' lookup3 `hashlittle`, the checksum HDF5 puts behind superblocks and version-2 object headers
    (checked against the checksums stored in the reference's test/data/*.jld2)."""
    n = len(data)
    a = b = c = (0xDEADBEEF + n + initval) & _M32
    i = 0
    while n > 12:
        a = (a + int.from_bytes(data[i:i + 4], "little")) & _M32
        b = (b + int.from_bytes(data[i + 4:i + 8], "little")) & _M32
        c = (c + int.from_bytes(data[i + 8:i + 12], "little")) & _M32
        a = (a - c) & _M32; a ^= _rot(c, 4); c = (c + b) & _M32
        b = (b - a) & _M32; b ^= _rot(a, 6); a = (a + c) & _M32
        c = (c - b) & _M32; c ^= _rot(b, 8); b = (b + a) & _M32
        a = (a - c) & _M32; a ^= _rot(c, 16); c = (c + b) & _M32
        b = (b - a) & _M32; b ^= _rot(a, 19); a = (a + c) & _M32
        c = (c - b) & _M32; c ^= _rot(b, 4); b = (b + a) & _M32
        i += 12
        n -= 12
    if n == 0:
        return c
    t = bytes(data[i:]) + b"\0" * (12 - n)
    a = (a + int.from_bytes(t[0:4], "little")) & _M32
    b = (b + int.from_bytes(t[4:8], "little")) & _M32
    c = (c + int.from_bytes(t[8:12], "little")) & _M32
    c ^= b; c = (c - _rot(b, 14)) & _M32
    a ^= c; a = (a - _rot(c, 11)) & _M32
    b ^= a; b = (b - _rot(a, 25)) & _M32
    c ^= b; c = (c - _rot(b, 16)) & _M32
    a ^= c; a = (a - _rot(c, 4)) & _M32
    b ^= a; b = (b - _rot(a, 14)) & _M32
    c ^= b; c = (c - _rot(b, 24)) & _M32
    return c


_HEADER_TEXT = b"HDF5-based Julia Data Format, version 0.2.0\x00 (lfb200, Julia-compatible layout, 64-bit LE)\x00"
_BASE = 512
# datatype messages exactly as JLD2.jl writes them (version 3, class 1 IEEE float / class 0 fixed point)
_DT_F64 = bytes.fromhex("31203f000800000000004000340b0034ff030000")
_DT_F32 = bytes.fromhex("31201f000400000000002000170800177f000000")
_DT_I64 = bytes.fromhex("30080000080000000000" "4000")          # class 0, signed, size 8, bit offset 0, precision 64


def _message(mtype: int, body: bytes) -> bytes:
    return struct.pack("<BHB", mtype, len(body), 0) + body


def _object_header(messages: List[bytes]) -> bytes:
    """Version-2 object header with one chunk (4-byte chunk size field) and its lookup3 checksum."""
    body = b"".join(messages)
    head = b"OHDR" + bytes([2, 0x02]) + struct.pack("<I", len(body)) + body
    return head + struct.pack("<I", lookup3(head))


class JLD2Writer:
    """Writes the HDF5 subset JLD2.jl itself emits for plain arrays and scalars: a 512-byte
    text header, a version-2 superblock, version-2 object headers with lookup3 checksums, groups as compact link
    messages, contiguous datasets of IEEE floats / Int64 with the dimensions listed reversed (Julia order), compact
    rank-0 scalars.  `write(path, value)` with `path` like "timeseries/u/12"; groups are created on the way.  The
    file is laid out when `close()` is called (or the `with` block ends).

    Julia-serialised structs (`serialized/grid`, boundary conditions ...) need Julia's type system and are not
    written: `jldopen(file)["timeseries/u/12"]` works from Julia, `FieldTimeSeries(file, "u")` needs the grid passed
    in.  JLD2File (above) reads everything written here."""

    def __init__(self, path: str):
        self.path = path
        self.tree: Dict[str, object] = {}
        self.closed = False

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *a):
        if exc_type is None:
            self.close()

    def write(self, path: str, value):
        parts = [p for p in path.split("/") if p]
        if not parts:
            raise ValueError("empty dataset path")
        node = self.tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise ValueError(f"{path!r}: {p!r} is a dataset, not a group")
        if isinstance(value, (bool, np.bool_)):
            value = int(value)
        if isinstance(value, (int, np.integer)):
            value = np.int64(value)
        elif isinstance(value, (float, np.floating)) and not isinstance(value, np.float32):
            value = np.float64(value)
        else:
            value = np.asarray(value)
            if value.dtype not in (np.float32, np.float64, np.int64):
                raise TypeError(f"{path!r}: dtype {value.dtype} is outside the written subset (float32/float64/int64)")
        node[parts[-1]] = value

    # ------------------------------------------------------------------ layout
    def close(self):
        if self.closed:
            return
        self.closed = True
        chunks: List[bytes] = []
        pos = 48                                              # first object behind the superblock (relative address)

        def emit(b: bytes) -> int:
            nonlocal pos
            pad = (-pos) % 8
            if pad:
                chunks.append(b"\0" * pad)
                pos += pad
            addr = pos
            chunks.append(b)
            pos += len(b)
            return addr

        def dataset(value) -> int:
            dt = {np.dtype(np.float64): _DT_F64, np.dtype(np.float32): _DT_F32, np.dtype(np.int64): _DT_I64}[np.asarray(value).dtype]
            msgs = [_message(0x05, bytes([3, 0x09]))]
            arr = np.asarray(value)
            if arr.ndim == 0:
                msgs.append(_message(MSG_DATASPACE, bytes([2, 0, 0, 0])))
                msgs.append(_message(MSG_DATATYPE, dt))
                raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
                msgs.append(_message(MSG_LAYOUT, bytes([4, 0]) + struct.pack("<H", len(raw)) + raw))
            else:
                raw = np.asfortranarray(arr).astype(arr.dtype.newbyteorder("<"), copy=False).tobytes(order="F")
                dims = tuple(reversed(arr.shape))
                msgs.append(_message(MSG_DATASPACE, bytes([2, len(dims), 0, 1]) + struct.pack(f"<{len(dims)}Q", *dims)))
                msgs.append(_message(MSG_DATATYPE, dt))
                daddr = emit(raw)
                msgs.append(_message(MSG_LAYOUT, bytes([4, 1]) + struct.pack("<QQ", daddr, len(raw))))
            return emit(_object_header(msgs))

        def group(node: Dict[str, object]) -> int:
            links = []
            for name, child in node.items():
                addr = group(child) if isinstance(child, dict) else dataset(child)
                nb = name.encode()
                if len(nb) < 256:
                    links.append(_message(MSG_LINK, bytes([1, 0x10, 1, len(nb)]) + nb + struct.pack("<Q", addr)))
                else:
                    links.append(_message(MSG_LINK, bytes([1, 0x11, 1]) + struct.pack("<H", len(nb)) + nb + struct.pack("<Q", addr)))
            msgs = [_message(0x02, bytes([0, 0]) + b"\xff" * 16), _message(0x0A, bytes([0, 0]))] + links
            return emit(_object_header(msgs))

        root = group(self.tree)
        body = b"".join(chunks)
        eof = _BASE + 48 + len(body)
        sb = _SIGNATURE + bytes([2, 8, 8, 0]) + struct.pack("<QQQQ", _BASE, 0xFFFFFFFFFFFFFFFF, eof, root)
        sb += struct.pack("<I", lookup3(sb))
        with open(self.path, "wb") as fh:
            fh.write(_HEADER_TEXT.ljust(_BASE, b"\0"))
            fh.write(sb)
            fh.write(body)
