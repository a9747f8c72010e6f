"""Read-only importer of the reference's ``docs/`` store, so that a database folder WRITTEN BY THE REFERENCE opens here.

The reference keeps ``{"doc": text, "int8": ndarray}`` (CohereEnhancedVectorDB.py:220-221) or ``{"doc", "emb_int8",
"min_max"}`` (VectorDBInt8.py:179-183) per document in ``rocksdict.Rdict(folder/docs)``, i.e. RocksDB block-based table
files (``*.sst``) whose values are pickles.  rocksdict is not installable here and a storage engine is out of scope, but
the documents and quantised vectors of an existing database are DATA this path needs - so this module walks the public
table format (footer -> metaindex -> index block -> data blocks, raw or Snappy-compressed) of every ``*.sst`` in the
folder, applies RocksDB's "highest sequence number wins / deletions hide" rule across files, and returns a plain dict.
``CohereEnhancedVectorDB`` / ``VectorDB*`` call it on open when the folder holds a RocksDB store and no payload sidecar;
the vectors then go to the device-resident index, the texts to the ``DocStore``.  Nothing is ever written back in
RocksDB's format (``save()`` writes index.bin + the sidecar + docs.log).

rocksdict encodes keys and values with a one-byte type tag: 1 bytes, 2 str, 3 int (big-endian two's complement),
4 float, 5 bool, 6 pickle.  Pickles are loaded through a whitelist (numpy arrays and scalars only): database folders
are untrusted input.
"""
from __future__ import annotations

import io
import os
import pickle
import re
import struct
from typing import Any, Dict, Iterator, Tuple

_MAGIC = bytes.fromhex("f7cff485b741e288")  # kBlockBasedTableMagicNumber, little-endian
_FOOTER = 53

_ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"),
            ("numpy", "dtype"), ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar")}


class RocksImportError(Exception):
    pass


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) not in _ALLOWED:
            raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name} from a database folder")
        return super().find_class(module, name)


def _varint(b: bytes, o: int) -> Tuple[int, int]:
    r = s = 0
    while True:
        c = b[o]
        o += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, o


def _snappy(b: bytes) -> bytes:
    """Snappy raw format: varint length, then literal / copy elements."""
    n, o = _varint(b, 0)
    out = bytearray()
    while o < len(b):
        tag = b[o]
        o += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(b[o:o + nb], "little")
                o += nb
            ln += 1
            out += b[o:o + ln]
            o += ln
            continue
        if kind == 1:
            ln, off = ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | b[o]
            o += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(b[o:o + 2], "little")
            o += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(b[o:o + 4], "little")
            o += 4
        if off == 0 or off > len(out):
            raise RocksImportError("corrupt Snappy block")
        if off >= ln:
            start = len(out) - off
            out += out[start:start + ln]
        else:  # overlapping copy = run-length expansion
            for _ in range(ln):
                out.append(out[-off])
    if len(out) != n:
        raise RocksImportError("corrupt Snappy block (length)")
    return bytes(out)


def _block(b: bytes, off: int, size: int) -> bytes:
    """Block contents followed by a 5-byte trailer: compression type + checksum."""
    if off + size + 1 > len(b):
        raise RocksImportError("block handle beyond the end of the file")
    ctype = b[off + size]
    raw = b[off:off + size]
    if ctype == 0:
        return raw
    if ctype == 1:
        return _snappy(raw)
    raise RocksImportError(f"unsupported block compression type {ctype} (only none / Snappy are read)")


def _entries(raw: bytes, index: bool = False) -> Iterator[Tuple[bytes, Any]]:
    """Prefix-compressed entries of a data block (key, value) or of an index block (key, block handle)."""
    nrestarts = struct.unpack_from("<I", raw, len(raw) - 4)[0] & 0x7FFFFFFF
    end, o, key = len(raw) - 4 - 4 * nrestarts, 0, b""
    while o < end:
        shared, o = _varint(raw, o)
        nonshared, o = _varint(raw, o)
        if index:
            key = key[:shared] + raw[o:o + nonshared]
            o += nonshared
            h_off, o = _varint(raw, o)
            h_size, o = _varint(raw, o)
            yield key, (h_off, h_size)
        else:
            vlen, o = _varint(raw, o)
            key = key[:shared] + raw[o:o + nonshared]
            o += nonshared
            yield key, raw[o:o + vlen]
            o += vlen


def _sst_entries(path: str) -> Iterator[Tuple[bytes, int, int, bytes]]:
    """(user key, sequence number, value type, value) of every entry of one table file."""
    b = open(path, "rb").read()
    n = len(b)
    if n < _FOOTER or b[-8:] != _MAGIC:
        raise RocksImportError(f"{path}: not a RocksDB block-based table")
    meta_size = struct.unpack_from("<I", b, n - _FOOTER + 13)[0]
    meta = dict(_entries(_block(b, n - _FOOTER - 5 - meta_size, meta_size)))
    if b"rocksdb.index" not in meta:
        raise RocksImportError(f"{path}: table format not understood (no rocksdb.index in the metaindex block)")
    v = meta[b"rocksdb.index"]
    i_off, o = _varint(v, 0)
    i_size, _ = _varint(v, o)
    for _, (d_off, d_size) in _entries(_block(b, i_off, i_size), index=True):
        for ikey, val in _entries(_block(b, d_off, d_size)):
            if len(ikey) < 8:
                raise RocksImportError(f"{path}: short internal key")
            trailer = int.from_bytes(ikey[-8:], "little")
            yield ikey[:-8], trailer >> 8, trailer & 0xFF, val


def _decode(tagged: bytes, what: str) -> Any:
    if not tagged:
        raise RocksImportError(f"empty {what}")
    tag, body = tagged[0], tagged[1:]
    if tag == 1:
        return bytes(body)
    if tag == 2:
        return body.decode("utf-8")
    if tag == 3:
        return int.from_bytes(body, "big", signed=True)
    if tag == 4:
        return struct.unpack(">d", body)[0] if len(body) == 8 else struct.unpack("<d", body[:8])[0]
    if tag == 5:
        return body != b"\x00"
    if tag == 6:
        return _Unpickler(io.BytesIO(body)).load()
    raise RocksImportError(f"unknown rocksdict type tag {tag} in a {what}")


def is_rocksdict_folder(path: str) -> bool:
    return os.path.isfile(os.path.join(path, "CURRENT")) and any(f.endswith(".sst") for f in os.listdir(path))


def read_rocksdict_folder(path: str) -> Dict[Any, Any]:
    """{key: value} of a rocksdict store, read straight from its table files.  Raises RocksImportError when the folder
    holds unflushed writes (a non-empty write-ahead log) - those live only in RocksDB's own log format."""
    files = sorted(os.listdir(path))
    for f in files:
        if re.fullmatch(r"\d+\.log", f) and os.path.getsize(os.path.join(path, f)) > 0:
            raise RocksImportError(f"{path}/{f}: the store has unflushed writes in its write-ahead log; open and close it once with "
                                   "rocksdict (which flushes the log into a table file) before importing")
    best: Dict[bytes, Tuple[int, int, bytes]] = {}
    for f in files:
        if not f.endswith(".sst"):
            continue
        for ukey, seq, vtype, val in _sst_entries(os.path.join(path, f)):
            cur = best.get(ukey)
            if cur is None or seq >= cur[0]:
                best[ukey] = (seq, vtype, val)
    out: Dict[Any, Any] = {}
    for ukey, (seq, vtype, val) in best.items():
        if vtype == 0:  # kTypeDeletion
            continue
        if vtype != 1:  # kTypeValue
            raise RocksImportError(f"unsupported RocksDB value type {vtype} (merge operands / blob indexes are not read)")
        out[_decode(ukey, "key")] = _decode(val, "value")
    return out
