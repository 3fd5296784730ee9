"""Reader / writer of TensorFlow "V2" checkpoints (the tensor bundle ``tf.train.Saver`` writes: ``<prefix>.index`` +
``<prefix>.data-00000-of-0000N``) in pure Python + NumPy -- the reference's checkpoints (trainval_model.py:46-56, 136-142 save them,
:185-190 restore them; README.md:22 publishes them) can be opened without TensorFlow.

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*, both third party and not vendored by the reference;
restated from their published on-disk layout):
* ``.index`` is a leveldb-style sorted string table: blocks of prefix-compressed entries ``varint32 shared | varint32 non_shared |
  varint32 value_len | key suffix | value`` followed by a uint32 restart array and its length; every block is followed by a 5-byte
  trailer (compression type, masked CRC32C); the file ends in a 48-byte footer holding the block handles (varint64 offset / size)
  of the meta-index and index blocks and the magic number 0xdb4775248b80fb57.  The index block maps separator keys to data-block
  handles.  Key "" holds a ``BundleHeaderProto`` (num_shards, endianness, version); every other key is a tensor name and holds a
  ``BundleEntryProto`` (dtype, shape, shard_id, offset, size, crc32c).
* ``.data-XXXXX-of-YYYYY`` holds the raw little-endian tensor bytes at those offsets.
Only uncompressed blocks (what the bundle writer emits), full (unsliced) tensors and numeric dtypes are supported; anything else
raises.  CRCs are verified on the index blocks always and on tensor data when ``verify=True``.
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 17: np.uint16,
           19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_OF = {np.dtype(v): k for k, v in _DTYPES.items()}

# ---- CRC32C (Castagnoli), table driven; slicing over NumPy for large buffers -----------------------------------------------------
_POLY = 0x82F63B78
_T = np.zeros((8, 256), dtype=np.uint32)
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _T[0, _i] = _c
for _k in range(1, 8):
    _T[_k] = (_T[_k - 1] >> 8) ^ _T[0][_T[_k - 1] & 0xFF]
_T0 = [int(x) for x in _T[0]]


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC32C of `data`.  Buffers above a few kB go through a NumPy formulation: the CRC is linear over GF(2), so the contributions
    of the bytes of a fixed-size chunk are table look-ups that can be XOR-reduced in bulk; chunks are then chained bytewise-free with
    the 'advance by n zero bytes' operator applied as 4 table look-ups per chunk."""
    crc ^= 0xFFFFFFFF
    mv = memoryview(data)
    n = len(mv)
    if n >= 4096:
        crc = _crc_bulk(np.frombuffer(mv, dtype=np.uint8), crc)
    else:
        for b in mv:
            crc = (crc >> 8) ^ _T0[(crc ^ b) & 0xFF]
    return crc ^ 0xFFFFFFFF


_ZERO_ADV: Dict[int, np.ndarray] = {}


def _advance_table(nbytes: int) -> np.ndarray:
    """tables A[j][v] = state reached from register value (v << 8j) after feeding `nbytes` zero bytes (linear operator)"""
    if nbytes not in _ZERO_ADV:
        basis = np.array([1 << i for i in range(32)], dtype=np.uint32)
        cur = basis.copy()
        for _ in range(nbytes):                       # advance the 32 basis vectors bytewise
            cur = (cur >> 8) ^ _T[0][cur & 0xFF]
        tab = np.zeros((4, 256), dtype=np.uint32)
        for j in range(4):
            for v in range(256):
                acc = 0
                for bit in range(8):
                    if v >> bit & 1:
                        acc ^= int(cur[8 * j + bit])
                tab[j, v] = acc
        _ZERO_ADV[nbytes] = tab
    return _ZERO_ADV[nbytes]


def _crc_bulk(a: np.ndarray, crc: int) -> int:
    CH = 1024
    n_full = len(a) // CH
    if n_full:
        blk = a[:n_full * CH].reshape(n_full, CH)
        # CRC (zero initial register) of every chunk at once: process the chunk 8 bytes at a time, slicing-by-8, vectorised over chunks
        reg = np.zeros(n_full, dtype=np.uint32)
        for off in range(0, CH, 8):
            w = blk[:, off:off + 8].astype(np.uint32)
            lo = reg ^ (w[:, 0] | (w[:, 1] << 8) | (w[:, 2] << 16) | (w[:, 3] << 24))
            reg = (_T[7][lo & 0xFF] ^ _T[6][(lo >> 8) & 0xFF] ^ _T[5][(lo >> 16) & 0xFF] ^ _T[4][lo >> 24] ^
                   _T[3][w[:, 4]] ^ _T[2][w[:, 5]] ^ _T[1][w[:, 6]] ^ _T[0][w[:, 7]])
        adv = _advance_table(CH)
        for r in reg.tolist():                        # chain: crc' = advance(crc, CH zero bytes) ^ chunk_crc
            crc = int(adv[0][crc & 0xFF] ^ adv[1][(crc >> 8) & 0xFF] ^ adv[2][(crc >> 16) & 0xFF] ^ adv[3][crc >> 24]) ^ r
    for b in a[n_full * CH:].tolist():
        crc = (crc >> 8) ^ _T0[(crc ^ b) & 0xFF]
    return crc


def _mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _unmask(m: int) -> int:
    r = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ---- varints / protobuf wire format ------------------------------------------------------------------------------------------------
def _get_varint(buf, pos: int) -> Tuple[int, int]:
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if b < 0x80:
            return val, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf) -> Dict[int, list]:
    """field number -> list of raw values (int for varint / fixed, bytes for length-delimited)"""
    out: Dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.setdefault(f, []).append(v)
    return out


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_entry(val: bytes):
    p = _parse_proto(val)
    dtype = p.get(1, [0])[0]
    shape = []
    if 2 in p:
        sp = _parse_proto(p[2][0])
        if sp.get(3, [0])[0]:
            raise ValueError("tensor of unknown rank in the bundle")
        for d in sp.get(2, []):
            shape.append(_signed(_parse_proto(d).get(1, [0])[0]))
    if 7 in p:
        raise ValueError("sliced (partitioned) variables are not supported")
    return dict(dtype=dtype, shape=tuple(shape), shard=p.get(3, [0])[0], offset=p.get(4, [0])[0], size=p.get(5, [0])[0],
                crc=p.get(6, [None])[0])


# ---- table -----------------------------------------------------------------------------------------------------------------------------
def _read_block(buf, offset: int, size: int, what: str):
    data = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if _unmask(stored) != crc32c(bytes(buf[offset:offset + size + 1])):
        raise ValueError(f"{what}: block checksum mismatch (corrupt .index file)")
    if ctype != 0:
        raise ValueError(f"{what}: compressed blocks (type {ctype}) are not supported")
    return data


def _block_entries(block) -> Iterable[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared]); pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(prefix: str) -> Tuple[Dict[str, dict], dict]:
    """-> ({tensor name: entry}, header) of the bundle `<prefix>.index`"""
    buf = memoryview(Path(str(prefix) + ".index").read_bytes())
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError(f"{prefix}.index is not a TensorFlow V2 checkpoint index (bad magic)")
    foot = buf[len(buf) - 48:]
    pos = 0
    _, pos = _get_varint(foot, pos); _, pos = _get_varint(foot, pos)            # meta-index handle
    ioff, pos = _get_varint(foot, pos); isz, pos = _get_varint(foot, pos)
    entries, header = {}, None
    for _, handle in _block_entries(_read_block(buf, ioff, isz, "index block")):
        boff, p2 = _get_varint(handle, 0)
        bsz, _ = _get_varint(handle, p2)
        for key, val in _block_entries(_read_block(buf, boff, bsz, "data block")):
            if key == b"":
                h = _parse_proto(val)
                header = dict(num_shards=h.get(1, [1])[0], endianness=h.get(2, [0])[0])
            else:
                entries[key.decode("utf-8")] = _parse_entry(val)
    if header is None:
        raise ValueError("bundle header missing")
    if header["endianness"] != 0:
        raise ValueError("big-endian bundles are not supported")
    return entries, header


def read_bundle(prefix: str, names: Optional[Iterable[str]] = None, *, verify: bool = False) -> Dict[str, np.ndarray]:
    """Loads the tensors of the checkpoint `<prefix>` (all, or `names`) as NumPy arrays."""
    entries, header = read_index(prefix)
    want = list(entries) if names is None else list(names)
    out, shards = {}, {}
    for name in want:
        if name not in entries:
            raise KeyError(f"{name!r} not in checkpoint {prefix}")
        e = entries[name]
        if e["dtype"] not in _DTYPES:
            raise ValueError(f"{name}: unsupported dtype enum {e['dtype']} (strings / resources / quantised types)")
        dt = np.dtype(_DTYPES[e["dtype"]])
        f = shards.get(e["shard"])
        if f is None:
            f = shards[e["shard"]] = np.memmap(f"{prefix}.data-{e['shard']:05d}-of-{header['num_shards']:05d}", dtype=np.uint8, mode="r")
        raw = f[e["offset"]:e["offset"] + e["size"]]
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if n * dt.itemsize != e["size"]:
            raise ValueError(f"{name}: {e['size']} bytes stored for shape {e['shape']} of {dt}")
        if verify and e["crc"] is not None and _unmask(e["crc"]) != crc32c(raw.tobytes()):
            raise ValueError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e["shape"])
    return out


# ---- writer (what tf.train.Saver / tf.train.load_checkpoint can read back) --------------------------------------------------------
def _pb_varint(field: int, v: int) -> bytes:
    return _put_varint(field << 3) + _put_varint(v)


def _pb_bytes(field: int, b: bytes) -> bytes:
    return _put_varint(field << 3 | 2) + _put_varint(len(b)) + b


def _block(items) -> bytes:
    """one table block, every entry a restart point (no prefix compression: simplest valid encoding)"""
    out, restarts = bytearray(), []
    for k, v in items:
        restarts.append(len(out))
        out += _put_varint(0) + _put_varint(len(k)) + _put_varint(len(v)) + k + v
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Writes `<prefix>.index` and `<prefix>.data-00000-of-00001` (one shard, one data block)."""
    data = bytearray()
    items = [(b"", _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1)))]      # header: num_shards = 1, little endian, version.producer = 1
    for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
        a = np.asarray(tensors[name], order="C")              # (ascontiguousarray would promote a scalar to shape (1,))
        if a.dtype not in _DT_OF:
            raise ValueError(f"{name}: unsupported dtype {a.dtype}")
        raw = a.tobytes()
        shape = b"".join(_pb_bytes(2, _pb_varint(1, int(d))) for d in a.shape)
        entry = _pb_varint(1, _DT_OF[a.dtype]) + _pb_bytes(2, shape)
        if len(data):
            entry += _pb_varint(4, len(data))
        entry += _pb_varint(5, len(raw)) + _put_varint(6 << 3 | 5) + struct.pack("<I", _mask(crc32c(raw)))
        items.append((name.encode("utf-8"), entry))
        data += raw
    Path(f"{prefix}.data-00000-of-00001").write_bytes(bytes(data))

    def with_trailer(blk: bytes) -> bytes:
        return blk + b"\x00" + struct.pack("<I", _mask(crc32c(blk + b"\x00")))
    out = bytearray()
    dblk = _block(items)
    dh = _put_varint(0) + _put_varint(len(dblk))
    out += with_trailer(dblk)
    mblk = _block([])
    mh = _put_varint(len(out)) + _put_varint(len(mblk))
    out += with_trailer(mblk)
    iblk = _block([(items[-1][0] + b"\x00", dh)])                            # separator key >= the last key of the data block
    ih = _put_varint(len(out)) + _put_varint(len(iblk))
    out += with_trailer(iblk)
    foot = mh + ih
    out += foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", MAGIC)
    Path(str(prefix) + ".index").write_bytes(bytes(out))
