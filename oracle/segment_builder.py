"""Builds byte-exact Pinot index buffers from column values -- TEST INFRASTRUCTURE (part of the oracle).

Restates the three creators the hot path's formats come from:

* dictionary: sorted distinct values, fixed width, BIG-endian
  (``pinot-segment-local/.../segment/creator/impl/SegmentDictionaryCreator.java:110-180``; strings padded with ``\\0`` to
  the longest entry);
* forward index: MSB-first big-endian bit stream of dictIds, ``ceil(N*b/8)`` bytes
  (``.../io/writer/impl/FixedBitSVForwardIndexWriter.java:39-46``, ``.../io/util/PinotDataBitSet.java:143-170``); a
  SORTED column instead stores (startDocId, endDocId) BE int pairs per dictId (``SortedIndexReaderImpl``);
* inverted index: (card+1) BE u32 offsets + RoaringBitmap portable serialization per dictId
  (``.../segment/creator/impl/inv/BitmapInvertedIndexWriter.java:33-50,90-97``);
* raw (no dictionary) fixed-byte forward index, version 4-less "v2/v3" chunk header with PASS_THROUGH compression
  (``.../segment/index/readers/forward/BaseChunkForwardIndexReader.java:60-106``).

The numpy packer is cross-checked against the oracle's byte-at-a-time ``po_bitset_write`` in tests/test_oracle_formats.py.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

INT, LONG, FLOAT, DOUBLE, STRING = 0, 1, 2, 3, 4
_NP_BE = {INT: ">i4", LONG: ">i8", FLOAT: ">f4", DOUBLE: ">f8"}


def num_bits_per_value(max_value: int) -> int:
    """PinotDataBitSet.getNumBitsPerValue (pinot-segment-local/.../io/util/PinotDataBitSet.java:61-72)."""
    return 1 if max_value <= 1 else int(max_value).bit_length()


def pack_fixed_bits(values: np.ndarray, bits: int) -> np.ndarray:
    """MSB-first bit stream, ceil(n*bits/8) bytes."""
    v = np.asarray(values, dtype=np.uint32)
    n = v.shape[0]
    out = np.zeros((n * bits + 7) // 8, dtype=np.uint8)
    chunk = 1 << 20
    assert (chunk * bits) % 8 == 0
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        shifts = np.arange(bits - 1, -1, -1, dtype=np.uint32)
        b = ((v[s:e, None] >> shifts) & 1).astype(np.uint8).reshape(-1)
        packed = np.packbits(b)
        o = s * bits // 8
        out[o:o + packed.shape[0]] = packed
    return out


def unpack_fixed_bits(buf: np.ndarray, n: int, bits: int) -> np.ndarray:
    b = np.unpackbits(np.asarray(buf, dtype=np.uint8))[: n * bits].reshape(n, bits).astype(np.uint32)
    w = (1 << np.arange(bits - 1, -1, -1, dtype=np.uint32))
    return (b * w).sum(axis=1).astype(np.int32)


@dataclass
class ColumnData:
    name: str
    data_type: int
    has_dictionary: bool
    bits: int
    cardinality: int
    is_sorted: bool
    dict_entry_bytes: int
    fwd: np.ndarray  # uint8
    dict: Optional[np.ndarray]  # uint8
    inv: Optional[np.ndarray]  # uint8
    dict_values: Optional[np.ndarray] = None  # decoded dictionary (numeric array or array of bytes)
    dict_ids: Optional[np.ndarray] = None  # int32 per doc (kept for tests)
    raw_values: Optional[np.ndarray] = None
    # raw column written with compressed chunks: `fwd` is that file (what a loader gets), `oracle_fwd` the same values as
    # PASS_THROUGH chunks -- the C oracle reads those; the chunk codecs are the oracle's Python half (oracle/chunk_codecs.py)
    oracle_fwd: Optional[np.ndarray] = None


@dataclass
class SegmentData:
    name: str
    num_docs: int
    columns: List[ColumnData] = field(default_factory=list)

    def column_index(self, name: str) -> int:
        for i, c in enumerate(self.columns):
            if c.name == name:
                return i
        raise KeyError(name)

    def column(self, name: str) -> ColumnData:
        return self.columns[self.column_index(name)]

    def value_of(self, column: str, dict_id: int):
        c = self.column(column)
        v = c.dict_values[dict_id]
        return v.decode("utf-8") if c.data_type == STRING else v.item()


def _np_type_of(values: np.ndarray) -> int:
    if values.dtype.kind == "S":
        return STRING
    if values.dtype == np.int32:
        return INT
    if values.dtype == np.int64:
        return LONG
    if values.dtype == np.float32:
        return FLOAT
    if values.dtype == np.float64:
        return DOUBLE
    raise TypeError(values.dtype)


def build_raw_column(name: str, values: np.ndarray, version: int = 2, compression: int = 0, docs_per_chunk: int = 1000) -> ColumnData:
    """No-dictionary fixed-byte SV column (FixedByteChunkForwardIndexWriter layout); `compression` is a ChunkCompressionType
    ordinal (0 PASS_THROUGH, 1 SNAPPY, 3 LZ4, 4 LZ4_LENGTH_PREFIXED)."""
    from . import chunk_codecs as cc
    values = np.asarray(values)
    dt = _np_type_of(values)
    width = np.dtype(_NP_BE[dt]).itemsize
    n = values.shape[0]
    body = values.astype(_NP_BE[dt]).tobytes()
    plain = cc.encode_fixed_byte_forward(body, width, n, cc.PASS_THROUGH, 2 if version == 1 else version, docs_per_chunk)
    if compression == cc.PASS_THROUGH and version != 1:
        return ColumnData(name, dt, False, 0, 0, False, width, plain, None, None, raw_values=values)
    fwd = cc.encode_fixed_byte_forward(body, width, n, compression, version, docs_per_chunk)
    assert cc.decode_fixed_byte_forward(fwd, width, n) == body
    return ColumnData(name, dt, False, 0, 0, False, width, fwd, None, None, raw_values=values, oracle_fwd=plain)


def build_column(name: str, values: np.ndarray, inverted: bool = False, lib=None) -> ColumnData:
    values = np.asarray(values)
    dt = _np_type_of(values)
    dict_values, dict_ids = np.unique(values, return_inverse=True)
    dict_ids = dict_ids.astype(np.int32)
    card = int(dict_values.shape[0])
    bits = num_bits_per_value(card - 1)
    n = values.shape[0]
    if dt == STRING:
        width = max(1, dict_values.dtype.itemsize)
        dbytes = np.frombuffer(dict_values.astype(f"S{width}").tobytes(), dtype=np.uint8).copy()
    else:
        width = np.dtype(_NP_BE[dt]).itemsize
        dbytes = np.frombuffer(dict_values.astype(_NP_BE[dt]).tobytes(), dtype=np.uint8).copy()
    is_sorted = bool(n > 0 and np.all(dict_ids[1:] >= dict_ids[:-1]))
    if is_sorted:
        # SortedIndexReaderImpl: per dictId (start, end) inclusive doc ids
        starts = np.searchsorted(dict_ids, np.arange(card), side="left")
        ends = np.searchsorted(dict_ids, np.arange(card), side="right") - 1
        fwd = np.frombuffer(np.stack([starts, ends], axis=1).astype(">i4").tobytes(), dtype=np.uint8).copy()
    else:
        fwd = pack_fixed_bits(dict_ids, bits)
    inv = None
    if inverted and not is_sorted:
        assert lib is not None, "building an inverted index needs the oracle library"
        inv = lib.inverted_index_build(dict_ids, card)
    return ColumnData(name, dt, True, bits, card, is_sorted, width, fwd, dbytes, inv, dict_values, dict_ids)


def build_segment(name: str, columns: Dict[str, np.ndarray], inverted: Sequence[str] = (), raw: Sequence[str] = (),
                  lib=None, raw_compression: Optional[Dict[str, int]] = None) -> SegmentData:
    n = None
    cols = []
    for cname, vals in columns.items():
        vals = np.asarray(vals)
        n = vals.shape[0] if n is None else n
        assert vals.shape[0] == n
        if cname in raw:
            cols.append(build_raw_column(cname, vals, compression=(raw_compression or {}).get(cname, 0)))
        else:
            cols.append(build_column(cname, vals, inverted=cname in inverted, lib=lib))
    return SegmentData(name, int(n), cols)
