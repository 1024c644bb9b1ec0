"""TEST INFRASTRUCTURE (oracle): chunk codecs of the reference's raw (no-dictionary) forward indexes, restated in Python.

Reference: ``BaseChunkForwardIndexReader.java:60-106`` (header), ``:204-240`` (one chunk = bytes up to the next chunk offset,
decompressed as a whole), ``FixedByteChunkSVForwardIndexReader.java:52-100`` (value i of a chunk at i * width),
``ChunkCompressionType.java`` (PASS_THROUGH 0, SNAPPY 1, ZSTANDARD 2, LZ4 3, LZ4_LENGTH_PREFIXED 4, GZIP 5).  The
compression libraries themselves are third-party dependencies of the reference (org.xerial.snappy:snappy-java,
org.lz4:lz4-java, versions pinned in the reference's pom.xml) and are not in /root/reference: the decoders below restate
the published formats (Snappy raw format; LZ4 block format).

Pinned: Snappy + the header handling against the reference's own fixtures ``fixedByteCompressed.v2`` (version 2, SNAPPY),
``fixedByteSVRDoubles.v1`` (version 1, always SNAPPY) and ``fixedByteRaw.v2`` (PASS_THROUGH), copied to ``tests/golden/``
with the expected values of ``FixedByteChunkSVForwardIndexTest.java:340-377`` (value i == i + startValue).
LZ4: PARITY UNPINNED -- the reference holds no LZ4 fixture for this reader; the decoder is only checked against the
compressor below (format-level round trip).

The compressors here exist to produce test inputs; they emit valid streams, not the byte-identical output of the Java
libraries.
"""
from __future__ import annotations

import struct

import numpy as np

PASS_THROUGH, SNAPPY, ZSTANDARD, LZ4, LZ4_LENGTH_PREFIXED, GZIP = range(6)


def snappy_decode(src: bytes) -> bytes:
    ip, n = 0, len(src)
    ulen, shift = 0, 0
    while True:
        b = src[ip]; ip += 1
        ulen |= (b & 0x7F) << shift
        if not b & 0x80:
            break
        shift += 7
    out = bytearray()
    while ip < n:
        tag = src[ip]; ip += 1
        kind = tag & 3
        if kind == 0:
            ln = (tag >> 2) + 1
            if ln > 60:
                nb = ln - 60
                ln = int.from_bytes(src[ip:ip + nb], "little") + 1
                ip += nb
            out += src[ip:ip + ln]
            ip += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | src[ip]; ip += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[ip:ip + 2], "little"); ip += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[ip:ip + 4], "little"); ip += 4
        if off == 0 or off > len(out):
            raise ValueError("snappy: bad copy offset")
        for _ in range(ln):   # copies may overlap their own output
            out.append(out[-off])
    if len(out) != ulen:
        raise ValueError(f"snappy: {len(out)} bytes produced, header says {ulen}")
    return bytes(out)


def lz4_block_decode(src: bytes) -> bytes:
    ip, n = 0, len(src)
    out = bytearray()
    while ip < n:
        token = src[ip]; ip += 1
        lit = token >> 4
        if lit == 15:
            while True:
                b = src[ip]; ip += 1
                lit += b
                if b != 255:
                    break
        out += src[ip:ip + lit]
        ip += lit
        if ip >= n:
            break
        off = src[ip] | (src[ip + 1] << 8); ip += 2
        ml = token & 15
        if ml == 15:
            while True:
                b = src[ip]; ip += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        if off == 0 or off > len(out):
            raise ValueError("lz4: bad match offset")
        for _ in range(ml):
            out.append(out[-off])
    return bytes(out)


def _find_matches(data: bytes, min_match: int, max_off: int, last_start=None, last_end=None):
    """Greedy single-candidate matcher shared by the two test compressors: yields (literal_start, literal_end, offset, length)
    for every match, then (literal_start, len(data), 0, 0) for the trailing literals."""
    table = {}
    i, anchor, n = 0, 0, len(data)
    last_start = n - min_match if last_start is None else last_start
    last_end = n if last_end is None else last_end
    while i <= last_start and i + min_match <= last_end:
        key = data[i:i + 4]
        j = table.get(key, -1)
        table[key] = i
        if j >= 0 and i - j <= max_off and data[j:j + min_match] == data[i:i + min_match]:
            ln = min_match
            while i + ln < last_end and data[j + ln] == data[i + ln]:
                ln += 1
            yield anchor, i, i - j, ln
            i += ln
            anchor = i
        else:
            i += 1
    yield anchor, n, 0, 0


def lz4_block_encode(data: bytes) -> bytes:
    out = bytearray()

    def ext(v):
        while v >= 255:
            out.append(255); v -= 255
        out.append(v)
    n = len(data)
    # block format end conditions: the last 5 bytes are literals, no match starts within the last 12 bytes
    for a, e, off, ln in _find_matches(data, 4, 65535, last_start=n - 12, last_end=n - 5):
        lit = e - a
        out.append((min(lit, 15) << 4) | (min(ln - 4, 15) if ln else 0))
        if lit >= 15:
            ext(lit - 15)
        out += data[a:e]
        if ln:
            out += struct.pack("<H", off)
            if ln - 4 >= 15:
                ext(ln - 4 - 15)
    return bytes(out)


def snappy_encode(data: bytes) -> bytes:
    out = bytearray()
    v = len(data)
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80); v >>= 7
    out.append(v)

    def literal(chunk):
        ln = len(chunk)
        if ln == 0:
            return
        if ln <= 60:
            out.append((ln - 1) << 2)
        else:
            nb = (max(ln - 1, 1).bit_length() + 7) // 8
            out.append((59 + nb) << 2)
            out.extend((ln - 1).to_bytes(nb, "little"))
        out.extend(chunk)
    for a, e, off, ln in _find_matches(data, 4, 65535):
        literal(data[a:e])
        while ln > 0:
            step = min(ln, 64)
            if ln - step in (1, 2, 3):   # never leave a copy shorter than 4 behind
                step -= 4
            out.append(((step - 1) << 2) | 2)
            out.extend(struct.pack("<H", off))
            ln -= step
    return bytes(out)


def decode_fixed_byte_forward(file_bytes, width: int, num_docs: int) -> bytes:
    """All values of a fixed-width raw SV forward index as one big-endian byte string (what getInt/getLong/getFloat/getDouble
    of FixedByteChunkSVForwardIndexReader return doc by doc)."""
    b = bytes(file_bytes)
    version, num_chunks, per_chunk, entry = struct.unpack(">4i", b[:16])
    if entry != width:
        raise ValueError(f"entry width {entry} != {width}")
    compression, header = SNAPPY, 16
    if version > 1:
        _total, compression, header = struct.unpack(">3i", b[16:28])
    osz = 4 if version <= 2 else 8
    offs = [int.from_bytes(b[header + c * osz: header + (c + 1) * osz], "big") for c in range(num_chunks)] + [len(b)]
    data_start = header + num_chunks * osz
    if compression == PASS_THROUGH:
        return b[data_start: data_start + num_docs * width]
    out = bytearray()
    for c in range(num_chunks):
        chunk = b[offs[c]: offs[c + 1]]
        if compression == SNAPPY:
            raw = snappy_decode(chunk)
        elif compression == LZ4:
            raw = lz4_block_decode(chunk)
        elif compression == LZ4_LENGTH_PREFIXED:
            raw = lz4_block_decode(chunk[4:])
            if len(raw) != int.from_bytes(chunk[:4], "little"):
                raise ValueError("lz4: length prefix mismatch")
        else:
            raise NotImplementedError(f"chunk compression {compression}")
        out += raw[: per_chunk * width]
    return bytes(out[: num_docs * width])


def encode_fixed_byte_forward(values_be: bytes, width: int, num_docs: int, compression: int, version: int = 2,
                              docs_per_chunk: int = 1000) -> np.ndarray:
    """A raw SV forward index file in FixedByteChunkForwardIndexWriter layout (test inputs)."""
    num_chunks = (num_docs + docs_per_chunk - 1) // docs_per_chunk
    osz = 4 if version <= 2 else 8
    if version == 1:
        assert compression == SNAPPY
        hdr = struct.pack(">4i", 1, num_chunks, docs_per_chunk, width)
    else:
        hdr = struct.pack(">7i", version, num_chunks, docs_per_chunk, width, num_docs, compression, 28)
    chunks = []
    for c in range(num_chunks):
        raw = values_be[c * docs_per_chunk * width: (c + 1) * docs_per_chunk * width]
        if compression == PASS_THROUGH:
            chunks.append(raw)
        elif compression == SNAPPY:
            chunks.append(snappy_encode(raw))
        elif compression == LZ4:
            chunks.append(lz4_block_encode(raw))
        elif compression == LZ4_LENGTH_PREFIXED:
            chunks.append(struct.pack("<i", len(raw)) + lz4_block_encode(raw))
        else:
            raise NotImplementedError(compression)
    pos = len(hdr) + num_chunks * osz
    offs = bytearray()
    for ch in chunks:
        offs += pos.to_bytes(osz, "big")
        pos += len(ch)
    return np.frombuffer(hdr + bytes(offs) + b"".join(chunks), dtype=np.uint8).copy()
