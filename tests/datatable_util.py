"""Test-side READER of DataTableImplV4 bytes (pinot-common/.../datatable/DataTableImplV4.java:108-215 deserialisation,
DataTableUtils.computeColumnOffsets :41-65, ObjectSerDeUtils object types) -- restated independently of the writer in
pinot_b200/csrc/host/datatable.cpp so that a layout mistake on one side shows.  Byte-level parity with the reference's own
writer is UNPINNED: the reference holds no golden DataTable bytes and cannot run here (no JVM)."""
import struct


class _Buf:
    def __init__(self, b, p=0):
        self.b, self.p = b, p

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def i64(self):
        v = struct.unpack_from(">q", self.b, self.p)[0]
        self.p += 8
        return v

    def f32(self):
        v = struct.unpack_from(">f", self.b, self.p)[0]
        self.p += 4
        return v

    def f64(self):
        v = struct.unpack_from(">d", self.b, self.p)[0]
        self.p += 8
        return v

    def string(self):
        n = self.i32()
        s = self.b[self.p:self.p + n].decode("utf-8")
        self.p += n
        return s


METADATA_KEYS = {2: ("numDocsScanned", "long"), 3: ("numEntriesScannedInFilter", "long"), 4: ("numEntriesScannedPostFilter", "long"),
                 10: ("totalDocs", "long"), 11: ("numGroupsLimitReached", "string")}


def parse(data: bytes):
    """-> {version, column_names, column_types, rows (python values; AVG -> (sum, count); sets -> frozenset), metadata}"""
    h = _Buf(data)
    version, num_rows, num_cols = h.i32(), h.i32(), h.i32()
    sect = [(h.i32(), h.i32()) for _ in range(5)]   # exceptions, dictionary, schema, fixed, variable
    assert version == 4 and h.p == 52
    (ex_s, ex_l), (di_s, di_l), (sc_s, sc_l), (fx_s, fx_l), (va_s, va_l) = sect
    assert ex_s == 52 and di_s == ex_s + ex_l and sc_s == di_s + di_l and fx_s == sc_s + sc_l and va_s == fx_s + fx_l
    assert _Buf(data, ex_s).i32() == 0  # no exceptions
    sdict = []
    if di_l:
        d = _Buf(data, di_s)
        sdict = [d.string() for _ in range(d.i32())]
        assert d.p == di_s + di_l
    s = _Buf(data, sc_s)
    n = s.i32()
    assert n == num_cols
    names = [s.string() for _ in range(n)]
    types = [s.string() for _ in range(n)]
    assert s.p == sc_s + sc_l
    offsets, size = [], 0
    for t in types:
        offsets.append(size)
        size += 4 if t in ("INT", "FLOAT", "STRING") else 8
    assert fx_l == num_rows * size
    rows = []
    for r in range(num_rows):
        row = []
        for c, t in enumerate(types):
            f = _Buf(data, fx_s + r * size + offsets[c])
            if t == "INT":
                row.append(f.i32())
            elif t == "LONG":
                row.append(f.i64())
            elif t == "FLOAT":
                row.append(f.f32())
            elif t == "DOUBLE":
                row.append(f.f64())
            elif t == "STRING":
                row.append(sdict[f.i32()])
            else:  # OBJECT
                pos, length = f.i32(), f.i32()
                v = _Buf(data, va_s + pos)
                code = v.i32()
                end = v.p + length
                if code == 4:
                    row.append((v.f64(), v.i64()))
                elif code in (9, 15, 16, 17, 18):
                    rd = {9: v.i32, 15: v.i64, 16: v.f32, 17: v.f64, 18: v.string}[code]
                    row.append(frozenset(rd() for _ in range(v.i32())))
                else:
                    raise ValueError(f"object type {code}")
                assert v.p == end
        rows.append(row)
    m = _Buf(data, va_s + va_l)
    mlen = m.i32()
    assert m.p + mlen == len(data)
    meta = {}
    for _ in range(m.i32()):
        name, kind = METADATA_KEYS[m.i32()]
        meta[name] = m.i64() if kind == "long" else m.string()
    assert m.p == len(data)
    return {"version": version, "column_names": names, "column_types": types, "rows": rows, "metadata": meta}
