#!/usr/bin/env python
"""Regenerates the committed golden fixtures under tests/golden/ from the reference's own test resources.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (all small, committed):

  test_data_sv.npz       the 11 columns that BaseSingleValueQueriesTest selects out of
                         pinot-core/src/test/resources/data/test_data-sv.avro (30 000 rows)
                         (pinot-core/src/test/java/org/apache/pinot/queries/BaseSingleValueQueriesTest.java:73-90).
                         INT columns as int32 arrays, STRING columns as fixed-width byte strings.
  padding_old_*.bin      the real forward-index / dictionary files of the pre-built v1 segment
                         pinot-core/src/test/resources/data/paddingOld.tar.gz (5 docs, 3-bit dictIds): golden BYTES that pin
                         the fixed-bit layout and the big-endian dictionary layout.
  star_tree_index.bin    pinot-segment-local/src/test/resources/data/startree/segment/star_tree_index (+ its index map as
                         star_tree_index_map.txt): a real star-tree built by the reference, golden bytes for
                         OffHeapStarTree / fixed-bit dims / raw fixed-byte metric chunks.

  raw_forward/           pinot-segment-local/src/test/resources/data/{fixedByteRaw.v2, fixedByteCompressed.v2,
                         fixedByteSVRDoubles.v1}: raw (no-dictionary) DOUBLE forward indexes written by old versions of the
                         reference (PASS_THROUGH v2, SNAPPY v2, SNAPPY v1); expected values per
                         FixedByteChunkSVForwardIndexTest.java:340-377: value i == i + 100.2356 (2000 docs) / i + 0 (10009 docs).

Only DATA is copied, never reference source code.  The Avro container reader below is written from the Avro 1.x
specification (null codec, zig-zag varints, union branch index per field, 16-byte sync marker per block).
"""
import io
import json
import os
import shutil
import tarfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

SV_COLUMNS = ["column1", "column3", "column5", "column6", "column7", "column9", "column11", "column12", "column17",
              "column18", "daysSinceEpoch"]


class _Reader:
    def __init__(self, data):
        self.d = data
        self.p = 0

    def long(self):
        r = 0
        s = 0
        while True:
            b = self.d[self.p]
            self.p += 1
            r |= (b & 0x7F) << s
            s += 7
            if not b & 0x80:
                break
        return (r >> 1) ^ -(r & 1)

    def bytes_(self):
        n = self.long()
        v = self.d[self.p:self.p + n]
        self.p += n
        return v


def read_avro(path):
    """Returns (field names, list of row tuples) of a null-codec Avro object container file."""
    r = _Reader(open(path, "rb").read())
    assert r.d[:4] == b"Obj\x01"
    r.p = 4
    meta = {}
    while True:
        n = r.long()
        if n == 0:
            break
        if n < 0:
            n = -n
            r.long()
        for _ in range(n):
            k = r.bytes_().decode()
            meta[k] = r.bytes_()
    assert meta.get("avro.codec", b"null") == b"null"
    schema = json.loads(meta["avro.schema"])
    fields = schema["fields"]
    sync = r.d[r.p:r.p + 16]
    r.p += 16
    rows = []
    while r.p < len(r.d):
        count = r.long()
        r.long()  # block size in bytes
        for _ in range(count):
            row = []
            for f in fields:
                t = f["type"]
                branch = t[r.long()] if isinstance(t, list) else t
                if branch == "null":
                    row.append(None)
                elif branch in ("int", "long"):
                    row.append(r.long())
                elif branch == "string":
                    row.append(r.bytes_().decode("utf-8"))
                else:
                    raise NotImplementedError(branch)
            rows.append(tuple(row))
        assert r.d[r.p:r.p + 16] == sync
        r.p += 16
    return [f["name"] for f in fields], rows


def main():
    names, rows = read_avro(f"{REF}/pinot-core/src/test/resources/data/test_data-sv.avro")
    assert len(rows) == 30000
    out = {}
    for c in SV_COLUMNS:
        i = names.index(c)
        col = [row[i] for row in rows]
        assert all(v is not None for v in col)
        if isinstance(col[0], int):
            out[c] = np.asarray(col, dtype=np.int32)
        else:
            out[c] = np.asarray([v.encode("utf-8") for v in col], dtype="S")
    np.savez_compressed(os.path.join(HERE, "test_data_sv.npz"), **out)

    with tarfile.open(f"{REF}/pinot-core/src/test/resources/data/paddingOld.tar.gz") as tf:
        for m in tf.getmembers():
            base = os.path.basename(m.name)
            if base in ("age.dict", "age.sv.unsorted.fwd", "percent.dict", "percent.sv.unsorted.fwd",
                        "outgoingName1.dict", "outgoingName1.sv.unsorted.fwd", "name.dict", "name.sv.unsorted.fwd"):
                data = tf.extractfile(m).read()
                open(os.path.join(HERE, "padding_old_" + base.replace(".", "_") + ".bin"), "wb").write(data)

    # ... and the whole directory as the reference wrote it (v1 layout: metadata.properties + one file per index), for the
    # native segment-directory loader (tests/test_gpu_segment_dir.py)
    with tarfile.open(f"{REF}/pinot-core/src/test/resources/data/paddingOld.tar.gz") as tf:
        os.makedirs(os.path.join(HERE, "paddingOld"), exist_ok=True)
        for m in tf.getmembers():
            if m.isfile():
                open(os.path.join(HERE, "paddingOld", os.path.basename(m.name)), "wb").write(tf.extractfile(m).read())

    os.makedirs(os.path.join(HERE, "raw_forward"), exist_ok=True)
    for name in ("fixedByteRaw.v2", "fixedByteCompressed.v2", "fixedByteSVRDoubles.v1"):
        shutil.copyfile(f"{REF}/pinot-segment-local/src/test/resources/data/{name}", os.path.join(HERE, "raw_forward", name))
        os.chmod(os.path.join(HERE, "raw_forward", name), 0o644)

    st = f"{REF}/pinot-segment-local/src/test/resources/data/startree/segment"
    shutil.copyfile(f"{st}/star_tree_index", os.path.join(HERE, "star_tree_index.bin"))
    with open(f"{st}/star_tree_index_map") as f, open(os.path.join(HERE, "star_tree_index_map.txt"), "w") as g:
        for line in f:
            if line.strip() and not line.startswith("#"):
                g.write(line)
    # the star-tree metadata lines of the segment's metadata.properties (dimension split order, function pairs)
    with open(f"{st}/metadata.properties") as f, open(os.path.join(HERE, "star_tree_metadata.txt"), "w") as g:
        for line in f:
            if line.startswith("startree.") or line.startswith("segment.total.docs") or any(
                    line.startswith(f"column.{c}.") for c in ("AirlineID", "Origin", "Dest", "ArrDelay")):
                g.write(line)
    for fn in sorted(os.listdir(HERE)):
        if os.path.isfile(os.path.join(HERE, fn)):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
