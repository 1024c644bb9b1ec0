"""Writes oracle-built index buffers as a Pinot segment DIRECTORY (test helper for the native segment loader).

Layouts restated from the reference (the bytes inside the files come from oracle/segment_builder.py, which is pinned to
golden bytes): v1 = one file per index, ``<col>.sv.unsorted.fwd`` / ``.sv.sorted.fwd`` / ``.sv.raw.fwd`` / ``.dict`` /
``.bitmap.inv`` + ``metadata.properties`` (``pinot-segment-spi/.../V1Constants.java:49-99``); v3 = ``v3/columns.psf`` in which
every index entry starts with the 8-byte magic 0xdeadbeefdeafbead, + ``v3/index_map`` with ``<col>.<index>.startOffset|size``
(``pinot-segment-local/.../store/SingleFileIndexDirectory.java:72-73,175-206,306``); star-tree = ``star_tree_index`` +
``star_tree_index_map`` + ``startree.v2.*`` metadata keys (``.../startree/v2/store/StarTreeIndexMapUtils.java``).
"""
from __future__ import annotations

import os
import struct

TYPE_NAMES = {0: "INT", 1: "LONG", 2: "FLOAT", 3: "DOUBLE", 4: "STRING"}
MAGIC = struct.pack(">Q", 0xDEADBEEFDEAFBEAD)


def _metadata_lines(seg, star_tree=None):
    # segment.padding.character: SegmentColumnarIndexCreator.java:484 writes the zero pad character, which the properties
    # writer escapes as \\u0000; the loaders (the reference's and ours) accept STRING dictionaries only with it
    lines = [f"segment.name = {seg.name}", f"segment.total.docs = {seg.num_docs}", "segment.padding.character = \\u0000"]
    for c in seg.columns:
        p = f"column.{c.name}."
        lines += [p + f"cardinality = {c.cardinality}", p + f"totalDocs = {seg.num_docs}", p + f"dataType = {TYPE_NAMES[c.data_type]}",
                  p + f"bitsPerElement = {c.bits}", p + f"lengthOfEachEntry = {c.dict_entry_bytes if c.data_type == 4 else 0}",
                  p + f"isSorted = {'true' if c.is_sorted else 'false'}", p + f"hasDictionary = {'true' if c.has_dictionary else 'false'}",
                  p + "isSingleValues = true"]
    if star_tree is not None:
        lines += ["startree.v2.count = 1", f"startree.v2.0.total.docs = {star_tree.num_docs}",
                  f"startree.v2.0.max.leaf.records = {star_tree.max_leaf_records}"]
        lines += [f"startree.v2.0.split.order = {d}" for d in star_tree.dimensions]  # repeated keys, as the reference writes lists
        lines += [f"startree.v2.0.function.column.pairs = {star_tree.metric_name(i)}" for i in range(len(star_tree.function_pairs))]
    return lines


def _write_star_tree(directory, st):
    nd = len(st.dimensions)
    parts = [("null", "STAR_TREE", st.tree.tobytes())]
    parts += [(st.dimensions[j], "FORWARD_INDEX", st.segment.columns[j].fwd.tobytes()) for j in range(nd)]
    parts += [(st.metric_name(i), "FORWARD_INDEX", st.segment.columns[nd + i].fwd.tobytes()) for i in range(len(st.function_pairs))]
    off, blob, lines = 0, b"", []
    for col, kind, data in parts:
        lines += [f"0.{col}.{kind}.OFFSET = {off}", f"0.{col}.{kind}.SIZE = {len(data)}"]
        blob += data
        off += len(data)
    with open(os.path.join(directory, "star_tree_index"), "wb") as f:
        f.write(blob)
    with open(os.path.join(directory, "star_tree_index_map"), "w") as f:
        f.write("\n".join(lines) + "\n")


def write_segment_dir(path: str, seg, version: str = "v1", star_tree=None) -> str:
    """Returns the directory to hand to IndexSegment.load (the segment root for both versions)."""
    os.makedirs(path, exist_ok=True)
    directory = path
    if version == "v3":
        directory = os.path.join(path, "v3")
        os.makedirs(directory, exist_ok=True)
        blob, lines = b"", []
        for c in seg.columns:
            entries = [("forward_index", c.fwd)]
            if c.has_dictionary:
                entries.append(("dictionary", c.dict))
            if c.inv is not None:
                entries.append(("inverted_index", c.inv))
            for kind, buf in entries:
                data = MAGIC + buf.tobytes()
                lines += [f"{c.name}.{kind}.startOffset = {len(blob)}", f"{c.name}.{kind}.size = {len(data)}"]
                blob += data
        with open(os.path.join(directory, "columns.psf"), "wb") as f:
            f.write(blob)
        with open(os.path.join(directory, "index_map"), "w") as f:
            f.write("\n".join(lines) + "\n")
    else:
        for c in seg.columns:
            ext = ".sv.raw.fwd" if not c.has_dictionary else ".sv.sorted.fwd" if c.is_sorted else ".sv.unsorted.fwd"
            with open(os.path.join(directory, c.name + ext), "wb") as f:
                f.write(c.fwd.tobytes())
            if c.has_dictionary:
                with open(os.path.join(directory, c.name + ".dict"), "wb") as f:
                    f.write(c.dict.tobytes())
            if c.inv is not None:
                with open(os.path.join(directory, c.name + ".bitmap.inv"), "wb") as f:
                    f.write(c.inv.tobytes())
    with open(os.path.join(directory, "metadata.properties"), "w") as f:
        f.write("\n".join(_metadata_lines(seg, star_tree)) + "\n")
    if star_tree is not None:
        _write_star_tree(directory, star_tree)
    return path
