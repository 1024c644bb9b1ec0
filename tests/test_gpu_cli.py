"""-m gpu: the command line face end to end -- Pinot segment directories on disk (v1 and v3, with a compressed raw column)
-> pb200h_segment_load_dir -> one submission -> merge by value -> final rows, against the oracle-side merge."""
import json

import numpy as np
import pytest

from pinot_b200 import cli, sql
from reduce_util import combine, normalise, reduce_rows
from segment_dir_util import write_segment_dir

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("on_device", [False, True])
def test_cli_over_segment_directories(oracle, tmp_path, capsys, on_device):
    rng = np.random.default_rng(21)
    segs, dirs = [], []
    for i, version in enumerate(["v1", "v3", "v3"]):
        n = 9_000 + 13 * i
        seg = oracle.build_segment(f"cli{i}", {
            "k": rng.integers(0, 6 + i, size=n).astype(np.int32) * 5, "v": rng.integers(-300, 300, size=n).astype(np.int32),
            "price": (rng.integers(0, 500, size=n) / 4.0).astype(np.float64)}, raw=["price"], raw_compression={"price": 3})
        segs.append(seg)
        dirs.append(write_segment_dir(str(tmp_path / f"seg{i}"), seg, version))
    text = "SELECT COUNT(*), SUM(price), MAX(v), AVG(v), DISTINCTCOUNT(v) FROM t WHERE v > -250 GROUP BY k LIMIT 50"
    argv = [text, "--json"] + [x for d in dirs for x in ("--segment-dir", d)] + (["--merge-on-device"] if on_device else [])
    assert cli.main(argv) == 0
    out = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    q = sql.parse(text)
    fns = [a.function for a in q.aggregations]
    want = dict(reduce_rows(fns, combine(fns, [normalise(s, q, r.num_groups, r.keys, r.doubles, r.longs, r.distinct)
                                                for s, r in ((s, oracle.execute(s, q)) for s in segs)])))
    assert len(out["rows"]) == len(want) <= 50
    for row in out["rows"]:
        key, vals = (row[0],), row[1:]
        w = want[key]
        assert vals[0] == w[0] and vals[2] == w[2] and vals[4] == w[4], (key, vals, w)
        assert abs(vals[1] - w[1]) <= 1e-6 * max(1.0, abs(w[1])) and abs(vals[3] - w[3]) <= 1e-9 * max(1.0, abs(w[3]))
    assert out["stats"]["numSegments"] == 3 and out["stats"]["totalDocs"] == sum(s.num_docs for s in segs)
