"""The table behind k_peano_keys (toycluster_b200/csrc/peano_lut.cuh): Peano_Key (peano.c:128-203)
restated as a 48-state transducer over bit planes by scripts/make_peano_lut.py.  CPU only: the
table, driven exactly as the kernel drives it, must give the oracle's keys -- including the
golden table of SURVEY 8c and the edge where a coordinate equals Boxsize (plane 63 set) -- and
the committed header must be what the generator writes."""
import importlib.util
import json
import os
import re

import numpy as np

from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("make_peano_lut", os.path.join(ROOT, "scripts", "make_peano_lut.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def _key(x, y, z, lut1, lut2):
    m = 1 << 63
    return gen.key_from_tables([int(y * m), int(z * m), int(x * m)], lut1, lut2)


def test_table_reproduces_the_oracle_keys():
    n, lut1, lut2 = gen.tables()
    assert n == 48 and len(lut1) == 48 * 8 and len(lut2) == 48 * 64
    rng = np.random.default_rng(5)
    pts = [(0, 0, 0), (1, 1, 1), (1.0, 0.3, 0.3), (0.3, 1.0, 0.3), (0.3, 0.3, 1.0), (0.5, 0.5, 0.5),
           (0.25, 0.75, 0.25), (0.999999, 1e-06, 0.5)]
    box = 13923.0
    pts += [tuple(float(np.float32(v * box)) / box for v in rng.random(3)) for _ in range(5000)]
    pts += [tuple(rng.random(3)) for _ in range(5000)]
    for x, y, z in pts:
        hi, lo = port.peano_key(float(x), float(y), float(z))
        assert _key(float(x), float(y), float(z), lut1, lut2) == (hi << 64 | lo), (x, y, z)


def test_golden_table_rows():
    path = os.path.join(ROOT, "tests", "golden", "peano_table.json")
    rows = json.load(open(path))
    n, lut1, lut2 = gen.tables()
    assert len(rows) >= 10
    for r in rows:
        x, y, z = (float(v) for v in r["xyz"])
        want = int(r["key"][0], 16) << 64 | int(r["key"][1], 16)
        assert _key(x, y, z, lut1, lut2) == want, r


def test_committed_header_is_the_generated_one():
    n, lut1, lut2 = gen.tables()
    text = open(os.path.join(ROOT, "toycluster_b200", "csrc", "peano_lut.cuh")).read()
    for name, lut in (("PEANO_LUT1", lut1), ("PEANO_LUT2", lut2)):
        body = re.search(name + r"\[\d+\] = \{(.*?)\};", text, re.S).group(1)
        assert [int(v) for v in body.replace("\n", " ").split(",") if v.strip()] == lut, name
