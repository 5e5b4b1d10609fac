"""The CPU restatement (oracle/toy_oracle.c) against the golden fixtures produced by the
reference's own unmodified code (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import port
from toycluster_b200 import workloads

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "merger_4096.npz"))


@pytest.fixture(scope="module")
def wl(gold):
    w = workloads.make("merger_1e6", n_gas=int(gold["n_gas"]), with_positions=False)
    # the fixture carries the exact scalars it was generated with
    assert w.boxsize == float(gold["boxsize"])
    assert np.array_equal(w.halo_table(), gold["halo_table"])
    assert w.mpart_gas == float(gold["mpart_gas"]) and w.mtotal == float(gold["mtotal"])
    return w


def test_peano_known_answers():
    with open(os.path.join(GOLD, "peano_table.json")) as f:
        table = json.load(f)
    assert len(table) >= 50
    for row in table:
        hi, lo = port.peano_key(*row["xyz"])
        assert [f"{hi:016x}", f"{lo:016x}"] == row["key"], row
        hi, lo = port.peano_key(*row["xyz"], reversed_=True)
        assert [f"{hi:016x}", f"{lo:016x}"] == row["reversed"], row


def test_survey_table_vectors():
    """The hand-checkable rows of SURVEY.md section 8(c)."""
    assert port.peano_key(0.25, 0.25, 0.75) == (0x3400000000000000, 0)
    assert port.peano_key(0.5, 0.5, 0.5) == (0xa000000000000000, 0)
    assert port.peano_key(0.1, 0.2, 0.3) == (0x1da0de8fc85c0de8, 0xfc85c0de8fc85c0c)
    assert port.peano_key(1.0, 0.3, 0.3) == (0x0808808808808808, 0x8088088088088088)   # x == 1 edge
    assert port.peano_key(0.5, 0.5, 0.5, True) == (0, 0x28)


def test_sort_order_and_keys(gold):
    perm, hi, lo, dup = port.sort(gold["pos0"], float(gold["boxsize"]))
    assert dup == 0
    assert np.array_equal(perm, gold["sort_id"])
    assert np.array_equal(hi, gold["sort_key_hi"]) and np.array_equal(lo, gold["sort_key_lo"])
    key = (hi.astype(object) << 64) | lo.astype(object)
    assert all(key[k] < key[k + 1] for k in range(len(key) - 1))       # sortedness


def test_guess_hsml(gold):
    pos_sorted = gold["pos0"][gold["sort_id"]]
    assert np.array_equal(port.guess_hsml(pos_sorted, float(gold["boxsize"])), gold["guess2"])


def test_neighbour_lists(gold):
    pos_sorted = gold["pos0"][gold["sort_id"]]
    off = 0
    saw_full = False
    for i, h, cnt in gold["ngb_queries"]:
        want = gold["ngb_lists"][off:off + int(cnt)]
        off += int(cnt)
        got = port.find_ngb(pos_sorted, float(gold["boxsize"]), int(i), np.float32(h))
        assert np.array_equal(got, want), (i, h)
        assert np.all(np.diff(got) > 0)
        saw_full |= len(got) == port.NGBMAX
    assert saw_full      # the fixture exercises the NGBMAX cut of tree.c:91-92


def test_wvt_iterations_bit_exact(gold, wl):
    rows, state, states = port.regularise(wl, gold["pos0"], max_iters=3, keep=True)
    log = gold["log"]
    for it in range(3):
        s = states[it]
        assert np.array_equal(s["id"], gold[f"it{it}_id"])
        for name in ("hsml", "rho", "varhsml", "rho_model", "hw", "delta", "pos"):
            assert np.array_equal(s[name], gold[f"it{it}_{name}"]), (it, name)
        assert float("%g" % rows[it]["max"]) == log[it][1]
        assert float("%g" % rows[it]["mean"]) == log[it][2]
        assert rows[it]["step"] == log[it][4]


def test_final_density_and_rotA(gold, wl):
    st = port.find_sph_quantities(wl, gold["it3_pos"], gold["it3_hsml"], gold["it3_id"])
    assert np.array_equal(st["id"], gold["final_id"])
    for name in ("pos", "hsml", "rho", "varhsml"):
        assert np.array_equal(st[name], gold[f"final_{name}"]), name
    b = port.bfld_from_rotA(wl, st, gold["final_apot"])
    assert np.array_equal(b, gold["final_bfld"])


def test_displacement_noise_floor(gold, wl):
    """How far the reference's own float accumulation (wvt_relax.c:167-169) sits from the
    exactly summed displacement: this is the floor for any other summation order."""
    it = 2
    pos = gold[f"it{it}_pos"]          # only used for shape; recompute from the state before
    st = port.find_sph_quantities(wl, gold["it1_pos"], gold["it1_hsml"], gold["it1_id"])
    out = port.wvt_iteration(wl, None, None, float(gold["log"][it + 1][4]), dens=st)
    assert np.array_equal(out["delta"], gold[f"it{it}_delta"])
    assert pos.shape == out["pos"].shape
