"""Pins the CPU restatement against the reference itself (oracle/_ref: the unmodified
sources compiled by oracle/Makefile) on seeded random workloads.  CPU only; skipped where the
compiled reference is not available."""
import numpy as np
import pytest

from oracle import port, ref
from toycluster_b200 import workloads

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("name,n,seed", [("single_1e5", 4096, 1), ("merger_1e6", 6144, 2)])
def test_port_matches_reference(name, n, seed):
    w = workloads.make(name, n_gas=n, seed=seed)
    r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), 4)
    r.load(w.pos)
    snaps = []

    def cb(it):
        s = r.read()
        if it > 0:
            s["hw"], s["delta"] = r.wvt_scratch()
        snaps.append(s)
        return 0

    niter = 3
    r.regularise(niter, cb)
    log = ref.parse_log(r.log())
    rows, state, states = port.regularise(w, w.pos, max_iters=niter, keep=True)
    assert len(rows) == len(log) == niter
    for it in range(niter):
        a, b = states[it], snaps[it + 1]
        assert np.array_equal(a["id"], b["id"])
        for k in ("hsml", "rho", "varhsml", "rho_model", "hw", "delta", "pos"):
            assert np.array_equal(a[k], b[k]), (it, k)
        assert float("%g" % rows[it]["mean"]) == log[it]["mean"]


def test_empty_search_and_ragged_tail():
    """Edge cases: a radius that finds only the particle itself, one that finds everything
    (cut at NGBMAX), and a particle count that is not a multiple of the 32-wide groups."""
    w = workloads.make("merger_1e6", n_gas=4099, seed=3)
    r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), 4)
    r.load(w.pos)
    r.sort()
    r.build_tree()
    d = r.read()
    for i in (0, 4098, 2000):
        for h in (1e-3, 50.0, 1e5):
            a = r.find_ngb_tree(i, h)
            b = port.find_ngb(d["pos"], w.boxsize, i, h)
            assert np.array_equal(a, b), (i, h)
    assert np.array_equal(port.find_ngb(d["pos"], w.boxsize, 5, 1e-3), [5])
    assert len(port.find_ngb(d["pos"], w.boxsize, 5, 1e5)) == port.NGBMAX
