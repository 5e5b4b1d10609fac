"""The reference octree displaces a node when the particle that creates it lies exactly on a
centre plane of the parent cell (tree.c:298-310), and Find_ngb_tree then prunes in-reach
particles underneath it (tree.c:56-58).  workloads.snap_to_cell_planes makes that common at
test sizes; these tests pin the restatement (CPU) and libtoygpu (GPU) against the unmodified
reference on such input."""
import numpy as np
import pytest

import toycluster_b200 as tc
from oracle import port, ref
from toycluster_b200 import workloads

N = 8192
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def wl():
    w = workloads.make("merger_1e6", n_gas=N, seed=3)
    w.pos = workloads.snap_to_cell_planes(w.pos, w.boxsize, 300)
    return w


def _ref(w, threads=4):
    return ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), threads)


@needs_ref
def test_reference_tree_really_prunes(wl):
    """The premise: on this input Find_ngb_tree != Find_ngb_simple for many targets."""
    r = _ref(wl)
    r.load(wl.pos)
    r.find_sph_quantities()
    d = r.read()
    first, count = port.tree_displaced(d["pos"], wl.boxsize)
    assert len(first) > 20
    differ = 0
    for i in range(0, N, 37):
        h = float(d["hsml"][i])
        a, s = r.find_ngb_tree(i, h), r.find_ngb_simple(i, h)
        assert set(a) <= set(s)
        differ += len(a) != len(s)
        assert np.array_equal(port.find_ngb(d["pos"], wl.boxsize, i, h), a)
        assert np.array_equal(port.find_ngb_simple(d["pos"], wl.boxsize, i, h), s)
    assert differ > 10


@needs_ref
def test_port_matches_reference_with_displaced_nodes(wl):
    r = _ref(wl)
    r.load(wl.pos)
    snaps = []

    def cb(it):
        s = r.read()
        if it > 0:
            s["hw"], s["delta"] = r.wvt_scratch()
        snaps.append(s)
        return 0

    niter = 2
    r.regularise(niter, cb)
    rows, state, states = port.regularise(wl, wl.pos, max_iters=niter, keep=True)
    for it in range(niter):
        a, b = states[it], snaps[it + 1]
        assert np.array_equal(a["id"], b["id"])
        for k in ("hsml", "rho", "varhsml", "rho_model", "hw", "delta", "pos"):
            assert np.array_equal(a[k], b[k]), (it, k)


# ---------------------------------------------------------------------------- GPU


@pytest.mark.gpu
@needs_ref
def test_gpu_neighbour_sets_follow_the_reference_tree(wl):
    r = _ref(wl)
    r.load(wl.pos)
    r.find_sph_quantities()
    d = r.read()
    g = tc.HotPath.from_workload(wl)
    g.upload(wl.pos)
    g.sort()
    st = g.stats()
    assert st["displaced_nodes"] > 20 and st["displaced_particles"] > 50
    assert st["displaced_overflow"] == 0
    e = tc.HotPath.from_workload(wl, flags=tc.EXACT_NEIGHBOURS)
    e.upload(wl.pos)
    e.sort()
    assert e.stats()["displaced_nodes"] == 0
    differ = 0
    for i in range(0, N, 23):
        for h in (float(d["hsml"][i]), float(d["hsml"][i]) * 1.7, 0.2 * wl.boxsize):
            a = r.find_ngb_tree(i, h)
            assert np.array_equal(g.find_ngb(i, h), a), (i, h)
            s = r.find_ngb_simple(i, h)
            assert np.array_equal(e.find_ngb(i, h), s), (i, h)
            differ += len(a) != len(s)
    assert differ > 20


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("tiles", [True, False])
def test_gpu_iterations_bit_exact_with_displaced_nodes(wl, tiles, monkeypatch):
    """Cold start + 3 WVT iterations in the reference's accumulation order: every array
    bit-identical, on the tile path (hand-backs) and on the generic path alone."""
    if not tiles:
        monkeypatch.setenv("TOYGPU_NO_TILES", "1")
    r = _ref(wl)
    r.load(wl.pos)
    after = []

    def cb(it):
        s = r.read()
        if it > 0:
            s["hw"], s["delta"] = r.wvt_scratch()
            after.append(s)
        return 0

    niter = 3
    r.regularise(niter + 1, cb)
    log = ref.parse_log(r.log())
    g = tc.HotPath.from_workload(wl, flags=tc.WVT_SEQUENTIAL)
    g.upload(wl.pos)
    for it in range(niter):
        g.wvt_iteration(log[it + 1]["step"])
        s, o = after[it], g.download()
        hw, dl = g.wvt_scratch()
        assert np.array_equal(o["id"], s["id"]), it
        for k in ("rho_model", "hsml", "rho", "varhsml", "pos"):
            assert np.array_equal(o[k], s[k]), (it, k, (o[k] != s[k]).mean())
        assert np.array_equal(hw, s["hw"]), it
        assert np.array_equal(dl, s["delta"]), it


@pytest.mark.gpu
@needs_ref
def test_gpu_default_mode_and_rotA_with_displaced_nodes(wl):
    """Default (tile + tree-sum) mode: rho/hsml bit-exact, displacement within 1e-5; rot(A)
    over the pruned neighbour sets within 1e-5 of the reference."""
    r = _ref(wl)
    r.load(wl.pos)
    r.find_sph_quantities()
    d0 = r.read()
    g = tc.HotPath.from_workload(wl)
    g.upload(wl.pos)
    g.find_sph_quantities()
    o = g.download()
    for k in ("hsml", "rho", "varhsml"):
        assert np.array_equal(o[k], d0[k]), k
    # warm pass from the reference's state: tile path
    r.find_sph_quantities()
    d1 = r.read()
    g.find_sph_quantities()
    o = g.download()
    assert g.stats()["displaced_particles"] > 50      # the tile sweep applies the open tests itself
    for k in ("hsml", "rho", "varhsml"):
        assert np.array_equal(o[k], d1[k]), k
    rng = np.random.default_rng(5)
    apot = rng.standard_normal((N, 3)).astype(np.float32)
    assert np.array_equal(o["id"], d1["id"])
    r.set_apot(apot)                        # current (Peano) order on both sides
    r.bfld_from_rotA()
    want = r.read()["bfld"]
    g.set_apot(apot)
    g.bfld_from_rotA_sph()
    got = g.download(bfld=True)["bfld"]
    scale = np.abs(want).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(got - want) / scale).max() < 1e-5


@pytest.mark.gpu
@needs_ref
def test_gpu_particle_on_the_box_face_is_filed_elsewhere():
    """Pos == Boxsize is legal (the wrap of wvt_relax.c:200 is inclusive) but scales to 2^63,
    whose Peano key is not the key of the cell the particle sits in: the reference files it in
    a leaf somewhere else and its neighbours do not find it.  Found by scripts/fuzz_parity.py
    with periodically shifted inputs."""
    w = workloads.make("merger_1e6", n_gas=20011, seed=5)
    off = np.array([0.31, 0.57, 0.83]) * w.boxsize
    pos = np.mod(w.pos.astype(np.float64) + off, w.boxsize).astype(np.float32)
    pos[pos >= np.float32(w.boxsize)] = 0
    # the densest region now straddles the faces: put a few of its particles exactly ON a face
    d = np.abs(pos - np.float32(w.boxsize)).min(axis=1)
    pick = np.argsort(d)[:6]
    for j, i in enumerate(pick):
        pos[i, j % 3] = np.float32(w.boxsize)
    w.pos = pos
    r = _ref(w, 8)
    r.load(w.pos)
    after = []

    def cb(it):
        s = r.read()
        if it > 0:
            s["hw"], s["delta"] = r.wvt_scratch()
            after.append(s)
        return 0

    r.regularise(3, cb)
    g = tc.HotPath.from_workload(w, flags=tc.WVT_SEQUENTIAL)
    g.upload(w.pos)
    for it in range(2):
        g.wvt_iteration(0.0085)
        if it == 0:
            assert g.stats()["displaced_particles"] >= 6
        s, o = after[it], g.download()
        hw, dl = g.wvt_scratch()
        for k in ("id", "rho_model", "hsml", "rho", "varhsml", "pos"):
            assert np.array_equal(o[k], s[k]), (it, k, int((o[k] != s[k]).sum()))
        assert np.array_equal(dl, s["delta"]), it


@needs_ref
def test_port_files_a_particle_on_the_box_face_like_the_reference():
    """CPU: a coordinate equal to Boxsize has the key of 2^63, not of its cell; the restated
    tree (like the reference's) then misses the particle from where it really is."""
    w = workloads.make("merger_1e6", n_gas=6000, seed=5)
    pos = np.mod(w.pos.astype(np.float64) + np.array([0.31, 0.57, 0.83]) * w.boxsize, w.boxsize).astype(np.float32)
    pos[pos >= np.float32(w.boxsize)] = 0
    pick = np.argsort(np.abs(pos - np.float32(w.boxsize)).min(axis=1))[:4]
    for j, i in enumerate(pick):
        pos[i, j % 3] = np.float32(w.boxsize)
    w.pos = pos
    r = _ref(w)
    r.load(w.pos)
    r.find_sph_quantities()
    d = r.read()
    o = port.find_sph_quantities(w, w.pos)
    for k in ("id", "pos", "hsml", "rho", "varhsml"):
        assert np.array_equal(o[k], d[k]), k
    on_face = np.flatnonzero((d["pos"] == np.float32(w.boxsize)).any(axis=1))
    assert len(on_face) >= 4
    missed = 0
    for i in on_face:
        # its nearest neighbours look for it with a generous radius
        dist = np.abs(d["pos"] - d["pos"][i])
        dist = np.minimum(dist, np.float32(w.boxsize) - dist)
        near = np.argsort((dist ** 2).sum(1))[1:6]
        for t in near:
            h = float(2 * d["hsml"][t])
            a, s = r.find_ngb_tree(int(t), h), r.find_ngb_simple(int(t), h)
            assert np.array_equal(port.find_ngb(d["pos"], w.boxsize, int(t), h), a)
            missed += (i in s) and (i not in a)
    assert missed > 0


@pytest.mark.gpu
def test_gpu_lattice_input_overflows_the_event_table_loudly():
    """A lattice whose planes are cell-centre planes displaces nodes everywhere: more events than
    the table (n/4 entries) holds.  The neighbour sets could then no longer be the reference's,
    so the call fails with a message instead of returning something else silently;
    with TG_EXACT_NEIGHBOURS nothing is flagged at all and the run completes."""
    w = workloads.make("single_1e5", n_gas=32768)
    k = np.arange(32, dtype=np.float64)
    gx, gy, gz = np.meshgrid(k, k, k, indexing="ij")
    lattice = np.stack([gx, gy, gz], -1).reshape(-1, 3) / 32 * w.boxsize     # j/32: every point on a centre plane
    rng = np.random.default_rng(3)
    pos = lattice.astype(np.float32)
    jitter = rng.uniform(-0.2, 0.2, pos.shape) * w.boxsize / 32
    move = rng.random(len(pos)) < 0.5                     # half the particles leave the planes
    pos[move] = np.mod(pos[move].astype(np.float64) + jitter[move], w.boxsize).astype(np.float32)
    pos = np.clip(pos, 0, np.float32(w.boxsize))
    g = tc.HotPath.from_workload(w)
    g.upload(pos)
    with pytest.raises(tc.ToyGpuError, match="displaced reference-tree nodes"):
        g.find_sph_quantities()
    with pytest.raises(tc.ToyGpuError, match="upload the particles again"):
        g.find_sph_quantities()                           # ids and positions disagree after a failure
    g.upload(pos)
    assert g.sort() is not None                           # sorting alone does not need the paths
    e = tc.HotPath.from_workload(w, flags=tc.EXACT_NEIGHBOURS)
    e.upload(pos)
    e.find_sph_quantities()
    assert e.stats()["displaced_nodes"] == 0 and e.stats()["displaced_overflow"] == 0
    o = e.download()
    assert np.isfinite(o["hsml"]).all() and np.isfinite(o["rho"]).all() and (o["hsml"] > 0).all()
