"""GPU parity against the committed golden fixtures (no reference needed at run time), plus
the properties that do not depend on size: tile path == generic path, rank slices == one rank,
AoS records travel with their particle."""
import os

import numpy as np
import pytest

import toycluster_b200 as tc
from toycluster_b200 import workloads
from toycluster_b200.dist import rank_slice

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "merger_4096.npz"))


def _ctx(gold, **kw):
    return tc.HotPath(int(gold["n_gas"]), float(gold["boxsize"]), float(gold["mpart_gas"]),
                      float(gold["mtotal"]), gold["halo_table"], **kw)


def test_golden_sort_guess_neighbours(gold):
    g = _ctx(gold)
    g.upload(gold["pos0"])
    assert np.array_equal(g.sort(), gold["sort_id"])
    assert np.array_equal(g.guess_hsml(), gold["guess2"])
    off = 0
    for i, h, cnt in gold["ngb_queries"]:
        want = gold["ngb_lists"][off:off + int(cnt)]
        off += int(cnt)
        assert np.array_equal(g.find_ngb(int(i), np.float32(h)), want), (i, h)


def test_golden_iterations_sequential(gold):
    g = _ctx(gold, flags=tc.WVT_SEQUENTIAL)
    g.upload(gold["pos0"])
    for it in range(4):
        g.wvt_iteration(float(gold["log"][it + 1][4]) if it + 1 < len(gold["log"]) else 0.0085)
        o = g.download()
        hw, dl = g.wvt_scratch()
        assert np.array_equal(o["id"], gold[f"it{it}_id"])
        for k in ("hsml", "rho", "varhsml", "rho_model", "pos"):
            assert np.array_equal(o[k], gold[f"it{it}_{k}"]), (it, k)
        assert np.array_equal(hw, gold[f"it{it}_hw"]) and np.array_equal(dl, gold[f"it{it}_delta"])
        if it == 2:
            break


def test_golden_final_density_and_rotA(gold):
    """Find_sph_quantities + Bfld_from_rotA_SPH as main.c:54-56 calls them."""
    g = _ctx(gold)
    g.upload(gold["it3_pos"], gold["it3_hsml"])
    g.find_sph_quantities()
    o = g.download()
    assert np.array_equal(gold["it3_id"][o["id"]], gold["final_id"])
    for k in ("pos", "hsml", "rho", "varhsml"):
        assert np.array_equal(o[k], gold[f"final_{k}"]), k
    g.set_apot(gold["final_apot"])           # current (Peano) order, like magnetic_field.c:33-69
    g.bfld_from_rotA_sph()
    b = g.download(bfld=True)["bfld"]
    want = gold["final_bfld"]
    rel = np.abs(b.astype(np.float64) - want) / np.maximum(np.abs(want), np.abs(want).max() * 1e-6)
    assert rel.max() <= 1e-5, rel.max()
    assert (b == want).mean() > 0.99


def test_tile_path_equals_generic_path():
    """The tile sweep and the generic sweep are two schedules of the same arithmetic."""
    w = workloads.make("merger_1e6", n_gas=150_000)
    g1 = tc.HotPath.from_workload(w)
    os.environ["TOYGPU_NO_TILES"] = "1"
    try:
        g2 = tc.HotPath.from_workload(w)
    finally:
        del os.environ["TOYGPU_NO_TILES"]
    for g in (g1, g2):
        g.upload(w.pos)
    tiled = 0
    for it in range(4):
        for g in (g1, g2):
            g.find_sph_quantities()
        a, b = g1.download(), g2.download()
        for k in ("id", "pos", "hsml", "rho", "varhsml"):
            assert np.array_equal(a[k], b[k]), (it, k)
        if it > 0:
            tiled += w.n_gas - g1.stats()["handed_back"]
    assert tiled > 2 * w.n_gas          # the fast path really ran


def test_rank_slices_equal_single_rank():
    """Two contexts on one GPU play ranks 0 and 1 of 2 (host copies stand in for the
    all-gather): positions, Hsml and densities equal the one-rank run bit for bit."""
    w = workloads.make("merger_1e6", n_gas=40_003)          # ragged on purpose
    one = tc.HotPath.from_workload(w, flags=tc.WVT_SEQUENTIAL)
    parts = [tc.HotPath.from_workload(w, flags=tc.WVT_SEQUENTIAL, rank=r, nranks=2) for r in range(2)]
    state_pos, state_h = w.pos, None
    one.upload(w.pos)
    for it in range(3):
        one.wvt_iteration(0.0085)
        ref = one.download()
        merged = {}
        for r, g in enumerate(parts):
            g.upload(state_pos, state_h)
            g.wvt_iteration(0.0085)
            o = g.download()
            lo, hi, _ = rank_slice(w.n_gas, r, 2)
            for k in ("pos", "hsml", "rho", "varhsml", "id"):
                merged.setdefault(k, np.empty_like(o[k]))[lo:hi] = o[k][lo:hi]
        # ids of the slice runs are relative to the re-uploaded order; map back
        if it == 0:
            ids = merged["id"]
        else:
            ids = prev_ids[merged["id"]]
        for k in ("pos", "hsml", "rho", "varhsml"):
            assert np.array_equal(merged[k], ref[k]), (it, k)
        assert np.array_equal(ids, ref["id"])
        prev_ids, state_pos, state_h = ids, merged["pos"], merged["hsml"]


def test_records_travel_with_the_particle():
    """tg_upload / tg_download on the driver's AoS records (globals.h:161-180)."""
    w = workloads.make("merger_1e6", n_gas=20_000)
    n = w.n_gas
    P = np.zeros(n, dtype=np.dtype([("Pos", "3f4"), ("Vel", "3f4"), ("ID", "i4"), ("Type", "i4"),
                                     ("Key", "2u8"), ("Tree_Parent", "i4"), ("pad", "3i4")]))
    S = np.zeros(n, dtype=np.dtype([("U", "f4"), ("Rho", "f4"), ("Hsml", "f4"), ("VarHsmlFac", "f4"),
                                     ("Bfld", "3f4"), ("Apot", "3f4"), ("ID", "f4"),
                                     ("Rho_Model", "f4"), ("Rs", "3f4")]))
    assert P.itemsize == 64 and S.itemsize == 60
    P["Pos"], P["ID"], P["Vel"][:, 0] = w.pos, np.arange(n) * 7 + 3, np.arange(n)
    S["U"], S["ID"] = np.arange(n) * 0.5, np.arange(n) * 7 + 3
    g = tc.HotPath.from_workload(w)
    g.upload_records(P, S)
    g.find_sph_quantities()
    o = g.download()
    g.download_records(P, S)
    assert np.array_equal(P["ID"], o["id"] * 7 + 3)             # whole records were permuted
    assert np.array_equal(P["Vel"][:, 0], o["id"].astype(np.float32))
    assert np.array_equal(S["U"], o["id"] * np.float32(0.5)) and np.array_equal(S["ID"], P["ID"])
    assert np.array_equal(P["Pos"], o["pos"]) and np.array_equal(P["Pos"], w.pos[o["id"]])
    assert np.array_equal(S["Hsml"], o["hsml"]) and np.array_equal(S["Rho"], o["rho"])
    assert np.array_equal(S["VarHsmlFac"], o["varhsml"]) and np.array_equal(S["Rho_Model"], o["rho_model"])
    key = (P["Key"][:, 1].astype(object) << 64) | P["Key"][:, 0].astype(object)
    assert all(key[k] < key[k + 1] for k in range(n - 1))       # Peano order of the last sort


def test_error_paths_and_edges():
    """Failure is loud (status + message), like the reference's Assert (aux.c:57-83)."""
    w = workloads.make("merger_1e6", n_gas=5000)
    # fewer gas particles than DESNNGB in reach: sph.c:36-64 would spin forever
    few = tc.HotPath(200, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table())
    few.upload(w.pos[:200])
    with pytest.raises(tc.ToyGpuError, match="did not terminate"):
        few.find_sph_quantities()
    # a coordinate outside [0, Boxsize]: peano.c:130-132 asserts
    bad = w.pos.copy()
    bad[17, 1] = np.float32(w.boxsize * 1.5)
    g = tc.HotPath.from_workload(w)
    g.upload(bad)
    with pytest.raises(tc.ToyGpuError, match="outside"):
        g.find_sph_quantities()
    # no halo table / no index / no Apot
    h = tc.HotPath.from_workload(w)
    h.upload(w.pos)
    with pytest.raises(tc.ToyGpuError, match="index"):
        h.bfld_from_rotA_sph()
    # x == Boxsize exactly is legal (wvt_relax.c:200 leaves it) and keeps its odd key
    edge = w.pos.copy()
    edge[5] = (np.float32(w.boxsize), np.float32(0.3 * w.boxsize), np.float32(0.3 * w.boxsize))
    g2 = tc.HotPath.from_workload(w)
    g2.upload(edge)
    hi, lo = g2.peano_keys()
    from oracle import port
    assert (int(hi[5]), int(lo[5])) == port.peano_key(1.0, float(edge[5, 1]) / w.boxsize,
                                                      float(edge[5, 2]) / w.boxsize)
    g2.find_sph_quantities()
    want = port.find_sph_quantities(w, edge)
    got = g2.download()
    assert np.array_equal(got["id"], want["id"])
    for k in ("hsml", "rho", "varhsml"):
        assert np.array_equal(got[k], want[k]), k


def test_make_magnetic_field_equals_its_three_stages():
    """tg_make_magnetic_field (magnetic_field.c:12-131 on the device) against the same three
    stages done separately: the vector potential restated in numpy, the rot(A) operator that
    test_golden_final_density_and_rotA pins against the reference, and the normalisation /
    cap in numpy.  (The whole-program check is tests/test_driver_e2e.py.)"""
    from toycluster_b200 import workloads
    w = workloads.make("merger_1e6", n_gas=12288)
    g = tc.HotPath.from_workload(w)
    g.upload(w.pos)
    g.find_sph_quantities()
    o = g.download()
    pos = o["pos"]
    boxhalf = np.float32(0.5 * w.boxsize)
    # magnetic_field.c:33-69
    amax = np.zeros(len(pos))
    for h in w.halos:
        if h.mass_gas == 0:
            continue
        d = (pos.astype(np.float64) - np.asarray(h.dcom) - np.float64(boxhalf)).astype(np.float32)
        r2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]          # float expression
        rho = workloads.gas_density_profile(np.sqrt(r2.astype(np.float64)), h.rho0, h.beta, h.rcore, h.rcut)
        amax = np.maximum(amax, np.power(rho / h.rho0, 0.5))
    apot = np.repeat(amax.astype(np.float32)[:, None], 3, 1)
    g.set_apot(apot)
    g.bfld_from_rotA_sph()
    b = g.download(bfld=True)["bfld"]
    for bnorm in (20e-6, 80e-6):
        norm, capped = g.make_magnetic_field(bnorm, 0.5)
        got_a, got_b = g.get_apot(), g.download(bfld=True)["bfld"]
        assert (got_a == apot).mean() > 0.999 and np.abs(got_a - apot).max() <= 1e-6 * apot.max()
        b2 = ((b[:, 0] * b[:, 0] + b[:, 1] * b[:, 1]) + b[:, 2] * b[:, 2]).astype(np.float64)
        want_norm = bnorm / np.sqrt(b2.max()) / np.sqrt(3.0)
        assert abs(norm - want_norm) <= 1e-6 * want_norm
        wb = (b.astype(np.float64) * want_norm).astype(np.float32)
        B2 = ((wb[:, 0] * wb[:, 0] + wb[:, 1] * wb[:, 1]) + wb[:, 2] * wb[:, 2]).astype(np.float64)
        over = B2 > 18e-6 ** 2
        wb[over] = (wb[over].astype(np.float64) * (18e-6 / np.sqrt(B2[over]))[:, None]).astype(np.float32)
        assert abs(capped - int(over.sum())) <= 2
        assert (capped > 0) == (bnorm > 40e-6)
        scale = np.abs(wb).max()
        assert np.abs(got_b - wb).max() <= 2e-6 * scale


def test_sort_with_long_runs_of_nearly_equal_keys():
    """The radix passes only cover the top key bits; runs that agree in them are ordered by the
    full 128-bit key (and equal keys by upload index) in k_fix_ties.  Clumps of particles a
    few parsec apart make such runs long and frequent."""
    from oracle import port
    w = workloads.make("merger_1e6", n_gas=2000)
    rng = np.random.default_rng(11)
    base = w.pos[:2000]
    clumps = [base]
    for _ in range(5):
        clumps.append((base + rng.uniform(-3e-3, 3e-3, base.shape)).astype(np.float32))
    pos = np.clip(np.concatenate(clumps), 0, np.float32(w.boxsize)).astype(np.float32)
    pos[7000] = pos[123]                      # bit-identical positions: tie broken by index
    pos[9000] = pos[123]
    n = len(pos)
    g = tc.HotPath(n, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table())
    g.upload(pos)
    perm = g.sort()
    want, hi, lo, dup = port.sort(pos, w.boxsize)
    assert dup >= 2
    assert np.array_equal(perm, want)
    assert np.array_equal(g.download()["pos"], pos[want])


def test_sort_with_thousands_of_particles_in_one_sort_cell():
    """Runs longer than the serial fix-up's cap (RS_TIE_CAP = 256): 3000 particles inside one
    2^-16 Boxsize cell -- distinct floats, plus 700 exact duplicates of one of them -- are ordered
    by a whole block (k_fix_long_ties) like every other run: full 128-bit key, then index."""
    from oracle import port
    w = workloads.make("merger_1e6", n_gas=6000)
    rng = np.random.default_rng(12)
    box = np.float32(w.boxsize)
    cell = w.boxsize / 65536.0
    corner = np.array([40000, 50001, 33333], np.float64) * cell
    clump = (corner + rng.uniform(0.02, 0.98, (3000, 3)) * cell).astype(np.float32)
    dup = np.repeat(clump[17:18], 700, axis=0)
    pos = np.concatenate([w.pos[:6000], clump, dup]).astype(np.float32)
    pos = pos[rng.permutation(len(pos))]
    n = len(pos)
    assert len(np.unique(clump, axis=0)) > 2500          # the cell really holds distinct positions
    g = tc.HotPath(n, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table())
    g.upload(pos)
    perm = g.sort()
    want, hi, lo, ndup = port.sort(pos, w.boxsize)
    assert ndup >= 700
    assert np.array_equal(perm, want)
    assert np.array_equal(g.download()["pos"], pos[want])


def test_records_device_path_is_idempotent_and_pinned():
    """tg_upload / tg_download keep the driver's records resident on the device: a second
    download without an upload in between returns the same bytes (the records on the device are
    already in the new order), with and without page-locked host arrays."""
    w = workloads.make("merger_1e6", n_gas=30_001)
    n = w.n_gas
    Pdt = np.dtype([("Pos", "3f4"), ("Vel", "3f4"), ("ID", "i4"), ("Type", "i4"),
                    ("Key", "2u8"), ("Tree_Parent", "i4"), ("pad", "3i4")])
    Sdt = np.dtype([("U", "f4"), ("Rho", "f4"), ("Hsml", "f4"), ("VarHsmlFac", "f4"),
                    ("Bfld", "3f4"), ("Apot", "3f4"), ("ID", "f4"), ("Rho_Model", "f4"), ("Rs", "3f4")])
    res = []
    for pin in (False, True):
        P, S = np.zeros(n, Pdt), np.zeros(n, Sdt)
        P["Pos"], P["ID"], S["U"] = w.pos, np.arange(n) * 3 + 1, np.arange(n) * 0.25
        g = tc.HotPath.from_workload(w)
        if pin:
            g.pin_host(P)
            g.pin_host(S)
        g.upload_records(P, S)
        g.wvt_iteration(0.0085)
        g.wvt_iteration(0.0085)
        g.download_records(P, S)
        first = (P.tobytes(), S.tobytes())
        o = g.download()
        assert np.array_equal(o["id"], np.arange(n))        # records and state are in the same order now
        g.download_records(P, S)
        assert (P.tobytes(), S.tobytes()) == first
        assert np.array_equal(S["U"], (P["ID"] - 1) / 3 * np.float32(0.25))
        if pin:
            g.unpin_host(P)
            g.unpin_host(S)
        g.close()
        res.append(first)
    assert res[0] == res[1]


def test_gadget_blocks_from_the_soa_state():
    """tg_set_output_order / tg_fill_block (io.c:85-133): a block's write buffer is
    the field of particle order[k] at position k -- identity, a random file order, a bad index."""
    w = workloads.make("merger_1e6", n_gas=20_001)
    n = w.n_gas
    g = tc.HotPath.from_workload(w)
    g.upload(w.pos)
    g.find_sph_quantities()
    g.set_apot(np.repeat(np.linspace(0.1, 1, n, dtype=np.float32)[:, None], 3, 1) * [1, 0.5, 0.25])
    g.bfld_from_rotA_sph()
    o = g.download(bfld=True)
    names = dict(POS="pos", RHO="rho", HSML="hsml", BFLD="bfld", RHOM="rho_model")
    for label, key in names.items():
        assert np.array_equal(g.fill_block(label), o[key]), label
    order = np.random.default_rng(5).permutation(n)
    g.set_output_order(order)
    for label, key in names.items():
        assert np.array_equal(g.fill_block(label), o[key][order]), label
    g.set_output_order(None)
    assert np.array_equal(g.fill_block("RHO"), o["rho"])
    bad = order.copy()
    bad[7] = n
    with pytest.raises(tc.ToyGpuError):
        g.set_output_order(bad)
    # the failed call did not disturb the state: the sorted keys are still the records' keys
    assert np.array_equal(g.download()["rho"], o["rho"])
    # a file order refers to the device order it was given in: after another operator (a new sort)
    # it is stale, and the writer says so instead of gathering the wrong particles
    g.set_output_order(order)
    g.find_sph_quantities()
    with pytest.raises(tc.ToyGpuError, match="order changed"):
        g.fill_block("RHO")
    g.set_output_order(None)
    assert np.array_equal(g.fill_block("RHO"), g.download()["rho"])
