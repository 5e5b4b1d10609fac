"""Synthetic cluster.par workloads: derived halo tables match SURVEY.md section 8(d)."""
import numpy as np

from toycluster_b200 import workloads


def test_single_halo_scalars():
    w = workloads.make("single_1e5", with_positions=False)
    assert w.boxsize == 13923.0 and w.n_gas == 100_000
    h = w.halos[0]
    assert abs(h.r200 - 1856.5) < 0.5 and abs(h.rcore - 255.4) < 0.1 and abs(h.rcut - 2599.0) < 0.5
    assert abs(h.rho0 / 7.442e-6 - 1) < 2e-3
    assert abs(w.mpart_gas / 0.3175 - 1) < 2e-3
    assert w.mtotal > 1e5          # wvt_relax.c:53: the step is not halved


def test_merger_scalars_and_positions():
    w = workloads.make("merger_1e6", n_gas=20000)
    assert w.boxsize == 12716.0
    a, b = w.halos
    assert abs(a.dcom[0] + 609.9) < 0.5 and abs(b.dcom[0] - 1951.7) < 0.5
    assert abs(a.rho0 / 7.736e-6 - 1) < 2e-3 and abs(b.rho0 / 9.143e-6 - 1) < 2e-3
    assert a.npart_gas + b.npart_gas == 20000
    assert w.pos.dtype == np.float32 and w.pos.shape == (20000, 3)
    assert w.pos.min() >= 0 and w.pos.max() <= w.boxsize
    # seeded: the same call gives the same particles
    assert np.array_equal(w.pos, workloads.make("merger_1e6", n_gas=20000).pos)
    # the denser halo centre holds more particles than a box corner
    c0 = np.array(a.dcom) + w.boxsize / 2
    near = (np.linalg.norm(w.pos - c0, axis=1) < 300).sum()
    corner = (np.linalg.norm(w.pos, axis=1) < 300).sum()
    assert near > 50 * max(corner, 1)
