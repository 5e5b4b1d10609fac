"""Small TG_FAST run for compute-sanitizer: ragged n, power-of-two box (periodic wrap in the
tile kernels), cold + warm density, two WVT iterations, rot(A)."""
import sys
import numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
rng = np.random.default_rng(9)
n, box = 30003, 4096.0
pos = rng.random((n, 3)).astype(np.float32) * np.float32(box)
k = 0
for cx in (0.0, box):
    for cy in (0.0, box):
        for cz in (0.0, box):
            off = (rng.random((600, 3)) * 30).astype(np.float32)
            c = np.array([cx, cy, cz], np.float32)
            pos[k:k + 600] = np.where(c == 0, off, np.float32(box) - off)
            k += 600
halo = np.array([[0, 0, 0, 1e-6, 0.54, 300.0, 3000.0, 0, 1.0]])
g = tc.HotPath(n, box, 1.0, 1e5, halo, flags=tc.FAST)
g.upload(pos)
g.find_sph_quantities()
g.find_sph_quantities()
for _ in range(2):
    g.wvt_iteration(0.0085)
g.find_sph_quantities()
g.set_apot(np.ones((n, 3), np.float32) * rng.random((n, 1)).astype(np.float32))
g.bfld_from_rotA_sph()
o = g.download(bfld=True)
assert np.isfinite(o["rho"]).all() and np.isfinite(o["bfld"]).all()
print("sanitize_small ok", g.stats()["handed_back"])
