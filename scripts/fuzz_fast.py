"""Randomised sweep of TG_FAST against the exact default mode (both on the GPU, so no CPU
reference run is needed and many cases fit a minute): seeds x sizes x (particles snapped onto
cell planes -> displaced reference-tree nodes) x (periodic shift -> every wrap path).  Every warm
iteration restarts the fast context from the exact one's state and must stay inside the
distribution bound of tests/test_gpu_fast.py; the searches per particle must be identical.
usage: fuzz_fast.py [seconds] [sizes] [first seed]"""
import sys, time, itertools
sys.path.insert(0, '.')
import numpy as np
import toycluster_b200 as tc
from toycluster_b200 import workloads

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
sizes = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (60011, 150000, 400003)
seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t0 = time.time(); bad = 0; ncase = 0
def rel(a, b): return np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b.astype(np.float64)), 1e-300)
for seed, name, n, snap, shift in itertools.product(range(seed0, 200), ("merger_1e6", "single_1e5"), sizes, (0, 3000), (0, 1)):
    if time.time() - t0 > budget: break
    w = workloads.make(name, n_gas=n, seed=seed)
    if shift:
        off = np.random.default_rng(seed).uniform(0, w.boxsize, 3)
        w.pos = np.mod(w.pos.astype(np.float64) + off, w.boxsize).astype(np.float32)
        w.pos[w.pos >= np.float32(w.boxsize)] = 0
    e = tc.HotPath.from_workload(w); f = tc.HotPath.from_workload(w, flags=tc.FAST)
    e.upload(w.pos); e.wvt_iteration(0.0085)
    ok = True; worst = [0, 0, 0]; hb = []
    for it in range(3):
        s = e.download()
        pos = workloads.snap_to_cell_planes(s["pos"], w.boxsize, snap, seed=seed + it, levels=(4, 5, 6, 7, 8)) if snap else s["pos"]
        for g in (e, f):
            g.upload(pos, s["hsml"]); g.wvt_iteration(0.0085)
        a, b = e.download(), f.download()
        _, da = e.wvt_scratch(); _, db = f.wvt_scratch()
        se, sf = e.stats(), f.stats()
        hb.append(sf["handed_back"])
        if not (np.array_equal(a["id"], b["id"]) and np.array_equal(a["rho_model"], b["rho_model"])): ok = False; print("   order / rho_model differ", it)
        if se["searches"] != sf["searches"]: ok = False; print("   searches differ", it, se["searches"], sf["searches"])
        for k in ("hsml", "rho", "varhsml"):
            r = rel(b[k], a[k]); worst[0] = max(worst[0], r.max())
            if (r <= 1e-5).mean() < 0.999 or r.max() > 5e-4: ok = False; print("   %s: within 1e-5 %.5f max %.2e" % (k, (r <= 1e-5).mean(), r.max()), it)
        sc = np.maximum(np.linalg.norm(da, axis=1), 1e-30)
        err = np.linalg.norm(db.astype(np.float64) - da, axis=1) / sc
        q99, q999 = np.quantile(err, 0.99), np.quantile(err, 0.999); worst[1] = max(worst[1], q99); worst[2] = max(worst[2], q999)
        if q99 > 1.5e-5 or q999 > 3e-5: ok = False; print("   displacement q99 %.2e q99.9 %.2e" % (q99, q999), it)
    ncase += 1; bad += not ok
    print("%-11s n=%6d seed=%3d snap=%4d shift=%d displaced=%5d handed_back=%s max rel %.1e  delta q99 %.1e q99.9 %.1e %s"
          % (name, n, seed, snap, shift, f.stats()["displaced_particles"], hb, worst[0], worst[1], worst[2], "ok" if ok else "FAIL"), flush=True)
    e.close(); f.close()
print("cases", ncase, "failed", bad)
sys.exit(1 if bad else 0)
