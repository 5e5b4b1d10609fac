"""Randomised parity sweep: seeds x sizes x workloads x (with / without particles snapped onto cell
planes), cold start + 3 WVT iterations in the reference's accumulation order, every array compared
bit for bit with the compiled reference.  Prints one line per case; exit code 1 on any mismatch."""
import sys, time, itertools
sys.path.insert(0, '.')
import numpy as np
import toycluster_b200 as tc
from toycluster_b200 import workloads
from oracle import ref

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
sizes = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (30011, 70000, 150003)
seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 1
t0 = time.time(); bad = 0; ncase = 0
cases = itertools.product(range(seed0, 100), ("merger_1e6", "single_1e5"), sizes, (0, 300), (0, 1))
for seed, name, n, snap, shift in cases:
    if time.time() - t0 > budget: break
    w = workloads.make(name, n_gas=n, seed=seed)
    if shift:      # periodic shift: the cluster straddles the box faces, every wrap path is taken
        off = np.random.default_rng(seed).uniform(0, w.boxsize, 3)
        w.pos = np.mod(w.pos.astype(np.float64) + off, w.boxsize).astype(np.float32)
        w.pos[w.pos >= np.float32(w.boxsize)] = 0
    if snap: w.pos = workloads.snap_to_cell_planes(w.pos, w.boxsize, snap, seed=seed)
    r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), 16)
    r.load(w.pos); after = []
    def cb(it):
        s = r.read()
        if it > 0:
            s["hw"], s["delta"] = r.wvt_scratch(); after.append(s)
        return 0
    niter = 3
    r.regularise(niter + 1, cb)
    log = ref.parse_log(r.log())
    g = tc.HotPath.from_workload(w, flags=tc.WVT_SEQUENTIAL); g.upload(w.pos)
    ok = True; hb = []
    for it in range(niter):
        g.wvt_iteration(log[it + 1]["step"])      # 0.0085 for the first iterations: exact in %g
        s, o = after[it], g.download(); hw, dl = g.wvt_scratch()
        hb.append(g.stats()["handed_back"])
        for k in ("id", "rho_model", "hsml", "rho", "varhsml", "pos"):
            if not np.array_equal(o[k], s[k]): ok = False; print("   MISMATCH", it, k, int((o[k] != s[k]).sum()))
        if not np.array_equal(hw, s["hw"]) or not np.array_equal(dl, s["delta"]):
            ok = False; print("   MISMATCH", it, "hw/delta", int((dl != s["delta"]).any(1).sum()))
    ncase += 1; bad += not ok
    print("%-11s n=%6d seed=%2d snap=%3d shift=%d displaced=%4d handed_back=%s %s" % (name, n, seed, snap, shift, g.stats()["displaced_particles"], hb, "ok" if ok else "FAIL"), flush=True)
print("cases", ncase, "failed", bad)
sys.exit(1 if bad else 0)
