"""Third randomised sweep: the WHOLE Regularise_sph_particles loop (library's own control flow,
tg_regularise) to the reference's termination, sequential mode: iteration count, every printed
number (as printed) and the final state bit for bit."""
import sys, time, itertools
sys.path.insert(0, '.')
import numpy as np
import toycluster_b200 as tc
from toycluster_b200 import workloads
from oracle import ref

def same_as_printed(v, p):
    return float("%g" % v) == p or abs(v - p) <= 1.01e-5 * abs(p)

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 200.0
t0 = time.time(); bad = 0; ncase = 0
for seed, name, n, shift in itertools.product(range(41, 100), ("merger_1e6", "single_1e5"), (9000, 30011, 60000), (0, 1)):
    if time.time() - t0 > budget: break
    w = workloads.make(name, n_gas=n, seed=seed)
    if shift:
        off = np.random.default_rng(seed).uniform(0, w.boxsize, 3)
        w.pos = np.mod(w.pos.astype(np.float64) + off, w.boxsize).astype(np.float32)
        w.pos[w.pos >= np.float32(w.boxsize)] = np.float32(w.boxsize)
    r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), 16)
    r.load(w.pos); r.regularise(); log = ref.parse_log(r.log()); r.find_sph_quantities(); want = r.read()
    g = tc.HotPath.from_workload(w, flags=tc.WVT_SEQUENTIAL); g.upload(w.pos)
    done, rows = g.regularise_sph_particles(); g.find_sph_quantities(); got = g.download()
    ok = done == len(log)
    for a, b in zip(rows, log):
        for k in ("max", "mean", "diff", "step"):
            if not same_as_printed(a[k], b[k]): ok = False; print("   LOG", a["it"], k, a[k], b[k])
    for k in ("id", "pos", "hsml", "rho", "varhsml"):
        if not np.array_equal(got[k], want[k]): ok = False; print("   STATE", k, int((got[k] != want[k]).sum()))
    ncase += 1; bad += not ok
    print("%-11s n=%6d seed=%2d shift=%d iterations %2d/%2d %s" % (name, n, seed, shift, done, len(log), "ok" if ok else "FAIL"), flush=True)
print("cases", ncase, "failed", bad)
sys.exit(1 if bad else 0)
