"""Partition invariance: R rank contexts on one GPU (host copies stand in for the all-gather)
against the one-rank run, bit for bit, over seeds x sizes x rank counts x periodic shifts."""
import sys, time, itertools
sys.path.insert(0, '.')
import numpy as np
import toycluster_b200 as tc
from toycluster_b200 import workloads
from toycluster_b200.dist import rank_slice
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
t0 = time.time(); bad = 0; ncase = 0
for seed, n, R, shift in itertools.product(range(61, 100), (20011, 64000), (2, 3, 8), (0, 1)):
    if time.time() - t0 > budget: break
    w = workloads.make("merger_1e6", n_gas=n, seed=seed)
    if shift:
        off = np.random.default_rng(seed).uniform(0, w.boxsize, 3)
        w.pos = np.mod(w.pos.astype(np.float64) + off, w.boxsize).astype(np.float32)
        w.pos[w.pos >= np.float32(w.boxsize)] = np.float32(w.boxsize)
    one = tc.HotPath.from_workload(w); one.upload(w.pos)
    parts = [tc.HotPath.from_workload(w, rank=r, nranks=R) for r in range(R)]
    state_pos, state_h, prev_ids, ok = w.pos, None, None, True
    for it in range(3):
        one.wvt_iteration(0.0085); ref_ = one.download(); merged = {}
        for r, g in enumerate(parts):
            g.upload(state_pos, state_h); g.wvt_iteration(0.0085); o = g.download()
            lo, hi, _ = rank_slice(n, r, R)
            for k in ("pos", "hsml", "rho", "varhsml", "id"):
                merged.setdefault(k, np.empty_like(o[k]))[lo:hi] = o[k][lo:hi]
        ids = merged["id"] if it == 0 else prev_ids[merged["id"]]
        for k in ("pos", "hsml", "rho", "varhsml"):
            if not np.array_equal(merged[k], ref_[k]): ok = False; print("   MISMATCH", it, k, int((merged[k] != ref_[k]).sum()))
        if not np.array_equal(ids, ref_["id"]): ok = False; print("   MISMATCH ids", it)
        prev_ids, state_pos, state_h = ids, merged["pos"], merged["hsml"]
    ncase += 1; bad += not ok
    print("n=%6d seed=%2d ranks=%d shift=%d %s" % (n, seed, R, shift, "ok" if ok else "FAIL"), flush=True)
print("cases", ncase, "failed", bad)
sys.exit(1 if bad else 0)
