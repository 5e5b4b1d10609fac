"""Full-size parity run (BASELINE configs[1]: 1 M gas merger): the reference's own code on the
host cores vs libtoygpu on the GPU, iteration by iteration.  Too slow for the test suite
(minutes of CPU); its output is kept under profiles/."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
from oracle import ref

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = workloads.make("merger_1e7" if n > 2_000_000 else "merger_1e6", n_gas=n)
r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), 0)
r.load(w.pos)
snaps, stamps = [], []
def cb(it):
    stamps.append(time.perf_counter())
    s = r.read()
    if it > 0:
        s["hw"], s["delta"] = r.wvt_scratch()
    snaps.append(s)
    return 0
t0 = time.perf_counter(); r.regularise(niter, cb); t_ref = time.perf_counter() - t0
log = ref.parse_log(r.log())
niter = min(niter, len(snaps) - 1)          # the reference may terminate by itself (wvt_relax.c:95-98)
import math
def exact_step(printed, base=0.0085):
    """The log prints step with %g (6 digits); the value itself is base * 0.8 * 0.8 * ... in
    double (wvt_relax.c:100)."""
    m = round(math.log(printed / base) / math.log(0.8))
    v = base
    for _ in range(m):
        v *= 0.8
    assert abs(v - printed) <= 1e-5 * printed
    return v
for row in log:
    row["step"] = exact_step(row["step"])
out = {"n_gas": n, "iterations": niter, "reference_threads": r.nthreads,
       "reference_s_per_iteration": list(np.diff(stamps))}
modes = (("sequential", tc.WVT_SEQUENTIAL), ("default", 0)) if len(sys.argv) < 4 else ((sys.argv[3], tc.WVT_SEQUENTIAL if sys.argv[3] == "sequential" else 0),)
for mode, flags in modes:
    g = tc.HotPath.from_workload(w, flags=flags)
    g.upload(w.pos)
    rows = []
    for it in range(niter):
        step = log[it + 1]["step"] if it + 1 < len(log) else log[it]["step"]
        if mode == "default" and it > 0:       # per-iteration parity: restart from the reference's state
            g.upload(snaps[it]["pos"], snaps[it]["hsml"])
        t1 = time.perf_counter(); g.wvt_iteration(step); dt = time.perf_counter() - t1
        s, o = snaps[it + 1], g.download()
        hw, dl = g.wvt_scratch()
        ids = o["id"] if (mode == "sequential" or it == 0) else snaps[it]["id"][o["id"]]
        row = {"it": it, "gpu_ms": dt * 1e3, "handed_back": g.stats()["handed_back"],
               "order_equal": bool(np.array_equal(ids, s["id"]))}
        for k in ("hsml", "rho", "varhsml", "rho_model", "pos"):
            row[k + "_bit_equal_frac"] = float((o[k] == s[k]).mean())
            if k != "pos":
                rel = np.abs(o[k].astype(np.float64) - s[k]) / np.abs(s[k])
                row[k + "_max_rel"] = float(rel.max())
        row["hw_bit_equal_frac"] = float((hw == s["hw"]).mean())
        sc = np.linalg.norm(s["delta"], axis=1)
        err = np.linalg.norm(dl.astype(np.float64) - s["delta"], axis=1) / np.maximum(sc, 1e-30)
        row["delta_bit_equal_frac"] = float((dl == s["delta"]).all(1).mean())
        row["delta_rel_q50_q99_q999_max"] = [float(v) for v in np.quantile(err, [0.5, 0.99, 0.999, 1.0])]
        rows.append(row)
        print(mode, json.dumps(row))
    out[mode] = rows
json.dump(out, open("gpurun_out/parity_full_size.json", "w"), indent=1)
print("reference total %.1f s" % t_ref)
