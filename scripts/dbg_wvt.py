import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import toycluster_b200 as tc
from toycluster_b200 import workloads
import test_gpu_parity as T
w = workloads.make("merger_1e6", n_gas=int(sys.argv[1]))
niter = 4
start, after, steps, log = T._reference_iterations(w, niter)
g = tc.HotPath.from_workload(w)
for it in range(niter):
    st = start[it]
    g.upload(st["pos"], st["hsml"] if it > 0 else None)
    g.wvt_iteration(steps[it])
    s, o = after[it], g.download()
    hw, dl = g.wvt_scratch()
    scale = np.linalg.norm(s["delta"], axis=1)
    err = np.linalg.norm(dl.astype(np.float64) - s["delta"], axis=1) / np.maximum(scale, 1e-30)
    print(it, "back", g.stats()["handed_back"], "q50 %.2e q99 %.2e q999 %.2e max %.2e" % tuple(np.quantile(err, [0.5, 0.99, 0.999, 1.0])),
          "pos equal %.4f" % (o["pos"] == s["pos"]).all(1).mean(), "rho eq", (o["rho"]==s["rho"]).mean())
