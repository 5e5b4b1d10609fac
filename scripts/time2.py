import sys, time, numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
n = int(sys.argv[1])
w = workloads.make("merger_1e6", n_gas=n)
g = tc.HotPath.from_workload(w); g.upload(w.pos)
for it in range(4): g.wvt_iteration(0.0085)
s = g.stats(); print("fused   sweep %.2f ms step %.2f" % (s["sweep_ms"], s["step_ms"]))
for it in range(3):
    g.find_sph_quantities(); s = g.stats()
    print("density sweep %.2f ms step %.2f evals/part %.0f back %d" % (s["sweep_ms"], s["step_ms"], s["pair_evals"]/n, s["handed_back"]))
