import sys, time, numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
name, n = sys.argv[1], int(sys.argv[2])
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t=time.time(); w = workloads.make(name, n_gas=n); print("gen", time.time()-t)
g = tc.HotPath.from_workload(w, flags=flags)
t=time.time(); g.upload(w.pos); print("upload", time.time()-t)
step = 0.0085
for it in range(int(sys.argv[4]) if len(sys.argv) > 4 else 6):
    t=time.time(); emax, emean = g.wvt_iteration(step); dt=time.time()-t
    s = g.stats()
    print(it, "wall %.1f ms step %.1f ms sweep %.1f ms  evals/part %.0f gath/part %.0f searches/part %.2f iters/part %.2f err %.4g %.4g back %d" % (
        dt*1e3, s["step_ms"], s["sweep_ms"], s["pair_evals"]/n, s["gathered"]/n, s["searches"]/n, s["hsml_iters"]/n, emax, emean, s["handed_back"]))
