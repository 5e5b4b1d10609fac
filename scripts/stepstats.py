import sys
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
w = workloads.make(sys.argv[1] if len(sys.argv) > 1 else "merger_1e7")
g = tc.HotPath.from_workload(w); g.upload(w.pos); n = w.n_gas
for it in range(6):
    g.wvt_iteration(0.0085); s = g.stats()
    print(it, "step %.1f sweep %.1f | per particle: evals %.0f gathered %.0f searches %.3f iters %.3f" % (
        s["step_ms"], s["sweep_ms"], s["pair_evals"]/n, s["gathered"]/n, s["searches"]/n, s["hsml_iters"]/n))
