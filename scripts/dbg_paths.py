import os, sys, numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
n = int(sys.argv[1])
w = workloads.make("merger_1e6", n_gas=n)
g1 = tc.HotPath.from_workload(w)
os.environ["TOYGPU_NO_TILES"] = "1"
g2 = tc.HotPath.from_workload(w)
del os.environ["TOYGPU_NO_TILES"]
for g in (g1, g2): g.upload(w.pos)
for it in range(5):
    for g in (g1, g2): g.wvt_iteration(0.0085)
    a, b = g1.download(), g2.download()
    # keep both trajectories identical: copy generic state into tile ctx
    bad = np.flatnonzero((a["rho"] != b["rho"]) | (a["hsml"] != b["hsml"]) | (a["varhsml"] != b["varhsml"]))
    print(it, "back", g1.stats()["handed_back"], "n differing", len(bad), "pos eq", (a["pos"]==b["pos"]).all(1).mean())
    for k in bad[:5]:
        print("   ", k, a["hsml"][k], b["hsml"][k], a["rho"][k], b["rho"][k], a["varhsml"][k], b["varhsml"][k])
    g1.upload(b["pos"], b["hsml"])
    g2.upload(b["pos"], b["hsml"])
