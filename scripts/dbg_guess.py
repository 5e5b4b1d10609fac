import sys, numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
n = 1_000_000
w = workloads.make("merger_1e6", n_gas=n)
want = np.load("scripts/tmp/guess_ref_1e6.npy")
g = tc.HotPath.from_workload(w); g.upload(w.pos); g.sort()
got = g.guess_hsml()
bad = np.flatnonzero(got != want)
print("gpu vs ref guess mismatches", len(bad))
print(bad[:20]); print(got[bad[:20]]); print(want[bad[:20]])
# context
for b in bad[:6]:
    print(b, "ratio^3", (got[b]/want[b])**3)
