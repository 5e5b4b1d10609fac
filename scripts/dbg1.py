import numpy as np, sys
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
from oracle import ref
w = workloads.make("merger_1e6", n_gas=20000)
r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), 8)
r.load(w.pos)
snaps=[]
def cb(it):
    if it>0:
        h,d = r.wvt_scratch(); s=r.read(); s["hw"]=h; s["delta"]=d; snaps.append(s)
    return 0
r.regularise(6, cb)
log = ref.parse_log(r.log())
g = tc.HotPath.from_workload(w, flags=1)
g.upload(w.pos)
for it in range(6):
    g.wvt_iteration(log[it]["step"])
    o = g.download(); hw, dl = g.wvt_scratch(); s = snaps[it]
    ratio = hw.astype(np.float64)/s["hw"]
    print(it, "ids", np.array_equal(o["id"], s["id"]), "hw eq", (hw==s["hw"]).mean(), "ratio", ratio.min(), ratio.max(),
          "rhom eq", (o["rho_model"]==s["rho_model"]).mean(), "pos eq", (o["pos"]==s["pos"]).all(1).mean(),
          "delta eq", (dl==s["delta"]).all(1).mean(), "rho eq", (o["rho"]==s["rho"]).mean(), "hsml eq", (o["hsml"]==s["hsml"]).mean())
    bad = np.flatnonzero(o["rho_model"]!=s["rho_model"])[:5]
    print("   bad rhom", bad, o["rho_model"][bad], s["rho_model"][bad])

    badd = np.flatnonzero((dl!=s["delta"]).any(1))[:5]
    for b in badd:
        print("   bad delta", b, dl[b], s["delta"][b], "hw", hw[b], "pos", o["pos"][b], len(g.find_ngb(b, hw[b]*w.boxsize)) if False else "")
