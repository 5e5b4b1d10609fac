import sys, numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
from oracle import ref, port
n = int(sys.argv[1])
w = workloads.make("merger_1e6", n_gas=n)
r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), 0)
r.load(w.pos); r.find_sph_quantities(); d = r.read()
import os
if len(sys.argv) > 2: os.environ["TOYGPU_NO_TILES"] = "1"
g = tc.HotPath.from_workload(w); g.upload(w.pos); g.find_sph_quantities(); o = g.download()
bad = np.flatnonzero(o["hsml"] != d["hsml"])
print("n", n, "mismatch", len(bad), "ids equal", np.array_equal(o["id"], d["id"]))
guess = g.guess_hsml()
for b in bad[:12]:
    c_ref = len(port.find_ngb(d["pos"], w.boxsize, b, d["hsml"][b])); c_gpu = len(port.find_ngb(d["pos"], w.boxsize, b, o["hsml"][b]))
    print(b, "ref h %.6g rho %.6g | gpu h %.6g rho %.6g | ratio %.5f guess %.6g cnt(ref h) %d cnt(gpu h) %d pos %s" % (d["hsml"][b], d["rho"][b], o["hsml"][b], o["rho"][b], o["hsml"][b]/d["hsml"][b], guess[b], c_ref, c_gpu, d["pos"][b]))
np.save("gpurun_out/cold_bad.npy", bad)
