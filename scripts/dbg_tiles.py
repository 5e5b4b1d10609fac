import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
n = int(sys.argv[1])
w = workloads.make("merger_1e6", n_gas=n)
g = tc.HotPath.from_workload(w); g.upload(w.pos)
for it in range(3): g.wvt_iteration(0.0085)
nt = C.c_int(); g.lib.tg_debug_tile_counts(g._ctx, None, C.byref(nt))
out = np.empty(nt.value, np.int32); g.lib.tg_debug_tile_counts(g._ctx, out.ctypes.data_as(C.c_void_p), C.byref(nt))
ng = np.where(out < 0, -out, (out >> 12) & 0xffff)
print("tiles", nt.value, "quantiles runs:", np.quantile(ng, [0.01,0.1,0.25,0.5,0.75,0.9,0.99]), "max", ng.max(), "frac<=384", (ng<=384).mean(), "<=512", (ng<=512).mean(), "<=640", (ng<=640).mean(), "back", (out<0).mean())
