"""Histogram of the tile walk's verdicts (tile_ng codes) after a few warm steps: why whole tiles
go to the generic sweep, and how many candidate runs / boxes the others carry (developer script)."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
name = sys.argv[1] if len(sys.argv) > 1 else "merger_1e7"
w = workloads.make(name)
g = tc.HotPath.from_workload(w, flags=tc.FAST)
g.upload(w.pos)
for it in range(4):
    g.wvt_iteration(0.0085)
nt = C.c_int()
g.lib.tg_debug_tile_counts.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
g.lib.tg_debug_tile_counts(g._ctx, None, C.byref(nt))
codes = np.empty(nt.value, np.int32)
g.lib.tg_debug_tile_counts(g._ctx, codes.ctypes.data_as(C.c_void_p), C.byref(nt))
o = g.download()
neg = codes[codes < 0]
print("tiles", nt.value, "handed back whole", len(neg), " cold / R >= 0.49 box (-1):", (neg == -1).sum(),
      " queue overflow (-2):", (neg == -2).sum(), " too many runs / boxes:", (neg < -2).sum())
if (neg < -2).any():
    print("  runs of those (quantiles 0.1 0.5 0.9 max):", np.quantile(-neg[neg < -2], [0.1, 0.5, 0.9, 1.0]))
ok = codes[codes >= 0]
nent, nruns = ok & 0xfff, (ok >> 12) & 0xffff
print("tiled: boxes per tile q50 %.0f q99 %.0f max %d; runs per tile q50 %.0f q90 %.0f q99 %.0f max %d; interior %.3f"
      % (np.median(nent), np.quantile(nent, .99), nent.max(), np.median(nruns), np.quantile(nruns, .9), np.quantile(nruns, .99), nruns.max(), ((ok >> 30) & 1).mean()))
# where the handed-back tiles are: radius from the box centre, Hsml
idx = np.nonzero(codes < 0)[0]
if len(idx):
    i0 = np.minimum(idx * 32, w.n_gas - 1)
    print("  Hsml / Boxsize of their first target: q10 %.4f q50 %.4f q90 %.4f   (all particles: q50 %.4f q99 %.4f)" %
          (tuple(np.quantile(o["hsml"][i0], [.1, .5, .9]) / w.boxsize) + tuple(np.quantile(o["hsml"], [.5, .99]) / w.boxsize)))
