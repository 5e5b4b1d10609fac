"""Why the tile sweep hands targets back to the generic sweep (tg_stats.handback_why)."""
import sys
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
name = sys.argv[1] if len(sys.argv) > 1 else "merger_1e7"
w = workloads.make(name)
g = tc.HotPath.from_workload(w, flags=tc.FAST)
g.upload(w.pos)
for it in range(4):
    g.wvt_iteration(0.0085)
    s = g.stats()
    print(it, "step_ms %.1f sweep_ms %.1f handed_back %d why %s displaced nodes %d particles %d" % (
        s["step_ms"], s["sweep_ms"], s["handed_back"], s["handback_why"], s["displaced_nodes"], s["displaced_particles"]))
