"""Wall / device time of Make_magnetic_field and of the cold first step (developer script)."""
import sys, time
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
name = sys.argv[1] if len(sys.argv) > 1 else "merger_1e6"
w = workloads.make(name)
g = tc.HotPath.from_workload(w, flags=tc.FAST)
g.upload(w.pos)
for it in range(3):
    t0 = time.perf_counter(); g.wvt_iteration(0.0085); dt = time.perf_counter() - t0
    s = g.stats()
    print("step", it, "wall %.1f ms  step_ms %.1f index %.2f sweep %.1f tail %.2f" % (dt * 1e3, s["step_ms"], s["index_ms"], s["sweep_ms"], s["tail_ms"]), flush=True)
t0 = time.perf_counter(); g.find_sph_quantities(); dt = time.perf_counter() - t0
s = g.stats(); print("find_sph wall %.1f ms step_ms %.1f" % (dt * 1e3, s["step_ms"]))
for k in range(3):
    t0 = time.perf_counter()
    norm, capped = g.make_magnetic_field(20e-6, 0.5, r_sample_gas=[1e30] * len(w.halos))
    dt = time.perf_counter() - t0
    s = g.stats()
    print("make_magnetic_field call", k, "wall %.1f ms  device step_ms %.1f sweep_ms %.1f index(ev0-ev2) %.2f tail %.2f" % (dt * 1e3, s["step_ms"], s["sweep_ms"], s["index_ms"], s["tail_ms"]), flush=True)
