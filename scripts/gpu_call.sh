python scripts/dbg_bfield_time.py merger_1e6 2>&1 | grep "step 0"
TOYGPU_LIB=$PWD/toycluster_b200/variants/libtoygpu_0old.so python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads
w = workloads.make("merger_1e6")
g = tc.HotPath.from_workload(w, flags=tc.FAST)
g.upload(w.pos)
g.wvt_iteration(0.0085); s = g.stats()
print("OLD step 0 step_ms %.1f sweep %.1f" % (s["step_ms"], s["sweep_ms"]))
PY
