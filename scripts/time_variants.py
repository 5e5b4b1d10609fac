"""Time A/B builds (scripts/variants.sh) on the 10 M merger (or TOYGPU_BENCH_WORKLOAD): one process per variant."""
import os, subprocess, sys, glob
here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(here)
code = '''
import sys; sys.path.insert(0, %r)
import toycluster_b200 as tc
from toycluster_b200 import workloads
import numpy as np
w = workloads.make(%r)
g = tc.HotPath.from_workload(w, flags=int(%r)); g.upload(w.pos)
ms = []
for it in range(7):
    g.wvt_iteration(0.0085); s = g.stats(); ms.append((s["step_ms"], s["sweep_ms"]))
o = g.download()
n = w.n_gas
print("step %%.2f sweep %%.2f  checksum %%.6f  searches/p %%.3f iters/p %%.3f evals/p %%.1f gathered/p %%.1f handed_back %%d why %%s" %% (np.mean([m[0] for m in ms[3:]]), np.mean([m[1] for m in ms[3:]]), float(o["pos"].astype(np.float64).sum() + o["hsml"].astype(np.float64).sum()), s["searches"] / n, s["hsml_iters"] / n, s["pair_evals"] / n, s["gathered"] / n, s["handed_back"], s["handback_why"]))
''' % (root, os.environ.get('TOYGPU_BENCH_WORKLOAD', 'merger_1e7'), os.environ.get('TOYGPU_BENCH_FLAGS', '4'))
libs = sorted(glob.glob(os.path.join(root, "toycluster_b200", "variants", "*%s*.so" % os.environ.get("TOYGPU_VARIANTS", ""))))
for lib in [os.path.join(root, "toycluster_b200", "libtoygpu.so")] + libs:
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, TOYGPU_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip(), r.stderr.strip()[-300:])
