"""Whole-program wall clock: the reference driver with its own hot path vs with gpu_shim +
libtoygpu (oracle/_ref/Toycluster_cpu / _gpu_b), same parameter file (BASELINE configs[1]:
two-cluster merger, 1 M gas + 1 M DM), all host threads.  Output -> profiles/."""
import json, os, subprocess, sys, time
here = os.path.dirname(os.path.abspath(__file__)); root = os.path.dirname(here)
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, "tests"))
from test_driver_e2e import PAR, read_gadget2
import numpy as np
ntotal = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
mass_ratio = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3125
sequential = len(sys.argv) > 3 and sys.argv[3] == "sequential"    # TG_WVT_SEQUENTIAL: blocks must be byte-identical
work = "/tmp/whole"; os.makedirs(work, exist_ok=True)
out = {"ntotal": ntotal, "threads": os.cpu_count()}
out["mass_ratio"], out["sequential"] = mass_ratio, sequential
for tag, exe, env in (("gpu", "Toycluster_gpu_b", {"TOYGPU_FLAGS": "1"} if sequential else {}), ("cpu", "Toycluster_cpu", {})):
    open(f"{work}/{tag}.par", "w").write(PAR.format(out=f"IC_{tag}", ntotal=ntotal, mass_ratio=mass_ratio, bnorm="20e-6"))
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(root, "oracle", "_ref", exe), f"{tag}.par"], cwd=work,
                       env=dict(os.environ, **env), capture_output=True, text=True)
    dt = time.perf_counter() - t0
    its = [l.strip() for l in r.stdout.splitlines() if l.lstrip().startswith("#")]
    out[tag] = {"seconds": dt, "rc": r.returncode, "iterations": len(its), "first": its[:2], "last": its[-1:]}
    print(tag, "%.1f s" % dt, len(its), "iterations", its[-1:], flush=True)
c, g = read_gadget2(f"{work}/IC_cpu"), read_gadget2(f"{work}/IC_gpu")
for lab in ("RHO ", "HSML"):
    a, b = np.frombuffer(c[lab], np.float32), np.frombuffer(g[lab], np.float32)
    out["median_" + lab.strip()] = [float(np.median(a)), float(np.median(b))]
out["blocks_byte_identical"] = {lab.strip(): bool(c[lab] == g[lab]) for lab in c}   # VEL: the reference's own
# velocity sampling is not reproducible run to run with more than one thread
bc, bg = np.frombuffer(c["BFLD"], np.float32).reshape(-1, 3), np.frombuffer(g["BFLD"], np.float32).reshape(-1, 3)
out["bfld_max_rel_of_scale"] = float((np.abs(bg - bc) / (np.abs(bc).max(axis=1, keepdims=True) + 1e-30)).max())
out["speedup"] = out["cpu"]["seconds"] / out["gpu"]["seconds"]
json.dump(out, open(os.path.join(root, "gpurun_out", "whole_program_%d.json" % ntotal), "w"), indent=1)
print(json.dumps(out)[-700:])
