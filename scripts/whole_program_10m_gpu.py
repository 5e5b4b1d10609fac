import os, subprocess, sys, time
root = "/root/repo"; sys.path.insert(0, os.path.join(root, "tests"))
from test_driver_e2e import PAR
work = "/tmp/whole10"; os.makedirs(work, exist_ok=True)
open(f"{work}/g.par", "w").write(PAR.format(out="IC_g", ntotal=20_000_000, mass_ratio=0.3125, bnorm="20e-6"))
t0 = time.perf_counter()
p = subprocess.Popen([os.path.join(root, "oracle/_ref/Toycluster_gpu_b"), "g.par"], cwd=work, stdout=subprocess.PIPE, text=True)
for line in p.stdout:
    s = line.strip()
    if s and (s.startswith("#") or s.startswith("Starting") or s.startswith("done") or "Output" in s or "Setting" in s or "Magnetic" in s or "Sampling" in s or "Bfld" in s):
        print("%7.2f s | %s" % (time.perf_counter() - t0, s[:90]), flush=True)
p.wait(); print("total %.1f s rc %d" % (time.perf_counter() - t0, p.returncode))
