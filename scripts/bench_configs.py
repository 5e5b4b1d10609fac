"""BASELINE.json configs[0..4] through libtoygpu in TG_FAST: steady-state step time, the cold
first step, and (config 4) Make_magnetic_field with the rot(A) sweep.  -> gpurun_out/configs.json
usage: python scripts/bench_configs.py [tag]"""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
from toycluster_b200 import workloads

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
out = {}
for cfg, name, n in (("configs[0] single halo 1e5", "single_1e5", None), ("configs[1] merger 1e6", "merger_1e6", None),
                     ("configs[2] merger 1e7", "merger_1e7", None), ("configs[4] merger + substructure 1e7", "merger_sub_1e7", None)):
    w = workloads.make(name) if n is None else workloads.make(name, n_gas=n)
    g = tc.HotPath.from_workload(w, flags=tc.FAST)
    g.upload(w.pos)
    rows = []
    for it in range(8):
        g.wvt_iteration(0.0085)
        rows.append(g.stats())
    warm = rows[3:]
    n = w.n_gas
    step_ms = float(np.mean([r["step_ms"] for r in warm])); sweep_ms = float(np.mean([r["sweep_ms"] for r in warm]))
    gath = float(np.mean([r["gathered"] for r in warm]))
    row = {"workload": name, "n_gas": n, "halo_rows": len(w.halos), "cold_step_ms": rows[0]["step_ms"],
           "step_ms": step_ms, "sweep_ms": sweep_ms, "steps_per_s": 1e3 / step_ms,
           "interactions_per_s": float(np.mean([r["pair_evals"] for r in warm])) / step_ms * 1e3,
           "sweep_roofline_frac": (40.0 * n + 16.0 * gath) / (sweep_ms * 1e-3) / 1e9 / PEAK,
           "handed_back": int(warm[-1]["handed_back"])}
    if name in ("merger_1e6", "merger_1e7"):       # configs[3]: + Bonafede magnetic field (SPH rot A on the GPU)
        g.find_sph_quantities()
        # (first call: allocates the vector-potential arrays -- cudaMalloc time, not the operator's)
        t0 = time.perf_counter()
        g.make_magnetic_field(20e-6, 0.5, r_sample_gas=[1e30] * len(w.halos))
        dt_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        norm, capped = g.make_magnetic_field(20e-6, 0.5, r_sample_gas=[1e30] * len(w.halos))
        dt = time.perf_counter() - t0
        s = g.stats()
        row["configs[3] magnetic field"] = {"make_magnetic_field_ms": dt * 1e3, "first_call_ms": dt_first * 1e3,
                                           "device_ms": s["step_ms"], "rotA_sweep_ms": s["sweep_ms"],
                                           "rotA_pairs": int(s["pair_evals"]), "rotA_handed_back": int(s["handed_back"]),
                                           "rotA_roofline_frac": (28.0 * n + 28.0 * s["gathered"]) / (s["sweep_ms"] * 1e-3) / 1e9 / PEAK,
                                           "bfld_norm": norm, "capped": capped}
    out[cfg] = row
    print(cfg, json.dumps(row), flush=True)
    g.close()
json.dump(out, open("gpurun_out/configs%s.json" % (("_" + sys.argv[1]) if len(sys.argv) > 1 else ""), "w"), indent=1)
