"""Generate the golden fixtures from the reference's OWN code (oracle/_ref = the unmodified
hot-path sources compiled by oracle/Makefile).  Run here, where /root/reference exists:

    python tests/golden/make_golden.py

peano_table.json      known-answer Peano keys (the table of SURVEY.md 8c and more)
merger_4096.npz       a 4096-particle two-cluster merger followed through Guess_hsml, neighbour
                      lists, the cold density pass, three WVT iterations (every scratch array)
                      and rot(A)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref  # noqa: E402
from toycluster_b200 import workloads  # noqa: E402

N, THREADS, NITER = 4096, 4, 3


def main():
    ref.build()
    w = workloads.make("merger_1e6", n_gas=N)
    r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), THREADS)

    # ---- Peano keys -----------------------------------------------------------------
    pts = [(0, 0, 0), (1, 1, 1), (0.25, 0.25, 0.25), (0.25, 0.25, 0.75), (0.25, 0.75, 0.25),
           (0.25, 0.75, 0.75), (0.75, 0.25, 0.25), (0.75, 0.25, 0.75), (0.75, 0.75, 0.25),
           (0.75, 0.75, 0.75), (0.5, 0.5, 0.5), (0.1, 0.2, 0.3), (0.7312, 0.1234, 0.9876),
           (0.999999, 1e-06, 0.5), (1.0, 0.3, 0.3), (0.0, 0.3, 0.3), (0.3, 1.0, 0.3),
           (0.3, 0.0, 0.3), (0.3, 0.3, 1.0),
           (float(np.float32(6961.5)) / 13923.0, float(np.float32(123.456)) / 13923.0,
            float(np.float32(13922.99)) / 13923.0)]
    rng = np.random.default_rng(7)
    pts += [tuple(map(float, rng.random(3))) for _ in range(40)]
    table = []
    for p in pts:
        hi, lo = r.peano_key(*p)
        rhi, rlo = r.peano_key(*p, reversed_=True)
        table.append(dict(xyz=list(p), key=[f"{hi:016x}", f"{lo:016x}"],
                          reversed=[f"{rhi:016x}", f"{rlo:016x}"]))
    with open(os.path.join(HERE, "peano_table.json"), "w") as f:
        json.dump(table, f, indent=1)

    # ---- the 4096-particle case ----------------------------------------------------------
    out = dict(pos0=w.pos, n_gas=N, boxsize=w.boxsize, mpart_gas=w.mpart_gas, mtotal=w.mtotal,
               halo_table=w.halo_table())
    r.load(w.pos)
    r.sort()
    r.build_tree()
    d = r.read()
    out.update(sort_id=d["id"], sort_key_hi=d["key_hi"], sort_key_lo=d["key_lo"],
               guess2=np.array([2 * r.guess_hsml(i) for i in range(N)], np.float32))
    queries, lists = [], []
    for i in (0, 1, 77, 1234, 2048, 4095):
        for h in (0.02 * w.boxsize, 0.08 * w.boxsize, 0.45 * w.boxsize):
            lst = r.find_ngb_tree(i, h)
            queries.append((i, h, len(lst)))
            lists.append(lst)
    out["ngb_queries"] = np.array(queries, np.float64)
    out["ngb_lists"] = np.concatenate(lists).astype(np.int32)

    r = ref.Ref(w.n_gas, w.boxsize, w.mpart_gas, w.mtotal, w.halo_table(), THREADS)
    r.load(w.pos)

    def cb(it):
        s = r.read()
        if it > 0:
            hw, dl = r.wvt_scratch()
            k = it - 1
            for name in ("id", "hsml", "rho", "varhsml", "rho_model", "pos"):
                out[f"it{k}_{name}"] = s[name]
            out[f"it{k}_hw"], out[f"it{k}_delta"] = hw, dl
        return 0

    r.regularise(NITER + 1, cb)
    log = ref.parse_log(r.log())
    out["log"] = np.array([[row[k] for k in ("it", "max", "mean", "diff", "step")] for row in log])

    # final density + rot(A) on the relaxed state (main.c:54-56)
    r.find_sph_quantities()
    s = r.read()
    apot = np.repeat(np.power(s["rho_model"] / s["rho_model"].max(), 0.5)[:, None], 3, 1).astype(np.float32)
    apot[:, 1] *= 0.5          # make the three components differ so every cross term is live
    apot[:, 2] *= 0.25
    r.set_apot(apot)
    r.bfld_from_rotA()
    s2 = r.read()
    for name in ("id", "pos", "hsml", "rho", "varhsml"):
        out[f"final_{name}"] = s[name]
    out["final_apot"], out["final_bfld"] = apot, s2["bfld"]
    np.savez_compressed(os.path.join(HERE, "merger_4096.npz"), **out)
    print("wrote", os.path.join(HERE, "merger_4096.npz"),
          os.path.getsize(os.path.join(HERE, "merger_4096.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
