import sys, os
import numpy as np
sys.path.insert(0, '.')
import toycluster_b200 as tc
rng = np.random.default_rng(9)
n, box = 30003, 4096.0
pos = rng.random((n, 3)).astype(np.float32) * np.float32(box)
k = 0
for cx in (0.0, box):
    for cy in (0.0, box):
        for cz in (0.0, box):
            off = (rng.random((600, 3)) * 30).astype(np.float32)
            c = np.array([cx, cy, cz], np.float32)
            pos[k:k + 600] = np.where(c == 0, off, np.float32(box) - off)
            k += 600
halo = np.array([[0, 0, 0, 1e-6, 0.54, 300.0, 3000.0, 0, 1.0]])
def run(flags, ncalls, hs0=None, notiles=False):
    if notiles: os.environ["TOYGPU_NO_TILES"] = "1"
    else: os.environ.pop("TOYGPU_NO_TILES", None)
    g = tc.HotPath(n, box, 1.0, 1e5, halo, flags=flags)
    g.upload(pos, hs0)
    for _ in range(ncalls): g.find_sph_quantities()
    o = g.download(); st = g.stats(); g.close()
    return o, st
def cmp(a, b, tag):
    assert np.array_equal(a["id"], b["id"])
    for k in ("hsml", "rho"):
        rel = np.abs(a[k].astype(np.float64) - b[k]) / np.abs(b[k])
        print(tag, k, "median %.2e q90 %.2e q99 %.2e max %.2e frac<=1e-5 %.4f" % (np.median(rel), np.quantile(rel, .9), np.quantile(rel, .99), rel.max(), (rel <= 1e-5).mean()))
ex1, _ = run(0, 1); fa1, s1 = run(tc.FAST, 1)
cmp(fa1, ex1, "cold: fast-generic vs exact      ")
ex2, _ = run(0, 2); fa2, s2 = run(tc.FAST, 2)
cmp(fa2, ex2, "cold+warm: fast vs exact         ")
# warm only from the SAME (exact cold) state
inv = np.empty(n, np.int64); inv[ex1["id"]] = np.arange(n)
h0 = ex1["hsml"][inv]
exw, _ = run(0, 1, h0); faw, sw = run(tc.FAST, 1, h0); fag, sg = run(tc.FAST, 1, h0, notiles=True)
cmp(faw, exw, "warm from exact cold: fast(tile) ")
cmp(fag, exw, "warm from exact cold: fast(gen)  ")
print("handed back", sw["handed_back"], sw["handback_why"], "searches", sw["searches"], "iters", sw["hsml_iters"], "exact iters", _["hsml_iters"] if False else "")

# ---- displacement: one WVT iteration from the same warm state, fast vs exact, path by path
def wvt(flags, notiles=False):
    if notiles: os.environ["TOYGPU_NO_TILES"] = "1"
    else: os.environ.pop("TOYGPU_NO_TILES", None)
    g = tc.HotPath(n, box, 1.0, 1e5, halo, flags=flags)
    g.upload(pos, h0)
    g.wvt_iteration(0.0085)
    hw, dl = g.wvt_scratch(); o = g.download(); st = g.stats(); g.close()
    return dl, o, st
de, oe, se = wvt(0, True)
for tag, fl, nt in (("exact tile", 0, False), ("fast tile ", tc.FAST, False), ("fast gen  ", tc.FAST, True)):
    d, o, st = wvt(fl, nt)
    sc = np.linalg.norm(de, axis=1)
    err = np.linalg.norm(d.astype(np.float64) - de, axis=1) / np.maximum(sc, 1e-30)
    print(tag, "delta rel err: median %.2e q99 %.2e q999 %.2e max %.2e | |delta|*box median %.3g max %.3g | handed back %d pairs %d vs %d" % (
        np.median(err), np.quantile(err, .99), np.quantile(err, .999), err.max(), np.median(sc) * box, sc.max() * box, st["handed_back"], st["pair_evals"], se["pair_evals"]))
    worst = np.argsort(err)[-3:]
    print("   worst:", [(int(k), float(err[k]), de[k].tolist(), d[k].tolist()) for k in worst])
