"""Group the per-line attribution of the tile sweep into phases."""
import re, sys, collections
groups = [("tile.cuh", 160, 232, "tile setup"), ("tile.cuh", 233, 280, "phase1 bit matrix"),
          ("tile.cuh", 281, 297, "phase2 setup"), ("tile.cuh", 298, 322, "expand"),
          ("tile.cuh", 323, 375, "classify+lists"), ("tile.cuh", 376, 405, "outer loop/compact"),
          ("tile.cuh", 406, 440, "results"), ("tile.cuh", 155, 159, "list get"),
          ("sph.cuh", 62, 74, "ngb_pred"), ("sph.cuh", 75, 90, "pair_r"), ("sph.cuh", 131, 156, "fdiv"),
          ("sph.cuh", 157, 163, "round_to_float"), ("sph.cuh", 164, 254, "find_hsml"),
          ("sph.cuh", 294, 325, "wvt_pair_fast"), ("common.cuh", 0, 999, "common helpers")]
agg = collections.Counter(); lanes = collections.Counter(); smp = collections.Counter()
for line in open(sys.argv[1]):
    m = re.match(r"(\S+)\s*:\s*(\d+) inst\s+([\d.]+)%\s+lanes\s+([\d.]+)\s+samples\s+([\d.]+)%", line)
    if not m: continue
    f, ln, pct, la, sp = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))
    name = "other (" + f + ")"
    for gf, lo, hi, gname in groups:
        if f == gf and lo <= ln <= hi: name = gname; break
    agg[name] += pct; lanes[name] += pct * la; smp[name] += sp
for k, v in agg.most_common():
    print(f"{k:28s} inst {v:5.1f}%  lanes {lanes[k]/v:4.1f}  samples {smp[k]:5.1f}%")
