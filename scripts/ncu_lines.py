"""Join an ncu SASS-page CSV with nvdisasm line info: instructions executed / stall samples per source line.
usage: ncu_lines.py prof_sass.csv all.sass <mangled kernel name fragment> [topN]"""
import csv, collections, re, sys
sass_csv, disasm, frag = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 50
# address -> (file,line) from nvdisasm
addr2line = {}
infunc = False; cur = None
for l in open(disasm):
    if l.startswith("//-----") and ".text." in l:
        infunc = frag in l
        continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
blk = rows[start[0] + 1:start[1]] if len(start) > 1 else rows[start[0] + 1:]
hdr = blk[0]; ix = {h: i for i, h in enumerate(hdr)}
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for r in blk[1:]:
    if len(r) < len(hdr): continue
    a = int(r[ix["Address"]], 16)
    if base is None: base = a
    e = float(r[ix["Instructions Executed"]] or 0); t = float(r[ix["Thread Instructions Executed"]] or 0)
    s = float(r[ix["# Samples"]] or 0)
    k = addr2line.get(a - base, ("?", 0))
    agg[k][0] += e; agg[k][1] += t; agg[k][2] += s
tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print("total warp-inst %.3g  samples %d" % (tot, ts))
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    f, ln = k
    if f not in src:
        try: src[f] = open("/root/repo/gpurun_out/r02s_" + f if f == "tile_fast.cuh" else "/root/repo/toycluster_b200/csrc/" + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    print(f"{f:12s}:{ln:4d} inst {v[0]/tot*100:5.1f}%  lanes {v[1]/max(v[0],1):4.1f}  samples {v[2]/ts*100:5.1f}% | {text}")
