"""Static SASS of a source-line range of one kernel: sass_region.py <kernel fragment> <file> <first marker text> <last marker text>"""
import re, subprocess, sys, os, tempfile
frag, fname, m0, m1 = sys.argv[1:5]
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "toycluster_b200", "libtoygpu.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
infunc = False; cur = None; out = []
for l in sass.splitlines():
    if l.startswith("//-----") and ".text." in l:
        infunc = frag in l; continue
    if not infunc: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: out.append((int(m.group(1), 16), cur, m.group(2)))
src = open(os.path.join(os.path.dirname(so), "csrc", fname)).read().splitlines()
l0 = [i + 1 for i, t in enumerate(src) if m0 in t][0]
l1 = [i + 1 for i, t in enumerate(src) if m1 in t][0]
idx = [i for i, (a, c, t) in enumerate(out) if c and c[0] == fname and l0 <= c[1] <= l1]
lo, hi = min(idx), max(idx)
print(f"lines {l0}-{l1}: {hi - lo + 1} static instructions (whole kernel {len(out)})")
for a, c, t in out[lo:hi + 1]:
    print(f'{a:05x} {c[0][:10]:10s}:{c[1]:4d}  {t}')
