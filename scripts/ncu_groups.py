"""Group the per-line output of ncu_lines.py by code region of tile_fast.cuh: ncu_groups.py lines.txt <inst per 1%> [source snapshot]"""
import re, collections, sys
per = float(sys.argv[2]) if len(sys.argv) > 2 else 42.0
src = open(sys.argv[3] if len(sys.argv) > 3 else '/root/repo/gpurun_out/r02q_tile_fast.cuh').read().splitlines()
def find(t): return [i + 1 for i, l in enumerate(src) if t in l][0]
marks = [(find('static __device__ __forceinline__ float rcp_approx'), 'find_hsml_fast'),
         (find('static __device__ __forceinline__ void tf_phase1_words'), 'phase 1'),
         (find('__global__ void __launch_bounds__(TF_WARPS * 32, TF_BLOCKS) k_sweep_tile_fast'), 'kernel setup / tile pipeline control'),
         (find('candidate runs of the tile'), 'run list'),
         (find('phase 1: lane = target'), 'phase 1 tickets'),
         (find('phase 2: one warp per target'), 'phase2 setup'),
         (find('(1) expand the bit row'), 'expand'),
         (find('float4 pi = a.pw[i];'), 'target constants'),
         (find('(2) classify every hit'), 'classify+wvt'),
         (find('for (int k = lane; k < nU + lane'), 'loop ctl / slow pass'),
         (find('(3) the outer loop'), 'outer loop / B copy'),
         (find('(4) results'), 'results')]
g = collections.defaultdict(lambda: [0.0, 0.0])
for l in open(sys.argv[1]):
    m = re.match(r'(\S+)\s*:\s*(\d+) inst\s+([\d.]+)%\s+lanes\s+([\d.]+)\s+samples\s+([\d.]+)%', l)
    if not m: continue
    f, ln, inst, smp = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(5))
    if f == 'tile_fast.cuh':
        k = 'header'
        for a, name in marks:
            if ln >= a: k = name
    elif f == 'tile.cuh': k = 'phase1 (tile.cuh)'
    else: k = f
    g[k][0] += inst; g[k][1] += smp
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0]):
    print(f'{k:30s} inst {v[0]:5.1f}%  {v[0]*per:6.0f}/target   samples {v[1]:5.1f}%')
