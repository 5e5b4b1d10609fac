// defect.cuh -- displaced nodes of the reference octree, reproduced.
//
// Find_ngb_tree (tree.c:25-111) is not the brute-force predicate: membership of a tree node
// is decided by key triplets, but the node's CENTRE is placed by comparing the position of
// the particle that creates it with the centre of the parent (tree.c:298-310).  A particle
// lying exactly on a centre plane of its parent cell (float positions: ~8 per 1e6 particles
// in the merger workloads), or within the rounding of a deep float centre, belongs to the
// upper cell while `Pos > centre` is false.  That node, and every descendant (their centres
// derive from it), is displaced by one cell size, and the open test of tree.c:56-58 then
// prunes particles that are within reach.  The reference's neighbour sets, and through them
// rho, hsml and the displacements of ~50 targets per event, are what they are only if this
// is reproduced.
//
// The box hierarchy of bvh.cuh finds the exact predicate set S_true.  The reference finds
//   S_ref = { j in S_true : every node on the path root -> leaf(j) passes the open test },
// and for a path whose centres are where the keys say, the open test cannot fail for a
// particle within reach.  So only particles underneath a displaced node need their path:
//
// A second way into the same situation: a coordinate equal to Boxsize scales to 2^63, whose key
// is not the key of the cell the particle sits in (peano.cuh), so the reference files the
// particle in a leaf somewhere else in the box.  It gets its own path (event level DF_SELF).
//
//   k_defect_detect  one thread per particle i: i creates the nodes of levels
//                    cpl[i]+1 .. max(cpl[i], cpl[i+1])+1 (cpl = common key triplets with the
//                    predecessor, guess.cuh).  Follow i's own coordinate bits down the float
//                    centre chain and compare `Pos > centre` (tree.c:298-302) with the bit:
//                    a disagreement is an event (i, level).  A superset is harmless.
//   k_defect_paths   every particle k of an event's cell: the LITERAL path of the reference,
//                    c_q = fl(c_{q-1} +- size_q/2) with the sign taken from the first
//                    particle of k's level-q cell (binary search on the sorted keys), for
//                    q = 1 .. leaf level (collapse rule of tree.c:201-226 as in guess.cuh);
//                    the sign bit of pw[k].w flags the particle, dmap[k] points at the path.
//   defect_open      the open tests of a flagged candidate's path, in the arithmetic of
//                    tree.c:37-58; the sweeps AND it into the particle predicate.
//
// Not reproduced (documented in DESIGN.md): nodes on the right spine of the tree keep
// DNext == 0, so a failed open test there steps into the children instead of pruning
// (tree.c:107); tree levels >= 31 (`1 << lvl` is undefined in tree.c:304); bit-identical keys.
#pragma once
#include "common.cuh"
#include "guess.cuh"

#define DF_MAX_LEVEL 30
#define DF_SELF 99                  // event level: only the particle itself needs a path

struct DefectTab {
    int2 *events;        // (first particle, level)
    int2 *big;           // [DF_BIG_CAP] (first, end) of cells too large for one block
    int *counts;         // [0] events, [1] path nodes used, [2] overflow, [3] flagged particles,
                         // [4] big cells
    float4 *nodes;       // (centre, size) per path level, terminated by size == 0
    int *dmap;           // [n] path offset of a flagged particle
    float *pwp;          // the pair-interleaved copy of pw (tile_fast.cuh): flagged there as well
    unsigned char *boxflag;  // [n / 32 + 1] level-0 boxes (32 particles) that hold a flagged particle
    int cap_events, cap_nodes;
};

// tree.c:304,308-310: size = (float)(Boxsize / (1 << lvl)); Pos = parent.Pos + sign * size * 0.5
static __device__ __forceinline__ float df_size(double box, int lvl)
{
    return (float)(box / (double)(1 << lvl));
}
static __device__ __forceinline__ float df_child(float parent, bool upper, float size)
{
    return (float)__dadd_rn((double)parent, (double)(upper ? size : -size) * 0.5);
}

__global__ void k_defect_detect(int n, const float4 *__restrict__ pw, double box,
                                const signed char *__restrict__ cpl, DefectTab d)
{
    // the node sizes of tree.c:304, one FP64 divide per level and block instead of per particle
    __shared__ float s_size[DF_MAX_LEVEL + 2];
    if (threadIdx.x <= DF_MAX_LEVEL) s_size[threadIdx.x] = df_size(box, threadIdx.x);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c0 = cpl[i], c1 = i + 1 < n ? (int)cpl[i + 1] : -1;
    const int lo = max(c0 + 1, 1);
    const int hi = min(max(c0, c1) + 1, DF_MAX_LEVEL);
    const float4 p = pw[i];
    const double scale = 9223372036854775808.0;   // 2^63, peano.c:134-136
    const uint64_t X = __double2ull_rz((double)p.x / box * scale);
    const uint64_t Y = __double2ull_rz((double)p.y / box * scale);
    const uint64_t Z = __double2ull_rz((double)p.z / box * scale);
    if ((X | Y | Z) >> 63) {
        // A coordinate equal to Boxsize (legal: the wrap of wvt_relax.c:200 is inclusive) scales
        // to 2^63, whose key is NOT that of the cell the particle sits in (peano.cuh): the
        // reference files the particle in a leaf somewhere else in the box, and finds it only
        // from there.  Its own literal path reproduces that; level DF_SELF = just this particle.
        const int e = atomicAdd(&d.counts[0], 1);
        if (e < d.cap_events) d.events[e] = make_int2(i, DF_SELF);
        // ... and the nodes it creates are placed by its position, far from their other members:
        // the sign test below disagrees with its (all-zero) low bits and flags them
    }
    if (hi < lo) return;
    float cx = (float)(box / 2), cy = cx, cz = cx;     // tree.c:133
    for (int q = 1; q <= hi; q++) {
        const bool bx = (X >> (63 - q)) & 1, by = (Y >> (63 - q)) & 1, bz = (Z >> (63 - q)) & 1;
        if (q >= lo && ((p.x > cx) != bx || (p.y > cy) != by || (p.z > cz) != bz)) {
            const int e = atomicAdd(&d.counts[0], 1);
            if (e < d.cap_events) d.events[e] = make_int2(i, q);
            return;                               // deeper nodes of i lie underneath this one
        }
        const float s = s_size[q];
        cx = df_child(cx, bx, s);
        cy = df_child(cy, by, s);
        cz = df_child(cz, bz, s);
    }
}

// The prefix mask of the first `level` key triplets.
static __device__ __forceinline__ void df_mask(int level, uint64_t &mh, uint64_t &ml)
{
    const int bits = 3 * level;
    if (bits == 0) { mh = 0; ml = 0; }
    else if (bits < 64) { mh = ~(~0ull >> bits); ml = 0; }
    else if (bits == 64) { mh = ~0ull; ml = 0; }
    else { mh = ~0ull; ml = ~(~0ull >> (bits - 64)); }
}

// Level of the leaf that holds particle p in the finished reference tree.
static __device__ int df_leaf_level(int p, int n, const signed char *__restrict__ cpl)
{
    for (int b = min(p + 8, n - 1); b > p; b--) {
        int start, level, count;
        collapse_event(b, cpl, start, level, count);
        if (start >= 0 && start <= p) return level;
    }
    int level = max((int)cpl[p], p + 1 < n ? (int)cpl[p + 1] : -1);
    if (level < 0) level = 0;
    return level + 1;
}

// Literal path of particle k (levels 1 .. leaf level), flag and map entry.
static __device__ void df_emit_path(int k, int n, float4 *__restrict__ pw, double box,
                                    const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                                    const signed char *__restrict__ cpl, const DefectTab &d)
{
    const int L = min(df_leaf_level(k, n, cpl), DF_MAX_LEVEL);
    const int off = atomicAdd(&d.counts[1], L + 1);
    if (off + L + 1 > d.cap_nodes) { d.counts[2] = 1; return; }
    const uint64_t kh = hi[k], kl = lo[k];
    float cx = (float)(box / 2), cy = cx, cz = cx;
    int f = 0;
    for (int q = 1; q <= L; q++) {
        uint64_t mh, ml;
        df_mask(q, mh, ml);
        const uint64_t ph = kh & mh, pl = kl & ml;          // smallest key of the cell
        int a = f, z = k;                                   // first particle of the cell
        while (a < z) {
            const int mid = (a + z) >> 1;
            if (key_less(hi[mid], lo[mid], ph, pl)) a = mid + 1; else z = mid;
        }
        f = a;
        const float4 pf = pw[f];
        const float s = df_size(box, q);
        cx = df_child(cx, pf.x > cx, s);                    // tree.c:298-310
        cy = df_child(cy, pf.y > cy, s);
        cz = df_child(cz, pf.z > cz, s);
        d.nodes[off + q - 1] = make_float4(cx, cy, cz, s);
    }
    d.nodes[off + L] = make_float4(0, 0, 0, 0);
    d.dmap[k] = off;
    float *w = &pw[k].w;
    *w = __int_as_float(__float_as_int(*w) | 0x80000000);
#ifndef TF_NO_PWP_FLAG
    // ... and in the copy the packed phase 2 of the fast tile sweep gathers from: pair k / 2 is
    // {x0, x1, y0, y1}, {z0, z1, w0, w1}
    float *w2 = d.pwp + 8 * (size_t)(k >> 1) + (k & 1) + 6;
    *w2 = __int_as_float(__float_as_int(*w2) | 0x80000000);
#endif
    d.boxflag[k >> 5] = 1;
    atomicAdd(&d.counts[3], 1);
}

#define DF_BIG 8192          // cells above this size are shared by the whole grid (second kernel)
#define DF_BIG_CAP 64

// One block per event.  Cells larger than DF_BIG are only recorded in d.big.
__global__ void k_defect_paths(int n, float4 *__restrict__ pw, double box,
                               const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                               const signed char *__restrict__ cpl, DefectTab d)
{
    __shared__ int s_first, s_end;
    const int nev = min(d.counts[0], d.cap_events);
    if (blockIdx.x == 0 && threadIdx.x == 0 && d.counts[0] > d.cap_events) d.counts[2] = 1;
    for (int e = blockIdx.x; e < nev; e += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const int2 ev = d.events[e];
            const int i = ev.x, m = ev.y;
            int end = i;
            if (m == DF_SELF) end = i + 1;
            else if (df_leaf_level(i, n, cpl) >= m) {  // else the node was collapsed away
                // one past the last particle of i's level-m cell
                uint64_t mh, ml;
                df_mask(m, mh, ml);
                const uint64_t uh = hi[i] | ~mh, ul = lo[i] | ~ml;   // largest key of the cell
                int a = i + 1, z = n;
                while (a < z) {
                    const int mid = (a + z) >> 1;
                    if (!key_less(uh, ul, hi[mid], lo[mid])) a = mid + 1; else z = mid;
                }
                end = a;
                if (end - i > DF_BIG) {
                    const int b = atomicAdd(&d.counts[4], 1);
                    if (b < DF_BIG_CAP) d.big[b] = make_int2(i, end); else d.counts[2] = 1;
                    end = i;
                }
            }
            s_first = i; s_end = end;
        }
        __syncthreads();
        const int first = s_first, end = s_end;
        for (int k = first + threadIdx.x; k < end; k += blockDim.x)
            df_emit_path(k, n, pw, box, hi, lo, cpl, d);
    }
}

__global__ void k_defect_paths_big(int n, float4 *__restrict__ pw, double box,
                                   const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                                   const signed char *__restrict__ cpl, DefectTab d)
{
    const int nbig = min(d.counts[4], DF_BIG_CAP);
    for (int b = 0; b < nbig; b++) {
        const int2 r = d.big[b];
        for (int k = r.x + blockIdx.x * blockDim.x + threadIdx.x; k < r.y; k += gridDim.x * blockDim.x)
            df_emit_path(k, n, pw, box, hi, lo, cpl, d);
    }
}

// tree.c:37-58 for every node on the path of a flagged candidate.
static __device__ __noinline__ bool defect_open(const float4 *__restrict__ path, float xi, float yi,
                                                float zi, float h, float box, float boxhalf)
{
    for (;; path++) {
        const float4 c = *path;
        if (c.w == 0.f) return true;
        float dx = fabsf(__fsub_rn(xi, c.x)), dy = fabsf(__fsub_rn(yi, c.y)), dz = fabsf(__fsub_rn(zi, c.z));
        if (dx > boxhalf) dx = __fsub_rn(dx, box);
        if (dy > boxhalf) dy = __fsub_rn(dy, box);
        if (dz > boxhalf) dz = __fsub_rn(dz, box);
        const float dl = (float)__dadd_rn(__dmul_rn(0.5 * K_SQRT3, (double)c.w), (double)h);
        if (!(sq3_nofma(dx, dy, dz) < __fmul_rn(dl, dl))) return false;
    }
}

static __device__ __forceinline__ bool df_flagged(float w) { return __float_as_int(w) < 0; }
