// bvh.cuh -- the neighbour index: a pointer-free bounding-box hierarchy over the Peano-sorted
// particle array, replacing the serial octree of tree.c:124-271.
//
// Because the particles are in Peano-Hilbert order, any run of consecutive particles is
// spatially compact.  Level 0 boxes bound runs of 32 particles (one coalesced 512-byte float4
// load per visit), level l+1 boxes bound 32 level-l boxes.  Boxes are tight (min/max of the
// actual positions), so the structure never depends on key <-> cell geometry (e.g. the
// x == Boxsize edge case of peano.c) and is rebuilt in O(N) with no sorting or pointers.
//
// The contract of Find_ngb_tree (tree.c:25-111) is its particle predicate, not its node
// walk: all j with periodic minimum-image r^2 < h^2 evaluated in FLOAT WITHOUT FMA
// (tree.c:67-88), ascending in j, stopping at NGBMAX.  The box test below only has to be
// conservative with respect to that predicate; boxes are inflated by 4e-7*Boxsize and the
// radius by 1e-5 relative to cover every rounding on either side.
#pragma once
#include "common.cuh"

// level 0: one warp per group of 32 particles; also the boxes of its four runs of 8
__global__ void k_bvh_leaves(int n, const float4 *__restrict__ pw, int n0, float pad,
                             float *cx, float *cy, float *cz, float *hx, float *hy, float *hz,
                             float *scx, float *scy, float *scz, float *shx, float *shy, float *shz)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= n0) return;
    const int lane = lane_id();
    const int k = g * 32 + lane;
    float lx = 3.0e38f, ly = 3.0e38f, lz = 3.0e38f, ux = -3.0e38f, uy = -3.0e38f, uz = -3.0e38f;
    if (k < n) {
        const float4 p = pw[k];
        lx = ux = p.x; ly = uy = p.y; lz = uz = p.z;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        lx = fminf(lx, __shfl_xor_sync(FULL_MASK, lx, o));
        ly = fminf(ly, __shfl_xor_sync(FULL_MASK, ly, o));
        lz = fminf(lz, __shfl_xor_sync(FULL_MASK, lz, o));
        ux = fmaxf(ux, __shfl_xor_sync(FULL_MASK, ux, o));
        uy = fmaxf(uy, __shfl_xor_sync(FULL_MASK, uy, o));
        uz = fmaxf(uz, __shfl_xor_sync(FULL_MASK, uz, o));
        if (o == 4 && (lane & 7) == 0) {      // reduced over aligned runs of 8 lanes
            const int s = 4 * g + (lane >> 3);
            scx[s] = 0.5f * (lx + ux); shx[s] = 0.5f * (ux - lx) + pad;
            scy[s] = 0.5f * (ly + uy); shy[s] = 0.5f * (uy - ly) + pad;
            scz[s] = 0.5f * (lz + uz); shz[s] = 0.5f * (uz - lz) + pad;
        }
    }
    if (lane == 0) {
        cx[g] = 0.5f * (lx + ux); hx[g] = 0.5f * (ux - lx) + pad;
        cy[g] = 0.5f * (ly + uy); hy[g] = 0.5f * (uy - ly) + pad;
        cz[g] = 0.5f * (lz + uz); hz[g] = 0.5f * (uz - lz) + pad;
    }
}

// level l+1 from level l: one warp per parent, lane = child
__global__ void k_bvh_up(int n_child, int n_parent, float pad,
                         const float *ccx, const float *ccy, const float *ccz,
                         const float *chx, const float *chy, const float *chz,
                         float *cx, float *cy, float *cz, float *hx, float *hy, float *hz)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= n_parent) return;
    const int lane = lane_id();
    const int k = g * 32 + lane;
    float lx = 3.0e38f, ly = 3.0e38f, lz = 3.0e38f, ux = -3.0e38f, uy = -3.0e38f, uz = -3.0e38f;
    if (k < n_child) {
        lx = ccx[k] - chx[k]; ux = ccx[k] + chx[k];
        ly = ccy[k] - chy[k]; uy = ccy[k] + chy[k];
        lz = ccz[k] - chz[k]; uz = ccz[k] + chz[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lx = fminf(lx, __shfl_xor_sync(FULL_MASK, lx, o));
        ly = fminf(ly, __shfl_xor_sync(FULL_MASK, ly, o));
        lz = fminf(lz, __shfl_xor_sync(FULL_MASK, lz, o));
        ux = fmaxf(ux, __shfl_xor_sync(FULL_MASK, ux, o));
        uy = fmaxf(uy, __shfl_xor_sync(FULL_MASK, uy, o));
        uz = fmaxf(uz, __shfl_xor_sync(FULL_MASK, uz, o));
    }
    if (lane == 0) {
        cx[g] = 0.5f * (lx + ux); hx[g] = 0.5f * (ux - lx) + pad;
        cy[g] = 0.5f * (ly + uy); hy[g] = 0.5f * (uy - ly) + pad;
        cz[g] = 0.5f * (lz + uz); hz[g] = 0.5f * (uz - lz) + pad;
    }
}

// Periodic distance^2 from a point to a box (centre c, half-width h), per axis through the
// nearest image of the box centre.
static __device__ __forceinline__ float box_dist2(float x, float y, float z, float cx, float cy,
                                                  float cz, float hx, float hy, float hz,
                                                  float box, float boxhalf)
{
    float dx = fabsf(x - cx), dy = fabsf(y - cy), dz = fabsf(z - cz);
    if (dx > boxhalf) dx = box - dx;
    if (dy > boxhalf) dy = box - dy;
    if (dz > boxhalf) dz = box - dz;
    dx = fmaxf(dx - hx, 0.f);
    dy = fmaxf(dy - hy, 0.f);
    dz = fmaxf(dz - hz, 0.f);
    return dx * dx + dy * dy + dz * dz;
}

// Warp-cooperative ordered walk: calls visit(group) for every level-0 box accepted by
// test(o) (o = index into the box arrays; the test must be monotone, i.e. accept every
// ancestor of an accepted box), in ascending group order -- hence ascending particle index,
// like the depth-first walk of tree.c:35-108.  visit returns false to stop early.  All 32
// lanes must call this with identical arguments.  Per-level state (pending-children mask,
// parent index) lives in the registers of lane == level, so the walk uses no local or shared
// memory.
template <class Test, class Visit>
static __device__ __forceinline__ void bvh_walk_pred(const Bvh &t, Test &&test, Visit &&visit)
{
    const int lane = lane_id();
    unsigned my_mask = 0;
    int my_parent = 0;

    auto test_children = [&](int level, int parent) -> unsigned {
        const int k = parent * 32 + lane;
        bool hit = false;
        if (k < t.lvl_n[level]) hit = test(t.lvl_off[level] + k);
        return __ballot_sync(FULL_MASK, hit);
    };

    int level = t.top;
    {
        const unsigned m = test_children(level, 0);
        if (lane == level) { my_mask = m; my_parent = 0; }
    }
    for (;;) {
        unsigned m = __shfl_sync(FULL_MASK, my_mask, level);
        if (m == 0) {
            if (level == t.top) break;
            level++;
            continue;
        }
        const int parent = __shfl_sync(FULL_MASK, my_parent, level);
        const int child = parent * 32 + (__ffs(m) - 1);
        m &= m - 1;
        if (lane == level) my_mask = m;
        if (level == 0) {
            if (!visit(child)) break;
        } else {
            level--;
            const unsigned cm = test_children(level, child);
            if (lane == level) { my_mask = cm; my_parent = child; }
        }
    }
}

// Boxes within `radius` of the point (x,y,z).
template <class Visit>
static __device__ __forceinline__ void bvh_walk(const Bvh &t, const Box &bx, float x, float y,
                                                float z, float radius, Visit &&visit)
{
    const float r2 = radius * radius * 1.00001f;
    bvh_walk_pred(t, [&](int o) -> bool {
        return box_dist2(x, y, z, t.cx[o], t.cy[o], t.cz[o], t.hx[o], t.hy[o], t.hz[o],
                         bx.box_f, bx.boxhalf_f) <= r2;
    }, visit);
}
