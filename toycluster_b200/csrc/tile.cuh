// tile.cuh -- sweep v2: the warm-start fast path of the neighbour sweep.
//
// ncu on sweep v1 (profiles/r01_v1_*) showed an issue-bound kernel spending ~22 % of its
// instructions walking the box hierarchy once per target and ~25 % on a candidate scan that
// is only 8 % efficient.  Consecutive Peano-ordered targets have almost the same
// neighbourhood, so v2 shares that work across a TILE of 32 consecutive targets (one leaf
// box of the index):
//
//   k_tile_walk   one warp per tile: R = max over the tile of the largest radius any of its
//                 targets can need without a third search (1.23*Hsml, sph.c:51, and the WVT
//                 radius, wvt_relax.c:135); breadth-first descent (order kept) with a box-to-box distance test, at
//                 the leaves run-against-run on 8-particle sub-boxes -> ascending list of
//                 (box, run mask) entries in global memory (~370 runs = 2900 candidates).
//   k_sweep_tile  one 8-warp block per tile, 3 blocks per SM:
//     phase 1  one LANE per TARGET, all lanes read the same candidate (one broadcast 16-byte
//              load per candidate, served by L1/L2): an FMA-contracted SUPERSET of the float
//              predicate of tree.c:67-88 at radius R_i -> one bit per (target, candidate) in
//              a shared-memory bit matrix.  No divergence, no ballots, no per-target walk.
//     phase 2  one WARP per target: expand the target's bit row into a compact hit list,
//              re-evaluate every hit with the exact FMA-free predicate against Hsml,
//              1.23*Hsml and the WVT radius (and, for a hit underneath a displaced reference
//              node, the open tests of defect.cuh), build the FP64 separation list with full
//              lanes, queue the displacement partners and evaluate them 32 at a time, run
//              Find_hsml (sph.c:80-214) -- the same device functions as the generic sweep.
//   Anything outside the fast path's envelope (cold start, list overflow, a third search,
//   Find_hsml not converging on the frozen list) is pushed to a work list and redone from
//   scratch by the generic kernel, so results do not depend on which path ran
//   (tests/test_gpu_golden.py::test_tile_path_equals_generic_path).
//
// Tuning record (B200, 1 M-particle merger, sweep ms; see profiles/):
//   v1 generic 25.4 | tiles, 2 blocks/SM 22.8 | + hoisted divide, 2-way ILP, FMA prefilter,
//   aliased lists, 3 blocks/SM 14.7 | + sub-box walk 14.0 | + XU relief 13.3 | + exact u-clamp,
//   interior tiles 12.8.  Tried and dropped: 4 blocks/SM with the list tail in global memory
//   (13.2: the register cap and the split accessor cost more than the occupancy gains),
//   larger caps 640 runs / 768 hits (13.3: fewer hand-backs but a smaller L1 carve-out),
//   smaller caps (13.2-15.3: hand-backs), software-pipelined candidate fetch (13.1),
//   per-sub-run candidate lists (only 23 % fewer candidates for a second bit matrix).
// 10 M-particle merger, whole step ms: 129.6 | packed FP32 (FADD2/FMUL2/FFMA2) in Find_hsml
//   and phase 1 121.0 | hit-list expansion as a uniform unrolled bit loop (was a loop over set
//   bits: 17 % of all instructions at 6.8 active lanes) 112.2 | displacement partners queued
//   for full-lane evaluation 110.8.  Tried and dropped: lane = bit expansion with two popc per
//   word (120.7: the XU pipe), Hsml entries first + separations of the "1.23*Hsml only" hits
//   deferred to the second search + skipping them in Find_hsml iterations below Hsml (121.6:
//   13 % fewer kernel evaluations, but the second gather of those hits is exposed L2 latency),
//   6 / 5 / 4 / 12 warps per block at 4 / 4 / 5 / 2 blocks per SM (112.6 / 114.1 / 113.5 / 114.6
//   against 108.9 for 8 x 3), four list entries per lane and iteration in Find_hsml (111.4),
//   caps 640 and 768 hits (108.6, 110.4).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "bvh.cuh"
#include "sph.cuh"
#include "f32x2.cuh"

#ifndef TL_WARPS
#define TL_WARPS 8
#endif
#ifndef TL_BLOCKS
#define TL_BLOCKS 3                      // resident blocks per SM the kernel is compiled for
#endif
#ifndef TL_ENT
#define TL_ENT 320                       // candidate level-0 boxes per tile (global list entries)
#endif
#ifndef TL_RUNS
#define TL_RUNS 768                      // candidate runs of 8 particles per tile (a multiple of 128).  512 sent 0.3 %
                                         // of the merger's tiles -- 521 to 813 runs -- to the generic sweep, 0.8 ms of
                                         // its latency-bound tail per step at 10 M; 768 keeps four fifths of them
#endif
#ifndef TL_CAP
#define TL_CAP 704                       // hits within R_i / density list entries per target
#endif
#define TL_WORDS (TL_RUNS / 4)           // bit-matrix words per target (4 runs = 32 candidates)
#define TL_UCAP TL_CAP
#define TL_LCAP TL_CAP
#define TL_MSTRIDE 33                    // bit-matrix row stride (words), odd => no bank conflicts

// Shared memory layout (bytes): bit matrix, run list, one list region per warp.  The region
// holds TL_CAP doubles (the separation list); the 16-bit hit list lives in its last quarter:
// entry k of the hit list is dead once batch k/32 has been read, and the separation list
// can only have grown to 8*(k+32) <= 6*TL_CAP + 2*(k+32) bytes by then.
#define TL_OFF_MASK 0
#define TL_OFF_RUN (TL_OFF_MASK + TL_WORDS * TL_MSTRIDE * 4)
#define TL_OFF_RL (TL_OFF_RUN + TL_RUNS * 4)
#define TL_OFF_MISC (TL_OFF_RL + TL_WARPS * TL_CAP * 8)
#define TL_OFF_WQ (TL_OFF_MISC + 64)           // misc: 2 ints
#define TL_SMEM (TL_OFF_WQ + TL_WARPS * 128)   // per warp: queue of 64 displacement partners

// Periodic gap^2 between two boxes (centre/half-width form).
template <bool WRAP = true>
static __device__ __forceinline__ float box_box_dist2(float ax, float ay, float az, float ahx,
                                                      float ahy, float ahz, float bx, float by,
                                                      float bz, float bhx, float bhy, float bhz,
                                                      float box, float boxhalf)
{
    float dx = fabsf(ax - bx), dy = fabsf(ay - by), dz = fabsf(az - bz);
    if (WRAP) {
        if (dx > boxhalf) dx = box - dx;
        if (dy > boxhalf) dy = box - dy;
        if (dz > boxhalf) dz = box - dz;
    }
    dx = fmaxf(dx - (ahx + bhx), 0.f);
    dy = fmaxf(dy - (ahy + bhy), 0.f);
    dz = fmaxf(dz - (ahz + bhz), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

// Largest search radius target i can need on the fast path.
// rmode 1: one search at Hsml itself (rot A, sph.c:229).
static __device__ __forceinline__ float tile_radius(float hA, float hw_raw, float norm, double box, int rmode = 0)
{
    if (rmode == 1) return hA;
    const float hB = (float)((double)hA * 1.23);                   // sph.c:51
    const float hsw = (float)((double)__fmul_rn(fabsf(hw_raw), norm) * box);   // wvt_relax.c:124,135
    return fmaxf(hB, hsw);
}

// Candidate list of a tile.  tile_ng[tile] = number of entries (bit 30 set: no target of the
// tile can see a periodic image, so phase 1 may skip the wrap), or <= -1 when the tile must
// take the generic path.  An entry is (level-0 box << 4) | mask of its four 8-particle runs
// that lie within reach of the tile.  "Reach" is tested run against run: the tile's own four
// sub-boxes, each with the largest radius of its 8 targets, against the candidate's four.
// One warp per tile.  The hierarchy is descended BREADTH first, level by level, through a
// per-warp queue in shared memory: the accepted nodes of a level stay in ascending order
// (ballot compaction), so the result equals the ordered depth-first walk, but a level's box
// tests are independent loads instead of one dependent chain per visited node -- the
// depth-first version spent ~300 serialised L2 round trips per tile (5.2 ms per step at 10 M).
#define TW_WARPS 8
#define TW_QCAP (TL_ENT + 64)            // accepted nodes per level a tile may have (TL_ENT + slack)

__global__ void __launch_bounds__(TW_WARPS * 32)
k_tile_walk(Bvh t, Box bx, const float4 *__restrict__ pw, const float *__restrict__ hsml_in,
            const double *__restrict__ vsum, int tile_lo, int tile_hi, int *__restrict__ tile_ng,
            int *__restrict__ tile_groups, int rmode)
{
    __shared__ int s_queue[TW_WARPS][2][TW_QCAP];
    const int wib = threadIdx.x >> 5;
    const int tile = tile_lo + blockIdx.x * TW_WARPS + wib;
    if (tile >= tile_hi) return;
    const int lane = lane_id();
    const unsigned lt = (1u << lane) - 1;
    const int i = tile * 32 + lane;
    const float norm = (float)pow(TG_DESNNGB / *vsum / K_FOURPITHIRD, 1.0 / 3.0);

    float R = 0;
    bool cold = false;
    if (i < t.n) {
        const float hA = hsml_in[i];
        cold = hA == 0;
        R = tile_radius(hA, pw[i].w, norm, bx.box_d, rmode);
    }
    R = fmaxf(R, __shfl_xor_sync(FULL_MASK, R, 1));
    R = fmaxf(R, __shfl_xor_sync(FULL_MASK, R, 2));
    R = fmaxf(R, __shfl_xor_sync(FULL_MASK, R, 4));
    const float Rsub = R;                       // max over this lane's run of 8 targets
    R = fmaxf(R, __shfl_xor_sync(FULL_MASK, R, 8));
    R = fmaxf(R, __shfl_xor_sync(FULL_MASK, R, 16));
    if (__any_sync(FULL_MASK, cold) || !(R < 0.98f * bx.boxhalf_f)) {     // (0.49 Boxsize: tile_fast.cuh's wrap)
        if (lane == 0) tile_ng[tile] = -1;
        return;
    }
    const float ax = t.cx[tile], ay = t.cy[tile], az = t.cz[tile];
    const float ahx = t.hx[tile], ahy = t.hy[tile], ahz = t.hz[tile];
    const float R2 = R * R * 1.00001f;
    // no target of the tile can see a periodic image (98.7 % of the tiles of the merger): then a
    // box within reach is within reach WITHOUT the wrap, and the rest of the walk -- issue bound,
    // ~5 k instructions per tile, most of them box distances -- is compiled without it
    const float mrg = R * 1.0001f;     // margin: a wrapped pair must stay a miss after rounding
    const bool interior = ax - ahx - mrg >= 0 && ax + ahx + mrg <= bx.box_f &&
                          ay - ahy - mrg >= 0 && ay + ahy + mrg <= bx.box_f &&
                          az - ahz - mrg >= 0 && az + ahz + mrg <= bx.box_f;
    auto body = [&](auto wrap_tag) {
    constexpr bool WRAP = decltype(wrap_tag)::value;
    auto test = [&](int o) -> bool {
        return box_box_dist2<WRAP>(ax, ay, az, ahx, ahy, ahz, t.cx[o], t.cy[o], t.cz[o], t.hx[o], t.hy[o],
                                   t.hz[o], bx.box_f, bx.boxhalf_f) <= R2;
    };

    // ---- descend: queue of accepted nodes per level, ascending -------------------------
    int *cur = s_queue[wib][0], *nxt = s_queue[wib][1];
    int ncur;
    {
        const int level = t.top;
        const bool hit = lane < t.lvl_n[level] && test(t.lvl_off[level] + lane);
        const unsigned m = __ballot_sync(FULL_MASK, hit);
        if (hit) cur[__popc(m & lt)] = lane;
        ncur = __popc(m);
    }
    bool overflow = false;
    for (int level = t.top - 1; level >= 0 && !overflow; level--) {
        __syncwarp();
        const int off = t.lvl_off[level], nl = t.lvl_n[level];
        int nnext = 0;
        for (int p = 0; p < ncur; p += 2) {              // two parents in flight
            const int k0 = cur[p] * 32 + lane;
            const int k1 = p + 1 < ncur ? cur[p + 1] * 32 + lane : nl;
            const bool h0 = k0 < nl && test(off + k0);
            const bool h1 = k1 < nl && test(off + k1);
            const unsigned m0 = __ballot_sync(FULL_MASK, h0), m1 = __ballot_sync(FULL_MASK, h1);
            const int c0 = __popc(m0), c1 = __popc(m1);
            if (nnext + c0 + c1 > TW_QCAP) { overflow = true; break; }
            if (h0) nxt[nnext + __popc(m0 & lt)] = k0;
            if (h1) nxt[nnext + c0 + __popc(m1 & lt)] = k1;
            nnext += c0 + c1;
        }
        int *sw = cur; cur = nxt; nxt = sw;
        ncur = nnext;
    }
    __syncwarp();
    if (overflow) {
        if (lane == 0) tile_ng[tile] = -2;
        return;
    }

    // ---- level-0 boxes: run against run, EIGHT boxes per pass ---------------------------
    // lane = (candidate box c = lane / 4 of the pass, its run b = lane % 4); every lane holds the
    // tile's own four sub-boxes (each with the largest radius of its 8 targets) in registers and
    // tests its candidate run against all four, so a pass is one round of independent loads for
    // eight boxes (it was one per two boxes: ~50 dependent L2 round trips per tile).
    const int cb = lane >> 2, rb = lane & 3;
    float tcx[4], tcy[4], tcz[4], thx[4], thy[4], thz[4], Ra2[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int sa = 4 * tile + q;
        tcx[q] = t.scx[sa]; tcy[q] = t.scy[sa]; tcz[q] = t.scz[sa];
        thx[q] = t.shx[sa]; thy[q] = t.shy[sa]; thz[q] = t.shz[sa];
        const float Ra = __shfl_sync(FULL_MASK, Rsub, q * 8);
        Ra2[q] = Ra * Ra * 1.00001f;
    }
    int ng = 0, nruns = 0;
    int *out = tile_groups + (size_t)tile * TL_ENT;
    for (int j = 0; j < ncur; j += 8) {
        const bool have = j + cb < ncur;
        const int g = have ? cur[j + cb] : 0;
        const int sb = 4 * g + rb;
        const float bx_ = t.scx[sb], by_ = t.scy[sb], bz_ = t.scz[sb];
        const float bhx = t.shx[sb], bhy = t.shy[sb], bhz = t.shz[sb];
        bool near = false;
#pragma unroll
        for (int q = 0; q < 4; q++)
            near |= box_box_dist2<WRAP>(tcx[q], tcy[q], tcz[q], thx[q], thy[q], thz[q], bx_, by_, bz_, bhx, bhy, bhz,
                                        bx.box_f, bx.boxhalf_f) <= Ra2[q];
        near &= have;
        const unsigned m = __ballot_sync(FULL_MASK, near);
        // boxes of the pass with at least one run in reach, in ascending order
        unsigned nz = (m | (m >> 1) | (m >> 2) | (m >> 3)) & 0x11111111u;      // bit 4c: box c has a run
        const int mine = (int)((m >> (4 * cb)) & 0xfu);
        if (rb == 0 && mine) {
            const int pos = ng + __popc(nz & ((1u << (4 * cb)) - 1u));
            if (pos < TL_ENT) out[pos] = (g << 4) | mine;
        }
        ng += __popc(nz);
        nruns += __popc(m);
    }
    if (lane == 0) {
        if (ng > TL_ENT || nruns > TL_RUNS) tile_ng[tile] = -max(nruns, 2);
        else tile_ng[tile] = ng | (nruns << 12) | (interior ? (1 << 30) : 0);
    }
    };      // body
    if (interior) body(std::false_type{});
    else body(std::true_type{});
}

// Phase 1 of the tile sweep: lane = target, one bit per (target, candidate).
// The bit only has to be a SUPERSET of the exact predicate (phase 2 re-evaluates every hit
// with the FMA-free arithmetic of tree.c:88), so the sum of squares is contracted to two FMAs
// and compared against a radius inflated by 2e-6.  Two candidates per instruction: the SoA
// copy of the positions delivers aligned pairs, and FADD2/FMUL2/FFMA2 (f32x2.cuh) work on both
// at once (the target's coordinate is the scalar operand).  The periodic wrap is
// d - Boxsize * rint(d / Boxsize) by the 1.5*2^23 trick, which differs from the reference's
// only for |d| within 1e-7 of Boxsize/2 -- where it changes d^2 by less than the inflation.
template <bool INTERIOR, int WARPS = TL_WARPS>
static __device__ __forceinline__ void tile_phase1(const float *__restrict__ sx, const float *__restrict__ sy,
                                                   const float *__restrict__ sz, const int *s_run,
                                                   unsigned *s_mask, int w, int lane, int ng, int nruns,
                                                   float xi, float yi, float zi, float R2, float box)
{
    const float R2p = R2 * 1.000002f;
    const f32x2 xi2 = pack2(xi, xi), yi2 = pack2(yi, yi), zi2 = pack2(zi, zi);
    const float ibox = 1.f / box;
    const f32x2 ib2 = pack2(ibox, ibox), mg2 = pack2(12582912.f, 12582912.f);
    const f32x2 nb2 = pack2(-box, -box);
    for (int q = w; q < ng; q += WARPS) {
        unsigned word = 0;
#pragma unroll 1
        for (int c = 0; c < 4; c++) {            // four runs of 8 candidates per word
            const int r8 = 4 * q + c;
            if (r8 >= nruns) break;
            const int first = s_run[r8];         // multiple of 8: 32-byte aligned rows
            // same address in every lane: broadcast loads, six in flight
            const ulonglong2 x0 = __ldg((const ulonglong2 *)(sx + first)), x1 = __ldg((const ulonglong2 *)(sx + first) + 1);
            const ulonglong2 y0 = __ldg((const ulonglong2 *)(sy + first)), y1 = __ldg((const ulonglong2 *)(sy + first) + 1);
            const ulonglong2 z0 = __ldg((const ulonglong2 *)(sz + first)), z1 = __ldg((const ulonglong2 *)(sz + first) + 1);
            unsigned sub = 0;
            auto test2 = [&](f32x2 X, f32x2 Y, f32x2 Z, unsigned b0, unsigned b1) {
                f32x2 dx = sub2(xi2, X), dy = sub2(yi2, Y), dz = sub2(zi2, Z);
                if (!INTERIOR) {
                    dx = fma2(sub2(fma2(dx, ib2, mg2), mg2), nb2, dx);
                    dy = fma2(sub2(fma2(dy, ib2, mg2), mg2), nb2, dy);
                    dz = fma2(sub2(fma2(dz, ib2, mg2), mg2), nb2, dz);
                }
                float s0, s1;
                unpack2(fma2(dz, dz, fma2(dy, dy, mul2(dx, dx))), s0, s1);
                if (s0 < R2p) sub |= b0;
                if (s1 < R2p) sub |= b1;
            };
            test2(x0.x, y0.x, z0.x, 1u, 2u);
            test2(x0.y, y0.y, z0.y, 4u, 8u);
            test2(x1.x, y1.x, z1.x, 16u, 32u);
            test2(x1.y, y1.y, z1.y, 64u, 128u);
            word |= sub << (8 * c);              // the pad of a short last run is far away
        }
        s_mask[q * TL_MSTRIDE + lane] = word;
    }
}

struct TileList {     // density list: r as double; the sign bit marks "outside the Hsml list"
    double *sm;
    __device__ __forceinline__ double get(int k) const { return fabs(sm[k]); }
};

template <int MODE>
__global__ void __launch_bounds__(TL_WARPS * 32, TL_BLOCKS) k_sweep_tile(const SweepArgs a, int tile_lo,
                                                                   int tile_hi)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *s_mask = (unsigned *)(smem + TL_OFF_MASK);
    int *s_run = (int *)(smem + TL_OFF_RUN);        // first particle of each candidate run
    int *s_misc = (int *)(smem + TL_OFF_MISC);       // [0] tile, [1] next target

    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1;
    double *rl = (double *)(smem + TL_OFF_RL) + w * TL_CAP;
    unsigned short *ul = (unsigned short *)(rl + (TL_CAP / 4) * 3);   // last quarter of the region
    TileList L{rl};
    unsigned short *wq = (unsigned short *)(smem + TL_OFF_WQ) + w * 64;

    const float norm = (float)pow(TG_DESNNGB / *a.vsum / K_FOURPITHIRD, 1.0 / 3.0);   // wvt_relax.c:120
    const float box = a.bx.box_f, boxhalf = a.bx.boxhalf_f;
    const float binv_hi = (float)a.bx.boxinv_d, binv_lo = (float)(a.bx.boxinv_d - (double)binv_hi);
    const int n = a.t.n;


    unsigned long long c_evals = 0, c_gath = 0, c_pairs = 0;
    unsigned c_search = 0, c_iters = 0;

    auto hand_back = [&](int i, int why) {     // redo target i on the generic path
        if (lane == 0) {
            a.worklist[atomicAdd(a.nwork, 1)] = i;
            atomicAdd(&a.counters[4 + why], 1ull);
        }
    };

    for (;;) {
        __syncthreads();              // previous tile fully consumed
        if (threadIdx.x == 0) { s_misc[0] = atomicAdd(a.next, 1); s_misc[1] = 0; }
        __syncthreads();
        const int tile = tile_lo + s_misc[0];
        if (tile >= tile_hi) break;
        const int code = a.tile_ng[tile];
        if (code < 0) {               // whole tile to the generic path
            if (w == 0) {
                const int i = tile * 32 + lane;
                if (i < n) {
                    a.worklist[atomicAdd(a.nwork, 1)] = i;
                    atomicAdd(&a.counters[4], 1ull);
                }
            }
            continue;
        }
        const int nent = code & 0xfff, nruns = (code >> 12) & 0xffff;
        const int ng = (nruns + 3) >> 2;           // bit-matrix words
        const bool interior = (code >> 30) & 1;

        // ---- candidate runs of the tile (ascending, so particle order is preserved) ------
        if (w == 0) {
            const int *ent = a.tile_groups + (size_t)tile * TL_ENT;
            int base = 0;
            for (int e0 = 0; e0 < nent; e0 += 32) {
                const int e = e0 + lane;
                const int v = e < nent ? ent[e] : 0;
                const int c = __popc(v & 0xf);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(FULL_MASK, incl, o);
                    if (lane >= o) incl += u;
                }
                int off = base + incl - c;
                for (int b = 0; b < 4; b++)
                    if (v & (1 << b)) s_run[off++] = (v >> 4) * 32 + 8 * b;
                base += __shfl_sync(FULL_MASK, incl, 31);
            }
        }
        __syncthreads();

        // ---- phase 1: lane = target, bit matrix of the exact predicate at radius R_i ---
        {
            const int i = tile * 32 + lane;
            float xi = 0, yi = 0, zi = 0, R2 = -1.f;
            if (i < n) {
                const float4 pi = a.pw[i];
                xi = pi.x; yi = pi.y; zi = pi.z;
                const float R = tile_radius(a.hsml_in[i], pi.w, norm, a.bx.box_d);
                R2 = __fmul_rn(R, R);
            }
            const size_t n8 = ((size_t)n + 7) & ~(size_t)7;      // stride of the SoA copy
            if (interior) tile_phase1<true>(a.soa, a.soa + n8, a.soa + 2 * n8, s_run, s_mask, w, lane, ng, nruns, xi, yi, zi, R2, box);
            else tile_phase1<false>(a.soa, a.soa + n8, a.soa + 2 * n8, s_run, s_mask, w, lane, ng, nruns, xi, yi, zi, R2, box);
        }
        __syncthreads();

        // ---- phase 2: one warp per target ------------------------------------------------
        for (;;) {
            int tsel = 0;
            if (lane == 0) tsel = atomicAdd(&s_misc[1], 1);
            tsel = __shfl_sync(FULL_MASK, tsel, 0);
            if (tsel >= 32) break;
            const int i = tile * 32 + tsel;
            if (i >= n) continue;

            float4 pi = a.pw[i];
            pi.w = fabsf(pi.w);                    // the sign bit is the displaced-node flag
            const float hA = a.hsml_in[i];
            const float hB = (float)((double)hA * 1.23);                        // sph.c:51
            const float hi_w = __fmul_rn(pi.w, norm);                           // wvt_relax.c:124
            const float hsw = (float)((double)hi_w * a.bx.box_d);               // wvt_relax.c:135
            const float hA2 = __fmul_rn(hA, hA), hB2 = __fmul_rn(hB, hB), hsw2 = __fmul_rn(hsw, hsw);
            const double A = a.step * (double)hi_w;

            // (1) expand the bit row into a compact candidate-slot list.  Every lane owns the
            //     words lane, lane+32, ... of the row and writes their hits to one contiguous
            //     stretch (order inside the list is irrelevant: all sums below are trees).  The
            //     bits are visited by a uniform, fully unrolled loop with predicated stores --
            //     no divergence and no per-hit popc/ffs on the XU pipe.  (A loop over the SET
            //     bits of each lane's words spent 17 % of the kernel's instructions here at 6.8
            //     active lanes; ranking lane = bit with two popc per word was no faster.)
            constexpr int NW = TL_WORDS / 32;
            unsigned wd[NW];
            int c = 0;
#pragma unroll
            for (int j = 0; j < NW; j++) {
                const int q = j * 32 + lane;
                wd[j] = q < ng ? s_mask[q * TL_MSTRIDE + tsel] : 0u;
                c += __popc(wd[j]);
            }
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += v;
            }
            const int nU = __shfl_sync(FULL_MASK, incl, 31);
            if (nU <= TL_UCAP) {
                unsigned short *out = ul + (incl - c);
#pragma unroll
                for (int j = 0; j < NW; j++) {
                    if (j * 32 >= ng) break;               // warp-uniform
                    const unsigned word = wd[j];
                    const int sbase = (j * 32 + lane) * 32;
#pragma unroll
                    for (int b = 0; b < 32; b++)
                        if (word & (1u << b)) *out++ = (unsigned short)(sbase + b);
                }
            }
            if (nU > TL_UCAP) { hand_back(i, 1); continue; }
            __syncwarp();

            // (2) classify every hit, build the separation list, sum the displacement
            int cntA = 0, cntB = 0, cntW = 0;      // cntA, cntW: per-lane until reduced below
            bool ranges_ok = true;                 // every separation fit for the hoisted divide
            float sx = 0, sy = 0, sz = 0;
            const float Af = (float)A;
            int qn = 0;                            // queued displacement partners (candidate slots)
            auto wvt_batch = [&](int cntq) {       // wvt_relax.c:137-170 for wq[0 .. cntq)
                if (lane < cntq) {
                    const int sl = wq[lane];
                    const float4 pj = a.pw[s_run[sl >> 3] + (sl & 7)];
                    float tx, ty, tz;
                    if (wvt_pair_fast(pi, pj, hi_w, norm, Af, binv_hi, binv_lo, tx, ty, tz)) {
                        sx += tx; sy += ty; sz += tz;
                        c_pairs++;
                    }
                }
            };
            for (int base = 0; base < nU; base += 32) {
                const int k = base + lane;
                const bool live = k < nU;
                const int slot = live ? ul[k] : 0;
                __syncwarp();                       // hit list batch read before the region is reused
                const int gidx = s_run[slot >> 3] + (slot & 7);
                const float4 pj = a.pw[gidx];
                const float xj = pj.x, yj = pj.y, zj = pj.z;
                float dx = fabsf(__fsub_rn(pi.x, xj)), dy = fabsf(__fsub_rn(pi.y, yj)),
                      dz = fabsf(__fsub_rn(pi.z, zj));
                if (!interior) {       // interior tile: no hit can be a periodic image
                    if (dx > boxhalf) dx = __fsub_rn(dx, box);
                    if (dy > boxhalf) dy = __fsub_rn(dy, box);
                    if (dz > boxhalf) dz = __fsub_rn(dz, box);
                }
                const float r2 = sq3_nofma(dx, dy, dz);                          // tree.c:88
                bool inA = live && r2 < hA2, inB = live && r2 < hB2, inW = live && r2 < hsw2;
                if ((inB | inW) && df_flagged(pj.w)) {
                    // underneath a displaced reference node (defect.cuh): Find_ngb_tree finds it
                    // only if every node of its path opens at the radius of that search
                    const float4 *path = a.dnodes + a.dmap[gidx];
                    if (inA) inA = defect_open(path, pi.x, pi.y, pi.z, hA, box, boxhalf);
                    if (inB) inB = inA || defect_open(path, pi.x, pi.y, pi.z, hB, box, boxhalf);
                    if (inW) inW = defect_open(path, pi.x, pi.y, pi.z, hsw, box, boxhalf);
                }
                const unsigned mB = __ballot_sync(FULL_MASK, inB);
                if (MODE & MODE_DENSITY) {
                    if (inB) {
                        double r = interior
                            ? pair_r<false>(pi.x, pi.y, pi.z, xj, yj, zj, a.bx.box_d, a.bx.boxhalf_d)
                            : pair_r<true>(pi.x, pi.y, pi.z, xj, yj, zj, a.bx.box_d, a.bx.boxhalf_d);
                        ranges_ok &= fdiv_range_ok(r);
                        if (!inA) r = __longlong_as_double(__double_as_longlong(r) | (1ll << 63));
                        const int pos = cntB + __popc(mB & lt);
                        if (pos < TL_LCAP) rl[pos] = r;
                    }
                }
                cntA += inA;
                cntB += __popc(mB);
                cntW += inW;
                if (MODE & MODE_WVT) {
                    // nU <= TL_CAP < NGBMAX: the list cut of tree.c:91 cannot bite here.
                    // Only about half of the hits are displacement partners, so they are queued
                    // and evaluated 32 at a time with full lanes.
                    const bool useW = inW && gidx != i;                          // wvt_relax.c:141
                    const unsigned mW = __ballot_sync(FULL_MASK, useW);
                    if (useW) wq[qn + __popc(mW & lt)] = (unsigned short)slot;
                    qn += __popc(mW);
                    __syncwarp();
                    if (qn >= 32) {
                        wvt_batch(32);
                        qn -= 32;
                        __syncwarp();
                        const unsigned short rest = wq[32 + lane];
                        __syncwarp();
                        wq[lane] = rest;
                        __syncwarp();
                    }
                }
            }
            if ((MODE & MODE_WVT) && qn > 0) wvt_batch(qn);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                cntA += __shfl_xor_sync(FULL_MASK, cntA, o);
                cntW += __shfl_xor_sync(FULL_MASK, cntW, o);
            }
            ranges_ok = __all_sync(FULL_MASK, ranges_ok);
            __syncwarp();

            float h = hA, rho = 0, drho = 0;
            bool ok = true;
            int cnt = 0;
            if (MODE & MODE_DENSITY) {
                // (3) the outer loop of sph.c:36-64, as far as the two prepared radii carry it
                if (cntA >= TG_DESNNGB) {                    // first search succeeds: Hsml list
                    if (cntA != cntB) {                      // drop the entries beyond Hsml, in place
                        int kept = 0;
                        const int tot = min(cntB, TL_LCAP);
                        for (int base = 0; base < tot; base += 32) {
                            const int k = base + lane;
                            const double r = k < tot ? rl[k] : -1.0;
                            const bool keep = __double_as_longlong(r) >= 0;
                            const unsigned mk = __ballot_sync(FULL_MASK, keep);
                            __syncwarp();
                            if (keep) rl[kept + __popc(mk & lt)] = r;
                            kept += __popc(mk);
                            __syncwarp();
                        }
                    }
                    cnt = cntA; h = hA; c_search += 1;
                    ok = cntB <= TL_LCAP;            // else list entries were dropped
                } else if (cntB >= TG_DESNNGB && cntB <= TL_LCAP) {   // second search, 1.23*Hsml
                    cnt = cntB; h = hB; c_search += 2;
                } else ok = false;                           // a third search: generic path
                int why = 3;                                 // a third search / list overflow
                if (ok) {
                    __syncwarp();
                    ok = find_hsml<(MODE & MODE_EXACT) != 0>(a, L, cnt, h, rho, drho, c_evals, c_iters, ranges_ok);
                    why = 4;                                 // no convergence on the frozen list
                }
                if (!ok) { hand_back(i, why); continue; }
            }
            c_search += (MODE & MODE_WVT) ? 1 : 0;
            c_gath += max(cnt, cntW);

            // (4) results
            double dsx = 0, dsy = 0, dsz = 0;
            if (MODE & MODE_WVT) { dsx = warp_sum((double)sx); dsy = warp_sum((double)sy); dsz = warp_sum((double)sz); }
            if (lane == 0) {
                if (MODE & MODE_DENSITY) {                                       // sph.c:66-70
                    const float q = __fmul_rn(__fdiv_rn(h, __fmul_rn(3.f, rho)), drho);
                    a.hsml_out[i] = h;
                    a.rho_out[i] = rho;
                    a.varh_out[i] = (float)(1.0 / (double)__fadd_rn(1.f, q));
                }
                if (MODE & MODE_WVT) {
                    a.delta[i] = (float)dsx;
                    a.delta[n + i] = (float)dsy;
                    a.delta[2 * (size_t)n + i] = (float)dsz;
                }
            }
        }
    }

    const unsigned long long pairs = warp_sum_u64(c_pairs);
    if (lane == 0) {
        atomicAdd(&a.counters[0], c_evals + pairs);
        atomicAdd(&a.counters[1], c_gath);
        atomicAdd(&a.counters[2], (unsigned long long)c_search);
        atomicAdd(&a.counters[3], (unsigned long long)c_iters);
    }
}
