// tile_fast.cuh -- TG_FAST: the tile sweep with FP32 kernel arithmetic.
//
// The exact modes (tile.cuh, sph.cuh) replay the reference's mixed precision operation for
// operation: FP64 separations, W and W' as FP64 polynomial chains rounded to float, FP64 sums.
// ncu showed what that costs on B200: 8.3 k warp instructions per target, a third of them on
// the FP64 / conversion pipes, and the sweep is issue bound (profiles/r01_final_*).  The
// north-star tolerance for rho, hsml and the displacement is 1e-5, not bit equality, so this
// mode keeps everything that DECIDES something exactly as the reference has it and evaluates
// the smooth parts in float:
//
//   kept exact   the neighbour SETS: float, FMA-free predicate of tree.c:67-88 against Hsml,
//                1.23*Hsml and the WVT radius, displaced-node open tests (defect.cuh), the
//                frozen list of sph.c:40,56, the outer retry loop (sph.c:36-64), the Newton /
//                bisection control flow of sph.c:156-195 in FP64 on FP64 totals.
//   float        r = sqrt(r2) from the predicate's own float r2 (one Newton step on MUFU.RSQ:
//                ~1 ulp) instead of an FP64 sqrt of FP64 differences; u = r * (1/h);
//                w(u) = (1-u)^8 (1+8u+25u^2+32u^3) and v(u) = u^2 (1-u)^7 (16u^2+7u+1) as
//                packed FP32 polynomials (FMUL2/FFMA2, two list entries per instruction),
//                per-lane float partial sums, FP64 tree over the lanes.
//   algebra      the sums are dimensionless (sph.c:149-153 with the constants taken out):
//                  wkNgb = 4pi/3 * kW * Sw,  rho = m kW/h^3 * Sw,
//                  dRho/dh = -m kW/h^4 (3 Sw - 22 Sv)   =>   omega = 22 Sv / (3 Sw),
//                so one Find_hsml iteration needs no per-entry h^3, h^4 factors and one FP64
//                divide.
//   displacement evaluated in place while the hit is classified (same r2, same gather):
//                u = r / (0.5 (h_i + h_j) Boxsize), clamped to 1 (the skip of
//                wvt_relax.c:160 is where W = 0), addend = step h_i W / r * d in float.
//
// Per-value error ~1e-7 (u) amplified by |W'/W(0)| <= 2.75: ~3e-7 W(0) per entry, random
// over ~400 entries => ~1e-7 on the sums, ~3e-8 on hsml per Newton step.  What can differ
// from the reference by more is a DECISION taken within that noise of its threshold
// (|wkNgb - 295| < 0.05, sph.c:161): one more or one fewer Newton step, i.e. up to the
// reference's own convergence slack of ~5e-5 in hsml, for ~1e-4 of the particles.
// tests/test_gpu_fast.py states this as a distribution against the compiled reference.
//
// Everything outside the envelope (cold start, list overflow, a third search, no convergence
// on the frozen list) goes to the same work list as in tile.cuh and is redone by the exact
// generic sweep.
#pragma once
#include <type_traits>
#include "tile.cuh"

// Block shape.  The kernel needs 72 registers, so 28 warps fit an SM; with the bit matrix out of
// shared memory (below) the warps of a block share nothing but the run list, and small blocks
// keep more tiles in flight: measured at 10 M (sweep incl. tile walk and hand-backs) 4 x 7:
// 50.5 ms, 3 x 9: 50.6, 2 x 14: 50.2 (shorter hit lists), 7 x 4: 51.1, 5 x 5: 52.2, 6 x 4 at 80
// registers: 52.9, 8 x 4 at 64 registers: 53.4, 12 x 2: 56.5, 24 x 1: 64.6.
#ifndef TF_WARPS
#define TF_WARPS 4
#endif
#ifndef TF_BLOCKS
#define TF_BLOCKS 7
#endif

// Tiles in flight per block.  The bit matrix of a tile lives in a block-private stretch of
// global memory (L2 resident: TL_WORDS x 32 words = 24 KB per tile), so that TF_NB of them cost no
// shared memory and
// the warps of a block need no common barrier: a warp that finds no target left in tile t goes
// on to the run list / phase 1 / targets of tile t+1 while the others finish (see the kernel).
#ifndef TF_NB
#define TF_NB 2
#endif
// hits within R_i / density list entries per target (a multiple of 64; < NGBMAX, so the list cut
// of tree.c:91-92 cannot bite).  With the bit matrix out of shared memory the lists and the run lists
// (TF_NB x TL_RUNS ints) are what is left: 7 blocks x 31.6 KB fill the SM, so 768 runs per tile cost 64
// of the 832 list entries that fitted beside 512.
#ifndef TF_CAP
#define TF_CAP 768
#endif
#ifndef TF_P1_CHUNK
#define TF_P1_CHUNK 4
#endif
#ifndef TF_POLL_NS
#define TF_POLL_NS 40                    // a waiting warp looks at its control word this often
#endif
#define TF_MASK_WORDS (TL_WORDS * 32)    // per tile: [word][target]

// Shared memory: run lists of the tiles in flight, per warp a hit list (particle indices) and a
// float separation list, the tile pipeline's control words, per-warp statistics.
#define TF_OFF_RUN 0
#define TF_OFF_UL (TF_OFF_RUN + TF_NB * TL_RUNS * 4)
#define TF_OFF_RL (TF_OFF_UL + TF_WARPS * TF_CAP * 4)
#define TF_OFF_MISC (TF_OFF_RL + TF_WARPS * TF_CAP * 4)
#define TF_OFF_CNT (TF_OFF_MISC + 256)
#define TF_OFF_FL (TF_OFF_CNT + TF_WARPS * 32)
#define TF_FLAG_CAP 32                   // per warp: hits underneath displaced nodes, kept for the second pass
#define TF_SMEM (TF_OFF_FL + TF_WARPS * TF_FLAG_CAP * 4)

#define TF_KW (1365.0 / (64 * K_PI))

static __device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
static __device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 1 / x to ~4e-15: MUFU.RCP on the float of x and one Newton step in double.  For the FP64
// quotients of the Find_hsml control flow (a correctly rounded FP64 divide is ~25 instructions
// on every lane, and TG_FAST needs those quotients to ~1e-9 at most).
static __device__ __forceinline__ double rcp_fast(double x)
{
    const double y = (double)rcp_approx((float)x);
    return fma(y, fma(-x, y, 1.0), y);
}

// 256-bit read-only load (LDG.E.256 on sm_100a): the 32-byte pair record of the packed phase 2
// and the eight-float rows of phase 1 in one request instead of two.
struct __align__(32) u256 { unsigned long long a, b, c, d; };
static __device__ __forceinline__ u256 ldg256(const void *p)
{
    u256 r;
#ifdef TF_NO_LDG256
    const ulonglong2 lo = __ldg((const ulonglong2 *)p), hi = __ldg((const ulonglong2 *)p + 1);
    r.a = lo.x; r.b = lo.y; r.c = hi.x; r.d = hi.y;
#else
    asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
#endif
    return r;
}

// Closest-image separation xi - xj -+ Boxsize, accurate in float.  The reference's neighbour
// predicate works on fl(xi - xj) -+ Boxsize (tree.c:67-78; reproduced for the predicate), but
// Find_hsml and the displacement take the separation from FP64 differences (sph.c:111-126).  For
// a pair across the periodic boundary fl(xi - xj) has already lost the low bits (error 2^-24 of
// Boxsize, 1e-5 of a separation of Boxsize/300); going through xi - Boxsize (exact for xi >
// Boxsize/2, Sterbenz) or xj - Boxsize keeps the float result within an ulp of the FP64 one.
static __device__ __forceinline__ float wrap_sep(float xi, float xj, float box, float boxhalf)
{
    const float d = xi - xj;
    if (d > boxhalf) return (xi - box) - xj;
    if (d < -boxhalf) return xi - (xj - box);
    return d;
}

// Sum of two doubles per lane over the warp with 6 shuffle steps instead of 10: the first
// step trades one quantity for the other between lane pairs, the last one trades the totals.
static __device__ __forceinline__ void warp_sum2(double &a, double &b)
{
    const bool odd = lane_id() & 1;
    const double give = odd ? a : b, keep = odd ? b : a;
    double v = keep + __shfl_xor_sync(FULL_MASK, give, 1);      // even lanes: a, odd lanes: b
#pragma unroll
    for (int o = 2; o <= 16; o <<= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    const double other = __shfl_xor_sync(FULL_MASK, v, 1);
    a = odd ? other : v;
    b = odd ? v : other;
}

// sph.c:80-214 on the frozen float list r[0 .. cnt).  Same control flow as find_hsml
// (sph.cuh); the per-entry arithmetic is the packed FP32 evaluation described above.
// u = r * inv with inv ~ 1/h from MUFU.RCP; its relative error e (exactly inv*h - 1, one FMA)
// shifts every u alike, which is a first-order correction of the sum:
//   Sw(u) = Sw(u (1+e)) - e * sum u w'(u) = Sw(u (1+e)) + 22 e Sv      (u w'(u) = -22 v(u)).
static __device__ __forceinline__ bool find_hsml_fast(const SweepArgs &a, const float *r, int cnt,
                                                      float &h_io, float &rho_out, float &drho_out,
                                                      unsigned &evals, unsigned &iters)
{
    const int lane = lane_id();
    const double mpart = a.bx.mpart;
    const double KN = K_FOURPITHIRD * TF_KW;

    double upper = (double)h_io * K_SQRT3, lower = 0;
    double hs = h_io, Sw = 0, Sv = 0;
    const int wshift = __clz(cnt) - 1;                                  // cnt < 2^(32 - clz) => cnt << wshift < 2^31
    const float wscale = __int_as_float((127 + wshift) << 23);          // 2^wshift
    const double wunscale = __longlong_as_double((long long)(1023 - wshift) << 52);
    int it = 0;
    bool done = false;

    const f32x2 one2 = pack2(1.f, 1.f);
    const f32x2 c32 = pack2(32.f, 32.f), c25 = pack2(25.f, 25.f), c8 = pack2(8.f, 8.f);
    const f32x2 c16 = pack2(16.f, 16.f), c7 = pack2(7.f, 7.f);

    for (;;) {
        const float hf = (float)hs;
        const float inv = rcp_approx(hf);
        const float einv = fmaf(inv, hf, -1.f);
        const f32x2 inv2 = pack2(inv, inv);
        f32x2 sw2 = pack2(0.f, 0.f), sv2 = pack2(0.f, 0.f);
        it++;

        auto pair = [&](float r0, float r1) {
            float u0, u1;
            unpack2(mul2(pack2(r0, r1), inv2), u0, u1);
            // r > hs contributes nothing (sph.c:135); clamping u is exact: w(1) = v(1) = 0
            const f32x2 u = pack2(fminf(u0, 1.f), fminf(u1, 1.f));
            const f32x2 t = sub2(one2, u);
            const f32x2 t2 = mul2(t, t), t3 = mul2(t2, t), t4 = mul2(t2, t2);
            const f32x2 t7 = mul2(t4, t3), t8 = mul2(t7, t);
            const f32x2 P = fma2(fma2(fma2(c32, u, c25), u, c8), u, one2);
            const f32x2 Q = fma2(fma2(c16, u, c7), u, one2);
            sw2 = fma2(t8, P, sw2);
            sv2 = fma2(mul2(u, u), mul2(t7, Q), sv2);
        };
        // the caller padded the list to a multiple of 64 with entries far outside (u clamps to 1)
#pragma unroll 2
        for (int k = lane; k < cnt; k += 64) pair(r[k], r[k + 32]);
        float w0, w1, v0, v1;
        unpack2(sw2, w0, w1);
        unpack2(sv2, v0, v1);
#ifdef TF_FP64_LANE_SUM
        Sw = (double)w0 + (double)w1;
        Sv = (double)v0 + (double)v1;
        warp_sum2(Sw, Sv);
#else
        // across the lanes in fixed point: one REDUX each instead of six 64-bit shuffle steps, and
        // the result does not depend on the order.  0 <= w <= 1 and 0 <= v < 1/16 per entry, so
        // cnt * 2^wshift < 2^31; the resolution (2.4e-7 for 300 entries, per lane) is below the
        // rounding of the float partial sums it converts (~1e-6 at a partial sum of ~10).
        Sw = (double)__reduce_add_sync(FULL_MASK, __float2int_rn((w0 + w1) * wscale)) * wunscale;
        Sv = (double)__reduce_add_sync(FULL_MASK, __float2int_rn((v0 + v1) * (16.f * wscale))) * (0.0625 * wunscale);
#endif
        Sw = fma(22.0 * (double)einv, Sv, Sw);
        evals += cnt;

        // sph.c:149 with the reference's own factors: the kernels see the float h, p3(hsml) is
        // the double -- their ratio hs^3 / (float)h^3 = 1 + g, |g| < 3e-7, is part of what the
        // iteration converges on.  g = (hs^3 - h3f) / h3f; the float inv^3 is plenty for 1/h3f.
        const float h3f = __fmul_rn(__fmul_rn(hf, hf), hf);
        const double g = (hs * hs * hs - (double)h3f) * (double)(inv * inv * inv);
        const double wkNgb = fma(KN * Sw, g, KN * Sw);

        if (it > 128) break;                                            // sph.c:156
        const double dev = fabs(wkNgb - TG_DESNNGB);
        if (dev < 0.05) { done = true; break; }                         // sph.c:161
        if (fabs(upper - lower) < 1e-4) { hs *= 1.26; break; }          // sph.c:168
        if (dev < 0.5 * TG_DESNNGB) {                                   // Newton-Raphson
            // omega = 1 + dRhodHsml * hs / (3 rho) = 22 Sv / (3 Sw)  (m, c1 and hs cancel)
            // fac = 1 - (wkNgb - 295) / (3 wkNgb omega)
            double fac = 1 - (wkNgb - TG_DESNNGB) * Sw * rcp_fast(wkNgb * 22.0 * Sv);
            fac = fmin(1.24, fac);
            fac = fmax(1 / 1.24, fac);
            hs *= fac;
        } else {                                                        // bisection in h^3
            if (wkNgb > TG_DESNNGB) upper = hs;
            if (wkNgb < TG_DESNNGB) lower = hs;
            // pow(x, 1/3) (sph.c:194) to ~1e-14: float cube root + one Newton step in double.  Every
            // target that needs the second search (45 % of them) starts with this step, and CUDA's
            // double pow is ~170 instructions
            const double x = 0.5 * (lower * lower * lower + upper * upper * upper);
            const float y0 = cbrtf((float)x);
            const double y = (double)y0;
            hs = y - (y * y * y - x) * (double)(1.f / (3.f * y0 * y0));
        }
    }
    iters += it;

    h_io = (float)hs;
    if (done) {                                                         // sph.c:151-153, 202-210
        const float hf = (float)hs;
        const double c1 = TF_KW * rcp_fast((double)__fmul_rn(__fmul_rn(hf, hf), hf));
        const double rho = mpart * c1 * Sw;
        const double drho = -mpart * c1 * (3.0 * Sw - 22.0 * Sv) * rcp_fast(hs);
        rho_out = (float)rho;
        drho_out = (float)drho;
        const float w0 = (float)c1;                                     // sph_kernel_WC6(0, hsml)
        const double bias = a.bias_const * mpart * (double)w0;
        rho_out = (float)((double)rho_out + bias);
    }
    return done;
}

// Phase 1 for the bit-matrix words [q0, q1) of this lane's target (four runs of eight candidates
// per word): tile.cuh's tile_phase1 with the words handed out by ticket instead of by warp
// number, and software-pipelined -- the three rows of run r + 1 are in flight while run r is
// tested (ncu: the first use of the rows held 19 % of the kernel's stall samples when every run
// waited for its own loads).
template <bool INTERIOR>
static __device__ __forceinline__ void tf_phase1_words(const float *__restrict__ sx, const float *__restrict__ sy,
                                                       const float *__restrict__ sz, const int *s_run, int q0, int q1,
                                                       int nruns, f32x2 xi2, f32x2 yi2, f32x2 zi2, float R2p, float box,
                                                       unsigned *__restrict__ gmask_lane)
{
    const float ibox = 1.f / box;
    const f32x2 ib2 = pack2(ibox, ibox), mg2 = pack2(12582912.f, 12582912.f);
    const f32x2 nb2 = pack2(-box, -box);
    const int rend = min(4 * q1, nruns);     // runs [4 q0, rend)
    auto test_run = [&](const u256 &X, const u256 &Y, const u256 &Z) -> unsigned {
        unsigned sub = 0;
        auto test2 = [&](f32x2 A, f32x2 B, f32x2 C, unsigned b0, unsigned b1) {
            f32x2 dx = sub2(xi2, A), dy = sub2(yi2, B), dz = sub2(zi2, C);
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
        test2(X.a, Y.a, Z.a, 1u, 2u);
        test2(X.b, Y.b, Z.b, 4u, 8u);
        test2(X.c, Y.c, Z.c, 16u, 32u);
        test2(X.d, Y.d, Z.d, 64u, 128u);
        return sub;                          // the pad of a short last run is far away
    };
    // same address in every lane: broadcast loads of whole rows (a run starts at a multiple of
    // 8, so the rows are 32-byte aligned).  Two runs per trip on two row sets A and B that take
    // turns, so that the hand-over of the rows in flight costs no register moves.
    int r = 4 * q0;
    int f = s_run[r];
    u256 AX = ldg256(sx + f), AY = ldg256(sy + f), AZ = ldg256(sz + f);
    unsigned word = 0;
#pragma unroll 1
    for (; r + 1 < rend; r += 2) {           // r is even
        f = s_run[r + 1];
        const u256 BX = ldg256(sx + f), BY = ldg256(sy + f), BZ = ldg256(sz + f);
        const int sh = 8 * (r & 3);          // 0 or 16
        word |= test_run(AX, AY, AZ) << sh;
        f = s_run[min(r + 2, rend - 1)];     // (past the end: re-read, unused)
        AX = ldg256(sx + f); AY = ldg256(sy + f); AZ = ldg256(sz + f);
        word |= test_run(BX, BY, BZ) << (sh + 8);
        if (r & 2) {                         // runs 4q .. 4q + 3 done
            gmask_lane[(r >> 2) * 32] = word;
            word = 0;
        }
    }
    if (r < rend) word |= test_run(AX, AY, AZ) << (8 * (r & 3));       // odd number of runs
    if (rend & 3) gmask_lane[((rend - 1) >> 2) * 32] = word;            // the short last word
}

template <int MODE>
__global__ void __launch_bounds__(TF_WARPS * 32, TF_BLOCKS) k_sweep_tile_fast(const SweepArgs a, int tile_lo,
                                                                        int tile_hi)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *s_cnt = (unsigned long long *)(smem + TF_OFF_CNT);   // per warp: evals, gathered, searches, iters

    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1;
    int *ul = (int *)(smem + TF_OFF_UL) + w * TF_CAP;          // hit list: particle indices
    float *rl = (float *)(smem + TF_OFF_RL) + w * TF_CAP;
    int *fl = (int *)(smem + TF_OFF_FL) + w * TF_FLAG_CAP;     // flagged hits of the current target
    unsigned long long *cw = s_cnt + w * 4;
    if (lane < 4) cw[lane] = 0;

    const float norm = (float)pow(TG_DESNNGB / *a.vsum / K_FOURPITHIRD, 1.0 / 3.0);   // wvt_relax.c:120
    const float box = a.bx.box_f, boxhalf = a.bx.boxhalf_f;
    const int n = a.t.n;

    auto hand_back = [&](int i, int why) {     // redo target i on the generic (exact) path
        if (lane == 0) {
            a.worklist[atomicAdd(a.nwork, 1)] = i;
            atomicAdd(&a.counters[4 + why], 1ull);
        }
    };

    // ---- the tile pipeline ---------------------------------------------------------------
    // Every warp walks the block's tile sequence t = 0, 1, 2, ... on its own; tile t uses buffer
    // t % TF_NB (run list in shared memory, bit matrix in global).  Per buffer:
    //   CLAIM   t+1 of the last tile claimed: the FIRST warp to arrive at t fetches the tile and
    //           builds its run list -- after every warp has LEFT the tile that used the buffer
    //           before (LEFT is cumulative: TF_WARPS per tile) -- then publishes READY = t+1;
    //   P1NEXT / P1DONE   phase 1 in tickets of TF_P1_CHUNK bit-matrix words, taken by whichever
    //           warps are there; everybody waits until all words are written;
    //   TNEXT   phase 2: the targets, one ticket per target as before.
    // Fast warps therefore run ahead by up to TF_NB - 1 tiles instead of idling at a block
    // barrier while the last targets of a tile are finished (round 2 ncu: 9 % of the warp time).
    enum { C_TILE = 0, C_CODE, C_READY, C_CLAIM, C_LEFT, C_P1NEXT, C_P1DONE, C_TNEXT, C_STRIDE };
    static_assert(TF_NB * C_STRIDE * 4 <= 256, "control words");
    int *s_ctl = (int *)(smem + TF_OFF_MISC);
    if (threadIdx.x < TF_NB * C_STRIDE) s_ctl[threadIdx.x] = 0;
    __syncthreads();                           // the only block barrier of the kernel
    unsigned *gm_block = a.tile_mask + (size_t)blockIdx.x * (TF_NB * TF_MASK_WORDS);

    auto wait_for = [&](const int *p, int need) {            // lane 0 polls, the warp follows
        if (lane == 0)
            while (*(volatile const int *)p < need) __nanosleep(TF_POLL_NS);
        __syncwarp();
        __threadfence_block();
    };

    for (int t = 0;; t++) {
        const int b = t % TF_NB;
        int *ctl = s_ctl + b * C_STRIDE;
        int *s_run = (int *)(smem + TF_OFF_RUN) + b * TL_RUNS;       // first particle of each candidate run
        unsigned *gmask = gm_block + b * TF_MASK_WORDS;
        auto leave = [&]() {                   // this warp is done with the buffer
            __threadfence_block();
            __syncwarp();
            if (lane == 0) atomicAdd(&ctl[C_LEFT], 1);
        };

        int first = 0;
        if (lane == 0) first = atomicMax(&ctl[C_CLAIM], t + 1) <= t;
        first = __shfl_sync(FULL_MASK, first, 0);
        if (first) {
            wait_for(&ctl[C_LEFT], TF_WARPS * (t / TF_NB));
            int tl = 0;
            if (lane == 0) tl = atomicAdd(a.next, 1);
            tl = tile_lo + __shfl_sync(FULL_MASK, tl, 0);
            int cd = 0;
            if (tl >= tile_hi) tl = -1;
            else {
                cd = a.tile_ng[tl];
                if (cd < 0) {                  // whole tile to the generic path
                    const int i = tl * 32 + lane;
                    if (i < n) {
                        a.worklist[atomicAdd(a.nwork, 1)] = i;
                        atomicAdd(&a.counters[4], 1ull);
                    }
                } else {
                    // ---- candidate runs of the tile (ascending) ----------------------------------
                    const int nent = cd & 0xfff, nruns = (cd >> 12) & 0xffff;
                    const int *ent = a.tile_groups + (size_t)tl * TL_ENT;
                    int base = 0;
                    bool anyflag = false;          // a candidate box with a particle under a displaced node
                    for (int e0 = 0; e0 < nent; e0 += 32) {
                        const int e = e0 + lane;
                        const int v = e < nent ? ent[e] : 0;
                        if (e < nent) anyflag |= a.boxflag[v >> 4] != 0;
                        const int c = __popc(v & 0xf);
                        int incl = c;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int u = __shfl_up_sync(FULL_MASK, incl, o);
                            if (lane >= o) incl += u;
                        }
                        int off = base + incl - c;
                        for (int bb = 0; bb < 4; bb++)
                            if (v & (1 << bb)) s_run[off++] = (v >> 4) * 32 + 8 * bb;
                        base += __shfl_sync(FULL_MASK, incl, 31);
                    }
                    // pad to a whole word of four runs (the expansion reads them four at a time)
                    const int ngp = (nruns + 3) >> 2;
                    if (lane < 4 && base + lane < 4 * ngp) s_run[base + lane] = 0;
                    if (__any_sync(FULL_MASK, anyflag)) cd |= 1 << 28;
                }
            }
            if (lane == 0) { ctl[C_TILE] = tl; ctl[C_CODE] = cd; ctl[C_P1NEXT] = 0; ctl[C_P1DONE] = 0; ctl[C_TNEXT] = 0; }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *(volatile int *)&ctl[C_READY] = t + 1;
        } else
            wait_for(&ctl[C_READY], t + 1);

        const int tile = *(volatile int *)&ctl[C_TILE];
        if (tile < 0) break;                   // no tiles left: every warp sees this at the same t
        const int code = *(volatile int *)&ctl[C_CODE];
        if (code < 0) { leave(); continue; }
        const int nruns = (code >> 12) & 0xffff;
        const int ng = (nruns + 3) >> 2;           // bit-matrix words
        const bool interior = (code >> 30) & 1;
        const bool tile_flags = (code >> 28) & 1;

        // ---- phase 1: lane = target, superset bit matrix at radius R_i (tile.cuh) --------
        {
            bool loaded = false;
            float R2p = -1.f;
            f32x2 xi2 = pack2(0.f, 0.f), yi2 = xi2, zi2 = xi2;
            const size_t n8 = ((size_t)n + 7) & ~(size_t)7;      // stride of the SoA copy
            for (;;) {
                int q0 = 0;
                if (lane == 0) q0 = atomicAdd(&ctl[C_P1NEXT], TF_P1_CHUNK);
                q0 = __shfl_sync(FULL_MASK, q0, 0);
                if (q0 >= ng) break;
                if (!loaded) {
                    loaded = true;
                    const int i = tile * 32 + lane;
                    if (i < n) {
                        const float4 pi = a.pw[i];
                        xi2 = pack2(pi.x, pi.x); yi2 = pack2(pi.y, pi.y); zi2 = pack2(pi.z, pi.z);
                        const float R = tile_radius(a.hsml_in[i], pi.w, norm, a.bx.box_d, (MODE & MODE_ROTA) ? 1 : 0);
                        R2p = __fmul_rn(R, R) * 1.000002f;
                    }
                }
                const int q1 = min(q0 + TF_P1_CHUNK, ng);
                if (interior) tf_phase1_words<true>(a.soa, a.soa + n8, a.soa + 2 * n8, s_run, q0, q1, nruns, xi2, yi2, zi2, R2p, box, gmask + lane);
                else tf_phase1_words<false>(a.soa, a.soa + n8, a.soa + 2 * n8, s_run, q0, q1, nruns, xi2, yi2, zi2, R2p, box, gmask + lane);
                // the words went to GLOBAL memory and are read back (through L2, __ldcg) by the other
                // warps of THIS block only: a block-scope fence before the block-scope signal is what
                // the memory model asks for (a device-scope fence here costs 0.5 ms per step at 10 M)
#ifdef TF_GPU_FENCE
                __threadfence();
#else
                __threadfence_block();
#endif
                __syncwarp();
                if (lane == 0) atomicAdd(&ctl[C_P1DONE], q1 - q0);
            }
            wait_for(&ctl[C_P1DONE], ng);
        }

        // ---- phase 2: one warp per target ------------------------------------------------
        for (;;) {
            int tsel = 0;
            if (lane == 0) tsel = atomicAdd(&ctl[C_TNEXT], 1);
            tsel = __shfl_sync(FULL_MASK, tsel, 0);
            if (tsel >= 32) break;
            const int i = tile * 32 + tsel;
            if (i >= n) continue;
            // the target's own record and Hsml: in flight while the row is expanded
            const float4 pi_raw = a.pw[i];
            const float hA_in = a.hsml_in[i];

            // (1) expand the bit row into a compact list of particle indices.  Every lane owns
            //     the words lane, lane+32, ... of the row and writes their hits to one contiguous
            //     stretch (order inside the list is irrelevant: all sums below are trees); uniform,
            //     fully unrolled bit loop with predicated stores (tile.cuh).
            //     PAIRS (density / displacement modes): the list holds PAIRS of consecutive particles
            //     (2p, 2p+1) with at least one hit -- the hot loop below evaluates both halves at once
            //     with packed FP32 instructions on the pair-interleaved copy of the positions; the
            //     half that is no hit fails the exact predicate by itself.  Half the bits to visit.
            constexpr bool PAIRS = !(MODE & MODE_ROTA);
            constexpr int NW = TL_WORDS / 32;
            unsigned wd[NW];
            int c = 0, chits = 0;
#pragma unroll
            for (int j = 0; j < NW; j++) {
                const int q = j * 32 + lane;
                unsigned word = q < ng ? __ldcg(gmask + q * 32 + tsel) : 0u;
                chits += __popc(word);
                if (PAIRS) word = (word | (word >> 1)) & 0x55555555u;
                wd[j] = word;
                c += __popc(word);
            }
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += v;
            }
            const int nU = __shfl_sync(FULL_MASK, incl, 31);
            // the separation list must hold every hit: the same bound as in tile.cuh
            if ((PAIRS ? __reduce_add_sync(FULL_MASK, chits) : nU) > TF_CAP) { hand_back(i, 1); continue; }
            {
                int *out = ul + (incl - c);
#pragma unroll
                for (int j = 0; j < NW; j++) {
                    if (j * 32 >= ng) break;               // warp-uniform
                    const unsigned word = wd[j];
                    const int4 rs = *(const int4 *)(s_run + 4 * (j * 32 + lane));   // the word's four runs
                    if (PAIRS) {
#pragma unroll
                        for (int b = 0; b < 16; b++) {     // pair b of the word: run b / 4, pair b % 4 of it
                            const int first = b < 4 ? rs.x : (b < 8 ? rs.y : (b < 12 ? rs.z : rs.w));
                            if (word & (1u << (2 * b))) *out++ = (first >> 1) + (b & 3);
                        }
                    } else {
#pragma unroll
                        for (int b = 0; b < 32; b++) {
                            const int first = b < 8 ? rs.x : (b < 16 ? rs.y : (b < 24 ? rs.z : rs.w));
                            if (word & (1u << b)) *out++ = first + (b & 7);
                        }
                    }
                }
                // pad the last batch with the target itself (its pair), marked dead by the loop below
                if (nU + lane < ((nU + 31) & ~31))
                    ul[nU + lane] = PAIRS ? (interior && !tile_flags ? (int)((((size_t)n + 7) & ~(size_t)7) >> 1) : i >> 1) : i;
            }
            __syncwarp();
            // the gather of the first batch starts now, ahead of the per-target constants
            constexpr bool PAIRS_ = !(MODE & MODE_ROTA);
            int g0 = ul[lane];                       // (nU >= 1: the target itself is a hit)
            u256 P0;
            float4 p0;
            if (PAIRS_) P0 = ldg256((const u256 *)a.pwp + g0);
            else p0 = a.pw[g0];

            float4 pi = pi_raw;
            pi.w = fabsf(pi.w);                    // the sign bit is the displaced-node flag
            float hA2, hB2, hsw2, Afy, cn;
            float ax_i = 0, ay_i = 0, az_i = 0, inv_h = 0;        // rot(A): Apot_i, 1/Hsml
            if (MODE & MODE_ROTA) {
                const float hA = hA_in;
                hA2 = __fmul_rn(hA, hA); hB2 = hA2; hsw2 = 0; cn = 0;
                inv_h = __frcp_rn(hA);
                const float h2 = hA * hA;
                // sph.c:278-280: weight = -m / rho_i * W'(r, h) / r * VarHsmlFac, W' = kW / h^4 * -22 (...)
                Afy = (float)(-a.bx.mpart / (double)a.rho_in[i] * (double)a.varh_in[i] * (TF_KW * -22.0) / ((double)h2 * (double)h2));
                ax_i = a.apot[3 * (size_t)i]; ay_i = a.apot[3 * (size_t)i + 1]; az_i = a.apot[3 * (size_t)i + 2];
            } else {
                const float hA = hA_in;
                const float hB = (float)((double)hA * 1.23);                      // sph.c:51
                const float hi_w = __fmul_rn(pi.w, norm);                         // wvt_relax.c:124
                const float hsw = (float)((double)hi_w * a.bx.box_d);             // wvt_relax.c:135
                hA2 = __fmul_rn(hA, hA); hB2 = __fmul_rn(hB, hB); hsw2 = __fmul_rn(hsw, hsw);
                // wvt_relax.c:167: step * hsml_i * W, W = kW t^8 (...)
                Afy = (float)(a.step * (double)hi_w * TF_KW);
                cn = 0.5f * norm * box;            // pair h in length units: (w_i + w_j) * cn
                // keep these in registers: recomputing them per batch (what ptxas does under
                // register pressure) puts conversions and FP64 on the XU pipe inside the hot loop
                asm volatile("" : "+f"(hA2), "+f"(hB2), "+f"(hsw2), "+f"(Afy), "+f"(cn));
            }

            // (2) classify every hit exactly; separation list (A from the front, "1.23*Hsml
            //     only" from the back); displacement summed in place.  Hits underneath a displaced
            //     reference node (defect.cuh; ~1e-4 of the particles) are skipped by the hot loop
            //     and taken by a second pass with the open tests, so that the hot loop has no call.
            int cntA = 0, baseB = TF_CAP - 1, cntW = 0, npair = 0;   // cntW, npair: per-lane until reduced
            float sx = 0, sy = 0, sz = 0;
            bool sawflag = false;
            int nflag = 0;                         // flagged hits met by the packed pass (warp-uniform)
            auto batch = [&](const bool live, const int gidx, const float4 pj, auto slow_tag) {
                constexpr bool SLOW = decltype(slow_tag)::value;
                const float dx = __fsub_rn(pi.x, pj.x), dy = __fsub_rn(pi.y, pj.y), dz = __fsub_rn(pi.z, pj.z);
                float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
                if (!interior) {       // interior tile: no hit can be a periodic image
                    if (ax > boxhalf) ax = __fsub_rn(ax, box);
                    if (ay > boxhalf) ay = __fsub_rn(ay, box);
                    if (az > boxhalf) az = __fsub_rn(az, box);
                }
                float r2 = sq3_nofma(ax, ay, az);                                // tree.c:88
                // the value of r (and the direction of the displacement): accurate separations
                float ex = dx, ey = dy, ez = dz, r2a = r2;
                if (!interior) {
                    ex = wrap_sep(pi.x, pj.x, box, boxhalf); ey = wrap_sep(pi.y, pj.y, box, boxhalf);
                    ez = wrap_sep(pi.z, pj.z, box, boxhalf);
                    r2a = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
                }
                const bool flagged = df_flagged(pj.w);
                // dead lanes: the pad of the last batch, and the hits the other pass takes
                r2 = (live && flagged == SLOW) ? r2 : 3.0e38f;
                bool inA = r2 < hA2, inB = r2 < hB2, inW = r2 < hsw2;
                if (!SLOW) sawflag |= flagged;
                if (SLOW && (inB | inW)) {
                    const float4 *path = a.dnodes + a.dmap[gidx];
                    const float hA = a.hsml_in[i];
                    const float hB = (float)((double)hA * 1.23);
                    const float hsw = (float)((double)__fmul_rn(pi.w, norm) * a.bx.box_d);
                    if (inA) inA = defect_open(path, pi.x, pi.y, pi.z, hA, box, boxhalf);
                    if (inB) inB = inA || defect_open(path, pi.x, pi.y, pi.z, hB, box, boxhalf);
                    if (inW) inW = defect_open(path, pi.x, pi.y, pi.z, hsw, box, boxhalf);
                }
                // r = sqrt(r2): MUFU.RSQ and one Newton step; the target itself (r2 = 0) gives 0
                const float y = rsqrt_approx(fmaxf(r2a, 1e-35f));
                float r = r2a * y;
                r = fmaf(0.5f * y, fmaf(-r, r, r2a), r);
                if (MODE & MODE_DENSITY) {
                    const unsigned mA = __ballot_sync(FULL_MASK, inA);
                    const unsigned mBo = __ballot_sync(FULL_MASK, inB) & ~mA;
                    const int rank = __popc((inA ? mA : mBo) & lt);
                    if (inB) rl[inA ? cntA + rank : baseB - rank] = r;
                    cntA += __popc(mA);
                    baseB -= __popc(mBo);
                }
                cntW += inW;
                if (MODE & MODE_ROTA) {
                    // sph.c:236-295: B_i += weight * (d x (A_i - A_j)) over Find_ngb_tree(i, Hsml_i)
                    const float u = fminf(r * inv_h, 1.f);
                    const float t = 1.f - u, t2 = t * t, t3 = t2 * t, t4 = t2 * t2;
                    const float Q = fmaf(fmaf(16.f, u, 7.f), u, 1.f);
                    const bool use = inA && gidx != i;                           // sph.c:247 skips i itself
                    const float f = use ? (t4 * t3) * u * Q * (Afy * y) : 0.f;    // weight (dwk / r folded in)
                    const float dX = ex, dY = ey, dZ = ez;
                    const float dAx = ax_i - a.apot[3 * (size_t)gidx], dAy = ay_i - a.apot[3 * (size_t)gidx + 1],
                                dAz = az_i - a.apot[3 * (size_t)gidx + 2];
                    sx = fmaf(f, dZ * dAy - dY * dAz, sx);
                    sy = fmaf(f, dX * dAz - dZ * dAx, sy);
                    sz = fmaf(f, dY * dAx - dX * dAy, sz);
                    if (use) npair++;
                    cntW += inA;                                                 // gathered: the Hsml set
                }
                if (MODE & MODE_WVT) {
                    // wvt_relax.c:137-170; nU <= TF_CAP < NGBMAX: the list cut cannot bite
                    const float hp = (pi.w + fabsf(pj.w)) * cn;                   // :158, length units
                    const float u = fminf(r * rcp_approx(hp), 1.f);               // :160 skip <=> W = 0
                    const float t = 1.f - u, t2 = t * t, t4 = t2 * t2;
                    const float P = fmaf(fmaf(fmaf(32.f, u, 25.f), u, 8.f), u, 1.f);
                    const bool use = inW && gidx != i;                           // :141
                    const float f = use ? (t4 * t4) * P * (Afy * y) : 0.f;
                    sx = fmaf(f, ex, sx);
                    sy = fmaf(f, ey, sy);
                    sz = fmaf(f, ez, sz);
                    if (use && u < 1.f) npair++;
                }
            };
            // Two particles per lane: pair p = particles (2p, 2p+1) from the pair-interleaved copy
            // A = {x0, x1, y0, y1}, B = {z0, z1, w0, w1}; every float step below is one packed
            // instruction for both (each half an ordinary IEEE operation, f32x2.cuh).
            const f32x2 xi2 = pack2(pi.x, pi.x), yi2 = pack2(pi.y, pi.y), zi2 = pack2(pi.z, pi.z);
            f32x2 sx2 = pack2(0.f, 0.f), sy2 = sx2, sz2 = sx2;
            // INTERIOR_ / FLAGS_: tile-uniform facts the loop is compiled for -- no target of the tile
            // can see a periodic image (98.7 % of the tiles of the merger); some candidate box of the
            // tile holds a particle underneath a displaced node (~0.1 %: defect.cuh's box flags, read
            // by the builder).  The common instance has neither the wrap nor the flag handling.
            auto batch2 = [&](auto interior_tag, auto flags_tag, const int k, const int pidx, const ulonglong2 A,
                              const ulonglong2 B) {
                constexpr bool INTERIOR_ = decltype(interior_tag)::value, FLAGS_ = decltype(flags_tag)::value;
                f32x2 dx2 = sub2(xi2, A.x), dy2 = sub2(yi2, A.y), dz2 = sub2(zi2, B.x);
                f32x2 ex2 = dx2, ey2 = dy2, ez2 = dz2;    // accurate separations (wrap_sep): r, direction
                if (!INTERIOR_) {
                    // closest image for the PREDICATE: d - Boxsize * rint(d / Boxsize) -- one rounded
                    // subtraction, the value of tree.c:70-78 up to its sign; rint can only differ from
                    // the reference's compare for |d| within 1e-7 of Boxsize/2, which is no hit for
                    // either image (the tile walk only admits R < 0.49 Boxsize).
                    // ... and for the VALUE: (xi - k+ Boxsize) - (xj + k- Boxsize) with k = rint(...) in
                    // {-1, 0, 1}, k+ = (k k + k)/2, k- = (k - k k)/2: the intermediate is exact
                    const float ibox = 1.f / box;
                    const f32x2 ib2 = pack2(ibox, ibox), mg2 = pack2(12582912.f, 12582912.f), nb2 = pack2(-box, -box);
                    const f32x2 pb2 = pack2(box, box), half2 = pack2(0.5f, 0.5f);
                    auto wrap2 = [&](f32x2 &d, f32x2 &e, const f32x2 xi, const f32x2 xj) {
                        const f32x2 k = sub2(fma2(d, ib2, mg2), mg2);
                        const f32x2 kk = mul2(k, k);
                        e = sub2(fma2(mul2(add2(kk, k), half2), nb2, xi), fma2(mul2(sub2(k, kk), half2), pb2, xj));
                        d = fma2(k, nb2, d);
                    };
                    wrap2(dx2, ex2, xi2, A.x);
                    wrap2(dy2, ey2, yi2, A.y);
                    wrap2(dz2, ez2, zi2, B.x);
                }
                // tree.c:88 without FMA: packed products, SCALAR sums (ptxas would contract a packed
                // product feeding a packed add into one FFMA2)
                float px0, px1, py0, py1, pz0, pz1, w0, w1;
                unpack2(mul2(dx2, dx2), px0, px1);
                unpack2(mul2(dy2, dy2), py0, py1);
                unpack2(mul2(dz2, dz2), pz0, pz1);
                unpack2(B.y, w0, w1);
                float r20 = __fadd_rn(__fadd_rn(px0, py0), pz0), r21 = __fadd_rn(__fadd_rn(px1, py1), pz1);
                // the pad of the last batch: in an interior tile it is the ghost pair behind the last
                // real one (finite, far away: fails every radius by itself, weight 0); elsewhere the
                // target's own pair, which has to be masked (a far-away pad could wrap into reach)
                const bool live = INTERIOR_ || k < nU;
                const bool fl0 = FLAGS_ && df_flagged(w0), fl1 = FLAGS_ && df_flagged(w1);
                r20 = (live && !fl0) ? r20 : 3.0e38f;       // dead: the pad, hits of the second pass;
                r21 = (live && !fl1) ? r21 : 3.0e38f;       // (the pad of an odd n is finite and far away)
                if (FLAGS_ && __any_sync(FULL_MASK, live && (fl0 | fl1))) {   // rare: remember them for the second pass
                    const unsigned m0 = __ballot_sync(FULL_MASK, live && fl0), m1 = __ballot_sync(FULL_MASK, live && fl1);
                    const int p0 = nflag + __popc(m0 & lt), p1 = nflag + __popc(m0) + __popc(m1 & lt);
                    if (live && fl0 && p0 < TF_FLAG_CAP) fl[p0] = 2 * pidx;
                    if (live && fl1 && p1 < TF_FLAG_CAP) fl[p1] = 2 * pidx + 1;
                    nflag += __popc(m0) + __popc(m1);
                }
                const bool inA0 = r20 < hA2, inB0 = r20 < hB2, inW0 = r20 < hsw2;
                const bool inA1 = r21 < hA2, inB1 = r21 < hB2, inW1 = r21 < hsw2;
                // r = sqrt(r2): MUFU.RSQ and one Newton step
                float a20 = r20, a21 = r21;               // r^2 for the VALUE of r
                if (!INTERIOR_) unpack2(fma2(ez2, ez2, fma2(ey2, ey2, mul2(ex2, ex2))), a20, a21);
                const float y0 = rsqrt_approx(fmaxf(a20, 1e-35f)), y1 = rsqrt_approx(fmaxf(a21, 1e-35f));
                f32x2 r2v = pack2(r20, r21);
                if (!INTERIOR_) r2v = fma2(ez2, ez2, fma2(ey2, ey2, mul2(ex2, ex2)));
                const f32x2 y2 = pack2(y0, y1);
                f32x2 rr = mul2(r2v, y2);
                rr = fma2(mul2(y2, pack2(0.5f, 0.5f)), fma2(sub2(pack2(0.f, 0.f), rr), rr, r2v), rr);
                float r0, r1;
                unpack2(rr, r0, r1);
                if (MODE & MODE_DENSITY) {
                    const unsigned mA0 = __ballot_sync(FULL_MASK, inA0), mA1 = __ballot_sync(FULL_MASK, inA1);
                    const unsigned mBo0 = __ballot_sync(FULL_MASK, inB0) & ~mA0;
                    const unsigned mBo1 = __ballot_sync(FULL_MASK, inB1) & ~mA1;
                    const int nA0 = __popc(mA0), nB0 = __popc(mBo0);
                    const int rank0 = __popc((inA0 ? mA0 : mBo0) & lt), rank1 = __popc((inA1 ? mA1 : mBo1) & lt);
                    if (inB0) rl[inA0 ? cntA + rank0 : baseB - rank0] = r0;
                    if (inB1) rl[inA1 ? cntA + nA0 + rank1 : baseB - nB0 - rank1] = r1;
                    cntA += nA0 + __popc(mA1);
                    baseB -= nB0 + __popc(mBo1);
                }
                cntW += inW0 + inW1;
                if (MODE & MODE_WVT) {
                    // wvt_relax.c:137-170, both halves at once
                    const f32x2 aw2 = B.y & 0x7fffffff7fffffffull;                // |w|: sign = defect flag
                    const f32x2 hp2 = mul2(add2(pack2(pi.w, pi.w), aw2), pack2(cn, cn));   // :158, length units
                    float h0, h1, u0, u1;
                    unpack2(hp2, h0, h1);
                    unpack2(mul2(rr, pack2(rcp_approx(h0), rcp_approx(h1))), u0, u1);
                    const f32x2 u = pack2(fminf(u0, 1.f), fminf(u1, 1.f));        // :160 skip <=> W = 0
                    const f32x2 one2 = pack2(1.f, 1.f);
                    const f32x2 t = sub2(one2, u), t2 = mul2(t, t), t4 = mul2(t2, t2);
                    const f32x2 P = fma2(fma2(fma2(pack2(32.f, 32.f), u, pack2(25.f, 25.f)), u, pack2(8.f, 8.f)), u, one2);
                    float f0, f1;
                    unpack2(mul2(mul2(mul2(t4, t4), P), mul2(pack2(Afy, Afy), y2)), f0, f1);
                    const bool use0 = inW0 && 2 * pidx != i, use1 = inW1 && 2 * pidx + 1 != i;   // :141
                    const f32x2 f2 = pack2(use0 ? f0 : 0.f, use1 ? f1 : 0.f);
                    sx2 = fma2(f2, ex2, sx2);
                    sy2 = fma2(f2, ey2, sy2);
                    sz2 = fma2(f2, ez2, sz2);
                    // pairs with r < h_ij (wvt_relax.c:160): exactly those with a weight (t = 0 at u = 1)
                    npair += (use0 && f0 != 0.f) + (use1 && f1 != 0.f);
                }
            };
            if (PAIRS) {
                // the gather of the next batch is in flight while this one is evaluated (two batches
                // per trip, so the hand-over needs no register moves)
                const u256 *pp = (const u256 *)a.pwp;    // one 32-byte record per pair
                const int kend = nU + lane;
                auto pair_loop = [&](auto it, auto ft) {
                    for (int k = lane; k < kend; k += 64) {
                        const int k1 = k + 32 < kend ? k + 32 : k;       // (past the end: re-read, unused)
                        const int g1 = ul[k1];
                        const u256 P1 = ldg256(pp + g1);
                        batch2(it, ft, k, g0, make_ulonglong2(P0.a, P0.b), make_ulonglong2(P0.c, P0.d));
                        const int k2 = k + 64 < kend ? k + 64 : k;
                        g0 = ul[k2];
                        P0 = ldg256(pp + g0);
                        if (k + 32 < kend) batch2(it, ft, k + 32, g1, make_ulonglong2(P1.a, P1.b), make_ulonglong2(P1.c, P1.d));
                    }
                };
                // two instances only (code size: the kernel's instruction fetch stalls grew with a third):
                // the common one, and the general one, whose wrap is exact arithmetic with k = 0 inside
                if (interior && !tile_flags) pair_loop(std::true_type{}, std::false_type{});
                else pair_loop(std::false_type{}, std::true_type{});
                float a0, a1;
                unpack2(sx2, a0, a1); sx = a0 + a1;
                unpack2(sy2, a0, a1); sy = a0 + a1;
                unpack2(sz2, a0, a1); sz = a0 + a1;
                if (nflag > 0) {                          // the flagged hits, with the open tests
                    __syncwarp();
                    if (nflag <= TF_FLAG_CAP)
                        for (int k = lane; k < nflag + lane; k += 32) {
                            const int g = k < nflag ? fl[k] : i;
                            batch(k < nflag, g, a.pw[g], std::true_type{});
                        }
                    else                                  // more than the list holds: rescan every pair
                        for (int k = lane; k < nU + lane; k += 32) {
                            const int g = 2 * ul[k];
                            batch(k < nU, g, a.pw[g], std::true_type{});
                            batch(k < nU, g + 1, a.pw[g + 1], std::true_type{});
                        }
                }
            } else {
                const int kend = nU + lane;
                for (int k = lane; k < kend; k += 64) {
                    const int k1 = k + 32 < kend ? k + 32 : k;
                    const int g1 = ul[k1];
                    const float4 p1 = a.pw[g1];
                    batch(k < nU, g0, p0, std::false_type{});
                    const int k2 = k + 64 < kend ? k + 64 : k;
                    g0 = ul[k2];
                    p0 = a.pw[g0];
                    if (k + 32 < kend) batch(k + 32 < nU, g1, p1, std::false_type{});
                }
                if (__any_sync(FULL_MASK, sawflag))
                    for (int k = lane; k < nU + lane; k += 32) {
                        const int g = ul[k];
                        batch(k < nU, g, a.pw[g], std::true_type{});
                    }
            }
            const int cntBo = TF_CAP - 1 - baseB;
            cntW = __reduce_add_sync(FULL_MASK, cntW);
            __syncwarp();

            float h = 0, rho = 0, drho = 0;
            int cnt = 0;
            unsigned n_search = (MODE & MODE_WVT) ? 1 : 0, n_evals = 0, n_iters = 0;
            if (MODE & MODE_DENSITY) {
                // (3) the outer loop of sph.c:36-64, as far as the two prepared radii carry it
                const float hA = hA_in;
                bool ok = true;
                if (cntA >= TG_DESNNGB) {                    // first search succeeds: Hsml list
                    cnt = cntA; h = hA; n_search += 1;
                } else if (cntA + cntBo >= TG_DESNNGB) {     // second search, 1.23*Hsml
                    // bring the entries beyond Hsml next to the others (ascending: a write never
                    // lands on an entry that is still to be read)
                    const int src0 = TF_CAP - cntBo;
                    for (int b = 0; b < cntBo; b += 32) {
                        const int q = b + lane;
                        const float v = q < cntBo ? rl[src0 + q] : 0.f;
                        __syncwarp();
                        if (q < cntBo) rl[cntA + q] = v;
                        __syncwarp();
                    }
                    cnt = cntA + cntBo; h = (float)((double)hA * 1.23); n_search += 2;
                } else ok = false;                           // a third search: generic path
                int why = 3;
                if (ok) {
                    // pad to whole passes of 64 (TF_CAP is a multiple of 64)
                    const int cnt64 = (cnt + 63) & ~63;
                    if (cnt + lane < cnt64) rl[cnt + lane] = 3.0e38f;
                    if (cnt + 32 + lane < cnt64) rl[cnt + 32 + lane] = 3.0e38f;
                    __syncwarp();
                    ok = find_hsml_fast(a, rl, cnt, h, rho, drho, n_evals, n_iters);
                    why = 4;                                 // no convergence on the frozen list
                }
                if (!ok) { hand_back(i, why); continue; }
            }

            // (4) results
            double dsx = 0, dsy = 0, dsz = 0;
            if (MODE & MODE_ROTA) {
                dsx = warp_sum((double)sx); dsy = warp_sum((double)sy); dsz = warp_sum((double)sz);
                npair = __reduce_add_sync(FULL_MASK, npair);
                n_search = 1;
                if (lane == 0) {
                    a.bfld[3 * (size_t)i] = (float)dsx;
                    a.bfld[3 * (size_t)i + 1] = (float)dsy;
                    a.bfld[3 * (size_t)i + 2] = (float)dsz;
                }
            }
            if (MODE & MODE_WVT) {
                dsx = warp_sum((double)sx); dsy = warp_sum((double)sy); dsz = warp_sum((double)sz);
                npair = __reduce_add_sync(FULL_MASK, npair);
            }
            if (lane == 0) {
                if (MODE & MODE_DENSITY) {                                       // sph.c:66-70
                    const float q = __fmul_rn(__fdiv_rn(h, __fmul_rn(3.f, rho)), drho);
                    a.hsml_out[i] = h;
                    a.rho_out[i] = rho;
                    a.varh_out[i] = __frcp_rn(__fadd_rn(1.f, q));
                }
                if (MODE & MODE_WVT) {
                    a.delta[i] = (float)dsx;
                    a.delta[n + i] = (float)dsy;
                    a.delta[2 * (size_t)n + i] = (float)dsz;
                }
                cw[0] += n_evals + npair; cw[1] += max(cnt, cntW); cw[2] += n_search; cw[3] += n_iters;
            }
        }
        leave();
    }

    if (lane == 0) {
        atomicAdd(&a.counters[0], cw[0]);
        atomicAdd(&a.counters[1], cw[1]);
        atomicAdd(&a.counters[2], cw[2]);
        atomicAdd(&a.counters[3], cw[3]);
    }
}
