// sph_fast.cuh -- the generic sweep (one warp per target, own walk of the box hierarchy) with
// the FP32 arithmetic of TG_FAST (tile_fast.cuh).  It serves what the tile sweep cannot take in
// that mode: the cold first pass (every target starts from 2*Guess_hsml, ~1000 neighbours) and
// the targets the tiles hand back (candidate set or hit list beyond the tile's caps).  Same
// contract as k_sweep (sph.cuh): exact neighbour sets (float, FMA-free predicate of
// tree.c:67-88, displaced-node open tests, first-2360 cut), the retry loop of sph.c:36-64 and
// the control flow of Find_hsml in FP64; r, u and the WC6 polynomials in float.
#pragma once
#include "tile_fast.cuh"

#define SF_WARPS 8
#define SF_LCAP 2368                     // >= TG_NGBMAX, whole passes of 64
#define SF_SMEM (SF_WARPS * SF_LCAP * 4)

// tree.c:67-88 as a value: the float, FMA-free r^2 of the closest image for the predicate, and
// (r2a, signed separations) from wrap_sep for the value of r and the direction.
static __device__ __forceinline__ float ngb_r2(float xi, float yi, float zi, float xj, float yj, float zj,
                                               float box, float boxhalf, float &r2a, float &sdx, float &sdy,
                                               float &sdz)
{
    float ax = fabsf(__fsub_rn(xi, xj)), ay = fabsf(__fsub_rn(yi, yj)), az = fabsf(__fsub_rn(zi, zj));
    if (ax > boxhalf) ax = __fsub_rn(ax, box);
    if (ay > boxhalf) ay = __fsub_rn(ay, box);
    if (az > boxhalf) az = __fsub_rn(az, box);
    sdx = wrap_sep(xi, xj, box, boxhalf); sdy = wrap_sep(yi, yj, box, boxhalf); sdz = wrap_sep(zi, zj, box, boxhalf);
    r2a = fmaf(sdz, sdz, fmaf(sdy, sdy, sdx * sdx));
    return sq3_nofma(ax, ay, az);
}

template <int MODE, bool USE_LIST>
__global__ void __launch_bounds__(SF_WARPS * 32) k_sweep_fast(const SweepArgs a)
{
    extern __shared__ float s_rl[];      // [SF_WARPS][SF_LCAP]
    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    float *rl = s_rl + w * SF_LCAP;
    const unsigned lt = (1u << lane) - 1;
    const float box = a.bx.box_f, boxhalf = a.bx.boxhalf_f;

    unsigned long long c_evals = 0, c_gath = 0, c_pairs = 0;
    unsigned long long c_search = 0, c_iters = 0;

    float norm = 1.f;
    if (MODE & MODE_WVT) norm = (float)pow(TG_DESNNGB / *a.vsum / K_FOURPITHIRD, 1.0 / 3.0);   // wvt_relax.c:120
    const float cn = 0.5f * norm * box;

    for (;;) {
        const int total = USE_LIST ? *a.nwork : a.hi - a.lo;
        int first = 0;
        if (lane == 0) first = atomicAdd(a.next, SW_CHUNK);
        first = __shfl_sync(FULL_MASK, first, 0);
        if (first >= total) break;
        const int last = min(first + SW_CHUNK, total);

        for (int item = first; item < last; item++) {
            const int i = USE_LIST ? a.worklist[item] : a.lo + item;
            float4 pi = a.pw[i];
            pi.w = fabsf(pi.w);                    // the sign bit is the displaced-node flag
            unsigned g_dens = 0, g_wvt = 0;

            if (MODE & MODE_DENSITY) {
                float h = a.hsml_in[i];
                if (h == 0) h = a.guess[i];                            // sph.c:25-26
                float rho = 0, drho = 0;
                bool done = false;
                for (int guard = 0; guard < 4096 && !done; guard++) {  // sph.c:36-64
                    // Find_ngb_tree(i, h): ascending, stops when the list is full (tree.c:91-92)
                    const float h2 = __fmul_rn(h, h);
                    int cnt = 0;
                    bvh_walk(a.t, a.bx, pi.x, pi.y, pi.z, h, [&](int g) -> bool {
                        const int k = g * 32 + lane;
                        bool hit = false;
                        float r = 0;
                        if (k < a.t.n) {
                            const float4 p = a.pw[k];
                            float sx_, sy_, sz_, r2a;
                            const float r2 = ngb_r2(pi.x, pi.y, pi.z, p.x, p.y, p.z, box, boxhalf, r2a, sx_, sy_, sz_);
                            hit = r2 < h2 && (!df_flagged(p.w) ||
                                              defect_open(a.dnodes + a.dmap[k], pi.x, pi.y, pi.z, h, box, boxhalf));
                            const float y = rsqrt_approx(fmaxf(r2a, 1e-35f));
                            r = r2a * y;
                            r = fmaf(0.5f * y, fmaf(-r, r, r2a), r);
                        }
                        const unsigned m = __ballot_sync(FULL_MASK, hit);
                        const int slot = cnt + __popc(m & lt);
                        if (hit && slot < TG_NGBMAX) rl[slot] = r;
                        cnt += __popc(m);
                        return cnt < TG_NGBMAX;
                    });
                    cnt = min(cnt, TG_NGBMAX);
                    c_search++;
                    g_dens = cnt;
                    if (cnt == TG_NGBMAX) { h = (float)((double)h / 1.24); continue; }
                    if (cnt < TG_DESNNGB) { h = (float)((double)h * 1.23); continue; }
                    const int cnt64 = (cnt + 63) & ~63;
                    if (cnt + lane < cnt64) rl[cnt + lane] = 3.0e38f;
                    if (cnt + 32 + lane < cnt64) rl[cnt + 32 + lane] = 3.0e38f;
                    __syncwarp();
                    unsigned ev = 0, itc = 0;
                    done = find_hsml_fast(a, rl, cnt, h, rho, drho, ev, itc);
                    c_evals += ev; c_iters += itc;
                    __syncwarp();
                }
                if (!done && lane == 0) atomicExch(a.status, 1);
                if (lane == 0) {                                       // sph.c:66-70
                    const float q = __fmul_rn(__fdiv_rn(h, __fmul_rn(3.f, rho)), drho);
                    a.hsml_out[i] = h;
                    a.rho_out[i] = rho;
                    a.varh_out[i] = __frcp_rn(__fadd_rn(1.f, q));
                }
            }

            if (MODE & MODE_WVT) {
                // wvt_relax.c:128-171 with the pair arithmetic of tile_fast.cuh
                const float hi_w = __fmul_rn(pi.w, norm);              // wvt_relax.c:124
                const float hs = (float)((double)hi_w * a.bx.box_d);   // wvt_relax.c:135
                const float hs2 = __fmul_rn(hs, hs);
                const float Af = (float)(a.step * (double)hi_w * TF_KW);
                int cnt = 0;
                float sx = 0, sy = 0, sz = 0;
                bvh_walk(a.t, a.bx, pi.x, pi.y, pi.z, hs, [&](int g) -> bool {
                    const int k = g * 32 + lane;
                    bool hit = false;
                    float f = 0, dX = 0, dY = 0, dZ = 0;
                    if (k < a.t.n) {
                        const float4 p = a.pw[k];
                        float r2a;
                        const float r2 = ngb_r2(pi.x, pi.y, pi.z, p.x, p.y, p.z, box, boxhalf, r2a, dX, dY, dZ);
                        hit = r2 < hs2 && (!df_flagged(p.w) ||
                                           defect_open(a.dnodes + a.dmap[k], pi.x, pi.y, pi.z, hs, box, boxhalf));
                        const float y = rsqrt_approx(fmaxf(r2a, 1e-35f));
                        float r = r2a * y;
                        r = fmaf(0.5f * y, fmaf(-r, r, r2a), r);
                        const float hp = (pi.w + fabsf(p.w)) * cn;
                        const float u = fminf(r * rcp_approx(hp), 1.f);
                        const float t = 1.f - u, t2 = t * t, t4 = t2 * t2;
                        const float P = fmaf(fmaf(fmaf(32.f, u, 25.f), u, 8.f), u, 1.f);
                        f = (t4 * t4) * P * (Af * y);
                        if (u < 1.f && hit && k != i) c_pairs++;
                    }
                    const unsigned m = __ballot_sync(FULL_MASK, hit);
                    const int slot = cnt + __popc(m & lt);
                    cnt += __popc(m);
                    if (hit && slot < TG_NGBMAX && k != i) {            // tree.c:91, wvt_relax.c:141
                        sx = fmaf(f, dX, sx); sy = fmaf(f, dY, sy); sz = fmaf(f, dZ, sz);
                    }
                    return cnt < TG_NGBMAX;
                });
                g_wvt = min(cnt, TG_NGBMAX);
                c_search++;
                const double dsx = warp_sum((double)sx), dsy = warp_sum((double)sy), dsz = warp_sum((double)sz);
                if (lane == 0) {
                    a.delta[i] = (float)dsx;
                    a.delta[a.t.n + i] = (float)dsy;
                    a.delta[2 * (size_t)a.t.n + i] = (float)dsz;
                }
            }
            c_gath += max(g_dens, g_wvt);
        }
    }

    const unsigned long long pairs = warp_sum_u64(c_pairs);
    if (lane == 0) {
        atomicAdd(&a.counters[0], c_evals + pairs);
        atomicAdd(&a.counters[1], c_gath);
        atomicAdd(&a.counters[2], c_search);
        atomicAdd(&a.counters[3], c_iters);
    }
}
