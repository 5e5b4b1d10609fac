// common.cuh -- shared device helpers and the parameter block every kernel receives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// -DTG_CUBIC_SPLINE builds the library the way -DSPH_CUBIC_SPLINE builds the reference
// (Makefile:25): M4 kernel in Find_hsml (sph.c:140-146, 442-466), DESNNGB 50 and NGBMAX 400
// (globals.h:40-52), no bias correction and no dRhodHsml (sph.c:201), WVT step 0.035
// (wvt_relax.c:48-49).  It is a separate shared object, libtoygpu_m4.so.
#ifdef TG_CUBIC_SPLINE
#define TG_DESNNGB 50
#define TG_NGBMAX 400
#else
#define TG_DESNNGB 295
#define TG_NGBMAX 2360
#endif
#define FULL_MASK 0xffffffffu
#define MAX_LEVELS 8
#define MAX_HALOS 4096   // MAXHALOS, globals.h:58

// globals.h:62-63 -- the reference's literal constants, not the exact values
#define K_SQRT3 1.73205080756887719
#define K_FOURPITHIRD 4.18879032135009765
#define K_PI 3.14159265358979323846

struct Halo {            // one row of Global_density_model's table
    double cx, cy, cz;   // D_CoM; Boxsize/2 is subtracted at evaluation time (wvt_relax.c:240)
    double rho0, beta, rcore, rcut;
    double mass_gas;
    // -DDOUBLE_BETA_COOL_CORES (setup.c:604-612): rho0 * Rho0_Fac and rcore / Rc_Fac of a cuspy
    // halo; rho0_cc == 0 => no second beta component (the default build ignores Have_Cuspy)
    double rho0_cc, rc_cc;
};

// Sorted-order bounding-box hierarchy over groups of 32 consecutive particles.
// Level 0 = groups of 32 particles, level l+1 = groups of 32 level-l nodes.
struct Bvh {
    int n;                    // particles
    int top;                  // highest level; it has <= 32 nodes
    int lvl_n[MAX_LEVELS];    // nodes per level
    int lvl_off[MAX_LEVELS];  // offset of each level in the SoA arrays
    const float *cx, *cy, *cz;   // box centres
    const float *hx, *hy, *hz;   // box half-widths (already inflated for rounding)
    // four sub-boxes (runs of 8 particles) per level-0 box, index 4*g + s; used by the tile walk
    const float *scx, *scy, *scz, *shx, *shy, *shz;
};

struct Box {
    float box_f, boxhalf_f;             // tree.c:27-28
    double box_d, boxhalf_d, boxinv_d;  // sph.c:83-84, wvt_relax.c:28-29
    double mpart;                       // Param.Mpart[0]
};

static __device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

static __device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;   // xor butterfly: every lane ends with the bit-identical total
}

static __device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Correctly rounded float ops without FMA contraction: the reference is built -std=c99,
// i.e. -ffp-contract=off, so tree.c:88 is three rounded products and two rounded sums.
static __device__ __forceinline__ float sq3_nofma(float a, float b, float c)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}
