// bfield.cuh -- Make_magnetic_field() around the rot(A) sweep (magnetic_field.c:12-131,
// SURVEY 8f-2): the Bonafede-2010 vector potential A = max_halos (rho_gas/rho0)^eta per
// particle (:33-69), B = rot A (the MODE_ROTA sweep, sph.c:216-300), the global normalisation
// to Bfld_Norm / sqrt(3) at the field maximum and the per-particle cap (:71-131) -- all on the
// device, so configuration 4 needs no AoS round trip for Apot / Bfld between the three stages.
#pragma once
#include "common.cuh"
#include "model.cuh"

// setup.c:598-615 (default build): rho0 (1 + (r/rc)^2)^(-3/2 beta) / (1 + (r/rcut)^4), FP64.
static __device__ __forceinline__ double gas_density_profile(double r, const Halo &h)
{
    const double q = r / h.rcore, s = r / h.rcut;
    const double base = __dadd_rn(1.0, __dmul_rn(q, q));
    const double cut = __dadd_rn(1.0, __dmul_rn(__dmul_rn(__dmul_rn(s, s), s), s));
    double rho = h.rho0 * pow(base, -3.0 / 2.0 * h.beta) / cut;
    if (h.rho0_cc != 0) {                                       // setup.c:604-612
        const double qc = r / h.rc_cc;
        rho = __dadd_rn(rho, h.rho0_cc / __dadd_rn(1.0, __dmul_rn(qc, qc)) / cut);
    }
    return rho;
}

// magnetic_field.c:33-69.  pw = positions in the current (Peano) order.
__global__ void k_vector_potential(int n, const float4 *__restrict__ pw, const Halo *__restrict__ halos,
                                   int nhalos, float boxhalf_f, double eta, float *__restrict__ apot)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 p = pw[k];
    double amax = 0;
    for (int i = 0; i < nhalos; i++) {
        const Halo h = halos[i];
        if (h.mass_gas == 0) continue;                                   // :45
        // float dx = Pos - D_CoM - boxhalf: evaluated in double, stored as float (:48-50)
        const float dx = (float)__dsub_rn(__dsub_rn((double)p.x, h.cx), (double)boxhalf_f);
        const float dy = (float)__dsub_rn(__dsub_rn((double)p.y, h.cy), (double)boxhalf_f);
        const float dz = (float)__dsub_rn(__dsub_rn((double)p.z, h.cz), (double)boxhalf_f);
        const double r2 = (double)sq3_nofma(dx, dy, dz);                 // float expression (:52)
        const double rho = gas_density_profile(sqrt(r2), h);
        const double A = pow(rho / h.rho0, eta);                         // :58
        if (A > amax) amax = A;
    }
    const float a = (float)amax;
    apot[3 * (size_t)k] = a; apot[3 * (size_t)k + 1] = a; apot[3 * (size_t)k + 2] = a;
}

// magnetic_field.c:77-86: max over particles of |B|^2, the squares and their sum in float.
__global__ void __launch_bounds__(RED_THREADS)
k_bfld_max(int n, const float *__restrict__ bfld, double *__restrict__ partial)
{
    __shared__ double sm[32];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0;
    if (k < n) v = (double)sq3_nofma(bfld[3 * (size_t)k], bfld[3 * (size_t)k + 1], bfld[3 * (size_t)k + 2]);
    const double m = block_max(v, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
}

__global__ void __launch_bounds__(RED_THREADS) k_bfld_max_final(int nb, const double *__restrict__ partial,
                                                                double *__restrict__ out)
{
    __shared__ double sm[32];
    double v = 0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x) v = fmax(v, partial[k]);
    const double m = block_max(v, sm);
    if (threadIdx.x == 0) out[0] = m;
}

struct HaloExtra {           // what Halo_containing (positions.c:327-388) reads beyond `Halo`
    double r_sample_gas, r_sample_dm;
    int is_stripped;
};

// positions.c:327-388.  x, y, z relative to the box centre.
static __device__ int halo_containing(int type, float x, float y, float z, const Halo *__restrict__ halos,
                                      const HaloExtra *__restrict__ ex, int nhalos, int sub_first,
                                      double boxsize)
{
    if ((double)x > boxsize || (double)y > boxsize || (double)z > boxsize) return -1;
    auto dist = [&](const Halo &h) -> float {
        const double dx = (double)x - h.cx, dy = (double)y - h.cy, dz = (double)z - h.cz;
        return (float)sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    };
    int i = 0;
    if (type > 0) {                                                      // "DM" branch
        if (nhalos > 1) {     // the reference reads Halo[1] of its zero-initialised table otherwise
            const float r = dist(halos[1]);
            if ((double)r < ex[1].r_sample_dm && x > 0) i = 1;
        }
        for (int j = sub_first; j < nhalos; j++) {
            const float r = dist(halos[j]);
            if ((double)r < ex[j].r_sample_dm) { i = j; break; }
        }
    } else {                                                             // SPH branch
        double rho_max = 0;
        for (int j = 0; j < nhalos; j++) {
            if (ex[j].is_stripped) continue;
            const Halo h = halos[j];
            const float r = dist(h);
            const double rho = gas_density_profile((double)r, h);
            if (rho > rho_max && (double)r < ex[j].r_sample_gas) { i = j; rho_max = rho; }
        }
    }
    return i;
}

// magnetic_field.c:92-127: scale, then cap at BMAX (2e-6 inside subhaloes).  The reference
// passes the PARTICLE INDEX as Halo_containing's `type` argument (:109), so particle 0 takes
// the SPH branch and every other one the DM branch; reproduced.
__global__ void k_bfld_normalise(int n, const float4 *__restrict__ pw, float *__restrict__ bfld,
                                 double norm, float boxhalf_f, const Halo *__restrict__ halos,
                                 const HaloExtra *__restrict__ ex, int nhalos, int sub_first,
                                 double boxsize, double bmax_main, double bmax_sub,
                                 int *__restrict__ n_limited)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float b0 = (float)((double)bfld[3 * (size_t)k] * norm);
    float b1 = (float)((double)bfld[3 * (size_t)k + 1] * norm);
    float b2 = (float)((double)bfld[3 * (size_t)k + 2] * norm);
    const double B2 = (double)sq3_nofma(b0, b1, b2);
    const float4 p = pw[k];
    const int i = halo_containing(k, __fsub_rn(p.x, boxhalf_f), __fsub_rn(p.y, boxhalf_f),
                                  __fsub_rn(p.z, boxhalf_f), halos, ex, nhalos, sub_first, boxsize);
    const double bmax = i > 1 ? bmax_sub : bmax_main;
    if (B2 > bmax * bmax) {
        const double f = bmax / sqrt(B2);
        b0 = (float)((double)b0 * f); b1 = (float)((double)b1 * f); b2 = (float)((double)b2 * f);
        atomicAdd(n_limited, 1);
    }
    bfld[3 * (size_t)k] = b0; bfld[3 * (size_t)k + 1] = b1; bfld[3 * (size_t)k + 2] = b2;
}

// positions.c:264-283: haloID of every gas particle (Halo_containing with type 0: the halo of
// the largest gas density among those whose sampling radius holds the particle) and the
// per-halo counts.  posh = current state (x, y, z, Hsml).
__global__ void k_halo_ids(int n, const float4 *__restrict__ posh, float boxhalf_f,
                           const Halo *__restrict__ halos, const HaloExtra *__restrict__ ex, int nhalos,
                           int sub_first, double boxsize, int *__restrict__ ids,
                           unsigned long long *__restrict__ counts)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 p = posh[k];
    const int i = halo_containing(0, __fsub_rn(p.x, boxhalf_f), __fsub_rn(p.y, boxhalf_f),
                                  __fsub_rn(p.z, boxhalf_f), halos, ex, nhalos, sub_first, boxsize);
    ids[k] = i;
    if (i >= 0) atomicAdd(&counts[i], 1ull);
}
