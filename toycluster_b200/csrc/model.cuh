// model.cuh -- the O(N) passes around the sweep: reorder + density model + WVT hsml
// (peano.c:85-126, wvt_relax.c:108-118, 227-256), the error pass (wvt_relax.c:73-87) and the
// move + wrap pass (wvt_relax.c:175-214).
#pragma once
#include "common.cuh"

#define RED_THREADS 256

// wvt_relax.c:227-256 + setup.c:598-615: max over halos of the beta model with r^4 cut-off,
// FP64, returned as float. No periodic wrap around the halo centre (the reference has none).
static __device__ __forceinline__ double halo_density(const Halo &h, double x, double y, double z,
                                                      double boxhalf)
{
    const double dx = __dsub_rn(__dsub_rn(x, h.cx), boxhalf);   // wvt_relax.c:240-242
    const double dy = __dsub_rn(__dsub_rn(y, h.cy), boxhalf);
    const double dz = __dsub_rn(__dsub_rn(z, h.cz), boxhalf);
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const double r = sqrt(r2);
    const double q = r / h.rcore, s = r / h.rcut;
    // setup.c:601-602, product and sum kept un-contracted like the -std=c99 build
    const double base = __dadd_rn(1.0, __dmul_rn(q, q));
    const double cut = __dadd_rn(1.0, __dmul_rn(__dmul_rn(__dmul_rn(s, s), s), s));
    double rho = h.rho0 * pow(base, -3.0 / 2.0 * h.beta) / cut;
    if (h.rho0_cc != 0) {                                       // setup.c:604-612
        const double qc = r / h.rc_cc;
        rho = __dadd_rn(rho, h.rho0_cc / __dadd_rn(1.0, __dmul_rn(qc, qc)) / cut);
    }
    return rho;
}

// The same profile in fast float arithmetic, good to ~1e-5: only used to decide WHICH halos
// can hold the maximum, never for the value.
static __device__ __forceinline__ float halo_density_estimate(const Halo &h, float x, float y, float z,
                                                              float boxhalf)
{
    const float dx = x - (float)h.cx - boxhalf, dy = y - (float)h.cy - boxhalf,
                dz = z - (float)h.cz - boxhalf;
    const float r2 = dx * dx + dy * dy + dz * dz;
    const float rc = (float)h.rcore, rt = (float)h.rcut;
    const float s2 = r2 / (rt * rt);
    float rho = (float)h.rho0 * __powf(1.f + r2 / (rc * rc), -1.5f * (float)h.beta) / (1.f + s2 * s2);
    if (h.rho0_cc != 0) {
        const float rcc = (float)h.rc_cc;
        rho += (float)h.rho0_cc / (1.f + r2 / (rcc * rcc)) / (1.f + s2 * s2);
    }
    return rho;
}

static __device__ __forceinline__ float global_density_model(float xf, float yf, float zf,
                                                             const Halo *__restrict__ halos,
                                                             int nhalos, double boxhalf)
{
    const double x = xf, y = yf, z = zf;
    double rho = 0;
    if (nhalos <= 1) {
        for (int i = 0; i < nhalos; i++) {
            const Halo h = halos[i];
            if (h.mass_gas == 0) continue;                   // wvt_relax.c:237
            rho = fmax(halo_density(h, x, y, z, boxhalf), rho);
        }
        return (float)rho;
    }
    // Substructure runs carry ~70 rows (substructure.c:127), a merger two.  The result is a
    // maximum, so the FP64 pow is only needed for rows whose float estimate is within 1e-3 of the
    // best estimate (the estimate is good to ~1e-5): same value, ~1 pow per particle instead of
    // one per row.
    const float bh = (float)boxhalf;
    float best = 0;
    for (int i = 0; i < nhalos; i++) {
        const Halo h = halos[i];
        if (h.mass_gas == 0) continue;
        best = fmaxf(best, halo_density_estimate(h, xf, yf, zf, bh));
    }
    const float cut = best * 0.999f;
    for (int i = 0; i < nhalos; i++) {
        const Halo h = halos[i];
        if (h.mass_gas == 0) continue;
        if (!(halo_density_estimate(h, xf, yf, zf, bh) < cut))
            rho = fmax(halo_density(h, x, y, z, boxhalf), rho);
    }
    return (float)rho;
}

// Block-wide deterministic sum / max helpers (fixed tree order => run-to-run reproducible).
static __device__ __forceinline__ double block_sum(double v, double *sm)
{
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
        t = warp_sum(t);
    }
    __syncthreads();
    return t;   // valid in warp 0
}

static __device__ __forceinline__ double block_max(double v, double *sm)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(FULL_MASK, t, o));
    }
    __syncthreads();
    return t;
}

// Reorder into Peano order and evaluate everything that depends on position only.
//   in : posh (x,y,z,Hsml) and id in the previous order, idx = sort permutation
//   out: pw (x,y,z, raw WVT hsml), hsml, id, rho_model, sorted key_lo; partial sums of raw^3
__global__ void __launch_bounds__(RED_THREADS)
k_reorder_model(int n, const int *__restrict__ idx, const float4 *__restrict__ posh_in,
                const int *__restrict__ id_in, const uint64_t *__restrict__ key_lo_in,
                const float *__restrict__ apot_in, const float *__restrict__ rmstate_in,
                float *__restrict__ rmstate_out, float4 *__restrict__ pw, float *__restrict__ pwp, float *__restrict__ soa,
                float *__restrict__ hsml, int *__restrict__ id_out,
                float *__restrict__ rho_model, uint64_t *__restrict__ key_lo_out,
                float *__restrict__ apot_out,
                const Halo *__restrict__ halos, int nhalos, double mpart, double boxhalf,
                double *__restrict__ partial)
{
    __shared__ double sm[32];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    double cube = 0;
    if (k < n) {
        const int src = idx[k];
        const float4 p = posh_in[src];
        const float rm = global_density_model(p.x, p.y, p.z, halos, nhalos, boxhalf);
        // wvt_relax.c:115: hsml = pow(WVTNNGB * Mpart / rho / fourpithird, 1/3) -> float
        const float hw = (float)pow(TG_DESNNGB * mpart / (double)rm / K_FOURPITHIRD, 1. / 3.);
        pw[k] = make_float4(p.x, p.y, p.z, hw);
        {   // pair-interleaved copy: {x0, x1, y0, y1}, {z0, z1, w0, w1} per pair of particles
            float *q = pwp + 8 * (size_t)(k >> 1) + (k & 1);
            q[0] = p.x; q[2] = p.y; q[4] = p.z; q[6] = hw;
        }
        const size_t n8 = ((size_t)n + 7) & ~(size_t)7;
        soa[k] = p.x; soa[n8 + k] = p.y; soa[2 * n8 + k] = p.z;
        hsml[k] = p.w;
        id_out[k] = id_in[src];
        rho_model[k] = rm;
        key_lo_out[k] = key_lo_in[src];
        rmstate_out[k] = rmstate_in[src];      // SphP.Rho_Model travels with the record
        if (apot_in) {
            apot_out[3 * k] = apot_in[3 * src];
            apot_out[3 * k + 1] = apot_in[3 * src + 1];
            apot_out[3 * k + 2] = apot_in[3 * src + 2];
        }
        cube = (double)__fmul_rn(__fmul_rn(hw, hw), hw);   // p3() on the float, wvt_relax.c:117
    }
    else if (k < (int)(((size_t)n + 7) & ~(size_t)7)) {
        // pad of the last run of 8: NaN compares false against every radius, with or without the
        // periodic wrap of phase 1 (a large finite pad wraps to 0 when Boxsize is a power of two)
        const size_t n8 = ((size_t)n + 7) & ~(size_t)7;
        const float qnan = __int_as_float(0x7fc00000);
        soa[k] = qnan; soa[n8 + k] = qnan; soa[2 * n8 + k] = qnan;
        // the pair-interleaved copy gets a FINITE far-away pad: the partner of the last particle
        // is evaluated by the packed phase 2 with weight 0, and 0 * NaN would poison its sums
        float *q = pwp + 8 * (size_t)(k >> 1) + (k & 1);
        q[0] = 1e18f; q[2] = 1e18f; q[4] = 1e18f; q[6] = 0.f;
        pw[k] = make_float4(qnan, qnan, qnan, 0.f);          // pw has n8 entries: the pad is addressable
    }
    const double s = block_sum(cube, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Sum `count` partials in a fixed order with one block: out[0] = sum.
__global__ void __launch_bounds__(RED_THREADS)
k_final_sum(int count, const double *__restrict__ partial, double *__restrict__ out)
{
    __shared__ double sm[32];
    double v = 0;
    for (int k = threadIdx.x; k < count; k += RED_THREADS) v += partial[k];
    const double s = block_sum(v, sm);
    if (threadIdx.x == 0) out[0] = s;
}

// wvt_relax.c:73-87 over the slice [lo, hi): partial[2b] = sum err, partial[2b+1] = max err.
__global__ void __launch_bounds__(RED_THREADS)
k_error(int lo, int hi, const float *__restrict__ rho, const float *__restrict__ rho_model,
        double *__restrict__ partial)
{
    __shared__ double sm[32];
    const int k = lo + blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0;
    if (k < hi) {
        const float rm = rho_model[k];
        // float err = fabs(Rho - rho) / rho : float difference, double divide, float result
        e = (double)(float)(fabs((double)__fsub_rn(rho[k], rm)) / (double)rm);
    }
    const double s = block_sum(e, sm);
    const double m = block_max(e, sm);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s; partial[2 * blockIdx.x + 1] = m; }
}

__global__ void __launch_bounds__(RED_THREADS)
k_final_err(int count, const double *__restrict__ partial, double *__restrict__ out)
{
    __shared__ double sm[32];
    double s = 0, m = 0;
    for (int k = threadIdx.x; k < count; k += RED_THREADS) {
        s += partial[2 * k];
        m = fmax(m, partial[2 * k + 1]);
    }
    s = block_sum(s, sm);
    m = block_max(m, sm);
    if (threadIdx.x == 0) { out[0] = s; out[1] = m; }
}

// wvt_relax.c:193-213: Pos += (float)(delta * boxsize), wrap into [0, Boxsize]; the result
// (with the freshly solved Hsml) becomes the next iteration's unsorted input.
// `scale` rescales a displacement computed with a stale step (fused sweep); 1 otherwise.
__global__ void k_move(int lo, int hi, int n, const float4 *__restrict__ pw,
                       const float *__restrict__ hsml, float *__restrict__ delta, double box,
                       double scale, float4 *__restrict__ posh_out)
{
    const int k = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= hi) return;
    const float4 p = pw[k];
    float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float d = delta[(size_t)a * n + k];
        if (scale != 1.0) {      // the sweep ran with the step as it stood before wvt_relax.c:100
            d = (float)((double)d * scale);
            delta[(size_t)a * n + k] = d;
        }
        float x = __fadd_rn(c[a], (float)((double)d * box));
        for (int guard = 0; guard < 64 && (double)x < 0; guard++) x = (float)((double)x + box);
        for (int guard = 0; guard < 64 && (double)x > box; guard++) x = (float)((double)x - box);
        c[a] = x;
    }
    posh_out[k] = make_float4(c[0], c[1], c[2], hsml[k]);
}

// No displacement: carry (x, y, z, Hsml) of the sorted order into the next input buffer.
__global__ void k_carry(int lo, int hi, const float4 *__restrict__ pw,
                        const float *__restrict__ hsml, float4 *__restrict__ posh_out)
{
    const int k = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= hi) return;
    const float4 p = pw[k];
    posh_out[k] = make_float4(p.x, p.y, p.z, hsml[k]);
}

// Host SoA staging <-> packed state. cold_flag is raised when any Hsml is 0 (sph.c:25).
__global__ void k_pack_state(int n, const float *__restrict__ pos, const float *__restrict__ hsml,
                             float4 *__restrict__ posh, int *__restrict__ id, int *__restrict__ cold_flag)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float h = hsml ? hsml[k] : 0.f;
    posh[k] = make_float4(pos[3 * (size_t)k], pos[3 * (size_t)k + 1], pos[3 * (size_t)k + 2], h);
    id[k] = k;
    if (h == 0.f) *cold_flag = 1;
}

__global__ void k_unpack_state(int n, const float4 *__restrict__ posh, float *__restrict__ pos,
                               float *__restrict__ hsml)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 p = posh[k];
    pos[3 * (size_t)k] = p.x; pos[3 * (size_t)k + 1] = p.y; pos[3 * (size_t)k + 2] = p.z;
    hsml[k] = p.w;
}

__global__ void k_iota(int n, int *__restrict__ id)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) id[k] = k;
}
