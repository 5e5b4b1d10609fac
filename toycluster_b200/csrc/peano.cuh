// peano.cuh -- Peano-Hilbert keys on the GPU, bit-exact with peano.c:128-203.
//
// The reference maps (x,y,z) in [0,1] to three 64-bit integers scaled by 2^63 in the axis
// order {y,z,x} (peano.c:134-136), runs Skilling's axes->transpose transform over bit planes
// 63..1 (peano.c:140-162), Gray-encodes (peano.c:166-177) and interleaves bit planes
// 63..21 into a 128-bit integer, most significant plane first, finally shifting left by two
// (peano.c:183-200).  The plane-63 triplet is shifted out of the 128 bits entirely, so the
// key is the 42 triplets of planes 62..21 followed by two zero bits.  A coordinate of
// exactly 1.0 sets plane 63 and still perturbs the lower planes through the transform; that
// is reproduced here, not "fixed".
#pragma once
#include "common.cuh"
#include "peano_lut.cuh"

// The key from the 48-state transducer of peano_lut.cuh (scripts/make_peano_lut.py restates
// peano.c:140-198 as a state machine and checks the table against the oracle): two bit planes per
// table look-up instead of ~25 integer instructions per plane (0.91 -> ? ms per 10 M keys).  `lut2` is the
// two-plane table in shared memory.  X[0..2] = {y, z, x} * 2^63 (peano.c:134-136).
static __device__ __forceinline__ void peano_key_lut(const unsigned short *__restrict__ lut2,
                                                     uint64_t X0, uint64_t X1, uint64_t X2,
                                                     uint64_t &hi, uint64_t &lo)
{
    // plane 63 (set only by a coordinate equal to Boxsize): its triplet is shifted out of the
    // 128 bits, but its effect on the state stays (peano.cuh header)
    unsigned state = PEANO_LUT1[(unsigned)(X0 >> 63) << 2 | (unsigned)(X1 >> 63) << 1 | (unsigned)(X2 >> 63)] >> 6;
    uint64_t A = 0, B = 0;        // triplets of planes 62..42 and 41..21
#pragma unroll
    for (int p = 0; p < 21; p++) {
        const int s = 61 - 2 * p;                       // planes s+1, s
        const unsigned idx = ((unsigned)(X0 >> s) & 3u) << 4 | ((unsigned)(X1 >> s) & 3u) << 2 | ((unsigned)(X2 >> s) & 3u);
        const unsigned e = lut2[state * 64 + idx];
        state = e >> 6;
        const uint64_t o = e & 63u;
        if (p < 10) A = A << 6 | o;
        else if (p == 10) { A = A << 3 | (o >> 3); B = o & 7u; }
        else B = B << 6 | o;
    }
    hi = A << 1 | B >> 62;        // 126 bits, left-aligned: key <<= 2 (peano.c:200)
    lo = B << 2;
}

// One thread per particle: keys of pos/Boxsize (peano.c:63-71). posh = (x, y, z, hsml).
__global__ void __launch_bounds__(256)
k_peano_keys(int n, const float4 *__restrict__ posh, double box,
             uint64_t *__restrict__ key_hi, uint64_t *__restrict__ key_lo,
             int *__restrict__ idx, int *__restrict__ range_err)
{
    __shared__ unsigned short lut2[PEANO_STATES * 64];
    for (int k = threadIdx.x; k < PEANO_STATES * 64; k += blockDim.x) lut2[k] = PEANO_LUT2[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = posh[i];
    const double x = (double)p.x / box, y = (double)p.y / box, z = (double)p.z / box;
    if (!(x >= 0 && x <= 1 && y >= 0 && y <= 1 && z >= 0 && z <= 1)) {
        atomicExch(range_err, 1);   // peano.c:130-132 would Assert
        key_hi[i] = ~0ull; key_lo[i] = ~0ull; idx[i] = i;
        return;
    }
    const double scale = 9223372036854775808.0;   // 2^63
    uint64_t hi, lo;
    peano_key_lut(lut2, __double2ull_rz(y * scale), __double2ull_rz(z * scale), __double2ull_rz(x * scale), hi, lo);
    key_hi[i] = hi;
    key_lo[i] = lo;
    idx[i] = i;
}
