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

struct Transposed { uint64_t a, b, c; };   // X[0], X[1], X[2] after the transform

// LOW = lowest bit plane the caller reads.  Plane q of the transform only rewrites bits below q
// and the Gray code only propagates from high to low bits, so planes below LOW can be left
// undone: the bits >= LOW of the result are those of the full loop (peano.c:140-162 runs to 1).
template <int LOW = 1>
static __device__ __forceinline__ Transposed hilbert_transpose(double x, double y, double z)
{
    const double scale = 9223372036854775808.0;   // 2^63
    uint64_t a = __double2ull_rz(y * scale);
    uint64_t b = __double2ull_rz(z * scale);
    uint64_t c = __double2ull_rz(x * scale);

    // planes 63 .. 1: conditional invert of the low bits of `a`, or exchange with b / c
    for (int plane = 63; plane >= (LOW > 1 ? LOW : 1); plane--) {
        const uint64_t q = 1ull << plane;
        const uint64_t low = q - 1;

        if (a & q) a ^= low;

        if (b & q) { a ^= low; }
        else { uint64_t t = (a ^ b) & low; a ^= t; b ^= t; }

        if (c & q) { a ^= low; }
        else { uint64_t t = (a ^ c) & low; a ^= t; c ^= t; }
    }

    // Gray encode
    b ^= a;
    c ^= b;
    uint64_t g = c;
    g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8; g ^= g >> 16; g ^= g >> 32;
    const uint64_t t = c ^ g;    // peano.c:169-174: t = X2_before ^ prefix-xor(X2)
    c = g;
    b ^= t;
    a ^= t;
    return {a, b, c};
}

// 128-bit key as (hi, lo).
static __device__ __forceinline__ void peano_key(double x, double y, double z,
                                                 uint64_t &hi, uint64_t &lo)
{
    const Transposed T = hilbert_transpose<21>(x, y, z);
    uint64_t h = 0, l = 0;
    // planes 62..21 -> 126 bits; the first 21 triplets + 1 bit land in hi.
#pragma unroll 1
    for (int plane = 62; plane >= 21; plane--) {
        const uint64_t tri = (((T.a >> plane) & 1) << 2) | (((T.b >> plane) & 1) << 1) |
                             ((T.c >> plane) & 1);
        h = (h << 3) | (l >> 61);
        l = (l << 3) | tri;
    }
    hi = (h << 2) | (l >> 62);
    lo = l << 2;
}

// One thread per particle: keys of pos/Boxsize (peano.c:63-71). posh = (x, y, z, hsml).
__global__ void k_peano_keys(int n, const float4 *__restrict__ posh, double box,
                             uint64_t *__restrict__ key_hi, uint64_t *__restrict__ key_lo,
                             int *__restrict__ idx, int *__restrict__ range_err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = posh[i];
    const double x = (double)p.x / box, y = (double)p.y / box, z = (double)p.z / box;
    if (!(x >= 0 && x <= 1 && y >= 0 && y <= 1 && z >= 0 && z <= 1)) {
        atomicExch(range_err, 1);   // peano.c:130-132 would Assert
        key_hi[i] = ~0ull; key_lo[i] = ~0ull; idx[i] = i;
        return;
    }
    uint64_t hi, lo;
    peano_key(x, y, z, hi, lo);
    key_hi[i] = hi;
    key_lo[i] = lo;
    idx[i] = i;
}
