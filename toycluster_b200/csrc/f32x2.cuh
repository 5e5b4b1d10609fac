// f32x2.cuh -- Blackwell's packed FP32 instructions (FADD2 / FMUL2 / FFMA2, PTX *.f32x2).
// Each half is an ordinary IEEE round-to-nearest operation, so a packed op on two values is
// bit-identical to the two scalar ops it replaces; it just issues once.  The sweep is
// instruction-issue bound, which is why the float parts of its inner loops are written on
// pairs.
//
// CAUTION (ptxas 12.9): a mul.rn.f32x2 whose result feeds an add.rn.f32x2 is contracted into
// one FFMA2 despite the explicit rounding modifiers (even with -Xptxas -fmad=false), i.e. one
// rounding instead of two.  Where the reference's arithmetic needs both roundings, write the
// step so that no rounded product is the direct operand of a packed add (see find_hsml).
#pragma once
#include <cuda_runtime.h>

typedef unsigned long long f32x2;   // two floats in one 64-bit register pair: {lo, hi}

static __device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
static __device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
static __device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
static __device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
static __device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
static __device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
