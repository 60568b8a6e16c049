// mg_exact.cuh -- rounding-exact arithmetic helpers shared by all kernels.
//
// The parity contract is bit-equality with the reference CPU solver, which is compiled for
// baseline x86-64: every +,-,*,/ is a separately rounded IEEE operation (no FMA contraction,
// SURVEY.md 7.2).  nvcc contracts a*b+c into FMA by default, so the kernels never write infix
// arithmetic on field values: they go through these intrinsics, which ptxas may not contract.
#pragma once
#include <cuda_runtime.h>

namespace mgx {

__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Correctly rounded a/b from a correctly rounded reciprocal y = RN(1/b) computed once on the host
// (Markstein 1990: q = RN(a*y); r = a - b*q exactly with one FMA; q' = RN(q + r*y) is the correctly
// rounded quotient).  Three pipe operations instead of the ~20-instruction IEEE division sequence;
// bit-identical to a/b whenever q is finite and no intermediate underflows.  Non-finite q (inf/NaN
// input) is passed through so that overflowed REF_COMPAT runs still match the reference.
template <typename T>
__device__ __forceinline__ T div_by_const(T a, T b, T y)
{
    T q = mul(a, y);
    T r = fma_(-b, q, a);
    T q2 = fma_(r, y, q);
    return (q - q == T(0)) ? q2 : q;  // q - q is NaN for inf/NaN
}

}  // namespace mgx
