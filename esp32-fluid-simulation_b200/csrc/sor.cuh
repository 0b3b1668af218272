// SOR node update — device restatement of poisson.cpp:63-112, shared by the
// half-sweep, temporally-blocked and ensemble kernels.
#pragma once

#include "fs_common.cuh"

namespace fs {

struct SorCoef {
    float dx;        // pois_context.dx
    float omega;     // pois_context.omega
    float keep;      // (1 - omega), formed in float (poisson.cpp:98,111)
    float neg_half;  // neg_a_ii_inv[2..4], poisson.cpp:67: double literals narrowed to float
    float neg_third;
    float neg_quarter;
};

static inline SorCoef make_sor_coef(float dx, float omega)
{
    SorCoef k;
    k.dx = dx;
    k.omega = omega;
    k.keep = 1 - omega;
    k.neg_half = (float)(-1.0 / 2.0);
    k.neg_third = (float)(-1.0 / 3.0);
    k.neg_quarter = (float)(-1.0 / 4.0);
    return k;
}

// pois_sor_fast, poisson.cpp:101-112.  dxd = dx * d_ij (the same product every
// iteration, so callers may hoist it).  sum = ((L + R) + D) + U.
__device__ __forceinline__ float sor_update_interior(float pc, float l, float r, float d, float u,
                                                     float dxd, const SorCoef &k)
{
    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(l, r), d), u);
    const float gs = __fmul_rn(-0.25f, __fsub_rn(dxd, sum));
    return __fadd_rn(__fmul_rn(k.keep, pc), __fmul_rn(k.omega, gs));
}

// The same update with the Gauss-Seidel coefficient as an operand.  With every MISSING neighbour
// passed as +0.0f, ((L + R) + D) + U is bit-identical to pois_gs_safe's running sum (poisson.cpp:
// 71-86: start from +0, add the existing neighbours in the order L, R, D, U) — adding +0 is exact and
// neither chain can produce -0 once a +0 took part — and coef = neg_a_ii_inv[#neighbours]
// (poisson.cpp:67,88) makes it pois_sor_safe; coef = -0.25f makes it pois_sor_fast.
__device__ __forceinline__ float sor_update_coef(float pc, float l, float r, float d, float u, float dxd,
                                                 float coef, const SorCoef &k)
{
    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(l, r), d), u);
    const float gs = __fmul_rn(coef, __fsub_rn(dxd, sum));
    return __fadd_rn(__fmul_rn(k.keep, pc), __fmul_rn(k.omega, gs));
}

// pois_sor_safe / pois_gs_safe, poisson.cpp:63-99: start from 0, add the
// EXISTING neighbours in order L, R, D, U, divide by their count via the table.
__device__ __forceinline__ float sor_update_wall(float pc, float l, float r, float d, float u,
                                                 bool hl, bool hr, bool hd, bool hu, float dxd,
                                                 const SorCoef &k)
{
    float sum = 0.0f;
    int a = 0;
    if (hl) { sum = __fadd_rn(sum, l); a++; }
    if (hr) { sum = __fadd_rn(sum, r); a++; }
    if (hd) { sum = __fadd_rn(sum, d); a++; }
    if (hu) { sum = __fadd_rn(sum, u); a++; }
    const float coef = a == 4 ? k.neg_quarter : a == 3 ? k.neg_third : a == 2 ? k.neg_half : 0.0f;
    const float gs = __fmul_rn(coef, __fsub_rn(dxd, sum));
    return __fadd_rn(__fmul_rn(k.keep, pc), __fmul_rn(k.omega, gs));
}

}  // namespace fs
