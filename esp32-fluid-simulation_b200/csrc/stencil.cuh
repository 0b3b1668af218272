// Per-node finite differences (finitediff.cpp:9-73) shared by the stand-alone stencil kernels and
// the fused advect kernels.
#pragma once

#include "fs_common.cuh"

namespace fs {

// v - grad p at local node (lx,ly): grad_sub_expr_fast/safe, finitediff.cpp:41-73.  A missing
// neighbour is replaced by the node's own pressure (finitediff.cpp:51-54).
__device__ __forceinline__ float2 grad_sub_value(float2 c, const float *__restrict__ p, const Geo &g, int lx, int ly,
                                                 float two_dx_inv)
{
    const size_t l = (size_t)ly * g.nx + lx;
    const int gi = g.ox + lx, gj = g.oy + ly;
    const float pc = __ldg(&p[l]);
    const float pl = gi > 0 ? __ldg(&p[l - 1]) : pc;
    const float pr = gi < g.GX - 1 ? __ldg(&p[l + 1]) : pc;
    const float pd = gj > 0 ? __ldg(&p[l - g.nx]) : pc;
    const float pu = gj < g.GY - 1 ? __ldg(&p[l + g.nx]) : pc;
    c.x = __fsub_rn(c.x, __fmul_rn(__fsub_rn(pr, pl), two_dx_inv));
    c.y = __fsub_rn(c.y, __fmul_rn(__fsub_rn(pu, pd), two_dx_inv));
    return c;
}

// divergence from the four neighbours' values (div_expr_fast / div_expr_safe, finitediff.cpp:9-31):
// c = the node, lft/rgt/dwn/upp = its neighbours where they exist.
__device__ __forceinline__ float div_value(float2 c, float lft_x, float rgt_x, float dwn_y, float upp_y, int gi,
                                           int gj, int GX, int GY, float two_dx_inv)
{
    const int i_max = GX - 1, j_max = GY - 1;
    float s;
    if (gi > 0 && gi < i_max && gj > 0 && gj < j_max) {
        s = __fadd_rn(__fadd_rn(-lft_x, rgt_x), __fadd_rn(-dwn_y, upp_y));
    } else {   // ghost velocity is the negated wall node; strictly left-to-right sum from zero
        s = 0.0f;
        s = __fadd_rn(s, gi > 0 ? -lft_x : c.x);
        s = __fadd_rn(s, gi < i_max ? rgt_x : -c.x);
        s = __fadd_rn(s, gj > 0 ? -dwn_y : c.y);
        s = __fadd_rn(s, gj < j_max ? upp_y : -c.y);
    }
    return __fmul_rn(s, two_dx_inv);
}

}  // namespace fs
