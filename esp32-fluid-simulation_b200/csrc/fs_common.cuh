// Shared device/host helpers for the stable-fluids kernels (sm_100a).
//
// Bit-exactness rules (SURVEY.md §7 hard part 1): the reference's results are
// only reproducible with FMA contraction OFF and the reference's operand
// association.  Every arithmetic step that exists in the reference is written
// here with the non-contractable intrinsics __fmul_rn/__fadd_rn/__fsub_rn, so
// parity does not depend on a compiler flag (the build passes -fmad=false as
// well).  The path is HBM-bound; losing FMA costs nothing.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fluid_b200.h"

namespace fs {

// Geometry of one launch.  A single-GPU grid is the special case ox=oy=0,
// nx=GX, ny=GY, rectangle = everything.  For a decomposed grid the arrays are a
// rank's padded local window (fs_tile in the C ABI): wall rules, red/black
// parity and advect coordinates are all evaluated in GLOBAL coordinates.
struct Geo {
    int GX, GY;          // global grid dims (dim_x = fast axis, dim_y)
    int ox, oy;          // global coordinate of local element (0,0)
    int nx, ny;          // local array extents; pitch = nx
    int x0, y0, x1, y1;  // compute rectangle in local coordinates, [x0,x1) x [y0,y1)
    int vx0, vy0, vx1, vy1;  // part of the window that holds VALID data for an advect's gathers (a fetch
                             // outside it raises FS_ERR_HALO_OVERRUN); the whole window unless a
                             // decomposed step refreshed only part of the ghosts
};

static inline Geo geo_full(int dim_x, int dim_y)
{
    Geo g;
    g.GX = dim_x; g.GY = dim_y; g.ox = 0; g.oy = 0; g.nx = dim_x; g.ny = dim_y;
    g.x0 = 0; g.y0 = 0; g.x1 = dim_x; g.y1 = dim_y;
    g.vx0 = 0; g.vy0 = 0; g.vx1 = dim_x; g.vy1 = dim_y;
    return g;
}

static inline Geo geo_tile(const fs_tile &t)
{
    Geo g;
    g.GX = t.gdim_x; g.GY = t.gdim_y; g.ox = t.ox; g.oy = t.oy; g.nx = t.nx; g.ny = t.ny;
    g.x0 = t.x0; g.y0 = t.y0; g.x1 = t.x1; g.y1 = t.y1;
    g.vx0 = 0; g.vy0 = 0; g.vx1 = t.nx; g.vy1 = t.ny;
    return g;
}

// ---- arithmetic that must match the reference bit for bit --------------------

// lerp, advect.h:13-16: p1*(1-d) + p2*d with (1-d) formed once in float
__device__ __forceinline__ float mixf(float w, float d, float a, float b)
{
    return __fadd_rn(__fmul_rn(a, w), __fmul_rn(b, d));
}

// UQ32(float), uq32.h:13: raw = (uint32_t)(x + 0.5f); cvt.rzi.u32.f32 saturates
__device__ __forceinline__ uint32_t uq32_from_float(float x)
{
    return __float2uint_rz(__fadd_rn(x, 0.5f));
}

// UQ32::operator float, uq32.h:15
__device__ __forceinline__ float uq32_to_float(uint32_t raw) { return __uint2float_rn(raw); }

// no-slip discount, advect.h:64-65 / 68-69
__device__ __forceinline__ float discount(float o)
{
    return o < 0.5f ? __fsub_rn(1.0f, __fmul_rn(2.0f, o)) : 0.0f;
}

// ---- streaming loads/stores ---------------------------------------------------
__device__ __forceinline__ float2 ldg_f2(const float2 *p) { return __ldg(p); }

}  // namespace fs

#define FS_CUDA_TRY(expr)                                  \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return (int)_e;             \
    } while (0)
