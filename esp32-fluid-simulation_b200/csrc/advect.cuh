// Semi-Lagrangian backtrace + bilinear sample — device-side restatement of
// advect.h:24-85 shared by every advect kernel (gather, TMA-tiled, ensemble).
#pragma once

#include "fs_common.cuh"

namespace fs {

// Payload traits: Vector2<float> (ino:253) and Vector3<UQ32> (ino:282).
struct Vec2Payload {
    static constexpr int NC = 2;
    using raw_t = float;
    __device__ static __forceinline__ float to_float(float r) { return r; }
    __device__ static __forceinline__ float from_float(float f) { return f; }
};
struct RgbPayload {
    static constexpr int NC = 3;
    using raw_t = uint32_t;
    // UQ32 -> float on every read (vector.h:116-118 via uq32.h:15); float -> UQ32
    // exactly once per converting assignment (vector.h:72-73 via uq32.h:13)
    __device__ static __forceinline__ float to_float(uint32_t r) { return uq32_to_float(r); }
    __device__ static __forceinline__ uint32_t from_float(float f) { return uq32_from_float(f); }
};

// sample<T>, advect.h:24-72.  `fetch(gi, gj, out[NC])` returns the raw payload of
// GLOBAL node (gi,gj); where it comes from (global memory, a shared-memory tile,
// a peer window) is the caller's business.
template <class P, class Fetch>
__device__ __forceinline__ void sample(typename P::raw_t (&out)[P::NC], const Fetch &fetch,
                                       float i, float j, int GX, int GY, bool no_slip)
{
    using raw_t = typename P::raw_t;
    constexpr int NC = P::NC;
    const bool x_under = i < 0.0f;                      // advect.h:26-29
    const bool x_over = i >= (float)(GX - 1);
    const bool y_under = j < 0.0f;
    const bool y_over = j >= (float)(GY - 1);
    const bool x_oob = x_under || x_over, y_oob = y_under || y_over;

    const float i_floor = floorf(i), j_floor = floorf(j);  // advect.h:34-35
    const float di = __fsub_rn(i, i_floor), dj = __fsub_rn(j, j_floor);
    const float wi = __fsub_rn(1.0f, di), wj = __fsub_rn(1.0f, dj);

    if (!x_oob && !y_oob) {                             // advect.h:38-42
        const int gi = (int)i_floor, gj = (int)j_floor;
        raw_t p11[NC], p12[NC], p21[NC], p22[NC];
        fetch(gi, gj, p11);
        fetch(gi, gj + 1, p12);
        fetch(gi + 1, gj, p21);
        fetch(gi + 1, gj + 1, p22);
#pragma unroll
        for (int ch = 0; ch < NC; ch++) {
            float a = mixf(wj, dj, P::to_float(p11[ch]), P::to_float(p12[ch]));
            float b = mixf(wj, dj, P::to_float(p21[ch]), P::to_float(p22[ch]));
            out[ch] = P::from_float(mixf(wi, di, a, b));
        }
        return;
    }

    raw_t e[NC];                                        // T p_edge, advect.h:45
    if (x_oob && y_oob) {                               // corner copy, advect.h:46-48
        fetch(x_under ? 0 : GX - 1, y_under ? 0 : GY - 1, e);
    } else if (x_oob) {                                 // advect.h:49-51
        const int gi = x_under ? 0 : GX - 1, gj = (int)j_floor;
        raw_t a[NC], b[NC];
        fetch(gi, gj, a);
        fetch(gi, gj + 1, b);
#pragma unroll
        for (int ch = 0; ch < NC; ch++)
            e[ch] = P::from_float(mixf(wj, dj, P::to_float(a[ch]), P::to_float(b[ch])));
    } else {                                            // advect.h:52-54
        const int gi = (int)i_floor, gj = y_under ? 0 : GY - 1;
        raw_t a[NC], b[NC];
        fetch(gi, gj, a);
        fetch(gi + 1, gj, b);
#pragma unroll
        for (int ch = 0; ch < NC; ch++)
            e[ch] = P::from_float(mixf(wi, di, P::to_float(a[ch]), P::to_float(b[ch])));
    }
    if (!no_slip) {                                     // advect.h:57-59
#pragma unroll
        for (int ch = 0; ch < NC; ch++) out[ch] = e[ch];
        return;
    }
    float f = 1.0f;                                     // advect.h:62-70
    if (x_oob) {
        float o = x_under ? -i : __fsub_rn(i, (float)(GX - 1));
        f = __fmul_rn(f, discount(o));
    }
    if (y_oob) {
        float o = y_under ? -j : __fsub_rn(j, (float)(GY - 1));
        f = __fmul_rn(f, discount(o));
    }
#pragma unroll
    for (int ch = 0; ch < NC; ch++)                     // advect.h:71
        out[ch] = P::from_float(__fmul_rn(P::to_float(e[ch]), f));
}

// advect.h:81 — source = (i,j) - vel*dt in float
__device__ __forceinline__ void backtrace(float &si, float &sj, int gi, int gj, float2 vel,
                                          float dt)
{
    si = __fsub_rn((float)gi, __fmul_rn(vel.x, dt));
    sj = __fsub_rn((float)gj, __fmul_rn(vel.y, dt));
}

// Fetch from the local window in global memory; a backtrace that leaves the
// window (only possible on a decomposed grid) raises the status flag.
template <class P>
struct GlobalFetch {
    const typename P::raw_t *__restrict__ base;
    int ox, oy, nx;
    int vx0, vy0, vw, vh;    // valid part of the window (Geo::vx0..)
    int *status;
    __device__ __forceinline__ void operator()(int gi, int gj, typename P::raw_t (&o)[P::NC]) const
    {
        const int lx = gi - ox, ly = gj - oy;
        if ((unsigned)(lx - vx0) >= (unsigned)vw || (unsigned)(ly - vy0) >= (unsigned)vh) {
            if (status) atomicCAS(status, 0, FS_ERR_HALO_OVERRUN);   // the first error wins
#pragma unroll
            for (int ch = 0; ch < P::NC; ch++) o[ch] = 0;
            return;
        }
        const typename P::raw_t *q = base + ((size_t)ly * nx + lx) * P::NC;
        if constexpr (P::NC == 2) {
            float2 t = __ldg(reinterpret_cast<const float2 *>(q));
            o[0] = t.x;
            o[1] = t.y;
        } else {
#pragma unroll
            for (int ch = 0; ch < P::NC; ch++) o[ch] = __ldg(q + ch);
        }
    }
};

}  // namespace fs
