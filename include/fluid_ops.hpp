// fluid_ops.hpp — C++ mirror of the reference's sim-step operator surface over the
// C ABI of fluid_b200.h.
//
// Same names, argument order and meaning as the reference's free functions, so a call
// site in loop() (ESP32-fluid-simulation.ino:249-289) switches implementation by
// changing a namespace:
//
//     advect(v_temp, velocity_field, velocity_field, N_ROWS, N_COLS, DT, true);        // reference
//     fluid_b200::advect(v_temp, velocity_field, velocity_field, N_ROWS, N_COLS, DT, true);
//
// Pointers are HOST pointers, exactly what the reference holds; each call copies in,
// runs the CUDA kernels and copies out (fsh_* entry points).  The element types are
// templates constrained by size only, so the reference's own Vector2<float> (8 B) and
// Vector3<UQ32> (12 B) bind without this header including the reference's headers.
// Errors: the reference's operators return void and cannot fail; here a failing CUDA
// call throws fluid_b200::error (there is no CPU fallback).
#ifndef FLUID_OPS_HPP
#define FLUID_OPS_HPP

#include <stdexcept>
#include <string>

#include "fluid_b200.h"

namespace fluid_b200 {

struct error : std::runtime_error {
    int code;
    error(int c, const char *what) : std::runtime_error(std::string(what) + ": " + fs_error_string(c)), code(c) {}
};

inline void check(int code, const char *what)
{
    if (code != FS_OK) throw error(code, what);
}

// One lazily created context per thread (device 0, legacy default stream).
inline fs_ctx *context()
{
    struct holder {
        fs_ctx *ctx = nullptr;
        ~holder() { if (ctx) fs_ctx_destroy(ctx); }
    };
    static thread_local holder h;
    if (!h.ctx) check(fs_ctx_create(&h.ctx, 0, nullptr), "fs_ctx_create");
    return h.ctx;
}

// advect<T,U>, advect.h:74-76.  T = Vector2<float> (8 B) or Vector3<UQ32> (12 B); V = Vector2<float>.
template <class T, class V>
inline void advect(T *next_p, T *p, V *vel, int dim_x, int dim_y, float dt, bool no_slip)
{
    static_assert(sizeof(V) == sizeof(fs_vec2f), "velocity must be Vector2<float>");
    static_assert(sizeof(T) == sizeof(fs_vec2f) || sizeof(T) == sizeof(fs_rgb_uq32),
                  "payload must be Vector2<float> or Vector3<UQ32>");
    if (sizeof(T) == sizeof(fs_vec2f))
        check(fsh_advect_vec2f((fs_vec2f *)next_p, (const fs_vec2f *)p, (const fs_vec2f *)vel, dim_x, dim_y, dt,
                               no_slip, context()), "advect<Vector2<float>>");
    else
        check(fsh_advect_rgb_uq32((fs_rgb_uq32 *)next_p, (const fs_rgb_uq32 *)p, (const fs_vec2f *)vel, dim_x,
                                  dim_y, dt, no_slip, context()), "advect<Vector3<UQ32>>");
}

// calculate_divergence, finitediff.h:6-7
template <class V>
inline void calculate_divergence(float *div, V *v, int dim_x, int dim_y, float dx)
{
    static_assert(sizeof(V) == sizeof(fs_vec2f), "v must be Vector2<float>");
    check(fsh_calculate_divergence(div, (const fs_vec2f *)v, dim_x, dim_y, dx, context()), "calculate_divergence");
}

// subtract_gradient, finitediff.h:9-10 (v in place)
template <class V>
inline void subtract_gradient(V *v, float *p, int dim_x, int dim_y, float dx)
{
    static_assert(sizeof(V) == sizeof(fs_vec2f), "v must be Vector2<float>");
    check(fsh_subtract_gradient((fs_vec2f *)v, p, dim_x, dim_y, dx, context()), "subtract_gradient");
}

// poisson_solve, poisson.h:4-5
inline void poisson_solve(float *p, float *div, int dim_x, int dim_y, float dx, int iters, float omega)
{
    check(fsh_poisson_solve(p, div, dim_x, dim_y, dx, iters, omega, context()), "poisson_solve");
}

// the whole loop() body, ino:249-289, v and c in place (state stays host-side like the reference's)
template <class V, class C>
inline void step(V *v, C *c, const fs_drag *drags, int n_drags, int dim_x, int dim_y, float dt, float dx,
                 int iters, float omega)
{
    static_assert(sizeof(V) == sizeof(fs_vec2f) && sizeof(C) == sizeof(fs_rgb_uq32), "layout");
    check(fsh_step((fs_vec2f *)v, (fs_rgb_uq32 *)c, drags, n_drags, dim_x, dim_y, dt, dx, iters, omega, nullptr,
                   nullptr, context()), "step");
}

}  // namespace fluid_b200
#endif
