// TEST INFRASTRUCTURE — not part of the product path.
//
// extern "C" shim over the UNMODIFIED reference sim core, compiled from the
// sources where they lie under /root/reference (never copied into this repo):
//   advect.h, vector.h, uq32.h, operations.h (header-only, included below)
//   finitediff.cpp, poisson.cpp              (compiled beside this file)
// Built by oracle/Makefile into oracle/_ref/libfluid_ref.so with
//   g++ -O2 -ffp-contract=off -mavx512f
// (contraction off => results independent of -O level / ISA; avx512f => the
// out-of-range float->uint32 in uq32.h:13 saturates like CUDA's cvt.rzi.u32.f32
// and the ESP32 FPU instead of wrapping).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load the resulting library.
//
// The only thing restated here (because the .ino cannot compile off-device) is
// the step ORDER of loop() — ino:249-289 — around calls into the reference's
// own operators.

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "advect.h"
#include "finitediff.h"
#include "poisson.h"
#include "uq32.h"
#include "vector.h"

static_assert(sizeof(Vector2<float>) == 8, "Vector2<float> must be 8 B AoS");
static_assert(sizeof(Vector3<UQ32>) == 12, "Vector3<UQ32> must be 12 B AoS");

struct ref_drag {  // same layout as `struct drag`, ino:45-48
    Vector2<uint16_t> coords;
    Vector2<float> velocity;
};
static_assert(sizeof(ref_drag) == 12, "struct drag is 12 B");

extern "C" {

// 1 when this host saturates out-of-range float->uint32 (needs AVX-512F at run time)
int ref_saturates(void) { return __builtin_cpu_supports("avx512f") ? 1 : 0; }

// advect.h:74-85 with T = Vector2<float>, U = float (ino:253)
void ref_advect_vec2f(float *next_p, float *p, float *vel, int dim_x, int dim_y,
                      float dt, int no_slip)
{
    advect<Vector2<float>, float>((Vector2<float> *)next_p, (Vector2<float> *)p,
                                  (Vector2<float> *)vel, dim_x, dim_y, dt,
                                  no_slip != 0);
}

// advect.h:74-85 with T = Vector3<UQ32>, U = float (ino:282)
void ref_advect_rgb_uq32(uint32_t *next_c, uint32_t *c, float *vel, int dim_x,
                         int dim_y, float dt, int no_slip)
{
    advect<Vector3<UQ32>, float>((Vector3<UQ32> *)next_c, (Vector3<UQ32> *)c,
                                 (Vector2<float> *)vel, dim_x, dim_y, dt,
                                 no_slip != 0);
}

// sample() itself (advect.h:24-72), for edge-semantics known-answer tests
void ref_sample_vec2f(float *out, float *p, float i, float j, int dim_x,
                      int dim_y, int no_slip)
{
    Vector2<float> r = sample((Vector2<float> *)p, i, j, dim_x, dim_y, no_slip != 0);
    out[0] = r.x;
    out[1] = r.y;
}

void ref_sample_rgb_uq32(uint32_t *out, uint32_t *c, float i, float j, int dim_x,
                         int dim_y, int no_slip)
{
    Vector3<UQ32> r = sample((Vector3<UQ32> *)c, i, j, dim_x, dim_y, no_slip != 0);
    out[0] = r.x.raw;
    out[1] = r.y.raw;
    out[2] = r.z.raw;
}

// uq32.h:13 and uq32.h:15
uint32_t ref_uq32_from_float(float x) { return UQ32(x).raw; }
float ref_uq32_to_float(uint32_t raw)
{
    UQ32 q;
    q.raw = raw;
    return (float)q;
}

// finitediff.cpp:33-39
void ref_calculate_divergence(float *div, float *v, int dim_x, int dim_y, float dx)
{
    calculate_divergence(div, (Vector2<float> *)v, dim_x, dim_y, dx);
}

// finitediff.cpp:75-82
void ref_subtract_gradient(float *v, float *p, int dim_x, int dim_y, float dx)
{
    subtract_gradient((Vector2<float> *)v, p, dim_x, dim_y, dx);
}

// poisson.cpp:114-125
void ref_poisson_solve(float *p, float *div, int dim_x, int dim_y, float dx,
                       int iters, float omega)
{
    poisson_solve(p, div, dim_x, dim_y, dx, iters, omega);
}

// ino:264-269 (queue drained in order; SET not add; x/y swapped). The reference
// has no bounds check (a latent OOB write); here out-of-range records are
// dropped so the shim cannot corrupt the heap — in-range behaviour is identical.
void ref_apply_drags(float *v, const void *drags, int n, int dim_x, int dim_y)
{
    const ref_drag *d = (const ref_drag *)drags;
    Vector2<float> *vf = (Vector2<float> *)v;
    for (int k = 0; k < n; k++) {
        ref_drag msg = d[k];
        if (msg.coords.y >= dim_x || msg.coords.x >= dim_y) continue;
        int ij = index(msg.coords.y, msg.coords.x, dim_x);
        Vector2<float> swapped(msg.velocity.y, msg.velocity.x);
        vf[ij] = swapped;
    }
}

// loop(), ino:249-289: advect v (no_slip) -> drags -> divergence -> SOR ->
// gradient-subtract -> advect dye (free-slip sampling). v and c are updated in
// place (the reference ping-pongs pointers; the caller-visible effect is the
// same). p_out/div_out (each dim_x*dim_y floats) receive the step's last
// pressure and divergence fields; either may be NULL.
void ref_step(float *v, uint32_t *c, const void *drags, int n_drags, int dim_x,
              int dim_y, float dt, float dx, int iters, float omega,
              float *p_out, float *div_out)
{
    size_t n = (size_t)dim_x * dim_y;
    Vector2<float> *vf = (Vector2<float> *)v;

    Vector2<float> *v_temp = new Vector2<float>[n];                    // ino:252
    advect(v_temp, vf, vf, dim_x, dim_y, dt, true);                    // ino:253
    memcpy(vf, v_temp, n * sizeof(Vector2<float>));                    // ino:255
    delete[] v_temp;                                                   // ino:256

    ref_apply_drags(v, drags, n_drags, dim_x, dim_y);                  // ino:264-269

    float *div_v = div_out ? div_out : new float[n];                   // ino:272
    float *p = p_out ? p_out : new float[n];                           // ino:273
    calculate_divergence(div_v, vf, dim_x, dim_y, dx);                 // ino:274
    poisson_solve(p, div_v, dim_x, dim_y, dx, iters, omega);           // ino:275
    subtract_gradient(vf, p, dim_x, dim_y, dx);                        // ino:276
    if (!div_out) delete[] div_v;                                      // ino:277
    if (!p_out) delete[] p;                                            // ino:278

    Vector3<UQ32> *cf = (Vector3<UQ32> *)c;
    Vector3<UQ32> *c_temp = new Vector3<UQ32>[n];                      // ino:281
    advect(c_temp, cf, vf, dim_x, dim_y, dt, false);                   // ino:282
    memcpy(cf, c_temp, n * sizeof(Vector3<UQ32>));                     // ino:286
    delete[] c_temp;                                                   // ino:287
}

// SURVEY.md §8(c) fact 6 input recipe: glibc srand(seed); per node in index
// order v.x, v.y = (rand()/(float)RAND_MAX - 0.5f)*200, then c.{x,y,z}.raw =
// rand()*2u.  Lives here (not in Python) because it depends on glibc rand().
void ref_fill_rand(float *v, uint32_t *c, int dim_x, int dim_y, unsigned seed)
{
    srand(seed);
    size_t n = (size_t)dim_x * dim_y;
    for (size_t k = 0; k < n; k++) {
        v[2 * k + 0] = (rand() / (float)RAND_MAX - 0.5f) * 200;
        v[2 * k + 1] = (rand() / (float)RAND_MAX - 0.5f) * 200;
        c[3 * k + 0] = rand() * 2u;
        c[3 * k + 1] = rand() * 2u;
        c[3 * k + 2] = rand() * 2u;
    }
}

}  // extern "C"
