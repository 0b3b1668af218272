/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference's stable-
 * fluids step.  Not part of the product path: only tests/, __graft_entry__.
 * smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pin: the reference ships no tests or golden vectors, so
 * this restatement is pinned against the reference's OWN sources compiled on
 * the host (oracle/_ref/libfluid_ref.so, see oracle/Makefile) — bit-for-bit on
 * every field — and against fixtures generated from that library
 * (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Layout contract (operations.h:7-9, vector.h, uq32.h): dense, unpadded,
 * node (i,j) at ij = dim_x*j + i; velocity = float[2] AoS, dye = uint32[3] AoS.
 */
#ifndef FLUID_ORACLE_H
#define FLUID_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {            /* struct drag, ino:45-48 */
    uint16_t cx, cy;        /* coords.x (column, slow axis j), coords.y (row, fast axis i) */
    float vx, vy;           /* velocity.x, velocity.y (graphics frame) */
} oracle_drag;

typedef struct {            /* same fields as fs_tile (include/fluid_b200.h) */
    int gdim_x, gdim_y, ox, oy, nx, ny, x0, y0, x1, y1;
} oracle_tile;

uint32_t oracle_uq32_from_float(float x);   /* uq32.h:13, saturating */
float oracle_uq32_to_float(uint32_t raw);   /* uq32.h:15 */

void oracle_sample_vec2f(float *out, const float *p, float i, float j, int dim_x,
                         int dim_y, int no_slip);              /* advect.h:24-72 */
void oracle_sample_rgb_uq32(uint32_t *out, const uint32_t *c, float i, float j,
                            int dim_x, int dim_y, int no_slip);

void oracle_advect_vec2f(float *next_p, const float *p, const float *vel, int dim_x,
                         int dim_y, float dt, int no_slip);    /* advect.h:74-85 */
void oracle_advect_rgb_uq32(uint32_t *next_c, const uint32_t *c, const float *vel,
                            int dim_x, int dim_y, float dt, int no_slip);
void oracle_calculate_divergence(float *div, const float *v, int dim_x, int dim_y,
                                 float dx);                    /* finitediff.cpp:9-39 */
void oracle_subtract_gradient(float *v, const float *p, int dim_x, int dim_y,
                              float dx);                       /* finitediff.cpp:41-82 */
void oracle_sor_half_sweep(float *p, const float *div, int dim_x, int dim_y, float dx,
                           float omega, int parity);           /* poisson.cpp:14-112 */
void oracle_poisson_solve(float *p, const float *div, int dim_x, int dim_y, float dx,
                          int iters, float omega);             /* poisson.cpp:114-125 */
void oracle_apply_drags(float *v, const oracle_drag *drags, int n, int dim_x,
                        int dim_y);                            /* ino:264-269 */
void oracle_step(float *v, uint32_t *c, const oracle_drag *drags, int n_drags,
                 int dim_x, int dim_y, float dt, float dx, int iters, float omega,
                 float *p_out, float *div_out);                /* ino:249-289 */
void oracle_upscale4_rgb565(uint16_t *out, const uint32_t *c, int dim_x,
                            int dim_y);                        /* ino:116-177 */
void oracle_init_color_wheel(float *v, uint32_t *c, int dim_x, int dim_y);
                                                               /* ino:196-241 */
/* the same operators over a window of a global grid (for the decomposed path);
 * the whole-grid functions above are the ox=oy=0 special case of these */
int oracle_tile_advect_vec2f(float *next_p, const float *p, const float *vel,
                             const oracle_tile *t, float dt, int no_slip);
int oracle_tile_advect_rgb_uq32(uint32_t *next_c, const uint32_t *c, const float *vel,
                                const oracle_tile *t, float dt, int no_slip);
void oracle_tile_calculate_divergence(float *div, const float *v, const oracle_tile *t, float dx);
void oracle_tile_subtract_gradient(float *v, const float *p, const oracle_tile *t, float dx);
void oracle_tile_sor_sweeps(float *p_out, const float *p_in, const float *div,
                            const oracle_tile *t, float dx, float omega, int first_parity,
                            int n_half);
void oracle_tile_apply_drags(float *v, const oracle_drag *drags, int n, const oracle_tile *t);

uint64_t oracle_fnv1a64(const void *data, uint64_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
