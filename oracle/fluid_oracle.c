/* TEST INFRASTRUCTURE — see fluid_oracle.h.  Plain-C restatement of the
 * reference's sim step; every function cites the reference lines it follows.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: the reference's own
 * results change when contraction is on, SURVEY.md §7 hard part 1).
 *
 * Pinned bit-for-bit against oracle/_ref (the reference's own sources) by
 * tests/test_oracle_vs_ref.py and against tests/golden/ fixtures.  The one
 * exception is oracle_init_color_wheel (ino:196-241): the .ino does not compile
 * off-device, so that function is "parity unpinned".
 */
#include "fluid_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* operations.h:7-9 */
static inline long node(int i, int j, int dim_x) { return (long)dim_x * j + i; }

/* A window of a global grid (same fields as fs_tile, include/fluid_b200.h): local
 * element (lx,ly) at ly*nx+lx is global node (ox+lx, oy+ly).  Wall rules, colour
 * parity and advect coordinates are GLOBAL; a whole grid is ox=oy=0, nx=GX, ny=GY. */
static oracle_tile whole(int dim_x, int dim_y)
{
    oracle_tile t = {dim_x, dim_y, 0, 0, dim_x, dim_y, 0, 0, dim_x, dim_y};
    return t;
}

/* local index of GLOBAL node (gi,gj), or -1 (and *overrun set) when it is outside the window */
static inline long wnode(const oracle_tile *t, int gi, int gj, int *overrun)
{
    int lx = gi - t->ox, ly = gj - t->oy;
    if (lx < 0 || lx >= t->nx || ly < 0 || ly >= t->ny) {
        if (overrun) *overrun = 1;
        return -1;
    }
    return (long)ly * t->nx + lx;
}

/* uq32.h:13 — raw = (uint32_t)(x + 0.5f).  Out-of-range conversion is UB in the
 * reference; the build pins SATURATION (CUDA cvt.rzi.u32.f32, ESP32 utrunc.s,
 * x86 vcvttss2usi): >= 2^32 -> 0xFFFFFFFF, negative/NaN -> 0. */
uint32_t oracle_uq32_from_float(float x)
{
    float y = x + 0.5f;
    if (!(y > 0.0f)) return 0u;
    if (y >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)y;
}

/* uq32.h:15 — round-to-nearest-even to 24 significant bits */
float oracle_uq32_to_float(uint32_t raw) { return (float)raw; }

/* advect.h:13-16: p1*(1-d) + p2*d, (1-d) formed once in float */
static inline float mix(float d, float a, float b)
{
    float w = 1 - d;
    return a * w + b * d;
}

/* no-slip discount, advect.h:61-70 */
static inline float discount(float overshoot)
{
    return overshoot < 0.5 ? (1 - 2 * overshoot) : 0;
}

/* advect.h:24-72, one payload of `nc` channels.  For dye (is_uq) corner values
 * are held as float(raw) and `out_f` is converted by the caller; `exact` is set
 * when the result is a verbatim copy of one node (corner, free-slip) so the
 * caller must not re-round it. */
typedef struct {
    float i_floor, j_floor, di, dj;
    int x_under, x_over, y_under, y_over;
} trace_t;

static inline trace_t classify(float i, float j, int dim_x, int dim_y)
{
    trace_t t;
    t.x_under = i < 0;                    /* advect.h:26-29 */
    t.x_over = i >= dim_x - 1;
    t.y_under = j < 0;
    t.y_over = j >= dim_y - 1;
    t.i_floor = floorf(i);                /* advect.h:34-35 */
    t.j_floor = floorf(j);
    t.di = i - t.i_floor;
    t.dj = j - t.j_floor;
    return t;
}

static inline float overshoot_factor(const trace_t *t, float i, float j, int dim_x,
                                     int dim_y)
{
    float f = 1.0f;                       /* advect.h:62-70 */
    if (t->x_under || t->x_over) {
        float o = t->x_under ? -i : i - (dim_x - 1);
        f *= discount(o);
    }
    if (t->y_under || t->y_over) {
        float o = t->y_under ? -j : j - (dim_y - 1);
        f *= discount(o);
    }
    return f;
}

static inline float wf(const float *p, long n, int ch) { return n < 0 ? 0.0f : p[2 * n + ch]; }
static inline uint32_t wu(const uint32_t *c, long n, int ch) { return n < 0 ? 0u : c[3 * n + ch]; }

static void tile_sample_vec2f(float *out, const float *p, float i, float j, const oracle_tile *w,
                              int no_slip, int *overrun)
{
    int dim_x = w->gdim_x, dim_y = w->gdim_y;
    trace_t t = classify(i, j, dim_x, dim_y);
    int x_oob = t.x_under || t.x_over, y_oob = t.y_under || t.y_over;
    float e[2];
    if (!x_oob && !y_oob) {               /* advect.h:38-42 */
        int gi = (int)t.i_floor, gj = (int)t.j_floor;
        long n11 = wnode(w, gi, gj, overrun), n12 = wnode(w, gi, gj + 1, overrun);
        long n21 = wnode(w, gi + 1, gj, overrun), n22 = wnode(w, gi + 1, gj + 1, overrun);
        for (int ch = 0; ch < 2; ch++)
            out[ch] = mix(t.di, mix(t.dj, wf(p, n11, ch), wf(p, n12, ch)),
                          mix(t.dj, wf(p, n21, ch), wf(p, n22, ch)));
        return;
    }
    if (x_oob && y_oob) {                 /* advect.h:46-48 */
        long n = wnode(w, t.x_under ? 0 : dim_x - 1, t.y_under ? 0 : dim_y - 1, overrun);
        e[0] = wf(p, n, 0);
        e[1] = wf(p, n, 1);
    } else if (x_oob) {                   /* advect.h:49-51 */
        int gi = t.x_under ? 0 : dim_x - 1, gj = (int)t.j_floor;
        long a = wnode(w, gi, gj, overrun), b = wnode(w, gi, gj + 1, overrun);
        for (int ch = 0; ch < 2; ch++) e[ch] = mix(t.dj, wf(p, a, ch), wf(p, b, ch));
    } else {                              /* advect.h:52-54 */
        int gi = (int)t.i_floor, gj = t.y_under ? 0 : dim_y - 1;
        long a = wnode(w, gi, gj, overrun), b = wnode(w, gi + 1, gj, overrun);
        for (int ch = 0; ch < 2; ch++) e[ch] = mix(t.di, wf(p, a, ch), wf(p, b, ch));
    }
    if (!no_slip) {                       /* advect.h:57-59 */
        out[0] = e[0];
        out[1] = e[1];
        return;
    }
    float f = overshoot_factor(&t, i, j, dim_x, dim_y);
    out[0] = e[0] * f;                    /* advect.h:71 */
    out[1] = e[1] * f;
}

void oracle_sample_vec2f(float *out, const float *p, float i, float j, int dim_x,
                         int dim_y, int no_slip)
{
    oracle_tile w = whole(dim_x, dim_y);
    tile_sample_vec2f(out, p, i, j, &w, no_slip, NULL);
}

static void tile_sample_rgb_uq32(uint32_t *out, const uint32_t *c, float i, float j,
                                 const oracle_tile *w, int no_slip, int *overrun)
{
    int dim_x = w->gdim_x, dim_y = w->gdim_y;
    trace_t t = classify(i, j, dim_x, dim_y);
    int x_oob = t.x_under || t.x_over, y_oob = t.y_under || t.y_over;
    uint32_t e[3];
    if (!x_oob && !y_oob) {
        /* inner lerps stay float (TPromoted, advect.h:10-11); one rounding at
         * the final Vector3<float> -> Vector3<UQ32> conversion */
        int gi = (int)t.i_floor, gj = (int)t.j_floor;
        long n11 = wnode(w, gi, gj, overrun), n12 = wnode(w, gi, gj + 1, overrun);
        long n21 = wnode(w, gi + 1, gj, overrun), n22 = wnode(w, gi + 1, gj + 1, overrun);
        for (int ch = 0; ch < 3; ch++)
            out[ch] = oracle_uq32_from_float(
                mix(t.di, mix(t.dj, (float)wu(c, n11, ch), (float)wu(c, n12, ch)),
                    mix(t.dj, (float)wu(c, n21, ch), (float)wu(c, n22, ch))));
        return;
    }
    if (x_oob && y_oob) {                 /* verbatim copy, no rounding */
        long n = wnode(w, t.x_under ? 0 : dim_x - 1, t.y_under ? 0 : dim_y - 1, overrun);
        for (int ch = 0; ch < 3; ch++) e[ch] = wu(c, n, ch);
    } else if (x_oob) {                   /* lerp result rounds INTO T p_edge (advect.h:45,51) */
        int gi = t.x_under ? 0 : dim_x - 1, gj = (int)t.j_floor;
        long a = wnode(w, gi, gj, overrun), b = wnode(w, gi, gj + 1, overrun);
        for (int ch = 0; ch < 3; ch++)
            e[ch] = oracle_uq32_from_float(mix(t.dj, (float)wu(c, a, ch), (float)wu(c, b, ch)));
    } else {
        int gi = (int)t.i_floor, gj = t.y_under ? 0 : dim_y - 1;
        long a = wnode(w, gi, gj, overrun), b = wnode(w, gi + 1, gj, overrun);
        for (int ch = 0; ch < 3; ch++)
            e[ch] = oracle_uq32_from_float(mix(t.di, (float)wu(c, a, ch), (float)wu(c, b, ch)));
    }
    if (!no_slip) {
        for (int ch = 0; ch < 3; ch++) out[ch] = e[ch];
        return;
    }
    /* factor * Vector3<UQ32> -> Vector3<float> -> rounds again on return */
    float f = overshoot_factor(&t, i, j, dim_x, dim_y);
    for (int ch = 0; ch < 3; ch++) out[ch] = oracle_uq32_from_float((float)e[ch] * f);
}

void oracle_sample_rgb_uq32(uint32_t *out, const uint32_t *c, float i, float j,
                            int dim_x, int dim_y, int no_slip)
{
    oracle_tile w = whole(dim_x, dim_y);
    tile_sample_rgb_uq32(out, c, i, j, &w, no_slip, NULL);
}

/* advect.h:74-85.  The reference nests i outer / j inner; the result does not
 * depend on the visiting order (next_p never aliases p or vel), so rows are
 * walked contiguously here.  Returns 1 if a backtrace left the window. */
int oracle_tile_advect_vec2f(float *next_p, const float *p, const float *vel,
                             const oracle_tile *t, float dt, int no_slip)
{
    int overrun = 0;
    for (int ly = t->y0; ly < t->y1; ly++)
        for (int lx = t->x0; lx < t->x1; lx++) {
            long l = (long)ly * t->nx + lx;
            float sx = (float)(t->ox + lx) - vel[2 * l] * dt;        /* advect.h:81 */
            float sy = (float)(t->oy + ly) - vel[2 * l + 1] * dt;
            tile_sample_vec2f(next_p + 2 * l, p, sx, sy, t, no_slip, &overrun);
        }
    return overrun;
}

int oracle_tile_advect_rgb_uq32(uint32_t *next_c, const uint32_t *c, const float *vel,
                                const oracle_tile *t, float dt, int no_slip)
{
    int overrun = 0;
    for (int ly = t->y0; ly < t->y1; ly++)
        for (int lx = t->x0; lx < t->x1; lx++) {
            long l = (long)ly * t->nx + lx;
            float sx = (float)(t->ox + lx) - vel[2 * l] * dt;
            float sy = (float)(t->oy + ly) - vel[2 * l + 1] * dt;
            tile_sample_rgb_uq32(next_c + 3 * l, c, sx, sy, t, no_slip, &overrun);
        }
    return overrun;
}

void oracle_advect_vec2f(float *next_p, const float *p, const float *vel, int dim_x,
                         int dim_y, float dt, int no_slip)
{
    oracle_tile w = whole(dim_x, dim_y);
    oracle_tile_advect_vec2f(next_p, p, vel, &w, dt, no_slip);
}

void oracle_advect_rgb_uq32(uint32_t *next_c, const uint32_t *c, const float *vel,
                            int dim_x, int dim_y, float dt, int no_slip)
{
    oracle_tile w = whole(dim_x, dim_y);
    oracle_tile_advect_rgb_uq32(next_c, c, vel, &w, dt, no_slip);
}

/* finitediff.cpp:9-39.  Interior and wall nodes associate the four terms
 * DIFFERENTLY (fast: (a+b)+(c+d); safe: (((0+a)+b)+c)+d) — kept as is. */
void oracle_tile_calculate_divergence(float *div, const float *v, const oracle_tile *t, float dx)
{
    float two_dx_inv = 1.0f / (2.0f * dx);                 /* finitediff.cpp:36 */
    int i_max = t->gdim_x - 1, j_max = t->gdim_y - 1;
    long pitch = t->nx;
    for (int ly = t->y0; ly < t->y1; ly++)
        for (int lx = t->x0; lx < t->x1; lx++) {
            int i = t->ox + lx, j = t->oy + ly;
            long l = (long)ly * pitch + lx;
            const float *c = v + 2 * l;
            int wall = (i == 0) || (i == i_max) || (j == 0) || (j == j_max);
            float s;
            if (!wall) {                                   /* finitediff.cpp:29 */
                s = (-c[-2] + c[2]) + (-c[-2 * pitch + 1] + c[2 * pitch + 1]);
            } else {                                       /* finitediff.cpp:16-20 */
                s = 0;
                s += (i > 0) ? -c[-2] : c[0];
                s += (i < i_max) ? c[2] : -c[0];
                s += (j > 0) ? -c[-2 * pitch + 1] : c[1];
                s += (j < j_max) ? c[2 * pitch + 1] : -c[1];
            }
            div[l] = s * two_dx_inv;
        }
}

void oracle_calculate_divergence(float *div, const float *v, int dim_x, int dim_y, float dx)
{
    oracle_tile w = whole(dim_x, dim_y);
    oracle_tile_calculate_divergence(div, v, &w, dx);
}

/* finitediff.cpp:41-82, in place; a missing neighbour is replaced by the node's
 * own pressure and the one-sided difference is still scaled by 1/(2dx). */
void oracle_tile_subtract_gradient(float *v, const float *p, const oracle_tile *t, float dx)
{
    float two_dx_inv = 1.0f / (2.0f * dx);                 /* finitediff.cpp:79 */
    int i_max = t->gdim_x - 1, j_max = t->gdim_y - 1;
    long pitch = t->nx;
    for (int ly = t->y0; ly < t->y1; ly++)
        for (int lx = t->x0; lx < t->x1; lx++) {
            int i = t->ox + lx, j = t->oy + ly;
            long l = (long)ly * pitch + lx;
            float pl = (i > 0) ? p[l - 1] : p[l];
            float pr = (i < i_max) ? p[l + 1] : p[l];
            float pd = (j > 0) ? p[l - pitch] : p[l];
            float pu = (j < j_max) ? p[l + pitch] : p[l];
            float gx = (pr - pl) * two_dx_inv;
            float gy = (pu - pd) * two_dx_inv;
            v[2 * l] = v[2 * l] - gx;
            v[2 * l + 1] = v[2 * l + 1] - gy;
        }
}

void oracle_subtract_gradient(float *v, const float *p, int dim_x, int dim_y, float dx)
{
    oracle_tile w = whole(dim_x, dim_y);
    oracle_tile_subtract_gradient(v, p, &w, dx);
}

/* One colour of poisson.cpp:14-61 over the rectangle [x0,x1)x[y0,y1) of a window:
 * parity 0 = GLOBAL (i+j) even (the reference's FIRST pass, on_red=false), parity
 * 1 = odd.  Every update of one colour reads only the other colour, so the
 * visiting order inside a half-sweep is free. */
static void tile_half_sweep(float *p, const float *div, const oracle_tile *t, int x0, int y0,
                            int x1, int y1, float dx, float omega, int parity)
{
    /* poisson.cpp:67 — double literals narrowed to float */
    static const float neg_inv[5] = {0, 0, -1.0 / 2.0, -1.0 / 3.0, -1.0 / 4.0};
    int i_max = t->gdim_x - 1, j_max = t->gdim_y - 1;
    long pitch = t->nx;
    float keep = 1 - omega;                                /* poisson.cpp:98,111 */
    for (int ly = y0; ly < y1; ly++)
        for (int lx = x0 + ((t->ox + x0 + t->oy + ly + parity) & 1); lx < x1; lx += 2) {
            int i = t->ox + lx, j = t->oy + ly;
            long l = (long)ly * pitch + lx;
            float gs;
            if (i > 0 && i < i_max && j > 0 && j < j_max) {    /* poisson.cpp:107-109 */
                float sum = p[l - 1] + p[l + 1] + p[l - pitch] + p[l + pitch];
                gs = -0.25f * (dx * div[l] - sum);
            } else {                                       /* poisson.cpp:69-89 */
                float sum = 0;
                int a = 0;
                if (i > 0) { sum += p[l - 1]; a++; }
                if (i < i_max) { sum += p[l + 1]; a++; }
                if (j > 0) { sum += p[l - pitch]; a++; }
                if (j < j_max) { sum += p[l + pitch]; a++; }
                gs = neg_inv[a] * (dx * div[l] - sum);
            }
            p[l] = keep * p[l] + omega * gs;
        }
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* `n_half` consecutive half-sweeps starting with colour first_parity, p_in -> p_out
 * (p_in == NULL: zero), valid on the rectangle when p_in is valid on the rectangle
 * grown by n_half (clipped to the global grid): sweep s runs on the rectangle grown
 * by n_half-1-s.  Same contract as fs_tile_sor_sweeps. */
void oracle_tile_sor_sweeps(float *p_out, const float *p_in, const float *div,
                            const oracle_tile *t, float dx, float omega, int first_parity,
                            int n_half)
{
    int sx0 = clampi(t->x0 - n_half, 0, t->nx), sx1 = clampi(t->x1 + n_half, 0, t->nx);
    int sy0 = clampi(t->y0 - n_half, 0, t->ny), sy1 = clampi(t->y1 + n_half, 0, t->ny);
    for (int ly = sy0; ly < sy1; ly++)
        for (int lx = sx0; lx < sx1; lx++) {
            long l = (long)ly * t->nx + lx;
            p_out[l] = p_in ? p_in[l] : 0.0f;
        }
    for (int s = 0; s < n_half; s++) {
        int r = n_half - 1 - s;
        tile_half_sweep(p_out, div, t, clampi(t->x0 - r, 0, t->nx), clampi(t->y0 - r, 0, t->ny),
                        clampi(t->x1 + r, 0, t->nx), clampi(t->y1 + r, 0, t->ny), dx, omega,
                        (first_parity + s) & 1);
    }
}

void oracle_sor_half_sweep(float *p, const float *div, int dim_x, int dim_y, float dx,
                           float omega, int parity)
{
    oracle_tile w = whole(dim_x, dim_y);
    tile_half_sweep(p, div, &w, 0, 0, dim_x, dim_y, dx, omega, parity);
}

/* poisson.cpp:114-125: no warm start; iters x (even colour, then odd colour) */
void oracle_poisson_solve(float *p, const float *div, int dim_x, int dim_y, float dx,
                          int iters, float omega)
{
    memset(p, 0, sizeof(float) * (size_t)dim_x * dim_y);
    for (int k = 0; k < iters; k++) {
        oracle_sor_half_sweep(p, div, dim_x, dim_y, dx, omega, 0);
        oracle_sor_half_sweep(p, div, dim_x, dim_y, dx, omega, 1);
    }
}

/* ino:264-269: queue order, SET not add, x/y swapped.  The reference does not
 * bounds-check; out-of-range records are dropped here (and in the product).  On a
 * window only the records that land in the rectangle are applied. */
void oracle_tile_apply_drags(float *v, const oracle_drag *drags, int n, const oracle_tile *t)
{
    for (int k = 0; k < n; k++) {
        if (drags[k].cy >= t->gdim_x || drags[k].cx >= t->gdim_y) continue;
        int lx = drags[k].cy - t->ox, ly = drags[k].cx - t->oy;
        if (lx < t->x0 || lx >= t->x1 || ly < t->y0 || ly >= t->y1) continue;
        long l = (long)ly * t->nx + lx;
        v[2 * l] = drags[k].vy;
        v[2 * l + 1] = drags[k].vx;
    }
}

void oracle_apply_drags(float *v, const oracle_drag *drags, int n, int dim_x, int dim_y)
{
    oracle_tile w = whole(dim_x, dim_y);
    oracle_tile_apply_drags(v, drags, n, &w);
}

/* ino:249-289 */
void oracle_step(float *v, uint32_t *c, const oracle_drag *drags, int n_drags,
                 int dim_x, int dim_y, float dt, float dx, int iters, float omega,
                 float *p_out, float *div_out)
{
    size_t n = (size_t)dim_x * dim_y;
    float *v_next = (float *)malloc(n * 2 * sizeof(float));
    oracle_advect_vec2f(v_next, v, v, dim_x, dim_y, dt, 1);
    memcpy(v, v_next, n * 2 * sizeof(float));
    free(v_next);

    oracle_apply_drags(v, drags, n_drags, dim_x, dim_y);

    float *d = div_out ? div_out : (float *)malloc(n * sizeof(float));
    float *p = p_out ? p_out : (float *)malloc(n * sizeof(float));
    oracle_calculate_divergence(d, v, dim_x, dim_y, dx);
    oracle_poisson_solve(p, d, dim_x, dim_y, dx, iters, omega);
    oracle_subtract_gradient(v, p, dim_x, dim_y, dx);
    if (!div_out) free(d);
    if (!p_out) free(p);

    uint32_t *c_next = (uint32_t *)malloc(n * 3 * sizeof(uint32_t));
    oracle_advect_rgb_uq32(c_next, c, v, dim_x, dim_y, dt, 0);
    memcpy(c, c_next, n * 3 * sizeof(uint32_t));
    free(c_next);
}

/* ino:116-177.  Cell (i,j), i < dim_x-1, j < dim_y-1, becomes 4x4 pixels at image
 * rows 4i..4i+3 (along the sim's FAST axis) and columns 4j..4j+3; row pitch =
 * 4*(dim_y-1).  The ramps are accumulated (c += dc), not c + k*dc, and the left
 * column of cell j>0 is the previous cell's right column — which is the same
 * arithmetic on the same two nodes, so it is simply recomputed here. */
void oracle_upscale4_rgb565(uint16_t *out, const uint32_t *c, int dim_x, int dim_y)
{
    long pitch = 4L * (dim_y - 1);
    for (int i = 0; i < dim_x - 1; i++)
        for (int j = 0; j < dim_y - 1; j++) {
            long n11 = node(i, j, dim_x), n12 = node(i, j + 1, dim_x);
            long n21 = node(i + 1, j, dim_x), n22 = node(i + 1, j + 1, dim_x);
            uint32_t px[4][4][3];
            for (int ch = 0; ch < 3; ch++) {
                float left[4], right[4];
                float a = (float)c[3 * n11 + ch];
                float da = ((float)c[3 * n21 + ch] - a) * 0.25f;      /* ino:133-138 */
                for (int ii = 0; ii < 4; ii++) { left[ii] = a; a += da; }
                float b = (float)c[3 * n12 + ch];
                float db = ((float)c[3 * n22 + ch] - b) * 0.25f;      /* ino:147-152 */
                for (int ii = 0; ii < 4; ii++) { right[ii] = b; b += db; }
                for (int ii = 0; ii < 4; ii++) {                      /* ino:155-162 */
                    float r = left[ii];
                    float dr = (right[ii] - r) * 0.25f;
                    for (int jj = 0; jj < 4; jj++) {
                        px[ii][jj][ch] = oracle_uq32_from_float(r);   /* ino:168 */
                        r += dr;
                    }
                }
            }
            for (int ii = 0; ii < 4; ii++)
                for (int jj = 0; jj < 4; jj++) {
                    uint16_t w = (uint16_t)(((px[ii][jj][0] & 0xF8000000u) >> 16) |
                                            ((px[ii][jj][1] & 0xFC000000u) >> 21) |
                                            ((px[ii][jj][2] & 0xF8000000u) >> 27));
                    w = (uint16_t)((w << 8) | (w >> 8));              /* ino:173 */
                    out[(4L * i + ii) * pitch + 4L * j + jj] = w;
                }
        }
}

/* ino:196-241 (PARITY UNPINNED — the sketch cannot compile off-device).  Zero
 * velocity; three-sector colour wheel; two in-place, order-dependent 1-2-1
 * smoothing passes (first along j, then along i, both walked i-outer/j-inner so
 * the "previous" neighbour is already smoothed). */
void oracle_init_color_wheel(float *v, uint32_t *c, int dim_x, int dim_y)
{
    const double pi = 3.1415926535897932384626433832795;   /* Arduino PI */
    size_t n = (size_t)dim_x * dim_y;
    memset(v, 0, n * 2 * sizeof(float));
    int ci = dim_x / 2, cj = dim_y / 2;
    uint32_t full = oracle_uq32_from_float((float)UINT32_MAX);
    for (int i = 0; i < dim_x; i++)
        for (int j = 0; j < dim_y; j++) {
            float ang = atan2f((float)(-(i - ci)), (float)(j - cj));
            uint32_t *q = c + 3 * node(i, j, dim_x);
            q[0] = q[1] = q[2] = oracle_uq32_from_float(0.0f);
            if (ang < -pi / 3) q[0] = full;
            else if (ang >= -pi / 3 && ang < pi / 3) q[1] = full;
            else q[2] = full;
        }
    for (int i = 0; i < dim_x; i++)
        for (int j = 0; j < dim_y; j++)
            for (int ch = 0; ch < 3; ch++) {
                uint32_t mid = c[3 * node(i, j, dim_x) + ch];
                uint32_t lo = (j == 0) ? mid : c[3 * node(i, j - 1, dim_x) + ch];
                uint32_t hi = (j == dim_y - 1) ? mid : c[3 * node(i, j + 1, dim_x) + ch];
                float s = 0.25f * (float)lo + 0.5f * (float)mid + 0.25f * (float)hi;
                c[3 * node(i, j, dim_x) + ch] = oracle_uq32_from_float(s);
            }
    for (int i = 0; i < dim_x; i++)
        for (int j = 0; j < dim_y; j++)
            for (int ch = 0; ch < 3; ch++) {
                uint32_t mid = c[3 * node(i, j, dim_x) + ch];
                uint32_t lo = (i == 0) ? mid : c[3 * node(i - 1, j, dim_x) + ch];
                uint32_t hi = (i == dim_x - 1) ? mid : c[3 * node(i + 1, j, dim_x) + ch];
                float s = 0.25f * (float)lo + 0.5f * (float)mid + 0.25f * (float)hi;
                c[3 * node(i, j, dim_x) + ch] = oracle_uq32_from_float(s);
            }
}

uint64_t oracle_fnv1a64(const void *data, uint64_t nbytes)
{
    const unsigned char *b = (const unsigned char *)data;
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t k = 0; k < nbytes; k++) {
        h ^= b[k];
        h *= 0x100000001b3ull;
    }
    return h;
}
