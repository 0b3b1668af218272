/* fluid_b200.h — C ABI of the B200-native stable-fluids step.
 *
 * Drop-in boundary for the sim-step operators of colonelwatch/ESP32-fluid-
 * simulation.  The reference has no FFI layer; its de-facto boundary is five
 * free functions over caller-owned dense arrays, called only from loop()
 * (ESP32-fluid-simulation.ino:249-289).  Every entry point below names the
 * reference interface it replaces.  Signatures keep the reference's argument
 * order and meaning; the only additions are an `int` status return and a
 * trailing opaque context (device + stream + scratch).
 *
 * Layout contract (identical to the reference, operations.h:7-9, vector.h,
 * uq32.h): dense, unpadded, node (i,j) at ij = dim_x*j + i (i is the fast
 * axis); velocity = {float x,y} (8 B), dye = {uint32 r,g,b} UQ32 raw words
 * (12 B, align 4), scalars = float.  Caller allocates and frees everything.
 *
 *   fs_*   : pointers are DEVICE pointers; work is enqueued on the context's
 *            stream and the call returns without synchronising.
 *   fsh_*  : pointers are HOST pointers (what reference code holds today); the
 *            call copies in, runs the same kernels, copies out and returns when
 *            the result is in host memory — a literal drop-in for the reference
 *            call of the same name.
 *
 * Results are bit-identical to the reference compiled with FMA contraction off
 * (g++ -O2 -ffp-contract=off), for floats as well as for UQ32 words.  The one
 * defined deviation: float->UQ32 conversion SATURATES (>= 2^32 -> 0xFFFFFFFF,
 * negative/NaN -> 0) where uq32.h:13 is undefined behaviour.
 *
 * There is no CPU fallback: every call fails with a CUDA error code when no
 * sm_100 device is usable.
 */
#ifndef FLUID_B200_H
#define FLUID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS_OK 0
#define FS_ERR_INVALID_ARG (-1)   /* NULL pointer, dim < 2, iters < 0, aliasing violation */
#define FS_ERR_NO_CONTEXT (-2)
#define FS_ERR_UNSUPPORTED (-3)
#define FS_ERR_HALO_OVERRUN (-4)  /* decomposed advect: backtrace left the local window */
#define FS_ERR_WOULD_BLOCK (-6)   /* fs_sim: both frame slots wait for the consumer / no frame produced yet */
#define FS_ERR_HALO_TIMEOUT (-5)  /* fs_halo_exchange: a neighbour never signalled (see "halo_timeout_ms") */
/* positive values are cudaError_t codes */

typedef struct fs_vec2f { float x, y; } fs_vec2f;             /* Vector2<float>, vector.h:4-57 */
typedef struct fs_rgb_uq32 { uint32_t r, g, b; } fs_rgb_uq32; /* Vector3<UQ32>, vector.h:63-122 + uq32.h:8-16 */
typedef struct fs_drag {                                      /* struct drag, ino:45-48 */
    uint16_t cx, cy;   /* coords.x = column (slow axis j), coords.y = row (fast axis i) */
    float vx, vy;      /* velocity.x, velocity.y in the graphics frame (swapped on apply) */
} fs_drag;

typedef struct fs_ctx fs_ctx;   /* opaque: device, stream, scratch buffers, TMA descriptors */

/* ---- context -------------------------------------------------------------- */
/* `stream` is a cudaStream_t (NULL = the legacy default stream), or FS_STREAM_NEW: the context creates
 * (and owns) a non-blocking stream — for hosts that do not link the CUDA runtime themselves. */
#define FS_STREAM_NEW ((void *)(intptr_t)-1)
int  fs_ctx_create(fs_ctx **out, int device, void *stream);
int  fs_ctx_destroy(fs_ctx *ctx);
int  fs_ctx_synchronize(fs_ctx *ctx);
/* Kernel-variant switches for A/B measurement (all variants are bit-identical):
 *   "sor"    : 0 = one half-sweep per launch, 1 = temporally blocked in registers + shared memory (default)
 *   "sor_t"  : full iterations fused per HBM round trip (1..8, default 6)
 *   "sor_shape": CTA region and loader of the blocked solver: 0 = 128x96 nodes, 1 = 128x192, both
 *              loaded straight from global; 2 / 3 = the same regions with persistent CTAs whose
 *              next tile is prefetched by TMA (fall back to 1 when rows are
 *              not 16-byte multiples); 4 / 5 / 6 / 7 = TMA variants with 12x12, 16x10, 12x16,
 *              16x11 (warps x rows per warp) strips; 7 is the default (spill-free, best wave count at 4096^2)
 *   "sor_one_launch": 1 = all passes of a solve run in ONE persistent launch with tile-level
 *              dependencies between passes (shapes 2, 3, 5); 0 (default) = one launch per pass.
 *              Measured: the dependency probes and release fences on the issuing thread cost more
 *              than the launch gaps and per-pass tails they remove (0.98 vs 0.87 ms at 4096^2)
 *   "halo_timeout_ms": how long fs_halo_exchange waits for a neighbour's flag before raising
 *              FS_ERR_HALO_TIMEOUT in the context's status (reported by fs_tile_check); default 10000, 0 = forever
 *   "advect" : 0 = direct L1/L2 gather, 1 = TMA-staged shared-memory tile (default where legal)
 *   "fuse"   : bit mask for fs_step: 1 = drags + divergence folded into the velocity advect, 2 =
 *              gradient-subtract folded into the dye advect (measured slower than the stand-alone
 *              gradient kernel, which runs at the HBM roofline), 4 = the RGB565 frame rendered inside
 *              the dye advect (fs_step_frame, fs_advect_rgb_frame, fs_dist with frame = 1);
 *              0 = one kernel per operator; default 5
 *   "sor_grid_limit": cap on the persistent SOR grid, 0 (default) = one CTA per SM; used when several
 *              emulated ranks share one device
 *   "e2e_bands": row bands the dye of fsh_step travels in, 1..16 (default 8; fewer when the grid has
 *              under 128 rows per band): band b is advected and sent home once bands 0..b+1 have
 *              arrived, so both PCIe directions stay busy.  A backtrace that reaches a band still
 *              in flight is detected and the dye advect redone in one piece ("e2e_redos",
 *              read-only, counts those calls)
 *   "ensemble": kernel behind fs_ensemble_step*: 0 (default) = automatic — the register-tiled projection
 *              (csrc/ensemble_reg.cuh), first-generation kernel for shapes it does not take; 1-4 = first
 *              generation with the dye streamed through L1/L2, 5 = first generation, automatic;
 *              6/7/8/9 = register-tiled with 2/4/6/8 rows per thread; 12/14/16/18 = the same, dye streamed;
 *              20/21 = automatic with the bulk-copy pipelined state I/O forced on / off (default: on for calls
 *              of up to 3 steps)
 *   "num_sms": (read-only) SMs of the context's device */
int  fs_ctx_set_option(fs_ctx *ctx, const char *name, int value);
int  fs_ctx_get_option(fs_ctx *ctx, const char *name, int *value);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t fs_ctx_launch_count(fs_ctx *ctx);
const char *fs_version(void);
const char *fs_error_string(int code);

/* pinned host memory for the fsh_* path / harness */
int  fs_host_alloc(void **out, size_t bytes);
int  fs_host_free(void *p);

/* ---- sim-step operators, device pointers ----------------------------------- */
/* advect<Vector2<float>,float>, advect.h:74-85 (ino:253).  next_p must not alias p/vel; p may alias vel. */
int fs_advect_vec2f(fs_vec2f *next_p, const fs_vec2f *p, const fs_vec2f *vel,
                    int dim_x, int dim_y, float dt, int no_slip, fs_ctx *ctx);
/* advect<Vector3<UQ32>,float>, advect.h:74-85 (ino:282). */
int fs_advect_rgb_uq32(fs_rgb_uq32 *next_c, const fs_rgb_uq32 *c, const fs_vec2f *vel,
                       int dim_x, int dim_y, float dt, int no_slip, fs_ctx *ctx);
/* calculate_divergence, finitediff.h:6-7 / finitediff.cpp:33-39. */
int fs_calculate_divergence(float *div, const fs_vec2f *v, int dim_x, int dim_y,
                            float dx, fs_ctx *ctx);
/* subtract_gradient, finitediff.h:9-10 / finitediff.cpp:75-82 (v in place). */
int fs_subtract_gradient(fs_vec2f *v, const float *p, int dim_x, int dim_y, float dx,
                         fs_ctx *ctx);
/* poisson_solve, poisson.h:4-5 / poisson.cpp:114-125 (p overwritten from zero; p != div). */
int fs_poisson_solve(float *p, const float *div, int dim_x, int dim_y, float dx,
                     int iters, float omega, fs_ctx *ctx);
/* one colour of domain_iter_red_black, poisson.cpp:14-61 (parity 0 = (i+j) even = the reference's first pass). */
int fs_sor_half_sweep(float *p, const float *div, int dim_x, int dim_y, float dx,
                      float omega, int parity, fs_ctx *ctx);
/* Residual of the system poisson_solve relaxes (not in the reference; a convergence probe):
 * r_ij = gs_ij(p) - p_ij with gs the Gauss-Seidel value of poisson.cpp:63-90,101-109.
 * *max_abs = max |r| (exact), *l2 = sqrt(sum r^2) accumulated in double; reduced on the device
 * with warp shuffles.  Results land in HOST memory; synchronises the stream. */
int fs_poisson_residual(float *max_abs, double *l2, const float *p, const float *div, int dim_x,
                        int dim_y, float dx, fs_ctx *ctx);
/* drag overwrite, ino:264-269.  `drags` is a HOST array (it is the touch queue);
 * records outside the grid are dropped (the reference writes out of bounds). */
int fs_apply_drags(fs_vec2f *v, const fs_drag *drags, int n, int dim_x, int dim_y,
                   fs_ctx *ctx);
/* loop() body, ino:249-289: advect v (no-slip) -> drags -> divergence -> SOR ->
 * gradient-subtract -> advect dye (free-slip).  v and c are updated in place.
 * p_out/div_out (dim_x*dim_y floats each, device) receive the step's pressure
 * and divergence, or may be NULL. */
int fs_step(fs_vec2f *v, fs_rgb_uq32 *c, const fs_drag *drags, int n_drags,
            int dim_x, int dim_y, float dt, float dx, int iters, float omega,
            float *p_out, float *div_out, fs_ctx *ctx);
/* The first three operators of loop() in ONE pass over the grid (what fs_step runs): v_out = advect(v_in)
 * with no-slip walls (ino:253), the drag records applied to v_out in queue order (ino:264-269), div =
 * calculate_divergence(v_out) (ino:274).  v_out must not alias v_in; `drags` is a HOST array. */
int fs_advect_drags_divergence(fs_vec2f *v_out, float *div, const fs_vec2f *v_in, const fs_drag *drags,
                               int n_drags, int dim_x, int dim_y, float dt, float dx, fs_ctx *ctx);
/* The same loop() body with the reference's dye pointer swap (ino:281-287) left to the caller:
 * the dye is read from c_in and the advected dye written to c_out (distinct buffers; the caller
 * alternates them from step to step like SWAP(c_temp, color_field)).  Saves fs_step's copy-back
 * of the dye (12 B/node read + written). */
int fs_step_pingpong(fs_vec2f *v, const fs_rgb_uq32 *c_in, fs_rgb_uq32 *c_out, const fs_drag *drags,
                     int n_drags, int dim_x, int dim_y, float dt, float dx, int iters, float omega,
                     float *p_out, float *div_out, fs_ctx *ctx);
/* Dye advect (ino:282) with the 4x RGB565 frame of the ADVECTED dye (draw_routine arithmetic,
 * ino:116-177) produced in the same kernel: the frame is rendered from the tile while it is still in
 * shared memory instead of re-reading the dye (fuse bit 4; otherwise advect + fs_upscale4_rgb565).
 * frame layout as fs_upscale4_rgb565. */
int fs_advect_rgb_frame(fs_rgb_uq32 *next_c, uint16_t *frame, const fs_rgb_uq32 *c, const fs_vec2f *vel,
                        int dim_x, int dim_y, float dt, int no_slip, fs_ctx *ctx);
/* fs_step_pingpong that also hands the frame of the new dye to the display side (the sketch's
 * color_produced hand-off, ino:285-288): loop() + draw_routine's arithmetic in one call. */
int fs_step_frame(fs_vec2f *v, const fs_rgb_uq32 *c_in, fs_rgb_uq32 *c_out, uint16_t *frame,
                  const fs_drag *drags, int n_drags, int dim_x, int dim_y, float dt, float dx, int iters,
                  float omega, float *p_out, float *div_out, fs_ctx *ctx);
/* draw_routine arithmetic, ino:116-177: 4x bilinear upscale of the dye, UQ32
 * round, RGB565 pack, byte swap.  out is (dim_x-1)*4 rows x (dim_y-1)*4 columns
 * of uint16, row pitch (dim_y-1)*4 (image rows run along the sim's fast axis). */
int fs_upscale4_rgb565(uint16_t *out, const fs_rgb_uq32 *c, int dim_x, int dim_y,
                       fs_ctx *ctx);
/* `batch` independent grids, `n_steps` consecutive loop() bodies each, one CTA per
 * grid with the grid's velocity / divergence / pressure resident in shared memory for the
 * whole call and its dye streamed through L2 (BASELINE.json configs[1]).  v, c hold the
 * grids back to back.  drags: HOST array laid out [n_steps][batch][max_drags],
 * drag_counts: HOST array [n_steps][batch] (both may be NULL with max_drags = 0).
 * Needs 16*dim_x*dim_y bytes of shared memory (two grids per SM up to ~7,000 nodes):
 * FS_ERR_UNSUPPORTED beyond 6,144 nodes (use fs_step per grid there). */
int fs_ensemble_step(fs_vec2f *v, fs_rgb_uq32 *c, const fs_drag *drags,
                     const int *drag_counts, int max_drags, int batch, int dim_x,
                     int dim_y, float dt, float dx, int iters, float omega,
                     int n_steps, fs_ctx *ctx);

/* fs_ensemble_step with the drag records already ON THE DEVICE ([n_steps][batch][max_drags] records,
 * [n_steps][batch] counts) — e.g. written by fs_touch_to_drags — so an ensemble is driven without host
 * round trips. */
int fs_ensemble_step_dev(fs_vec2f *v, fs_rgb_uq32 *c, const fs_drag *drags_dev, const int *counts_dev,
                         int max_drags, int batch, int dim_x, int dim_y, float dt, float dx, int iters,
                         float omega, int n_steps, fs_ctx *ctx);
/* setup(), ino:196-241, on the device: zero velocity, the three-sector colour wheel (atan2f of the
 * node's offset from the centre against +-PI/3) and the two IN-PLACE 1-2-1 smoothing passes (first
 * along j, then along i; each reads its already-smoothed predecessor).  Every one of the `batch`
 * grids (laid out back to back) receives the same initial condition. */
int fs_init_color_wheel(fs_vec2f *v, fs_rgb_uq32 *c, int batch, int dim_x, int dim_y, fs_ctx *ctx);
/* touch_routine(), ino:63-96, on the device.  samples: [batch][n_samples][3] int32 = {touched, raw x,
 * raw y}, one per polling period; every touched sample whose predecessor was touched becomes a drag
 * record {coords = map() of the raw reading onto [0,n_cols] x [0,n_rows] (ino:77-78), velocity =
 * delta * 1000.f / polling_ms (ino:82-83)}, in order, at most max_drags per grid (the sketch's queue
 * holds 10 and drops the rest, ino:49,85).  cal == NULL: the sketch's calibration (ino:17-21). */
typedef struct fs_touch_cal { int min_x, max_x, min_y, max_y, polling_ms; } fs_touch_cal;
int fs_touch_to_drags(fs_drag *drags_out, int *counts_out, const int *samples, int n_samples, int batch,
                      int max_drags, int n_rows, int n_cols, const fs_touch_cal *cal, fs_ctx *ctx);

/* ---- the same operators over HOST pointers (reference drop-ins) ------------- */
int fsh_advect_vec2f(fs_vec2f *next_p, const fs_vec2f *p, const fs_vec2f *vel,
                     int dim_x, int dim_y, float dt, int no_slip, fs_ctx *ctx);
int fsh_advect_rgb_uq32(fs_rgb_uq32 *next_c, const fs_rgb_uq32 *c, const fs_vec2f *vel,
                        int dim_x, int dim_y, float dt, int no_slip, fs_ctx *ctx);
int fsh_calculate_divergence(float *div, const fs_vec2f *v, int dim_x, int dim_y,
                             float dx, fs_ctx *ctx);
int fsh_subtract_gradient(fs_vec2f *v, const float *p, int dim_x, int dim_y, float dx,
                          fs_ctx *ctx);
int fsh_poisson_solve(float *p, const float *div, int dim_x, int dim_y, float dx,
                      int iters, float omega, fs_ctx *ctx);
int fsh_step(fs_vec2f *v, fs_rgb_uq32 *c, const fs_drag *drags, int n_drags,
             int dim_x, int dim_y, float dt, float dx, int iters, float omega,
             float *p_out, float *div_out, fs_ctx *ctx);
int fsh_upscale4_rgb565(uint16_t *out, const fs_rgb_uq32 *c, int dim_x, int dim_y,
                        fs_ctx *ctx);

/* ---- decomposed grids (2-D block decomposition) -------------- */
/* A rank's padded local window of a global grid.  Local array element (lx,ly)
 * lives at ly*nx + lx and is global node (ox+lx, oy+ly).  Wall rules apply only
 * where the window touches the GLOBAL boundary; red/black parity and advect
 * coordinates are global. */
typedef struct fs_tile {
    int gdim_x, gdim_y;   /* global grid */
    int ox, oy;           /* global coordinate of local (0,0); may be negative only if clipped by caller */
    int nx, ny;           /* local (padded) extents; pitch = nx */
    int x0, y0, x1, y1;   /* compute rectangle [x0,x1) x [y0,y1) in local coordinates */
} fs_tile;

int fs_tile_advect_vec2f(fs_vec2f *next_p, const fs_vec2f *p, const fs_vec2f *vel,
                         const fs_tile *t, float dt, int no_slip, fs_ctx *ctx);
int fs_tile_advect_rgb_uq32(fs_rgb_uq32 *next_c, const fs_rgb_uq32 *c, const fs_vec2f *vel,
                            const fs_tile *t, float dt, int no_slip, fs_ctx *ctx);
int fs_tile_calculate_divergence(float *div, const fs_vec2f *v, const fs_tile *t,
                                 float dx, fs_ctx *ctx);
int fs_tile_subtract_gradient(fs_vec2f *v, const float *p, const fs_tile *t, float dx,
                              fs_ctx *ctx);
/* `n_half` consecutive half-sweeps starting with colour `first_parity`, from
 * p_in to p_out (distinct buffers), valid on the compute rectangle provided
 * p_in is valid on that rectangle grown by n_half nodes (clipped to the global
 * grid).  p_in == NULL means "all zero" (poisson.cpp:117-119). */
int fs_tile_sor_sweeps(float *p_out, const float *p_in, const float *div,
                       const fs_tile *t, float dx, float omega, int first_parity,
                       int n_half, fs_ctx *ctx);
int fs_tile_apply_drags(fs_vec2f *v, const fs_drag *drags, int n, const fs_tile *t,
                        fs_ctx *ctx);
/* Reads (and clears) the context's device-side status flag set by tile advects
 * whose backtrace left the local window: returns FS_OK or FS_ERR_HALO_OVERRUN.
 * Synchronises the stream. */
int fs_tile_check(fs_ctx *ctx);
/* max over the compute rectangle of max(|v.x|,|v.y|)*dt, rounded up to whole
 * nodes: the halo a subsequent advect needs.  Synchronises the stream. */
int fs_tile_max_displacement(int *out_nodes, const fs_vec2f *vel, const fs_tile *t,
                             float dt, fs_ctx *ctx);

/* ---- halo exchange over NVLink peer memory (no NCCL on the critical path) -------- */
/* Field windows that take part in exchanges live in an arena from fs_arena_alloc
 * (plain cudaMalloc, zero-filled; its first bytes are the caller's flag slots).  A
 * rank exports its arena with fs_ipc_export, ships the 64-byte handle to its
 * neighbours by any means (torch.distributed), and they map it with fs_ipc_open. */
int fs_arena_alloc(void **out, size_t bytes, fs_ctx *ctx);
int fs_arena_free(void *arena, fs_ctx *ctx);
int fs_ipc_export(void *arena, unsigned char handle[64], fs_ctx *ctx);
int fs_ipc_open(void **peer_arena, const unsigned char handle[64], fs_ctx *ctx);
int fs_ipc_close(void *peer_arena, fs_ctx *ctx);

#define FS_HALO_MAX_COPIES 32
#define FS_HALO_MAX_PEERS 8
typedef struct fs_halo_copy {       /* one pitched 2-D strip, sizes in bytes (multiples of 4) */
    const void *src;                /* in this rank's window */
    void *dst;                      /* in a neighbour's ghost region (pointer into its mapped arena) */
    long long src_pitch, dst_pitch;
    int row_bytes, rows;
} fs_halo_copy;
/* ONE kernel: store the strips into the neighbours' ghosts, then write `seq` to each
 * signal_flags[k] (a uint64 slot in neighbour k's arena) and wait until every
 * wait_flags[k] (uint64 slots in this rank's arena) holds a value >= seq.  `seq` must
 * grow by one per exchange, identically on all ranks. */
int fs_halo_exchange(const fs_halo_copy *copies, int n_copies, void *const *signal_flags,
                     void *const *wait_flags, int n_peers, unsigned long long seq, fs_ctx *ctx);
/* Re-target the context at another stream (e.g. a capturing stream). */
int fs_ctx_set_stream(fs_ctx *ctx, void *stream);


/* ---- a device-resident simulation: one CUDA-graph launch per step + an asynchronous frame stream -----
 * (SURVEY.md §8f #2).  The sketch's loop() task hands every step's dye to the draw task through the
 * color_consumed / color_produced semaphore pair (ino:285-288) — a double buffer.  fs_sim keeps the
 * state on the device, runs a step as ONE cudaGraphLaunch (the per-step drag records are written into
 * the captured kernel node with cudaGraphExecKernelNodeSetParams; two graphs stand for the dye pointer
 * swap of ino:286), renders the step's RGB565 frame inside the dye advect (frame = 1) and streams it
 * to one of two pinned host buffers on a copy stream while the next step computes.
 *   producer: fs_sim_step        (FS_ERR_WOULD_BLOCK when two frames are waiting for the consumer)
 *   consumer: fs_sim_acquire_frame (waits for the oldest frame's copy; "take color_produced")
 *             fs_sim_release_frame ("give color_consumed")
 * Small or odd-pitched grids that the fused kernels cannot take are stepped without a graph; results
 * are identical either way. */
typedef struct fs_sim fs_sim;
int fs_sim_create(fs_sim **out, int dim_x, int dim_y, float dt, float dx, int iters, float omega, int frame,
                  fs_ctx *ctx);
int fs_sim_destroy(fs_sim *s);
int fs_sim_upload(fs_sim *s, const fs_vec2f *v, const fs_rgb_uq32 *c);                 /* host or device pointers */
int fs_sim_download(fs_sim *s, fs_vec2f *v, fs_rgb_uq32 *c, float *p, float *div);     /* any may be NULL; synchronises */
int fs_sim_step(fs_sim *s, const fs_drag *drags, int n_drags);                         /* `drags`: HOST array */
int fs_sim_acquire_frame(fs_sim *s, const uint16_t **frame, int *rows, int *cols);     /* pinned host memory */
int fs_sim_release_frame(fs_sim *s);
int fs_sim_stats(const fs_sim *s, unsigned long long *steps, unsigned long long *graph_launches,
                 unsigned long long *eager_steps);

/* ---- the decomposed step as ONE object per rank (fs_dist_*) -----------------------
 * fs_dist owns a rank's windows of every field (one IPC-exportable arena), knows its neighbours and
 * runs the whole loop() body (ino:249-289) of its rectangle per call:
 *   fused advect+drags+divergence on the rectangle grown by the SOR passes' redundancy ring  ->
 *   blocked SOR passes, each FUSED with its halo exchange (rim tiles store straight into the
 *   neighbours' ghosts over NVLink, interior tiles overlap the hand-shake)  ->  gradient-subtract  ->
 *   one exchange kernel for projected velocity + dye  ->  dye advect.
 * Results are bit-identical to fs_step on the whole grid.  Every rank must make the same calls.
 * Bootstrap: create on every rank, exchange the 64-byte fs_dist_ipc_handle()s by any means
 * (torch.distributed, MPI, a file), fs_dist_connect() with all `world` handles in rank order. */
typedef struct fs_dist fs_dist;
typedef struct fs_dist_config {
    int gdim_x, gdim_y;     /* global grid */
    int world, rank;
    int px, py;             /* ranks along dim_x / dim_y; 0 = choose (1x1, 1x2, 2x2, 2x4) */
    int ghost;              /* ghost nodes towards each neighbouring rank, multiple of 4 (64) */
    int advect_halo;        /* A: how far (nodes) a backtrace may leave the rectangle; beyond it the
                               step raises FS_ERR_HALO_OVERRUN.  Needs A + 2*sor_t + 2 <= ghost */
    int iters;              /* SOR iterations per step (the context's "sor_t" of them per pass) */
    float dt, dx, omega;
    int frame;              /* 1 = also render this rank's part of the 4x RGB565 frame every step (fused into
                               the dye advect, ino:116-177); read it with fs_dist_frame */
} fs_dist_config;
typedef struct fs_dist_info_t {
    int px, py, n_neighbours;
    int sor_passes, sor_t;
    int div_ring;           /* the divergence is formed on the rectangle grown by this many nodes */
    int velocity_halo, dye_halo;   /* ghost widths refreshed every step */
    int exchanges_per_step;
    unsigned long long exchanges;  /* hand-shakes since creation */
    size_t arena_bytes;
    float phase_ms[5];      /* the LAST step on this rank, CUDA events: advect+drags+divergence | SOR passes incl.
                               their fused exchanges | gradient | wait for the dye halo (the velocity and dye exchanges run
                               on side streams under the dye advect / the next step's advect + SOR) | dye advect (+ frame);
                               -1 before the first step.  Querying synchronises with that step. */
} fs_dist_info_t;
/* (A rank keeps three streams busy with kernels that wait for other ranks' flags: run the process with
 * CUDA_DEVICE_MAX_CONNECTIONS=32 so that each gets a hardware work queue of its own.) */
int fs_dist_create(fs_dist **out, const fs_dist_config *cfg, fs_ctx *ctx);
int fs_dist_destroy(fs_dist *d);
int fs_dist_window(const fs_dist *d, fs_tile *out);          /* this rank's window (pitch = nx) */
int fs_dist_info(const fs_dist *d, fs_dist_info_t *out);
int fs_dist_ipc_handle(fs_dist *d, unsigned char handle[64]);
/* `handles`: world x 64 bytes, rank order; maps the (up to 8) neighbours' arenas */
int fs_dist_connect(fs_dist *d, const unsigned char *handles);
/* all ranks live in THIS process on one device (tests, single-GPU emulation): plain pointers */
int fs_dist_connect_local(fs_dist *d, fs_dist *const *all_ranks);
/* whole windows (nx*ny elements, ghosts included; their contents are refreshed before use), host or
 * device pointers; asynchronous on the context's stream */
int fs_dist_upload(fs_dist *d, const fs_vec2f *v_window, const fs_rgb_uq32 *c_window);
/* the OWNED rectangle of the current state as dense (x1-x0) x (y1-y0) arrays (host or device
 * pointers; any may be NULL); p/div = the last step's pressure and divergence.  Synchronises. */
int fs_dist_download(fs_dist *d, fs_vec2f *v_rect, fs_rgb_uq32 *c_rect, float *p_rect, float *div_rect);
/* device pointers of the CURRENT windows (they alternate from step to step like ino:255,286) */
int fs_dist_device_fields(fs_dist *d, fs_vec2f **v, fs_rgb_uq32 **c, float **p, float **div);
/* this rank's part of the frame: the cells that START in its rectangle, 4*cells_x rows x 4*cells_y columns
 * of RGB565 (device pointer, dense, row pitch 4*cells_y); global pixel (4*gx0 + r, 4*gy0 + c) */
int fs_dist_frame(fs_dist *d, uint16_t **frame, int *rows, int *cols);
/* one loop() body; `drags` = the step's whole queue (HOST array; every rank passes all records) */
int fs_dist_step(fs_dist *d, const fs_drag *drags, int n_drags);
/* FS_OK, FS_ERR_HALO_OVERRUN or FS_ERR_HALO_TIMEOUT since the last check; synchronises */
int fs_dist_check(fs_dist *d);

#ifdef __cplusplus
}
#endif
#endif /* FLUID_B200_H */
